// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI shim around the UNMODIFIED reference CPU sources, which are compiled where they lie
// (/root/reference/src/{mlp_cpu,mlp_grid,phys_cpu}.cpp) by oracle/Makefile into
// oracle/_ref/libphysref.so.  Nothing here restates reference arithmetic: every function below
// forwards to the reference's own symbol.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
//
// The one non-forwarding piece is ref_fused_loss_mt: a std::thread driver that calls the
// reference's own mlp_infer_cpu on disjoint point ranges (the MLP is pointwise, so the bits are
// identical to one big call -- SURVEY.md section 8d) so that 128^3/256^3 grids finish in
// reasonable time and without mlp_forward<ExecCpu>'s 2*B*H scratch (src/mlp_cpu.cpp:16).
#include "backend.h"
#include "mlp.h"
#include "mlp_grid.h"
#include "phys.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

// src/mlp_grid.cpp:50 references mlp_forward<ExecCuda>; a CPU-only oracle must still link.
// (mlp_backward<ExecCuda> is referenced by nothing in the three CPU sources.)
template <>
void mlp_forward<ExecCuda>(const float*, const float*, const float*, const float*, const float*, float*,
                           std::size_t, std::size_t, std::size_t, std::size_t) {
    std::fprintf(stderr, "oracle/_ref: mlp_forward<ExecCuda> is a link stub in the CPU oracle\n");
    std::abort();
}

namespace {

struct CGrid {  // mirrors phys::GridSpec (include/phys.h:8-13) with a C-friendly bool
    int nx, ny, nz;
    float hx, hy, hz, dt;
    int periodic;
};

phys::GridSpec to_spec(const CGrid* g) {
    phys::GridSpec s;
    s.nx = g->nx; s.ny = g->ny; s.nz = g->nz;
    s.hx = g->hx; s.hy = g->hy; s.hz = g->hz;
    s.dt = g->dt; s.periodic = g->periodic != 0;
    return s;
}

phys::MLPWeights to_weights(int In, int H, int Out, const float* W1, const float* b1, const float* W2, const float* b2) {
    phys::MLPWeights w;
    w.W1.assign(W1, W1 + std::size_t(H) * In);
    w.b1.assign(b1, b1 + H);
    w.W2.assign(W2, W2 + std::size_t(Out) * H);
    w.b2.assign(b2, b2 + Out);
    return w;
}

phys::MLPGridConfig to_cfg(int In, int H, int Out, int norm_minus_one_to_one) {
    phys::MLPGridConfig c;
    c.dims.In = In; c.dims.H = H; c.dims.Out = Out;
    c.norm = norm_minus_one_to_one ? phys::CoordNorm::MinusOneToOne : phys::CoordNorm::ZeroToOne;
    return c;
}

}  // namespace

extern "C" {

void ref_mlp_random_init(int In, int H, int Out, unsigned seed, float scale, float* W1, float* b1, float* W2, float* b2) {
    phys::MLPWeights w;
    phys::MLPDims d; d.In = In; d.H = H; d.Out = Out;
    phys::mlp_random_init(w, d, seed, scale);
    std::memcpy(W1, w.W1.data(), w.W1.size() * sizeof(float));
    std::memcpy(b1, w.b1.data(), w.b1.size() * sizeof(float));
    std::memcpy(W2, w.W2.data(), w.W2.size() * sizeof(float));
    std::memcpy(b2, w.b2.data(), w.b2.size() * sizeof(float));
}

void ref_make_grid_coords(const CGrid* g, float t, int norm_m1p1, float* coords /* N*4 */) {
    std::vector<float> c;
    phys::make_grid_coords(to_spec(g), t, norm_m1p1 ? phys::CoordNorm::MinusOneToOne : phys::CoordNorm::ZeroToOne, c);
    std::memcpy(coords, c.data(), c.size() * sizeof(float));
}

void ref_mlp_forward_cpu(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* y,
                         size_t B, size_t In, size_t H, size_t Out) {
    mlp_forward<ExecCpu>(x, W1, b1, W2, b2, y, B, In, H, Out);
}

void ref_mlp_backward_cpu(const float* x, const float* y_target, const float* W1, const float* b1, const float* W2,
                          const float* b2, float* dW1, float* db1, float* dW2, float* db2, size_t B, size_t In, size_t H,
                          size_t Out) {
    mlp_backward<ExecCpu>(x, y_target, W1, b1, W2, b2, dW1, db1, dW2, db2, B, In, H, Out);
}

void ref_mlp_grid_infer_cpu(const CGrid* g, int In, int H, int Out, int norm_m1p1, const float* W1, const float* b1,
                            const float* W2, const float* b2, float t, float* out /* N*Out */) {
    std::vector<float> o;
    phys::mlp_grid_infer_cpu(to_spec(g), to_cfg(In, H, Out, norm_m1p1), to_weights(In, H, Out, W1, b1, W2, b2), t, o);
    std::memcpy(out, o.data(), o.size() * sizeof(float));
}

void ref_mlp_generate_fields_cpu(const CGrid* g, int In, int H, int Out, int norm_m1p1, const float* W1, const float* b1,
                                 const float* W2, const float* b2, float t, float dt, float* s_m, float* s_0, float* s_p,
                                 float* u_m, float* u_0, float* u_p) {
    std::vector<float> a, b, c, d, e, f;
    phys::mlp_generate_fields_cpu(to_spec(g), to_cfg(In, H, Out, norm_m1p1), to_weights(In, H, Out, W1, b1, W2, b2), t, dt,
                                  a, b, c, d, e, f);
    std::memcpy(s_m, a.data(), a.size() * sizeof(float));
    std::memcpy(s_0, b.data(), b.size() * sizeof(float));
    std::memcpy(s_p, c.data(), c.size() * sizeof(float));
    std::memcpy(u_m, d.data(), d.size() * sizeof(float));
    std::memcpy(u_0, e.data(), e.size() * sizeof(float));
    std::memcpy(u_p, f.data(), f.size() * sizeof(float));
}

void ref_phys_residuals(const CGrid* g, const float* s_m, const float* s_0, const float* s_p, const float* u_m,
                        const float* u_0, const float* u_p, float* Rs, float* Rx, float* Ry, float* Rz) {
    phys::cpu_phys_residuals(to_spec(g), s_m, s_0, s_p, u_m, u_0, u_p, Rs, Rx, Ry, Rz);
}

void ref_phys_loss_forward(const CGrid* g, float w_sigma, float w_u, const float* s_m, const float* s_0, const float* s_p,
                           const float* u_m, const float* u_0, const float* u_p, float* loss_sigma, float* loss_u,
                           float* Rs, float* Rx, float* Ry, float* Rz) {
    phys::PhysWeights w; w.w_sigma = w_sigma; w.w_u = w_u;
    phys::cpu_phys_loss_forward(to_spec(g), w, s_m, s_0, s_p, u_m, u_0, u_p, loss_sigma, loss_u, Rs, Rx, Ry, Rz);
}

void ref_phys_loss_backward(const CGrid* g, float w_sigma, float w_u, const float* Rs, const float* Rx, const float* Ry,
                            const float* Rz, float* gs, float* gx, float* gy, float* gz) {
    phys::PhysWeights w; w.w_sigma = w_sigma; w.w_u = w_u;
    phys::cpu_phys_loss_backward(to_spec(g), w, Rs, Rx, Ry, Rz, gs, gx, gy, gz);
}

// Whole hot path on the CPU with `threads` host threads: reference make_grid_coords + mlp_infer_cpu on
// chunks (bit-identical to mlp_generate_fields_cpu, see header), the reference's channel-major split
// layout (src/mlp_grid.cpp:69-80), then the reference's single-threaded cpu_phys_loss_forward.
// Residual outputs are optional (may be null).  Returns 0, or -1 on bad arguments.
int ref_fused_loss_mt(const CGrid* g, int In, int H, int Out, int norm_m1p1, const float* W1, const float* b1,
                      const float* W2, const float* b2, float t, float dt, float w_sigma, float w_u, int threads,
                      float* loss_sigma, float* loss_u, float* Rs, float* Rx, float* Ry, float* Rz) {
    if (In != 4 || Out != 4 || threads < 1) return -1;
    const phys::GridSpec spec = to_spec(g);
    const phys::MLPGridConfig cfg = to_cfg(In, H, Out, norm_m1p1);
    const phys::MLPWeights w = to_weights(In, H, Out, W1, b1, W2, b2);
    const std::size_t N = std::size_t(spec.nx) * spec.ny * spec.nz;
    const float ts[3] = {t - dt, t, t + dt};   // same float expressions as src/mlp_grid.cpp:87-89
    std::vector<float> sig[3], u[3];
    for (int s = 0; s < 3; ++s) { sig[s].resize(N); u[s].resize(3 * N); }

    // The coordinate array of a time slice is generated once with the reference's own make_grid_coords (the z
    // normalisation uses the global nz) and the workers take disjoint chunks of points from it.
    for (int s = 0; s < 3; ++s) {
        std::vector<float> coords;
        phys::make_grid_coords(spec, ts[s], cfg.norm, coords);
        const std::size_t chunk = 1 << 11;  // 2*chunk*H floats of reference scratch stay cache-resident
        const std::size_t nchunks = (N + chunk - 1) / chunk;
        auto work = [&](int tid) {
            std::vector<float> y(chunk * 4);
            for (std::size_t c = tid; c < nchunks; c += threads) {
                const std::size_t i0 = c * chunk, n = std::min(chunk, N - i0);
                phys::mlp_infer_cpu(cfg.dims, w, coords.data() + i0 * 4, n, y.data());
                for (std::size_t i = 0; i < n; ++i) {
                    sig[s][i0 + i]       = y[i * 4 + 0];
                    u[s][i0 + i]         = y[i * 4 + 1];
                    u[s][N + i0 + i]     = y[i * 4 + 2];
                    u[s][2 * N + i0 + i] = y[i * 4 + 3];
                }
            }
        };
        std::vector<std::thread> pool;
        for (int k = 1; k < threads; ++k) pool.emplace_back(work, k);
        work(0);
        for (auto& th : pool) th.join();
    }
    phys::PhysWeights pw; pw.w_sigma = w_sigma; pw.w_u = w_u;
    phys::cpu_phys_loss_forward(spec, pw, sig[0].data(), sig[1].data(), sig[2].data(), u[0].data(), u[1].data(),
                                u[2].data(), loss_sigma, loss_u, Rs, Rx, Ry, Rz);
    return 0;
}

int ref_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }

}  // extern "C"
