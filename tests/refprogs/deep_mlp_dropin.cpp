// The additive deep-network entry point through the C++ layer (ours, in the style of the reference's test/*.cpp):
// phys::mlp_phys_loss_deep_cuda must (1) with ONE hidden layer reproduce the reference's CPU path
// (mlp_generate_fields_cpu + cpu_phys_loss_forward: src/mlp_grid.cpp:82-106, src/phys_cpu.cpp:112-149) to 1e-6, and
// (2) with three hidden layers agree with the reference's layer rule (src/mlp_cpu.cpp:19-24) applied again on the CPU
// -- strict mode to 1e-6, tensor-core mode to 1e-4 (the north-star's loss tolerance).  Built against this repository's
// include/*.h and linked with the reference's CPU objects.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "phys_b200.h"

// 4 -> H -> ... -> H -> 4 on explicit coordinates, every layer as src/mlp_cpu.cpp: from the bias, + W[g,h] * a[h], h ascending
static void deep_forward_cpu(const phys::DeepMLPWeights& w, std::size_t H, const std::vector<float>& coords, std::vector<float>& y) {
    const std::size_t B = coords.size() / 4;
    y.resize(B * 4);
    std::vector<float> a(H), b(H);
    for (std::size_t i = 0; i < B; ++i) {
        for (std::size_t h = 0; h < H; ++h) {
            float s = w.b1[h];
            for (std::size_t k = 0; k < 4; ++k) s += w.W1[h * 4 + k] * coords[i * 4 + k];
            a[h] = s > 0.f ? s : 0.f;
        }
        for (int l = 0; l + 1 < w.hidden_layers; ++l) {
            for (std::size_t g = 0; g < H; ++g) {
                float s = w.bh[l * H + g];
                for (std::size_t h = 0; h < H; ++h) s += w.Wh[(l * H + g) * H + h] * a[h];
                b[g] = s > 0.f ? s : 0.f;
            }
            a.swap(b);
        }
        for (std::size_t o = 0; o < 4; ++o) {
            float s = w.b2[o];
            for (std::size_t h = 0; h < H; ++h) s += w.W2[o * H + h] * a[h];
            y[i * 4 + o] = s;
        }
    }
}

int main() {
    phys::GridSpec g;
    g.nx = 37; g.ny = 20; g.nz = 9; g.hx = g.hy = g.hz = 1.0f; g.dt = 2e-3f; g.periodic = true;
    const std::size_t N = std::size_t(g.nx) * g.ny * g.nz, H = 64;
    phys::MLPGridConfig cfg;
    cfg.dims.H = H;
    phys::MLPWeights w1;
    phys::mlp_random_init(w1, cfg.dims, 321u, 0.25f);
    phys::PhysWeights pw;
    const float t = 0.25f, dt = 2e-3f;
    int fails = 0;
    auto close = [](float a, float b, float tol) { return std::fabs(a - b) <= tol * std::fabs(a); };

    {   // (1) one hidden layer == the reference
        std::vector<float> c[6];
        phys::mlp_generate_fields_cpu(g, cfg, w1, t, dt, c[0], c[1], c[2], c[3], c[4], c[5]);
        float ls_c = 0, lu_c = 0, ls_d = 0, lu_d = 0;
        phys::cpu_phys_loss_forward(g, pw, c[0].data(), c[1].data(), c[2].data(), c[3].data(), c[4].data(), c[5].data(), &ls_c, &lu_c);
        phys::DeepMLPWeights d;
        d.hidden_layers = 1; d.W1 = w1.W1; d.b1 = w1.b1; d.W2 = w1.W2; d.b2 = w1.b2;
        phys::mlp_phys_loss_deep_cuda(g, cfg, d, pw, t, dt, &ls_d, &lu_d);
        const bool ok = close(ls_c, ls_d, 1e-6f) && close(lu_c, lu_d, 1e-6f);
        std::printf("L=1 reference cpu (%.9g, %.9g) vs mlp_phys_loss_deep_cuda (%.9g, %.9g): %s\n", ls_c, lu_c, ls_d, lu_d, ok ? "[PASS]" : "[FAIL]");
        fails += !ok;
    }
    {   // (2) three hidden layers: the layer rule again on the CPU, then the reference's loss
        phys::DeepMLPWeights d;
        d.hidden_layers = 3; d.W1 = w1.W1; d.b1 = w1.b1; d.W2 = w1.W2; d.b2 = w1.b2;
        std::mt19937 gen(5);
        std::uniform_real_distribution<float> dist(-0.2f, 0.2f);
        d.Wh.resize(2 * H * H); d.bh.resize(2 * H);
        for (float& v : d.Wh) v = dist(gen);
        for (float& v : d.bh) v = dist(gen);
        std::vector<float> sig[3], vel[3];
        const float ts[3] = {t - dt, t, t + dt};
        for (int s = 0; s < 3; ++s) {
            std::vector<float> coords, y;
            phys::make_grid_coords(g, ts[s], cfg.norm, coords);
            deep_forward_cpu(d, H, coords, y);
            sig[s].resize(N); vel[s].resize(3 * N);
            for (std::size_t i = 0; i < N; ++i) {
                sig[s][i] = y[i * 4];
                for (int k = 0; k < 3; ++k) vel[s][k * N + i] = y[i * 4 + 1 + k];
            }
        }
        float ls_c = 0, lu_c = 0;
        phys::cpu_phys_loss_forward(g, pw, sig[0].data(), sig[1].data(), sig[2].data(), vel[0].data(), vel[1].data(), vel[2].data(), &ls_c, &lu_c);
        for (int fast = 0; fast < 2; ++fast) {
            float ls_d = 0, lu_d = 0;
            phys::mlp_phys_loss_deep_cuda(g, cfg, d, pw, t, dt, &ls_d, &lu_d, fast != 0);
            const float tol = fast ? 1e-4f : 1e-6f;
            const bool ok = close(ls_c, ls_d, tol) && close(lu_c, lu_d, tol);
            std::printf("L=3 %s cpu (%.9g, %.9g) vs cuda (%.9g, %.9g) tol %g: %s\n", fast ? "tensor-core" : "strict", ls_c, lu_c, ls_d, lu_d,
                        double(tol), ok ? "[PASS]" : "[FAIL]");
            fails += !ok;
        }
    }
    return fails ? 1 : 0;
}
