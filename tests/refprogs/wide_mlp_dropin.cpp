// Drop-in behaviour for shapes off the tuned path (ours, in the style of the reference's test/*.cpp): a hidden
// width the grid kernels are not instantiated for (H = 192 > 128) must still work through the reference's names --
// phys::mlp_generate_fields_cuda, phys::mlp_grid_infer_cuda and the additive phys::mlp_phys_loss_fused_cuda -- and
// agree with the reference's CPU path (src/mlp_grid.cpp:82-106, src/phys_cpu.cpp:112-149): fields bit for bit
// (strict fp32 operator), losses to 1e-6.  Built against this repository's include/*.h.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "phys_b200.h"

static bool same_bits(const std::vector<float>& a, const std::vector<float>& b) {
    return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(float)) == 0;
}

int main() {
    phys::GridSpec g;
    g.nx = 37; g.ny = 20; g.nz = 9; g.hx = g.hy = g.hz = 1.0f; g.dt = 2e-3f; g.periodic = false;
    int fails = 0;
    for (std::size_t H : {std::size_t(192), std::size_t(130)}) {
        phys::MLPGridConfig cfg;
        cfg.dims.H = H;
        phys::MLPWeights w;
        phys::mlp_random_init(w, cfg.dims, 321u, 0.25f);
        std::vector<float> c[6], d[6];
        phys::mlp_generate_fields_cpu(g, cfg, w, 0.25f, 2e-3f, c[0], c[1], c[2], c[3], c[4], c[5]);
        phys::mlp_generate_fields_cuda(g, cfg, w, 0.25f, 2e-3f, d[0], d[1], d[2], d[3], d[4], d[5]);
        bool ok = true;
        for (int k = 0; k < 6; ++k) ok = ok && same_bits(c[k], d[k]);
        std::vector<float> yc, yd;
        phys::mlp_grid_infer_cpu(g, cfg, w, 0.3f, yc);
        phys::mlp_grid_infer_cuda(g, cfg, w, 0.3f, yd);
        ok = ok && same_bits(yc, yd);
        std::printf("H=%zu fields/grid_infer cpu == cuda bitwise: %s\n", H, ok ? "[PASS]" : "[FAIL]");
        fails += !ok;
        phys::PhysWeights pw;
        float ls_c = 0, lu_c = 0, ls_d = 0, lu_d = 0;
        phys::cpu_phys_loss_forward(g, pw, c[0].data(), c[1].data(), c[2].data(), c[3].data(), c[4].data(), c[5].data(), &ls_c, &lu_c);
        phys::mlp_phys_loss_fused_cuda(g, cfg, w, pw, 0.25f, 2e-3f, &ls_d, &lu_d);
        const bool lok = std::fabs(ls_c - ls_d) <= 1e-6f * std::fabs(ls_c) && std::fabs(lu_c - lu_d) <= 1e-6f * std::fabs(lu_c);
        std::printf("H=%zu loss cpu (%.9g, %.9g) vs mlp_phys_loss_fused_cuda (%.9g, %.9g): %s\n", H, ls_c, lu_c, ls_d, lu_d,
                    lok ? "[PASS]" : "[FAIL]");
        fails += !lok;
    }
    return fails ? 1 : 0;
}
