// The closed loop of the reference's plan (REQUIREMENT.md:155-169, milestone M6) written against the C++ API a
// user of the reference would call: include/mlp_grid.h + phys.h types, phys::mlp_random_init, and the additive
// phys::mlp_phys_loss_grad_cuda (include/phys_b200.h).  Adam on the host (580 parameters), K steps; the plan's
// acceptance criterion is "L falls by >= 90 % within K steps".  Links against libphysad_b200.so only.
#include <cmath>
#include <cstdio>
#include <vector>

#include "phys_b200.h"

int main() {
    using namespace phys;
    GridSpec g{48, 48, 32, 1.f, 1.f, 1.f, 2e-3f, true};
    MLPGridConfig cfg;  // {4, 64, 4}, MinusOneToOne
    MLPWeights w, grad;
    mlp_random_init(w, cfg.dims, 777, 0.25f);
    const PhysWeights pw{1.f, 1.f};
    std::vector<float>* par[4] = {&w.W1, &w.b1, &w.W2, &w.b2};
    std::vector<float>* gr[4] = {&grad.W1, &grad.b1, &grad.W2, &grad.b2};
    std::vector<std::vector<double>> m(4), v(4);
    for (int k = 0; k < 4; ++k) { m[k].assign(par[k]->size(), 0.0); v[k].assign(par[k]->size(), 0.0); }
    const int K = 120;
    const double lr = 3e-3, b1 = 0.9, b2 = 0.999;
    double first = 0.0, last = 0.0;
    for (int it = 1; it <= K; ++it) {
        float ls = 0.f, lu = 0.f;
        mlp_phys_loss_grad_cuda(g, cfg, w, pw, 0.25f, g.dt, &ls, &lu, grad);
        last = double(ls) + double(lu);
        if (it == 1) first = last;
        if (it == 1 || it % 20 == 0) std::printf("step %3d  L_sigma %.6e  L_u %.6e\n", it, ls, lu);
        for (int k = 0; k < 4; ++k)
            for (size_t i = 0; i < par[k]->size(); ++i) {
                const double gg = (*gr[k])[i];
                m[k][i] = b1 * m[k][i] + (1 - b1) * gg;
                v[k][i] = b2 * v[k][i] + (1 - b2) * gg * gg;
                const double mh = m[k][i] / (1 - std::pow(b1, it)), vh = v[k][i] / (1 - std::pow(b2, it));
                (*par[k])[i] = float(double((*par[k])[i]) - lr * mh / (std::sqrt(vh) + 1e-8));
            }
    }
    const bool ok = std::isfinite(last) && last <= 0.1 * first;
    std::printf("%s closed loop: L %.6e -> %.6e (%.1f %% lower) in %d steps\n", ok ? "[PASS]" : "[FAIL]", first, last,
                100.0 * (1.0 - last / first), K);
    return ok ? 0 : 1;
}
