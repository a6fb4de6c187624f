// Physics-loss operators: transport / momentum residuals of a (sigma, u) field by central
// differences, weighted mean-square loss, and the loss VJP with respect to the residuals.
// API-compatible with the reference's include/phys.h (same names, argument order, defaults, and
// the host-pointer contract: every pointer is caller-owned HOST memory, the callee handles all
// device work).  All cuda_* functions are implemented on the sm_100a kernels of this repository;
// the cpu_* functions are declared for source compatibility and defined only by the reference.
//
//   R_sigma = d_t sigma + u . grad(sigma) + sigma * div(u)
//   R_u     = d_t u + (u . grad) u
//   d_t f ~ (f(t+dt) - f(t-dt)) / (2 dt),  d_x f ~ (f[x+1] - f[x-1]) / (2 hx)  (wrap if periodic, else clamp)
//   L_sigma = w_sigma * mean(R_sigma^2),  L_u = w_u * mean(|R_u|^2),  g = 2 w / N * R
//
// Layout: scalar fields are N = nx*ny*nz floats with linear index (z*ny + y)*nx + x; vector fields
// are 3N floats, channel-major [ux | uy | uz].
#ifndef PHYS_AUTODIFF_PHYS_H
#define PHYS_AUTODIFF_PHYS_H

#include <cstddef>

namespace phys {

struct GridSpec {
    int nx{0}, ny{0}, nz{0};
    float hx{1.f}, hy{1.f}, hz{1.f};
    float dt{1.f};
    bool periodic{true};
};

struct PhysWeights {
    float w_sigma{1.f};
    float w_u{1.f};
};

// ---- CUDA, "non-fused" family (reference include/phys.h:67-117).  Same results as the fused
// family; kept as distinct symbols because callers select the backend by function name.
void cuda_phys_residuals_nonfused(const GridSpec& g, const float* sigma_tm1, const float* sigma_t, const float* sigma_tp1,
                                  const float* u_tm1, const float* u_t, const float* u_tp1,
                                  float* R_sigma, float* R_ux, float* R_uy, float* R_uz);

// Loss with the reduction on the device (the reference reduces on the host); out_* and opt_* may be null.
void cuda_phys_loss_forward_nonfused(const GridSpec& g, const PhysWeights& w, const float* sigma_tm1, const float* sigma_t,
                                     const float* sigma_tp1, const float* u_tm1, const float* u_t, const float* u_tp1,
                                     float* out_loss_sigma, float* out_loss_u, float* opt_R_sigma = nullptr,
                                     float* opt_R_ux = nullptr, float* opt_R_uy = nullptr, float* opt_R_uz = nullptr);

void cuda_phys_loss_backward_nonfused(const GridSpec& g, const PhysWeights& w, const float* R_sigma, const float* R_ux,
                                      const float* R_uy, const float* R_uz,
                                      float* g_sigma, float* g_ux, float* g_uy, float* g_uz);

// *_timed: kernel_ms receives the kernel-only time from CUDA events (may be null).
void cuda_phys_residuals_nonfused_timed(const GridSpec& g, const float* sigma_tm1, const float* sigma_t,
                                        const float* sigma_tp1, const float* u_tm1, const float* u_t, const float* u_tp1,
                                        float* R_sigma, float* R_ux, float* R_uy, float* R_uz, float* kernel_ms);

// ---- CUDA, "fused" family (reference include/phys.h:120-156): one stencil kernel.
void cuda_phys_residuals_fused(const GridSpec& g, const float* sigma_tm1, const float* sigma_t, const float* sigma_tp1,
                               const float* u_tm1, const float* u_t, const float* u_tp1,
                               float* R_sigma, float* R_ux, float* R_uy, float* R_uz);

// Takes the FIELDS (not residuals) and recomputes the residual inside the VJP kernel.
void cuda_phys_loss_backward_fused(const GridSpec& g, const PhysWeights& w, const float* sigma_tm1, const float* sigma_t,
                                   const float* sigma_tp1, const float* u_tm1, const float* u_t, const float* u_tp1,
                                   float* g_sigma, float* g_ux, float* g_uy, float* g_uz);

void cuda_phys_residuals_fused_timed(const GridSpec& g, const float* sigma_tm1, const float* sigma_t,
                                     const float* sigma_tp1, const float* u_tm1, const float* u_t, const float* u_tp1,
                                     float* R_sigma, float* R_ux, float* R_uy, float* R_uz, float* kernel_ms);

// ---- CPU reference family (reference include/phys.h:26-64): declarations only.
void cpu_phys_residuals(const GridSpec& g, const float* sigma_tm1, const float* sigma_t, const float* sigma_tp1,
                        const float* u_tm1, const float* u_t, const float* u_tp1,
                        float* R_sigma, float* R_ux, float* R_uy, float* R_uz);
void cpu_phys_loss_forward(const GridSpec& g, const PhysWeights& w, const float* sigma_tm1, const float* sigma_t,
                           const float* sigma_tp1, const float* u_tm1, const float* u_t, const float* u_tp1,
                           float* out_loss_sigma, float* out_loss_u, float* opt_R_sigma = nullptr,
                           float* opt_R_ux = nullptr, float* opt_R_uy = nullptr, float* opt_R_uz = nullptr);
void cpu_phys_loss_backward(const GridSpec& g, const PhysWeights& w, const float* R_sigma, const float* R_ux,
                            const float* R_uy, const float* R_uz, float* g_sigma, float* g_ux, float* g_uy, float* g_uz);

}  // namespace phys

#endif  // PHYS_AUTODIFF_PHYS_H
