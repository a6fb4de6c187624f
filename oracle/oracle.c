/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle ("port") for the phys-autodiff hot path.
 *
 * Plain-C restatement of the reference's CPU algorithm for: weight init, grid coordinates, the
 * two-layer ReLU MLP, the channel-major field split, the central-difference PDE residuals, the
 * weighted mean-square loss and its residual-VJP.  Every function cites the reference lines it
 * follows (paths relative to /root/reference).  It is pinned against the reference itself:
 * tests/test_oracle_cpu.py compares it bit-for-bit with oracle/_ref/libphysref.so (the unmodified
 * reference sources, built by oracle/Makefile) where that library is present, and against the
 * golden vectors in tests/golden/ (generated from the reference by tests/golden/make_golden.py)
 * everywhere else.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library.  The product path (phys_autodiff_b200/csrc) never links or calls it.
 *
 * Build: gcc -std=c11 -O3 -ffp-contract=off (no -march=native, no -ffast-math), so that every
 * `s += w * x` is a separately rounded multiply and add, as in the reference's x86-64 build.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int nx, ny, nz;
    float hx, hy, hz, dt;
    int periodic;
} oracle_grid; /* mirrors phys::GridSpec, include/phys.h:8-13 (bool widened to int) */

/* ------------------------------------------------------------------------------------------
 * Weight init.  src/mlp_grid.cpp:8-19 draws W1, b1, W2, b2 (in that order) from
 * std::uniform_real_distribution<float>(-scale, scale) over std::mt19937(seed).  The stream is
 * standard-library specific; this restates libstdc++ 13's: generate_canonical<float,24> takes ONE
 * 32-bit draw, divides by 2^32 in float, clamps a result of 1.0f to nextafterf(1,0), and the
 * distribution returns canon * (b - a) + a.  Pinned against the reference function for many
 * seeds in tests/test_oracle_cpu.py (anchor: seed 777, scale .25, H=64 -> W1[0] = -0.173668131).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt19937_state;

static void mt_seed(mt19937_state* s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}

static uint32_t mt_next(mt19937_state* s) {
    if (s->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
            uint32_t v = s->mt[(i + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            s->mt[i] = v;
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static float uniform_float(mt19937_state* s, float a, float b) {
    float canon = (float)mt_next(s) / 4294967296.0f;
    if (canon >= 1.0f) canon = nextafterf(1.0f, 0.0f);
    return canon * (b - a) + a;
}

void oracle_mlp_random_init(int In, int H, int Out, uint32_t seed, float scale, float* W1, float* b1, float* W2, float* b2) {
    mt19937_state st;
    mt_seed(&st, seed);
    for (int i = 0; i < H * In; ++i) W1[i] = uniform_float(&st, -scale, scale);
    for (int i = 0; i < H; ++i) b1[i] = uniform_float(&st, -scale, scale);
    for (int i = 0; i < Out * H; ++i) W2[i] = uniform_float(&st, -scale, scale);
    for (int i = 0; i < Out; ++i) b2[i] = uniform_float(&st, -scale, scale);
}

/* ------------------------------------------------------------------------------------------
 * Grid coordinates, src/mlp_grid.cpp:21-43.  Axis value for index i of n: 0 when n<=1, else
 * u = float(i)/float(n-1); MinusOneToOne maps u -> 2u-1 and leaves t; ZeroToOne keeps u and
 * uses t+0.5.  Points are ordered z-outer, y, x-inner; each point is [x,y,z,t].
 * ---------------------------------------------------------------------------------------- */
static float axis_coord(int i, int n, int minus_one_to_one) {
    if (n <= 1) return 0.0f;
    float u = (float)i / (float)(n - 1);
    return minus_one_to_one ? 2.f * u - 1.f : u;
}

/* coords for the linear point range [p0, p0+count) of the grid */
static void coords_range(const oracle_grid* g, float t, int m1p1, size_t p0, size_t count, float* c) {
    const float tt = m1p1 ? t : (t + 0.5f);
    for (size_t q = 0; q < count; ++q) {
        size_t p = p0 + q;
        int x = (int)(p % (size_t)g->nx);
        int y = (int)((p / (size_t)g->nx) % (size_t)g->ny);
        int z = (int)(p / ((size_t)g->nx * (size_t)g->ny));
        c[q * 4 + 0] = axis_coord(x, g->nx, m1p1);
        c[q * 4 + 1] = axis_coord(y, g->ny, m1p1);
        c[q * 4 + 2] = axis_coord(z, g->nz, m1p1);
        c[q * 4 + 3] = tt;
    }
}

void oracle_make_grid_coords(const oracle_grid* g, float t, int m1p1, float* coords) {
    coords_range(g, t, m1p1, 0, (size_t)g->nx * g->ny * g->nz, coords);
}

/* ------------------------------------------------------------------------------------------
 * MLP forward, src/mlp_cpu.cpp:14-36.  Hidden unit: start from b1[h], add W1[h,k]*x[k] for k
 * ascending, ReLU is (s > 0 ? s : 0) (:7-9).  Output: start from b2[o], add W2[o,h]*a[h] for h
 * ascending.  All fp32, multiply and add rounded separately.  (The reference stores z1/a1 for the
 * whole batch, :16; per-point scratch gives the same values.)
 * ---------------------------------------------------------------------------------------- */
void oracle_mlp_forward(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* y,
                        size_t B, size_t In, size_t H, size_t Out) {
    float* a = (float*)malloc(sizeof(float) * (H ? H : 1));
    for (size_t i = 0; i < B; ++i) {
        const float* xi = x + i * In;
        for (size_t h = 0; h < H; ++h) {
            float s = b1[h];
            for (size_t k = 0; k < In; ++k) s += W1[h * In + k] * xi[k];
            a[h] = s > 0.f ? s : 0.f;
        }
        for (size_t o = 0; o < Out; ++o) {
            float s = b2[o];
            for (size_t h = 0; h < H; ++h) s += W2[o * H + h] * a[h];
            y[i * Out + o] = s;
        }
    }
    free(a);
}

/* Deeper MLP: In -> H -> H -> ... -> H -> Out with L >= 1 hidden layers.  PARITY UNPINNED for L > 1: the reference
 * has exactly one hidden layer (include/mlp.h:5-6), so there is nothing of the reference's to check this against.
 * Every layer repeats the reference's layer rule (src/mlp_cpu.cpp:19-24,29-32): start from the bias, add
 * W[g,h]*a[h] for h ascending with separately rounded fp32 multiply and add, ReLU on hidden layers.  For L == 1
 * it must equal oracle_mlp_forward bit for bit (tests/test_oracle_cpu.py), which IS pinned to the reference.
 * Wh: (L-1) matrices [H x H] row-major [g][h]; bh: (L-1) x H. */
void oracle_mlp_forward_deep(const float* x, int L, const float* W1, const float* b1, const float* Wh, const float* bh,
                             const float* W2, const float* b2, float* y, size_t B, size_t In, size_t H, size_t Out) {
    float* a = (float*)malloc(sizeof(float) * (H ? H : 1));
    float* a2 = (float*)malloc(sizeof(float) * (H ? H : 1));
    for (size_t i = 0; i < B; ++i) {
        const float* xi = x + i * In;
        for (size_t h = 0; h < H; ++h) {
            float s = b1[h];
            for (size_t k = 0; k < In; ++k) s += W1[h * In + k] * xi[k];
            a[h] = s > 0.f ? s : 0.f;
        }
        for (int l = 0; l + 1 < L; ++l) {
            const float* W = Wh + (size_t)l * H * H;
            const float* b = bh + (size_t)l * H;
            for (size_t g = 0; g < H; ++g) {
                float s = b[g];
                for (size_t h = 0; h < H; ++h) s += W[g * H + h] * a[h];
                a2[g] = s > 0.f ? s : 0.f;
            }
            float* t = a; a = a2; a2 = t;
        }
        for (size_t o = 0; o < Out; ++o) {
            float s = b2[o];
            for (size_t h = 0; h < H; ++h) s += W2[o * H + h] * a[h];
            y[i * Out + o] = s;
        }
    }
    free(a);
    free(a2);
}

/* MLP over the grid at time t with a deep network (AoS out), coordinates as oracle_mlp_grid_infer. */
void oracle_mlp_grid_infer_deep(const oracle_grid* g, int H, int L, int m1p1, const float* W1, const float* b1, const float* Wh,
                                const float* bh, const float* W2, const float* b2, float t, float* out) {
    const size_t N = (size_t)g->nx * g->ny * g->nz, chunk = 4096;
    float* c = (float*)malloc(sizeof(float) * chunk * 4);
    for (size_t p = 0; p < N; p += chunk) {
        size_t n = N - p < chunk ? N - p : chunk;
        coords_range(g, t, m1p1, p, n, c);
        oracle_mlp_forward_deep(c, L, W1, b1, Wh, bh, W2, b2, out + p * 4, n, 4, (size_t)H, 4);
    }
    free(c);
}

/* MLP backward (MSE weight gradients), src/mlp_cpu.cpp:38-85: forward pass, gz2 = (2/float(B*Out))*(y - target),
 * then every gradient entry accumulated sequentially over the batch in fp32 (i ascending), gz1 =
 * (sum_o gz2[i,o]*W2[o,h]) * (z1 > 0). */
void oracle_mlp_backward(const float* x, const float* y_target, const float* W1, const float* b1, const float* W2,
                         const float* b2, float* dW1, float* db1, float* dW2, float* db2, size_t B, size_t In, size_t H,
                         size_t Out) {
    float* z1 = (float*)malloc(sizeof(float) * B * H);
    float* a1 = (float*)malloc(sizeof(float) * B * H);
    float* gz2 = (float*)malloc(sizeof(float) * B * Out);
    float* gz1 = (float*)malloc(sizeof(float) * B * H);
    for (size_t i = 0; i < B; ++i)
        for (size_t h = 0; h < H; ++h) {
            float s = b1[h];
            for (size_t k = 0; k < In; ++k) s += W1[h * In + k] * x[i * In + k];
            z1[i * H + h] = s;
            a1[i * H + h] = s > 0.f ? s : 0.f;
        }
    const float scale = 2.f / (float)(B * Out);
    for (size_t i = 0; i < B; ++i)
        for (size_t o = 0; o < Out; ++o) {
            float s = b2[o];
            for (size_t h = 0; h < H; ++h) s += W2[o * H + h] * a1[i * H + h];
            gz2[i * Out + o] = scale * (s - y_target[i * Out + o]);
        }
    for (size_t o = 0; o < Out; ++o)
        for (size_t h = 0; h < H; ++h) {
            float s = 0.f;
            for (size_t i = 0; i < B; ++i) s += gz2[i * Out + o] * a1[i * H + h];
            dW2[o * H + h] = s;
        }
    for (size_t o = 0; o < Out; ++o) {
        float s = 0.f;
        for (size_t i = 0; i < B; ++i) s += gz2[i * Out + o];
        db2[o] = s;
    }
    for (size_t i = 0; i < B; ++i)
        for (size_t h = 0; h < H; ++h) {
            float s = 0.f;
            for (size_t o = 0; o < Out; ++o) s += gz2[i * Out + o] * W2[o * H + h];
            gz1[i * H + h] = s * (z1[i * H + h] > 0.f ? 1.f : 0.f);
        }
    for (size_t h = 0; h < H; ++h)
        for (size_t k = 0; k < In; ++k) {
            float s = 0.f;
            for (size_t i = 0; i < B; ++i) s += gz1[i * H + h] * x[i * In + k];
            dW1[h * In + k] = s;
        }
    for (size_t h = 0; h < H; ++h) {
        float s = 0.f;
        for (size_t i = 0; i < B; ++i) s += gz1[i * H + h];
        db1[h] = s;
    }
    free(z1); free(a1); free(gz2); free(gz1);
}

/* MLP over the grid at time t -> AoS [sigma,ux,uy,uz] per point.  src/mlp_grid.cpp:53-59. */
void oracle_mlp_grid_infer(const oracle_grid* g, int In, int H, int Out, int m1p1, const float* W1, const float* b1,
                           const float* W2, const float* b2, float t, float* out) {
    const size_t N = (size_t)g->nx * g->ny * g->nz, chunk = 4096;
    float* c = (float*)malloc(sizeof(float) * chunk * 4);
    for (size_t p = 0; p < N; p += chunk) {
        size_t n = N - p < chunk ? N - p : chunk;
        coords_range(g, t, m1p1, p, n, c);
        oracle_mlp_forward(c, W1, b1, W2, b2, out + p * (size_t)Out, n, (size_t)In, (size_t)H, (size_t)Out);
    }
    free(c);
}

/* AoS outputs (stride 4) -> sigma[N], u[3N] channel-major.  src/mlp_grid.cpp:69-80. */
void oracle_split_fields(const float* y, size_t N, float* sigma, float* u) {
    for (size_t i = 0; i < N; ++i) {
        sigma[i] = y[i * 4 + 0];
        u[i] = y[i * 4 + 1];
        u[N + i] = y[i * 4 + 2];
        u[2 * N + i] = y[i * 4 + 3];
    }
}

/* Three time slices t-dt, t, t+dt (float expressions as on src/mlp_grid.cpp:87-89). */
void oracle_generate_fields(const oracle_grid* g, int In, int H, int Out, int m1p1, const float* W1, const float* b1,
                            const float* W2, const float* b2, float t, float dt, float* s_m, float* s_0, float* s_p,
                            float* u_m, float* u_0, float* u_p) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    float* y = (float*)malloc(sizeof(float) * N * 4);
    const float ts[3] = {t - dt, t, t + dt};
    float* sig[3] = {s_m, s_0, s_p};
    float* vel[3] = {u_m, u_0, u_p};
    for (int s = 0; s < 3; ++s) {
        oracle_mlp_grid_infer(g, In, H, Out, m1p1, W1, b1, W2, b2, ts[s], y);
        oracle_split_fields(y, N, sig[s], vel[s]);
    }
    free(y);
}

/* ------------------------------------------------------------------------------------------
 * PDE residuals, src/phys_cpu.cpp:25-110.  Neighbour index: periodic -> mathematical modulo
 * (:12-15), else clamp to [0,n-1] (:8-10) -- the divisor stays 1/(2h) at clamped edges.  All
 * arithmetic in double on float loads, inverse spacings 1.0/(2.0*double(h)) (:38-41);
 *   R_sigma = d_t sigma + u.grad(sigma) + sigma*div(u)           (:96-103)
 *   R_u[c]  = d_t u_c + (u.grad) u_c                              (:98-106)
 * each sum evaluated left to right as written there, result rounded to float.
 * ---------------------------------------------------------------------------------------- */
static int nb(int v, int n, int periodic) {
    if (periodic) {
        int r = v % n;
        return r < 0 ? r + n : r;
    }
    return v < 0 ? 0 : (v > n - 1 ? n - 1 : v);
}

void oracle_phys_residuals(const oracle_grid* g, const float* s_m, const float* s_0, const float* s_p, const float* u_m,
                           const float* u_0, const float* u_p, float* Rs, float* Rx, float* Ry, float* Rz) {
    const int nx = g->nx, ny = g->ny, nz = g->nz, per = g->periodic;
    const size_t N = (size_t)nx * ny * nz;
    const double i2t = 1.0 / (2.0 * (double)g->dt);
    const double i2x = 1.0 / (2.0 * (double)g->hx);
    const double i2y = 1.0 / (2.0 * (double)g->hy);
    const double i2z = 1.0 / (2.0 * (double)g->hz);
#define LIN(X, Y, Z) ((size_t)(((Z) * ny + (Y)) * nx + (X)))
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t i = LIN(x, y, z);
                const size_t xp = LIN(nb(x + 1, nx, per), y, z), xm = LIN(nb(x - 1, nx, per), y, z);
                const size_t yp = LIN(x, nb(y + 1, ny, per), z), ym = LIN(x, nb(y - 1, ny, per), z);
                const size_t zp = LIN(x, y, nb(z + 1, nz, per)), zm = LIN(x, y, nb(z - 1, nz, per));
                const double st = ((double)s_p[i] - (double)s_m[i]) * i2t;
                const double u[3] = {(double)u_0[i], (double)u_0[N + i], (double)u_0[2 * N + i]};
                double ut[3], gs[3], gu[3][3]; /* gu[c][d] = d u_c / d x_d */
                for (int c = 0; c < 3; ++c) ut[c] = ((double)u_p[c * N + i] - (double)u_m[c * N + i]) * i2t;
                gs[0] = ((double)s_0[xp] - (double)s_0[xm]) * i2x;
                gs[1] = ((double)s_0[yp] - (double)s_0[ym]) * i2y;
                gs[2] = ((double)s_0[zp] - (double)s_0[zm]) * i2z;
                for (int c = 0; c < 3; ++c) {
                    const float* f = u_0 + (size_t)c * N;
                    gu[c][0] = ((double)f[xp] - (double)f[xm]) * i2x;
                    gu[c][1] = ((double)f[yp] - (double)f[ym]) * i2y;
                    gu[c][2] = ((double)f[zp] - (double)f[zm]) * i2z;
                }
                const double div = gu[0][0] + gu[1][1] + gu[2][2];
                const double adv_s = u[0] * gs[0] + u[1] * gs[1] + u[2] * gs[2];
                Rs[i] = (float)(st + adv_s + (double)s_0[i] * div);
                float* Ru[3] = {Rx, Ry, Rz};
                for (int c = 0; c < 3; ++c) {
                    const double adv = u[0] * gu[c][0] + u[1] * gu[c][1] + u[2] * gu[c][2];
                    Ru[c][i] = (float)(ut[c] + adv);
                }
            }
#undef LIN
}

/* First-order UPWIND advection (the "upwind switch" the reference plans and never ships, REQUIREMENT.md:123-134 --
 * PARITY UNPINNED: there is no reference implementation; this restatement is the checker).  Everything is as in
 * oracle_phys_residuals (src/phys_cpu.cpp:66-109: float loads widened to double, central time difference, the
 * divergence in sigma * div(u) by central differences) except the advective derivatives of u . grad(f):
 *     d f / d x_j  ->  (f(x) - f(x - e_j)) / h_j   if u_j(x) > 0,   (f(x + e_j) - f(x)) / h_j   otherwise
 * with the same wrap / clamp neighbour rule (a clamped neighbour is the point itself: the one-sided difference is 0). */
void oracle_phys_residuals_upwind(const oracle_grid* g, const float* s_m, const float* s_0, const float* s_p, const float* u_m,
                                  const float* u_0, const float* u_p, float* Rs, float* Rx, float* Ry, float* Rz) {
    const int nx = g->nx, ny = g->ny, nz = g->nz, per = g->periodic;
    const size_t N = (size_t)nx * ny * nz;
    const double i2t = 1.0 / (2.0 * (double)g->dt);
    const double i2[3] = {1.0 / (2.0 * (double)g->hx), 1.0 / (2.0 * (double)g->hy), 1.0 / (2.0 * (double)g->hz)};
    const double i1[3] = {1.0 / (double)g->hx, 1.0 / (double)g->hy, 1.0 / (double)g->hz};
#define LIN(X, Y, Z) ((size_t)(((Z) * ny + (Y)) * nx + (X)))
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t i = LIN(x, y, z);
                const size_t ip[3] = {LIN(nb(x + 1, nx, per), y, z), LIN(x, nb(y + 1, ny, per), z), LIN(x, y, nb(z + 1, nz, per))};
                const size_t im[3] = {LIN(nb(x - 1, nx, per), y, z), LIN(x, nb(y - 1, ny, per), z), LIN(x, y, nb(z - 1, nz, per))};
                const float* f[4] = {s_0, u_0, u_0 + N, u_0 + 2 * N};
                const float* fp[4] = {s_p, u_p, u_p + N, u_p + 2 * N};
                const float* fm[4] = {s_m, u_m, u_m + N, u_m + 2 * N};
                const double u[3] = {(double)u_0[i], (double)u_0[N + i], (double)u_0[2 * N + i]};
                double div = 0.0;
                for (int d = 0; d < 3; ++d) div += ((double)f[1 + d][ip[d]] - (double)f[1 + d][im[d]]) * i2[d];
                float* R[4] = {Rs, Rx, Ry, Rz};
                for (int c = 0; c < 4; ++c) {
                    double adv = 0.0;
                    for (int d = 0; d < 3; ++d) {
                        const double back = ((double)f[c][i] - (double)f[c][im[d]]) * i1[d];
                        const double fwd = ((double)f[c][ip[d]] - (double)f[c][i]) * i1[d];
                        adv += u[d] * (u[d] > 0.0 ? back : fwd);
                    }
                    double r = ((double)fp[c][i] - (double)fm[c][i]) * i2t + adv;
                    if (c == 0) r += (double)s_0[i] * div;
                    R[c][i] = (float)r;
                }
            }
#undef LIN
}

/* Analytic ("tangent") loss -- PARITY UNPINNED, there is no reference implementation: the reference differences MLP
 * outputs on the grid (src/phys_cpu.cpp:71-93), this propagates the input-derivatives through the network in forward
 * mode (BASELINE.json north_star's literal wording; include/physad_b200.h: physad_tangent_loss_dev).  The hidden
 * pre-activation z_h is formed in fp32 exactly as the forward path does (src/mlp_cpu.cpp:19-22), so the ReLU mask is the
 * forward's; everything after it is in double.  Space derivatives refer to the physical coordinate x = i*hx of the
 * normalised input c = norm(i/(n-1)): dc/dx = (2 or 1) / ((n-1) hx).  Writes the four residual arrays (may be null) and
 * the two double sums. */
void oracle_tangent_loss(const oracle_grid* g, int H, int m1p1, const float* W1, const float* b1, const float* W2, const float* b2,
                         float t, double* acc_s, double* acc_u, float* Rs, float* Rx, float* Ry, float* Rz) {
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double f = m1p1 ? 2.0 : 1.0;
    const double s[3] = {nx > 1 ? f / ((double)(nx - 1) * (double)g->hx) : 0.0, ny > 1 ? f / ((double)(ny - 1) * (double)g->hy) : 0.0,
                         nz > 1 ? f / ((double)(nz - 1) * (double)g->hz) : 0.0};
    const float ct = m1p1 ? t : t + 0.5f; /* src/mlp_grid.cpp:38 */
    double as = 0.0, au = 0.0;
    size_t i = 0;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x, ++i) {
                const float c[4] = {axis_coord(x, nx, m1p1), axis_coord(y, ny, m1p1), axis_coord(z, nz, m1p1), ct};
                double yv[4], J[4][4];
                for (int o = 0; o < 4; ++o) {
                    yv[o] = (double)b2[o];
                    for (int k = 0; k < 4; ++k) J[o][k] = 0.0;
                }
                for (int h = 0; h < H; ++h) {
                    float zh = b1[h];
                    for (int k = 0; k < 4; ++k) zh += W1[h * 4 + k] * c[k];
                    if (zh > 0.f)
                        for (int o = 0; o < 4; ++o) {
                            yv[o] += (double)W2[o * H + h] * (double)zh;
                            for (int k = 0; k < 4; ++k) J[o][k] += (double)W2[o * H + h] * (double)W1[h * 4 + k];
                        }
                }
                double gr[4][3];
                for (int o = 0; o < 4; ++o)
                    for (int j = 0; j < 3; ++j) gr[o][j] = J[o][j] * s[j];
                const double div = gr[1][0] + gr[2][1] + gr[3][2];
                double R[4];
                for (int o = 0; o < 4; ++o) R[o] = J[o][3] + (yv[1] * gr[o][0] + yv[2] * gr[o][1] + yv[3] * gr[o][2]);
                R[0] += yv[0] * div;
                const float r[4] = {(float)R[0], (float)R[1], (float)R[2], (float)R[3]};
                if (Rs) { Rs[i] = r[0]; Rx[i] = r[1]; Ry[i] = r[2]; Rz[i] = r[3]; }
                as += (double)r[0] * r[0];
                au += (double)r[1] * r[1] + (double)r[2] * r[2] + (double)r[3] * r[3];
            }
    *acc_s = as;
    *acc_u = au;
}

/* Sum of squares over a point range, sequential in double as src/phys_cpu.cpp:140-145. */
void oracle_sumsq(const float* Rs, const float* Rx, const float* Ry, const float* Rz, size_t i0, size_t i1, double* acc_s,
                  double* acc_u) {
    double as = 0.0, au = 0.0;
    for (size_t i = i0; i < i1; ++i) {
        as += (double)Rs[i] * Rs[i];
        au += (double)Rx[i] * Rx[i] + (double)Ry[i] * Ry[i] + (double)Rz[i] * Rz[i];
    }
    *acc_s = as;
    *acc_u = au;
}

/* Loss forward, src/phys_cpu.cpp:112-149: L = float(w * acc * (1.0/N)).  Output pointers and the
 * four residual pointers may each be null (:128-136, :147-148). */
void oracle_phys_loss_forward(const oracle_grid* g, float w_sigma, float w_u, const float* s_m, const float* s_0,
                              const float* s_p, const float* u_m, const float* u_0, const float* u_p, float* loss_sigma,
                              float* loss_u, float* Rs, float* Rx, float* Ry, float* Rz) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    float* own[4] = {0, 0, 0, 0};
    float** req[4] = {&Rs, &Rx, &Ry, &Rz};
    for (int k = 0; k < 4; ++k)
        if (!*req[k]) *req[k] = own[k] = (float*)malloc(sizeof(float) * (N ? N : 1));
    oracle_phys_residuals(g, s_m, s_0, s_p, u_m, u_0, u_p, Rs, Rx, Ry, Rz);
    double as, au;
    oracle_sumsq(Rs, Rx, Ry, Rz, 0, N, &as, &au);
    const double invN = 1.0 / (double)N;
    if (loss_sigma) *loss_sigma = (float)(w_sigma * as * invN);
    if (loss_u) *loss_u = (float)(w_u * au * invN);
    for (int k = 0; k < 4; ++k) free(own[k]);
}

/* Residual VJP, src/phys_cpu.cpp:151-170: scale = 2.f*w/float(N) formed in fp32, g = scale*R. */
void oracle_phys_loss_backward(const oracle_grid* g, float w_sigma, float w_u, const float* Rs, const float* Rx,
                               const float* Ry, const float* Rz, float* gs, float* gx, float* gy, float* gz) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    const float ks = 2.f * w_sigma / (float)N, ku = 2.f * w_u / (float)N;
    for (size_t i = 0; i < N; ++i) {
        gs[i] = ks * Rs[i];
        gx[i] = ku * Rx[i];
        gy[i] = ku * Ry[i];
        gz[i] = ku * Rz[i];
    }
}

/* The whole hot path: generate_fields -> loss_forward (call stack A of SURVEY.md section 3).
 * Also returns the raw double sums so multi-rank partial-sum tests have something to compare. */
int oracle_fused_loss(const oracle_grid* g, int In, int H, int Out, int m1p1, const float* W1, const float* b1,
                      const float* W2, const float* b2, float t, float dt, float w_sigma, float w_u, float* loss_sigma,
                      float* loss_u, double* acc_sigma, double* acc_u, float* Rs, float* Rx, float* Ry, float* Rz) {
    if (In != 4 || Out != 4) return -1; /* split_outputs_to_fields hard-codes stride 4 */
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    float* f = (float*)malloc(sizeof(float) * 12 * (N ? N : 1));
    float* own = 0;
    if (!Rs || !Rx || !Ry || !Rz) {
        own = (float*)malloc(sizeof(float) * 4 * (N ? N : 1));
        Rs = own; Rx = own + N; Ry = own + 2 * N; Rz = own + 3 * N;
    }
    oracle_generate_fields(g, In, H, Out, m1p1, W1, b1, W2, b2, t, dt, f, f + N, f + 2 * N, f + 3 * N, f + 6 * N, f + 9 * N);
    oracle_phys_residuals(g, f, f + N, f + 2 * N, f + 3 * N, f + 6 * N, f + 9 * N, Rs, Rx, Ry, Rz);
    double as, au;
    oracle_sumsq(Rs, Rx, Ry, Rz, 0, N, &as, &au);
    if (acc_sigma) *acc_sigma = as;
    if (acc_u) *acc_u = au;
    const double invN = 1.0 / (double)N;
    if (loss_sigma) *loss_sigma = (float)(w_sigma * as * invN);
    if (loss_u) *loss_u = (float)(w_u * au * invN);
    free(f);
    free(own);
    return 0;
}
