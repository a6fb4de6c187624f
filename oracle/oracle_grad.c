/*
 * TEST INFRASTRUCTURE ONLY -- CPU checker for the closed-loop gradient  d(L_sigma + L_u) / d(MLP weights).
 *
 * PARITY UNPINNED: the reference has no such function.  Its backward stops at dL/dR
 * (src/phys_cpu.cpp:151-170) and its MLP backward (src/mlp_cpu.cpp:38-85) is the MSE one; chaining
 * them through the stencil is the planned "closed loop" of REQUIREMENT.md:155-169 (SURVEY.md 8f
 * rank 1).  What this file restates from the reference is the forward it differentiates (same
 * formulas and evaluation order as oracle.c: MLP src/mlp_cpu.cpp:14-36, residuals
 * src/phys_cpu.cpp:25-110, loss :140-148, dL/dR scale :162-163).  The adjoint itself is validated by
 * central finite differences of an all-double loss (tests/test_oracle_cpu.py), not by a reference.
 *
 * Two modes:
 *   all_double = 0  fields by the strict-fp32 MLP, residuals rounded to float, g = (2w/float(N))*R in
 *                   fp32 -- i.e. exactly the numbers the reference forward/backward would hand to an
 *                   MLP backward; the adjoint of the stencil and of the MLP then runs in double.
 *                   This is what the CUDA kernel is compared with.
 *   all_double = 1  everything in double (weights promoted): the differentiable function whose
 *                   finite differences pin the adjoint.
 * Gradient layout: dW1[H*4] (row-major, as W1), db1[H], dW2[4*H] (as W2), db2[4]  -> 9H+4 doubles.
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int nx, ny, nz;
    float hx, hy, hz, dt;
    int periodic;
} oracle_grid;

static float axis_coord(int i, int n, int m1p1) { /* src/mlp_grid.cpp:25-29 */
    if (n <= 1) return 0.0f;
    float u = (float)i / (float)(n - 1);
    return m1p1 ? 2.f * u - 1.f : u;
}

static int nb(int v, int n, int periodic) { /* src/phys_cpu.cpp:8-15 */
    if (periodic) {
        int r = v % n;
        return r < 0 ? r + n : r;
    }
    return v < 0 ? 0 : (v > n - 1 ? n - 1 : v);
}

/* hidden pre-activations and outputs of one point: fp32 strict (th_f) or double (th_d) */
static void mlp_point_f(int H, const float* th, const float x[4], float* z, double y[4]) {
    const float *W1 = th, *b1 = th + 4 * H, *W2 = th + 5 * H, *b2 = th + 9 * H;
    for (int h = 0; h < H; ++h) {
        float s = b1[h];
        for (int k = 0; k < 4; ++k) s += W1[h * 4 + k] * x[k];
        z[h] = s;
    }
    for (int o = 0; o < 4; ++o) {
        float s = b2[o];
        for (int h = 0; h < H; ++h) s += W2[o * H + h] * (z[h] > 0.f ? z[h] : 0.f);
        y[o] = (double)s;
    }
}

static void mlp_point_d(int H, const double* th, const float x[4], double* z, double y[4]) {
    const double *W1 = th, *b1 = th + 4 * H, *W2 = th + 5 * H, *b2 = th + 9 * H;
    for (int h = 0; h < H; ++h) {
        double s = b1[h];
        for (int k = 0; k < 4; ++k) s += W1[h * 4 + k] * (double)x[k];
        z[h] = s;
    }
    for (int o = 0; o < 4; ++o) {
        double s = b2[o];
        for (int h = 0; h < H; ++h) s += W2[o * H + h] * (z[h] > 0.0 ? z[h] : 0.0);
        y[o] = s;
    }
}

typedef struct {
    const oracle_grid* g;
    int H, m1p1, all_double;
    const float* th_f;
    const double* th_d;
    float ts[3]; /* network time input of slices t-dt, t, t+dt (src/mlp_grid.cpp:38, :87-89) */
} model;

static void point_coords(const model* m, size_t p, int s, float x[4]) {
    const oracle_grid* g = m->g;
    x[0] = axis_coord((int)(p % (size_t)g->nx), g->nx, m->m1p1);
    x[1] = axis_coord((int)((p / (size_t)g->nx) % (size_t)g->ny), g->ny, m->m1p1);
    x[2] = axis_coord((int)(p / ((size_t)g->nx * (size_t)g->ny)), g->nz, m->m1p1);
    x[3] = m->ts[s];
}

/* F[(s*4 + c)*N + i]: channel c (sigma, ux, uy, uz) of slice s.  R[c*N + i].  Returns sum of squares. */
static void forward_all(const model* m, double* F, double* R, double* acc_s, double* acc_u) {
    const oracle_grid* g = m->g;
    const int nx = g->nx, ny = g->ny, nz = g->nz, per = g->periodic, H = m->H;
    const size_t N = (size_t)nx * ny * nz;
    float* zf = (float*)malloc(sizeof(float) * H);
    double* zd = (double*)malloc(sizeof(double) * H);
    for (int s = 0; s < 3; ++s)
        for (size_t i = 0; i < N; ++i) {
            float x[4];
            double y[4];
            point_coords(m, i, s, x);
            if (m->all_double) mlp_point_d(H, m->th_d, x, zd, y);
            else mlp_point_f(H, m->th_f, x, zf, y);
            for (int c = 0; c < 4; ++c) F[((size_t)s * 4 + c) * N + i] = y[c];
        }
    free(zf);
    free(zd);
    const double i2t = 1.0 / (2.0 * (double)g->dt), i2x = 1.0 / (2.0 * (double)g->hx);
    const double i2y = 1.0 / (2.0 * (double)g->hy), i2z = 1.0 / (2.0 * (double)g->hz);
    const double *Fm = F, *F0 = F + 4 * N, *Fp = F + 8 * N;
    double as = 0.0, au = 0.0;
#define LIN(X, Y, Z) ((size_t)(((Z) * ny + (Y)) * nx + (X)))
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t i = LIN(x, y, z);
                const size_t xp = LIN(nb(x + 1, nx, per), y, z), xm = LIN(nb(x - 1, nx, per), y, z);
                const size_t yp = LIN(x, nb(y + 1, ny, per), z), ym = LIN(x, nb(y - 1, ny, per), z);
                const size_t zp = LIN(x, y, nb(z + 1, nz, per)), zm = LIN(x, y, nb(z - 1, nz, per));
                double dt_[4], gr[4][3];
                for (int c = 0; c < 4; ++c) {
                    const double* f = F0 + (size_t)c * N;
                    dt_[c] = (Fp[(size_t)c * N + i] - Fm[(size_t)c * N + i]) * i2t;
                    gr[c][0] = (f[xp] - f[xm]) * i2x;
                    gr[c][1] = (f[yp] - f[ym]) * i2y;
                    gr[c][2] = (f[zp] - f[zm]) * i2z;
                }
                const double u[3] = {F0[N + i], F0[2 * N + i], F0[3 * N + i]};
                const double div = gr[1][0] + gr[2][1] + gr[3][2];
                const double adv_s = u[0] * gr[0][0] + u[1] * gr[0][1] + u[2] * gr[0][2];
                double r[4];
                r[0] = dt_[0] + adv_s + F0[i] * div; /* src/phys_cpu.cpp:103 */
                for (int c = 1; c < 4; ++c) r[c] = dt_[c] + (u[0] * gr[c][0] + u[1] * gr[c][1] + u[2] * gr[c][2]); /* :104-106 */
                for (int c = 0; c < 4; ++c) {
                    if (!m->all_double) r[c] = (double)(float)r[c];
                    R[(size_t)c * N + i] = r[c];
                }
                as += r[0] * r[0];
                au += r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
            }
#undef LIN
    *acc_s = as;
    *acc_u = au;
}

/* All-double loss L_sigma + L_u of double weights: the function the finite differences probe. */
int oracle_phys_loss_double(const oracle_grid* g, int H, int m1p1, const double* theta, float t, float dt, float w_sigma,
                            float w_u, double* loss_sigma, double* loss_u) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    if (!N) return -1;
    model m = {g, H, m1p1, 1, 0, theta, {0, 0, 0}};
    const float off = m1p1 ? 0.f : 0.5f;
    m.ts[0] = (t - dt) + off; m.ts[1] = t + off; m.ts[2] = (t + dt) + off;
    double* F = (double*)malloc(sizeof(double) * 12 * N);
    double* R = (double*)malloc(sizeof(double) * 4 * N);
    double as, au;
    forward_all(&m, F, R, &as, &au);
    *loss_sigma = (double)w_sigma * as / (double)N;
    *loss_u = (double)w_u * au / (double)N;
    free(F);
    free(R);
    return 0;
}

/* transposed central difference along one axis: sum over the points p whose +1 neighbour is q minus
 * those whose -1 neighbour is q, of c(p).  Periodic: p = q-1 and q+1 (wrapped).  Clamped (the reference's
 * rule keeps the 1/(2h) divisor at the edge, src/phys_cpu.cpp:8-10): the edge point is its own neighbour. */
static double transposed_diff(const double* c, size_t base, size_t stride, int q, int n, int per) {
    if (per) return c[base + (size_t)nb(q - 1, n, 1) * stride] - c[base + (size_t)nb(q + 1, n, 1) * stride];
    const double lo = q >= 1 ? c[base + (size_t)(q - 1) * stride] : -c[base + (size_t)q * stride];
    const double hi = q <= n - 2 ? c[base + (size_t)(q + 1) * stride] : -c[base + (size_t)q * stride];
    return lo - hi;
}

int oracle_phys_loss_grad(const oracle_grid* g, int H, int m1p1, const float* W1, const float* b1, const float* W2,
                          const float* b2, float t, float dt, float w_sigma, float w_u, int all_double,
                          double* loss_sigma, double* loss_u, double* grad) {
    const int nx = g->nx, ny = g->ny, nz = g->nz, per = g->periodic;
    const size_t N = (size_t)nx * ny * nz;
    const int NG = 9 * H + 4;
    if (!N) return -1;
    float* th_f = (float*)malloc(sizeof(float) * NG);
    double* th_d = (double*)malloc(sizeof(double) * NG);
    memcpy(th_f, W1, sizeof(float) * 4 * H);
    memcpy(th_f + 4 * H, b1, sizeof(float) * H);
    memcpy(th_f + 5 * H, W2, sizeof(float) * 4 * H);
    memcpy(th_f + 9 * H, b2, sizeof(float) * 4);
    for (int i = 0; i < NG; ++i) th_d[i] = (double)th_f[i];
    model m = {g, H, m1p1, all_double, th_f, th_d, {0, 0, 0}};
    const float off = m1p1 ? 0.f : 0.5f;
    m.ts[0] = (t - dt) + off; m.ts[1] = t + off; m.ts[2] = (t + dt) + off;

    double* F = (double*)malloc(sizeof(double) * 12 * N);
    double* R = (double*)malloc(sizeof(double) * 4 * N);
    double as, au;
    forward_all(&m, F, R, &as, &au);
    if (loss_sigma) *loss_sigma = (double)w_sigma * as / (double)N;
    if (loss_u) *loss_u = (double)w_u * au / (double)N;

    /* g = dL/dR.  Mode 0 follows src/phys_cpu.cpp:162-163: the scale is formed and applied in fp32. */
    double* G = (double*)malloc(sizeof(double) * 4 * N);
    for (int c = 0; c < 4; ++c) {
        const float w = c == 0 ? w_sigma : w_u;
        if (all_double) {
            const double k = 2.0 * (double)w / (double)N;
            for (size_t i = 0; i < N; ++i) G[(size_t)c * N + i] = k * R[(size_t)c * N + i];
        } else {
            const float k = 2.f * w / (float)N;
            for (size_t i = 0; i < N; ++i) G[(size_t)c * N + i] = (double)(k * (float)R[(size_t)c * N + i]);
        }
    }
    /* stencil fluxes at every point: C[(j*4 + c)*N + p] multiplies d_j(field c) evaluated at p in the loss:
     *   R_sigma = ... + u_j d_j sigma + sigma d_j u_j     -> c=0: g_s u_j ;   c=j+1 gets + g_s sigma
     *   R_ui    = ... + u_j d_j u_i                       -> c=i+1: g_ui u_j                                  */
    const double* F0 = F + 4 * N;
    double* Cf = (double*)malloc(sizeof(double) * 12 * N);
    for (int j = 0; j < 3; ++j)
        for (size_t p = 0; p < N; ++p) {
            const double uj = F0[(size_t)(j + 1) * N + p];
            Cf[((size_t)j * 4 + 0) * N + p] = G[p] * uj;
            for (int i = 0; i < 3; ++i)
                Cf[((size_t)j * 4 + i + 1) * N + p] = G[(size_t)(i + 1) * N + p] * uj + (i == j ? G[p] * F0[p] : 0.0);
        }
    const double i2t = 1.0 / (2.0 * (double)g->dt);
    const double i2h[3] = {1.0 / (2.0 * (double)g->hx), 1.0 / (2.0 * (double)g->hy), 1.0 / (2.0 * (double)g->hz)};
    for (int i = 0; i < NG; ++i) grad[i] = 0.0;
    double *dW1 = grad, *db1 = grad + 4 * H, *dW2 = grad + 5 * H, *db2 = grad + 9 * H;
    float* zf = (float*)malloc(sizeof(float) * H);
    double* zd = (double*)malloc(sizeof(double) * H);
#define LIN(X, Y, Z) ((size_t)(((Z) * ny + (Y)) * nx + (X)))
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const size_t q = LIN(x, y, z);
                const size_t xp = LIN(nb(x + 1, nx, per), y, z), xm = LIN(nb(x - 1, nx, per), y, z);
                const size_t yp = LIN(x, nb(y + 1, ny, per), z), ym = LIN(x, nb(y - 1, ny, per), z);
                const size_t zp = LIN(x, y, nb(z + 1, nz, per)), zm = LIN(x, y, nb(z - 1, nz, per));
                const size_t ip[3] = {xp, yp, zp}, im[3] = {xm, ym, zm};
                /* adjoint of the time-t fields at q */
                double gr[4][3];
                for (int c = 0; c < 4; ++c)
                    for (int j = 0; j < 3; ++j) gr[c][j] = (F0[(size_t)c * N + ip[j]] - F0[(size_t)c * N + im[j]]) * i2h[j];
                double A[3][4]; /* A[s][c]: dL/d field c of slice s at q */
                A[1][0] = G[q] * (gr[1][0] + gr[2][1] + gr[3][2]);
                for (int j = 0; j < 3; ++j)
                    A[1][j + 1] = G[q] * gr[0][j] + G[N + q] * gr[1][j] + G[2 * N + q] * gr[2][j] + G[3 * N + q] * gr[3][j];
                const int qi[3] = {x, y, z}, nn[3] = {nx, ny, nz};
                const size_t stride[3] = {1, (size_t)nx, (size_t)nx * ny};
                for (int j = 0; j < 3; ++j) {
                    const size_t base = q - (size_t)qi[j] * stride[j];
                    for (int c = 0; c < 4; ++c)
                        A[1][c] += i2h[j] * transposed_diff(Cf + ((size_t)j * 4 + c) * N, base, stride[j], qi[j], nn[j], per);
                }
                for (int c = 0; c < 4; ++c) {
                    A[2][c] = G[(size_t)c * N + q] * i2t;
                    A[0][c] = -A[2][c];
                }
                /* MLP backward of each slice with upstream gradient A[s] */
                for (int s = 0; s < 3; ++s) {
                    float xin[4];
                    double yy[4];
                    point_coords(&m, q, s, xin);
                    if (all_double) mlp_point_d(H, th_d, xin, zd, yy);
                    else {
                        mlp_point_f(H, th_f, xin, zf, yy);
                        for (int h = 0; h < H; ++h) zd[h] = (double)zf[h];
                    }
                    for (int o = 0; o < 4; ++o) db2[o] += A[s][o];
                    for (int h = 0; h < H; ++h) {
                        const double a = zd[h] > 0.0 ? zd[h] : 0.0;
                        double da = 0.0;
                        for (int o = 0; o < 4; ++o) {
                            dW2[o * H + h] += A[s][o] * a;
                            da += th_d[5 * H + o * H + h] * A[s][o];
                        }
                        const double dz = zd[h] > 0.0 ? da : 0.0;
                        db1[h] += dz;
                        for (int k = 0; k < 4; ++k) dW1[h * 4 + k] += dz * (double)xin[k];
                    }
                }
            }
#undef LIN
    free(zf); free(zd); free(Cf); free(G); free(R); free(F); free(th_f); free(th_d);
    return 0;
}
