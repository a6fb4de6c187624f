// See tangent_kernels.cuh.
#include "tangent_kernels.cuh"

#include "fused_loss.cuh"

namespace physad {

namespace {

__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int H>
__global__ void __launch_bounds__(TANGENT_THREADS, 2) k_tangent_loss(const __grid_constant__ TangentConst<H> w, const TangentArgs a) {
    constexpr int PP = 2;   // points per thread and pass: 2 x 20 independent accumulator chains
    __shared__ double2 s_red[TANGENT_THREADS / 32];
    __shared__ unsigned int s_flag;
    const long long n = (long long)(a.z_end - a.z_begin) * a.ny * a.nx;
    const long long stride = (long long)gridDim.x * TANGENT_THREADS;
    const int plane = a.nx * a.ny;
    double acc_s = 0.0, acc_u = 0.0;
    for (long long p0 = (long long)blockIdx.x * TANGENT_THREADS + threadIdx.x; p0 < n; p0 += PP * stride) {
        // packed f32x2 throughout (one issue slot per two lane-operations): the two points of a pass share every packed
        // instruction of the pre-activation; y and the Jacobian rows are pairs {y0,y1},{y2,y3} and {J[o][0],J[o][1]},{J[o][2],J[o][3]}
        static_assert(PP == 2, "the pre-activation is packed over the two points of a pass");
        float c[PP][3];
        f32x2 y01[PP], y23[PP], Ja[PP][4], Jb[PP][4];
        bool live[PP];
#pragma unroll
        for (int q = 0; q < PP; ++q) {
            const long long p = p0 + q * stride;
            live[q] = p < n;
            const long long pc = live[q] ? p : n - 1;
            const int zl = int(pc / plane), rem = int(pc - (long long)zl * plane);
            const int yy = rem / a.nx, xx = rem - yy * a.nx;
            c[q][0] = __ldg(a.cxs + xx); c[q][1] = __ldg(a.cys + yy); c[q][2] = __ldg(a.czs + a.z_begin + zl);
            y01[q] = pack2(w.b2.x, w.b2.y); y23[q] = pack2(w.b2.z, w.b2.w);
#pragma unroll
            for (int o = 0; o < 4; ++o) { Ja[q][o] = 0ull; Jb[q][o] = 0ull; }
        }
        const f32x2 cx2 = pack2(c[0][0], c[1][0]), cy2 = pack2(c[0][1], c[1][1]), cz2 = pack2(c[0][2], c[1][2]);
#pragma unroll 4
        for (int h = 0; h < H; ++h) {
            const float4 w1 = w.w1[h], w2 = w.w2[h];
            const float b1 = w.b1[h];
            // the forward path's roundings (src/mlp_cpu.cpp:19-22), two points per instruction: the masks equal the reference
            // forward's bit for bit (separate multiply and add: the half-swap keeps ptxas from contracting them)
            f32x2 z2 = add2_rn_swapped(bcast2(b1), mul2_rn(bcast2(w1.x), pack2(c[1][0], c[0][0])));
            z2 = add2_rn_swapped(z2, mul2_rn(bcast2(w1.y), pack2(c[1][1], c[0][1])));
            z2 = add2_rn_swapped(z2, mul2_rn(bcast2(w1.z), pack2(c[1][2], c[0][2])));
            z2 = add2_rn(z2, bcast2(w1.w));
            float zq[PP];
            unpack2(z2, zq[0], zq[1]);
            const f32x2 w2a = pack2(w2.x, w2.y), w2b = pack2(w2.z, w2.w);
#pragma unroll
            for (int q = 0; q < PP; ++q) {
                const f32x2 act = bcast2(fmaxf(zq[q], 0.f)), m = bcast2(zq[q] > 0.f ? 1.f : 0.f);
                y01[q] = fma2(w2a, act, y01[q]);
                y23[q] = fma2(w2b, act, y23[q]);
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float4 pv = w.p[h][o];
                    Ja[q][o] = fma2(m, pack2(pv.x, pv.y), Ja[q][o]);
                    Jb[q][o] = fma2(m, pack2(pv.z, pv.w), Jb[q][o]);
                }
            }
        }
        (void)cx2; (void)cy2; (void)cz2;
        float y[PP][4], J[PP][4][4];
#pragma unroll
        for (int q = 0; q < PP; ++q) {
            unpack2(y01[q], y[q][0], y[q][1]);
            unpack2(y23[q], y[q][2], y[q][3]);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                unpack2(Ja[q][o], J[q][o][0], J[q][o][1]);
                unpack2(Jb[q][o], J[q][o][2], J[q][o][3]);
            }
        }
#pragma unroll
        for (int q = 0; q < PP; ++q) {
            if (!live[q]) continue;
            const float s[3] = {a.sx, a.sy, a.sz};
            float g[4][3];   // d y_o / d x_j
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int j = 0; j < 3; ++j) g[o][j] = J[q][o][j] * s[j];
            const float div = (g[1][0] + g[2][1]) + g[3][2];
            float R[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) R[o] = J[q][o][3] + ((y[q][1] * g[o][0] + y[q][2] * g[o][1]) + y[q][3] * g[o][2]);
            R[0] += y[q][0] * div;
            acc_s += double(R[0]) * double(R[0]);
            acc_u += double(R[1]) * double(R[1]) + double(R[2]) * double(R[2]) + double(R[3]) * double(R[3]);
            if (a.R[0]) {
                const long long p = p0 + q * stride;
                a.R[0][p] = R[0]; a.R[1][p] = R[1]; a.R[2][p] = R[2]; a.R[3][p] = R[3];
            }
        }
    }
    grid_reduce2<TANGENT_THREADS / 32>(acc_s, acc_u, a.partials, a.ticket, a.acc_out, s_red, &s_flag);
}

template <int H>
int launch_t(const void* k, const TangentArgs& a, int blocks, cudaStream_t st) {
    k_tangent_loss<H><<<blocks, TANGENT_THREADS, 0, st>>>(*static_cast<const TangentConst<H>*>(k), a);
    return int(cudaGetLastError());
}

}  // namespace

int tangent_launch(int H, const void* tangent_const, const TangentArgs& a, int blocks, cudaStream_t st) {
    switch (H) {
        case 32: return launch_t<32>(tangent_const, a, blocks, st);
        case 64: return launch_t<64>(tangent_const, a, blocks, st);
        case 128: return launch_t<128>(tangent_const, a, blocks, st);
    }
    return int(cudaErrorInvalidValue);
}

}  // namespace physad
