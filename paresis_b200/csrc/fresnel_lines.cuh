// One line of the Fresnel propagator's per-axis convolution, start to finish in shared memory (see fresnel.cu for the
// mathematics): load the n core samples of the line, zero-pad to M = 2^k, forward transform, multiply by the kernel's
// spectrum, inverse transform, store the first n samples.  HBM sees one read and one write of the line; the cuFFT
// formulation of the same thing (pad kernel, transform, multiply kernel, transform) moves the M-padded line seven times.
//
//  * decimation in frequency on the way in, decimation in time on the way out, both in place: the spectrum is met in
//    digit-reversed order, which costs nothing because the kernel's spectrum is stored in that order once, when the
//    transfer function is prepared (paresis_fresnel_kernel_create) -- no reordering pass, no second buffer;
//  * radix-8 passes with the line held as float2 in shared memory (index i lives at i + i/16: the last passes walk 8 or
//    16 consecutive elements per thread, the padding spreads them over all banks); the kernel is compiled per log2(M),
//    so every stride is an immediate;
//  * twiddles: one small table per pass, laid out so that consecutive threads read consecutive entries -- W^1 .. W^4
//    of the pass's own root of unity, built in fp64; the other three powers are one complex product away.  (A single
//    table of M-th roots indexed j * p * M/S costs 7 scattered loads per butterfly: measured 4x the shared-memory
//    wavefronts of the transform itself.)
//  * the first pass reads global memory directly (the upper half of the padded line is zero and never loaded), the last
//    pass writes the n wanted samples directly; the innermost pass of both transforms is fused with the spectrum
//    multiply (radix 4, 8 or 16 in registers: forward butterfly, multiply, inverse butterfly), the spectrum stored
//    pair-major so that a warp reads 512 consecutive bytes.
#pragma once
#include "common.cuh"

namespace paresis {

constexpr int FL_MIN_LOG = 9, FL_MAX_LOG = 14;

// log2(M) = 3 * outer + mid, mid in {2, 3, 4}
__host__ __device__ constexpr int fl_log_mid(int log_m) { return log_m % 3 == 0 ? 3 : (log_m % 3 == 1 ? 4 : 2); }
__host__ __device__ constexpr int fl_outer(int log_m) { return (log_m - fl_log_mid(log_m)) / 3; }
__host__ __device__ constexpr int fl_threads(int log_m) { return log_m >= 14 ? 1024 : (log_m >= 12 ? 512 : (log_m >= 11 ? 256 : 64)); }
// twiddle table of pass k (sub-transform size S = M / 8^k): S/8 entries of four float2 (W^1 .. W^4), passes back to back
__host__ __device__ constexpr int fl_tw_offset(int log_m, int k) {
    int off = 0;
    for (int i = 0; i < k; ++i) off += 4 << (log_m - 3 * i - 3);
    return off;
}
__host__ __device__ constexpr int fl_tw_entries(int log_m) { return fl_tw_offset(log_m, fl_outer(log_m)); }
inline size_t fl_smem_bytes(int log_m) { return sizeof(float2) * (size_t)((1 << log_m) + (1 << (log_m - 4)) + 16); }

__device__ __forceinline__ int fl_pad(int i) { return i + (i >> 4); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a * conj(b)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// a * (DIR * i): DIR = -1 is the forward transform's W_4 = -i
template <int DIR> __device__ __forceinline__ float2 mul_qi(float2 a) { return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }

// r-point DFTs on registers, natural order in and out; DIR = -1 forward, +1 inverse (unnormalised)
template <int DIR> __device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_qi<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2); a1 = cadd(t1, t3); a3 = csub(t1, t3);
}
template <int DIR> __device__ __forceinline__ void dft8(float2* v) {
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<DIR>(e0, e1, e2, e3);
    dft4<DIR>(o0, o1, o2, o3);
    constexpr float h = 0.70710678118654752f;
    // W_8^k = exp(DIR * 2 pi i k / 8)
    o1 = cmul(o1, make_float2(h, DIR * h));
    o2 = mul_qi<DIR>(o2);
    o3 = cmul(o3, make_float2(-h, DIR * h));
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
template <int DIR> __device__ __forceinline__ void dft16(float2* v) {
    float2 e[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    dft8<DIR>(e);
    dft8<DIR>(o);
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    const float2 w[8] = {make_float2(1.f, 0.f), make_float2(c1, DIR * s1), make_float2(h, DIR * h), make_float2(s1, DIR * c1),
                         make_float2(0.f, DIR * 1.f), make_float2(-s1, DIR * c1), make_float2(-h, DIR * h), make_float2(-c1, DIR * s1)};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float2 t = k == 0 ? o[0] : (k == 4 ? mul_qi<DIR>(o[4]) : cmul(o[k], w[k]));
        v[k] = cadd(e[k], t);
        v[k + 8] = csub(e[k], t);
    }
}
template <int LR, int DIR> __device__ __forceinline__ void dft_small(float2* v) {
    if (LR == 2) dft4<DIR>(v[0], v[1], v[2], v[3]);
    else if (LR == 3) dft8<DIR>(v);
    else dft16<DIR>(v);
}

// the seven twiddles W^p of one butterfly from the table's (W, W^2, W^3, W^4)
__device__ __forceinline__ void fl_twiddles(const float2* __restrict__ t, float2* w) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(t)), b = __ldg(reinterpret_cast<const float4*>(t) + 1);
    w[1] = make_float2(a.x, a.y);
    w[2] = make_float2(a.z, a.w);
    w[3] = make_float2(b.x, b.y);
    w[4] = make_float2(b.z, b.w);
    w[5] = cmul(w[4], w[1]);
    w[6] = cmul(w[4], w[2]);
    w[7] = cmul(w[4], w[3]);
}

// Forward radix-8 pass over sub-transforms of size 2^LOG_S: x[j + q S/8] -> W_S^(j p) DFT_8[p] at p S/8 + j.
// Element i of the line is at line[fl_pad(i)]; the eight addresses are fl_pad(base) + q S/8 + (q S/8) / 16, whatever S
// (base = block * S + j with j < S/8, and S/8 either divides 16 or is a multiple of it).
template <int LOG_M, int LOG_S, int THREADS, bool FROM_GLOBAL>
__device__ __forceinline__ void pass_forward(float2* line, const float2* __restrict__ src, int n, const float2* __restrict__ tw, int tid) {
    constexpr int LOG_Q = LOG_S - 3, Q = 1 << LOG_Q;
#pragma unroll 1
    for (int b0 = 0; b0 < (1 << (LOG_M - 3)); b0 += THREADS) {
        const int b = b0 + tid;
        const int j = b & (Q - 1), base = ((b >> LOG_Q) << LOG_S) + j;
        float2* const p = line + fl_pad(base);
        PARESIS_BOUND(fl_pad(base) + 7 * Q + ((7 * Q) >> 4), (1 << LOG_M) + (1 << (LOG_M - 4)) + 16);
        PARESIS_BOUND(fl_pad(base) + 7 * Q + ((7 * Q) >> 4) - fl_pad(base + 7 * Q), 1);      // the closed form IS fl_pad
        float2 v[8];
        if (FROM_GLOBAL) {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = base + q * Q < n ? __ldg(src + base + q * Q) : make_float2(0.f, 0.f);
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = p[q * Q + ((q * Q) >> 4)];
        }
        float2 w[8];
        fl_twiddles(tw + 4 * j, w);
        dft8<-1>(v);
        p[0] = v[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) p[k * Q + ((k * Q) >> 4)] = cmul(v[k], w[k]);
    }
}

// Its inverse: W_S^(-j p) x[p S/8 + j] -> inverse DFT_8 -> j + q S/8
template <int LOG_M, int LOG_S, int THREADS, bool TO_GLOBAL>
__device__ __forceinline__ void pass_inverse(float2* line, float2* __restrict__ dst, int n, const float2* __restrict__ tw, int tid) {
    constexpr int LOG_Q = LOG_S - 3, Q = 1 << LOG_Q;
#pragma unroll 1
    for (int b0 = 0; b0 < (1 << (LOG_M - 3)); b0 += THREADS) {
        const int b = b0 + tid;
        const int j = b & (Q - 1), base = ((b >> LOG_Q) << LOG_S) + j;
        const float2* const p = line + fl_pad(base);
        PARESIS_BOUND(fl_pad(base) + 7 * Q + ((7 * Q) >> 4), (1 << LOG_M) + (1 << (LOG_M - 4)) + 16);
        PARESIS_BOUND(fl_pad(base) + 7 * Q + ((7 * Q) >> 4) - fl_pad(base + 7 * Q), 1);
        float2 w[8];
        fl_twiddles(tw + 4 * j, w);
        float2 v[8];
        v[0] = p[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) v[k] = cmul_conj(p[k * Q + ((k * Q) >> 4)], w[k]);
        dft8<1>(v);
        if (TO_GLOBAL) {
#pragma unroll
            for (int q = 0; q < 8; ++q) if (base + q * Q < n) dst[base + q * Q] = v[q];
        } else {
            float2* const o = line + fl_pad(base);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q * Q + ((q * Q) >> 4)] = v[q];
        }
    }
}

// The innermost pass of both transforms around the spectrum multiply: groups of R = 2^LR consecutive elements.
// g_mid[h * (M/R) + b] = (G_dr[b R + 2h], G_dr[b R + 2h + 1]): consecutive threads read consecutive 16-byte entries.
template <int LOG_M, int LR, int THREADS>
__device__ __forceinline__ void pass_middle(float2* line, const float4* __restrict__ g_mid, int tid) {
    constexpr int R = 1 << LR, GROUPS = 1 << (LOG_M - LR);
#pragma unroll 1
    for (int b0 = 0; b0 < GROUPS; b0 += THREADS) {
        const int b = b0 + tid;
        if (GROUPS < THREADS && b >= GROUPS) break;
        float2* const p = line + b * R + ((b * R) >> 4);      // R <= 16: a group never straddles a padding slot
        PARESIS_BOUND(b * R + ((b * R) >> 4) + R - 1, (1 << LOG_M) + (1 << (LOG_M - 4)) + 16);
        PARESIS_BOUND(fl_pad(b * R + R - 1) - (b * R + ((b * R) >> 4) + R - 1), 1);
        PARESIS_BOUND((R / 2 - 1) * GROUPS + b, 1 << (LOG_M - 1));
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = p[q];
        dft_small<LR, -1>(v);
#pragma unroll
        for (int h = 0; h < R / 2; ++h) {
            const float4 g = __ldg(g_mid + h * GROUPS + b);
            v[2 * h] = cmul(v[2 * h], make_float2(g.x, g.y));
            v[2 * h + 1] = cmul(v[2 * h + 1], make_float2(g.z, g.w));
        }
        dft_small<LR, 1>(v);
#pragma unroll
        for (int q = 0; q < R; ++q) p[q] = v[q];
    }
}

template <int LOG_M, int K, int OUTER, int THREADS> struct FlOuter {
    // passes K .. OUTER-1 forward, the middle pass, then the same passes backwards
    static __device__ __forceinline__ void run(float2* line, const float2* src, float2* dst, int n, const float2* tw, const float4* g_mid, int tid) {
        constexpr int LOG_S = LOG_M - 3 * K;
        if (K > 0) __syncthreads();
        pass_forward<LOG_M, LOG_S, THREADS, K == 0>(line, src, n, tw + fl_tw_offset(LOG_M, K), tid);
        FlOuter<LOG_M, K + 1, OUTER, THREADS>::run(line, src, dst, n, tw, g_mid, tid);
        __syncthreads();
        pass_inverse<LOG_M, LOG_S, THREADS, K == 0>(line, dst, n, tw + fl_tw_offset(LOG_M, K), tid);
    }
};
template <int LOG_M, int OUTER, int THREADS> struct FlOuter<LOG_M, OUTER, OUTER, THREADS> {
    static __device__ __forceinline__ void run(float2* line, const float2*, float2*, int, const float2*, const float4* g_mid, int tid) {
        __syncthreads();
        pass_middle<LOG_M, fl_log_mid(LOG_M), THREADS>(line, g_mid, tid);
    }
};

// in: lines x n (row-major), out: lines x n = the first n samples of the circular convolution of the zero-padded line
// with the kernel whose digit-reversed, pair-major spectrum (scaled by 1/M) is g_mid.  One block per line.
template <int LOG_M>
__global__ void __launch_bounds__(fl_threads(LOG_M), fl_threads(LOG_M) >= 1024 ? 1 : 2)
line_convolve_kernel(const float2* __restrict__ in, int n, const float2* __restrict__ tw, const float4* __restrict__ g_mid,
                     float2* __restrict__ out) {
    extern __shared__ __align__(16) float2 fl_line[];
    FlOuter<LOG_M, 0, fl_outer(LOG_M), fl_threads(LOG_M)>::run(fl_line, in + (size_t)blockIdx.x * n, out + (size_t)blockIdx.x * n, n, tw,
                                                               g_mid, threadIdx.x);
}

// position in the digit-reversed spectrum -> frequency index, for the pass structure above
__host__ __device__ inline int fl_frequency_of(int pos, int log_m) {
    // pos = p_1 (M/8) + p_2 (M/64) + ... + p_mid;  k = p_1 + 8 (p_2 + 8 (... + 8^(outer-1) p_mid))
    const int outer = fl_outer(log_m), log_mid = fl_log_mid(log_m);
    int k = 0, weight_log = 0, shift = log_m;
    for (int s = 0; s < outer; ++s) {
        shift -= 3;
        k += ((pos >> shift) & 7) << weight_log;
        weight_log += 3;
    }
    k += (pos & ((1 << log_mid) - 1)) << weight_log;
    return k;
}

// log2(M) if M is a power of two the kernel is compiled for, else 0
inline int line_fft_log(int M) {
    int log_m = 0;
    while ((1 << log_m) < M) ++log_m;
    return ((1 << log_m) == M && log_m >= FL_MIN_LOG && log_m <= FL_MAX_LOG) ? log_m : 0;
}

}  // namespace paresis
