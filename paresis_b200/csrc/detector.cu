// Detector model: source blur -> os x os sum binning -> PSF blur -> crop, then Poisson noise.
//
// Reference: Detector.py:79-119 (detection), :185-198 (resize), :201-220 (create_gaussian_shape).
// The reference pads by reflection (15*os), convolves with scipy.signal.fftconvolve(mode='same')
// (zero extension beyond the padded frame), bins, convolves again and crops 15 detector pixels.
// create_gaussian_shape is an outer product of a normalised 1-D Gaussian with itself, so both
// blurs are done as exact separable direct convolutions; blur + binning collapse into ONE strided
// pass per axis with the composite kernel W(d) = sum_{a<os} g(d - a).  The reflect padding is
// index arithmetic, never materialised.  Everything is fp32, HBM-bound streaming.
#include <math.h>

#include "common.cuh"

namespace paresis {

constexpr int DET_PAD = 15;     // Detector.py:92
constexpr int DET_THREADS = 128;
constexpr int MAX_TAPS = 1024;  // os + 2*half of the composite kernel, kept in shared memory

__device__ __forceinline__ int reflect_index(int q, int n) {
    // numpy.pad(mode='reflect'): edge sample not repeated; one bounce is enough for pad <= n-1
    if (q < 0) q = -q;
    if (q >= n) q = 2 * (n - 1) - q;
    return q;
}

__device__ __forceinline__ void build_composite(float* W, const float* __restrict__ g, int half, int os) {
    // W[t], t = d + half, d in [-half, os-1+half]
    const int taps = os + 2 * half;
    for (int t = threadIdx.x; t < taps; t += blockDim.x) {
        const int d = t - half;
        float s = 0.f;
        for (int a = 0; a < os; ++a) {
            const int e = d - a;
            if (e >= -half && e <= half) s += g ? g[e + half] : 1.f;
        }
        W[t] = s;
    }
    __syncthreads();
}

// Pass A: along columns.  T1[xp][v] = sum_d W(d) * P(xp, v*os + d), P = reflect-padded image, 0 outside.
__global__ void __launch_bounds__(DET_THREADS)
blur_bin_cols_kernel(const float* __restrict__ img, int nx, int ny, int os, const float* __restrict__ g, int half,
                     float* __restrict__ T1, int out_cols) {
    __shared__ float W[MAX_TAPS];
    build_composite(W, g, half, os);
    const int v = blockIdx.x * DET_THREADS + threadIdx.x;
    const int xp = blockIdx.y;
    if (v >= out_cols) return;
    const int pad = DET_PAD * os, npy = ny + 2 * pad;
    const float* row = img + (size_t)reflect_index(xp - pad, nx) * ny;
    const int taps = os + 2 * half;
    float acc = 0.f;
    for (int t = 0; t < taps; ++t) {
        const int yp = v * os + t - half;
        if (yp >= 0 && yp < npy) acc = fmaf(W[t], __ldg(row + reflect_index(yp - pad, ny)), acc);
    }
    T1[(size_t)xp * out_cols + v] = acc;
}

// Pass B: along rows.  B[u][v] = sum_d W(d) * T1[u*os + d][v], 0 outside.
__global__ void __launch_bounds__(DET_THREADS)
blur_bin_rows_kernel(const float* __restrict__ T1, int npx, int os, const float* __restrict__ g, int half,
                     float* __restrict__ B, int out_cols) {
    __shared__ float W[MAX_TAPS];
    build_composite(W, g, half, os);
    const int v = blockIdx.x * DET_THREADS + threadIdx.x;
    const int u = blockIdx.y;
    if (v >= out_cols) return;
    const int taps = os + 2 * half;
    float acc = 0.f;
    for (int t = 0; t < taps; ++t) {
        const int xp = u * os + t - half;
        if (xp >= 0 && xp < npx) acc = fmaf(W[t], __ldg(T1 + (size_t)xp * out_cols + v), acc);
    }
    B[(size_t)u * out_cols + v] = acc;
}

// PSF pass along columns with crop: C[u][b] = sum_e g(e) B[u][b + 15 + e], 0 outside [0, in_cols).
__global__ void __launch_bounds__(DET_THREADS)
psf_cols_kernel(const float* __restrict__ B, int in_cols, const float* __restrict__ g, int half,
                float* __restrict__ C, int det_y) {
    const int b = blockIdx.x * DET_THREADS + threadIdx.x;
    const int u = blockIdx.y;
    if (b >= det_y) return;
    float acc = 0.f;
    for (int e = -half; e <= half; ++e) {
        const int v = b + DET_PAD + e;
        if (v >= 0 && v < in_cols) acc = fmaf(__ldg(g + e + half), __ldg(B + (size_t)u * in_cols + v), acc);
    }
    C[(size_t)u * det_y + b] = acc;
}

// PSF pass along rows with crop: out[a][b] = sum_e g(e) C[a + 15 + e][b], 0 outside [0, in_rows).
__global__ void __launch_bounds__(DET_THREADS)
psf_rows_kernel(const float* __restrict__ C, int in_rows, const float* __restrict__ g, int half,
                float* __restrict__ out, int det_y) {
    const int b = blockIdx.x * DET_THREADS + threadIdx.x;
    const int a = blockIdx.y;
    if (b >= det_y) return;
    float acc = 0.f;
    for (int e = -half; e <= half; ++e) {
        const int u = a + DET_PAD + e;
        if (u >= 0 && u < in_rows) acc = fmaf(__ldg(g + e + half), __ldg(C + (size_t)u * det_y + b), acc);
    }
    out[(size_t)a * det_y + b] = acc;
}

__global__ void __launch_bounds__(DET_THREADS)
crop_kernel(const float* __restrict__ B, int in_cols, float* __restrict__ out, int det_y) {
    const int b = blockIdx.x * DET_THREADS + threadIdx.x;
    const int a = blockIdx.y;
    if (b < det_y) out[(size_t)a * det_y + b] = B[(size_t)(a + DET_PAD) * in_cols + b + DET_PAD];
}

// Detector.py:185-198 stand-alone: s = nx / size_x, plain block sums.
__global__ void __launch_bounds__(DET_THREADS)
bin_sum_kernel(const float* __restrict__ img, int nx, int ny, int s, float* __restrict__ out, int sx, int sy) {
    const int b = blockIdx.x * DET_THREADS + threadIdx.x;
    const int a = blockIdx.y;
    if (b >= sy) return;
    float acc = 0.f;
    for (int u = 0; u < s; ++u)
        for (int w = 0; w < s; ++w) {
            const int r = a * s + u, c = b * s + w;
            if (r < nx && c < ny) acc += __ldg(img + (size_t)r * ny + c);
        }
    out[(size_t)a * sy + b] = acc;
}

// ---------------------------------------------------------------------------------------------
// Poisson noise: Philox4x32-10 counter-based generator + inversion (lam < 10) / PTRS (lam >= 10)
// ---------------------------------------------------------------------------------------------
struct Philox {
    uint32_t c[4], k[2];
    __device__ __forceinline__ void round() {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    __device__ __forceinline__ void generate(uint32_t out[4]) {
        Philox s = *this;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            s.round();
            s.k[0] += 0x9E3779B9u;
            s.k[1] += 0xBB67AE85u;
        }
        out[0] = s.c[0]; out[1] = s.c[1]; out[2] = s.c[2]; out[3] = s.c[3];
    }
};

__device__ __forceinline__ float u01f(uint32_t x) {
    return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // 24-bit uniform in (0, 1)
}
__device__ __forceinline__ double u01d(uint32_t hi, uint32_t lo) {
    const uint64_t x = ((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11);   // 53-bit uniform in (0, 1)
    return ((double)(x & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

// log(k!) for k < 16
__constant__ float LOG_FACT[16] = {0.f, 0.f, 0.69314718f, 1.79175947f, 3.17805383f, 4.78749174f, 6.57925121f, 8.52516136f,
                                   10.60460290f, 12.80182748f, 15.10441257f, 17.50230785f, 19.98721450f, 22.55216385f,
                                   25.19122118f, 27.89927138f};

// log of the Poisson pmf at k for mean lam >= 10, in fp32 without cancellation:
//   k >= 16: Stirling,  log p = lam * g(x) - log(2 pi k)/2 - 1/(12k) + 1/(360k^3),  x = (k - lam)/lam,
//            g(x) = x - (1+x) log(1+x)  (series for small x: the two O(lam) terms never meet)
//   k <  16: -lam + k log(lam) - log(k!) from the table (all terms are small there).
__device__ __forceinline__ float log_poisson_pmf(float k, float lam) {
    if (k < 16.f) return -lam + k * logf(lam) - LOG_FACT[(int)k];
    const float x = (k - lam) / lam;
    float g;
    if (fabsf(x) < 0.05f) {
        const float x2 = x * x;
        g = x2 * (-0.5f + x * (1.f / 6.f + x * (-1.f / 12.f + x * (0.05f + x * (-1.f / 30.f)))));
    } else {
        g = x - (1.f + x) * log1pf(x);
    }
    const float ik = 1.f / k;
    return lam * g - 0.5f * logf(6.28318530718f * k) - ik * (1.f / 12.f - ik * ik * (1.f / 360.f));
}

// One Poisson variate with mean lam, a pure function of (seed, sequence, pixel).
//   lam < 10 : inversion by sequential search on one uniform;
//   lam >= 10: PTRS -- W. Hoermann, "The transformed rejection method for generating Poisson random
//              variables", Insur. Math. Econ. 12 (1993).  ~86 % of the draws end at the quick
//              acceptance test; the full test uses the cancellation-free fp32 log-pmf above, so a
//              warp never waits on fp64 transcendentals.
__device__ __forceinline__ Philox poisson_stream(uint64_t seed, uint64_t seq, uint64_t pixel) {
    Philox g;
    g.c[0] = (uint32_t)pixel;
    g.c[1] = 0u;
    g.c[2] = (uint32_t)seq;
    g.c[3] = (uint32_t)(seq >> 32) ^ (uint32_t)(pixel >> 32);
    g.k[0] = (uint32_t)seed;
    g.k[1] = (uint32_t)(seed >> 32);
    return g;
}

struct PtrsSetup {
    float b, a, vr;
    __device__ __forceinline__ explicit PtrsSetup(float lam) {
        const float slam = sqrtf(lam);
        b = 0.931f + 2.53f * slam;
        a = -0.059f + 0.02483f * b;
        vr = 0.9277f - 3.6224f / (b - 2.0f);
    }
    // k = floor((2a/us + b) U + lam + 0.43): the sum is formed in fp64 so large means keep unit resolution
    __device__ __forceinline__ float candidate(float lam, float U, float us) const {
        return (float)floor((double)((2.0f * a / us + b) * U) + (double)lam + 0.43);
    }
};

// The cheap part of a draw: everything except PTRS trials that fail the quick acceptance test.
// Returns true and the variate when it is decided here (lam <= 0, lam < 10, or trial 0 accepted).
__device__ __forceinline__ bool poisson_quick(float lam, uint64_t seed, uint64_t seq, uint64_t pixel, float& result) {
    if (!(lam > 0.f)) { result = 0.f; return true; }
    Philox g = poisson_stream(seed, seq, pixel);
    uint32_t r[4];
    g.generate(r);
    if (lam < 10.f) {
        const float u = (float)u01d(r[0], r[1]);
        float p = expf(-lam), F = p;
        int x = 0;
        while (u > F && x < 200) {
            ++x;
            p *= lam / (float)x;
            F += p;
        }
        result = (float)x;
        return true;
    }
    const PtrsSetup t(lam);
    const float U = u01f(r[0]) - 0.5f, V = u01f(r[1]);
    const float us = 0.5f - fabsf(U);
    result = t.candidate(lam, U, us);
    return us >= 0.07f && V <= t.vr;
}

// The full PTRS loop for lam >= 10 (same stream as poisson_quick: trial 0 is replayed first).
__device__ float poisson_ptrs(float lam, uint64_t seed, uint64_t seq, uint64_t pixel) {
    Philox g = poisson_stream(seed, seq, pixel);
    uint32_t r[4];
    const PtrsSetup t(lam);
    const float log_invalpha = logf(1.1239f + 1.1328f / (t.b - 3.4f));
    for (uint32_t trial = 0; trial < 64; ++trial) {
        g.c[1] = trial;
        g.generate(r);
        const float U = u01f(r[0]) - 0.5f, V = u01f(r[1]);
        const float us = 0.5f - fabsf(U);
        const float k = t.candidate(lam, U, us);
        if (us >= 0.07f && V <= t.vr) return k;
        if (k < 0.f || (us < 0.013f && V > us)) continue;
        if (logf(V) + log_invalpha - logf(t.a / (us * us) + t.b) <= log_poisson_pmf(k, lam)) return k;
    }
    return floorf(lam + 0.5f);  // unreachable in practice (acceptance > 0.9 per trial)
}

__device__ float poisson_draw(float lam, uint64_t seed, uint64_t seq, uint64_t pixel) {
    float x;
    if (poisson_quick(lam, seed, seq, pixel, x)) return x;
    return poisson_ptrs(lam, seed, seq, pixel);
}

__global__ void __launch_bounds__(256)
poisson_kernel(const float* __restrict__ expect, float* __restrict__ counts, size_t n, uint64_t seed, uint64_t seq) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        counts[p] = poisson_draw(expect[p], seed, seq, p);
}

// ---------------------------------------------------------------------------------------------
// Fused detector: one kernel from the oversampled image to detector counts.
//
// A block owns a TB x TB tile of detector pixels.  It stages the source window it needs in
// shared memory -- reflect padding resolved by index arithmetic, zeros beyond the padded frame
// exactly like fftconvolve(mode='same') sees them -- then runs blur+bin along columns, along
// rows, the PSF along columns, along rows, and (optionally) the Poisson draw, all on chip.
// Global traffic: the image once (plus the halo re-reads, served by L2) and the counts once.
// ---------------------------------------------------------------------------------------------
constexpr int TB = 32;
constexpr int FUSED_TX = 32, FUSED_TY = 8;
constexpr int FUSED_THREADS = FUSED_TX * FUSED_TY;

struct FusedShape {
    int nx, ny, os, det_x, det_y, src_half, psf_half;
    int bw;      // binned window side   = TB + 2*psf_half
    int sw;      // source window side   = bw*os + 2*src_half
};

// TAPS = os + 2*src_half when it is small (column offsets and weights live in registers);
// TAPS = 0 is the generic version for wider source kernels.
template <bool NOISE, int TAPS>
__global__ void __launch_bounds__(FUSED_THREADS)
detect_fused_kernel(const float* __restrict__ img, FusedShape s, const float* __restrict__ gsrc,
                    const float* __restrict__ gpsf, float* __restrict__ out, uint64_t seed, uint64_t seq) {
    extern __shared__ float smem[];
    float* A = smem;                              // [sw][bw]   blurred+binned along columns
    float* B = A + s.sw * s.bw;                   // [bw][bw]   blurred+binned both ways
    float* C = A;                                 // [bw][TB]   PSF along columns (reuses A)
    float* W = B + s.bw * s.bw;                   // composite kernel, os + 2*src_half taps
    float* P = W + (s.os + 2 * s.src_half);       // PSF kernel, 2*psf_half + 1 taps
    int* rowoff = (int*)(P + 2 * s.psf_half + 1); // [sw] source row offset (elements) or -1 = zero row
    int* queue = rowoff + s.sw;                   // [TB*TB][2] pixels whose Poisson draw needs the full PTRS test
    __shared__ int n_queued;
    if (threadIdx.x == 0 && threadIdx.y == 0) n_queued = 0;
    const int tid = threadIdx.y * FUSED_TX + threadIdx.x;
    const int taps = TAPS > 0 ? TAPS : s.os + 2 * s.src_half;
    const int pad = DET_PAD * s.os, npx = s.nx + 2 * pad, npy = s.ny + 2 * pad;
    const int bxn = s.det_x + 2 * DET_PAD, byn = s.det_y + 2 * DET_PAD;
    // first binned row / column of the window (coordinates of the padded, binned frame)
    const int u0 = blockIdx.y * TB + DET_PAD - s.psf_half, v0 = blockIdx.x * TB + DET_PAD - s.psf_half;
    const int x0 = u0 * s.os - s.src_half, y0 = v0 * s.os - s.src_half;
    for (int t = tid; t < taps; t += FUSED_THREADS) {
        const int d = t - s.src_half;
        float acc = 0.f;
        for (int a = 0; a < s.os; ++a) {
            const int e = d - a;
            if (e >= -s.src_half && e <= s.src_half) acc += gsrc ? gsrc[e + s.src_half] : 1.f;
        }
        W[t] = acc;
    }
    for (int t = tid; t < 2 * s.psf_half + 1; t += FUSED_THREADS) P[t] = gpsf ? gpsf[t] : 1.f;
    for (int a = tid; a < s.sw; a += FUSED_THREADS) {
        const int xp = x0 + a;
        rowoff[a] = (xp >= 0 && xp < npx) ? reflect_index(xp - pad, s.nx) * s.ny : -1;   // nx*ny < 2^30 (checked on the host)
    }
    __syncthreads();

    // blur + bin along columns, straight from global memory (each source pixel is read by <= 2 taps
    // of neighbouring threads: L1 serves the overlap).  A thread keeps one output column: its
    // source columns (reflected, or -1 beyond the padded frame) are fixed for the whole sweep.
    for (int v = threadIdx.x; v < s.bw; v += FUSED_TX) {
        const int yb = y0 + v * s.os;
        if (TAPS > 0) {
            int col[TAPS > 0 ? TAPS : 1];
            float w[TAPS > 0 ? TAPS : 1];
#pragma unroll
            for (int k = 0; k < TAPS; ++k) {
                const int yp = yb + k;
                const bool in = yp >= 0 && yp < npy;
                col[k] = in ? reflect_index(yp - pad, s.ny) : 0;
                w[k] = in ? W[k] : 0.f;
            }
            for (int a = threadIdx.y; a < s.sw; a += FUSED_TY) {
                const int ro = rowoff[a];
                float acc = 0.f;
                if (ro >= 0) {
                    const float* row = img + ro;
#pragma unroll
                    for (int k = 0; k < TAPS; ++k) acc = fmaf(w[k], __ldg(row + col[k]), acc);
                }
                A[a * s.bw + v] = acc;
            }
        } else {
            for (int a = threadIdx.y; a < s.sw; a += FUSED_TY) {
                const int ro = rowoff[a];
                float acc = 0.f;
                if (ro >= 0) {
                    const float* row = img + ro;
                    for (int k = 0; k < taps; ++k) {
                        const int yp = yb + k;
                        if (yp >= 0 && yp < npy) acc = fmaf(W[k], __ldg(row + reflect_index(yp - pad, s.ny)), acc);
                    }
                }
                A[a * s.bw + v] = acc;
            }
        }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < s.bw; v += FUSED_TX) {                    // along rows
        const bool vin = (v0 + v >= 0) & (v0 + v < byn);
        for (int u = threadIdx.y; u < s.bw; u += FUSED_TY) {
            const float* src = A + u * s.os * s.bw + v;
            float acc = 0.f;
            if (TAPS > 0) {
#pragma unroll
                for (int k = 0; k < TAPS; ++k) acc = fmaf(W[k], src[k * s.bw], acc);
            } else {
                for (int k = 0; k < taps; ++k) acc = fmaf(W[k], src[k * s.bw], acc);
            }
            // binned pixels outside the padded frame are the zeros fftconvolve(mode='same') extends with
            const bool in = vin & (u0 + u >= 0) & (u0 + u < bxn);
            B[u * s.bw + v] = in ? acc : 0.f;
        }
    }
    __syncthreads();
    const int np = 2 * s.psf_half + 1;
    for (int u = threadIdx.y; u < s.bw; u += FUSED_TY) {                    // PSF along columns
        const float* src = B + u * s.bw + threadIdx.x;
        float acc = 0.f;
        for (int k = 0; k < np; ++k) acc = fmaf(P[k], src[k], acc);
        C[u * TB + threadIdx.x] = acc;
    }
    __syncthreads();
    const int db = blockIdx.x * TB + threadIdx.x;
    for (int a = threadIdx.y; a < TB; a += FUSED_TY) {                      // PSF along rows, noise, store
        const int da = blockIdx.y * TB + a;
        if (da >= s.det_x || db >= s.det_y) continue;
        const float* src = C + a * TB + threadIdx.x;
        float acc = 0.f;
        for (int k = 0; k < np; ++k) acc = fmaf(P[k], src[k * TB], acc);
        const size_t p = (size_t)da * s.det_y + db;
        if (!NOISE) { out[p] = acc; continue; }
        // ~86 % of the draws finish at the quick acceptance test; the rest are queued so that the
        // expensive tail runs on dense warps instead of dragging every warp through it
        float x;
        if (poisson_quick(acc, seed, seq, p, x)) {
            out[p] = x;
        } else {
            const int slot = atomicAdd(&n_queued, 1);
            queue[2 * slot] = a * TB + threadIdx.x;
            queue[2 * slot + 1] = __float_as_int(acc);
        }
    }
    if (NOISE) {
        __syncthreads();
        for (int qi = tid; qi < n_queued; qi += FUSED_THREADS) {
            const int local = queue[2 * qi];
            const float lam = __int_as_float(queue[2 * qi + 1]);
            const size_t p = (size_t)(blockIdx.y * TB + local / TB) * s.det_y + (blockIdx.x * TB + local % TB);
            out[p] = poisson_ptrs(lam, seed, seq, p);
        }
    }
}

template <bool NOISE, int TAPS>
static int launch_fused(const float* image, const FusedShape& s, const float* gs, const float* gp, float* out,
                        uint64_t seed, uint64_t seq, size_t smem, cudaStream_t st) {
    static size_t configured = 0;
    if (smem > configured) {
        PARESIS_CUDA(cudaFuncSetAttribute(detect_fused_kernel<NOISE, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const dim3 grid(div_up(s.det_y, TB), div_up(s.det_x, TB)), block(FUSED_TX, FUSED_TY);
    detect_fused_kernel<NOISE, TAPS><<<grid, block, smem, st>>>(image, s, gs, gp, out, seed, seq);
    PARESIS_LAUNCH_CHECK("detect_fused_kernel");
    return PARESIS_OK;
}

template <bool NOISE>
static int dispatch_fused(int taps, const float* image, const FusedShape& s, const float* gs, const float* gp, float* out,
                          uint64_t seed, uint64_t seq, size_t smem, cudaStream_t st) {
    switch (taps) {
        case 1: return launch_fused<NOISE, 1>(image, s, gs, gp, out, seed, seq, smem, st);
        case 2: return launch_fused<NOISE, 2>(image, s, gs, gp, out, seed, seq, smem, st);
        case 3: return launch_fused<NOISE, 3>(image, s, gs, gp, out, seed, seq, smem, st);
        case 4: return launch_fused<NOISE, 4>(image, s, gs, gp, out, seed, seq, smem, st);
        case 5: return launch_fused<NOISE, 5>(image, s, gs, gp, out, seed, seq, smem, st);
        case 6: return launch_fused<NOISE, 6>(image, s, gs, gp, out, seed, seq, smem, st);
        case 8: return launch_fused<NOISE, 8>(image, s, gs, gp, out, seed, seq, smem, st);
        default: return launch_fused<NOISE, 0>(image, s, gs, gp, out, seed, seq, smem, st);
    }
}

}  // namespace paresis

using namespace paresis;

extern "C" size_t paresis_detect_work_floats(int nx, int ny, int os, int det_x, int det_y) {
    const size_t npx = (size_t)nx + 2 * DET_PAD * os;
    const size_t bx = det_x + 2 * DET_PAD, by = det_y + 2 * DET_PAD;
    (void)ny;
    return npx * by + bx * by + bx * (size_t)det_y;
}

extern "C" int paresis_detect(const float* image, int nx, int ny, int os, int det_x, int det_y,
                              const float* src_kernel, int src_half, const float* psf_kernel, int psf_half,
                              float* work, float* expect_out, paresis_stream stream) {
    if (!image || !work || !expect_out || os < 1 || det_x < 1 || det_y < 1) {
        set_last_error("paresis_detect: bad arguments");
        return PARESIS_ERR_ARG;
    }
    if (nx != det_x * os || ny != det_y * os) {
        set_last_error("paresis_detect: study grid %dx%d is not detector %dx%d times oversampling %d", nx, ny, det_x, det_y, os);
        return PARESIS_ERR_ARG;
    }
    if (DET_PAD * os > nx - 1 || DET_PAD * os > ny - 1) {
        set_last_error("paresis_detect: reflect margin %d exceeds the image", DET_PAD * os);
        return PARESIS_ERR_ARG;
    }
    if (src_half < 0 || psf_half < 0 || os + 2 * src_half > MAX_TAPS || (src_half > 0 && !src_kernel) || (psf_half > 0 && !psf_kernel)) {
        set_last_error("paresis_detect: kernel half-widths %d / %d unsupported", src_half, psf_half);
        return PARESIS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int pad = DET_PAD * os, npx = nx + 2 * pad;
    const int bx = det_x + 2 * DET_PAD, by = det_y + 2 * DET_PAD;
    float* T1 = work;
    float* B = T1 + (size_t)npx * by;
    float* C = B + (size_t)bx * by;
    const float* g = src_half > 0 ? src_kernel : nullptr;
    blur_bin_cols_kernel<<<dim3(div_up(by, DET_THREADS), npx), DET_THREADS, 0, s>>>(image, nx, ny, os, g, src_half, T1, by);
    PARESIS_LAUNCH_CHECK("blur_bin_cols_kernel");
    blur_bin_rows_kernel<<<dim3(div_up(by, DET_THREADS), bx), DET_THREADS, 0, s>>>(T1, npx, os, g, src_half, B, by);
    PARESIS_LAUNCH_CHECK("blur_bin_rows_kernel");
    if (psf_half > 0) {
        psf_cols_kernel<<<dim3(div_up(det_y, DET_THREADS), bx), DET_THREADS, 0, s>>>(B, by, psf_kernel, psf_half, C, det_y);
        PARESIS_LAUNCH_CHECK("psf_cols_kernel");
        psf_rows_kernel<<<dim3(div_up(det_y, DET_THREADS), det_x), DET_THREADS, 0, s>>>(C, bx, psf_kernel, psf_half, expect_out, det_y);
        PARESIS_LAUNCH_CHECK("psf_rows_kernel");
    } else {
        crop_kernel<<<dim3(div_up(det_y, DET_THREADS), det_x), DET_THREADS, 0, s>>>(B, by, expect_out, det_y);
        PARESIS_LAUNCH_CHECK("crop_kernel");
    }
    return PARESIS_OK;
}

static size_t fused_smem_bytes(const FusedShape& s) {
    const size_t floats = (size_t)s.sw * s.bw + (size_t)s.bw * s.bw + (s.os + 2 * s.src_half) + (2 * s.psf_half + 1) + s.sw
                          + 2 * TB * TB;
    return floats * sizeof(float);
}

extern "C" int paresis_detect_counts(const float* image, int nx, int ny, int os, int det_x, int det_y,
                                     const float* src_kernel, int src_half, const float* psf_kernel, int psf_half,
                                     float* work, float* out, int noise, uint64_t seed, uint64_t sequence,
                                     paresis_stream stream) {
    if (!image || !out || os < 1 || det_x < 1 || det_y < 1 || nx != det_x * os || ny != det_y * os ||
        DET_PAD * os > nx - 1 || DET_PAD * os > ny - 1 || src_half < 0 || psf_half < 0 ||
        (src_half > 0 && !src_kernel) || (psf_half > 0 && !psf_kernel) || (long)nx * ny >= (1L << 30)) {
        set_last_error("paresis_detect_counts: bad arguments");
        return PARESIS_ERR_ARG;
    }
    FusedShape s{nx, ny, os, det_x, det_y, src_half, psf_half, TB + 2 * psf_half, (TB + 2 * psf_half) * os + 2 * src_half};
    const size_t smem = fused_smem_bytes(s);
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 200 * 1024) {
        // very wide kernels: separable passes through global memory, then the draw in place
        if (!work) { set_last_error("paresis_detect_counts: this kernel size needs the work buffer"); return PARESIS_ERR_ARG; }
        int rc = paresis_detect(image, nx, ny, os, det_x, det_y, src_kernel, src_half, psf_kernel, psf_half, work, out, stream);
        if (rc != PARESIS_OK || !noise) return rc;
        return paresis_poisson(out, out, (size_t)det_x * det_y, seed, sequence, stream);
    }
    const float* gs = src_half > 0 ? src_kernel : nullptr;
    const float* gp = psf_half > 0 ? psf_kernel : nullptr;
    const int taps = os + 2 * src_half;
    return noise ? dispatch_fused<true>(taps, image, s, gs, gp, out, seed, sequence, smem, st)
                 : dispatch_fused<false>(taps, image, s, gs, gp, out, seed, sequence, smem, st);
}

extern "C" int paresis_poisson(const float* expect, float* counts, size_t n, uint64_t seed, uint64_t sequence,
                               paresis_stream stream) {
    if (!expect || !counts) { set_last_error("paresis_poisson: null pointer"); return PARESIS_ERR_ARG; }
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    poisson_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(expect, counts, n, seed, sequence);
    PARESIS_LAUNCH_CHECK("poisson_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_bin_sum(const float* image, int nx, int ny, int size_x, int size_y, float* out,
                               paresis_stream stream) {
    if (!image || !out || size_x < 1 || size_y < 1 || nx < size_x) {
        set_last_error("paresis_bin_sum: bad arguments");
        return PARESIS_ERR_ARG;
    }
    const int sfac = nx / size_x;  // Detector.py:192
    bin_sum_kernel<<<dim3(div_up(size_y, DET_THREADS), size_x), DET_THREADS, 0, (cudaStream_t)stream>>>(
        image, nx, ny, sfac, out, size_x, size_y);
    PARESIS_LAUNCH_CHECK("bin_sum_kernel");
    return PARESIS_OK;
}
