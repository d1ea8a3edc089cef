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

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    // 53-bit uniform in (0, 1)
    const uint64_t x = ((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11);
    return ((double)(x & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ double poisson_draw(double lam, Philox& g) {
    uint32_t r[4];
    if (!(lam > 0.0)) return 0.0;
    if (lam < 10.0) {
        // inversion by sequential search on one uniform
        g.generate(r);
        const double u = u01(r[0], r[1]);
        double p = exp(-lam), F = p;
        int x = 0;
        while (u > F && x < 200) {
            ++x;
            p *= lam / x;
            F += p;
        }
        return (double)x;
    }
    // PTRS, W. Hoermann, "The transformed rejection method for generating Poisson random variables" (1993)
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (uint32_t trial = 0; trial < 64; ++trial) {
        g.c[1] = trial;
        g.generate(r);
        const double U = u01(r[0], r[1]) - 0.5, V = u01(r[2], r[3]);
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lam + 0.5);  // unreachable in practice (acceptance > 0.9 per trial)
}

__global__ void __launch_bounds__(256)
poisson_kernel(const float* __restrict__ expect, float* __restrict__ counts, size_t n, uint64_t seed, uint64_t seq) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        Philox g;
        g.c[0] = (uint32_t)p;
        g.c[1] = 0u;
        g.c[2] = (uint32_t)seq;
        g.c[3] = (uint32_t)(seq >> 32) ^ (uint32_t)(p >> 32);
        g.k[0] = (uint32_t)seed;
        g.k[1] = (uint32_t)(seed >> 32);
        counts[p] = (float)poisson_draw((double)expect[p], g);
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
