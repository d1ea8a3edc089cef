// Dark-field branch of the ray-tracing model.
//
// Reference: refractionFileNumba2.py:88-196 (fastRefractionDF) and Sample.py:322-343 (the Lung /
// 'cylinder_beeds' scattering-angle model).  The refraction part of fastRefractionDF is the ordinary splat
// run twice, on the rays with and without a scattering angle (:143-150); what is new is the
// per-pixel VARIABLE-WIDTH Gaussian scatter of the refracted dark-field intensity (:171-184, a Python
// double loop upstream), and the bookkeeping around it.  All three pieces are element-wise / small-stencil
// fp32 work on HBM-resident images.
#include <math.h>

#include "common.cuh"

namespace paresis {

constexpr int DF_THREADS = 256;

// Sample.py:328-332: newDf = 2 delta sqrt(Nsphere) sqrt(ln(2/delta) + 1), Nsphere = NsphereVol^(1/3) * t[um];
// times propagationDistance / (pixel * M) (refractionFileNumba2.py:114).  All scalars are folded into `coeff`.
__global__ void __launch_bounds__(DF_THREADS)
df_angle_kernel(const float* __restrict__ thickness, float coeff, float* __restrict__ df_px, size_t n) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const float t = thickness[p];
        df_px[p] = t > 0.f ? coeff * sqrtf(t) : 0.f;
    }
}

// refractionFileNumba2.py:130, :143-146: angles above `limit` pixels are dropped; rays are split by angle == 0.
__global__ void __launch_bounds__(DF_THREADS)
df_split_kernel(const float* __restrict__ intensity, float uniform, const float* __restrict__ df_px, float limit,
                float* __restrict__ i_plain, float* __restrict__ i_df, float* __restrict__ df_clean, size_t n) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        float d = df_px[p];
        if (d > limit) d = 0.f;
        const float v = intensity ? intensity[p] : uniform;
        i_plain[p] = d != 0.f ? 0.f : v;
        i_df[p] = d == 0.f ? 0.f : v;
        df_clean[p] = d;
    }
}

__device__ __forceinline__ int round_half_even(float x) { return __float2int_rn(x); }   // Python's round()

// refractionFileNumba2.py:171-184: every pixel of the refracted dark-field image spreads its value over a
// normalised Gaussian patch of sigma = df/2 (gaussian_shape, :14-23: half-width round(3 sigma), separable);
// df == 0 leaves the value in place.  One warp per source pixel row segment: a lane owns one source pixel and
// walks its patch; patches of neighbouring lanes overlap, so the REDs of a warp are unit-stride.
__global__ void __launch_bounds__(DF_THREADS)
df_scatter_kernel(const float* __restrict__ scattered, const float* __restrict__ df_px, float* __restrict__ out, int nx, int ny) {
    const int j = blockIdx.x * DF_THREADS + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= ny) return;
    const size_t p = (size_t)i * ny + j;
    const float v = scattered[p];
    if (v == 0.f) return;                                    // :173
    const float d = df_px[p];
    if (d == 0.f) { red_add(out + p, v); return; }           // :185-186
    const float sigma = 0.5f * d;                            // :176
    const int half = round_half_even(3.f * sigma);
    if (half == 0) { red_add(out + p, v); return; }          // 1 x 1 patch
    const float inv2s2 = 0.5f / (sigma * sigma);
    float norm = 0.f;
    for (int a = -half; a <= half; ++a) norm += expf(-(float)(a * a) * inv2s2);
    const float scale = v / (norm * norm);                   // g / sum(g), separable
    for (int a = -half; a <= half; ++a) {
        const int r = i + a;
        if ((unsigned)r >= (unsigned)nx) continue;           // beyond the frame: cropped upstream (:189)
        const float wa = scale * expf(-(float)(a * a) * inv2s2);
        float* row = out + (size_t)r * ny;
        for (int b = -half; b <= half; ++b) {
            const int c = j + b;
            if ((unsigned)c < (unsigned)ny) red_add(row + c, wa * expf(-(float)(b * b) * inv2s2));
        }
    }
}

static inline int df_blocks(size_t n) {
    size_t b = (n + DF_THREADS * 4 - 1) / (DF_THREADS * 4);
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace paresis

using namespace paresis;

extern "C" int paresis_df_angle(const float* thickness, double coeff, float* df_px, size_t n, paresis_stream stream) {
    if (!thickness || !df_px) { set_last_error("paresis_df_angle: null pointer"); return PARESIS_ERR_ARG; }
    df_angle_kernel<<<df_blocks(n), DF_THREADS, 0, (cudaStream_t)stream>>>(thickness, (float)coeff, df_px, n);
    PARESIS_LAUNCH_CHECK("df_angle_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_df_split(const float* intensity, float intensity_uniform, const float* df_px, float limit_px,
                                float* i_plain, float* i_df, float* df_clean, size_t n, paresis_stream stream) {
    if (!df_px || !i_plain || !i_df || !df_clean) { set_last_error("paresis_df_split: null pointer"); return PARESIS_ERR_ARG; }
    df_split_kernel<<<df_blocks(n), DF_THREADS, 0, (cudaStream_t)stream>>>(intensity, intensity_uniform, df_px, limit_px, i_plain,
                                                                           i_df, df_clean, n);
    PARESIS_LAUNCH_CHECK("df_split_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_df_scatter(const float* scattered, const float* df_px, float* out, int nx, int ny, paresis_stream stream) {
    if (!scattered || !df_px || !out || nx < 1 || ny < 1) { set_last_error("paresis_df_scatter: bad arguments"); return PARESIS_ERR_ARG; }
    df_scatter_kernel<<<dim3(div_up(ny, DF_THREADS), nx), DF_THREADS, 0, (cudaStream_t)stream>>>(scattered, df_px, out, nx, ny);
    PARESIS_LAUNCH_CHECK("df_scatter_kernel");
    return PARESIS_OK;
}
