// Library bookkeeping (version, error text) and the element-wise kernels of the path:
// stand-alone transmission (Sample.py:248-351), fill / axpy / mean used by the host shim.
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

namespace paresis {

static thread_local char g_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

constexpr int EW_THREADS = 256;

static inline int ew_blocks(size_t n, int per_thread = 4) {
    size_t b = (n + (size_t)EW_THREADS * per_thread - 1) / ((size_t)EW_THREADS * per_thread);
    const size_t cap = 148 * 16;  // grid-stride beyond ~2 waves of 8 blocks/SM
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

struct LayerPtrs {
    const float* t[PARESIS_MAX_LAYERS];
    double att[PARESIS_MAX_LAYERS];
    double phase[PARESIS_MAX_LAYERS];
    int n;
};

// Sample.py:347-348 -- I *= exp(-2 k beta t), phi -= k delta t, material by material.
__global__ void __launch_bounds__(EW_THREADS)
transmit_rt_kernel(const float* __restrict__ I_in, const double* __restrict__ phi_in, LayerPtrs L,
                   float* __restrict__ I_out, double* __restrict__ phi_out, size_t n) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        double arg = 0.0, ph = phi_in ? phi_in[p] : 0.0;
        for (int m = 0; m < L.n; ++m) {
            const double t = (double)ld_stream(L.t[m] + p);
            arg += L.att[m] * t;
            ph -= L.phase[m] * t;
        }
        if (I_out) I_out[p] = (float)((double)I_in[p] * exp(-arg));
        if (phi_out) phi_out[p] = ph;
    }
}

// Sample.py:279 -- wave *= exp((-i k delta - k beta) t).  The phase reaches 1e2-1e3 rad, so it
// is accumulated and reduced in fp64 before the fp32 sincos.
__global__ void __launch_bounds__(EW_THREADS)
transmit_wave_kernel(const float2* __restrict__ w_in, float amp, LayerPtrs L, float2* __restrict__ w_out, size_t n) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        double arg = 0.0, ph = 0.0;
        for (int m = 0; m < L.n; ++m) {
            const double t = (double)ld_stream(L.t[m] + p);
            arg += L.att[m] * t;
            ph -= L.phase[m] * t;
        }
        ph -= 6.283185307179586476925 * rint(ph * 0.15915494309189533577);
        float sn, cs;
        sincosf((float)ph, &sn, &cs);
        const float mag = expf((float)-arg);
        float2 w = w_in ? w_in[p] : make_float2(amp, 0.f);
        w_out[p] = make_float2(mag * (w.x * cs - w.y * sn), mag * (w.x * sn + w.y * cs));
    }
}

__global__ void __launch_bounds__(EW_THREADS) fill_kernel(float* dst, float v, size_t n) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) dst[p] = v;
}

__global__ void __launch_bounds__(EW_THREADS) axpy_kernel(float* dst, const float* __restrict__ src, float a, size_t n) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        dst[p] = fmaf(a, src[p], dst[p]);
}

__global__ void __launch_bounds__(EW_THREADS) sum_kernel(const float* __restrict__ src, size_t n, double scale, double* out) {
    // float4 loads (cudaMalloc'd images are 16-byte aligned), fp32 partial sums of 4, fp64 across iterations
    double acc = 0.0;
    const size_t n4 = (((uintptr_t)src & 15) == 0) ? n / 4 : 0;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n4; p += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src4 + p);
        acc += (double)((v.x + v.y) + (v.z + v.w));
    }
    for (size_t p = 4 * n4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        acc += (double)src[p];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __shared__ double part[EW_THREADS / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < EW_THREADS / 32; ++w) s += part[w];
        atomicAdd(out, s * scale);
    }
}

static int fill_layers(LayerPtrs& L, const float* const* t, const double* att, const double* phase, int n, const char* who) {
    if (n < 0 || n > PARESIS_MAX_LAYERS || (n > 0 && (!t || !att || !phase))) {
        set_last_error("%s: need 0..%d layers", who, PARESIS_MAX_LAYERS);
        return PARESIS_ERR_ARG;
    }
    L.n = n;
    for (int m = 0; m < n; ++m) {
        if (!t[m]) { set_last_error("%s: null thickness map %d", who, m); return PARESIS_ERR_ARG; }
        L.t[m] = t[m];
        L.att[m] = att[m];
        L.phase[m] = phase[m];
    }
    return PARESIS_OK;
}

}  // namespace paresis

using namespace paresis;

extern "C" int paresis_version(void) { return 100; }
extern "C" const char* paresis_last_error(void) { return g_error; }

extern "C" int paresis_transmit_rt(const float* intensity_in, const double* phi_in,
                                   const float* const* thickness_host, const double* atten_host,
                                   const double* phase_host, int n_layers,
                                   float* intensity_out, double* phi_out, size_t n, paresis_stream stream) {
    if ((intensity_out && !intensity_in) || (!intensity_out && !phi_out)) {
        set_last_error("paresis_transmit_rt: nothing to do / missing input");
        return PARESIS_ERR_ARG;
    }
    LayerPtrs L{};
    int rc = fill_layers(L, thickness_host, atten_host, phase_host, n_layers, "paresis_transmit_rt");
    if (rc) return rc;
    transmit_rt_kernel<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>(intensity_in, phi_in, L, intensity_out, phi_out, n);
    PARESIS_LAUNCH_CHECK("transmit_rt_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_transmit_wave(const paresis_c32* wave_in, float amplitude_uniform,
                                     const float* const* thickness_host, const double* atten_host,
                                     const double* phase_host, int n_layers,
                                     paresis_c32* wave_out, size_t n, paresis_stream stream) {
    if (!wave_out) { set_last_error("paresis_transmit_wave: null output"); return PARESIS_ERR_ARG; }
    LayerPtrs L{};
    int rc = fill_layers(L, thickness_host, atten_host, phase_host, n_layers, "paresis_transmit_wave");
    if (rc) return rc;
    transmit_wave_kernel<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>(
        (const float2*)wave_in, amplitude_uniform, L, (float2*)wave_out, n);
    PARESIS_LAUNCH_CHECK("transmit_wave_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_fill(float* dst, float value, size_t n, paresis_stream stream) {
    if (!dst) { set_last_error("paresis_fill: null pointer"); return PARESIS_ERR_ARG; }
    fill_kernel<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>(dst, value, n);
    PARESIS_LAUNCH_CHECK("fill_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_axpy(float* dst, const float* src, float scale, size_t n, paresis_stream stream) {
    if (!dst || !src) { set_last_error("paresis_axpy: null pointer"); return PARESIS_ERR_ARG; }
    axpy_kernel<<<ew_blocks(n), EW_THREADS, 0, (cudaStream_t)stream>>>(dst, src, scale, n);
    PARESIS_LAUNCH_CHECK("axpy_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_sum_scaled(const float* src, size_t n, double scale, double* out, paresis_stream stream) {
    if (!src || !out || n == 0) { set_last_error("paresis_sum_scaled: bad arguments"); return PARESIS_ERR_ARG; }
    sum_kernel<<<ew_blocks(n, 16), EW_THREADS, 0, (cudaStream_t)stream>>>(src, n, scale, out);
    PARESIS_LAUNCH_CHECK("sum_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_mean(const float* src, size_t n, double* out, paresis_stream stream) {
    if (!src || !out || n == 0) { set_last_error("paresis_mean: bad arguments"); return PARESIS_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    PARESIS_CUDA(cudaMemsetAsync(out, 0, sizeof(double), s));
    sum_kernel<<<ew_blocks(n, 16), EW_THREADS, 0, s>>>(src, n, 1.0 / (double)n, out);
    PARESIS_LAUNCH_CHECK("sum_kernel");
    return PARESIS_OK;
}
