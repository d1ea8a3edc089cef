// Fresnel propagator: reflect-pad -> FFT -> separable transfer function -> inverse FFT -> crop.
//
// Reference: Experiment.py:219-252 (wavePropagation).  The transform size must stay
// (N + 30)^2 -- the reflect margin and the periodic wrap are part of the reference's answer --
// so the FFT itself is cuFFT (library GEMM-free work; sizes like 2078 = 2 x 1039 run Bluestein);
// the pad, the transfer-function multiply and the crop / |.|^2 accumulation around it are
// hand-written and fused so the padded field is touched once per stage.
#include <cufft.h>

#include "common.cuh"

struct paresis_fresnel_plan {
    cufftHandle fft;
    int nx, ny, margin, nxp, nyp;
    float2* buf;
    float2* spec;     // spectrum of the last paresis_fresnel_spectrum() input (allocated on first use)
    size_t work_bytes;
};

namespace paresis {

__device__ __forceinline__ int reflect_idx(int q, int n) {
    if (q < 0) q = -q;
    if (q >= n) q = 2 * (n - 1) - q;
    return q;
}

// np.pad(wave, 15, mode='reflect')  (Experiment.py:236-237)
__global__ void __launch_bounds__(256)
pad_reflect_kernel(const float2* __restrict__ in, int nx, int ny, int m, float2* __restrict__ buf, int nyp) {
    const int yp = blockIdx.x * blockDim.x + threadIdx.x;
    const int xp = blockIdx.y;
    if (yp >= nyp) return;
    buf[(size_t)xp * nyp + yp] = in[(size_t)reflect_idx(xp - m, nx) * ny + reflect_idx(yp - m, ny)];
}

// exp(-i z (u^2+v^2)/(2kM)) * spectrum, with fftshift/ifftshift folded into the vectors (:250)
__global__ void __launch_bounds__(256)
transfer_kernel(float2* __restrict__ buf, const float2* __restrict__ hx, const float2* __restrict__ hy, int nyp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (b >= nyp) return;
    const float2 p = hx[a], q = hy[b];
    const float2 h = make_float2(p.x * q.x - p.y * q.y, p.x * q.y + p.y * q.x);
    float2 w = buf[(size_t)a * nyp + b];
    buf[(size_t)a * nyp + b] = make_float2(w.x * h.x - w.y * h.y, w.x * h.y + w.y * h.x);
}

// the same, out of place: spectrum (kept) -> work buffer
__global__ void __launch_bounds__(256)
transfer_from_kernel(const float2* __restrict__ spec, float2* __restrict__ buf, const float2* __restrict__ hx, const float2* __restrict__ hy,
                     int nyp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (b >= nyp) return;
    const float2 p = hx[a], q = hy[b];
    const float2 h = make_float2(p.x * q.x - p.y * q.y, p.x * q.y + p.y * q.x);
    const float2 w = spec[(size_t)a * nyp + b];
    buf[(size_t)a * nyp + b] = make_float2(w.x * h.x - w.y * h.y, w.x * h.y + w.y * h.x);
}

// crop (:251), global phase (:250) and, optionally, |.|^2 accumulation (:351-358)
__global__ void __launch_bounds__(256)
crop_phase_kernel(const float2* __restrict__ buf, int nyp, int m, float2 phase, float2* __restrict__ out,
            float* __restrict__ acc, int ny) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= ny) return;
    const float2 w = buf[(size_t)(i + m) * nyp + (j + m)];
    const float2 r = make_float2(w.x * phase.x - w.y * phase.y, w.x * phase.y + w.y * phase.x);
    const size_t p = (size_t)i * ny + j;
    if (out) out[p] = r;
    if (acc) acc[p] += r.x * r.x + r.y * r.y;
}

}  // namespace paresis

using namespace paresis;

extern "C" int paresis_fresnel_plan_create(int nx, int ny, int margin, paresis_fresnel_plan** plan) {
    if (!plan || nx < 2 || ny < 2 || margin < 0 || margin > nx - 1 || margin > ny - 1) {
        set_last_error("paresis_fresnel_plan_create: bad arguments");
        return PARESIS_ERR_ARG;
    }
    paresis_fresnel_plan* p = new paresis_fresnel_plan();
    p->nx = nx; p->ny = ny; p->margin = margin;
    p->nxp = nx + 2 * margin; p->nyp = ny + 2 * margin;
    p->buf = nullptr; p->spec = nullptr; p->work_bytes = 0;
    cufftResult r = cufftCreate(&p->fft);
    if (r == CUFFT_SUCCESS) r = cufftMakePlan2d(p->fft, p->nxp, p->nyp, CUFFT_C2C, &p->work_bytes);
    if (r != CUFFT_SUCCESS) {
        set_last_error("cuFFT plan %d x %d failed (code %d)", p->nxp, p->nyp, (int)r);
        delete p;
        return PARESIS_ERR_CUFFT;
    }
    cudaError_t e = cudaMalloc(&p->buf, sizeof(float2) * (size_t)p->nxp * p->nyp);
    if (e != cudaSuccess) {
        cufftDestroy(p->fft);
        delete p;
        return check_cuda(e, "cudaMalloc(fresnel buffer)");
    }
    *plan = p;
    return PARESIS_OK;
}

extern "C" int paresis_fresnel_plan_destroy(paresis_fresnel_plan* p) {
    if (!p) return PARESIS_OK;
    cufftDestroy(p->fft);
    cudaFree(p->buf);
    if (p->spec) cudaFree(p->spec);
    delete p;
    return PARESIS_OK;
}

extern "C" size_t paresis_fresnel_plan_bytes(const paresis_fresnel_plan* p) {
    return p ? p->work_bytes + sizeof(float2) * (size_t)p->nxp * p->nyp : 0;
}

extern "C" int paresis_fresnel_propagate(paresis_fresnel_plan* p, const paresis_c32* wave_in,
                                         const paresis_c32* hx, const paresis_c32* hy, paresis_c32 phase,
                                         paresis_c32* wave_out, float* intensity_acc, paresis_stream stream) {
    if (!p || !wave_in || !hx || !hy || (!wave_out && !intensity_acc)) {
        set_last_error("paresis_fresnel_propagate: null pointer");
        return PARESIS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 gp((p->nyp + 255) / 256, p->nxp), gc((p->ny + 255) / 256, p->nx);
    pad_reflect_kernel<<<gp, 256, 0, s>>>((const float2*)wave_in, p->nx, p->ny, p->margin, p->buf, p->nyp);
    PARESIS_LAUNCH_CHECK("pad_reflect_kernel");
    cufftResult r = cufftSetStream(p->fft, s);
    if (r == CUFFT_SUCCESS) r = cufftExecC2C(p->fft, p->buf, p->buf, CUFFT_FORWARD);
    if (r != CUFFT_SUCCESS) { set_last_error("cufftExecC2C forward failed (code %d)", (int)r); return PARESIS_ERR_CUFFT; }
    transfer_kernel<<<gp, 256, 0, s>>>(p->buf, (const float2*)hx, (const float2*)hy, p->nyp);
    PARESIS_LAUNCH_CHECK("transfer_kernel");
    r = cufftExecC2C(p->fft, p->buf, p->buf, CUFFT_INVERSE);
    if (r != CUFFT_SUCCESS) { set_last_error("cufftExecC2C inverse failed (code %d)", (int)r); return PARESIS_ERR_CUFFT; }
    crop_phase_kernel<<<gc, 256, 0, s>>>(p->buf, p->nyp, p->margin, make_float2(phase.re, phase.im), (float2*)wave_out,
                                   intensity_acc, p->ny);
    PARESIS_LAUNCH_CHECK("crop_phase_kernel");
    return PARESIS_OK;
}

// Two propagations of the SAME field over different distances (Experiment.py:340-341 and :349 both start from the wave
// behind the membrane) share their reflect-pad and forward FFT: paresis_fresnel_spectrum() keeps fft2(pad(wave)) in the
// plan, paresis_fresnel_from_spectrum() applies a transfer function to it (out of place), transforms back and crops.
// On the Bluestein sizes this path is made of (4126 = 2 x 2063, 8222 = 2 x 4111) a 2-D transform is ~45 % of a
// propagation, so a position with three propagations saves one transform in six.
extern "C" int paresis_fresnel_spectrum(paresis_fresnel_plan* p, const paresis_c32* wave_in, paresis_stream stream) {
    if (!p || !wave_in) { set_last_error("paresis_fresnel_spectrum: null pointer"); return PARESIS_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (!p->spec) PARESIS_CUDA(cudaMalloc(&p->spec, sizeof(float2) * (size_t)p->nxp * p->nyp));
    const dim3 gp((p->nyp + 255) / 256, p->nxp);
    pad_reflect_kernel<<<gp, 256, 0, s>>>((const float2*)wave_in, p->nx, p->ny, p->margin, p->spec, p->nyp);
    PARESIS_LAUNCH_CHECK("pad_reflect_kernel");
    cufftResult r = cufftSetStream(p->fft, s);
    if (r == CUFFT_SUCCESS) r = cufftExecC2C(p->fft, p->spec, p->spec, CUFFT_FORWARD);
    if (r != CUFFT_SUCCESS) { set_last_error("cufftExecC2C forward failed (code %d)", (int)r); return PARESIS_ERR_CUFFT; }
    return PARESIS_OK;
}

extern "C" int paresis_fresnel_from_spectrum(paresis_fresnel_plan* p, const paresis_c32* hx, const paresis_c32* hy, paresis_c32 phase,
                                             paresis_c32* wave_out, float* intensity_acc, paresis_stream stream) {
    if (!p || !p->spec || !hx || !hy || (!wave_out && !intensity_acc)) {
        set_last_error("paresis_fresnel_from_spectrum: null pointer, or no spectrum stored (call paresis_fresnel_spectrum first)");
        return PARESIS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 gp((p->nyp + 255) / 256, p->nxp), gc((p->ny + 255) / 256, p->nx);
    transfer_from_kernel<<<gp, 256, 0, s>>>(p->spec, p->buf, (const float2*)hx, (const float2*)hy, p->nyp);
    PARESIS_LAUNCH_CHECK("transfer_from_kernel");
    cufftResult r = cufftSetStream(p->fft, s);
    if (r == CUFFT_SUCCESS) r = cufftExecC2C(p->fft, p->buf, p->buf, CUFFT_INVERSE);
    if (r != CUFFT_SUCCESS) { set_last_error("cufftExecC2C inverse failed (code %d)", (int)r); return PARESIS_ERR_CUFFT; }
    crop_phase_kernel<<<gc, 256, 0, s>>>(p->buf, p->nyp, p->margin, make_float2(phase.re, phase.im), (float2*)wave_out, intensity_acc, p->ny);
    PARESIS_LAUNCH_CHECK("crop_phase_kernel");
    return PARESIS_OK;
}
