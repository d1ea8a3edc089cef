// Argument block and ray clean-up shared by the fused refraction kernels (refraction.cu, refract_lean.cu).
#pragma once
#include "splat.cuh"

namespace paresis {

template <typename T>
struct RefractArgs {
    const T* map[PARESIS_MAX_LAYERS];
    float g_obj[PARESIS_MAX_LAYERS];
    float g_ref[PARESIS_MAX_LAYERS];
    float att[PARESIS_MAX_LAYERS];
    const float* I_in;
    float I_uniform;
    float* out_obj;
    float* out_ref;
    float* dx_pad;   // optional: cleaned object-beam displacement, stored at (+margin, +margin)
    float* dy_pad;
    Frame f;
    float clamp_x, clamp_y;   // rays with |D| above these are dropped (v2: nx, ny; v1: 1e3)
    int rows;
    int* flag;
    // pipeline extras (all optional): buffers the NEXT kernel accumulates into are zero-filled here, pixel
    // by pixel; the input intensity is cleared once read (so the next energy can scatter into it again);
    // the sum of everything the reference beam deposits inside the image is added to *sum_ref
    // (= N * mean of the reference image of this energy, Experiment.py:485-486).
    float* zero[3];
    bool clear_input;
    double* sum_ref;
    double* zero_scalar;
    float intensity_scale;   // > 0: nominal beam intensity, enables the fixed-point tile kernels
    bool full_tiles;         // tile hops: blocks of the full tile height instead of the wave-filling row count
};

// The tile hop for several membrane positions in one launch (blockIdx.z): what differs between the positions.
constexpr int LEAN_MAX_BATCH = 8;
struct LeanItem {
    const float* map[PARESIS_MAX_LAYERS];
    const float* I_in;
    float* out_obj;
    float* out_ref;
    float* zero[3];
    double* sum_ref;
    double* zero_scalar;
};
struct LeanArgs : RefractArgs<float> {
    LeanItem z[LEAN_MAX_BATCH];
};

// refractionFileNumba2.py:59-64: |D| < 1e-12 -> 0; |D| > N kills the ray (I = 0, D = 0).
__device__ __forceinline__ void clean(float& v, float& dx, float& dy, float cx, float cy) {
    if (fabsf(dx) < 1e-12f) dx = 0.f;
    if (fabsf(dy) < 1e-12f) dy = 0.f;
    const bool bx = fabsf(dx) > cx, by = fabsf(dy) > cy;
    if (bx | by) v = 0.f;
    if (bx) dx = 0.f;
    if (by) dy = 0.f;
}

// Fixed-point shared-memory tile kernel (refract_lean.cu): tiles of up to 16 source rows x 256 columns, halo 4, integer
// bilinear split, deferred misses.
int dispatch_refract_lean(int n_layers, const RefractArgs<float>& a, cudaStream_t s);
// ... for n_batch positions at once: a.z[0 .. n_batch) hold their images, the RefractArgs part the shared coefficients
// (its own map / image pointers are ignored, only out_ref != nullptr and I_in != nullptr select the kernel shape).
extern int g_lean_rows_override;      // paresis_set_tuning(3, rows): rows per block of the tile hops, 0 = pick_tile_rows
int dispatch_refract_lean_batch(int n_layers, const LeanArgs& a, int n_batch, cudaStream_t s);

// Stand-alone splat through fixed-point shared-memory tiles (splat_tile.cu): variant 3 of paresis_splat.
int launch_splat_tile(const float* I, const float* Dx, const float* Dy, float* out, const Frame& f, int* flag, cudaStream_t s);

// Stand-alone splat as an owner-computes rolling-strip kernel (splat_strip.cu): variants 4 (out = ...) and 5 (out += ...).
int launch_splat_strip(const float* I, const float* Dx, const float* Dy, float* out, const Frame& f, int* flag, bool accumulate,
                       cudaStream_t s);

}  // namespace paresis
