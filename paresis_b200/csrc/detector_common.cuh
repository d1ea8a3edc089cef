// Constants and index helpers shared by the detector translation units (Detector.py:79-119).
#pragma once
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

constexpr int DT_MAX_IMAGES = 8;

struct DetImages {
    const float* img[DT_MAX_IMAGES];
    float* out[DT_MAX_IMAGES];
    uint64_t seq[DT_MAX_IMAGES];
};

// Tile kernels are compiled per oversampling factor (detector_tile_os*.cu).  Each returns PARESIS_OK
// after launching, an error code, or -1 when (src_half, psf_half) has no compiled kernel.
#define PARESIS_DT_ARGS const DetImages& im, int n_images, int nx, int ny, int det_x, int det_y, const float* gs, \
                        const float* gp, int noise, uint64_t seed, cudaStream_t st
int dispatch_detect_tile_os1(int hs, int hp, PARESIS_DT_ARGS);
int dispatch_detect_tile_os2(int hs, int hp, PARESIS_DT_ARGS);
int dispatch_detect_tile_os3(int hs, int hp, PARESIS_DT_ARGS);
int dispatch_detect_tile_os4(int hs, int hp, PARESIS_DT_ARGS);

}  // namespace paresis
