// Shared helpers for the paresis_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/paresis_b200.h"

namespace paresis {

// Status flag bits written by kernels into a device int (checked lazily by the host shim).
enum : int {
    FLAG_NONFINITE = 1,  // a NaN/Inf intensity was deposited (refractionFileNumba2.py:81-82 guard)
};

void set_last_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return PARESIS_OK;
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return PARESIS_ERR_CUDA;
}

#define PARESIS_CUDA(call)                                         \
    do {                                                           \
        int _rc = ::paresis::check_cuda((call), #call);            \
        if (_rc != PARESIS_OK) return _rc;                         \
    } while (0)

#define PARESIS_LAUNCH_CHECK(name) PARESIS_CUDA(cudaPeekAtLastError())

// Index checks of the debug build (python -m paresis_b200.build --bounds-check -> libparesis_b200_checked.so, selected with
// PARESIS_B200_LIB): a device assert traps on the first index outside [0, n).  Nothing in the production library.
#ifdef PARESIS_BOUNDS_CHECK
#include <assert.h>
#define PARESIS_BOUND(i, n) assert((long long)(i) >= 0 && (long long)(i) < (long long)(n))
#else
#define PARESIS_BOUND(i, n) ((void)0)
#endif

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Streaming loads: every map on this path is read once per kernel, keep it out of L1.
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

// Fire-and-forget float add in L2 (SASS: REDG.E.ADD.F32).
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }

}  // namespace paresis
