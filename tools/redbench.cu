// Micro-benchmark: how fast can sm_100a add into global memory?  (design input for the splat flush)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/redbench.bin tools/redbench.cu
// Prints one JSON line per (kernel, size).  GB/s = 4 bytes per output element / time.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void red1(float* p, float a) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory"); }
__device__ __forceinline__ void red2(float* p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory"); }

// every kernel: grid-stride over n floats, each element touched exactly once
__global__ void k_store4(float* out, size_t n, float v) {
    size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4, st = (size_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += st) *reinterpret_cast<float4*>(out + i) = make_float4(v, v, v, v);
}
__global__ void k_rmw4(float* out, size_t n, float v) {
    size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4, st = (size_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += st) { float4 x = *reinterpret_cast<float4*>(out + i); x.x += v; x.y += v; x.z += v; x.w += v; *reinterpret_cast<float4*>(out + i) = x; }
}
__global__ void k_red1(float* out, size_t n, float v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) red1(out + i, v);
}
__global__ void k_red2(float* out, size_t n, float v) {
    size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2, st = (size_t)gridDim.x * blockDim.x * 2;
    for (; i < n; i += st) red2(out + i, v, v);
}
__global__ void k_red4(float* out, size_t n, float v) {
    size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4, st = (size_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += st) red4(out + i, v, v, v, v);
}
// only 8 of 32 lanes issue a v4 RED covering the warp's 32 floats (the shape a shuffle-gathered flush has)
__global__ void k_red4_8lanes(float* out, size_t n, float v) {
    const int lane = threadIdx.x & 31;
    size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = w * 32; i < n; i += nw * 32) if (lane < 8) red4(out + i + lane * 4, v, v, v, v);
}
// scalar RED, each lane 4 consecutive-row cells (the MODE 0 deposit pattern: 4 REDs per ray, unit stride per instruction)
__global__ void k_red1x4(float* out, size_t n, float v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 8193 < n; i += st) { red1(out + i, v); red1(out + i + 1, v); red1(out + i + 8192, v); red1(out + i + 8193, v); }
}
// read 3 streams (12 B/px) + one v4 RED per 4 px: the traffic shape of an ideal splat
__global__ void k_read3_red4(const float* a, const float* b, const float* c, float* out, size_t n) {
    size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4, st = (size_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += st) {
        float4 x = __ldg(reinterpret_cast<const float4*>(a + i)), y = __ldg(reinterpret_cast<const float4*>(b + i)), z = __ldg(reinterpret_cast<const float4*>(c + i));
        red4(out + i, x.x + y.x * z.x, x.y + y.y * z.y, x.z + y.z * z.z, x.w + y.w * z.w);
    }
}
__global__ void k_read3_red1(const float* a, const float* b, const float* c, float* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) red1(out + i, __ldg(a + i) + __ldg(b + i) * __ldg(c + i));
}
__global__ void k_read3_store4(const float* a, const float* b, const float* c, float* out, size_t n) {
    size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4, st = (size_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += st) {
        float4 x = __ldg(reinterpret_cast<const float4*>(a + i)), y = __ldg(reinterpret_cast<const float4*>(b + i)), z = __ldg(reinterpret_cast<const float4*>(c + i));
        *reinterpret_cast<float4*>(out + i) = make_float4(x.x + y.x * z.x, x.y + y.y * z.y, x.z + y.z * z.z, x.w + y.w * z.w);
    }
}
// shared memory: conflict-free adds, ITER per thread; result written so nothing is optimised away
template <int KIND>
__global__ void k_smem(float* out, int iters, float v) {
    __shared__ float s[4096];
    for (int t = threadIdx.x; t < 4096; t += blockDim.x) s[t] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* mine = s + warp * 512;           // 8 warps x 512 floats: warp-private region
    for (int it = 0; it < iters; ++it) {
        const int a = ((it * 33) & 448) + lane + (it & 31);   // unit stride across lanes
        if (KIND == 0) atomicAdd(mine + (a & 511), v);
        else if (KIND == 1) { float x = mine[a & 511]; mine[a & 511] = x + v; }
        else { float x = mine[a & 511]; mine[a & 511] = x + v; x = mine[(a + 1) & 511]; mine[(a + 1) & 511] = x + v; }
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s[threadIdx.x];
}

static float* flushbuf; static size_t flushn = 64u << 20;
template <typename F> static void run(const char* name, size_t n, double bytes_per_elem, F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0; const int reps = 7;
    for (int r = 0; r < reps + 2; ++r) {
        if (n * 4 > (100u << 20)) k_store4<<<148 * 8, 256>>>(flushbuf, flushn, 0.f);   // evict L2 for DRAM-sized cases
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    CK(cudaGetLastError());
    printf("{\"kernel\": \"%s\", \"n\": %zu, \"ms\": %.4f, \"ms_best\": %.4f, \"GBps\": %.1f, \"GBps_best\": %.1f}\n", name, n, sum / reps, best,
           n * bytes_per_elem / (sum / reps) / 1e6, n * bytes_per_elem / best / 1e6);
    fflush(stdout);
}

int main() {
    CK(cudaMalloc(&flushbuf, flushn * 4));
    for (size_t side : {2048, 4096, 8192}) {
        const size_t n = side * side;
        float *out, *a, *b, *c;
        CK(cudaMalloc(&out, n * 4)); CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4));
        CK(cudaMemset(out, 0, n * 4)); CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4)); CK(cudaMemset(c, 0, n * 4));
        for (int bps : {4, 8}) {
            const int g = 148 * bps;
            char nm[64];
            #define NAME(s) (snprintf(nm, sizeof nm, "%s_g%d", s, bps), nm)
            run(NAME("store4"), n, 4, [&] { k_store4<<<g, 256>>>(out, n, 1.f); });
            run(NAME("rmw4"), n, 8, [&] { k_rmw4<<<g, 256>>>(out, n, 1.f); });
            run(NAME("red1"), n, 4, [&] { k_red1<<<g, 256>>>(out, n, 1.f); });
            run(NAME("red2"), n, 4, [&] { k_red2<<<g, 256>>>(out, n, 1.f); });
            run(NAME("red4"), n, 4, [&] { k_red4<<<g, 256>>>(out, n, 1.f); });
            run(NAME("red4_8lanes"), n, 4, [&] { k_red4_8lanes<<<g, 256>>>(out, n, 1.f); });
            run(NAME("red1x4"), n, 4, [&] { k_red1x4<<<g, 256>>>(out, n, 1.f); });
            run(NAME("read3_store4"), n, 16, [&] { k_read3_store4<<<g, 256>>>(a, b, c, out, n); });
            run(NAME("read3_red4"), n, 16, [&] { k_read3_red4<<<g, 256>>>(a, b, c, out, n); });
            run(NAME("read3_red1"), n, 16, [&] { k_read3_red1<<<g, 256>>>(a, b, c, out, n); });
        }
        CK(cudaFree(out)); CK(cudaFree(a)); CK(cudaFree(b)); CK(cudaFree(c));
    }
    // shared memory adds: 148*4 blocks x 256 threads x iters adds
    float* o; CK(cudaMalloc(&o, 148 * 4 * 256 * 4));
    const int iters = 4096; const size_t adds = (size_t)148 * 4 * 256 * iters;
    run("smem_atomic_add", adds, 1, [&] { k_smem<0><<<148 * 4, 256>>>(o, iters, 1.f); });
    run("smem_plain_rmw", adds, 1, [&] { k_smem<1><<<148 * 4, 256>>>(o, iters, 1.f); });
    run("smem_plain_rmw_x2", adds * 2, 1, [&] { k_smem<2><<<148 * 4, 256>>>(o, iters, 1.f); });
    return 0;
}
