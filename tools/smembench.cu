// Micro-benchmark: shared-memory accumulate variants (design input for the tile splat).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/smembench.bin tools/smembench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// KIND 0: float atomicAdd (CAS loop)  1: u32 atomicAdd  2: u64 atomicAdd  3: plain float RMW  4: float->s64 convert + u64 atomicAdd
// PATTERN 0: unit stride, all lanes distinct; 1: lanes pairwise collide (addr = lane/2 ...), 2: "torn" rows: lane groups of 8 on 4 different rows
template <int KIND, int PATTERN>
__global__ void k_acc(float* out, int iters, float v) {
    __shared__ unsigned long long s64[2048];
    float* sf = reinterpret_cast<float*>(s64);
    unsigned* su = reinterpret_cast<unsigned*>(s64);
    for (int t = threadIdx.x; t < 2048; t += blockDim.x) s64[t] = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int it = 0; it < iters; ++it) {
        int a;
        if (PATTERN == 0) a = lane + (it & 31);
        else if (PATTERN == 1) a = (lane >> 1) + (it & 31);
        else a = (lane & 7) + (it & 15) + 64 * (lane >> 3);
        a = (a + warp * 200) & 2047;
        if (KIND == 0) atomicAdd(sf + a, v);
        else if (KIND == 1) atomicAdd(su + a, (unsigned)it);
        else if (KIND == 2) atomicAdd(s64 + a, (unsigned long long)it);
        else if (KIND == 3) { float x = sf[a]; sf[a] = x + v; }
        else atomicAdd(s64 + a, (unsigned long long)__float2ll_rn(v * (float)(it + 1)));
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s64[threadIdx.x];
}

template <typename F> static void run(const char* name, size_t adds, F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    // cycles per warp-op per SM at 1.965 GHz, 148 SMs
    const double warp_ops_per_sm = adds / 32.0 / 148.0;
    printf("{\"kernel\": \"%s\", \"ms\": %.4f, \"Gadds_per_s\": %.1f, \"cycles_per_warp_op_per_SM\": %.2f}\n", name, best, adds / best / 1e6,
           best * 1e-3 * 1.965e9 / warp_ops_per_sm);
    fflush(stdout);
}

int main() {
    float* o; CK(cudaMalloc(&o, 148 * 4 * 256 * 4));
    const int iters = 4096; const size_t adds = (size_t)148 * 4 * 256 * iters;
#define RUN(K, P, NAME) run(NAME, adds, [&] { k_acc<K, P><<<148 * 4, 256>>>(o, iters, 1.f); })
    RUN(0, 0, "f32_cas_distinct"); RUN(0, 1, "f32_cas_pairs"); RUN(0, 2, "f32_cas_torn");
    RUN(1, 0, "u32_add_distinct"); RUN(1, 1, "u32_add_pairs"); RUN(1, 2, "u32_add_torn");
    RUN(2, 0, "u64_add_distinct"); RUN(2, 1, "u64_add_pairs"); RUN(2, 2, "u64_add_torn");
    RUN(3, 0, "f32_plain_distinct"); RUN(3, 2, "f32_plain_torn");
    RUN(4, 0, "cvt_u64_add_distinct"); RUN(4, 1, "cvt_u64_add_pairs");
    return 0;
}
