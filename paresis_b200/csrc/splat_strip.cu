// Stand-alone splat (fastloopNumba: I, Dx, Dy -> out) as an OWNER-COMPUTES rolling-strip kernel: variants 4 and 5 of
// paresis_splat.  Reference: refractionFileNumba2.py:198-263 (same body: refractionFileNumba.py:70-135), called by
// fastRefraction on a freshly zeroed array (:70, :77).
//
//   variant 4: out  = splat(I, Dx, Dy)   -- fastRefraction's contract; 16 B/px of HBM traffic (read 12, store 4), no
//                                           zero-fill, no read-modify-write
//   variant 5: out += splat(I, Dx, Dy)   -- fastloopNumba's own contract; the owner adds its finished rows (20 B/px),
//                                           no atomics on the image
//
// The machinery (circular fixed-point tile, ownership windows, ray list + drain launch) is in strip.cuh.  The API
// carries no intensity scale: every block takes the same one from a fixed 16 x 16 lattice of the intensity image
// (twice the mean of its finite positive samples, rounded down to a power of two -- the scaling itself is exact), so
// that all blocks agree on which rays are tile rays.  Rays outside [2^-13, 2) of that scale, like rays that move
// further than H = 12 pixels or touch the image border, go through make_ray() -- the reference's loop-frame rules --
// in the drain launch.
#include <mutex>
#include <type_traits>

#include "strip.cuh"

namespace paresis {

constexpr int SPLAT_H = 12;     // reach of the tile path; on a membrane's field 0.16 % of the rays move further (1.9 % at H = 8)

template <int H, bool ACC>
__global__ void __launch_bounds__(STRIP_THREADS, 5)
splat_strip_kernel(const float* __restrict__ I, const float* __restrict__ Dx, const float* __restrict__ Dy, float* __restrict__ out,
                   Frame f, StripPlan p, uint4* __restrict__ far, unsigned* __restrict__ far_count, bool vec) {
    using S = Strip<H>;
    constexpr int U = STRIP_U;
    extern __shared__ __align__(16) unsigned strip_smem[];
    unsigned* const tile = strip_smem;
    unsigned* const misc = strip_smem + S::TILE_WORDS;      // [0..7] / [8..15]: warp partials of the scale sample

    const int tid = threadIdx.x, lane = tid & 31;
    const int C0 = blockIdx.x * p.oc;
    const int oc = min(p.oc, f.ny - C0);
    const int R0 = blockIdx.y * p.seg_rows, R1 = min(R0 + p.seg_rows, f.nx);
    const int j = C0 - H + tid;
    const bool live = j >= 0 && j < f.ny && tid < p.oc + 2 * H;
    const int jc = min(max(j, 0), f.ny - 1);
    const bool own_col = j >= C0 && j < C0 + oc;
    const unsigned wid = (blockIdx.y * gridDim.x + blockIdx.x) * STRIP_WARPS + (tid >> 5);
    uint4* const slice = far + (size_t)wid * p.far_cap;
    unsigned n_far = 0u;

    {
        uint4* z = reinterpret_cast<uint4*>(tile);
        for (int k = tid; k < S::TILE_WORDS / 4; k += STRIP_THREADS) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    // the intensity scale: the same 256 samples in every block
    {
        const int si = (int)(((long long)(2 * (tid >> 4) + 1) * f.nx) >> 5), sj = (int)(((long long)(2 * (tid & 15) + 1) * f.ny) >> 5);
        const float t = __ldg(I + (size_t)si * f.ny + sj);
        const bool good = t > 0.f && t < 3.0e38f;              // NaN fails the compares
        float sum = good ? t : 0.f;
        const unsigned cnt = __popc(__ballot_sync(FULL_MASK, good));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, d);
        if (lane == 0) { misc[tid >> 5] = __float_as_uint(sum); misc[8 + (tid >> 5)] = cnt; }
    }

    // source rows of this block and the first U of them
    const int s_begin = max(R0 - H, 0), s_end = min(R1 + H, f.nx);
    const int last_off = (s_end - 1) * f.ny + jc;               // nx * ny < 2^30 (host check)
    int pre = s_begin * f.ny + jc;
    float vq[U], dxq[U], dyq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        strip_prefetch(vq[u], I + pre); strip_prefetch(dxq[u], Dx + pre); strip_prefetch(dyq[u], Dy + pre);
        pre = min(pre + f.ny, last_off);
    }
    __syncthreads();
    float m;
    {
        float sum = 0.f; unsigned cnt = 0u;
#pragma unroll
        for (int w = 0; w < STRIP_WARPS; ++w) { sum += __uint_as_float(misc[w]); cnt += misc[8 + w]; }
        m = cnt ? 2.f * sum / (float)cnt : 0.f;
    }
    // 2 * mean = 1.x * 2^e: unit 2^(e + 1 - FIX) ... tile rays in [2^9, 2^(FIX+1)) units = [2^(e + 10 - FIX), 2^(e + 1))
    const unsigned mexp = __float_as_uint(m) >> 23;
    const bool fixed_ok = mexp >= 32u && mexp < 254u;
    const float scale = __uint_as_float((254u + S::FIX - mexp) << 23), inv_scale = __uint_as_float((mexp - S::FIX) << 23);
    const unsigned vmin_bits = (mexp + 9u - S::FIX) << 23, vspan = fixed_ok ? ((unsigned)(S::FIX - 8) << 23) : 0u;

    const ColWin cw = col_window<H>(j, C0, p.oc, f.ny, live, (unsigned)__cvta_generic_to_shared(tile));
    const FlushLane fl = flush_lane<S::W>(C0, oc);
    int flush_next = R0 - 1;
    // rows whose deposit window is the full [-H, H-1] and that this block owns as source rows: most of a segment
    const int in_lo = max(R0 + H - 1, H), in_hi = min(R1 - H, f.nx - 1 - H);
    // warp-uniform: within H columns of every lane there is no image border (all four cells of a reach-H ray are image cells)
    const bool cols_simple = __all_sync(FULL_MASK, j - H >= 0 && j + H <= f.ny - 1);
    const unsigned long long half = strip_half(f.nx);

    // The ray of source pixel (i, j): into the tile, or -- if this block owns the pixel and no block's tile takes the
    // ray -- on the list.  On interior rows away from the image border "no tile takes it" is a two-compare test, so the
    // warps at the edge of a strip (whose rays often belong to the neighbour's tile) cost what the others cost.
    auto ray = [&](auto IC, int i, const RowWin& rw, bool own_row, float v, float dx, float dy) {
        constexpr bool INTERIOR = decltype(IC)::value;
        unsigned bx, by;
        const bool ok = strip_deposit<S::W, false>(rw, cw, v, dx, dy, scale, vmin_bits, vspan, 0u, half, bx, by);
        if (INTERIOR && cols_simple) {
            const bool tile_ray = (bx - (STRIP_MAGIC - (unsigned)H)) < 2u * H && (by - (STRIP_MAGIC - (unsigned)H)) < 2u * H &&
                                  (__float_as_uint(v) - vmin_bits) < vspan;
            const bool push = !tile_ray && own_col && v != 0.f;
            if (__any_sync(FULL_MASK, push)) strip_push(slice, n_far, push, (unsigned)(i * f.ny + j), v, dx, dy);
        } else {
            const bool cand = !ok && own_row && own_col && v != 0.f;
            if (__any_sync(FULL_MASK, cand)) {
                const bool push = cand && !(tile_class<H>(i, j, bx, by, f.nx, f.ny) && (__float_as_uint(v) - vmin_bits) < vspan);
                strip_push(slice, n_far, push, (unsigned)(i * f.ny + j), v, dx, dy);
            }
        }
    };

    for (int s = s_begin; s < s_end; s += U) {
        if (s >= in_lo && s + U - 1 <= in_hi) {                    // block-uniform
#pragma unroll
            for (int u = 0; u < U; ++u) {
                ray(std::true_type{}, s + u, row_window_interior<H>(s + u), true, vq[u], dxq[u], dyq[u]);
                strip_prefetch(vq[u], I + pre); strip_prefetch(dxq[u], Dx + pre); strip_prefetch(dyq[u], Dy + pre);    // row s + u + U (clamped)
                pre = min(pre + f.ny, last_off);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = s + u;
                if (i >= s_end) break;                             // block-uniform
                ray(std::false_type{}, i, row_window<H>(i, R0, R1, f.nx), i >= R0 && i < R1, vq[u], dxq[u], dyq[u]);
                strip_prefetch(vq[u], I + pre); strip_prefetch(dxq[u], Dx + pre); strip_prefetch(dyq[u], Dy + pre);
                pre = min(pre + f.ny, last_off);
            }
        }
        __syncthreads();
        // rows that no later source row can reach are final: out they go, by all threads, and their slots are zeroed
        const int final_row = s + U >= s_end ? R1 - 1 : s + U - 1 - H;
        while (flush_next <= final_row) {
            const int rb = min(flush_next + 3, final_row);
            strip_flush_rows<S::W, ACC>(tile, out, fl, flush_next, rb, R0, R1, oc, C0, f.ny, inv_scale, vec);
            flush_next = rb + 1;
        }
    }
    if (lane == 0) far_count[wid] = n_far;
}

// The listed rays, in fp32, with the reference's loop-frame rules (make_ray).  One warp per list slice.
__global__ void __launch_bounds__(128)
splat_drain_kernel(const uint4* __restrict__ far, const unsigned* __restrict__ far_count, unsigned far_cap, unsigned n_slices,
                   float* __restrict__ out, Frame f, int* flag) {
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_slices) return;
    const unsigned n = far_count[w];
    if (n == 0u) return;
    const uint4* slice = far + (size_t)w * far_cap;
    Splatter<0> sp;
    sp.init(out, f.ny, flag);
    for (unsigned k = threadIdx.x & 31; k < n; k += 32) {
        const uint4 e = slice[k];
        const int idx = (int)(e.x & FAR_INDEX);
        const int i = idx / f.ny, j = idx - i * f.ny;
        const Ray q = make_ray(i, j, __uint_as_float(e.y), __uint_as_float(e.z), __uint_as_float(e.w), f);
        if (q.ok) {
            const float s = (q.w[0] + q.w[1]) + (q.w[2] + q.w[3]);
            sp.bad |= !(fabsf(s) <= 3.0e38f);
            sp.cells(q);
        }
    }
    sp.finish();
}

int launch_splat_drain(const uint4* far, const unsigned* far_count, unsigned far_cap, unsigned n_slices, float* out, const Frame& f, int* flag,
                       cudaStream_t s) {
    splat_drain_kernel<<<(n_slices + 3) / 4, 128, 0, s>>>(far, far_count, far_cap, n_slices, out, f, flag);
    PARESIS_LAUNCH_CHECK("splat_drain_kernel");
    return PARESIS_OK;
}

bool splat_strip2_fits(const float* I, const float* Dx, const float* Dy, const float* out, const Frame& f);      // splat_strip2.cu
int launch_splat_strip2(const float* I, const float* Dx, const float* Dy, float* out, const Frame& f, int* flag, bool accumulate, cudaStream_t s);

static DeviceSlots g_splat_slots[2];

template <bool ACC>
static int launch_variant(const float* I, const float* Dx, const float* Dy, float* out, const Frame& f, int* flag, cudaStream_t s) {
    constexpr int H = SPLAT_H;
    constexpr size_t smem = sizeof(unsigned) * (Strip<H>::TILE_WORDS + 32);
    int slots = 0;
    int rc = g_splat_slots[ACC ? 1 : 0].get(splat_strip_kernel<H, ACC>, STRIP_THREADS, smem, &slots);
    if (rc) return rc;
    const StripPlan p = plan_strips(f.nx, f.ny, H, slots);
    const size_t slices = (size_t)p.strips * p.segs * STRIP_WARPS;
    const size_t list_bytes = slices * p.far_cap * sizeof(uint4);
    void* scratch = nullptr;
    rc = strip_scratch_alloc(list_bytes + slices * sizeof(unsigned), &scratch, s);
    if (rc) return rc;
    uint4* far = reinterpret_cast<uint4*>(scratch);
    unsigned* far_count = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(scratch) + list_bytes);
    const bool vec = (f.ny & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    dim3 grid(p.strips, p.segs);
    splat_strip_kernel<H, ACC><<<grid, STRIP_THREADS, smem, s>>>(I, Dx, Dy, out, f, p, far, far_count, vec);
    PARESIS_LAUNCH_CHECK("splat_strip_kernel");
    splat_drain_kernel<<<(unsigned)div_up((int)slices, 4), 128, 0, s>>>(far, far_count, p.far_cap, (unsigned)slices, out, f, flag);
    PARESIS_LAUNCH_CHECK("splat_drain_kernel");
    return strip_scratch_free(scratch, s);
}

int launch_splat_strip(const float* I, const float* Dx, const float* Dy, float* out, const Frame& f, int* flag, bool accumulate,
                       cudaStream_t s) {
    // aligned images: two source columns per thread (splat_strip2.cu); anything else: one column per thread (this file)
    if (splat_strip2_fits(I, Dx, Dy, out, f)) return launch_splat_strip2(I, Dx, Dy, out, f, flag, accumulate, s);
    return accumulate ? launch_variant<true>(I, Dx, Dy, out, f, flag, s) : launch_variant<false>(I, Dx, Dy, out, f, flag, s);
}

// ---- scratch for the ray lists --------------------------------------------------------------------------------------
// One cached block per device (the library's per-device workspace): it grows on demand and is handed from stream to
// stream through an event, so a call costs no allocation and no synchronisation on the same stream.  A second call
// while the block is out (another host thread) takes a stream-ordered allocation of its own.  paresis_trim() frees it.
namespace {
struct ScratchCache {
    void* ptr = nullptr;
    size_t bytes = 0;
    cudaStream_t last = nullptr;
    cudaEvent_t released = nullptr;
    bool out = false;
};
ScratchCache g_scratch[32];
std::mutex g_scratch_lock;
}  // namespace

int strip_scratch_alloc(size_t bytes, void** ptr, cudaStream_t s) {
    int dev = 0;
    PARESIS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 32) dev = 0;
    if (bytes < 256) bytes = 256;
    std::lock_guard<std::mutex> guard(g_scratch_lock);
    ScratchCache& c = g_scratch[dev];
    if (!c.out) {
        if (c.ptr && c.bytes < bytes) {                 // too small: give it back in the order of its last user
            PARESIS_CUDA(cudaFreeAsync(c.ptr, c.last));
            c.ptr = nullptr;
        }
        if (!c.ptr) {
            if (!c.released) PARESIS_CUDA(cudaEventCreateWithFlags(&c.released, cudaEventDisableTiming));
            PARESIS_CUDA(cudaMallocAsync(&c.ptr, bytes, s));
            c.bytes = bytes;
        } else if (c.last != s) {
            PARESIS_CUDA(cudaStreamWaitEvent(s, c.released, 0));
        }
        c.out = true;
        c.last = s;
        *ptr = c.ptr;
        return PARESIS_OK;
    }
    PARESIS_CUDA(cudaMallocAsync(ptr, bytes, s));
    return PARESIS_OK;
}

int strip_scratch_free(void* ptr, cudaStream_t s) {
    int dev = 0;
    PARESIS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 32) dev = 0;
    std::lock_guard<std::mutex> guard(g_scratch_lock);
    ScratchCache& c = g_scratch[dev];
    if (c.out && ptr == c.ptr) {
        PARESIS_CUDA(cudaEventRecord(c.released, s));
        c.last = s;
        c.out = false;
        return PARESIS_OK;
    }
    PARESIS_CUDA(cudaFreeAsync(ptr, s));
    return PARESIS_OK;
}

int strip_scratch_trim() {
    std::lock_guard<std::mutex> guard(g_scratch_lock);
    for (int d = 0; d < 32; ++d) {
        ScratchCache& c = g_scratch[d];
        if (c.ptr && !c.out) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(d);
            cudaFreeAsync(c.ptr, c.last);
            cudaSetDevice(cur);
            c.ptr = nullptr; c.bytes = 0;
        }
    }
    return PARESIS_OK;
}

}  // namespace paresis

extern "C" int paresis_trim(void) { return paresis::strip_scratch_trim(); }
