// The owner-computes rolling-strip splat with TWO source columns per thread (variants 4 / 5 of paresis_splat on
// 16-byte-aligned images whose width is a multiple of 4; other shapes take the one-column kernel of splat_strip.cu).
// Reference: refractionFileNumba2.py:198-263 called by fastRefraction on a fresh zero array (:70, :77).
//
// Why two columns: the one-column kernel is instruction-bound (profiles/r02_summary.md: ~80 warp-instructions per 32 rays,
// of which ~35 are not the ray itself -- loop control, the list test, the flush, re-materialised constants).  A thread
// that owns two columns shares the row window, the loop and the list test between its two rays.  A block covers 512
// source columns and owns up to 488, so the column halo (2 x 12) costs 4.7 % instead of 9.4 %.
// The two columns are adjacent (2t, 2t+1): I, Dx, Dy arrive as 64-bit words.  (Columns 32 apart -- consecutive tile
// columns across the lanes of a warp -- were tried against shared-memory bank conflicts: same conflict count, it is the
// torn field that makes lanes collide, and 11 % more instructions: 231 vs 222 us at 8192^2, profiles/r02_summary.md.)
// Everything else -- ownership windows, circular fixed-point tile, per-warp ray lists, drain launch -- is strip.cuh's.
#include <type_traits>

#include "strip.cuh"

namespace paresis {

constexpr int S2_H = 12;                 // reach of the tile path
constexpr int S2_COLS = 2 * STRIP_THREADS;
constexpr int S2_OC_MAX = S2_COLS - 2 * S2_H;       // 488 owned columns
constexpr int S2_W = S2_OC_MAX + 12;     // words per tile row: pad quad + owned + garbage quad, +4 so that rows 2 apart sit on different banks
constexpr int S2_FIX = 21;
constexpr int S2_TILE_WORDS = STRIP_SLOTS * S2_W;
static_assert(S2_W % 4 == 0 && S2_W % 32 != 0 && S2_W % 32 != 16 && S2_W / 4 <= 128, "tile rows: quads, no bank alignment, two flush quads per thread");
static_assert(2 * STRIP_U + 2 * S2_H <= STRIP_SLOTS, "rows being flushed + rows being deposited must fit the circular tile");
static_assert(((unsigned long long)(2 * S2_H + 1) * (2 * S2_H + 1) << (S2_FIX + 1)) < (1ull << 32), "a cell cannot overflow");

// the two columns of a thread in one 64-bit streaming load whose place in the instruction stream is fixed (see strip_prefetch)
__device__ __forceinline__ void ld2(float2& dst, const float* p) {
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(dst.x), "=f"(dst.y) : "l"(p));
}

struct Ray2 {          // what one ray needs between its arithmetic and its atomics
    unsigned a0, a1, w0, w1, w2, w3, bx, by;
    bool ok;
};

// the arithmetic of one ray, branch-free: window tests, fixed-point split, tile addresses
__device__ __forceinline__ Ray2 ray_math(const RowWin& rw, unsigned csub, unsigned cspan, unsigned caddr, float v, float dx, float dy,
                                         float scale, unsigned vmin_bits, unsigned vspan, unsigned long long half) {
    constexpr float M = 12582912.f;
    Ray2 r;
    const float tx = __fadd_rd(dx, M), ty = __fadd_rd(dy, M);
    r.bx = __float_as_uint(tx); r.by = __float_as_uint(ty);
    r.ok = (r.bx - rw.sub) < rw.span && (r.by - csub) < cspan && (__float_as_uint(v) - vmin_bits) < vspan;
    const unsigned fx = __float_as_uint(__fadd_rz(dx, (M + 1.f) - tx)) << 9, fy = __float_as_uint(__fadd_rz(dy, (M + 1.f) - ty)) << 9;
    const unsigned V = __float_as_uint(fmaf(v, scale, 8388608.f)) - 0x4B000000u;      // round(v * scale), v * scale < 2^22
    const unsigned V1 = strip_mulhi_rn(V, fx, half), V0 = V - V1;
    r.w1 = strip_mulhi_rn(V0, fy, half); r.w0 = V0 - r.w1;
    r.w3 = strip_mulhi_rn(V1, fy, half); r.w2 = V1 - r.w3;
    const unsigned s0 = (r.bx + rw.slot) & (STRIP_SLOTS - 1), s1 = (r.bx + rw.slot + 1u) & (STRIP_SLOTS - 1);
    const unsigned col = caddr + r.by * 4u;
    r.a0 = col + s0 * (S2_W * 4u); r.a1 = col + s1 * (S2_W * 4u);
    return r;
}

template <bool ACC>
__global__ void __launch_bounds__(STRIP_THREADS, 3)
splat_strip2_kernel(const float* __restrict__ I, const float* __restrict__ Dx, const float* __restrict__ Dy, float* __restrict__ out,
                    Frame f, StripPlan p, uint4* __restrict__ far, unsigned* __restrict__ far_count) {
    constexpr int H = S2_H, U = STRIP_U, W = S2_W, Q = S2_W / 4;
    extern __shared__ __align__(16) unsigned strip_smem[];
    unsigned* const tile = strip_smem;
    unsigned* const misc = strip_smem + S2_TILE_WORDS;      // [0..7] / [8..15]: warp partials of the scale sample

    const int tid = threadIdx.x, lane = tid & 31;
    const int C0 = blockIdx.x * p.oc;
    const int oc = min(p.oc, f.ny - C0);
    const int R0 = blockIdx.y * p.seg_rows, R1 = min(R0 + p.seg_rows, f.nx);
    const int j0 = C0 - H + 2 * tid, j1 = j0 + 1;                // this thread's source columns (j0 is even)
    const bool live0 = j0 >= 0 && j0 < f.ny && 2 * tid < p.oc + 2 * H;
    const bool live1 = j1 >= 0 && j1 < f.ny && 2 * tid + 1 < p.oc + 2 * H;
    const int jc = min(max(j0, 0), f.ny - 2);
    const bool own0 = j0 >= C0 && j0 < C0 + oc, own1 = j1 >= C0 && j1 < C0 + oc;
    const unsigned wid = (blockIdx.y * gridDim.x + blockIdx.x) * STRIP_WARPS + (tid >> 5);
    uint4* const slice = far + (size_t)wid * p.far_cap;
    unsigned n_far = 0u;

    {
        uint4* z = reinterpret_cast<uint4*>(tile);
        for (int k = tid; k < S2_TILE_WORDS / 4; k += STRIP_THREADS) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    // the intensity scale: the same 256 samples in every block (splat_strip.cu)
    {
        const int si = (int)(((long long)(2 * (tid >> 4) + 1) * f.nx) >> 5), sj = (int)(((long long)(2 * (tid & 15) + 1) * f.ny) >> 5);
        const float t = __ldg(I + (size_t)si * f.ny + sj);
        const bool good = t > 0.f && t < 3.0e38f;
        float sum = good ? t : 0.f;
        const unsigned cnt = __popc(__ballot_sync(FULL_MASK, good));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, d);
        if (lane == 0) { misc[tid >> 5] = __float_as_uint(sum); misc[8 + (tid >> 5)] = cnt; }
    }

    const int s_begin = max(R0 - H, 0), s_end = min(R1 + H, f.nx);
    const int last_off = (s_end - 1) * f.ny + jc;               // nx * ny < 2^30 (host check)
    int pre = s_begin * f.ny + jc;
    float2 vq[U], dxq[U], dyq[U];                               // .x: column j0, .y: column j1
#pragma unroll
    for (int u = 0; u < U; ++u) {
        ld2(vq[u], I + pre); ld2(dxq[u], Dx + pre); ld2(dyq[u], Dy + pre);
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
    const unsigned mexp = __float_as_uint(m) >> 23;
    const bool fixed_ok = mexp >= 32u && mexp < 254u;
    const float scale = __uint_as_float((254u + S2_FIX - mexp) << 23), inv_scale = __uint_as_float((mexp - S2_FIX) << 23);
    const unsigned vmin_bits = (mexp + 9u - S2_FIX) << 23, vspan = fixed_ok ? ((unsigned)(S2_FIX - 8) << 23) : 0u;

    // column windows of the two rays (col_window(), with the block's 488-column ownership)
    const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
    unsigned csub0, cspan0, csub1, cspan1, caddr0;
    {
        const int lo0 = max(max(-H, C0 - 1 - j0), -j0), hi0 = min(min(H - 1, C0 + p.oc - 1 - j0), f.ny - 2 - j0);
        const int lo1 = max(max(-H, C0 - 1 - j1), -j1), hi1 = min(min(H - 1, C0 + p.oc - 1 - j1), f.ny - 2 - j1);
        csub0 = STRIP_MAGIC + (unsigned)lo0; cspan0 = live0 ? (unsigned)max(hi0 - lo0 + 1, 0) : 0u;
        csub1 = STRIP_MAGIC + (unsigned)lo1; cspan1 = live1 ? (unsigned)max(hi1 - lo1 + 1, 0) : 0u;
        caddr0 = tile_s + (unsigned)(j0 - C0 + STRIP_PAD) * 4u - STRIP_MAGIC * 4u;
        asm volatile("" : "+r"(csub0), "+r"(cspan0), "+r"(csub1), "+r"(cspan1), "+r"(caddr0));
    }
    const int in_lo = max(R0 + H - 1, H), in_hi = min(R1 - H, f.nx - 1 - H);
    const unsigned long long half = strip_half(f.nx);
    int flush_next = R0 - 1;

    // the two rays of source pixels (i, j0) and (i, j1): into the tile, or -- if this block owns the pixel and no
    // block's tile takes the ray -- on the warp's list
    auto rays = [&](int i, const RowWin& rw, bool own_row, const float2 v, const float2 dx, const float2 dy) {
        const Ray2 a = ray_math(rw, csub0, cspan0, caddr0, v.x, dx.x, dy.x, scale, vmin_bits, vspan, half);
        const Ray2 b = ray_math(rw, csub1, cspan1, caddr0 + 4u, v.y, dx.y, dy.y, scale, vmin_bits, vspan, half);
        if (a.ok) strip_reds4(a.a0, a.a1, a.w0, a.w1, a.w2, a.w3);
        if (b.ok) strip_reds4(b.a0, b.a1, b.w0, b.w1, b.w2, b.w3);
        const bool ca = !a.ok && own_row && own0 && v.x != 0.f, cb = !b.ok && own_row && own1 && v.y != 0.f;
        if (__any_sync(FULL_MASK, ca || cb)) {                      // warp-uniform; rare away from the strip's edges
            const bool pa = ca && !(tile_class<H>(i, j0, a.bx, a.by, f.nx, f.ny) && (__float_as_uint(v.x) - vmin_bits) < vspan);
            const bool pb = cb && !(tile_class<H>(i, j1, b.bx, b.by, f.nx, f.ny) && (__float_as_uint(v.y) - vmin_bits) < vspan);
            strip_push(slice, n_far, pa, (unsigned)(i * f.ny + j0), v.x, dx.x, dy.x);
            strip_push(slice, n_far, pb, (unsigned)(i * f.ny + j1), v.y, dx.y, dy.y);
        }
    };

    for (int s = s_begin; s < s_end; s += U) {
        if (s >= in_lo && s + U - 1 <= in_hi) {                    // block-uniform
#pragma unroll
            for (int u = 0; u < U; ++u) {
                rays(s + u, row_window_interior<H>(s + u), true, vq[u], dxq[u], dyq[u]);
                ld2(vq[u], I + pre); ld2(dxq[u], Dx + pre); ld2(dyq[u], Dy + pre);            // row s + u + U (clamped)
                pre = min(pre + f.ny, last_off);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = s + u;
                if (i >= s_end) break;                             // block-uniform
                rays(i, row_window<H>(i, R0, R1, f.nx), i >= R0 && i < R1, vq[u], dxq[u], dyq[u]);
                ld2(vq[u], I + pre); ld2(dxq[u], Dx + pre); ld2(dyq[u], Dy + pre);
                pre = min(pre + f.ny, last_off);
            }
        }
        __syncthreads();
        // rows that no later source row can reach are final: out they go, by all threads (two quads each), slots zeroed.
        // Quad 0 of a tile row is the pad + left garbage column, quads 1 .. oc/4 the owned columns, then the right garbage column.
        const int final_row = s + U >= s_end ? R1 - 1 : s + U - 1 - H;
        while (flush_next <= final_row) {
            const int rb = min(flush_next + 3, final_row);
            const int r = flush_next + (tid >> 6);
            if (r <= rb) {
                uint4* row = reinterpret_cast<uint4*>(tile + (r & (STRIP_SLOTS - 1)) * W);
                float4* g = reinterpret_cast<float4*>(out + (size_t)r * f.ny + C0 - 4);           // quad q starts at column C0 + 4 (q - 1)
                const bool store = r >= R0;                                                      // row R0 - 1: upper garbage row
                const int qa = tid & 63, qb = qa + 64;
                uint4 ua = row[qa], ub = make_uint4(0u, 0u, 0u, 0u);
                row[qa] = make_uint4(0u, 0u, 0u, 0u);
                if (qb < Q) { ub = row[qb]; row[qb] = make_uint4(0u, 0u, 0u, 0u); }
                const bool ma = store && qa >= 1 && 4 * (qa - 1) < oc, mb = store && qb < Q && 4 * (qb - 1) < oc;
                float4 oa = make_float4((float)ua.x * inv_scale, (float)ua.y * inv_scale, (float)ua.z * inv_scale, (float)ua.w * inv_scale);
                float4 ob = make_float4((float)ub.x * inv_scale, (float)ub.y * inv_scale, (float)ub.z * inv_scale, (float)ub.w * inv_scale);
                if (ACC) {
                    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
                    if (ma) ba = g[qa];
                    if (mb) bb = g[qb];
                    oa.x += ba.x; oa.y += ba.y; oa.z += ba.z; oa.w += ba.w;
                    ob.x += bb.x; ob.y += bb.y; ob.z += bb.z; ob.w += bb.w;
                }
                if (ma) g[qa] = oa;
                if (mb) g[qb] = ob;
            }
            flush_next = rb + 1;
        }
    }
    if (lane == 0) far_count[wid] = n_far;
}

int launch_splat_drain(const uint4* far, const unsigned* far_count, unsigned far_cap, unsigned n_slices, float* out, const Frame& f, int* flag,
                       cudaStream_t s);      // splat_strip.cu

static DeviceSlots g_splat2_slots[2];

template <bool ACC>
static int launch_variant2(const float* I, const float* Dx, const float* Dy, float* out, const Frame& f, int* flag, cudaStream_t s) {
    constexpr size_t smem = sizeof(unsigned) * (S2_TILE_WORDS + 32);
    int slots = 0;
    int rc = g_splat2_slots[ACC ? 1 : 0].get(splat_strip2_kernel<ACC>, STRIP_THREADS, smem, &slots);
    if (rc) return rc;
    StripPlan p;
    p.strips = div_up(f.ny, S2_OC_MAX);
    p.oc = (div_up(f.ny, p.strips) + 3) / 4 * 4;
    int segs = slots / p.strips;
    if (segs < 1) segs = 1;
    int rows = div_up(f.nx, segs);
    const int min_rows = 6 * S2_H;
    if (rows < min_rows) rows = min_rows < f.nx ? min_rows : f.nx;
    p.segs = div_up(f.nx, rows);
    p.seg_rows = div_up(f.nx, p.segs);
    p.far_cap = (unsigned)p.seg_rows * 64u;                    // a warp owns at most 64 columns x seg_rows pixels
    const size_t slices = (size_t)p.strips * p.segs * STRIP_WARPS;
    const size_t list_bytes = slices * p.far_cap * sizeof(uint4);
    void* scratch = nullptr;
    rc = strip_scratch_alloc(list_bytes + slices * sizeof(unsigned), &scratch, s);
    if (rc) return rc;
    uint4* far = reinterpret_cast<uint4*>(scratch);
    unsigned* far_count = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(scratch) + list_bytes);
    dim3 grid(p.strips, p.segs);
    splat_strip2_kernel<ACC><<<grid, STRIP_THREADS, smem, s>>>(I, Dx, Dy, out, f, p, far, far_count);
    PARESIS_LAUNCH_CHECK("splat_strip2_kernel");
    rc = launch_splat_drain(far, far_count, p.far_cap, (unsigned)slices, out, f, flag, s);
    if (rc) return rc;
    return strip_scratch_free(scratch, s);
}

// 16-byte-aligned images with ny % 4 == 0 and at least 64 columns: 64-bit loads and 128-bit stores line up
bool splat_strip2_fits(const float* I, const float* Dx, const float* Dy, const float* out, const Frame& f) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(I) | reinterpret_cast<uintptr_t>(Dx) | reinterpret_cast<uintptr_t>(Dy) |
                           reinterpret_cast<uintptr_t>(out);
    return (bits & 15) == 0 && (f.ny & 3) == 0 && f.ny >= 64 && f.nx >= 2;
}

int launch_splat_strip2(const float* I, const float* Dx, const float* Dy, float* out, const Frame& f, int* flag, bool accumulate,
                        cudaStream_t s) {
    return accumulate ? launch_variant2<true>(I, Dx, Dy, out, f, flag, s) : launch_variant2<false>(I, Dx, Dy, out, f, flag, s);
}

}  // namespace paresis
