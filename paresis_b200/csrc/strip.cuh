// Owner-computes ROLLING-STRIP scatter: the machinery shared by the stand-alone splat (splat_strip.cu, variants 4/5
// of paresis_splat) and the fused refraction hops (refract_strip.cuh).
//
// Reference contract: fastloopNumba (refractionFileNumba2.py:198-263) scatters into an array that fastRefraction
// has just allocated as zeros (:70, :77) -- through fastRefraction the contract is an OVERWRITE.  The tile kernels of
// round 1 (refract_lean.cuh, splat_tile.cu) are source-owner kernels: a block bins the rays of ITS source tile and
// adds the tile to the image with REDs, so the image has to be zero-filled first and every line is read-modify-
// written in L2/DRAM (20 B/px where the contract needs 16).  Here a block OWNS OUTPUT cells instead:
//
//   * a block of 256 threads owns `oc` (<= 256 - 2H) output columns of one row segment [R0, R1) and walks the source
//     rows R0-H .. R1+H-1 of the source columns C0-H .. C0+oc+H-1, one thread per source column, U rows per step;
//   * a ray whose lower cell lies within H pixels of its source (floor(D) in [-H, H-1] on both axes), whose four
//     cells are strictly inside the image and whose intensity is in fixed-point range is a TILE ray: every block
//     that owns one of its four cells has the source pixel inside its source window, re-computes the ray and
//     deposits it (native 32-bit integer shared-memory atomics, see refract_lean.cuh for why fixed point) --
//     cells it does not own fall into one garbage row / column around the owned window and are never stored;
//   * the tile is CIRCULAR over rows (32 slots): once the walk is H rows past an output row that row is final,
//     is converted and leaves the SM with plain 128-bit stores (or `out +=` in ACC mode), and its slot is zeroed for
//     re-use.  Rows cost no halo re-computation inside a segment; the only redundant work is the H-wide column
//     halo (256 / (256 - 2H)) and 2H rows per segment;
//   * every other ray (further than H, touching the image border, too bright / dim / negative / not finite) is
//     pushed by the block that owns its SOURCE pixel on that block's private slice of a global list and deposited
//     in fp32 with the reference's edge rules by a small second launch (strip_drain_kernel) -- after the stores.
//
// No zero-fill pass, no read-modify-write of the image, results independent of the block schedule for tile rays
// (integer sums); fp32 summation order only matters for the listed rays.
//
// Fixed point: unit = 2^-FIX of the intensity scale (FIX = 22 for H <= 8, 21 for H = 12); a cell collects at most
// (2H+1)^2 rays of < 2^(FIX+1) units, so it cannot overflow; rays below 2^9 units go on the list instead of being
// quantised coarsely.  The list is per WARP (its 32 columns x the segment's rows): appending needs no atomics.
#pragma once
#include "splat.cuh"

namespace paresis {

constexpr int STRIP_THREADS = 256;   // source columns per block = depositing threads
constexpr int STRIP_WARPS = STRIP_THREADS / 32;
constexpr int STRIP_FLUSH_WARPS = 1;
constexpr int STRIP_BLOCK = STRIP_THREADS + 32 * STRIP_FLUSH_WARPS;   // + the warps that only flush finished rows (see strip_flusher)
constexpr int STRIP_U = 4;           // source rows per step (one barrier per step)
constexpr int STRIP_SLOTS = 32;      // rows of the circular tile
constexpr int STRIP_PAD = 4;         // owned column 0 sits at word 4 of a tile row; word 3 is the left garbage column
constexpr unsigned STRIP_MAGIC = 0x4B400000u;   // bit pattern of 1.5 * 2^23
constexpr unsigned FAR_REF = 0x40000000u, FAR_TWIN = 0x80000000u, FAR_INDEX = 0x3FFFFFFFu;

template <int H>
struct Strip {
    static constexpr int OC_MAX = 256 - 2 * H < 240 ? 256 - 2 * H : 240;           // owned columns of a strip
    static constexpr int W0 = OC_MAX + STRIP_PAD + 4;                             // pad quad + owned + garbage quad
    static constexpr int W = (W0 % 32 == 0 || W0 % 32 == 16) ? W0 + 4 : W0;       // words per tile row, rows 1 and 2 apart on different banks
    static constexpr int TILE_WORDS = STRIP_SLOTS * W;
    // fixed-point units per intensity scale: a cell collects at most (2H+1)^2 rays of less than 2 scales each
    static constexpr int FIX = H <= 8 ? 22 : 21;
    static_assert(H == 4 || H == 8 || H == 12, "reach of the tile path");
    static_assert(OC_MAX % 4 == 0 && W % 4 == 0 && W / 4 <= 64, "tile rows: whole quads, one flush pass of 64 threads per row");
    static_assert(2 * STRIP_U + 2 * H <= STRIP_SLOTS, "rows being flushed + rows being deposited must fit the circular tile");
    static_assert(((unsigned long long)(2 * H + 1) * (2 * H + 1) << (FIX + 1)) < (1ull << 32), "a cell cannot overflow");
};

// How one image is cut into blocks (host side; the same numbers go to the kernel).
struct StripPlan {
    int strips, oc;          // column strips and owned columns per strip (multiple of 4)
    int segs, seg_rows;      // row segments per strip and rows per segment
    unsigned far_cap;        // list entries per warp and beam (= the 32 columns x seg_rows pixels a warp may own)
};

inline StripPlan plan_strips(int nx, int ny, int H, int slots, int batch = 1) {
    StripPlan p;
    const int oc_max = 256 - 2 * H < 240 ? 256 - 2 * H : 240;
    p.strips = div_up(ny, oc_max);
    p.oc = (div_up(ny, p.strips) + 3) / 4 * 4;
    // one wave of equal blocks when that leaves segments of a useful length, else as many waves as it takes
    int segs = slots / (p.strips * batch);
    if (segs < 1) segs = 1;
    int rows = div_up(nx, segs);
    const int min_rows = 6 * H;                      // keeps the 2H halo rows of a segment at a quarter of its work
    if (rows < min_rows) rows = min_rows < nx ? min_rows : nx;
    p.segs = div_up(nx, rows);
    p.seg_rows = div_up(nx, p.segs);
    p.far_cap = (unsigned)p.seg_rows * 32u;
    return p;
}

// ---- fixed-point split of one ray, bit-compatible with refract_lean.cuh ------------------------------------
// (a * b + 2^31) >> 32: product of a fixed-point value and a 0.32 fraction, rounded to nearest
__device__ __forceinline__ unsigned strip_mulhi_rn(unsigned a, unsigned b, unsigned long long half) {
    return (unsigned)(((unsigned long long)a * b + half) >> 32);
}

// 2^31 as a 64-bit value the assembler cannot fold (`small` = any kernel argument below 2^30, e.g. an image dimension),
// so that it stays in one register pair for the whole kernel instead of being re-materialised for every ray
__device__ __forceinline__ unsigned long long strip_half(int small) { return 0x80000000ull | (unsigned long long)((unsigned)small >> 30); }

__device__ __forceinline__ void strip_reds4(unsigned a0, unsigned a1, unsigned w0, unsigned w1, unsigned w2, unsigned w3) {
    asm volatile(
        "red.shared.add.u32 [%0], %2;\n\t"
        "red.shared.add.u32 [%0+4], %3;\n\t"
        "red.shared.add.u32 [%1], %4;\n\t"
        "red.shared.add.u32 [%1+4], %5;"
        ::"r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
        : "memory");
}

// Named barriers between the depositing warps and the flushing warp (PTX producer / consumer pattern: the side that
// does not need to wait only ARRIVES).  FULL[b]: "chunk k (k & 1 == b) is deposited"; EMPTY[b]: "its finished rows are
// flushed and zeroed".
constexpr int BAR_FULL = 1, BAR_EMPTY = 3;
// (barrier ids as immediates: with a register id the block would reserve all 16 hardware barriers)
__device__ __forceinline__ void named_sync(int base, int odd) {
    if (base == BAR_FULL) { if (odd) asm volatile("bar.sync 2, %0;" ::"n"(STRIP_BLOCK) : "memory"); else asm volatile("bar.sync 1, %0;" ::"n"(STRIP_BLOCK) : "memory"); }
    else { if (odd) asm volatile("bar.sync 4, %0;" ::"n"(STRIP_BLOCK) : "memory"); else asm volatile("bar.sync 3, %0;" ::"n"(STRIP_BLOCK) : "memory"); }
}
__device__ __forceinline__ void named_arrive(int base, int odd) {
    if (base == BAR_FULL) { if (odd) asm volatile("bar.arrive 2, %0;" ::"n"(STRIP_BLOCK) : "memory"); else asm volatile("bar.arrive 1, %0;" ::"n"(STRIP_BLOCK) : "memory"); }
    else { if (odd) asm volatile("bar.arrive 4, %0;" ::"n"(STRIP_BLOCK) : "memory"); else asm volatile("bar.arrive 3, %0;" ::"n"(STRIP_BLOCK) : "memory"); }
}

// A streaming load whose position in the instruction stream is fixed (volatile asm keeps its order against the
// shared-memory atomics): the prefetch of source row i + U is issued right AFTER row i has been deposited, into the
// very registers row i occupied -- the compiler otherwise loads into fresh registers and copies them back, and the
// copy waits for the load ("register ring").
__device__ __forceinline__ void strip_prefetch(float& dst, const float* p) {
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dst) : "l"(p));
}

// Block-uniform description of the rows a source row may deposit into.
struct RowWin {
    unsigned sub;    // STRIP_MAGIC + lo: (bits(tx) - sub) < span  <=>  floor(dx) in [lo, lo + span)
    unsigned span;
    unsigned slot;   // i - STRIP_MAGIC: (bits(tx) + slot) & 31 = circular slot of row i + floor(dx)
};

template <int H>
__device__ __forceinline__ RowWin row_window(int i, int R0, int R1, int nx) {
    const int lo = max(max(-H, R0 - 1 - i), -i);
    const int hi = min(min(H - 1, R1 - 1 - i), nx - 2 - i);
    RowWin w;
    w.sub = STRIP_MAGIC + (unsigned)lo;
    w.span = (unsigned)max(hi - lo + 1, 0);
    w.slot = (unsigned)i - STRIP_MAGIC;
    return w;
}

template <int H>
__device__ __forceinline__ RowWin row_window_interior(int i) {
    RowWin w;
    w.sub = STRIP_MAGIC - (unsigned)H;
    w.span = 2u * H;
    w.slot = (unsigned)i - STRIP_MAGIC;
    return w;
}

// Per-thread description of the columns this thread's rays may deposit into.
struct ColWin {
    unsigned sub;     // STRIP_MAGIC + lo
    unsigned span;    // 0 for a thread without a source column
    unsigned addr;    // shared byte address of (slot 0, the source column itself), minus 4 * STRIP_MAGIC
};

template <int H>
__device__ __forceinline__ ColWin col_window(int j, int C0, int oc_plan, int ny, bool live, unsigned tile_s) {
    const int lo = max(max(-H, C0 - 1 - j), -j);
    const int hi = min(min(H - 1, C0 + oc_plan - 1 - j), ny - 2 - j);
    ColWin w;
    w.sub = STRIP_MAGIC + (unsigned)lo;
    w.span = live ? (unsigned)max(hi - lo + 1, 0) : 0u;
    w.addr = tile_s + (unsigned)(j - C0 + STRIP_PAD) * 4u - STRIP_MAGIC * 4u;
    asm volatile("" : "+r"(w.addr), "+r"(w.sub), "+r"(w.span));      // keep these in registers: one add per ray instead of re-deriving them
    return w;
}

// Is (i, j, D) a TILE ray as far as geometry goes (block-independent)?  bx, by = bit patterns of D + 1.5 * 2^23 rounded down.
template <int H>
__device__ __forceinline__ bool tile_class(int i, int j, unsigned bx, unsigned by, int nx, int ny) {
    const int kx = (int)(bx - STRIP_MAGIC), ky = (int)(by - STRIP_MAGIC);
    return (unsigned)(kx + H) < 2u * H && (unsigned)(ky + H) < 2u * H && (unsigned)(i + kx) < (unsigned)(nx - 1) &&
           (unsigned)(j + ky) < (unsigned)(ny - 1);
}

// The deposit of one ray into the circular tile(s).  Returns whether the ray went into the tile; bx / by come back for
// the caller's list test.
//   t = D + 1.5 * 2^23 rounded down carries floor(D) in its low mantissa bits (|D| < 2^22; NaN, Inf and anything
//   larger land far outside every window), and u = D + ((M + 1) - t) rounded towards zero is 1 + (D - floor D) in
//   one rounding, below 2 by construction: its mantissa is the bilinear fraction in 0.23 fixed point.
template <int W, bool TWIN>
__device__ __forceinline__ bool strip_deposit(const RowWin& rw, const ColWin& cw, float v, float dx, float dy, float scale,
                                              unsigned vmin_bits, unsigned vspan, unsigned twin_off, unsigned long long half,
                                              unsigned& bx, unsigned& by) {
    constexpr float M = 12582912.f;
    const float tx = __fadd_rd(dx, M), ty = __fadd_rd(dy, M);
    bx = __float_as_uint(tx); by = __float_as_uint(ty);
    const bool ok = (bx - rw.sub) < rw.span && (by - cw.sub) < cw.span && (__float_as_uint(v) - vmin_bits) < vspan;
    if (ok) {
        const unsigned fx = __float_as_uint(__fadd_rz(dx, (M + 1.f) - tx)) << 9, fy = __float_as_uint(__fadd_rz(dy, (M + 1.f) - ty)) << 9;
        // round(v * scale) for v * scale < 2^23 on the FMA pipe: 2^23 + x has an ulp of 1
        const unsigned V = __float_as_uint(fmaf(v, scale, 8388608.f)) - 0x4B000000u;
        const unsigned V1 = strip_mulhi_rn(V, fx, half), V0 = V - V1;
        const unsigned w1 = strip_mulhi_rn(V0, fy, half), w0 = V0 - w1;
        const unsigned w3 = strip_mulhi_rn(V1, fy, half), w2 = V1 - w3;
        const unsigned s0 = (bx + rw.slot) & (STRIP_SLOTS - 1), s1 = (bx + rw.slot + 1u) & (STRIP_SLOTS - 1);
        // TWIN: the ray goes to the tile at +0 and to the one at +twin_off; otherwise only to the tile at +twin_off
        const unsigned col = cw.addr + by * 4u + (TWIN ? 0u : twin_off);
        const unsigned a0 = col + s0 * (W * 4u), a1 = col + s1 * (W * 4u);
        strip_reds4(a0, a1, w0, w1, w2, w3);
        if (TWIN) strip_reds4(a0 + twin_off, a1 + twin_off, w0, w1, w2, w3);
    }
    return ok;
}

// The rays a warp could not give to any tile go on the WARP's private slice of the list: no atomics, the count lives
// in a register.  Called by all 32 lanes (`push` = this lane has an entry).
__device__ __forceinline__ void strip_push(uint4* slice, unsigned& count, bool push, unsigned head, float v, float dx, float dy) {
    const unsigned mask = __ballot_sync(FULL_MASK, push);
    if (push) slice[count + __popc(mask & ((1u << (threadIdx.x & 31)) - 1u))] =
        make_uint4(head, __float_as_uint(v), __float_as_uint(dx), __float_as_uint(dy));
    count += __popc(mask);
}

// In-line flush by the depositing threads themselves (splat_strip.cu): thread (tid / 64, tid % 64) handles quad tid % 64
// of row ra + tid / 64 of a pass of up to 4 rows.  Quad 0 is the pad + left garbage column, quads 1 .. oc/4 the owned
// columns, the next one the right garbage column; rows outside [R0, R1) are garbage rows: zeroed, never stored.
struct FlushLane {
    int rr;            // row of the pass
    int word;          // word offset of the quad in a tile row, or -1 for a thread beyond the row
    int col;           // first image column of the quad; < 0: not an owned quad
};

template <int W>
__device__ __forceinline__ FlushLane flush_lane(int C0, int oc) {
    FlushLane l;
    const int q = threadIdx.x & 63;
    l.rr = threadIdx.x >> 6;
    l.word = q < W / 4 ? 4 * q : -1;
    l.col = (q >= 1 && 4 * (q - 1) < oc && q < W / 4) ? C0 + 4 * (q - 1) : -1;
    return l;
}

template <int W, bool ACC>
__device__ __forceinline__ void strip_flush_rows(unsigned* tile, float* out, const FlushLane& l, int ra, int rb, int R0, int R1, int oc,
                                                 int C0, int ny, float inv_scale, bool vec) {
    const int r = ra + l.rr;
    if (r > rb || l.word < 0) return;
    uint4* p = reinterpret_cast<uint4*>(tile + (r & (STRIP_SLOTS - 1)) * W + l.word);
    const uint4 u = *p;
    *p = make_uint4(0u, 0u, 0u, 0u);
    if (r < R0 || r >= R1 || l.col < 0) return;
    float4 o = make_float4((float)u.x * inv_scale, (float)u.y * inv_scale, (float)u.z * inv_scale, (float)u.w * inv_scale);
    float* g = out + (size_t)r * ny + l.col;
    if (vec) {
        if (ACC) {
            const float4 b = *reinterpret_cast<const float4*>(g);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        *reinterpret_cast<float4*>(g) = o;
    } else {
        const float e[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (l.col - C0 + k < oc) g[k] = ACC ? g[k] + e[k] : e[k];
    }
}

// The flushing warp's loop over the chunks of a segment: wait until chunk k is deposited, write the rows no later source
// row can reach (plain 128-bit stores, or `out +=` in ACC mode), zero their slots, release them.  Two rows at a time, a
// lane handling the quads lane, lane + 32 of each and issuing all its loads first: four shared-memory (and, in ACC mode,
// four global) loads in flight per lane, which is what lets ONE warp keep up with the eight depositing ones.
// NT tiles (beams) side by side in shared memory, one image each.  Returns the integer sum of what this lane stored from
// tile NT-1.
template <int H, int W, int NT, bool ACC, bool SUM>
__device__ __forceinline__ unsigned long long strip_flusher(unsigned* tiles, float* const* outs, int R0, int R1, int C0, int oc, int nx, int ny,
                                                            float inv_scale, bool vec) {
    constexpr int U = STRIP_U, Q = W / 4, NQ = (Q + 31) / 32, RB = 2;
    const int lane = threadIdx.x & 31;
    const int s_begin = max(R0 - H, 0), s_end = min(R1 + H, nx);
    const int n_chunks = (s_end - s_begin + U - 1) / U;
    int flush_next = R0 - 1;
    unsigned long long sum = 0ull;
    for (int k = 0; k < n_chunks; ++k) {
        const int s = s_begin + k * U;
        named_sync(BAR_FULL, k & 1);
        const int final_row = k == n_chunks - 1 ? R1 - 1 : s + U - 1 - H;
        for (int r0 = flush_next; r0 <= final_row; r0 += RB) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                uint4 u[RB][NQ];
                float4 b[RB][NQ];
                bool mine[RB][NQ];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const int r = r0 + rr;
                    uint4* row = reinterpret_cast<uint4*>(tiles + t * (STRIP_SLOTS * W) + (r & (STRIP_SLOTS - 1)) * W);
                    const float4* g = reinterpret_cast<const float4*>(outs[t] + (size_t)r * ny + C0 - 4);   // quad q starts at column C0 + 4 (q - 1)
#pragma unroll
                    for (int e = 0; e < NQ; ++e) {
                        const int q = 32 * e + lane;
                        u[rr][e] = make_uint4(0u, 0u, 0u, 0u);
                        b[rr][e] = make_float4(0.f, 0.f, 0.f, 0.f);
                        // row R0 - 1 is the upper garbage row, quad 0 the pad + left garbage column, quads beyond oc the right
                        // garbage column: zeroed, never stored
                        mine[rr][e] = r <= final_row && r >= R0 && q >= 1 && 4 * (q - 1) < oc;
                        if (r <= final_row && q < Q) { u[rr][e] = row[q]; row[q] = make_uint4(0u, 0u, 0u, 0u); }
                        if (ACC && vec && mine[rr][e]) b[rr][e] = g[q];
                    }
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    float* g = outs[t] + (size_t)(r0 + rr) * ny + C0 - 4;
#pragma unroll
                    for (int e = 0; e < NQ; ++e) {
                        if (!mine[rr][e]) continue;
                        const int q = 32 * e + lane;
                        const uint4 w = u[rr][e];
                        float4 o = make_float4((float)w.x * inv_scale, (float)w.y * inv_scale, (float)w.z * inv_scale, (float)w.w * inv_scale);
                        if (vec) {
                            if (ACC) { o.x += b[rr][e].x; o.y += b[rr][e].y; o.z += b[rr][e].z; o.w += b[rr][e].w; }
                            *(reinterpret_cast<float4*>(g) + q) = o;
                            if (SUM && t == NT - 1) sum += (unsigned long long)w.x + w.y + w.z + w.w;
                        } else {
                            const float ov[4] = {o.x, o.y, o.z, o.w};
                            const unsigned wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                            for (int e4 = 0; e4 < 4; ++e4)
                                if (4 * (q - 1) + e4 < oc) {
                                    float* ge = g + 4 * q + e4;
                                    *ge = ACC ? *ge + ov[e4] : ov[e4];
                                    if (SUM && t == NT - 1) sum += wv[e4];
                                }
                        }
                    }
                }
            }
        }
        if (final_row >= flush_next) flush_next = final_row + 1;
        if (k + 2 < n_chunks) named_arrive(BAR_EMPTY, k & 1);      // nobody waits for the last two
    }
    return sum;
}

// Per-device cache of launch facts (resident blocks per SM x SMs): cudaFuncSetAttribute and the occupancy query are
// per device, so the cache is keyed by the current device (a process may drive several GPUs).
struct DeviceSlots {
    int slots[32] = {0};
    template <typename K>
    int get(K kernel, int threads, size_t smem, int* out) {
        int dev = 0;
        PARESIS_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 32) dev = 0;
        if (!slots[dev]) {
            PARESIS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0, sms = 0;
            PARESIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
            PARESIS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            slots[dev] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
        }
        *out = slots[dev];
        return PARESIS_OK;
    }
};

// Scratch for the ray lists (entries + one counter per warp): a per-device block cached inside the library, handed from
// stream to stream through an event (splat_strip.cu).
int strip_scratch_alloc(size_t bytes, void** ptr, cudaStream_t s);
int strip_scratch_free(void* ptr, cudaStream_t s);
int strip_scratch_trim();

}  // namespace paresis
