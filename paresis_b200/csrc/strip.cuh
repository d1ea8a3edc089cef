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
// Fixed point: unit = 2^-22 of the intensity scale; a cell collects at most (2H+1)^2 rays of < 2^23 units, so it
// cannot overflow; rays below 2^9 units (2^-13 of the scale) go on the list instead of being quantised coarsely.
#pragma once
#include "splat.cuh"

namespace paresis {

constexpr int STRIP_THREADS = 256;   // source columns per block
constexpr int STRIP_U = 4;           // source rows per step (one barrier per step)
constexpr int STRIP_SLOTS = 32;      // rows of the circular tile
constexpr int STRIP_PAD = 4;         // owned column 0 sits at word 4 of a tile row; word 3 is the left garbage column
constexpr int STRIP_FIX = 22;        // fixed-point units per intensity scale = 2^22
constexpr unsigned STRIP_MAGIC = 0x4B400000u;   // bit pattern of 1.5 * 2^23
constexpr unsigned FAR_REF = 0x40000000u, FAR_TWIN = 0x80000000u, FAR_INDEX = 0x3FFFFFFFu;

template <int H>
struct Strip {
    static constexpr int OC_MAX = 240;                                   // owned columns of a strip (<= 256 - 2H)
    static constexpr int W = OC_MAX + STRIP_PAD + 4;                     // words per tile row: pad quad + owned + garbage quad
    static constexpr int TILE_WORDS = STRIP_SLOTS * W;
    static_assert(H == 4 || H == 8, "halo: 4 or 8 (16-byte aligned strips, 256 source columns per block)");
    static_assert(W % 4 == 0 && W % 32 != 0 && W / 4 <= 64, "tile rows: quads, no bank alignment, one flush pass");
    static_assert(3 * STRIP_U + 2 * H <= STRIP_SLOTS, "deposit window + rows in flight to memory must fit the circular tile");
    static_assert(((unsigned long long)(2 * H + 1) * (2 * H + 1) << (STRIP_FIX + 1)) < (1ull << 32), "a cell cannot overflow");
};

// How one image is cut into blocks (host side; the same numbers go to the kernel).
struct StripPlan {
    int strips, oc;          // column strips and owned columns per strip (multiple of 4)
    int segs, seg_rows;      // row segments per strip and rows per segment
    unsigned far_cap;        // list entries per block and beam (= owned pixels of a block)
};

inline StripPlan plan_strips(int nx, int ny, int H, int slots, int batch = 1) {
    StripPlan p;
    const int oc_max = 240;
    p.strips = div_up(ny, oc_max);
    p.oc = (div_up(ny, p.strips) + 3) / 4 * 4;
    // one wave of equal blocks when that leaves segments of a useful length, else as many waves as it takes
    int segs = slots / (p.strips * batch);
    if (segs < 1) segs = 1;
    int rows = div_up(nx, segs);
    const int min_rows = 6 * H;                      // keeps the 2H halo rows of a segment below a third of its work
    if (rows < min_rows) rows = min_rows < nx ? min_rows : nx;
    p.segs = div_up(nx, rows);
    p.seg_rows = div_up(nx, p.segs);
    p.far_cap = (unsigned)p.seg_rows * (unsigned)p.oc;
    return p;
}

// ---- fixed-point split of one ray, bit-compatible with refract_lean.cuh ------------------------------------
// (a * b + 2^31) >> 32: product of a fixed-point value and a 0.32 fraction, rounded to nearest
__device__ __forceinline__ unsigned strip_mulhi_rn(unsigned a, unsigned b, unsigned long long half) {
    return (unsigned)(((unsigned long long)a * b + half) >> 32);
}

// 2^31 in a register pair the compiler cannot see through (it would re-materialise the constant for every ray)
__device__ __forceinline__ unsigned long long strip_half() {
    unsigned long long h = 0x80000000ull;
    asm volatile("" : "+l"(h));
    return h;
}

__device__ __forceinline__ void strip_reds4(unsigned a0, unsigned a1, unsigned w0, unsigned w1, unsigned w2, unsigned w3) {
    asm volatile(
        "red.shared.add.u32 [%0], %2;\n\t"
        "red.shared.add.u32 [%0+4], %3;\n\t"
        "red.shared.add.u32 [%1], %4;\n\t"
        "red.shared.add.u32 [%1+4], %5;"
        ::"r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
        : "memory");
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

// Is (i, j, D) a TILE ray as far as geometry goes (block-independent)?  kx, ky = floor(D).
template <int H>
__device__ __forceinline__ bool tile_class(int i, int j, int kx, int ky, int nx, int ny) {
    return (unsigned)(kx + H) < 2u * H && (unsigned)(ky + H) < 2u * H && (unsigned)(i + kx) < (unsigned)(nx - 1) &&
           (unsigned)(j + ky) < (unsigned)(ny - 1);
}

// The deposit of one ray into the circular tile(s).  Everything up to the atomics is branch-free, so that the chains
// of consecutive rays interleave.  Returns whether the ray went into the tile.
//   t = D + 1.5 * 2^23 rounded down carries floor(D) in its low mantissa bits (|D| < 2^22; NaN, Inf and anything
//   larger land far outside every window), and u = D + ((M + 1) - t) rounded towards zero is 1 + (D - floor D) in
//   one rounding, below 2 by construction: its mantissa is the bilinear fraction in 0.23 fixed point.
template <int W, bool TWIN>
__device__ __forceinline__ bool strip_deposit(const RowWin& rw, const ColWin& cw, float v, float dx, float dy, float scale,
                                              unsigned vmin_bits, unsigned vspan, unsigned twin_off, unsigned long long half) {
    constexpr float M = 12582912.f;
    const float tx = __fadd_rd(dx, M), ty = __fadd_rd(dy, M);
    const unsigned bx = __float_as_uint(tx), by = __float_as_uint(ty);
    const bool ok = (bx - rw.sub) < rw.span && (by - cw.sub) < cw.span && (__float_as_uint(v) - vmin_bits) < vspan;
    const unsigned fx = __float_as_uint(__fadd_rz(dx, (M + 1.f) - tx)) << 9, fy = __float_as_uint(__fadd_rz(dy, (M + 1.f) - ty)) << 9;
    const unsigned V = __float2uint_rn(v * scale);
    const unsigned V1 = strip_mulhi_rn(V, fx, half), V0 = V - V1;
    const unsigned w1 = strip_mulhi_rn(V0, fy, half), w0 = V0 - w1;
    const unsigned w3 = strip_mulhi_rn(V1, fy, half), w2 = V1 - w3;
    const unsigned s0 = (bx + rw.slot) & (STRIP_SLOTS - 1), s1 = (bx + rw.slot + 1u) & (STRIP_SLOTS - 1);
    const unsigned col = cw.addr + by * 4u;
    const unsigned a0 = col + s0 * (W * 4u), a1 = col + s1 * (W * 4u);
    if (ok) {
        strip_reds4(a0, a1, w0, w1, w2, w3);
        if (TWIN) strip_reds4(a0 + twin_off, a1 + twin_off, w0, w1, w2, w3);
    }
    return ok;
}

// A ray for the list: 16 bytes on this block's slice.
__device__ __forceinline__ void strip_push(uint4* slice, unsigned* counter, unsigned flags, int index, float v, float dx, float dy) {
    const unsigned k = atomicAdd(counter, 1u);
    slice[k] = make_uint4(flags | (unsigned)index, __float_as_uint(v), __float_as_uint(dx), __float_as_uint(dy));
}

// Rows [ra, rb] (at most 4) of the circular tile -> image, and their slots back to zero.  Thread q of a 64-thread
// group handles words 4q .. 4q+3 of its row: quad 0 is the pad + left garbage column, quads 1 .. oc/4 the owned
// columns, the next one the right garbage column.  `store` rows are written, the others (garbage rows) only zeroed.
// Returns the integer sum of what this thread stored (for the running sum of a beam).
template <int W, bool ACC>
__device__ __forceinline__ unsigned long long strip_flush(unsigned* tile, float* out, int ra, int rb, int R0, int R1, int C0, int oc,
                                                          int ny, float inv_scale, bool vec) {
    const int tid = threadIdx.x;
    const int r = ra + (tid >> 6), q = tid & 63;
    unsigned long long sum = 0ull;
    if (r <= rb && q < W / 4) {
        uint4* p = reinterpret_cast<uint4*>(tile + (r & (STRIP_SLOTS - 1)) * W) + q;
        const uint4 u = *p;
        *p = make_uint4(0u, 0u, 0u, 0u);
        const int c = 4 * (q - 1);                     // first owned column of this quad, relative to C0
        if (r >= R0 && r < R1 && q >= 1 && c < oc) {
            float4 o = make_float4((float)u.x * inv_scale, (float)u.y * inv_scale, (float)u.z * inv_scale, (float)u.w * inv_scale);
            float* g = out + (size_t)r * ny + C0 + c;
            if (vec) {
                if (ACC) {
                    const float4 b = *reinterpret_cast<const float4*>(g);
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                *reinterpret_cast<float4*>(g) = o;
                sum = (unsigned long long)u.x + u.y + u.z + u.w;
            } else {
                const float e[4] = {o.x, o.y, o.z, o.w};
                const unsigned ue[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c + k < oc) { g[k] = ACC ? g[k] + e[k] : e[k]; sum += ue[k]; }
            }
        }
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

// Stream-ordered scratch for the ray lists (entries + one counter per block).  The library owns no long-lived
// buffers for this: the pool keeps freed blocks, so after the first call this is a pointer bump.
int strip_scratch_alloc(size_t bytes, void** ptr, cudaStream_t s);
int strip_scratch_free(void* ptr, cudaStream_t s);

}  // namespace paresis
