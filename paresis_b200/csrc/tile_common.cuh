// Helpers shared by the source-owner fixed-point TILE kernels (refract_lean.cuh, refract_tile_multi.cu, splat_tile.cu):
// tile layout constants, the per-ray deposit into a tile or straight to L2, the dense RED flush, the rows-per-tile choice.
//
// Why a tile: on a membrane the displacement field is torn every few pixels (sphere caps), so the REDs of a warp land in
// ~11 different 32-byte sectors per instruction, and L2 retires atomics per SECTOR (~170 sector-ops/ns on B200, measured:
// tools/redbench.cu, profiles/).  A block of 8 warps owns TR source rows x 256 source columns and bins its rays into a
// shared-memory tile that covers the source tile plus a halo of H pixels; the tile leaves the SM as dense 128-bit REDs.
// Why fixed point: sm_100a has no native fp32 add on shared memory (atomicAdd compiles to an ATOMS.CAST.SPIN loop), but
// 32-bit integer ATOMS.ADD is native and fire-and-forget (tools/smembench.cu: 2.1 vs 4.8-13.6 cycles per warp-op).
// (The first tile kernel that lived here, refract_tile_kernel, was superseded by refract_lean.cuh in round 1 and removed
// in round 2, once the lean kernel had its own oracle test: tests/test_gpu_kernels.py::test_refract_layers_vs_oracle.)
#pragma once
#include "refract_common.cuh"

namespace paresis {

constexpr int TILE_COLS = 256;   // 8 warps x 32 lanes

__device__ __forceinline__ void red_add4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// r, c of the lower cell and the four weights; `simple` = all four cells strictly inside the image
constexpr int FIX_BITS = 19;  // with 16 x 256 rays per tile: rays up to 2 x intensity_scale stay in fixed point

__device__ __forceinline__ float ex2_fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Block-uniform description of one output image: shared tile, global image, and the window of lower cells
// (r, c) whose four cells lie both inside the tile and strictly inside the image.
struct TileTarget {
    unsigned* tile;
    float* out;
    int r0, c0;            // first lower cell of the window (image coordinates)
    unsigned nr, nc;       // window extent: lower cells r0 .. r0+nr-1, c0 .. c0+nc-1
    int rlo, clo;          // image coordinates of tile cell (0, 0)
};

// One ray straight to L2 in fp32, cell by cell; cells outside the image are dropped, which is what zero-padding,
// scattering and cropping does (refractionFileNumba2.py:65-78; the |D| > N kill of :61-64 only removes rays that
// land outside anyway).  Returns what was deposited inside the image.  TWIN: the same ray goes to `out2` as well.
template <bool TWIN>
__device__ __forceinline__ float deposit_direct(float* out, float* out2, int i, int j, float v, float dx, float dy, int nx, int ny,
                                                bool& bad) {
    const float flx = floorf(dx), fly = floorf(dy);
    const float fx = dx - flx, fy = dy - fly;
    const int r = i + __float2int_rd(dx), c = j + __float2int_rd(dy);   // saturating
    const float v1 = v * fx, v0 = v - v1;
    const float w1 = v0 * fy, w0 = v0 - w1, w3 = v1 * fy, w2 = v1 - w3;
    bad |= !(fabsf(w0 + w3) <= 3.0e38f);
    const bool ra = (unsigned)r < (unsigned)nx, rb = (unsigned)(r + 1) < (unsigned)nx;
    const bool ca = (unsigned)c < (unsigned)ny, cb = (unsigned)(c + 1) < (unsigned)ny;
    const long long o = (long long)r * ny + c;
    float* p = out + o;
    float sum = 0.f;
    if (ra && ca && w0 != 0.f) { red_add(p, w0); sum += w0; }
    if (ra && cb && w1 != 0.f) { red_add(p + 1, w1); sum += w1; }
    if (rb && ca && w2 != 0.f) { red_add(p + ny, w2); sum += w2; }
    if (rb && cb && w3 != 0.f) { red_add(p + ny + 1, w3); sum += w3; }
    if (TWIN) {
        float* q = out2 + o;
        if (ra && ca && w0 != 0.f) red_add(q, w0);
        if (ra && cb && w1 != 0.f) red_add(q + 1, w1);
        if (rb && ca && w2 != 0.f) red_add(q + ny, w2);
        if (rb && cb && w3 != 0.f) red_add(q + ny + 1, w3);
    }
    return sum;
}

// One ray.  Fast path: fixed-point deposit into the tile (native ATOMS.ADD).  Everything else -- the ray leaves
// the tile, touches the image border, is too bright / negative / not finite -- goes to L2 in fp32, cell by cell;
// cells outside the image are dropped, which is what zero-padding, scattering and cropping does
// (refractionFileNumba2.py:65-78; the |D| > N kill of :61-64 only removes rays that land outside anyway).
// Returns what was deposited inside the image.
// TWIN: the same ray goes to a second image as well (`t2`: its tile sits `twin_off` words after the first).
template <int SC, bool TWIN>
__device__ __forceinline__ float deposit(const TileTarget& t, int i, int j, float v, float dx, float dy, int nx, int ny,
                                         float scale, unsigned vmax_bits, bool live, bool& bad, float* out2 = nullptr,
                                         int twin_off = 0) {
    const float flx = floorf(dx), fly = floorf(dy);
    const float fx = dx - flx, fy = dy - fly;
    const int r = i + __float2int_rd(dx), c = j + __float2int_rd(dy);   // saturating
    // v in [0, vmax) as one unsigned compare on the bit pattern (negative and NaN patterns are larger)
    const bool fast = live && (unsigned)(r - t.r0) < t.nr && (unsigned)(c - t.c0) < t.nc && __float_as_uint(v) < vmax_bits &&
                      fx + fy < 3.f;
    if (fast) {
        const float vs = v * scale;
        const float v1 = vs * fx, v0 = vs - v1;
        const float w1 = v0 * fy, w0 = v0 - w1, w3 = v1 * fy, w2 = v1 - w3;
        unsigned* p = t.tile + (r - t.rlo) * SC + (c - t.clo);
        const unsigned u0 = __float2uint_rn(w0), u1 = __float2uint_rn(w1), u2 = __float2uint_rn(w2), u3 = __float2uint_rn(w3);
        atomicAdd(p, u0);
        atomicAdd(p + 1, u1);
        atomicAdd(p + SC, u2);
        atomicAdd(p + SC + 1, u3);
        if (TWIN) {
            unsigned* q = p + twin_off;
            atomicAdd(q, u0);
            atomicAdd(q + 1, u1);
            atomicAdd(q + SC, u2);
            atomicAdd(q + SC + 1, u3);
        }
        return v;
    }
    return live ? deposit_direct<TWIN>(t.out, out2, i, j, v, dx, dy, nx, ny, bad) : 0.f;
}

// Tile -> image: dense 128-bit REDs, all-zero quads skipped; image borders and odd pitches fall back to scalars.
template <int SR, int SC>
__device__ __forceinline__ void flush_tile(const unsigned* tile, float* out, int rlo, int clo, int nx, int ny, float inv_scale,
                                           bool vec_ok, int nrows = SR) {
    constexpr int Q = SC / 4;
    for (int idx = threadIdx.x; idx < nrows * Q; idx += blockDim.x) {
        const int sr = idx / Q, q4 = idx - sr * Q;
        const int r = rlo + sr, c = clo + 4 * q4;
        if ((unsigned)r >= (unsigned)nx) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(tile + sr * SC + 4 * q4);
        if ((u.x | u.y | u.z | u.w) == 0u) continue;
        const float4 v = make_float4((float)u.x * inv_scale, (float)u.y * inv_scale, (float)u.z * inv_scale, (float)u.w * inv_scale);
        float* p = out + (size_t)r * ny + c;
        if (vec_ok && c >= 0 && c + 3 < ny) {
            red_add4(p, v);
        } else {
            if (u.x && (unsigned)c < (unsigned)ny) red_add(p, v.x);
            if (u.y && (unsigned)(c + 1) < (unsigned)ny) red_add(p + 1, v.y);
            if (u.z && (unsigned)(c + 2) < (unsigned)ny) red_add(p + 2, v.z);
            if (u.w && (unsigned)(c + 3) < (unsigned)ny) red_add(p + 3, v.w);
        }
    }
}

// Rows per block: all blocks cost the same, so a launch takes ceil(blocks / resident slots) rounds.  Pick the row count
// (<= TR, the size the tile was built for) whose last round is fullest, charging the fixed cost of a block (three
// extra thickness rows, halo flush) against short tiles.
inline int pick_tile_rows(int nx, int strips, int slots, int max_rows) {
    int best = max_rows;
    double best_score = -1.0;
    for (int rows = max_rows; rows >= (max_rows > 8 ? 6 : 4); --rows) {
        const long blocks = (long)strips * div_up(nx, rows);
        const long rounds = (blocks + slots - 1) / slots;
        const double score = (double)blocks / (double)(rounds * slots) * rows / (rows + 4.0);
        if (score > best_score + 1e-9) { best_score = score; best = rows; }
    }
    return best;
}

}  // namespace paresis
