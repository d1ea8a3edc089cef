// The fixed-point tile hop (source-owner tiles + dense RED flush; helpers in tile_common.cuh) -- the production hop.
// It spends 20-30 % fewer instructions per ray than the first tile kernel (what that bought, and what bounds the hops
// instead, is in DESIGN.md section 3 and profiles/r01_summary.md):
//
//  * floor and fraction of a displacement come from magic-number adds on the FMA pipe (see lean_deposit) instead of
//    FRND / F2I on the quarter-rate pipe; NaN, Inf and huge displacements fail the tile-window test by themselves.
//  * the ray is converted to fixed point once (V = round(v * S)) and split between its four cells with integer
//    multiply-high, rounded to nearest: V1 = (V * fx + 2^31) >> 32, V0 = V - V1, ... -- the four parts add up to V
//    exactly, whatever the order.
//  * both beams of a pixel are computed before any branch, so that their dependent chains interleave.
//  * row neighbours come from L1 (two more loads off the address of the row fetched one step earlier) instead of two
//    shuffles, two selects and a halo column per map.
//  * rays that cannot take the tile (outside its window, too bright, negative, not finite) are not handled where they
//    occur -- that made ~30 % of the warp-steps of a membrane walk through the long fp32 path for one or two lanes --
//    but pushed on a small shared-memory list and deposited densely once the block is through its rows.
//  * the first and last warp of an image row and the first / last block of rows run a second copy of the loop
//    with clamped loads and the edge rules of np.gradient; the interior copy has neither.
//  * the tile is zeroed and flushed only over the rows this block can reach, by a flat walk over its quads.
//
// The bilinear fractions are truncated to 23 bits and the weights rounded to one fixed-point unit; end-to-end images
// sit at the same distance from the reference as with the first tile kernel (profiles/r01_parity_distances.txt).
#pragma once
#include "tile_common.cuh"

namespace paresis {

__device__ __forceinline__ void reds4(unsigned addr, unsigned w0, unsigned w1, unsigned w2, unsigned w3, unsigned rowb) {
    asm volatile(
        "red.shared.add.u32 [%0], %1;\n\t"
        "red.shared.add.u32 [%0+4], %2;\n\t"
        "red.shared.add.u32 [%4], %3;\n\t"
        "red.shared.add.u32 [%4+4], %5;"
        ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(addr + rowb), "r"(w3)
        : "memory");
}

// (a * b + 2^31) >> 32: the product of a fixed-point value and a 0.32 fraction, rounded to nearest (one IMAD.HI with a 64-bit addend;
// a plain multiply-high would floor, i.e. move 2^-20 of every ray towards its lower cell)
__device__ __forceinline__ unsigned mulhi_rn(unsigned a, unsigned b) {
    return (unsigned)(((unsigned long long)a * b + 0x80000000ull) >> 32);
}

// Fast path of one ray; false = the ray has to go the long way.  The window of lower cells whose four cells lie
// inside the tile and strictly inside the image is nr x nc cells from `win_s` (shared address of its first cell);
// (trow, tcol): the source pixel in window coordinates.
// t = D + 1.5 * 2^23 rounded down carries floor(D) in its low mantissa bits (|D| < 2^22; anything else, NaN and Inf
// included, lands far outside the tile window), and u = D + ((M + 1) - t) rounded towards zero is 1 + (D - floor(D))
// in ONE rounding, always below 2: its mantissa is the bilinear fraction in 0.23 fixed point.
template <int SC, int SR, bool TWIN>
__device__ __forceinline__ bool lean_deposit(unsigned win_s, int trow, int tcol, unsigned nr, unsigned nc, float v, float dx, float dy,
                                             float scale, unsigned vmin_bits, unsigned vspan) {
    constexpr float M = 12582912.f;
    const float tx = __fadd_rd(dx, M), ty = __fadd_rd(dy, M);
    const int kx = (int)(__float_as_uint(tx) - 0x4B400000u) + trow, ky = (int)(__float_as_uint(ty) - 0x4B400000u) + tcol;
    // vmin <= v < vmax as one unsigned compare on the bit pattern (negative, NaN and zero patterns fall outside)
    const bool fast = (unsigned)kx < nr && (unsigned)ky < nc && (__float_as_uint(v) - vmin_bits) < vspan;
    // branch-free up to the atomics (a lane that fails computes garbage it does not use), so that the chains of the
    // two beams of a pixel interleave instead of waiting on each other
    const unsigned fx = __float_as_uint(__fadd_rz(dx, (M + 1.f) - tx)) << 9, fy = __float_as_uint(__fadd_rz(dy, (M + 1.f) - ty)) << 9;
    const unsigned V = __float2uint_rn(v * scale);
    const unsigned V1 = mulhi_rn(V, fx), V0 = V - V1;
    const unsigned w1 = mulhi_rn(V0, fy), w0 = V0 - w1;
    const unsigned w3 = mulhi_rn(V1, fy), w2 = V1 - w3;
    const unsigned addr = win_s + (unsigned)kx * (SC * 4u) + (unsigned)ky * 4u;
    if (fast) {
        reds4(addr, w0, w1, w2, w3, SC * 4u);
        if (TWIN) reds4(addr + SR * SC * 4u, w0, w1, w2, w3, SC * 4u);
    }
    return fast;
}

// Tile rows [sr0, sr1) -> image, for a tile whose columns all lie inside the image and 16-byte aligned rows:
// one flat walk over the quads (rows are contiguous in shared memory), row / column kept incrementally.
template <int SC>
__device__ __forceinline__ void flush_rows_fast(const unsigned* tile, float* out, int rlo, int clo, int sr0, int sr1, int ny,
                                                float inv_scale) {
    constexpr int Q = SC / 4, STEP_R = TILE_COLS / Q, STEP_Q = TILE_COLS - STEP_R * Q;
    int sr = sr0 + (int)threadIdx.x / Q, q4 = (int)threadIdx.x % Q;
    int goff = (rlo + sr) * ny + clo + 4 * q4;                 // nx * ny < 2^30 (host check)
    const uint4* t = reinterpret_cast<const uint4*>(tile) + sr * Q + q4;
    const int grow = STEP_R * ny + 4 * STEP_Q;
    while (sr < sr1) {
        const uint4 u = *t;
        if ((u.x | u.y | u.z | u.w) != 0u)
            red_add4(out + goff, make_float4((float)u.x * inv_scale, (float)u.y * inv_scale, (float)u.z * inv_scale, (float)u.w * inv_scale));
        t += TILE_COLS; sr += STEP_R; q4 += STEP_Q; goff += grow;
        if (q4 >= Q) { q4 -= Q; sr += 1; goff += ny - 4 * Q; }
    }
}

// resident blocks per SM the kernels are compiled for (register cap = 65536 / (256 * blocks)); -D overrides for A/B builds
#ifndef LEAN_MIN_BLOCKS_DUAL
#define LEAN_MIN_BLOCKS_DUAL 3      // 80 registers: no spills in the row loop; same-box A/B: +4.7 % on the job over 4 blocks x 64
#endif
#ifndef LEAN_MIN_BLOCKS_SINGLE
#define LEAN_MIN_BLOCKS_SINGLE 6
#endif
#define LEAN_MIN_BLOCKS(DUAL, NM) ((DUAL) ? LEAN_MIN_BLOCKS_DUAL : LEAN_MIN_BLOCKS_SINGLE)
// pixels a ray may move and still take the tile (the tile carries a halo of this many cells on every side); a multiple
// of 4 (vector flush), and 8 no longer leaves room for three blocks of the two-beam kernel: 4 is the only useful value
#ifndef LEAN_REACH
#define LEAN_REACH 4
#endif
// source rows a tile is built for (two-beam / one-beam kernels)
#ifndef LEAN_TR_DUAL
#define LEAN_TR_DUAL 16
#endif
#ifndef LEAN_TR_SINGLE
#define LEAN_TR_SINGLE 16
#endif
constexpr unsigned MISS_REF = 0x40000000u, MISS_TWIN = 0x80000000u;

template <int NM, bool DUAL, bool HAS_I, int TR, int MQ, bool ZB>
__global__ void __launch_bounds__(TILE_COLS, LEAN_MIN_BLOCKS(DUAL, NM))
refract_lean_kernel(const LeanArgs a) {
    // this block's membrane position; a single-position launch (ZB = false) addresses its pointers as plain kernel
    // parameters instead of through a register-indexed constant load per use
    const LeanItem& it = a.z[ZB ? blockIdx.z : 0];
    constexpr int H = LEAN_REACH;
    static_assert(H % 4 == 0, "the fast flush issues 128-bit REDs at column block * 256 - H: the reach must keep them aligned");
    constexpr int SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H, NT = DUAL ? 2 : 1;
    // a cell can receive every ray of the block: rays are admitted to the tile up to 2^32 / (TR * 256) units each (vmax_bits
    // below), brighter ones take the fp32 list.  TR = 16 admits 2 x intensity_scale, TR = 24 still 1.33 x.
    static_assert(((1ull << 32) / (TR * TILE_COLS)) >= (5ull << (FIX_BITS - 2)), "rays up to 1.25 x intensity_scale must fit the fixed-point tile");
    extern __shared__ __align__(16) unsigned tile_smem[];
    uint4* const queue = reinterpret_cast<uint4*>(tile_smem + NT * SR * SC);
    unsigned* const qcount = tile_smem + NT * SR * SC + 4 * MQ;

    const Frame f = a.f;
    const int tid = threadIdx.x, lane = tid & 31;
    const int j = blockIdx.x * TILE_COLS + tid;
    const int i0 = blockIdx.y * a.rows;          // a.rows <= TR: picked on the host to fill whole waves
    const int i1 = min(i0 + a.rows, f.nx);
    const bool live = j < f.ny;
    const int jc = live ? j : f.ny - 1;  // dead lanes read a valid address, contribute nothing
    const int jw0 = j - lane;
    // warp-uniform: every row and column the gradients of this warp read is interior (and row i+2 exists)
    const bool full_lean = jw0 >= 1 && jw0 + 31 <= f.ny - 2 && i0 >= 1 && i1 + 1 <= f.nx - 1;
    const bool inner_cols = __all_sync(FULL_MASK, live && j > 0 && j < f.ny - 1);

    const int used_rows = (i1 - i0) + 2 * H + 1;     // tile rows the rays of this block can reach (<= SR)
    {   // zero the tile(s) and the miss counter
        uint4* z = reinterpret_cast<uint4*>(tile_smem);
#pragma unroll
        for (int t = 0; t < NT; ++t)
            for (int k = tid; k < used_rows * (SC / 4); k += TILE_COLS) z[t * (SR * SC / 4) + k] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) *qcount = 0u;
    }

    // Row ring: slot k of step s holds row i-1+k, row i+2 is fetched now.  The row
    // neighbours (lr) of row i+1 are fetched now as well, from the lines the previous step brought into L1.
    constexpr int RING = 4;
    float row[RING][NM], inten[RING], lr[2][NM][2];
    int off = i0 * f.ny + jc;          // element offset of (i, jc); nx*ny < 2^30 (host check)
    const int last = (f.nx - 1) * f.ny + jc;   // offsets are clamped to the image instead of predicating the loads
#pragma unroll
    for (int k = 0; k < RING - 1; ++k) {
        const int o = min(max(off + (k - 1) * f.ny, jc), last);
#pragma unroll
        for (int m = 0; m < NM; ++m) row[k][m] = __ldg(it.map[m] + o);
        // plain loads: with clear_input the same thread stores to this address after reading it
        inten[k] = HAS_I ? it.I_in[o] : a.I_uniform;
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        lr[0][m][0] = __ldg(it.map[m] + off + (jc > 0 ? -1 : 0));
        lr[0][m][1] = __ldg(it.map[m] + off + (jc < f.ny - 1 ? 1 : 0));
    }

    // rays below vmax convert to less than 2^32 / (TR * 256) - 4 units: a whole tile cannot overflow one cell
    const float fix_scale = (float)(1u << FIX_BITS) / a.intensity_scale;
    const unsigned vmax_bits = __float_as_uint(a.intensity_scale * (float)((unsigned)((1ull << 32) / (TR * TILE_COLS)) - 8u) / (float)(1u << FIX_BITS));
    // rays below 2^9 units (2^-10 of the intensity scale) would be quantised to worse than 1e-3: they take the fp32 list
    const unsigned vmin_bits = __float_as_uint(a.intensity_scale * (1.0f / 1024.0f));
    const float neg_log2e = -1.4426950408889634f;
    // zero-fill of the buffers the next kernel scatters into: this block's rows x 256 columns, with 128-bit stores
    // when the layout allows it, else pixel by pixel in the row loop (the host fills unused slots with a used pointer)
    const bool zero_fill = it.zero[0] != nullptr;
    const bool zero_vec = zero_fill && (f.ny & 3) == 0 &&
                          ((reinterpret_cast<uintptr_t>(it.zero[0]) | reinterpret_cast<uintptr_t>(it.zero[1]) |
                            reinterpret_cast<uintptr_t>(it.zero[2])) & 15) == 0;
    if (zero_fill && !zero_vec && live) {
        for (int r = i0; r < i1; ++r) {
            const size_t o = (size_t)r * f.ny + j;
            it.zero[0][o] = 0.f; it.zero[1][o] = 0.f; it.zero[2][o] = 0.f;
        }
    }
    if (zero_vec) {
        const int q = tid & 63, c = blockIdx.x * TILE_COLS + 4 * q;
        if (c < f.ny) {
            for (int r = i0 + (tid >> 6); r < i1; r += TILE_COLS / 64) {
                const size_t o = (size_t)r * f.ny + c;
#pragma unroll
                for (int k = 0; k < 3; ++k) *reinterpret_cast<float4*>(it.zero[k] + o) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    const bool clear_in = HAS_I && a.clear_input;
    // window of lower cells (tile coordinates) whose four cells are tile cells and image cells
    const int rlo = i0 - H, clo = blockIdx.x * TILE_COLS - H;
    const int kx_lo = max(0, -rlo), ky_lo = max(0, -clo);
    const unsigned win_r = (unsigned)max(min(used_rows - 1, f.nx - 1 - rlo) - kx_lo, 0), win_c = (unsigned)max(min(SC - 1, f.ny - 1 - clo) - ky_lo, 0);
    const unsigned win_s = (unsigned)__cvta_generic_to_shared(tile_smem) + (unsigned)(kx_lo * SC + ky_lo) * 4u;
    const int tcol = tid + H - ky_lo;
    bool bad = false;
    float ref_sum = 0.f;
    if (it.zero_scalar && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *it.zero_scalar = 0.0;
    __syncthreads();

    // a ray that cannot take the tile: remember it (16 bytes), or deposit it now if the list is full
    auto miss = [&](unsigned flags, int i, float v, float dx, float dy) {
        if (v == 0.f) return;                          // nothing to deposit (and not worth a list entry)
        const unsigned slot = atomicAdd(qcount, 1u);
        if (slot < (unsigned)MQ) {
            queue[slot] = make_uint4(flags | ((unsigned)tid << 8) | (unsigned)(i - i0), __float_as_uint(v), __float_as_uint(dx),
                                     __float_as_uint(dy));
        } else if (flags & MISS_TWIN) {
            ref_sum += deposit_direct<true>(it.out_obj, it.out_ref, i, j, v, dx, dy, f.nx, f.ny, bad);
        } else {
            const float s = deposit_direct<false>((flags & MISS_REF) ? it.out_ref : it.out_obj, nullptr, i, j, v, dx, dy, f.nx, f.ny, bad);
            if (flags & MISS_REF) ref_sum += s;
        }
    };

    // the ray(s) of one source pixel: into the tile(s), or on the list.  `dead` = 0, or ~0u for a lane outside
    // the image (it must run the warp vote, and deposits nothing)
    auto emit = [&](int i, int trow, float vo, float vin, float dxo, float dyo, float dxr, float dyr, unsigned dead) {
        const unsigned vspan = (vmax_bits - vmin_bits) & ~dead;      // a dead lane fails the range test ...
        if (DUAL) {
            // outside the sample the two beams are the same ray: form it once, deposit it twice
            const bool same = dxo == dxr && dyo == dyr && vo == vin;
            if (__all_sync(FULL_MASK, same)) {
                if (lean_deposit<SC, SR, true>(win_s, trow, tcol, win_r, win_c, vo, dxo, dyo, fix_scale, vmin_bits, vspan)) ref_sum += vo;
                else if (!dead) miss(MISS_TWIN, i, vo, dxo, dyo);     // ... and is not a miss either
            } else {
                const bool fo = lean_deposit<SC, SR, false>(win_s, trow, tcol, win_r, win_c, vo, dxo, dyo, fix_scale, vmin_bits, vspan);
                const bool fr = lean_deposit<SC, SR, false>(win_s + SR * SC * 4u, trow, tcol, win_r, win_c, vin, dxr, dyr, fix_scale, vmin_bits, vspan);
                if (fr) ref_sum += vin;
                if (!(fo && fr) && !dead) {
                    if (!fo) miss(0u, i, vo, dxo, dyo);
                    if (!fr) miss(MISS_REF, i, vin, dxr, dyr);
                }
            }
        } else {
            if (!lean_deposit<SC, SR, false>(win_s, trow, tcol, win_r, win_c, vo, dxo, dyo, fix_scale, vmin_bits, vspan) && !dead) miss(0u, i, vo, dxo, dyo);
        }
    };

    if (full_lean) {
        // no clamps, no edge rules, deposits go to the tile
        int off2 = off + 2 * f.ny;       // (i+2, j)
        const float* pn[NM];             // (i+1, j) of each map: the address the previous step fetched from
#pragma unroll
        for (int m = 0; m < NM; ++m) pn[m] = it.map[m] + (off + f.ny);
        int trow = H - kx_lo;            // the source pixel in window coordinates: (trow, tcol)
        for (int ib = i0; ib < i1; ib += RING) {
#pragma unroll
            for (int s = 0; s < RING; ++s) {
                if (ib + s >= i1) break;                             // warp-uniform
                const int kup = s % RING, kmid = (s + 1) % RING, kdn = (s + 2) % RING, knew = (s + 3) % RING;
                const int lcur = s % 2, lnew = (s + 1) % 2;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    const float* p2 = it.map[m] + off2;
                    row[knew][m] = __ldg(p2);
                    lr[lnew][m][0] = __ldg(pn[m] - 1);
                    lr[lnew][m][1] = __ldg(pn[m] + 1);
                    pn[m] = p2;
                }
                inten[knew] = HAS_I ? it.I_in[off2] : a.I_uniform;
                float dxo = 0.f, dyo = 0.f, dxr = 0.f, dyr = 0.f, arg = 0.f;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    const float gx = row[kdn][m] - row[kup][m], gy = lr[lcur][m][1] - lr[lcur][m][0];
                    dxo = fmaf(a.g_obj[m], gx, dxo);
                    dyo = fmaf(a.g_obj[m], gy, dyo);
                    if (DUAL) {
                        dxr = fmaf(a.g_ref[m], gx, dxr);
                        dyr = fmaf(a.g_ref[m], gy, dyr);
                    }
                    arg = fmaf(a.att[m], row[kmid][m], arg);
                }
                const float vin = inten[kmid];
                // Sample.py:347; ex2.approx keeps ~2e-7 relative accuracy over the attenuation range
                const float vo = vin * ex2_fast(arg * neg_log2e);
                if (clear_in) const_cast<float*>(it.I_in)[off2 - 2 * f.ny] = 0.f;
                emit(ib + s, trow, vo, vin, dxo, dyo, dxr, dyr, 0u);
                off2 += f.ny;
                ++trow;
            }
        }
    } else {
        // border warps, first and last rows: clamped loads, edge rules of np.gradient
        int trow = H - kx_lo;
        const int dl = jc > 0 ? -1 : 0, dr = jc < f.ny - 1 ? 1 : 0;
        int onext = min(off + f.ny, last);
        for (int ib = i0; ib < i1; ib += RING) {
#pragma unroll
            for (int s = 0; s < RING; ++s) {
                const int i = ib + s;
                if (i >= i1) break;                                  // warp-uniform
                const int kup = s % RING, kmid = (s + 1) % RING, kdn = (s + 2) % RING, knew = (s + 3) % RING;
                const int lcur = s % 2, lnew = (s + 1) % 2;
                {
                    const int o1 = onext;                             // (i+1, jc), clamped
                    const int o2 = min(off + 2 * f.ny, last);         // (i+2, jc), clamped
                    onext = o2;
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        row[knew][m] = __ldg(it.map[m] + o2);
                        lr[lnew][m][0] = __ldg(it.map[m] + o1 + dl);
                        lr[lnew][m][1] = __ldg(it.map[m] + o1 + dr);
                    }
                    inten[knew] = HAS_I ? it.I_in[o2] : a.I_uniform;
                }
                const bool inner = inner_cols && i > 0 && i < f.nx - 1;   // warp-uniform
                float dxo = 0.f, dyo = 0.f, dxr = 0.f, dyr = 0.f, arg = 0.f;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    const float* t = it.map[m];
                    const float mid = row[kmid][m], up = row[kup][m], dn = row[kdn][m];
                    const float lf = lr[lcur][m][0], rt = lr[lcur][m][1];
                    float gy, gx;
                    if (inner) {
                        gy = rt - lf;
                        gx = dn - up;
                    } else {
                        // np.gradient(edge_order=2) numerators times 2h (refractionFileNumba2.py:54).  The third sample
                        // of a one-sided difference is two lanes away in this warp's own row (a load here would sit at L2
                        // latency on every row of every border warp, and the block waits for its slowest warp)
                        const float* r = t + (size_t)i * f.ny;
                        const float two_up = __shfl_sync(FULL_MASK, mid, 2), two_down = __shfl_sync(FULL_MASK, mid, (lane + 30) & 31);
                        if (jc == 0) gy = -3.f * mid + 4.f * rt - (f.ny > 2 ? two_up : __ldg(r + 2));
                        else if (jc == f.ny - 1) gy = 3.f * mid - 4.f * lf + (lane >= 2 && live ? two_down : __ldg(r + f.ny - 3));
                        else gy = rt - lf;
                        if (i == 0) gx = -3.f * mid + 4.f * dn - (f.nx > 2 ? row[knew][m] : __ldg(t + (size_t)2 * f.ny + jc));
                        else if (i == f.nx - 1) gx = 3.f * mid - 4.f * up + __ldg(t + (size_t)(f.nx - 3) * f.ny + jc);
                        else gx = dn - up;
                    }
                    dxo = fmaf(a.g_obj[m], gx, dxo);
                    dyo = fmaf(a.g_obj[m], gy, dyo);
                    if (DUAL) {
                        dxr = fmaf(a.g_ref[m], gx, dxr);
                        dyr = fmaf(a.g_ref[m], gy, dyr);
                    }
                    arg = fmaf(a.att[m], mid, arg);
                }
                const float vin = inten[kmid];
                const float vo = vin * ex2_fast(arg * neg_log2e);
                if (live && clear_in) const_cast<float*>(it.I_in)[off] = 0.f;
                emit(i, trow, vo, vin, dxo, dyo, dxr, dyr, live ? 0u : ~0u);
                off += f.ny;
                ++trow;
            }
        }
    }
    __syncthreads();
    {   // the rays that could not take the tile, densely
        const unsigned nq = min(*qcount, (unsigned)MQ);
        for (unsigned k = tid; k < nq; k += TILE_COLS) {
            const uint4 e = queue[k];
            const int i = i0 + (int)(e.x & 0xFFu), jj = blockIdx.x * TILE_COLS + (int)((e.x >> 8) & 0xFFu);
            const float v = __uint_as_float(e.y), dx = __uint_as_float(e.z), dy = __uint_as_float(e.w);
            if (DUAL && (e.x & MISS_TWIN)) {
                ref_sum += deposit_direct<true>(it.out_obj, it.out_ref, i, jj, v, dx, dy, f.nx, f.ny, bad);
            } else {
                const bool to_ref = DUAL && (e.x & MISS_REF);
                const float s = deposit_direct<false>(to_ref ? it.out_ref : it.out_obj, nullptr, i, jj, v, dx, dy, f.nx, f.ny, bad);
                if (to_ref) ref_sum += s;
            }
        }
    }
    if (bad && a.flag) atomicOr(a.flag, FLAG_NONFINITE);
    if (DUAL && it.sum_ref) {   // one double atomic per warp
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ref_sum += __shfl_xor_sync(FULL_MASK, ref_sum, d);
        if (lane == 0) atomicAdd(it.sum_ref, (double)ref_sum);
    }
    const bool vec_ok = (f.ny & 3) == 0;
    const float inv_scale = a.intensity_scale / (float)(1u << FIX_BITS);
    const bool cols_inside = clo >= 0 && clo + SC <= f.ny;                       // block-uniform
    const int sr0 = max(0, -rlo), sr1 = min(used_rows, f.nx - rlo);              // tile rows that are image rows
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        float* out = k == 0 ? it.out_obj : it.out_ref;
        const bool vec = vec_ok && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
        if (vec && cols_inside) flush_rows_fast<SC>(tile_smem + k * SR * SC, out, rlo, clo, sr0, sr1, f.ny, inv_scale);
        else flush_tile<SR, SC>(tile_smem + k * SR * SC, out, rlo, clo, f.nx, f.ny, inv_scale, vec, used_rows);
    }
}

template <int NM, bool DUAL, bool HAS_I, int TR>
static int launch_refract_lean(const LeanArgs& a_in, int n_batch, cudaStream_t s) {
    constexpr int H = LEAN_REACH, MQ = 256;   // two tiles + list of the two-beam hop: 58 KB
    constexpr int SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H;
    constexpr size_t smem = sizeof(unsigned) * (SR * SC * (DUAL ? 2 : 1) + 4 * MQ + 4);
    static int slots_of[32] = {0};   // resident blocks x SMs, per device (the attribute and the occupancy are per device)
    int dev = 0;
    PARESIS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 32) dev = 0;
    if (!slots_of[dev]) {
        PARESIS_CUDA(cudaFuncSetAttribute(refract_lean_kernel<NM, DUAL, HAS_I, TR, MQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PARESIS_CUDA(cudaFuncSetAttribute(refract_lean_kernel<NM, DUAL, HAS_I, TR, MQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        PARESIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refract_lean_kernel<NM, DUAL, HAS_I, TR, MQ, false>, TILE_COLS, smem));
        PARESIS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        slots_of[dev] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    LeanArgs a = a_in;
    const int strips = div_up(a.f.ny, TILE_COLS);
    // rows per block: the wave-filling count for a launch that runs alone, the whole tile when launches overlap anyway
    a.rows = g_lean_rows_override > 0 ? (g_lean_rows_override < TR ? g_lean_rows_override : TR)
                                      : (a.full_tiles ? TR : pick_tile_rows(a.f.nx, strips * n_batch, slots_of[dev], TR));
    dim3 grid(strips, div_up(a.f.nx, a.rows), n_batch);
    if (n_batch > 1) refract_lean_kernel<NM, DUAL, HAS_I, TR, MQ, true><<<grid, TILE_COLS, smem, s>>>(a);
    else refract_lean_kernel<NM, DUAL, HAS_I, TR, MQ, false><<<grid, TILE_COLS, smem, s>>>(a);
    PARESIS_LAUNCH_CHECK("refract_lean_kernel");
    return PARESIS_OK;
}

}  // namespace paresis
