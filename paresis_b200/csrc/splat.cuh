// Device-side ray deposit logic shared by the stand-alone splat kernel and the fused
// thickness -> transmission -> gradient -> displacement -> splat kernels.
//
// Semantics follow refractionFileNumba2.py:198-263 exactly (see oracle/oracle_loops.c for
// the scalar restatement this is tested against); the implementation is organised for the
// GPU: every lane of a warp owns one source column and walks down the rows, so
//   * deposits of neighbouring lanes that fall on neighbouring cells are merged with two
//     shuffles and leave the SM as ONE unit-stride REDG per row (warp-aggregated atomics),
//   * the lower-row half of each deposit is carried in a register and merged with the next
//     source row's upper-row half when they hit the same cell (the common case for a smooth
//     displacement field), so a ray costs ~1 REDG lane instead of 4.
// Anything irregular (frame borders, rays that tear apart) falls back to per-cell REDGs, so
// the result is the reference's for every input; only the fp32 summation order differs.
#pragma once
#include "common.cuh"

namespace paresis {

struct Frame {
    int nx, ny, margin;
};

// One ray in "lower cell" form: rows r, r+1 and columns c, c+1.
// w[0]=(r,c) w[1]=(r,c+1) w[2]=(r+1,c) w[3]=(r+1,c+1), already multiplied by the intensity.
// ok bit k set <=> cell k receives its deposit (reference frame rules AND inside the image).
struct Ray {
    int r, c;
    float w[4];
    unsigned ok;
};

__device__ __forceinline__ Ray empty_ray() {
    Ray q;
    q.r = 0; q.c = 0; q.ok = 0u;
    q.w[0] = q.w[1] = q.w[2] = q.w[3] = 0.f;
    return q;
}

__device__ __forceinline__ int floor_to_int_sat(float f) {
    // keeps i + floor(D) inside int32 for any D (Inf saturates, NaN -> 0)
    f = fminf(fmaxf(f, -1073741824.f), 1073741824.f);
    return __float2int_rd(f);
}

// (i, j): source pixel; v: intensity; (dx, dy): displacement in pixels along rows / columns.
__device__ __forceinline__ Ray make_ray(int i, int j, float v, float dx, float dy, const Frame& f) {
    Ray q;
    // :228-233 -- only |D| > 1 moves the base cell; the signed remainder picks the side
    int rb = i, cb = j;
    float fx = dx, fy = dy;
    if (fabsf(dx) > 1.f) { float fl = floorf(dx); rb += floor_to_int_sat(fl); fx = dx - fl; }
    if (fabsf(dy) > 1.f) { float fl = floorf(dy); cb += floor_to_int_sat(fl); fy = dy - fl; }
    const float ax = fabsf(fx), ay = fabsf(fy);
    const bool px = fx >= 0.f, py = fy >= 0.f;
    // :235-262 -- loop-frame rules (frame = image zero-padded by margin)
    const int fxn = f.nx + 2 * f.margin, fyn = f.ny + 2 * f.margin;
    const int rp = rb + f.margin, cp = cb + f.margin;
    const bool base_ok = (rp >= 0) & (rp < fxn) & (cp >= 0) & (cp < fyn);
    const bool row_ok = px ? (rp < fxn - 1) : (rp > 0);
    const bool col_ok = py ? (cp < fyn - 1) : (cp > 0);
    const bool nb_ok = base_ok & row_ok & col_ok;
    // lower-cell form: the base cell (:237) sits at row offset bi / column offset bj, the row
    // neighbour (:242/:255), column neighbour (:244/:249) and diagonal (:243/:248) fill the rest.
    // Products keep the reference's order (I * row factor) * column factor.
    q.r = px ? rb : rb - 1;
    q.c = py ? cb : cb - 1;
    const float rw0 = px ? 1.f - ax : ax, rw1 = px ? ax : 1.f - ax;
    const float cw0 = py ? 1.f - ay : ay, cw1 = py ? ay : 1.f - ay;
    const float v0 = v * rw0, v1 = v * rw1;
    q.w[0] = v0 * cw0; q.w[1] = v0 * cw1; q.w[2] = v1 * cw0; q.w[3] = v1 * cw1;
    const unsigned base_bit = 1u << (2 * (px ? 0 : 1) + (py ? 0 : 1));
    const unsigned ok = nb_ok ? 0xFu : (base_ok ? base_bit : 0u);
    // clip to the stored image (deposits into the virtual margin are the cropped ones, :78)
    const bool r0 = (q.r >= 0) & (q.r < f.nx), r1 = (q.r + 1 >= 0) & (q.r + 1 < f.nx);
    const bool c0 = (q.c >= 0) & (q.c < f.ny), c1 = (q.c + 1 >= 0) & (q.c + 1 < f.ny);
    unsigned inside = (r0 & c0 ? 1u : 0u) | (r0 & c1 ? 2u : 0u) | (r1 & c0 ? 4u : 0u) | (r1 & c1 ? 8u : 0u);
    q.ok = ok & inside;
    return q;
}

// The same ray by plain floor arithmetic: r = i + floor(dx), weights (1-f, f).  Numerically the
// reference's map for every displacement (|D| <= 1 included, see DESIGN.md "Splat semantics"); what
// it cannot express is the loop-frame edge rule of :235-262, so it is only used when all four cells
// lie strictly inside the image (`simple`), where that rule never fires.  Anything else goes
// through make_ray().
struct FastRay {
    int key;               // r * ny + c of the lower cell
    float w0, w1, w2, w3;  // (r,c) (r,c+1) (r+1,c) (r+1,c+1)
    bool simple;
};

__device__ __forceinline__ FastRay fast_ray(int i, int j, float v, float dx, float dy, int nx, int ny) {
    FastRay q;
    const float flx = floorf(dx), fly = floorf(dy);
    const float fx = dx - flx, fy = dy - fly;
    const int r = i + __float2int_rd(dx), c = j + __float2int_rd(dy);   // saturating; wrap-around fails `simple`
    q.simple = ((unsigned)r < (unsigned)(nx - 1)) & ((unsigned)c < (unsigned)(ny - 1));
    const float v1 = v * fx, v0 = v - v1;
    q.w1 = v0 * fy; q.w0 = v0 - q.w1;
    q.w3 = v1 * fy; q.w2 = v1 - q.w3;
    q.key = r * ny + c;
    return q;
}

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int NO_KEY = -(1 << 30);

// MODE 0: one REDG per non-zero deposit.  MODE 1: + lane merge.  MODE 2: + row carry.
template <int MODE>
struct Splatter {
    float* out;
    int ny;
    int lane;
    int carry_key;
    float carry_val;
    int* flag;
    bool bad;   // a non-finite deposit was seen; published once, in finish()

    __device__ __forceinline__ void init(float* out_, int ny_, int* flag_) {
        out = out_; ny = ny_; flag = flag_;
        lane = threadIdx.x & 31;
        carry_key = NO_KEY; carry_val = 0.f;
        bad = false;
    }

    __device__ __forceinline__ void cells(const Ray& q) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (((q.ok >> k) & 1u) && q.w[k] != 0.f)
                red_add(out + (size_t)(q.r + (k >> 1)) * ny + (q.c + (k & 1)), q.w[k]);
    }

    // Must be called by all 32 lanes together (lanes without a pixel pass empty_ray()).
    __device__ __forceinline__ void put(const Ray& q) {
        if (q.ok) {
            const float s = (q.w[0] + q.w[1]) + (q.w[2] + q.w[3]);
            bad |= !(fabsf(s) <= 3.0e38f);
        }
        if (MODE == 0) {
            cells(q);
            return;
        }
        const bool simple = q.ok == 0xFu;
        const int key = simple ? q.r * ny + q.c : NO_KEY;
        const int kprev = __shfl_up_sync(FULL_MASK, key, 1);
        const float w1p = __shfl_up_sync(FULL_MASK, q.w[1], 1);
        const float w3p = __shfl_up_sync(FULL_MASK, q.w[3], 1);
        const bool absorb = simple && lane > 0 && key == kprev + 1;   // left lane's right column is my left column
        const unsigned m = __ballot_sync(FULL_MASK, absorb);
        const bool given = ((m >> lane) >> 1) & 1u;                   // my right column was taken by lane+1
        float top = q.w[0], bot = q.w[2];
        if (absorb) { top += w1p; bot += w3p; }
        if (simple) {
            if (MODE == 2) {
                if (carry_key == key) top += carry_val;
                else if (carry_key != NO_KEY) red_add(out + carry_key, carry_val);
                if (bot != 0.f) { carry_key = key + ny; carry_val = bot; }
                else carry_key = NO_KEY;
            } else if (bot != 0.f) {
                red_add(out + key + ny, bot);
            }
            if (top != 0.f) red_add(out + key, top);
            if (!given) {
                if (q.w[1] != 0.f) red_add(out + key + 1, q.w[1]);
                if (q.w[3] != 0.f) red_add(out + key + ny + 1, q.w[3]);
            }
        } else {
            if (MODE == 2 && carry_key != NO_KEY) { red_add(out + carry_key, carry_val); carry_key = NO_KEY; }
            cells(q);
        }
    }

    // Fast path: every lane of the warp holds a `simple` ray (all four cells inside the image).
    __device__ __forceinline__ void put_simple(const FastRay& q) {
        bad |= !(fabsf(q.w0 + q.w3) <= 3.0e38f);
        if (MODE == 0) {
            float* p = out + q.key;
            if (q.w0 != 0.f) red_add(p, q.w0);
            if (q.w1 != 0.f) red_add(p + 1, q.w1);
            if (q.w2 != 0.f) red_add(p + ny, q.w2);
            if (q.w3 != 0.f) red_add(p + ny + 1, q.w3);
            return;
        }
        const int kprev = __shfl_up_sync(FULL_MASK, q.key, 1);
        const float w1p = __shfl_up_sync(FULL_MASK, q.w1, 1);
        const float w3p = __shfl_up_sync(FULL_MASK, q.w3, 1);
        const bool absorb = lane > 0 && q.key == kprev + 1;
        const unsigned m = __ballot_sync(FULL_MASK, absorb);
        const bool given = ((m >> lane) >> 1) & 1u;
        float top = q.w0, bot = q.w2;
        if (absorb) { top += w1p; bot += w3p; }
        float* p = out + q.key;
        if (MODE == 2) {
            if (carry_key == q.key) top += carry_val;
            else if (carry_key != NO_KEY) red_add(out + carry_key, carry_val);
            if (bot != 0.f) { carry_key = q.key + ny; carry_val = bot; }
            else carry_key = NO_KEY;
        } else if (bot != 0.f) {
            red_add(p + ny, bot);
        }
        if (top != 0.f) red_add(p, top);
        if (!given) {
            if (q.w1 != 0.f) red_add(p + 1, q.w1);
            if (q.w3 != 0.f) red_add(p + ny + 1, q.w3);
        }
    }

    __device__ __forceinline__ void finish() {
        if (MODE == 2 && carry_key != NO_KEY) { red_add(out + carry_key, carry_val); carry_key = NO_KEY; }
        if (bad && flag) atomicOr(flag, FLAG_NONFINITE);
    }
};

}  // namespace paresis
