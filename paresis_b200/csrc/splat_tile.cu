// Stand-alone splat (fastloopNumba: I, Dx, Dy -> out +=) through fixed-point shared-memory tiles: variant 3 of
// paresis_splat.  Reference: refractionFileNumba2.py:198-263 (same body: refractionFileNumba.py:70-135).
//
// Variants 0-2 (refraction.cu) send every deposit to L2 as a RED; on a membrane's displacement field, torn every few
// pixels, a warp's REDs land in ~11 sectors and L2 retires atomics per sector (profiles/r01_summary.md): 0.32 of the
// HBM roofline where a smooth field gets 0.74.  Here a block of 8 warps owns up to 16 source rows x 256 columns, bins
// its rays into a tile covering the source tile plus a 4-pixel halo with native 32-bit integer shared-memory atomics,
// and the tile leaves the SM as dense 128-bit REDs (refract_lean.cuh has the deposit and the flush).
//
// The API carries no intensity scale, so each block takes its own from a quarter of its rows: twice the mean of
// their finite positive intensities = 1.x * 2^e gives the unit 2^(e - 19), rays below 2^(e+1) take the tile, and
// 4096 rays of less than 2^20 units cannot overflow a 32-bit cell.  Rays that
// are brighter, leave the tile window, touch the image border, are negative or not finite go through make_ray() --
// the reference's loop-frame rules -- straight to L2, from a list, once the block is through its rows.
#include "refract_lean.cuh"

namespace paresis {

template <int TR, int MQ>
__global__ void __launch_bounds__(TILE_COLS, 6)
splat_tile_kernel(const float* __restrict__ I, const float* __restrict__ Dx, const float* __restrict__ Dy, float* __restrict__ out,
                  Frame f, int rows, int* flag) {
    constexpr int H = 4, SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H, U = 4;
    static_assert((unsigned long long)TR * TILE_COLS * ((1u << 20) - 1u) < (1ull << 32), "a tile of full-scale rays must fit a 32-bit cell");
    extern __shared__ __align__(16) unsigned tile_smem[];
    uint4* const queue = reinterpret_cast<uint4*>(tile_smem + SR * SC);
    unsigned* const qcount = tile_smem + SR * SC + 4 * MQ;
    float* const wsum = reinterpret_cast<float*>(qcount + 4);   // one slot per warp

    const int tid = threadIdx.x, lane = tid & 31;
    const int j = blockIdx.x * TILE_COLS + tid;
    const int i0 = blockIdx.y * rows, i1 = min(i0 + rows, f.nx);
    const bool live = j < f.ny;
    const int jc = live ? j : f.ny - 1;
    const int used_rows = (i1 - i0) + 2 * H + 1;
    {
        uint4* z = reinterpret_cast<uint4*>(tile_smem);
        for (int k = tid; k < used_rows * (SC / 4); k += TILE_COLS) z[k] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) *qcount = 0u;
    }
    const int nrows = i1 - i0;
    const int col0 = i0 * f.ny + jc, last = col0 + (nrows - 1) * f.ny;      // nx * ny < 2^30 (host check)

    // the scale of the tile from a quarter of its rows, and the first rows of the walk
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float t = __ldg(I + min(col0 + k * (TR / 4) * f.ny, last));
        if (live && t > 0.f && t < 3.0e38f) m += t;               // NaN fails the compares
    }
    float vq[U], dxq[U], dyq[U];
    int nxt = col0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        vq[u] = __ldg(I + nxt);
        dxq[u] = __ldg(Dx + nxt);
        dyq[u] = __ldg(Dy + nxt);
        nxt = min(nxt + f.ny, last);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m += __shfl_xor_sync(FULL_MASK, m, d);
    if (lane == 0) wsum[tid >> 5] = m;
    __syncthreads();
    m = 0.f;
#pragma unroll
    for (int w = 0; w < TILE_COLS / 32; ++w) m += wsum[w];
    // twice the mean = 1.x * 2^e sets a power-of-two unit 2^(e - 19) (the scaling itself is exact); rays below
    // 2^(e+1) convert to less than 2^20 units.  (Mean below 2^-107: no fixed-point path, every ray goes the long way.)
    m = 2.f * m / (float)(4 * min(TILE_COLS, f.ny - blockIdx.x * TILE_COLS));
    const unsigned mexp = __float_as_uint(m) >> 23;
    const bool fixed_ok = mexp >= 20u && mexp < 254u;
    const float scale = __uint_as_float((273u - mexp) << 23), inv_scale = __uint_as_float((mexp - 19u) << 23);
    const unsigned vmax_bits = fixed_ok && live ? (mexp + 1u) << 23 : 0u;   // 0 <= v < 2^(e+1) as one unsigned compare

    // window of lower cells (tile coordinates) whose four cells are tile cells and image cells (there the plain
    // floor form is the reference's map, splat.cuh: fast_ray)
    const int rlo = i0 - H, clo = blockIdx.x * TILE_COLS - H;
    const int kx_lo = max(0, -rlo), ky_lo = max(0, -clo);
    const unsigned win_r = (unsigned)max(min(used_rows - 1, f.nx - 1 - rlo) - kx_lo, 0), win_c = (unsigned)max(min(SC - 1, f.ny - 1 - clo) - ky_lo, 0);
    const unsigned win_s = (unsigned)__cvta_generic_to_shared(tile_smem) + (unsigned)(kx_lo * SC + ky_lo) * 4u;
    const int tcol = tid + H - ky_lo, trow0 = H - kx_lo;

    Splatter<0> slow;
    slow.init(out, f.ny, flag);
    auto long_way = [&](int i, int jj, float vv, float dx, float dy) {
        const Ray q = make_ray(i, jj, vv, dx, dy, f);
        if (q.ok) {
            const float s = (q.w[0] + q.w[1]) + (q.w[2] + q.w[3]);
            slow.bad |= !(fabsf(s) <= 3.0e38f);
            slow.cells(q);
        }
    };

    for (int ub = 0; ub < nrows; ub += U) {
#pragma unroll
        for (int s = 0; s < U; ++s) {
            const int u = ub + s;
            if (u >= nrows) break;                                // block-uniform
            const float v = vq[s], dx = dxq[s], dy = dyq[s];
            vq[s] = __ldg(I + nxt);                               // row u + U (clamped: the last rows are fetched again)
            dxq[s] = __ldg(Dx + nxt);
            dyq[s] = __ldg(Dy + nxt);
            nxt = min(nxt + f.ny, last);
            const bool fast = lean_deposit<SC, SR, false>(win_s, trow0 + u, tcol, win_r, win_c, v, dx, dy, scale, 0u, vmax_bits);
            if (!fast && live) {
                const unsigned slot = atomicAdd(qcount, 1u);
                if (slot < (unsigned)MQ) queue[slot] = make_uint4(((unsigned)tid << 8) | (unsigned)u, __float_as_uint(v), __float_as_uint(dx), __float_as_uint(dy));
                else long_way(i0 + u, j, v, dx, dy);
            }
        }
    }
    __syncthreads();
    {
        const unsigned nq = min(*qcount, (unsigned)MQ);
        for (unsigned k = tid; k < nq; k += TILE_COLS) {
            const uint4 e = queue[k];
            long_way(i0 + (int)(e.x & 0xFFu), blockIdx.x * TILE_COLS + (int)((e.x >> 8) & 0xFFu), __uint_as_float(e.y), __uint_as_float(e.z),
                     __uint_as_float(e.w));
        }
    }
    slow.finish();
    if (fixed_ok) {
        const bool vec = (f.ny & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
        if (vec && clo >= 0 && clo + SC <= f.ny) flush_rows_fast<SC>(tile_smem, out, rlo, clo, max(0, -rlo), min(used_rows, f.nx - rlo), f.ny, inv_scale);
        else flush_tile<SR, SC>(tile_smem, out, rlo, clo, f.nx, f.ny, inv_scale, vec, used_rows);
    }
}

int launch_splat_tile(const float* I, const float* Dx, const float* Dy, float* out, const Frame& f, int* flag, cudaStream_t s) {
    constexpr int TR = 16, MQ = 512, H = 4;
    constexpr int SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H;
    constexpr size_t smem = sizeof(unsigned) * (SR * SC + 4 * MQ + 4 + TILE_COLS / 32);
    static int slots_of[32] = {0};             // resident blocks x SMs, per device
    int dev = 0;
    PARESIS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 32) dev = 0;
    if (!slots_of[dev]) {
        PARESIS_CUDA(cudaFuncSetAttribute(splat_tile_kernel<TR, MQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        PARESIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, splat_tile_kernel<TR, MQ>, TILE_COLS, smem));
        PARESIS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        slots_of[dev] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    const int slots = slots_of[dev];
    const int strips = div_up(f.ny, TILE_COLS);
    const int rows = f.nx >= 64 ? pick_tile_rows(f.nx, strips, slots, TR) : min(f.nx, TR);
    dim3 grid(strips, div_up(f.nx, rows));
    splat_tile_kernel<TR, MQ><<<grid, TILE_COLS, smem, s>>>(I, Dx, Dy, out, f, rows, flag);
    PARESIS_LAUNCH_CHECK("splat_tile_kernel");
    return PARESIS_OK;
}

}  // namespace paresis
