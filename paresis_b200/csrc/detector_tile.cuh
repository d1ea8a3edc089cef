// Tiled detector kernel: oversampled image(s) -> detector counts in ONE pass, for the kernel sizes
// PARESIS actually produces (compile-time OS, HS = round(3 sigma_src), HP = round(3 PSF)).
//
// Reference: Detector.py:79-119.  Same arithmetic as detect_fused_kernel (detector.cu) -- reflect
// padding by index arithmetic, zeros beyond the padded frame, composite blur+bin kernel, separable
// PSF, crop, Poisson -- organised for throughput:
//   * a block owns a 32 x 64 tile of detector pixels; blockIdx.z selects the image, so the sample
//     and reference images (and propagation / white at position 0) share one launch;
//   * phase 1 folds "blur + bin along x" into the global loads: a thread streams 128-bit loads
//     down a column group and keeps the running sums in registers (the oversampled window never
//     touches shared memory);
//   * phases 2-4 (blur + bin along y, PSF along y, PSF along x) run out of shared memory with
//     compile-time taps held in registers, several outputs per thread;
//   * Poisson draws share one Philox block per pixel pair; draws that fail the quick PTRS test are
//     queued with their random words and finished by dense warps, step by step (poisson.cuh).
// Instantiated per oversampling factor in detector_tile_os*.cu.
#pragma once
#include "detector_common.cuh"
#include "poisson.cuh"

namespace paresis {

constexpr int DT_TX = 32, DT_TY = 64, DT_THREADS = 256;
#ifndef DT_MIN_BLOCKS
#define DT_MIN_BLOCKS 4
#endif

template <int OS, int HS, int HP>
struct DetTile {
    static constexpr int TAPS = OS + 2 * HS;
    static constexpr int NP = 2 * HP + 1;
    static constexpr int BWX = DT_TX + 2 * HP, BWY = DT_TY + 2 * HP;   // binned window
    static constexpr int SWX = BWX * OS + 2 * HS;                      // source rows of the window
    static constexpr int SWY = BWY * OS + 2 * HS;                      // source columns of the window
    static constexpr int CG = (SWY + 3 + 3) / 4;                       // float4 column groups (window start aligned down)
    static constexpr int R1W = CG * 4;                                 // row pitch of R1
    static constexpr int RCH = DT_THREADS / CG > BWX ? BWX : DT_THREADS / CG;   // row chunks processed side by side
    static constexpr int UCH = (BWX + RCH - 1) / RCH;                  // binned rows per chunk
    static constexpr int NR = UCH * OS + 2 * HS;                       // source rows a chunk reads
    static constexpr int R2W = (BWY + 3) / 4 * 4 + 4;                    // pitch of R2: 16 B rows + slack for 128-bit over-read
    static constexpr int R3W = DT_TY;
    static constexpr int R1_FLOATS = BWX * R1W;                        // also R3
    static constexpr int R2_FLOATS = BWX * R2W;
    // Poisson queues (after phase 4 nothing else lives in R1 / R2): queue A = one 16-byte entry per pixel the quick
    // test left open (worst case every pixel), queue B = the 8-byte entries of what step 0 rejected, behind it
    static constexpr int QA_FLOATS = 4 * DT_TX * DT_TY, QB_CAP = 1024;
    static constexpr int WORK_FLOATS = R1_FLOATS + R2_FLOATS > QA_FLOATS + 2 * QB_CAP ? R1_FLOATS + R2_FLOATS : QA_FLOATS + 2 * QB_CAP;
    static constexpr size_t SMEM = sizeof(float) * WORK_FLOATS + sizeof(int) * SWX;
    static_assert(CG <= DT_THREADS, "source window too wide for one block");
    static_assert(BWX * R3W <= R1_FLOATS, "R3 must fit in R1");
};

template <int OS, int HS, int HP>
__global__ void __launch_bounds__(DT_THREADS, DT_MIN_BLOCKS)
detect_tile_kernel(DetImages im, int nx, int ny, int det_x, int det_y, const float* __restrict__ gsrc,
                   const float* __restrict__ gpsf, int noise, uint64_t seed) {
    using T = DetTile<OS, HS, HP>;
    extern __shared__ __align__(16) float dt_smem[];
    float* R1 = dt_smem;
    float* R2 = R1 + T::R1_FLOATS;
    float* R3 = R1;
    int* rowoff = reinterpret_cast<int*>(dt_smem + T::WORK_FLOATS);
    __shared__ int n_queued, n_requeued;

    const int tid = threadIdx.x;
    const float* __restrict__ img = im.img[blockIdx.z];
    float* __restrict__ out = im.out[blockIdx.z];
    const uint64_t seq = im.seq[blockIdx.z];
    const int pad = DET_PAD * OS, npx = nx + 2 * pad, npy = ny + 2 * pad;
    const int bxn = det_x + 2 * DET_PAD, byn = det_y + 2 * DET_PAD;
    const int a0 = blockIdx.y * DT_TX, b0 = blockIdx.x * DT_TY;          // detector origin of the tile
    const int u0 = a0 + DET_PAD - HP, v0 = b0 + DET_PAD - HP;            // binned, padded frame
    const int x0 = u0 * OS - HS;                                         // padded source row of window row 0
    const int yi0 = (b0 - HP) * OS - HS;                                 // image column of window column 0
    const bool out_aligned = (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    const int o = ((yi0 % 4) + 4) % 4;
    const int ya = yi0 - o;                                              // aligned image column of R1 column 0

    // composite blur+bin taps W(d) = sum_{a<OS} g(d - a) and the PSF taps, in registers
    float W[T::TAPS], P[T::NP];
#pragma unroll
    for (int t = 0; t < T::TAPS; ++t) {
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < OS; ++a) {
            const int e = t - HS - a;
            if (e >= -HS && e <= HS) acc += HS > 0 ? __ldg(gsrc + e + HS) : 1.f;
        }
        W[t] = acc;
    }
#pragma unroll
    for (int t = 0; t < T::NP; ++t) P[t] = HP > 0 ? __ldg(gpsf + t) : 1.f;

    if (tid == 0) { n_queued = 0; n_requeued = 0; }
    for (int a = tid; a < T::SWX; a += DT_THREADS) {
        const int xp = x0 + a;
        rowoff[a] = (xp >= 0 && xp < npx) ? reflect_index(xp - pad, nx) * ny : -1;   // nx*ny < 2^30 (host check)
    }
    __syncthreads();

    // ---- phase 1: R1[u][c] = sum_s W[s] * src(u*OS + s, ya + c)          (blur + bin along x)
    {
        const int cg = tid % T::CG, rc = tid / T::CG;
        if (rc < T::RCH) {
            const int u_lo = rc * T::UCH;
            const bool fast = ya >= 0 && ya + T::R1W <= ny && (ny & 3) == 0 && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
            float4 acc[T::UCH];
#pragma unroll
            for (int k = 0; k < T::UCH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (fast && x0 >= 0 && x0 + T::SWX <= npx) {
                // every row of the window exists and every quad is one aligned 128-bit load: no branch between the
                // loads, so that they are all in flight together (rows past the window repeat its last row; their
                // taps feed accumulators that are not stored)
                const float* base = img + ya + 4 * cg;
                constexpr int BATCH = 8;
#pragma unroll
                for (int r0 = 0; r0 < T::NR; r0 += BATCH) {
                    float4 v[BATCH];
#pragma unroll
                    for (int b = 0; b < BATCH; ++b)
                        if (r0 + b < T::NR) {
                            PARESIS_BOUND(rowoff[min(u_lo * OS + r0 + b, T::SWX - 1)] + ya + 4 * cg + 3, nx * ny);
                            v[b] = __ldg(reinterpret_cast<const float4*>(base + rowoff[min(u_lo * OS + r0 + b, T::SWX - 1)]));
                        }
#pragma unroll
                    for (int b = 0; b < BATCH; ++b) {
                        const int r = r0 + b;
                        if (r >= T::NR) break;
#pragma unroll
                        for (int k = 0; k < T::UCH; ++k) {
                            const int t = r - k * OS;   // compile-time after unrolling
                            if (t >= 0 && t < T::TAPS) {
                                acc[k].x = fmaf(W[t], v[b].x, acc[k].x);
                                acc[k].y = fmaf(W[t], v[b].y, acc[k].y);
                                acc[k].z = fmaf(W[t], v[b].z, acc[k].z);
                                acc[k].w = fmaf(W[t], v[b].w, acc[k].w);
                            }
                        }
                    }
                }
            } else {
                int col[4];
                if (!fast) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int yp = ya + pad + 4 * cg + e;
                        col[e] = (yp >= 0 && yp < npy) ? reflect_index(yp - pad, ny) : -1;
                    }
                }
#pragma unroll
                for (int r = 0; r < T::NR; ++r) {
                    const int a = u_lo * OS + r;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int ro = a < T::SWX ? rowoff[a] : -1;
                    if (ro >= 0) {
                        if (fast) {
                            v = __ldg(reinterpret_cast<const float4*>(img + ro + ya + 4 * cg));
                        } else {
                            if (col[0] >= 0) v.x = __ldg(img + ro + col[0]);
                            if (col[1] >= 0) v.y = __ldg(img + ro + col[1]);
                            if (col[2] >= 0) v.z = __ldg(img + ro + col[2]);
                            if (col[3] >= 0) v.w = __ldg(img + ro + col[3]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < T::UCH; ++k) {
                        const int t = r - k * OS;   // compile-time after unrolling
                        if (t >= 0 && t < T::TAPS) {
                            acc[k].x = fmaf(W[t], v.x, acc[k].x);
                            acc[k].y = fmaf(W[t], v.y, acc[k].y);
                            acc[k].z = fmaf(W[t], v.z, acc[k].z);
                            acc[k].w = fmaf(W[t], v.w, acc[k].w);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < T::UCH; ++k)
                if (u_lo + k < T::BWX) *reinterpret_cast<float4*>(R1 + (u_lo + k) * T::R1W + 4 * cg) = acc[k];
        }
    }
    __syncthreads();

    // ---- phase 2: R2[u][v] = sum_t W[t] * R1[u][o + v*OS + t], zero outside the padded binned frame
    {
        constexpr int VB = T::BWY % 4 == 0 ? 4 : 2, PAIRS = T::BWY / VB;
        for (int task = tid; task < T::BWX * PAIRS; task += DT_THREADS) {
            const int u = task / PAIRS, vp = task - u * PAIRS;
            const float* src = R1 + u * T::R1W + o + vp * VB * OS;
            float in[VB * OS + 2 * HS];
#pragma unroll
            for (int k = 0; k < VB * OS + 2 * HS; ++k) in[k] = src[k];
            const bool uin = (unsigned)(u0 + u) < (unsigned)bxn;
#pragma unroll
            for (int k = 0; k < VB; ++k) {
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < T::TAPS; ++t) acc = fmaf(W[t], in[k * OS + t], acc);
                const int v = vp * VB + k;
                // fftconvolve(mode='same') sees zeros beyond the padded frame (Detector.py:106-108)
                R2[u * T::R2W + v] = (uin && (unsigned)(v0 + v) < (unsigned)byn) ? acc : 0.f;
            }
        }
    }
    __syncthreads();

    // ---- phase 3: R3[u][b] = sum_f P[f] * R2[u][b + f]                     (PSF along y)
    {
        constexpr int QUADS = DT_TY / 4, NIN = 4 + 2 * HP, NQ = (NIN + 3) / 4;
        for (int task = tid; task < T::BWX * QUADS; task += DT_THREADS) {
            const int u = task / QUADS, bq = task - u * QUADS;
            const float4* src = reinterpret_cast<const float4*>(R2 + u * T::R2W + 4 * bq);
            float in[NQ * 4];
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
                const float4 q = src[k];
                in[4 * k] = q.x; in[4 * k + 1] = q.y; in[4 * k + 2] = q.z; in[4 * k + 3] = q.w;
            }
            float r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float acc = 0.f;
#pragma unroll
                for (int f = 0; f < T::NP; ++f) acc = fmaf(P[f], in[k + f], acc);
                r[k] = acc;
            }
            *reinterpret_cast<float4*>(R3 + u * T::R3W + 4 * bq) = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
    __syncthreads();

    // ---- phase 4: out[a][b] = sum_e P[e] * R3[a + e][b]  (PSF along x), Poisson, store
    constexpr int RB = DT_TX / (DT_THREADS / 32);     // rows per thread (4)
    const int lane = tid & 31, warp = tid >> 5;
    const int al = warp * RB;                         // first local row
    float2 acc[RB];
    {
        const float2* src = reinterpret_cast<const float2*>(R3 + al * T::R3W + 2 * lane);
#pragma unroll
        for (int k = 0; k < RB; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < RB + 2 * HP; ++r) {
            const float2 v = src[r * (T::R3W / 2)];
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int e = r - k;
                if (e >= 0 && e < T::NP) { acc[k].x = fmaf(P[e], v.x, acc[k].x); acc[k].y = fmaf(P[e], v.y, acc[k].y); }
            }
        }
    }
    const int db = b0 + 2 * lane;
    if (!noise) {
#pragma unroll
        for (int k = 0; k < RB; ++k) {
            const int da = a0 + al + k;
            if (da >= det_x || db >= det_y) continue;
            const size_t p = (size_t)da * det_y + db;
            if (db + 1 < det_y && out_aligned && (p & 1) == 0) {
                *reinterpret_cast<float2*>(out + p) = acc[k];
            } else {
                out[p] = acc[k].x;
                if (db + 1 < det_y) out[p + 1] = acc[k].y;
            }
        }
        return;
    }
    __syncthreads();                                  // R3 has been read: R1 / R2 become the Poisson queues
    uint4* const qa = reinterpret_cast<uint4*>(dt_smem);
    int2* const qb = reinterpret_cast<int2*>(dt_smem + T::QA_FLOATS);
    {
        // quick test of every pixel (one Philox block per aligned pixel pair); what it leaves open is queued
        // together with its two random words, so that nothing is generated twice
#pragma unroll
        for (int k = 0; k < RB; ++k) {
            const int da = a0 + al + k;
            if (da >= det_x || db >= det_y) continue;
            const size_t p = (size_t)da * det_y + db;
            const bool two = db + 1 < det_y;
            const bool pair = two && (p & 1) == 0;
            uint32_t r[4];
            float x0v, x1v = 0.f;
            bool ok0, ok1 = true;
            if (pair) {
                Philox g = poisson_stream(seed, seq, p);
                g.generate(r);
                ok0 = poisson_decide(acc[k].x, r[0], r[1], x0v);
                ok1 = poisson_decide(acc[k].y, r[2], r[3], x1v);
            } else {
                poisson_words(seed, seq, p, r[0], r[1]);
                ok0 = poisson_decide(acc[k].x, r[0], r[1], x0v);
                if (two) {
                    poisson_words(seed, seq, p + 1, r[2], r[3]);
                    ok1 = poisson_decide(acc[k].y, r[2], r[3], x1v);
                }
            }
            if (pair && out_aligned) {
                *reinterpret_cast<float2*>(out + p) = make_float2(x0v, x1v);
            } else {
                out[p] = x0v;
                if (two) out[p + 1] = x1v;
            }
            if (!(ok0 && ok1)) {
                const int n = (ok0 ? 0 : 1) + (ok1 ? 0 : 1);
                int slot = atomicAdd(&n_queued, n);
                const unsigned local = (unsigned)((al + k) * DT_TY + 2 * lane);
                PARESIS_BOUND(slot + n - 1, DT_TX * DT_TY);
                if (!ok0) qa[slot++] = make_uint4(local, __float_as_uint(acc[k].x), r[0], r[1]);
                if (!ok1) qa[slot] = make_uint4(local + 1u, __float_as_uint(acc[k].y), r[2], r[3]);
            }
        }
    }
    __syncthreads();
    {
        // step 0 on dense warps: the full test of the candidate the quick test left open (or the whole draw of a
        // small mean); what it rejects goes to queue B (or, if that is full, is finished on the spot)
        const int nq = n_queued;
        for (int qi = tid; qi < nq; qi += DT_THREADS) {
            const uint4 e = qa[qi];
            PARESIS_BOUND(e.x, DT_TX * DT_TY); PARESIS_BOUND(a0 + e.x / DT_TY, det_x); PARESIS_BOUND(b0 + e.x % DT_TY, det_y);
            const size_t p = (size_t)(a0 + e.x / DT_TY) * det_y + (b0 + e.x % DT_TY);
            const float lam = __uint_as_float(e.y);
            float x;
            if (poisson_first(lam, e.z, e.w, x)) {
                out[p] = x;
            } else {
                const int slot = atomicAdd(&n_requeued, 1);
                if (slot < T::QB_CAP) qb[slot] = make_int2((int)e.x, (int)e.y);
                else out[p] = poisson_rest(lam, seed, seq, p, 1u);
            }
        }
    }
    // blocks 1, 2, ... of the pixels still open, two candidates each, re-queued between blocks so that the warps stay dense
    int2* qcur = qb;
    int2* qnext = reinterpret_cast<int2*>(dt_smem);    // queue A is dead once step 0 is through
    int* ncur = &n_requeued;
    int* nnext = &n_queued;
    for (uint32_t blk = 1; blk < 64; ++blk) {
        __syncthreads();                               // pushes into qcur are complete, qnext has been read
        const int nq = min(*ncur, T::QB_CAP);
        if (nq == 0) break;                            // block-uniform
        if (tid == 0) *nnext = 0;
        __syncthreads();
        for (int qi = tid; qi < nq; qi += DT_THREADS) {
            const int2 e = qcur[qi];
            const size_t p = (size_t)(a0 + e.x / DT_TY) * det_y + (b0 + e.x % DT_TY);
            float x;
            if (poisson_block(__int_as_float(e.y), seed, seq, p, blk, x)) out[p] = x;
            else {
                const int slot = atomicAdd(nnext, 1);   // at most nq <= QB_CAP entries
                PARESIS_BOUND(slot, T::QB_CAP);
                qnext[slot] = e;
            }
        }
        int2* tq = qcur; qcur = qnext; qnext = tq;
        int* tn = ncur; ncur = nnext; nnext = tn;
    }
}

template <int OS, int HS, int HP>
static int launch_detect_tile(const DetImages& im, int n_images, int nx, int ny, int det_x, int det_y, const float* gs,
                              const float* gp, int noise, uint64_t seed, cudaStream_t st) {
    using T = DetTile<OS, HS, HP>;
    static bool configured[32] = {false};      // per device: the attribute belongs to the device's copy of the kernel
    int dev = 0;
    PARESIS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 32) dev = 0;
    if (!configured[dev]) {
        PARESIS_CUDA(cudaFuncSetAttribute(detect_tile_kernel<OS, HS, HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
        configured[dev] = true;
    }
    const dim3 grid(div_up(det_y, DT_TY), div_up(det_x, DT_TX), n_images);
    detect_tile_kernel<OS, HS, HP><<<grid, DT_THREADS, T::SMEM, st>>>(im, nx, ny, det_x, det_y, gs, gp, noise, seed);
    PARESIS_LAUNCH_CHECK("detect_tile_kernel");
    return PARESIS_OK;
}

// One dispatcher per oversampling factor: HS = round(3 sigma_src) in {0, 1, 2}, HP = round(3 PSF) in {0, 2, 3, 4, 6}
// (PSF 0, 0.5-0.8, 1, 1.2-1.5, 2 px).  Anything else takes the generic kernel of detector.cu.
#define PARESIS_DT_DISPATCH(OS_)                                                                                       \
    int dispatch_detect_tile_os##OS_(int hs, int hp, PARESIS_DT_ARGS) {                                                \
        PARESIS_DT_ROW(OS_, 0) PARESIS_DT_ROW(OS_, 1) PARESIS_DT_ROW(OS_, 2)                                           \
        return -1;                                                                                                     \
    }
#define PARESIS_DT_CASE(OS_, HS_, HP_) \
    if (hs == HS_ && hp == HP_) return launch_detect_tile<OS_, HS_, HP_>(im, n_images, nx, ny, det_x, det_y, gs, gp, noise, seed, st);
#define PARESIS_DT_ROW(OS_, HS_) \
    PARESIS_DT_CASE(OS_, HS_, 0) PARESIS_DT_CASE(OS_, HS_, 2) PARESIS_DT_CASE(OS_, HS_, 3) PARESIS_DT_CASE(OS_, HS_, 4) \
    PARESIS_DT_CASE(OS_, HS_, 6)

}  // namespace paresis
