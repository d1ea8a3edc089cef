// Instantiations, drain launch and C entry points of the owner-computes rolling-strip hop (refract_strip.cuh).
// Reference: Experiment.py:463-474, Sample.py:285-351, refractionFileNumba2.py:25-86, :198-263.
#include "refract_strip.cuh"

namespace paresis {

__global__ void __launch_bounds__(128) refract_strip_drain_kernel(const HopDrainArgs a) {
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.n_slices) return;
    const unsigned n = a.far_count[w];
    if (n == 0u) return;
    const unsigned z = w / a.slices_per_item;
    const uint4* slice = a.far + (size_t)w * a.slice_cap;
    float* const out_obj = a.out_obj[z];
    float* const out_ref = a.out_ref[z];
    bool bad = false;
    float ref_sum = 0.f;
    for (unsigned k = threadIdx.x & 31; k < n; k += 32) {
        const uint4 e = slice[k];
        const int idx = (int)(e.x & FAR_INDEX);
        const int i = idx / a.f.ny, j = idx - i * a.f.ny;
        const float v = __uint_as_float(e.y), dx = __uint_as_float(e.z), dy = __uint_as_float(e.w);
        if (e.x & FAR_TWIN) {
            ref_sum += deposit_direct<true>(out_obj, out_ref, i, j, v, dx, dy, a.f.nx, a.f.ny, bad);
        } else if (e.x & FAR_REF) {
            ref_sum += deposit_direct<false>(out_ref, nullptr, i, j, v, dx, dy, a.f.nx, a.f.ny, bad);
        } else {
            deposit_direct<false>(out_obj, nullptr, i, j, v, dx, dy, a.f.nx, a.f.ny, bad);
        }
    }
    if (bad && a.flag) atomicOr(a.flag, FLAG_NONFINITE);
    if (a.sum_ref[z]) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ref_sum += __shfl_xor_sync(FULL_MASK, ref_sum, d);
        if ((threadIdx.x & 31) == 0 && ref_sum != 0.f) atomicAdd(a.sum_ref[z], (double)ref_sum);
    }
}

static size_t list_bytes(const StripPlan& p, int n_items, bool dual, size_t* slices_out) {
    const size_t slices = (size_t)p.strips * p.segs * STRIP_WARPS * n_items;
    if (slices_out) *slices_out = slices;
    return slices * p.far_cap * (dual ? 2 : 1) * sizeof(uint4) + slices * sizeof(unsigned);
}

template <int NM, bool DUAL, bool HAS_I, int H, bool ACC>
static int launch_refract_strip(HopArgs& a, int n_items, void* work, size_t work_bytes, cudaStream_t s, bool size_only, size_t* need) {
    constexpr size_t smem = sizeof(unsigned) * (Strip<H>::TILE_WORDS * (DUAL ? 2 : 1) + 8);
    static DeviceSlots ds;
    int slots = 0;
    int rc = ds.get(refract_strip_kernel<NM, DUAL, HAS_I, H, ACC>, STRIP_BLOCK, smem, &slots);
    if (rc) return rc;
    a.p = plan_strips(a.f.nx, a.f.ny, H, slots, n_items);
    size_t slices = 0;
    const size_t bytes = list_bytes(a.p, n_items, DUAL, &slices);
    if (size_only) { *need = bytes; return PARESIS_OK; }
    void* scratch = work;
    if (!work || work_bytes < bytes) {
        rc = strip_scratch_alloc(bytes, &scratch, s);
        if (rc) return rc;
    }
    a.far = reinterpret_cast<uint4*>(scratch);
    a.far_count = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(scratch) + slices * a.p.far_cap * (DUAL ? 2 : 1) * sizeof(uint4));
    dim3 grid(a.p.strips, a.p.segs, n_items);
    refract_strip_kernel<NM, DUAL, HAS_I, H, ACC><<<grid, STRIP_BLOCK, smem, s>>>(a);
    PARESIS_LAUNCH_CHECK("refract_strip_kernel");
    HopDrainArgs d{};
    for (int z = 0; z < n_items; ++z) { d.out_obj[z] = a.item[z].out_obj; d.out_ref[z] = a.item[z].out_ref; d.sum_ref[z] = DUAL ? a.item[z].sum_ref : nullptr; }
    d.far = a.far; d.far_count = a.far_count;
    d.slice_cap = a.p.far_cap * (DUAL ? 2 : 1);
    d.slices_per_item = (unsigned)(slices / n_items);
    d.n_slices = (unsigned)slices;
    d.f = a.f; d.flag = a.flag;
    refract_strip_drain_kernel<<<(unsigned)((slices + 3) / 4), 128, 0, s>>>(d);
    PARESIS_LAUNCH_CHECK("refract_strip_drain_kernel");
    if (scratch != work) return strip_scratch_free(scratch, s);
    return PARESIS_OK;
}

template <int NM, bool DUAL, int H>
static int dispatch_io(HopArgs& a, int n_items, bool has_i, bool acc, void* work, size_t work_bytes, cudaStream_t s, bool size_only, size_t* need) {
    if (has_i) return acc ? launch_refract_strip<NM, DUAL, true, H, true>(a, n_items, work, work_bytes, s, size_only, need)
                          : launch_refract_strip<NM, DUAL, true, H, false>(a, n_items, work, work_bytes, s, size_only, need);
    return acc ? launch_refract_strip<NM, DUAL, false, H, true>(a, n_items, work, work_bytes, s, size_only, need)
               : launch_refract_strip<NM, DUAL, false, H, false>(a, n_items, work, work_bytes, s, size_only, need);
}

// The shapes the per-energy loop uses: the membrane hop (one beam, uniform or mapped input, reach 8) and the object
// hop (one or two beams, mapped input, reach 12); other layer counts take the same kernels with more maps.
template <int NM>
static int dispatch_shape(HopArgs& a, int n_items, bool dual, bool has_i, bool acc, int reach, void* work, size_t work_bytes,
                          cudaStream_t s, bool size_only, size_t* need) {
    if (dual) return dispatch_io<NM, true, 12>(a, n_items, has_i, acc, work, work_bytes, s, size_only, need);
    if (reach <= 8) return dispatch_io<NM, false, 8>(a, n_items, has_i, acc, work, work_bytes, s, size_only, need);
    return dispatch_io<NM, false, 12>(a, n_items, has_i, acc, work, work_bytes, s, size_only, need);
}

static int dispatch_all(int n_layers, HopArgs& a, int n_items, bool dual, bool has_i, bool acc, int reach, void* work, size_t work_bytes,
                        cudaStream_t s, bool size_only, size_t* need) {
    switch (n_layers) {
        case 1: return dispatch_shape<1>(a, n_items, dual, has_i, acc, reach, work, work_bytes, s, size_only, need);
        case 2: return dispatch_shape<2>(a, n_items, dual, has_i, acc, reach, work, work_bytes, s, size_only, need);
        case 3: return dispatch_shape<3>(a, n_items, dual, has_i, acc, reach, work, work_bytes, s, size_only, need);
        default: return dispatch_shape<4>(a, n_items, dual, has_i, acc, reach, work, work_bytes, s, size_only, need);
    }
}

int dispatch_refract_strip(int n_layers, HopArgs& a, int n_items, bool dual, bool has_i, bool accumulate, int reach, void* work,
                           size_t work_bytes, cudaStream_t s) {
    return dispatch_all(n_layers, a, n_items, dual, has_i, accumulate, reach, work, work_bytes, s, false, nullptr);
}

size_t refract_strip_work_bytes(int nx, int ny, int n_layers, int n_items, bool dual, bool has_i, int reach) {
    HopArgs a{};
    a.f = Frame{nx, ny, 15};
    size_t need = 0;
    if (dispatch_all(n_layers, a, n_items, dual, has_i, false, reach, nullptr, 0, nullptr, true, &need) != PARESIS_OK) return 0;
    return need;
}

}  // namespace paresis

using namespace paresis;

extern "C" size_t paresis_refract_hop_work_bytes(int nx, int ny, int n_layers, int n_items, int dual, int has_intensity_map, int reach) {
    if (nx < 3 || ny < 3 || n_items < 1 || n_items > STRIP_MAX_BATCH || n_layers < 1 || n_layers > PARESIS_MAX_LAYERS) return 0;
    return refract_strip_work_bytes(nx, ny, n_layers, n_items, dual != 0, has_intensity_map != 0, reach);
}

extern "C" int paresis_refract_hop_batch(const paresis_hop_item* items_host, int n_items, const paresis_layer* layers_host, int n_layers,
                                         float intensity_uniform, float intensity_scale, int accumulate, int reach, int nx, int ny,
                                         void* work, size_t work_bytes, int* flag, paresis_stream stream) {
    if (!items_host || n_items < 1 || n_items > STRIP_MAX_BATCH || !layers_host || n_layers < 1 || n_layers > PARESIS_MAX_LAYERS) {
        set_last_error("paresis_refract_hop_batch: need 1..%d items and 1..%d layers", STRIP_MAX_BATCH, PARESIS_MAX_LAYERS);
        return PARESIS_ERR_ARG;
    }
    if (nx < 3 || ny < 3 || (long)nx * ny >= (1L << 30) || !(intensity_scale > 0.f) || !(intensity_scale < 1.0e37f)) {
        set_last_error("paresis_refract_hop_batch: need 3 <= n, nx*ny < 2^30 and a positive intensity scale");
        return PARESIS_ERR_ARG;
    }
    HopArgs a{};
    const bool dual = items_host[0].out_ref != nullptr, has_i = items_host[0].intensity_in != nullptr;
    bool vec = (ny & 3) == 0;
    for (int z = 0; z < n_items; ++z) {
        const paresis_hop_item& it = items_host[z];
        if (!it.out_obj || (it.out_ref != nullptr) != dual || (it.intensity_in != nullptr) != has_i) {
            set_last_error("paresis_refract_hop_batch: item %d: every item needs out_obj and the same set of optional images", z);
            return PARESIS_ERR_ARG;
        }
        for (int m = 0; m < n_layers; ++m) {
            if (!it.thickness[m]) { set_last_error("paresis_refract_hop_batch: item %d: null map %d", z, m); return PARESIS_ERR_ARG; }
            a.item[z].map[m] = it.thickness[m];
        }
        a.item[z].I_in = it.intensity_in;
        a.item[z].out_obj = it.out_obj;
        a.item[z].out_ref = it.out_ref;
        a.item[z].sum_ref = it.sum_ref;
        vec = vec && (reinterpret_cast<uintptr_t>(it.out_obj) & 15) == 0 && (reinterpret_cast<uintptr_t>(it.out_ref) & 15) == 0;
    }
    for (int m = 0; m < n_layers; ++m) {
        a.g_obj[m] = layers_host[m].grad_obj;
        a.g_ref[m] = layers_host[m].grad_ref;
        a.att[m] = layers_host[m].atten;
    }
    a.I_uniform = intensity_uniform;
    a.intensity_scale = intensity_scale;
    a.f = Frame{nx, ny, 15};
    a.flag = flag;
    a.vec = vec;
    return dispatch_refract_strip(n_layers, a, n_items, dual, has_i, accumulate != 0, reach, work, work_bytes, (cudaStream_t)stream);
}
