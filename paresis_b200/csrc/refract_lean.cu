// Instantiations of the lean fixed-point tile hop (refract_lean.cuh) and their dispatcher.
// Reference: refractionFileNumba2.py:25-86, :198-263; Sample.py:285-351; Experiment.py:463-474.
#include "refract_lean.cuh"

namespace paresis {

template <int NM>
static int dispatch_lean_shape(const LeanArgs& a, int n_batch, cudaStream_t s) {
    const bool dual = a.z[0].out_ref != nullptr, has_i = a.z[0].I_in != nullptr;
    if (dual) return has_i ? launch_refract_lean<NM, true, true, LEAN_TR_DUAL>(a, n_batch, s) : launch_refract_lean<NM, true, false, LEAN_TR_DUAL>(a, n_batch, s);
    return has_i ? launch_refract_lean<NM, false, true, LEAN_TR_SINGLE>(a, n_batch, s) : launch_refract_lean<NM, false, false, LEAN_TR_SINGLE>(a, n_batch, s);
}

int dispatch_refract_lean_batch(int n_layers, const LeanArgs& a, int n_batch, cudaStream_t s) {
    switch (n_layers) {
        case 1: return dispatch_lean_shape<1>(a, n_batch, s);
        case 2: return dispatch_lean_shape<2>(a, n_batch, s);
        case 3: return dispatch_lean_shape<3>(a, n_batch, s);
        default: return dispatch_lean_shape<4>(a, n_batch, s);
    }
}

int dispatch_refract_lean(int n_layers, const RefractArgs<float>& a, cudaStream_t s) {
    LeanArgs b{};
    static_cast<RefractArgs<float>&>(b) = a;
    for (int m = 0; m < PARESIS_MAX_LAYERS; ++m) b.z[0].map[m] = a.map[m];
    b.z[0].I_in = a.I_in; b.z[0].out_obj = a.out_obj; b.z[0].out_ref = a.out_ref;
    for (int k = 0; k < 3; ++k) b.z[0].zero[k] = a.zero[k];
    b.z[0].sum_ref = a.sum_ref; b.z[0].zero_scalar = a.zero_scalar;
    return dispatch_refract_lean_batch(n_layers, b, 1, s);
}

}  // namespace paresis

using namespace paresis;

// Experiment.py:463-474 for up to PARESIS_MAX_HOP_BATCH membrane positions in one launch of the tile hop (out +=).
extern "C" int paresis_refract_tile_batch(const paresis_tile_hop_item* items_host, int n_items, const paresis_layer* layers_host,
                                          int n_layers, float intensity_uniform, float intensity_scale, int clear_input, int nx, int ny,
                                          int* flag, paresis_stream stream) {
    if (!items_host || n_items < 1 || n_items > LEAN_MAX_BATCH || !layers_host || n_layers < 1 || n_layers > PARESIS_MAX_LAYERS) {
        set_last_error("paresis_refract_tile_batch: need 1..%d items and 1..%d layers", LEAN_MAX_BATCH, PARESIS_MAX_LAYERS);
        return PARESIS_ERR_ARG;
    }
    if (nx < 3 || ny < 3 || (long)nx * ny >= (1L << 30) || !(intensity_scale > 0.f) || !(intensity_scale < 3.0e38f)) {
        set_last_error("paresis_refract_tile_batch: need 3 <= n, nx*ny < 2^30 and a positive intensity scale");
        return PARESIS_ERR_ARG;
    }
    LeanArgs a{};
    const bool dual = items_host[0].out_ref != nullptr, has_i = items_host[0].intensity_in != nullptr;
    for (int z = 0; z < n_items; ++z) {
        const paresis_tile_hop_item& it = items_host[z];
        if (!it.out_obj || (it.out_ref != nullptr) != dual || (it.intensity_in != nullptr) != has_i) {
            set_last_error("paresis_refract_tile_batch: item %d: every item needs out_obj and the same set of optional images", z);
            return PARESIS_ERR_ARG;
        }
        for (int m = 0; m < n_layers; ++m) {
            if (!it.thickness[m]) { set_last_error("paresis_refract_tile_batch: item %d: null map %d", z, m); return PARESIS_ERR_ARG; }
            a.z[z].map[m] = it.thickness[m];
        }
        a.z[z].I_in = it.intensity_in;
        a.z[z].out_obj = it.out_obj;
        a.z[z].out_ref = it.out_ref;
        int nz = 0;
        for (int k = 0; k < 3; ++k) if (it.zero_fill[k]) a.z[z].zero[nz++] = it.zero_fill[k];
        for (int k = nz; k < 3 && nz > 0; ++k) a.z[z].zero[k] = a.z[z].zero[0];
        a.z[z].sum_ref = dual ? it.sum_ref : nullptr;
        a.z[z].zero_scalar = it.zero_scalar;
    }
    for (int m = 0; m < n_layers; ++m) {
        a.g_obj[m] = layers_host[m].grad_obj;
        a.g_ref[m] = layers_host[m].grad_ref;
        a.att[m] = layers_host[m].atten;
    }
    a.I_uniform = intensity_uniform;
    a.f = Frame{nx, ny, 15};
    a.clamp_x = (float)nx; a.clamp_y = (float)ny;
    a.flag = flag;
    a.clear_input = clear_input != 0 && has_i;
    a.intensity_scale = intensity_scale;
    return dispatch_refract_lean_batch(n_layers, a, n_items, (cudaStream_t)stream);
}
