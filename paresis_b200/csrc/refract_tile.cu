// Instantiations of the shared-memory tile refraction kernel (refract_tile.cuh) and their dispatcher.
// Reference: refractionFileNumba2.py:25-86, :198-263; Sample.py:285-351; Experiment.py:463-474.
#include "refract_tile.cuh"

namespace paresis {

template <int NM, int TR, int H>
static int dispatch_tile_shape(const RefractArgs<float>& a, cudaStream_t s) {
    const bool dual = a.out_ref != nullptr, has_i = a.I_in != nullptr;
    if (dual) return has_i ? launch_refract_tile<NM, true, true, true, TR, H>(a, s)
                           : launch_refract_tile<NM, true, false, true, TR, H>(a, s);
    return has_i ? launch_refract_tile<NM, false, true, true, TR, H>(a, s)
                 : launch_refract_tile<NM, false, false, true, TR, H>(a, s);
}

template <int TR, int H>
static int dispatch_tile_layers(int n_layers, const RefractArgs<float>& a, cudaStream_t s) {
    switch (n_layers) {
        case 1: return dispatch_tile_shape<1, TR, H>(a, s);
        case 2: return dispatch_tile_shape<2, TR, H>(a, s);
        case 3: return dispatch_tile_shape<3, TR, H>(a, s);
        default: return dispatch_tile_shape<4, TR, H>(a, s);
    }
}

int dispatch_refract_tile(int n_layers, const RefractArgs<float>& a, cudaStream_t s) {
    return dispatch_tile_layers<16, 4>(n_layers, a, s);
}

}  // namespace paresis
