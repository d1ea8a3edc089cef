// Instantiations of the two-columns-per-thread refraction kernel (refract_pair.cuh) and their dispatcher.
// Reference: refractionFileNumba2.py:25-86, :198-263; Sample.py:285-351; Experiment.py:463-474.
#include "refract_pair.cuh"

namespace paresis {

template <int NM>
static int dispatch_pair_shape(const RefractArgs<float>& a, int rows, cudaStream_t s) {
    const bool dual = a.out_ref != nullptr, has_i = a.I_in != nullptr;
    if (dual) return has_i ? launch_refract_pair<NM, true, true, true>(a, rows, s)
                           : launch_refract_pair<NM, true, false, true>(a, rows, s);
    return has_i ? launch_refract_pair<NM, false, true, true>(a, rows, s)
                 : launch_refract_pair<NM, false, false, true>(a, rows, s);
}

int dispatch_refract_pair(int n_layers, const RefractArgs<float>& a, int rows_override, cudaStream_t s) {
    switch (n_layers) {
        case 1: return dispatch_pair_shape<1>(a, rows_override, s);
        case 2: return dispatch_pair_shape<2>(a, rows_override, s);
        case 3: return dispatch_pair_shape<3>(a, rows_override, s);
        default: return dispatch_pair_shape<4>(a, rows_override, s);
    }
}

}  // namespace paresis
