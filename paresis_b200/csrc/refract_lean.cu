// Instantiations of the lean fixed-point tile hop (refract_lean.cuh) and their dispatcher.
// Reference: refractionFileNumba2.py:25-86, :198-263; Sample.py:285-351; Experiment.py:463-474.
#include "refract_lean.cuh"

namespace paresis {

template <int NM>
static int dispatch_lean_shape(const RefractArgs<float>& a, cudaStream_t s) {
    const bool dual = a.out_ref != nullptr, has_i = a.I_in != nullptr;
    if (dual) return has_i ? launch_refract_lean<NM, true, true, 16>(a, s) : launch_refract_lean<NM, true, false, 16>(a, s);
    return has_i ? launch_refract_lean<NM, false, true, 16>(a, s) : launch_refract_lean<NM, false, false, 16>(a, s);
}

int dispatch_refract_lean(int n_layers, const RefractArgs<float>& a, cudaStream_t s) {
    switch (n_layers) {
        case 1: return dispatch_lean_shape<1>(a, s);
        case 2: return dispatch_lean_shape<2>(a, s);
        case 3: return dispatch_lean_shape<3>(a, s);
        default: return dispatch_lean_shape<4>(a, s);
    }
}

}  // namespace paresis
