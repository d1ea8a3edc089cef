// Tiled detector kernels for oversampling 4 (Detector.py:79-119); see detector_tile.cuh.
#include "detector_tile.cuh"

namespace paresis {
PARESIS_DT_DISPATCH(4)
}  // namespace paresis
