// Tiled detector kernels for oversampling 3 (Detector.py:79-119); see detector_tile.cuh.
#include "detector_tile.cuh"

namespace paresis {
PARESIS_DT_DISPATCH(3)
}  // namespace paresis
