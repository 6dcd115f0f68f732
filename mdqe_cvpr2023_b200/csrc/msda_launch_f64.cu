// Sampling kernels for one dtype combination (value, location / weight, grad_value accumulator): see msda_launch.cuh.
#include "msda_launch.cuh"

namespace msda {
MSDA_LAUNCH_EXTERN(, double, double, double)
}  // namespace msda
