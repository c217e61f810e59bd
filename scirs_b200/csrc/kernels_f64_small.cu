// kernels_f64_small.cu — generated list of tile kernel instantiations (see kernel_inst.cuh)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f64_small(void (*add)(const KernelEntry&)) {
    SFC_ADD(double, 2, 256, false)
    SFC_ADD(double, 4, 256, false)
    SFC_ADD(double, 8, 256, false)
    SFC_ADD(double, 16, 256, false)
    SFC_ADD(double, 32, 128, false)
    SFC_ADD(double, 64, 64, false)
    SFC_ADD(double, 128, 32, false)
    SFC_ADD(double, 128, 16, false)
    SFC_ADD(double, 64, 32, false)
}
}  // namespace sfc
