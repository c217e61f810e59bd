// kernels_f32_small.cu — generated list of tile kernel instantiations (see kernel_inst.cuh)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f32_small(void (*add)(const KernelEntry&)) {
    SFC_ADD(float, 2, 256, false)
    SFC_ADD(float, 4, 256, false)
    SFC_ADD(float, 8, 256, false)
    SFC_ADD(float, 16, 256, false)
    SFC_ADD(float, 32, 128, false)
    SFC_ADD(float, 64, 64, false)
    SFC_ADD(float, 128, 32, false)
}
}  // namespace sfc
