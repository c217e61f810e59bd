// kernels_f32_dbl_a.cu — generated list of tile kernel instantiations (see kernel_inst.cuh)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f32_dbl_a(void (*add)(const KernelEntry&)) {
    SFC_ADD(float, 8, 256, true)
    SFC_ADD(float, 16, 256, true)
    SFC_ADD(float, 32, 128, true)
    SFC_ADD(float, 64, 64, true)
    SFC_ADD(float, 128, 32, true)
    SFC_ADD(float, 256, 16, true)
    SFC_ADD(float, 512, 8, true)
}
}  // namespace sfc
