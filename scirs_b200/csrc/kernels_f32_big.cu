// kernels_f32_big.cu — generated list of tile kernel instantiations (see kernel_inst.cuh)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f32_big(void (*add)(const KernelEntry&)) {
    SFC_ADD(float, 4096, 1, false)
    SFC_ADD(float, 4096, 4, false)
    SFC_ADD(float, 8192, 1, false)
    SFC_ADD(float, 8192, 2, false)
    SFC_ADD(float, 16384, 1, false)
}
}  // namespace sfc
