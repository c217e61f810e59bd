// kernels_f64_big.cu — generated list of tile kernel instantiations (see kernel_inst.cuh)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f64_big(void (*add)(const KernelEntry&)) {
    SFC_ADD(double, 2048, 4, false)
    SFC_ADD(double, 4096, 1, false)
    SFC_ADD(double, 4096, 2, false)
    SFC_ADD(double, 8192, 1, false)
}
}  // namespace sfc
