// kernels_f64_dbl_b.cu — generated list of tile kernel instantiations (see kernel_inst.cuh)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f64_dbl_b(void (*add)(const KernelEntry&)) {
    SFC_ADD(double, 1024, 4, true)
    SFC_ADD(double, 2048, 2, true)
    SFC_ADD(double, 2048, 1, true)
    SFC_ADD(double, 1024, 2, true)
    SFC_ADD(double, 4096, 1, true)
    SFC_ADD(double, 8192, 1, true)
}
}  // namespace sfc
