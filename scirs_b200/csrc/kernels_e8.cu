// kernels_e8.cu — radix-8 (8 points per thread) variants: half the registers, twice the warps
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_e8(void (*add)(const KernelEntry&)) {
    SFC_ADD_E(double, 4096, 1, false, 8)
    SFC_ADD_E(double, 1024, 4, false, 8)
    add(::sfc::KernelInst<double, 1024, 4, false, 8, 1>::entry());
    SFC_ADD_E(double, 2048, 2, false, 8)
    SFC_ADD_E(double, 512, 8, false, 8)
    SFC_ADD_E(double, 8192, 1, false, 8)
    SFC_ADD_E(float, 4096, 1, false, 8)
    SFC_ADD_E(float, 2048, 2, false, 8)
}
}  // namespace sfc
