// kernels_f64_mid.cu — generated list of tile kernel instantiations (see kernel_inst.cuh)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f64_mid(void (*add)(const KernelEntry&)) {
    SFC_ADD(double, 256, 16, false)
    SFC_ADD(double, 256, 8, false)
    SFC_ADD(double, 512, 8, false)
    SFC_ADD(double, 512, 4, false)
    SFC_ADD(double, 1024, 4, false)
    SFC_ADD(double, 1024, 8, false)
    SFC_ADD(double, 2048, 2, false)
    SFC_ADD(double, 2048, 1, false)
    SFC_ADD(double, 1024, 2, false)
    // one thread group (named barrier) per lane: contiguous-row tiles whose lanes are whole warps never synchronise the CTA
    add(::sfc::KernelInst<double, 512, 4, false, 16, 1, 4>::entry());
    add(::sfc::KernelInst<double, 1024, 2, false, 16, 1, 2>::entry());
}
}  // namespace sfc
