// kernels_f32_mid.cu — generated list of tile kernel instantiations (see kernel_inst.cuh)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f32_mid(void (*add)(const KernelEntry&)) {
    SFC_ADD(float, 256, 16, false)
    SFC_ADD(float, 512, 8, false)
    SFC_ADD(float, 512, 4, false)
    SFC_ADD(float, 512, 16, false)
    SFC_ADD(float, 1024, 4, false)
    SFC_ADD(float, 1024, 16, false)
    SFC_ADD(float, 2048, 2, false)
    SFC_ADD(float, 2048, 1, false)
    SFC_ADD(float, 1024, 2, false)
    SFC_ADD(float, 2048, 8, false)
    // one thread group (named barrier) per lane, as for f64 (kernels_f64_mid.cu)
    add(::sfc::KernelInst<float, 512, 4, false, 16, 1, 4>::entry());
    add(::sfc::KernelInst<float, 1024, 2, false, 16, 1, 2>::entry());
}
}  // namespace sfc
