// kernels_f64_real.cu — fused real-transform flavours (rfft / irfft fast paths), f64
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f64_real(void (*add)(const KernelEntry&)) {
    SFC_ADD_REAL(double, 32, 128)
    SFC_ADD_REAL(double, 64, 64)
    SFC_ADD_REAL(double, 64, 32)
    SFC_ADD_REAL(double, 128, 32)
    SFC_ADD_REAL(double, 128, 16)
    SFC_ADD_REAL(double, 256, 16)
    SFC_ADD_REAL(double, 256, 8)
    SFC_ADD_REAL(double, 512, 8)
    SFC_ADD_REAL(double, 512, 4)
    SFC_ADD_REAL(double, 1024, 4)
    SFC_ADD_REAL(double, 2048, 2)
    SFC_ADD_REAL(double, 2048, 1)
    SFC_ADD_REAL(double, 1024, 2)
    SFC_ADD_REAL(double, 4096, 1)
    SFC_ADD_REAL(double, 8192, 1)
}
}  // namespace sfc
