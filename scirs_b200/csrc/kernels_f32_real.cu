// kernels_f32_real.cu — fused real-transform flavours (rfft / irfft fast paths), f32
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_f32_real(void (*add)(const KernelEntry&)) {
    SFC_ADD_REAL(float, 32, 128)
    SFC_ADD_REAL(float, 64, 64)
    SFC_ADD_REAL(float, 128, 32)
    SFC_ADD_REAL(float, 256, 16)
    SFC_ADD_REAL(float, 512, 8)
    SFC_ADD_REAL(float, 512, 4)
    SFC_ADD_REAL(float, 1024, 4)
    SFC_ADD_REAL(float, 2048, 2)
    SFC_ADD_REAL(float, 2048, 1)
    SFC_ADD_REAL(float, 1024, 2)
    SFC_ADD_REAL(float, 4096, 1)
    SFC_ADD_REAL(float, 8192, 1)
    SFC_ADD_REAL(float, 16384, 1)
}
}  // namespace sfc
