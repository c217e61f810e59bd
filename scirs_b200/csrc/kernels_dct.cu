// kernels_dct.cu — TM_FAST_DCT2: DCT-II of real rows of N = 2L points in one kernel (dct.rs:523-559)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_dct(void (*add)(const KernelEntry&)) {
    SFC_ADD_DCT2(double, 64, 64)
    SFC_ADD_DCT2(double, 128, 32)
    SFC_ADD_DCT2(double, 256, 16)
    SFC_ADD_DCT2(double, 512, 8)
    SFC_ADD_DCT2(double, 512, 4)
    SFC_ADD_DCT2(double, 1024, 4)
    SFC_ADD_DCT2(double, 2048, 2)
    SFC_ADD_DCT2(double, 2048, 1)
    SFC_ADD_DCT2(double, 1024, 2)
    SFC_ADD_DCT2(double, 4096, 1)
    SFC_ADD_DCT2(double, 8192, 1)
    // TM_FAST_DCT4 (experimental, SFC_DCT4_FUSED=1): 32 KiB tiles for rows, the wide 64 KiB ones for strided axes
    SFC_ADD_DCT4(double, 64, 64)
    SFC_ADD_DCT4(double, 128, 32)
    SFC_ADD_DCT4(double, 256, 16)
    SFC_ADD_DCT4(double, 512, 8)
    SFC_ADD_DCT4(double, 1024, 4)
    SFC_ADD_DCT4(double, 2048, 2)
    SFC_ADD_DCT4(double, 64, 32)
    SFC_ADD_DCT4(double, 128, 16)
    SFC_ADD_DCT4(double, 256, 8)
    SFC_ADD_DCT4(double, 512, 4)
    SFC_ADD_DCT4(double, 1024, 2)
    SFC_ADD_DCT4(double, 2048, 1)
    SFC_ADD_DCT4(double, 4096, 1)
    SFC_ADD_DCT4(double, 8192, 1)
    // TM_FAST_2D (experimental, SFC_FFT2_TILE2D=1): 16 x 32 tile on 8 adjacent columns
    SFC_ADD_2D(double, 512, 8)
    SFC_ADD_2D(double, 256, 16)   // 4096 columns: 16 x 16 tile on 16 adjacent columns
    SFC_ADD_2D(double, 1024, 4)   // 16384 columns: 16 x 64 tile on 4 adjacent columns (64-byte segments)
}
}  // namespace sfc
