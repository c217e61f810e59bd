// kernels_r3.cu — power-of-three tiles (r3_tile.cuh), f64 and f32.  Two tiles per length: 243 threads / 35 KiB for contiguous
// rows (three CTAs per SM) and 486 threads / 70 KiB with twice the lanes for strided passes (two CTAs per SM, longer segments).
#include "r3_tile.cuh"
namespace sfc {
#define SFC_ADD_R3(T, L, TL) add(::sfc::R3Inst<T, L, TL>::entry());
#define SFC_ADD_R3_BOTH(T)                                     \
    SFC_ADD_R3(T, 9, 243) SFC_ADD_R3(T, 9, 486)                \
    SFC_ADD_R3(T, 27, 81) SFC_ADD_R3(T, 27, 162)               \
    SFC_ADD_R3(T, 81, 27) SFC_ADD_R3(T, 81, 54)                \
    SFC_ADD_R3(T, 243, 9) SFC_ADD_R3(T, 243, 18)               \
    SFC_ADD_R3(T, 729, 3) SFC_ADD_R3(T, 729, 6)                \
    SFC_ADD_R3(T, 2187, 1) SFC_ADD_R3(T, 2187, 2) SFC_ADD_R3(T, 2187, 3)
void register_kernels_r3(void (*add)(const KernelEntry&)) {
    SFC_ADD_R3_BOTH(double)
#ifndef SFC_HOST_EMUL  // the f32 path uses packed f32x2 PTX: not emulated on the host
    SFC_ADD_R3_BOTH(float)
#endif
}
}  // namespace sfc
