// kernels_r3.cu — power-of-three tiles (r3_tile.cuh): one wide and one narrow tile per length, f64 and f32
#include "r3_tile.cuh"
namespace sfc {
#define SFC_ADD_R3(T, L, TL) add(::sfc::R3Inst<T, L, TL>::entry());
void register_kernels_r3(void (*add)(const KernelEntry&)) {
    SFC_ADD_R3(double, 9, 729) SFC_ADD_R3(double, 9, 243)
    SFC_ADD_R3(double, 27, 243) SFC_ADD_R3(double, 27, 81)
    SFC_ADD_R3(double, 81, 81) SFC_ADD_R3(double, 81, 27)
    SFC_ADD_R3(double, 243, 27) SFC_ADD_R3(double, 243, 9)
    SFC_ADD_R3(double, 729, 9) SFC_ADD_R3(double, 729, 3)
    SFC_ADD_R3(double, 2187, 3) SFC_ADD_R3(double, 2187, 1)
#ifndef SFC_HOST_EMUL  // the f32 path uses packed f32x2 PTX: not emulated on the host
    SFC_ADD_R3(float, 9, 729) SFC_ADD_R3(float, 9, 243)
    SFC_ADD_R3(float, 27, 243) SFC_ADD_R3(float, 27, 81)
    SFC_ADD_R3(float, 81, 81) SFC_ADD_R3(float, 81, 27)
    SFC_ADD_R3(float, 243, 27) SFC_ADD_R3(float, 243, 9)
    SFC_ADD_R3(float, 729, 9) SFC_ADD_R3(float, 729, 3)
    SFC_ADD_R3(float, 2187, 3) SFC_ADD_R3(float, 2187, 1)
#endif
}
}  // namespace sfc
