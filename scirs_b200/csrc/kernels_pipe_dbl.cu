// kernels_pipe_dbl.cu — TM_PIPE_C2C flavour of the fused forward * table * inverse tiles (Bluestein pass B)
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_pipe_dbl(void (*add)(const KernelEntry&)) {
    SFC_ADD_PIPE(double, 4096, 1, true)
    SFC_ADD_PIPE(double, 2048, 2, true)
    SFC_ADD_PIPE(double, 1024, 4, true)
}
}  // namespace sfc
