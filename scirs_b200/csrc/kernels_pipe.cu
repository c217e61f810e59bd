// kernels_pipe.cu — persistent TMA-pipelined flavour (TM_PIPE_C2C) of the 64 KiB complex tiles
#include "kernel_inst.cuh"
namespace sfc {
void register_kernels_pipe(void (*add)(const KernelEntry&)) {
    SFC_ADD_PIPE(double, 4096, 1, false)
    SFC_ADD_PIPE(double, 2048, 2, false)
    SFC_ADD_PIPE(double, 1024, 4, false)
    SFC_ADD_PIPE(double, 512, 8, false)
    SFC_ADD_PIPE(double, 256, 16, false)
    SFC_ADD_PIPE(double, 8192, 1, false)
    SFC_ADD_GPIPE(double, 4096, 2, false)
    SFC_ADD_GPIPE(double, 2048, 4, false)
    SFC_ADD_GPIPE(double, 1024, 8, false)
}
}  // namespace sfc
