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
    SFC_ADD_PIPE_LATE(double, 4096, 1)
    SFC_ADD_PIPE_LATE(double, 8192, 1)
    SFC_ADD_PIPE_LATE(double, 2048, 2)
    SFC_ADD_PIPE_LATE(double, 2048, 1)
    SFC_ADD_PIPE_LATE(float, 8192, 1)
    // strided (column) tiles of the N-D axis passes, four-step and Bluestein passes: landed through a tensor map
    SFC_ADD_PIPE_LATE(double, 64, 64)
    SFC_ADD_PIPE_LATE(double, 128, 32)
    SFC_ADD_PIPE_LATE(double, 256, 16)
    SFC_ADD_PIPE_LATE(double, 512, 8)
    SFC_ADD_PIPE_LATE(double, 1024, 4)
    SFC_ADD_PIPE_LATE(double, 2048, 4)
    SFC_ADD_GPIPE(double, 4096, 2, false)
    SFC_ADD_GPIPE(double, 2048, 4, false)
    SFC_ADD_GPIPE(double, 1024, 8, false)
}
}  // namespace sfc
