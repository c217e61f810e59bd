"""scirs_b200 — B200-native drop-in for the data-parallel FFT hot path of scirs2-fft.

Host-side mirror of the reference's interface for that path (same names,
argument meaning and error behaviour); all arithmetic happens in
``lib/libscirs2_fft_cuda.so`` (hand-written sm_100a kernels behind a C ABI,
``include/scirs2_fft_cuda.h``).  No CPU fallback.
"""
from .error import (FFTError, ComputationError, DimensionError, ValueError_, NotImplementedError_, BackendError,
                    PlanError, CommunicationError, MemoryError_)
from .fft import (fft, ifft, rfft, irfft, fft2, ifft2, fft2_parallel, ifft2_parallel, rfft2, irfft2, fftn, ifftn,
                  rfftn, irfftn, fft_strided, fft_strided_complex, ifft_strided, fft_simd, ifft_simd, fft_adaptive,
                  ifft_adaptive, fft2_simd, fft2_adaptive, fftn_simd, fftn_adaptive, ifft2_simd, ifftn_simd,
                  rfft_simd, irfft_simd, rfft_adaptive, irfft_adaptive, rfft_batch, irfft_batch)
from .consumers import (DCTType, DSTType, dct, idct, dct2, idct2, dctn, idctn, dst, idst, dst2, idst2, dstn, idstn,
                        dht, idht, dht2, fht, hfft, ihfft, hilbert, get_window, stft, spectrogram, FftMode, fft_inplace,
                        process_in_chunks, fft2_efficient, fft_streaming, fftn_optimized, fftn_memory_efficient,
                        rfftn_optimized)
from .czt import CZT, czt, czt_points, zoom_fft
from . import signal  # the scirs2-signal callers keep their own namespace (their stft / spectrogram differ from scirs2-fft's)
from .plan import FftPlan, FftPlanExecutor
from .plan_serialization import (PlanInfo, PlanMetrics, PlanDatabaseStats, PlanSerializationManager,
                                 create_and_time_plan)
from .auto_tuning import (AutoTuner, AutoTuneConfig, BenchmarkResult, FftVariant, SizeRange, SizeStep, SystemInfo,
                          TuningDatabase, GpuPlanTuner)
from .plan_cache import PlanCache, CacheStats, get_global_cache
from .backend import FftBackend, CudaFftBackend, BackendManager, BackendContext, get_backend_manager
from .context import (WorkerConfig, WorkerPool, WorkerPoolInfo, get_global_pool, set_workers, get_workers, FftContext,
                      FftContextBuilder, fft_context, with_fft_settings, with_backend, with_workers, without_cache)
from .planning import (PlannerBackend, PlanningStrategy, AdvancedFftPlanner, PlanBuilder, ParallelExecutor,
                       ParallelPlanner, get_global_planner, plan_ahead_of_time, AdaptivePlanningConfig, AdaptivePlanner,
                       AdaptiveExecutor)

__all__ = [n for n in dir() if not n.startswith("_")]
