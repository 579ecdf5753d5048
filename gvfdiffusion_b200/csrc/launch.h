// launch.h -- kernel launch with the programmatic-dependent-launch attribute (PDL).
//
// One NFE of the DiT is ~260 dependent kernels of 10-300 us; with plain stream order every boundary pays
// the drain of the previous grid, the launch latency and the next kernel's prologue (barrier init, TMEM
// allocation, tensor-map fetch) back to back.  Kernels launched through launch_pdl() may start their
// prologue while the previous grid is still draining; they call pdl_wait() (tc_common.cuh) before touching
// global memory, so the data dependence is unchanged.  The attribute is kept by stream capture (a
// programmatic edge in the CUDA graph).
// MEASURED (B200, bench.py, graph replay): it does not pay here -- 296.3 ms / object with the attribute
// against 290.3 ms without (trigger at kernel start: 298.9 ms).  Graph replay already hides the launch
// latency and the early-resident dependent CTAs only add scheduling work.  So the default is OFF
// (gvf_set_pdl(1) / bench.py --pdl turn it on for A/B runs); the kernels keep their griddepcontrol.wait,
// which returns immediately for a normally launched grid.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace gvf {

extern int g_pdl_enabled;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace gvf
