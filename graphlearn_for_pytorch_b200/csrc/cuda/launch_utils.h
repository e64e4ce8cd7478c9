// Host-side launch helper: every engine kernel goes through launch_k(), which sets the programmatic-stream-
// serialization attribute (PDL, see device_utils.cuh) unless GLT_B200_PDL=0.  Under stream capture consecutive
// launches become programmatic edges of the CUDA graph.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

namespace glt {

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("GLT_B200_PDL");
    return e ? std::atoi(e) != 0 : true;
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace glt
