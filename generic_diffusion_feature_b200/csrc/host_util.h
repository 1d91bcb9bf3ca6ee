// Host-side helpers: error plumbing and TMA tensor-map construction (driver entry point fetched at run
// time through cudart so the library does not link libcuda at build time).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include "../../include/gdf.h"

namespace gdf {

// Error codes GDF_OK / GDF_ERR_* come from the C ABI header (include/gdf.h).

std::string& last_error();  // thread-local message of the last failure
int fail(int code, const char* fmt, ...);

#define GDF_CUDA(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::gdf::fail(GDF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,       \
                         cudaGetErrorString(_e));                                                \
  } while (0)

#define GDF_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

// Build a tiled bf16 tensor map. dims/strides innermost first; strides_bytes has rank-1 entries
// (stride of dims 1..rank-1); box innermost first. swizzle_bytes in {128, 64, 0}; zero fill out of bounds.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes = 128);

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining
// (its CTAs run their prologue and then block in pdl_wait() until the predecessor has completed and flushed), which
// hides the launch gap + prologue of the ~1000 back-to-back kernels of one extraction step. Every kernel launched this
// way calls pdl_wait() before touching global memory. GDF_PDL=0 turns the attribute off (A/B timing).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace gdf
