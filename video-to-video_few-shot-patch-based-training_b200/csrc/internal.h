// internal.h — shared host-side helpers of libpbt (not part of the C-ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/pbt.h"

namespace pbt {

void set_last_error(const char* msg);
int cuda_fail(cudaError_t e, const char* where);

#define PBT_CUDA_CHECK(expr)                                    \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return pbt::cuda_fail(_e, #expr);    \
  } while (0)

#define PBT_REQUIRE(cond, msg)       \
  do {                               \
    if (!(cond)) {                   \
      pbt::set_last_error(msg);      \
      return PBT_ERR_ARG;            \
    }                                \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency, so the
// library loads on a machine without a driver).
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode_tiled();

// 4-D tensor map over a P8 view: dims (w*8, h, planes, n), 16-bit elements, no swizzle, zero OOB fill.
int make_p8_tmap(CUtensorMap* map, const pbt_act_t& t, int box_w_px, int box_h, int box_planes);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline uint32_t round_up(uint32_t a, uint32_t b) { return (a + b - 1) / b * b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int num_sms();

// Programmatic dependent launch: every kernel of the library is launched with the "programmatic stream serialization"
// attribute and starts with griddepcontrol.launch_dependents / griddepcontrol.wait (pdl_sync), so the grid of launch
// i+1 is scheduled onto the SMs while launch i drains and only its memory accesses wait for launch i to complete.  The
// training step is ~200 short dependent launches; this hides the launch + block-scheduling latency between them (also
// inside a captured CUDA graph, where the edges become programmatic dependencies).  PBT_PDL=0 turns the attribute off.
bool pdl_enabled();

template <typename... Exp, typename... Act>
inline void launch(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Act>(args)...);   // errors surface in the caller's cudaGetLastError()
}

#ifdef __CUDACC__
// let the next launch of the stream start scheduling, then wait until every launch this one depends on has completed and
// its memory is visible.  Nothing before this point may touch global memory written by earlier launches.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

}  // namespace pbt
