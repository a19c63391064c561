// common.cuh -- shared helpers for libsarnet_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/sarnet.h"

namespace sar {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return SAR_OK;
}

#define SAR_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      sar::set_error(__VA_ARGS__);              \
      return (code);                            \
    }                                           \
  } while (0)

// ---- programmatic dependent launch (PDL): every kernel of the step is launched with
// programmaticStreamSerialization, so its CTAs may become resident (and run their prologue: barrier init, TMEM
// allocation, constant weights -> shared memory) while the previous kernel of the stream drains.  Device side:
// pdl_wait() before the first access to anything a previous kernel wrote (or that it may still read), then
// pdl_trigger() so the NEXT kernel's prologue overlaps this kernel's body.  SAR_NO_PDL=1 disables the attribute
// (the device instructions are no-ops then).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Cooperative launch: the driver guarantees that EVERY CTA of the grid is resident at once (or refuses the launch)
// -- required by kernels whose CTAs wait on each other (conv_tc_chain_kernel spins on tile counters that other CTAs
// of the same launch publish).  A cooperative launch is a full dependency on the previous kernel (no programmatic
// overlap), which the chain's griddepcontrol instructions tolerate (they are no-ops then).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_coop(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Dynamic shared memory above 48 KB is an opt-in per (device, function).  It is raised ONCE to the device maximum and
// never lowered: a per-launch cudaFuncSetAttribute(exact size) left the function's attribute at whatever the LAST
// launch of that instantiation needed, and tools that re-launch a captured graph's kernel nodes one by one (ncu's
// graph-node profiling) then saw LaunchFailed for nodes captured with a larger size (ncu_rc = 9).
int allow_max_smem_impl(const void* fn, const char* what);
template <typename... KArgs>
inline int allow_max_smem(void (*kern)(KArgs...), const char* what) {
  return allow_max_smem_impl(reinterpret_cast<const void*>(kern), what);
}
constexpr size_t SAR_MAX_DYN_SMEM = 227 * 1024;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// NaN-propagating max / ReLU (max.NaN.f32): fmaxf returns the non-NaN operand, which would turn an fp16 overflow of the
// tensor-core operand planes (|x| >= 65504 -> inf in the hi plane, inf - inf = NaN in the accumulator) into silent zeros
// at the next ReLU / max-pool.  With these the NaN reaches the model outputs, where the host checks for it (model._finite).
__device__ __forceinline__ float fmax_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float relu_nan(float a) { return fmax_nan(a, 0.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide reductions through a small smem scratch (>= 32 floats); all threads get the result
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace sar
