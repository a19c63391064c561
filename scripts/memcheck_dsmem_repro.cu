// Minimal reproduction for the compute-sanitizer memcheck report on bigru_tc_kernel ("Invalid __shared__ write of
// size 2048 bytes ... Access at 0x.... is not located in remote CTA" at its cp.async.bulk shared::cta -> shared::cluster
// push, profiles/r2_sanitizer_summary_v2.txt).  A cluster of 8 CTAs; each CTA fills 2 KB of its own shared memory and
// bulk-copies it into slot `rank` of EVERY CTA's buffer, itself included (addresses from mapa), completing on the
// receiver's per-slot mbarrier; each CTA then checks the 8 blocks it received -- the Bi-GRU's exchange pattern.  The kernel is correct by construction (the host verifies the result); if memcheck flags the same
// error here, the report is a tool limitation for this instruction form, not a defect of the Bi-GRU kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/dsmem_repro scripts/memcheck_dsmem_repro.cu
//   /tmp/dsmem_repro && compute-sanitizer --tool memcheck /tmp/dsmem_repro
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int CL = 8;                       // CTAs per cluster, as bigru_tc_kernel
__global__ void __cluster_dims__(CL, 1, 1) push_kernel(int* ok) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint32_t* dst = reinterpret_cast<uint32_t*>(smem);                      // [CL] blocks of 2 KB, block r written by CTA r
  uint32_t* src = reinterpret_cast<uint32_t*>(smem + CL * 2048);          // my 2 KB block
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (CL + 1) * 2048);    // [CL] one barrier per incoming block
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 512; i += blockDim.x) src[i] = (rank << 16) | (uint32_t)i;
  if (threadIdx.x < CL) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + threadIdx.x)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 2048;" ::"r"(smem_u32(bar + threadIdx.x)) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < CL) {                     // thread t pushes my block to CTA (rank + t) % CL (t = 0: myself)
    const uint32_t peer = (rank + threadIdx.x) % CL;
    uint32_t rdst, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(smem_u32(dst + rank * 512)), "r"(peer));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(bar + rank)), "r"(peer));
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rdst), "r"(smem_u32(src)), "r"(2048u), "r"(rbar) : "memory");
  }
  int good = 1;
  for (int r = 0; r < CL; ++r) {              // wait for every CTA's block, check it
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar + r)) : "memory");
    for (int i = threadIdx.x; i < 512; i += blockDim.x) good &= dst[r * 512 + i] == (((uint32_t)r << 16) | (uint32_t)i);
  }
  if (!good) atomicExch(ok, 0);
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  int* ok;
  cudaMallocManaged(&ok, sizeof(int));
  *ok = 1;
  cudaFuncSetAttribute(push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 45056);
  push_kernel<<<2 * CL, 320, 45056>>>(ok);
  cudaError_t e = cudaDeviceSynchronize();
  printf("dsmem bulk push repro: launch %s, data %s\n", cudaGetErrorString(e), *ok ? "correct" : "WRONG");
  return (e == cudaSuccess && *ok) ? 0 : 1;
}
