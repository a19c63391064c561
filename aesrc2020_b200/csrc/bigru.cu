// bigru.cu -- recurrent step loop of Bidirectional(CuDNNGRU) (model.py:44-50).
//
// The input projections x*W + b_input of all S steps and both directions are one GEMM done by
// the caller (sar_conv2d_fwd against the concatenated [forward | backward] kernels); this kernel
// runs the S strictly sequential steps
//     hp = h @ U + b_rec                                   (u x 3u, gate order z|r|h)
//     z = sigmoid(xz + hpz)   r = sigmoid(xr + hpr)   hh = tanh(xh + r * hph)
//     h' = z*h + (1-z)*hh                                  (reset_after / cuDNN form)
// as ONE persistent launch.  A thread-block CLUSTER of 8 CTAs owns 8 utterances of one direction
// for the whole sequence:
//   * the recurrent kernel U (256 x 768 fp32 = 786 KB) is split by hidden unit: CTA r owns units
//     [32r, 32r+32) = 96 columns (z, r and h of each unit) and keeps its 256 x 96 slice in
//     REGISTERS for all S steps (thread = 4 columns x 16 k's = 64 registers), so no weight byte
//     moves after the prologue;
//   * h (8 x 256) lives in every CTA's shared memory, double buffered, laid out so that the 16
//     k-slices of a warp read conflict-free 128-bit rows; a thread reads 16 rows x 8 utterances
//     (512 B) per step for 512 FMAs -- shared-memory and FMA pipes are balanced;
//   * the 16 k-slice partials are combined by a recursive-halving shuffle reduction (30 shuffles
//     for 32 values, each lane ends owning 2 complete sums);
//   * after the gates each CTA pushes its 32 new units into all 8 CTAs' next h buffer through
//     distributed shared memory; a split cluster barrier (arrive.release ... wait.acquire)
//     publishes the step while the x-projections of step t+2 are being prefetched;
//   * outputs are staged in shared memory and written to HBM once, coalesced, after the loop (no
//     global store sits in front of the per-step release).
// Both directions and all utterance groups run concurrently (grid = 8 * groups * 2 CTAs).
#include <cooperative_groups.h>
#include <stdlib.h>
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace sar {

#ifdef SAR_GRU_PROFILE       // SAR_NVCC_EXTRA=-DSAR_GRU_PROFILE: thread 0 of CTA 0 prints cycle stamps of steps 8..11
#define GRU_STAMP(k) if (blockIdx.x == 0 && t == 0 && step >= 8 && step < 12) stamps[(step - 8) * 8 + (k)] = clock64();
#else
#define GRU_STAMP(k)
#endif

constexpr int GRU_U = 256;
constexpr int GRU_CL = 8;                        // CTAs per cluster
constexpr int GRU_UPC = GRU_U / GRU_CL;          // 32 hidden units per CTA
constexpr int GRU_COLS = 3 * GRU_UPC;            // 96 recurrent columns per CTA
constexpr int GRU_CPT = 4;                       // columns per thread
constexpr int GRU_KQ = 16;                       // k-slices
constexpr int GRU_KPT = GRU_U / GRU_KQ;          // 16 k's per thread
constexpr int GRU_THREADS = (GRU_COLS / GRU_CPT) * GRU_KQ;   // 384
// BG = utterances per cluster (template parameter: 8 or 12).  A B200 can keep only 15 clusters of 8 CTAs
// resident (cudaOccupancyMaxActiveClusters), so B = 64 with BG = 8 (16 clusters) ran as TWO waves; BG = 12
// gives 12 clusters in one wave at 1.5x the FMA work per step.
constexpr int GRU_MAX_CLUSTERS = 15;

// shared::cta address -> the same location in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// asynchronous remote store that also counts 4 bytes on the destination CTA's mbarrier: the data and its
// "arrived" signal travel together, so a step needs no cluster-wide barrier (arrive.release alone cost ~1300 cycles)
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int GRU_BG>
__global__ void __cluster_dims__(GRU_CL, 1, 1) __launch_bounds__(GRU_THREADS, 1)
bigru_kernel(const float* __restrict__ xp, const float* __restrict__ rec, const float* __restrict__ rbias,
             float* __restrict__ out, int B, int S, int seq, int stage_out) {
  constexpr int U = GRU_U, U3 = 3 * GRU_U;
  constexpr int GRU_HSTRIDE = GRU_KPT * GRU_BG + 4;            // +4 banks per k-slice: conflict-free LDS.128
  constexpr int GRU_HBUF = GRU_KQ * GRU_HSTRIDE;               // floats per h buffer
  constexpr int NV = GRU_CPT * GRU_BG;                         // partial sums per thread
  constexpr int NOWN = NV / GRU_KQ;                            // complete sums a lane owns after the reduction
  static_assert(GRU_BG % 4 == 0 && NV % GRU_KQ == 0 && GRU_UPC * GRU_BG <= GRU_THREADS, "unsupported utterance group");
  __shared__ __align__(16) float h_s[2 * GRU_HBUF];                 // [buf][kq][i][b]
  __shared__ __align__(16) float hp_s[GRU_COLS][GRU_BG];
  __shared__ __align__(8) uint64_t hbar[2];                          // h buffer b has received all 8 CTAs' slices
  extern __shared__ __align__(16) float out_s[];                    // [S][gb][gj] when stage_out

#ifdef SAR_GRU_PROFILE
  const long long t_entry = clock64();
  long long t_glob0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_glob0));
#endif
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / GRU_CL;
  const int dir = cid & 1;
  const int b0 = (cid >> 1) * GRU_BG;
  const int t = threadIdx.x, lane = t & 31;
  const int j0 = rank * GRU_UPC;

  // ---- prologue: my 4 columns x 16 k's of the recurrent kernel -> registers
  const int kq = lane & 15;                               // k-slice
  const int cgp = (t >> 5) * 2 + (lane >> 4);             // column group 0..23
  const float* Ud = rec + (size_t)dir * U * U3;
  float u[GRU_CPT][GRU_KPT];
#pragma unroll
  for (int c = 0; c < GRU_CPT; ++c) {
    const int cc = cgp * GRU_CPT + c;                     // column within CTA: gate = cc/32, unit = cc%32
    const int col = (cc / GRU_UPC) * U + j0 + (cc % GRU_UPC);
#pragma unroll
    for (int i = 0; i < GRU_KPT; ++i) u[c][i] = __ldg(Ud + (size_t)(kq * GRU_KPT + i) * U3 + col);
  }
  // after the reduction lane kq owns the NOWN sums with index kq*NOWN + j in the thread's [column][utterance] space
  const int red_col = cgp * GRU_CPT + (kq * NOWN) / GRU_BG;
  const int red_b = (kq * NOWN) % GRU_BG;

  // gate-phase role: thread -> (unit gj, utterance gb)
  const bool gate_thread = t < GRU_UPC * GRU_BG;          // 256 threads
  const int gj = t / GRU_BG, gb = t % GRU_BG;            // gb fastest: a warp's DSMEM push covers whole h rows
  const int unit = j0 + gj;
  const bool bvalid = gate_thread && (b0 + gb) < B;
  float rbz = 0.f, rbr = 0.f, rbh = 0.f;
  if (gate_thread) {
    rbz = __ldg(rbias + dir * U3 + unit);
    rbr = __ldg(rbias + dir * U3 + U + unit);
    rbh = __ldg(rbias + dir * U3 + 2 * U + unit);
  }
  const int hpos = (unit / GRU_KPT) * GRU_HSTRIDE + (unit % GRU_KPT) * GRU_BG + gb;

  for (int i = t; i < 2 * GRU_HBUF; i += GRU_THREADS) h_s[i] = 0.f;
  if (t == 0) {
    mbar_init(&hbar[0], 1); mbar_init(&hbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t h_u = smem_u32(h_s), hbar_u = smem_u32(hbar);
  constexpr uint32_t STEP_BYTES = (uint32_t)(GRU_CL * GRU_UPC * GRU_BG * sizeof(float));   // one step's h' from all 8 CTAs

  pdl_wait();            // everything above read only the recurrent weights
  pdl_trigger();
  auto xrow = [&](int step) {
    const int tt = dir ? (S - 1 - step) : step;
    return xp + (((size_t)(b0 + gb) * S + tt) * 2 + dir) * U3 + unit;
  };
  // x-projection operands of steps `step` (x0) and `step+1` (x1) live in registers
  float x0z = 0.f, x0r = 0.f, x0h = 0.f, x1z = 0.f, x1r = 0.f, x1h = 0.f;
  if (bvalid) {
    const float* p = xrow(0); x0z = __ldg(p); x0r = __ldg(p + U); x0h = __ldg(p + 2 * U);
    if (S > 1) { const float* q = xrow(1); x1z = __ldg(q); x1r = __ldg(q + U); x1h = __ldg(q + 2 * U); }
  }
  cluster.sync();

  int cur = 0;
#ifdef SAR_GRU_PROFILE
  long long stamps[32];
  const long long t_loop = clock64();
#endif
  for (int step = 0; step < S; ++step) {
    // h(step) is complete in my buffer `cur` once all 8 CTAs' step-1 gate threads have pushed (step 0 reads zeros);
    // buffer cur^1 is then armed for h(step+1).  Double buffering + this data dependency make the h buffers race
    // free: nobody can push h(step+1) before every CTA has pushed h(step), i.e. finished reading h(step-1).
    if (step > 0) mbar_wait(&hbar[cur], (uint32_t)((step - 1) >> 1) & 1u);
    if (t == 0 && step + 1 < S) mbar_expect_tx(&hbar[cur ^ 1], STEP_BYTES);
    GRU_STAMP(0)
    // ---- partial dot products: 4 columns x 16 k's x 8 utterances
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.f;
    const float* hq = h_s + cur * GRU_HBUF + kq * GRU_HSTRIDE;
#pragma unroll
    for (int i = 0; i < GRU_KPT; ++i) {
      float hv[GRU_BG];
#pragma unroll
      for (int q = 0; q < GRU_BG / 4; ++q) {
        const float4 h4 = *reinterpret_cast<const float4*>(hq + i * GRU_BG + 4 * q);
        hv[4 * q] = h4.x; hv[4 * q + 1] = h4.y; hv[4 * q + 2] = h4.z; hv[4 * q + 3] = h4.w;
      }
#pragma unroll
      for (int c = 0; c < GRU_CPT; ++c)
#pragma unroll
        for (int b = 0; b < GRU_BG; ++b) v[c * GRU_BG + b] = fmaf(u[c][i], hv[b], v[c * GRU_BG + b]);
    }
    GRU_STAMP(1)
    // ---- recursive-halving reduction over the 16 k-slices (lanes kq = lane & 15): at every level a lane keeps
    // the half of its index range selected by the corresponding bit of kq and adds the partner's copy of it
    float w1[NV / 2], w2[NV / 4], w3[NV / 8], w4[NV / 16];
    {
      const bool up = (kq & 8) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        const float send = up ? v[i] : v[i + NV / 2], keep = up ? v[i + NV / 2] : v[i];
        w1[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    {
      const bool up = (kq & 4) != 0;
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        const float send = up ? w1[i] : w1[i + NV / 4], keep = up ? w1[i + NV / 4] : w1[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    }
    {
      const bool up = (kq & 2) != 0;
#pragma unroll
      for (int i = 0; i < NV / 8; ++i) {
        const float send = up ? w2[i] : w2[i + NV / 8], keep = up ? w2[i + NV / 8] : w2[i];
        w3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
    }
    {
      const bool up = (kq & 1) != 0;
#pragma unroll
      for (int i = 0; i < NV / 16; ++i) {
        const float send = up ? w3[i] : w3[i + NV / 16], keep = up ? w3[i + NV / 16] : w3[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
    }
    GRU_STAMP(2)
#pragma unroll
    for (int j = 0; j < NOWN; ++j) hp_s[red_col][red_b + j] = w4[j];
    __syncthreads();
    GRU_STAMP(3)

    // ---- gates for my 32 units x 8 utterances, push h' to every CTA of the cluster
    if (gate_thread) {
      const float z = sigmoidf_(x0z + hp_s[gj][gb] + rbz);
      const float r = sigmoidf_(x0r + hp_s[GRU_UPC + gj][gb] + rbr);
      const float hh = tanhf(x0h + r * (hp_s[2 * GRU_UPC + gj][gb] + rbh));
      const float hold = h_s[cur * GRU_HBUF + hpos];
      const float hn = z * hold + (1.f - z) * hh;
      if (step + 1 < S) {
        const uint32_t dst = h_u + (uint32_t)(((cur ^ 1) * GRU_HBUF + hpos) * sizeof(float));
        const uint32_t bar = hbar_u + (uint32_t)((cur ^ 1) * sizeof(uint64_t));
#pragma unroll
        for (int r2 = 0; r2 < GRU_CL; ++r2) st_async_f32(mapa_u32(dst, (uint32_t)r2), hn, mapa_u32(bar, (uint32_t)r2));
      }
      if (stage_out) {
        if (seq & 1) out_s[(size_t)step * (GRU_UPC * GRU_BG) + t] = hn;
        else if (step == S - 1) out_s[t] = hn;
      } else if (bvalid) {
        const int tt = dir ? (S - 1 - step) : step;
        if (seq) out[((size_t)(b0 + gb) * S + tt) * (2 * U) + dir * U + unit] = hn;
        else if (step == S - 1) out[(size_t)(b0 + gb) * (2 * U) + dir * U + unit] = hn;
      }
    }
    GRU_STAMP(4)
    // rotate the x-projection registers and prefetch step+2 while the pushes are in flight
    x0z = x1z; x0r = x1r; x0h = x1h;
    if (bvalid && step + 2 < S) { const float* p = xrow(step + 2); x1z = __ldg(p); x1r = __ldg(p + U); x1h = __ldg(p + 2 * U); }
    GRU_STAMP(5)
    GRU_STAMP(6)
    cur ^= 1;
  }
  cluster.sync();                        // nobody leaves while a peer could still address its shared memory
#ifdef SAR_GRU_PROFILE
  const long long t_loop_end = clock64();
#endif
#ifdef SAR_GRU_PROFILE
  if (blockIdx.x == 0 && t == 0)
    for (int i = 0; i < 4; ++i)
      printf("gru step %d: fma %lld  reduce %lld  sts+bar %lld  gates+push %lld  arrive+prefetch %lld  wait %lld | total %lld\n", 8 + i,
             stamps[i * 8 + 1] - stamps[i * 8], stamps[i * 8 + 2] - stamps[i * 8 + 1], stamps[i * 8 + 3] - stamps[i * 8 + 2],
             stamps[i * 8 + 4] - stamps[i * 8 + 3], stamps[i * 8 + 5] - stamps[i * 8 + 4], stamps[i * 8 + 6] - stamps[i * 8 + 5],
             stamps[i * 8 + 6] - stamps[i * 8]);
#endif

  // ---- write the staged outputs: 32 consecutive units (128 B) per (step, utterance)
  if (stage_out && gate_thread) {
    const int oj = t & 31, ob = t >> 5;                    // unit fastest for coalesced global stores (ob < GRU_BG)
    if (b0 + ob < B) {
      if (seq) {
        for (int step = 0; step < S; ++step) {
          const int tt = dir ? (S - 1 - step) : step;
          out[((size_t)(b0 + ob) * S + tt) * (2 * U) + dir * U + j0 + oj] =
              out_s[(size_t)step * (GRU_UPC * GRU_BG) + oj * GRU_BG + ob];
        }
      } else {
        out[(size_t)(b0 + ob) * (2 * U) + dir * U + j0 + oj] = out_s[oj * GRU_BG + ob];
      }
    }
  }
#ifdef SAR_GRU_PROFILE
  if ((blockIdx.x == 0 || blockIdx.x == 127) && t == 0) {
    long long t_glob1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_glob1));
    printf("gru cta %d: prologue %lld  loop %lld  epilogue %lld cycles; globaltimer %lld ns\n", (int)blockIdx.x, t_loop - t_entry, t_loop_end - t_loop,
           clock64() - t_loop_end, t_glob1 - t_glob0);
  }
#endif
}

}  // namespace sar

namespace sar { int bigru_tc_launch(const float*, const float*, const float*, float*, int, int, int, int, cudaStream_t); }

extern "C" int sar_bigru_fwd(const float* xp, const float* rec, const float* rbias, float* out,
                             int B, int S, int u, int seq, void* stream) {
  return sar_bigru_nb_fwd(xp, rec, rbias, out, B, S, u, seq, 0, stream);
}

extern "C" int sar_bigru_nb_fwd(const float* xp, const float* rec, const float* rbias, float* out,
                                int B, int S, int u, int seq, int nb, void* stream) {
  using namespace sar;
  SAR_REQUIRE(nb == 0 || nb == 16 || nb == 32, SAR_ERR_BAD_ARG, "sar_bigru_nb_fwd: nb must be 0 (auto), 16 or 32");
  SAR_REQUIRE(xp && rec && rbias && out, SAR_ERR_BAD_ARG, "sar_bigru_fwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0, SAR_ERR_BAD_ARG, "sar_bigru_fwd: non-positive dimension");
  SAR_REQUIRE(u == GRU_U, SAR_ERR_UNSUPPORTED, "sar_bigru_fwd: hidden size %d unsupported (this build: %d)", u, GRU_U);
  // default: the tcgen05 recurrence (bigru_tc.cu); SAR_GRU_FFMA=1 selects the CUDA-core kernel below
  static const bool use_ffma = getenv("SAR_GRU_FFMA") != nullptr;
  if (!use_ffma) {
    SAR_REQUIRE(aligned16(xp) && aligned16(rbias) && aligned16(out), SAR_ERR_BAD_ARG, "sar_bigru_fwd: pointers must be 16-byte aligned");
    return bigru_tc_launch(xp, rec, rbias, out, B, S, seq, nb, (cudaStream_t)stream);
  }
  // utterances per cluster: fewest waves of at most GRU_MAX_CLUSTERS resident clusters, weighted by the measured
  // step time of each variant (~3.9k cycles at 8 utterances, ~4.7k at 12)
  auto waves = [](int B_, int bg) { const int cl = 2 * ((B_ + bg - 1) / bg); return (cl + GRU_MAX_CLUSTERS - 1) / GRU_MAX_CLUSTERS; };
  const int bg = (waves(B, 8) * 39 <= waves(B, 12) * 47) ? 8 : 12;
  const int groups = (B + bg - 1) / bg;
  dim3 grid(GRU_CL * groups * 2);
  size_t smem = seq ? (size_t)S * GRU_UPC * bg * sizeof(float) : (size_t)GRU_UPC * bg * sizeof(float);
  int stage_out = 1;
  if (smem > 160 * 1024) { smem = 0; stage_out = 0; }      // very long sequences: write through
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(160 * 1024));
    if (e != cudaSuccess) { set_error("sar_bigru_fwd: %s", cudaGetErrorString(e)); return (int)e; }
    launch_k(kern, dim3(grid), dim3(GRU_THREADS), smem, (cudaStream_t)stream, xp, rec, rbias, out, B, S, seq, stage_out);
    return 0;
  };
  const int lrc = bg == 8 ? launch(bigru_kernel<8>) : launch(bigru_kernel<12>);
  if (lrc) return lrc;
  return check_launch("sar_bigru_fwd");
}

// diagnostic (not part of include/sarnet.h): how many 8-CTA clusters of bigru_kernel can be resident at once
extern "C" int sar_debug_bigru_max_clusters(int smem_bytes) {
  using namespace sar;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(GRU_CL * 64); cfg.blockDim = dim3(GRU_THREADS); cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = GRU_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaFuncSetAttribute(bigru_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(160 * 1024));
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, bigru_kernel<12>, &cfg);
  return e == cudaSuccess ? n : -(int)e;
}
