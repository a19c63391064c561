// bigru_tc.cu -- the Bidirectional(CuDNNGRU) recurrence (model.py:44-50) with the recurrent matmul on tcgen05.
//
// Same contract as bigru.cu (the caller has done the input projections x*W + b_input of all steps):
//     hp = h @ U + b_rec      z = sigmoid(xz + hpz)   r = sigmoid(xr + hpr)   hh = tanh(xh + r * hph)
//     h' = z*h + (1-z)*hh                                      (gate order z|r|h, reset_after / cuDNN form)
// A thread-block CLUSTER of 8 CTAs owns NB = 32 utterances of one direction for the whole sequence.
//   * CTA r owns hidden units [32r, 32r+32).  Its slice of U^T is the A operand of the step's MMA, resident in shared
//     memory for all S steps: rows m = gate*32 + j (96 live rows of M = 128), K = 256 = eight 64B-swizzled K-major
//     atoms of 32 k (atom a = the units CTA a owns), as fp16 hi and lo planes (U = hi + lo/2048) -- 128 KB, loaded once.
//   * h (NB x 256) is the B operand: every CTA holds all of it as fp16 hi/lo rows [h_hi (NB rows) ; h_lo (NB rows)]
//     per atom (4 KB, contiguous), double buffered (2 x 32 KB).  Per step one elected thread issues, per k16,
//         [acc0|acc1] (+)= A_hi x [h_hi ; h_lo]   (N = 2 NB)      acc1 += A_lo x h_hi   (N = NB)
//     (the 2^-22 A_lo x h_lo term is dropped, as in conv_tc.cu) -- 32 tcgen05.mma, ~1k cycles, against ~3k cycles
//     of FFMA + shuffle reduction in bigru.cu.
//   * 8 gate warps drain TMEM (lane = row = (gate, unit), columns = utterances), transpose hp through shared memory
//     to (utterance, 4 units) per thread, run the gates with the thread's h kept in registers, write h' to HBM, split
//     it to fp16 hi/lo into a 4 KB staging block laid out exactly like atom `rank` of a B buffer, and eight threads
//     PUSH that block into the next-step B buffer of the 8 CTAs with one cp.async.bulk (shared::cta ->
//     shared::cluster) each, completing on the destination's mbarrier: data and "arrived" signal travel together,
//     no cluster barrier in the step loop.  (First version: 4096 8-byte st.async per CTA and step -- 1450 cycles of
//     issue; the bulk copies take ~150.)  The MMA warp of every CTA waits on that mbarrier (32 KB per step) and
//     issues the next step.
//   * gates use ex2.approx-based sigmoid / tanh (|abs err| ~1e-7; the libm forms were 1290 of 5100 cycles per step).
// Both directions and all utterance groups run concurrently: grid = 8 * ceil(B/32) * 2 CTAs (B = 64: 32 SMs).
#include <cooperative_groups.h>
#include <stdlib.h>
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace sar {

#ifdef SAR_GRU_PROFILE       // SAR_NVCC_EXTRA=-DSAR_GRU_PROFILE: CTA 0 prints cycle stamps of steps 8..11
#define GT_STAMP(k) if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == GT_WARP_MMA) && step >= 8 && step < 12) stamps[(step - 8) * 8 + (k)] = clock64();
#else
#define GT_STAMP(k)
#endif

constexpr int GT_U = 256;
constexpr int GT_CL = 8;                          // CTAs per cluster
constexpr int GT_UPC = GT_U / GT_CL;              // 32 hidden units per CTA
constexpr int GT_GATE_THREADS = 256;              // thread -> (utterance, NB/8 consecutive units)
constexpr int GT_THREADS = GT_GATE_THREADS + 32;  // + the MMA-issuing warp
constexpr int GT_WARP_MMA = 8;
constexpr int GT_ROWB = 64;                       // bytes per h row of an atom: 32 k (fp16), SWIZZLE_64B
// tensor memory: columns [0, 2 NB) accumulators acc0 | acc1; [64, 192) U_hi; [192, 320) U_lo (lane = row m, a 32-bit
// column holds k = 2c, 2c+1 -> 8 columns per k16 slice).  The A operand never touches shared memory.
constexpr int GT_TMEM_COLS = 512;
constexpr int GT_TM_AHI = 64, GT_TM_ALO = 192;
// A B200 keeps at most 15 clusters of 8 CTAs resident: NB = 16 while both directions of the batch fit in one wave
// (B <= 112), NB = 32 beyond (half the clusters, twice the SM-to-SM bytes per CTA and step).
constexpr int GT_MAX_CLUSTERS = 15;
template <int NB> struct GtCfg {
  static constexpr int UPT = NB / 8;                        // units per gate thread (2 or 4)
  static constexpr int OPR = GT_UPC / UPT;                  // gate threads per utterance (16 or 8)
  static constexpr int B_ATOM = 2 * NB * GT_ROWB;           // [h_hi ; h_lo] rows x 32 k = one CTA's units
  static constexpr int B_BUF = GT_CL * B_ATOM;
  static constexpr size_t SMEM = 1024 + 2 * (size_t)B_BUF + 2 * (size_t)B_ATOM + 3 * NB * GT_UPC * sizeof(float) + 256;
};

__device__ __forceinline__ uint32_t gt_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// bulk copy from my shared memory into a peer CTA's, counting its bytes on the peer's mbarrier
__device__ __forceinline__ void bulk_push(uint32_t rdst, uint32_t src, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// tcgen05.mma with the A operand in tensor memory
__device__ __forceinline__ void umma_f16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ float gt_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gt_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }
// N consecutive floats (N = 2 or 4) through one vector access
template <int N> __device__ __forceinline__ void ldg_vec(const float* p, float (&v)[N]) {
  if constexpr (N == 4) { const float4 q = __ldg(reinterpret_cast<const float4*>(p)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
  else { const float2 q = __ldg(reinterpret_cast<const float2*>(p)); v[0] = q.x; v[1] = q.y; }
}
template <int N> __device__ __forceinline__ void ld_vec(const float* p, float (&v)[N]) {
  if constexpr (N == 4) { const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
  else { const float2 q = *reinterpret_cast<const float2*>(p); v[0] = q.x; v[1] = q.y; }
}
template <int N> __device__ __forceinline__ void st_vec(float* p, const float (&v)[N]) {
  if constexpr (N == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}

template <int NB>
__global__ void __cluster_dims__(GT_CL, 1, 1) __launch_bounds__(GT_THREADS, 1)
bigru_tc_kernel(const float* __restrict__ xp, const float* __restrict__ rec, const float* __restrict__ rbias,
                float* __restrict__ out, int B, int S, int seq) {
  using Cfg = GtCfg<NB>;
  constexpr int U = GT_U, U3 = 3 * GT_U, UPT = Cfg::UPT, OPR = Cfg::OPR, B_ATOM = Cfg::B_ATOM, B_BUF = Cfg::B_BUF;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // same offset in every CTA of the cluster
  uint8_t* b_base = smem;                                        // [2 buffers][8 atoms][2 NB rows][64 B]
  uint8_t* stage = b_base + 2 * B_BUF;                           // [2][2 NB rows][64 B]: my units' h' in atom layout
  float* hp_s = reinterpret_cast<float*>(stage + 2 * B_ATOM);    // [3 gates][NB][32 units]
  uint64_t* hbar = reinterpret_cast<uint64_t*>(hp_s + 3 * NB * GT_UPC);      // [2][8]: atom a of B buffer b has arrived
  uint64_t* mma_bar = hbar + 2 * GT_CL;                           // the step's accumulators are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / GT_CL;
  const int dir = cid & 1;
  const int b0 = (cid >> 1) * NB;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int j0 = rank * GT_UPC;

  // ---- prologue (reads only weights: runs under the previous kernel's tail with PDL)
  if (t == 0) {
    for (int i = 0; i < 2 * GT_CL; ++i) mbar_init(&hbar[i], 1);
    mbar_init(mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GT_WARP_MMA) tmem_alloc(tmem_slot, GT_TMEM_COLS);
  // both h buffers start as h(0) = 0
  for (int i = t; i < (2 * B_BUF) / 16; i += GT_THREADS) reinterpret_cast<uint4*>(b_base)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < GT_WARP_MMA) {
    // my row of U^T (row m = TMEM lane = gate*32 + j; rows 96..127 are zero) -> fp16 hi/lo pairs -> tensor memory:
    // warp (quad, half) fills lanes 32 quad.., k in [128 half, 128 half + 128), 16 k (8 columns) per tcgen05.st
    const int quad = warp & 3, half = warp >> 2;
    const float* Ud = rec + (size_t)dir * U * U3 + (quad < 3 ? quad * U + j0 + lane : 0);
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 2
    for (int ks = 8 * half; ks < 8 * half + 8; ++ks) {
      uint32_t hh[8], ll[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float w0 = 0.f, w1 = 0.f;
        if (quad < 3) { w0 = __ldg(Ud + (size_t)(ks * 16 + 2 * e) * U3); w1 = __ldg(Ud + (size_t)(ks * 16 + 2 * e + 1) * U3); }
        const __half2 h2 = __floats2half2_rn(w0, w1);
        const float2 hf = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn((w0 - hf.x) * 2048.f, (w1 - hf.y) * 2048.f);
        hh[e] = *reinterpret_cast<const uint32_t*>(&h2);
        ll[e] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      tmem_st8(trow + (uint32_t)(GT_TM_AHI + 8 * ks), hh);
      tmem_st8(trow + (uint32_t)(GT_TM_ALO + 8 * ks), ll);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t b_u = smem_u32(b_base), hbar_u = smem_u32(hbar);

  pdl_wait();
  pdl_trigger();
  cluster.sync();                                  // every CTA's barriers and h buffers are ready for remote pushes

  if (warp == GT_WARP_MMA) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * NB) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_n = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t acc0 = tmem_base, acc1 = tmem_base + (uint32_t)NB;
#ifdef SAR_GRU_PROFILE
    long long stamps[32];
#endif
    for (int step = 0; step < S; ++step) {
      GT_STAMP(0)
      const int cur = step & 1;
      // atom a of buffer cur^1 will receive CTA a's units of h(step+1); its previous phase (h(step-1)) completed
      // before step-1's MMAs.  A push may arrive before this arm (tx-count goes negative, the pending arrival keeps
      // the phase open).
      if (lane < GT_CL && step + 1 < S) mbar_expect_tx(&hbar[(cur ^ 1) * GT_CL + lane], (uint32_t)B_ATOM);
      __syncwarp();
      // Atom (rank - l) & 7 is polled by lane l (pushes are issued in that order: mine first, then rank-1's, ...);
      // every poll round issues the MMAs of the atoms that have landed since the last one, so the early atoms'
      // MMAs run while the later ones are still crossing the cluster.  (A blocking wait per atom cost ~190 cycles
      // of fixed overhead per atom on top of ~37 cycles per MMA.)
      // Atoms are ISSUED in a fixed order (fp32 accumulation order = bitwise reproducible results whatever the
      // arrival timing): after every poll round, the ready prefix of that order.
      uint32_t ready = 0;
      int next = 0;
      const uint32_t par = (uint32_t)((step - 1) >> 1) & 1u;
      const int my_atom = (rank - lane) & (GT_CL - 1);
      while (next < GT_CL) {
        bool r = false;
        if (lane < GT_CL && !((ready >> lane) & 1u)) r = (step == 0) || mbar_try_wait(&hbar[cur * GT_CL + my_atom], par);
        ready |= __ballot_sync(0xffffffffu, r);
        int upto = next;
        while (upto < GT_CL && ((ready >> upto) & 1u)) ++upto;
        if (upto == next) continue;
        if (next == 0) { GT_STAMP(1) }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int i = next; i < upto; ++i) {
            const int a = (rank - i) & (GT_CL - 1);
            const uint64_t db = make_desc(b_u + (uint32_t)(cur * B_BUF + a * B_ATOM), GT_ROWB);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t kcol = (uint32_t)(8 * (2 * a + kk));                         // k16 slice 2a + kk of U^T
              umma_f16_ta(acc0, tmem_base + GT_TM_AHI + kcol, db + 2u * kk, idesc_2n, (i | kk) ? 1u : 0u);   // [acc0|acc1] (+)= Uh x [hh;hl]
              umma_f16_ta(acc1, tmem_base + GT_TM_ALO + kcol, db + 2u * kk, idesc_n, 1u);                    // acc1 += Ul x hh
            }
          }
          if (upto == GT_CL) umma_commit(mma_bar);
        }
        __syncwarp();
        next = upto;
      }
      GT_STAMP(2)
    }
#ifdef SAR_GRU_PROFILE
    if (blockIdx.x == 0 && lane == 0)
      for (int i = 0; i < 4; ++i)
        printf("gru_tc mma  step %d: t0 %lld  wait-h %lld  issue %lld\n", 8 + i, stamps[i * 8], stamps[i * 8 + 1] - stamps[i * 8], stamps[i * 8 + 2] - stamps[i * 8 + 1]);
#endif
  } else {
    // ===================== gate warps =====================
    const int quad = warp & 3, chalf = warp >> 2;
    const int n = t / OPR, o = t % OPR;
    const int u0 = j0 + UPT * o;
    const bool bvalid = (b0 + n) < B;
    float rbz[UPT], rbr[UPT], rbh[UPT];
    ldg_vec<UPT>(rbias + dir * U3 + u0, rbz);
    ldg_vec<UPT>(rbias + dir * U3 + U + u0, rbr);
    ldg_vec<UPT>(rbias + dir * U3 + 2 * U + u0, rbh);
    auto xrow = [&](int step) {
      const int tt = dir ? (S - 1 - step) : step;
      return xp + (((size_t)(b0 + n) * S + tt) * 2 + dir) * U3 + u0;
    };
    float x0z[UPT], x0r[UPT], x0h[UPT], x1z[UPT], x1r[UPT], x1h[UPT], hold[UPT];
#pragma unroll
    for (int e = 0; e < UPT; ++e) x0z[e] = x0r[e] = x0h[e] = x1z[e] = x1r[e] = x1h[e] = hold[e] = 0.f;
    if (bvalid) {
      const float* p = xrow(0); ldg_vec<UPT>(p, x0z); ldg_vec<UPT>(p + U, x0r); ldg_vec<UPT>(p + 2 * U, x0h);
      if (S > 1) { const float* q = xrow(1); ldg_vec<UPT>(q, x1z); ldg_vec<UPT>(q + U, x1r); ldg_vec<UPT>(q + 2 * U, x1h); }
    }
    // my units in the staging block (= atom `rank` of a B buffer): row n (hi) / NB + n (lo), 16-byte chunk swizzled
    // by the row (NB is a multiple of 8: both rows have the same swizzle), UPT halves inside it
    const uint32_t stage_u = smem_u32(stage);
    const int ub = UPT * o * 2;                                  // byte offset of my units in the 64-byte row
    const uint32_t my_off = (uint32_t)(n * GT_ROWB + (((ub >> 4) ^ ((n >> 1) & 3)) << 4) + (ub & 15));
    // threads 0..7 push the block to CTA rank + t of the cluster (atom `rank` there, and that atom's barrier)
    // (both buffers' addresses are mapped one by one: compute-sanitizer's memcheck rejects an offset added to a
    // mapa result -- "not located in remote CTA" -- although the hardware accepts it)
    const uint32_t peer = (uint32_t)((rank + t) & 7);
    const uint32_t rdst_a = gt_mapa(b_u + (uint32_t)(rank * B_ATOM), peer), rdst_b = gt_mapa(b_u + (uint32_t)(rank * B_ATOM + B_BUF), peer);
    const uint32_t rbar_a = gt_mapa(hbar_u + (uint32_t)(rank * 8), peer), rbar_b = gt_mapa(hbar_u + (uint32_t)(rank * 8 + GT_CL * 8), peer);
#ifdef SAR_GRU_PROFILE
    long long stamps[32];
#endif
    for (int step = 0; step < S; ++step) {
      const int cur = step & 1;
      GT_STAMP(0)
      mbar_wait(mma_bar, (uint32_t)step & 1u);
      GT_STAMP(1)
      tc_fence_after();
      if (quad < 3) {                                // rows 96..127 of the tile are padding
        constexpr int CW = NB / 2;                  // columns (utterances) per warp
        const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chalf * CW);
        float* dst = hp_s + ((size_t)quad * NB + chalf * CW) * GT_UPC + lane;
#pragma unroll
        for (int c8 = 0; c8 < CW; c8 += 8) {
          uint32_t r0[8], r1[8];
          tmem_ld8(ta + (uint32_t)c8, r0);
          tmem_ld8(ta + (uint32_t)(NB + c8), r1);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) dst[(c8 + e) * GT_UPC] = fmaf(__uint_as_float(r1[e]), 1.f / 2048.f, __uint_as_float(r0[e]));
        }
      }
      tc_fence_before();
      named_bar_sync(1, GT_GATE_THREADS);
      GT_STAMP(2)
      float hz[UPT], hr[UPT], hc[UPT], hn[UPT];
      ld_vec<UPT>(hp_s + ((size_t)0 * NB + n) * GT_UPC + UPT * o, hz);
      ld_vec<UPT>(hp_s + ((size_t)1 * NB + n) * GT_UPC + UPT * o, hr);
      ld_vec<UPT>(hp_s + ((size_t)2 * NB + n) * GT_UPC + UPT * o, hc);
#pragma unroll
      for (int e = 0; e < UPT; ++e) {
        const float z = gt_sigmoid(x0z[e] + hz[e] + rbz[e]);
        const float r = gt_sigmoid(x0r[e] + hr[e] + rbr[e]);
        const float hh = gt_tanh(x0h[e] + r * (hc[e] + rbh[e]));
        hn[e] = z * hold[e] + (1.f - z) * hh;
        hold[e] = hn[e];
      }
      GT_STAMP(3)
      if (step + 1 < S) {
        // staging block step & 1: its previous push (step - 2) was consumed by every peer's step-1 MMAs, which all
        // of this step's inputs depend on
        const uint32_t sb = stage_u + (uint32_t)(cur * B_ATOM);
#pragma unroll
        for (int e = 0; e < UPT; e += 2) {
          const __half2 a = __floats2half2_rn(hn[e], hn[e + 1]);
          const float2 fa = __half22float2(a);
          const __half2 la = __floats2half2_rn((hn[e] - fa.x) * 2048.f, (hn[e + 1] - fa.y) * 2048.f);
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(sb + my_off + (uint32_t)(2 * e)), "r"(*reinterpret_cast<const uint32_t*>(&a)) : "memory");
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(sb + my_off + (uint32_t)(NB * GT_ROWB + 2 * e)), "r"(*reinterpret_cast<const uint32_t*>(&la)) : "memory");
        }
        fence_proxy_async();                        // the bulk copy reads shared memory through the async proxy
        named_bar_sync(2, GT_GATE_THREADS);
        if (t < GT_CL)
          bulk_push(cur ? rdst_a : rdst_b, sb, (uint32_t)B_ATOM, cur ? rbar_a : rbar_b);
      }
      if (bvalid) {
        const int tt = dir ? (S - 1 - step) : step;
        if (seq) st_vec<UPT>(out + ((size_t)(b0 + n) * S + tt) * (2 * U) + dir * U + u0, hn);
        else if (step == S - 1) st_vec<UPT>(out + (size_t)(b0 + n) * (2 * U) + dir * U + u0, hn);
      }
      GT_STAMP(4)
#pragma unroll
      for (int e = 0; e < UPT; ++e) { x0z[e] = x1z[e]; x0r[e] = x1r[e]; x0h[e] = x1h[e]; }
      if (bvalid && step + 2 < S) { const float* p = xrow(step + 2); ldg_vec<UPT>(p, x1z); ldg_vec<UPT>(p + U, x1r); ldg_vec<UPT>(p + 2 * U, x1h); }
      GT_STAMP(5)
    }
#ifdef SAR_GRU_PROFILE
    if (blockIdx.x == 0 && t == 0)
      for (int i = 0; i < 4; ++i)
        printf("gru_tc gate step %d: t0 %lld  wait-mma %lld  tmem+bar %lld  gates %lld  push %lld  out+prefetch %lld | total %lld\n", 8 + i, stamps[i * 8],
               stamps[i * 8 + 1] - stamps[i * 8], stamps[i * 8 + 2] - stamps[i * 8 + 1], stamps[i * 8 + 3] - stamps[i * 8 + 2],
               stamps[i * 8 + 4] - stamps[i * 8 + 3], stamps[i * 8 + 5] - stamps[i * 8 + 4], stamps[i * 8 + 5] - stamps[i * 8]);
#endif
  }
  tc_fence_before();
  cluster.sync();                                  // nobody leaves while a peer could still address its shared memory
  if (warp == GT_WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GT_TMEM_COLS);
  }
}

static int bigru_tc_launch(const float* xp, const float* rec, const float* rbias, float* out, int B, int S, int seq, int nb_req, cudaStream_t stream) {
  static const int force_nb = getenv("SAR_GRU_NB") ? atoi(getenv("SAR_GRU_NB")) : 0;     // experiments: 16 or 32
  const int nb = force_nb ? force_nb : (nb_req ? nb_req : (2 * ((B + 15) / 16) <= GT_MAX_CLUSTERS ? 16 : 32));
  auto launch = [&](auto kern, size_t smem, int NBv) -> int {
    { const int arc = allow_max_smem(kern, "sar_bigru_fwd"); if (arc) return arc; }
    const int groups = (B + NBv - 1) / NBv;
    launch_k(kern, dim3(GT_CL * groups * 2), dim3(GT_THREADS), smem, stream, xp, rec, rbias, out, B, S, seq);
    return 0;
  };
  const int rc = nb == 16 ? launch(bigru_tc_kernel<16>, GtCfg<16>::SMEM, 16) : launch(bigru_tc_kernel<32>, GtCfg<32>::SMEM, 32);
  if (rc) return rc;
  return check_launch("sar_bigru_fwd");
}

}  // namespace sar

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
  SAR_REQUIRE(u == GT_U, SAR_ERR_UNSUPPORTED, "sar_bigru_fwd: hidden size %d unsupported (this build: %d)", u, GT_U);
  SAR_REQUIRE(aligned16(xp) && aligned16(rbias) && aligned16(out), SAR_ERR_BAD_ARG, "sar_bigru_fwd: pointers must be 16-byte aligned");
  return bigru_tc_launch(xp, rec, rbias, out, B, S, seq, nb, (cudaStream_t)stream);
}
