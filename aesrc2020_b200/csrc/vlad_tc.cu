// vlad_tc.cu -- NetVLAD / GhostVLAD with both contractions on the 5th-gen tensor cores.
//
// Replaces the 1x1 assignment Conv2D (model.py:89-95 / 99-105) + VladPooling.call (VLAD.py:26-49), like vlad.cu, but
//     score = X @ Wa + ba            (S x 256 x (K+G))          VLAD.py:33 (the conv) -> softmax, VLAD.py:34-35
//     V^T   = X^T @ A[:, k-slice]    (256 x S x K)              VLAD.py:38-45 (sum_s A[s,k] * X[s]) - (sum_s A[s,k]) * c[k]
// are tcgen05.mma with fp32 accumulators in tensor memory.  fp32 accuracy comes from the same fp16 hi/lo operand
// split as the convolutions (conv_tc.cu): [acc0 | acc1] (+)= A_hi x [B_hi ; B_lo], acc1 += A_lo x B_hi, result
// acc0 + 2^-11 acc1.
//
// Data flow of one work item = (utterance b, cluster slice h):
//   * X (the LayerNorm'ed AR_DS output) arrives as fp16 hi/lo planes [2][B*S][256] written by layernorm_kernel; ONE TMA
//     box per 64-channel chunk lands it in shared memory, 128B-swizzled, rows = descriptors s.  The SAME tile is the
//     K-major A operand of the score GEMM (M = s, K = d) and the MN-major A operand of the residual GEMM (M = d, K = s):
//     no transpose anywhere.
//   * Wa^T hi/lo ([2][KGP][256], packed on the host) is loaded once per CTA and stays resident (persistent CTAs).
//   * scores: two threads per descriptor row (column halves); online softmax straight out of tensor memory; the probabilities of this item's
//     cluster slice go to shared memory as fp16 hi/lo rows [s][P_hi | P_lo] -- the MN-major B operand of the second
//     GEMM (N = cluster) -- and their column sums (sum_s A[s,k]) are reduced with warp shuffles.
//   * residual: thread = feature column d (two 128-column M tiles); V[k][d] = acc - asum[k] * c[k][d], squared norms
//     per cluster reduced across the CTA, rows written normalised: every store instruction covers 128 B (fp32) /
//     64 B (fp16 planes for the tensor-core AR_EMBEDDING GEMM) of one output row.
// HBM traffic is the algorithmic minimum (X read once per slice from L2/HBM, K*D written once).
//
// Roles: warps 0-7 compute (TMEM lane quadrant = warp & 3), warp 8 = TMA producer + MMA issuer (one elected lane).
// Limits: D == 256, S <= 128, K + G <= 128 and the tiles must fit shared memory; sar_vlad_tc_supported() tells, the
// CUDA-core kernel (vlad.cu) covers everything else (and the VladPooling([feat, score]) surface with given scores).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace sar {

constexpr int VT_THREADS = 288;              // 8 compute warps + the TMA / MMA warp
constexpr int VT_WARP_MMA = 8;
constexpr int VT_D = 256;
constexpr int VT_MISC_BYTES = 6144;
constexpr float VT_LO_INV = 1.f / 2048.f;
constexpr uint32_t VT_ACC2 = 256;                 // TMEM column of the residual accumulators (scores use [0, 2*KGP))

struct VtParams {
  int B, S, S_pad, K, G, KG, KGP, KL, nsplit, items, nxbuf;
  int xcb, wcb, acb, a_chunks;                    // bytes of one 64-channel chunk of X / Wa^T / the probability tile
  const float* ba; const float* centers;
  float* out; __half* out_planes;
};

// shared-memory matrix descriptor, 128B swizzle: K-major operands ignore LBO; MN-major: LBO = stride between 64-element
// groups along M/N, SBO = stride between 8-row groups along K (cute: ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)))
__device__ __forceinline__ uint64_t vt_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

// v[i] of every lane -> lane l returns sum over the 32 lanes of v[l] (31 shuffles instead of 32 x 5)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(VT_THREADS, 1)
vlad_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, const VtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;                                           // [nxbuf][plane 2][chunk 4][xcb]
  uint8_t* ws = xs + (size_t)p.nxbuf * 8 * p.xcb;               // [chunk 4][plane 2][wcb]: [B_hi ; B_lo] adjacent
  uint8_t* as_ = ws + (size_t)8 * p.wcb;                        // [a_chunks][acb]: rows s, columns [P_hi | P_lo]
  uint8_t* misc = as_ + (size_t)p.a_chunks * p.acb;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(misc);
  uint64_t* xfull = wfull + 1;                                  // [2]
  uint64_t* xempty = xfull + 2;                                 // [2]
  uint64_t* sfull = xempty + 2;                                 // scores complete
  uint64_t* afull = sfull + 1;                                  // probability tile written (256 arrivals)
  uint64_t* vfull = afull + 1;                                  // residual accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  float* s_ba = reinterpret_cast<float*>(misc + 128);           // [128]
  float* s_part = s_ba + 128;                                   // [4 warps][64] column sums of the probabilities
  float* s_asum = s_part + 256;                                 // [64]
  float* s_ss = s_asum + 64;                                    // [8 warps][32] squared-norm partials
  float* s_inv = s_ss + 256;                                    // [32]
  float* s_mx = s_inv + 32;                                     // [2 halves][128 rows] softmax partial max
  float* s_sm = s_mx + 256;                                     // [2 halves][128 rows] softmax partial sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int KGP = p.KGP, KL = p.KL;

  if (warp == VT_WARP_MMA) {
    if (lane == 0) {
      mbar_init(wfull, 1);
      mbar_init(&xfull[0], 1); mbar_init(&xfull[1], 1);
      mbar_init(&xempty[0], 1); mbar_init(&xempty[1], 1);
      mbar_init(sfull, 1); mbar_init(afull, 256); mbar_init(vfull, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      prefetch_tmap(&mapX); prefetch_tmap(&mapW);
    }
  } else if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
  }
  if (tid < 128) s_ba[tid] = (tid < p.KG) ? __ldg(p.ba + tid) : 0.f;      // constants: before the PDL wait
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == VT_WARP_MMA) {
    // ===================== TMA producer + MMA issuer =====================
    if (elect_one()) {                                          // Wa^T: a constant, in flight before the PDL wait
      mbar_expect_tx(wfull, (uint32_t)(8 * p.wcb));
      for (int c = 0; c < 4; ++c)
        for (int pl = 0; pl < 2; ++pl)
          tma_load_3d(&mapW, smem_u32(ws + (size_t)(2 * c + pl) * p.wcb), wfull, c * 64, 0, pl);
    }
    __syncwarp();
    pdl_wait();
    pdl_trigger();
    auto issue_x = [&](int it_idx, int item) {
      const int buf = it_idx % p.nxbuf;
      if (it_idx >= p.nxbuf) mbar_wait(&xempty[buf], (uint32_t)((it_idx / p.nxbuf) - 1) & 1u);
      if (elect_one()) {
        const int b = item / p.nsplit;
        mbar_expect_tx(&xfull[buf], (uint32_t)(8 * p.xcb));
        for (int pl = 0; pl < 2; ++pl)
          for (int c = 0; c < 4; ++c)
            tma_load_3d(&mapX, smem_u32(xs + ((size_t)buf * 8 + pl * 4 + c) * p.xcb), &xfull[buf], c * 64, b * p.S, pl);
      }
      __syncwarp();
    };
    const uint32_t idesc1_2n = (1u << 4) | ((uint32_t)((2 * KGP) >> 3) << 17) | (8u << 24);
    const uint32_t idesc1_n = (1u << 4) | ((uint32_t)(KGP >> 3) << 17) | (8u << 24);
    const uint32_t mn = (1u << 15) | (1u << 16);                // both operands MN-major
    const uint32_t idesc2_2n = (1u << 4) | mn | ((uint32_t)((2 * KL) >> 3) << 17) | (8u << 24);
    const uint32_t idesc2_n = (1u << 4) | mn | ((uint32_t)(KL >> 3) << 17) | (8u << 24);
    const int nk2 = p.S_pad >> 4;
    int it = 0;
    if ((int)blockIdx.x < p.items) issue_x(0, blockIdx.x);
    mbar_wait(wfull, 0);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int next = item + gridDim.x;
      if (p.nxbuf == 2 && next < p.items) issue_x(it + 1, next);
      const int buf = it % p.nxbuf;
      mbar_wait(&xfull[buf], (uint32_t)(it / p.nxbuf) & 1u);
      tc_fence_after();
      const uint32_t xb = smem_u32(xs + (size_t)buf * 8 * p.xcb);
      if (elect_one()) {                                        // ---- scores: M = s (128 rows), N = clusters, K = 256
        const uint32_t wb = smem_u32(ws);
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t dah = vt_desc(xb + (uint32_t)(c * p.xcb + kk * 32), 16, 1024);
            const uint64_t dal = vt_desc(xb + (uint32_t)((4 + c) * p.xcb + kk * 32), 16, 1024);
            const uint64_t dbh = vt_desc(wb + (uint32_t)(2 * c * p.wcb + kk * 32), 16, 1024);
            umma_f16(tmem_base, dah, dbh, idesc1_2n, (c | kk) ? 1u : 0u);          // [acc0 | acc1] (+)= Xh x [Wh ; Wl]
            umma_f16(tmem_base + (uint32_t)KGP, dal, dbh, idesc1_n, 1u);            // acc1 += Xl x Wh
          }
        umma_commit(sfull);
      }
      __syncwarp();
      mbar_wait(afull, (uint32_t)it & 1u);                      // probability tile of this item is in shared memory
      tc_fence_after();
      if (elect_one()) {                                        // ---- residual: M = d (2 tiles of 128), N = slice, K = s
        const uint32_t ab = smem_u32(as_);
        for (int t = 0; t < 2; ++t) {
          const uint32_t acc = tmem_base + VT_ACC2 + (uint32_t)(t * 2 * KL);
          for (int kk = 0; kk < nk2; ++kk) {
            const uint64_t dah = vt_desc(xb + (uint32_t)(2 * t * p.xcb + kk * 2048), (uint32_t)p.xcb, 1024);
            const uint64_t dal = vt_desc(xb + (uint32_t)((4 + 2 * t) * p.xcb + kk * 2048), (uint32_t)p.xcb, 1024);
            const uint64_t db = vt_desc(ab + (uint32_t)(kk * 2048), (uint32_t)p.acb, 1024);
            umma_f16(acc, dah, db, idesc2_2n, kk ? 1u : 0u);                         // [acc0 | acc1] (+)= Xh^T x [Ph | Pl]
            umma_f16(acc + (uint32_t)KL, dal, db, idesc2_n, 1u);                     // acc1 += Xl^T x Ph
          }
        }
        umma_commit(vfull);
        umma_commit(&xempty[buf]);                              // X tile free once both GEMMs have retired
      }
      __syncwarp();
      if (p.nxbuf == 1 && next < p.items) issue_x(it + 1, next);
    }
  } else {
    // ===================== compute warps =====================
    // 8 warps: quad = warp & 3 is the TMEM lane quadrant (rows quad*32 .. +31), hf = warp >> 2 splits the work of a
    // quadrant in two: the score columns (softmax partials are merged through shared memory), the 32-cluster blocks
    // of the probability tile, and the two 128-column M tiles of the residual epilogue.  Two warps per scheduler, and
    // half the per-thread work of a 4-warp layout: the phases are chains of dependent ALU ops, i.e. latency bound.
    pdl_wait();                                                 // centers are constants, the outputs are not
    const int quad = warp & 3, hf = warp >> 2;
    const int row = quad * 32 + lane;                           // descriptor row s (scores) / feature column d (residual)
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool row_valid = row < p.S;
    const uint32_t a_u = smem_u32(as_);
    constexpr float LOG2E = 1.4426950408889634f;
    int it = 0;
#ifdef SAR_VLAD_PROFILE
    long long st[8];
#define VT_STAMP(j) if (blockIdx.x == 0 && tid == 0 && it == 1) st[j] = clock64();
#else
#define VT_STAMP(j)
#endif
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int b = item / p.nsplit, h = item - b * p.nsplit;
      VT_STAMP(0)
      const int k_lo = h * KL;                                  // this item's clusters: [k_lo, min(K, k_lo + KL))
      mbar_wait(sfull, (uint32_t)it & 1u);
      tc_fence_after();
      VT_STAMP(1)
      // ---- softmax over the K+G scores of my row (VLAD.py:34-35): my half of the 16-column groups, online
      float mx = -INFINITY, sum = 0.f;
      for (int g = 16 * hf; g < KGP; g += 32) {
        uint32_t r0[16], r1[16];
        tmem_ld16(lane_base + (uint32_t)g, r0);
        tmem_ld16(lane_base + (uint32_t)(KGP + g), r1);
        tmem_ld_wait();
        float v[16];
        float gm = -INFINITY;
        const int nv = p.KG - g;                                // valid columns of this group (>= 16: all)
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          v[e] = (fmaf(__uint_as_float(r1[e]), VT_LO_INV, __uint_as_float(r0[e])) + s_ba[g + e]) * LOG2E;
          if (e >= nv) v[e] = -INFINITY;
          gm = fmaxf(gm, v[e]);
        }
        const float nm = fmaxf(mx, gm);
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) part += exp2f(v[e] - nm);  // exp2f(-inf) = 0 for the padded columns
        sum = (mx == -INFINITY ? 0.f : sum * exp2f(mx - nm)) + part;
        mx = nm;
      }
      s_mx[hf * 128 + row] = mx;
      s_sm[hf * 128 + row] = sum;
      named_bar_sync(1, 256);
      {
        const float m0 = s_mx[row], m1 = s_mx[128 + row], q0 = s_sm[row], q1 = s_sm[128 + row];
        mx = fmaxf(m0, m1);
        sum = (m0 == -INFINITY ? 0.f : q0 * exp2f(m0 - mx)) + (m1 == -INFINITY ? 0.f : q1 * exp2f(m1 - mx));
      }
      const float inv_sum = 1.0f / sum;
      VT_STAMP(2)
      // ---- probabilities of my cluster slice -> [P_hi | P_lo] rows (MN-major B operand), column sums by shuffles:
      //      32-cluster block jb = 32 * hf (+ 64 ...) is mine
      for (int jb = 32 * hf; jb < KL; jb += 64) {
        float pr[32];
#pragma unroll
        for (int g = 0; g < 32; g += 16) {
          if (jb + g < KL) {
            uint32_t r0[16], r1[16];
            tmem_ld16(lane_base + (uint32_t)(k_lo + jb + g), r0);
            tmem_ld16(lane_base + (uint32_t)(KGP + k_lo + jb + g), r1);
            tmem_ld_wait();
            const int nv = row_valid ? p.K - (k_lo + jb + g) : 0;          // clusters of this group that exist
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int k = k_lo + jb + g + e;
              const float v = (fmaf(__uint_as_float(r1[e]), VT_LO_INV, __uint_as_float(r0[e])) + s_ba[k & 127]) * LOG2E;
              pr[g + e] = (e < nv) ? exp2f(v - mx) * inv_sum : 0.f;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) pr[g + e] = 0.f;
          }
        }
        if (row < p.S_pad) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {                         // 8 probabilities = one 16-byte unit of hi and of lo
            if (jb + 8 * u < KL) {
              uint32_t hh[4], ll[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a0 = pr[8 * u + 2 * e], a1 = pr[8 * u + 2 * e + 1];
                const __half2 h2 = __floats2half2_rn(a0, a1);
                const float2 hf2 = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn((a0 - hf2.x) * 2048.f, (a1 - hf2.y) * 2048.f);
                hh[e] = *reinterpret_cast<const uint32_t*>(&h2);
                ll[e] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              const int jh = jb + 8 * u, jl = KL + jb + 8 * u;  // column of the unit in [P_hi | P_lo]
              const uint32_t ah = a_u + (uint32_t)((jh >> 6) * p.acb + row * 128 + ((((jh & 63) >> 3) ^ (row & 7)) << 4));
              const uint32_t al = a_u + (uint32_t)((jl >> 6) * p.acb + row * 128 + ((((jl & 63) >> 3) ^ (row & 7)) << 4));
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ah), "r"(hh[0]), "r"(hh[1]), "r"(hh[2]), "r"(hh[3]) : "memory");
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(al), "r"(ll[0]), "r"(ll[1]), "r"(ll[2]), "r"(ll[3]) : "memory");
            }
          }
        }
        const float cs = warp_transpose_sum(pr, lane);          // sum over this warp's 32 rows of column jb + lane
        s_part[quad * 64 + jb + lane] = cs;
      }
      VT_STAMP(3)
      fence_proxy_async();                                      // generic-proxy writes -> visible to tcgen05.mma
      tc_fence_before();
      mbar_arrive(afull);
      named_bar_sync(1, 256);
      if (tid < KL) s_asum[tid] = (s_part[tid] + s_part[64 + tid]) + (s_part[128 + tid] + s_part[192 + tid]);
      named_bar_sync(1, 256);
      // ---- residual epilogue: thread = feature column d = hf * 128 + row of M tile hf
      const int d = hf * 128 + row;
      for (int kb = 0; kb < KL; kb += 32) {
        float v[32];
        float ssq[32];
        const int kbase = k_lo + kb;
        int nk = KL - kb;                                       // clusters of this block that exist (uniform)
        if (p.K - kbase < nk) nk = p.K - kbase;
        if (nk > 32) nk = 32;
        // the cluster centres of this block: 32 independent loads per thread, in flight BEFORE the accumulators are
        // waited for (L2 hits: 72 KB shared by every CTA, more than the L1 left beside 190 KB of shared memory)
        const float* cptr = p.centers + (size_t)kbase * VT_D + d;
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = (e < nk) ? __ldg(cptr + e * VT_D) : 0.f;
        if (kb == 0) {
          VT_STAMP(4)
          mbar_wait(vfull, (uint32_t)it & 1u);
          tc_fence_after();
          VT_STAMP(5)
        }
#pragma unroll
        for (int g = 0; g < 32; g += 16) {
          if (kb + g < KL) {
            uint32_t r0[16], r1[16];
            tmem_ld16(lane_base + VT_ACC2 + (uint32_t)(hf * 2 * KL + kb + g), r0);
            tmem_ld16(lane_base + VT_ACC2 + (uint32_t)(hf * 2 * KL + KL + kb + g), r1);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float acc = fmaf(__uint_as_float(r1[e]), VT_LO_INV, __uint_as_float(r0[e]));
              const float x = (g + e < nk) ? fmaf(-s_asum[kb + g + e], v[g + e], acc) : 0.f;       // VLAD.py:38-45
              v[g + e] = x;
              ssq[g + e] = x * x;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) { v[g + e] = 0.f; ssq[g + e] = 0.f; }
          }
        }
        const float ws_ = warp_transpose_sum(ssq, lane);
        s_ss[warp * 32 + lane] = ws_;
        named_bar_sync(1, 256);
        if (tid < 32) {
          float tot = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) tot += s_ss[w8 * 32 + tid];
          s_inv[tid] = 1.0f / sqrtf(fmaxf(tot, 1e-12f));                                           // VLAD.py:47-48
        }
        named_bar_sync(1, 256);
        const size_t o0 = ((size_t)b * p.K + kbase) * VT_D + d;
        if (p.out) {
          float* po = p.out + o0;
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e < nk) po[e * VT_D] = v[e] * s_inv[e];
        }
        if (p.out_planes) {
          __half* ph = p.out_planes + o0;
          __half* pl = ph + (size_t)p.B * p.K * VT_D;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (e < nk) {
              const float o = v[e] * s_inv[e];
              const __half hi = __float2half_rn(o);
              ph[e * VT_D] = hi;
              pl[e * VT_D] = __float2half_rn((o - __half2float(hi)) * 2048.f);
            }
          }
        }
        named_bar_sync(1, 256);                                 // s_ss / s_inv are reused by the next cluster block
      }
      tc_fence_before();                                        // my TMEM reads are ordered before the next item's MMAs
      VT_STAMP(6)
#ifdef SAR_VLAD_PROFILE
      if (blockIdx.x == 0 && tid == 0 && it == 1)
        printf("vlad_tc item 1: wait scores %lld | softmax %lld | probs+tile %lld | asum+centers issue %lld | wait GEMM2 %lld | epilogue %lld | total %lld\n",
               st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[6] - st[5], st[6] - st[0]);
#endif
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// tiling decision shared by sar_vlad_tc_supported and the launch
static bool vt_plan(int B, int S, int D, int K, int G, int sms, VtParams& p, size_t& smem) {
  if (D != VT_D || S < 1 || S > 128 || K < 1 || G < 0 || K + G > 128 || B < 1) return false;
  p.B = B; p.S = S; p.K = K; p.G = G;
  p.S_pad = (S + 15) & ~15;
  p.KG = K + G;
  p.KGP = (p.KG + 15) & ~15;
  // cluster slices: enough items to fill the SMs at small batches; a slice is a multiple of 16 clusters, at most 64
  int nsplit = (sms + B - 1) / B;
  if (nsplit < 1) nsplit = 1;
  int KL = (((K + nsplit - 1) / nsplit) + 15) & ~15;
  if (KL > 64) KL = 64;
  if (KL < 16) KL = 16;
  p.KL = KL;
  p.nsplit = (K + KL - 1) / KL;
  p.items = B * p.nsplit;
  p.xcb = p.S_pad * 128;
  p.wcb = p.KGP * 128;
  p.acb = p.S_pad * 128;
  p.a_chunks = (2 * KL + 63) / 64;
  const size_t fixed = 1024 + (size_t)8 * p.wcb + (size_t)p.a_chunks * p.acb + VT_MISC_BYTES;
  // the score GEMM reads 128 rows of every X chunk: rows beyond S_pad run into what follows the tile, which must be
  // inside the allocation (Wa^T, the probability tile and the misc block follow the X buffers)
  const size_t overrun = (size_t)(128 - p.S_pad) * 128;
  for (p.nxbuf = 2; p.nxbuf >= 1; --p.nxbuf) {
    smem = fixed + (size_t)p.nxbuf * 8 * p.xcb;
    if (smem <= SAR_MAX_DYN_SMEM && (size_t)8 * p.wcb + (size_t)p.a_chunks * p.acb + VT_MISC_BYTES >= overrun) return true;
  }
  return false;
}

}  // namespace sar

extern "C" int sar_vlad_tc_supported(int B, int S, int D, int K, int G) {
  using namespace sar;
  VtParams p{};
  size_t smem = 0;
  return vt_plan(B, S, D, K, G, 148, p, smem) ? 1 : 0;
}

extern "C" int sar_vlad_tc_fwd(const void* x_planes, long long x_rows, const void* wa_packed, const float* b_assign,
                               const float* centers, float* out, void* out_planes, int B, int S, int D, int K, int G,
                               void* stream) {
  using namespace sar;
  SAR_REQUIRE(x_planes && wa_packed && b_assign && centers && (out || out_planes), SAR_ERR_BAD_ARG, "sar_vlad_tc_fwd: null pointer");
  SAR_REQUIRE(x_rows >= (long long)B * S, SAR_ERR_BAD_ARG, "sar_vlad_tc_fwd: x_planes has %lld rows, need B*S = %lld", x_rows, (long long)B * S);
  SAR_REQUIRE(aligned16(x_planes) && aligned16(wa_packed) && aligned16(centers) && (!out || aligned16(out)) &&
                  (!out_planes || aligned16(out_planes)), SAR_ERR_ALIGN, "sar_vlad_tc_fwd: unaligned pointer");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  VtParams p{};
  size_t smem = 0;
  SAR_REQUIRE(vt_plan(B, S, D, K, G, sms, p, smem), SAR_ERR_UNSUPPORTED,
              "sar_vlad_tc_fwd: unsupported shape (need D == 256, S <= 128, K+G <= 128 and tiles within shared memory; "
              "got S=%d D=%d K=%d G=%d) -- use sar_vlad_fwd", S, D, K, G);
  p.ba = b_assign; p.centers = centers; p.out = out; p.out_planes = reinterpret_cast<__half*>(out_planes);
  CUtensorMap mapX, mapW;
  int rc;
  if ((rc = tc_make_map(&mapX, x_planes, x_rows, VT_D, 2, 64, p.S_pad))) return rc;
  if ((rc = tc_make_map(&mapW, wa_packed, p.KGP, VT_D, 2, 64, p.KGP))) return rc;
  { const int arc = allow_max_smem(vlad_tc_kernel, "sar_vlad_tc_fwd"); if (arc) return arc; }
  const int grid = p.items < sms ? p.items : sms;
  launch_k(vlad_tc_kernel, dim3(grid), dim3(VT_THREADS), smem, (cudaStream_t)stream, mapX, mapW, p);
  return check_launch("sar_vlad_tc_fwd");
}
