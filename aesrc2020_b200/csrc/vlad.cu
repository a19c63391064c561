// vlad.cu -- NetVLAD / GhostVLAD in one pass over the utterance's descriptors.
//
// Replaces the 1x1 assignment Conv2D (model.py:89-95 / 99-105) + VladPooling.call
// (VLAD.py:26-49).  The reference materialises feat_res and weighted_res, two
// (B,1,S,K+G,D) fp32 tensors (VLAD.py:40-41; 14.2 MB per utterance at K+G=72, D=256, S=48).
// Here one CTA per utterance stages X (S x D) in shared memory ONCE and computes
//     score = X @ Wa + ba ;  A = softmax_k(score)                      VLAD.py:33-35
//     V[k] = sum_s A[s,k] * X[s] - (sum_s A[s,k]) * c[k]     k < K      VLAD.py:38-45
//     out[k] = V[k] / sqrt(max(|V[k]|^2, 1e-12))                        VLAD.py:47-48
// so HBM traffic is the algorithmic minimum: X read once, K*D written once (ghost clusters only
// take part in the softmax and are never accumulated).
//
// 512 threads (16 warps: the kernel is latency bound at one CTA per SM, see profiles/r1_summary_v9.md), register tiled:
//   scores   : thread = 4 descriptors x 4 clusters x HALF of d (warps 0-7: d < D/2 into A, warps 8-15: the rest
//              into A2; the softmax pass adds the halves), marching over d in float4 steps (8 LDS.128 per
//              64 FMA), Wa staged in shared memory when it fits;
//   residual : warp = (8 clusters, half of the columns), lane = 4 feature columns -> 32 accumulators per
//              thread, 3 LDS.128 per 32 FMA; the per-cluster L2 norm is a warp reduction + one exchange
//              between the two column halves, and the normalised rows go straight from registers to HBM.
#include "common.cuh"

namespace sar {

constexpr int VLAD_THREADS = 512;

__global__ void __launch_bounds__(VLAD_THREADS) vlad_kernel(const float* __restrict__ feat, const float* __restrict__ wa,
                                                             const float* __restrict__ ba, const float* __restrict__ score,
                                                             const float* __restrict__ centers, float* __restrict__ out,
                                                             __half* __restrict__ out_planes, int S, int D, int K, int G, int wa_smem) {
  extern __shared__ __align__(16) float smem[];
  const int KG = K + G;
  const int KGP = (KG + 7) & ~7;                  // padded cluster count (A row, zero-filled)
  const int XS = D + 4;                           // padded X row: +4 banks per descriptor
  const int SP = (S + 3) & ~3;                    // descriptors padded to the 4-row score tile
  float* X = smem;                                // [SP][XS]
  float* A = X + (size_t)SP * XS;                 // [SP][KGP]
  float* A2 = A + (size_t)SP * KGP;               // [SP][KGP] scores of the upper half of d
  float* nrm = A2 + (size_t)SP * KGP;             // [8 warps pairs][8 clusters][2 halves] squared-norm partials
  float* Ws = nrm + 128;                          // [D][KGP] (only when wa_smem)
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;

  // ---- constants first (overlaps the previous kernel's tail under programmatic dependent launch)
  for (int i = t; i < 2 * SP * KGP; i += VLAD_THREADS) A[i] = 0.f;       // A and A2
  if (!score && wa_smem) {
    for (int i = t; i < D * KGP; i += VLAD_THREADS) {
      const int d = i / KGP, k = i - d * KGP;
      Ws[i] = (k < KG) ? __ldg(wa + (size_t)d * KG + k) : 0.f;
    }
  }
  pdl_wait();
  pdl_trigger();
  // ---- stage X (coalesced 16B loads), zero the padded rows/columns
  {
    const float4* src = reinterpret_cast<const float4*>(feat + (size_t)b * S * D);
    const int d4 = D >> 2;
    for (int i = t; i < SP * d4; i += VLAD_THREADS) {
      const int s = i / d4, c = i - s * d4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < S) v = __ldg(src + (size_t)s * d4 + c);
      *reinterpret_cast<float4*>(X + (size_t)s * XS + 4 * c) = v;
    }
  }
  __syncthreads();

  // ---- scores
  if (score) {                                     // VladPooling called with external scores
    for (int i = t; i < S * KG; i += VLAD_THREADS) {
      const int s = i / KG, k = i - s * KG;
      A[s * KGP + k] = __ldg(score + ((size_t)b * S + s) * KG + k);
    }
  } else {                                         // fused 1x1 assignment conv, 4x4 register tiles
    const int kt = KGP >> 2, tiles = (SP >> 2) * kt;
    const int dh = t >> 8;                         // d half (warp-uniform)
    float* Ad = dh ? A2 : A;
    for (int tile = t & 255; tile < tiles; tile += 256) {
      const int kg4 = tile % kt, sg4 = tile / kt;
      const int k0 = 4 * kg4, s0 = 4 * sg4;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int d = dh * (D >> 1); d < (dh + 1) * (D >> 1); d += 4) {
        float4 xv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(X + (size_t)(s0 + i) * XS + d);
        if (wa_smem) {
#pragma unroll
          for (int e = 0; e < 4; ++e) wv[e] = *reinterpret_cast<const float4*>(Ws + (size_t)(d + e) * KGP + k0);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float* wr = wa + (size_t)(d + e) * KG + k0;
            wv[e] = make_float4(k0 < KG ? __ldg(wr) : 0.f, k0 + 1 < KG ? __ldg(wr + 1) : 0.f,
                                k0 + 2 < KG ? __ldg(wr + 2) : 0.f, k0 + 3 < KG ? __ldg(wr + 3) : 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xe[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[i][0] = fmaf(xe[e], wv[e].x, acc[i][0]);
            acc[i][1] = fmaf(xe[e], wv[e].y, acc[i][1]);
            acc[i][2] = fmaf(xe[e], wv[e].z, acc[i][2]);
            acc[i][3] = fmaf(xe[e], wv[e].w, acc[i][3]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (s0 + i < S && k0 + j < KG) Ad[(s0 + i) * KGP + k0 + j] = acc[i][j] + (dh ? 0.f : __ldg(ba + k0 + j));
    }
  }
  __syncthreads();

  // ---- softmax over clusters, one warp per descriptor row (padded columns stay 0)
  for (int s = warp; s < S; s += VLAD_THREADS / 32) {
    float m = -INFINITY;
    for (int k = lane; k < KG; k += 32) { A[s * KGP + k] += A2[s * KGP + k]; m = fmaxf(m, A[s * KGP + k]); }
    m = warp_max(m);
    float sum = 0.f;
    for (int k = lane; k < KG; k += 32) {
      float e = expf(A[s * KGP + k] - m);
      A[s * KGP + k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int k = lane; k < KG; k += 32) A[s * KGP + k] = A[s * KGP + k] / sum;
  }
  __syncthreads();

  // ---- residual accumulation: warp <- (8 clusters, 128 columns), lane <- 4 columns
  const int kgroups = (K + 7) >> 3;
  const int half = warp & 1, d0 = half * 128 + 4 * lane;
  for (int kg0 = 0; kg0 < kgroups; kg0 += VLAD_THREADS / 64) {          // uniform trip count: the loop holds a __syncthreads
    const int kgp = kg0 + (warp >> 1);
    const int k0 = 8 * kgp;
    const bool live = kgp < kgroups;                                       // warp-uniform
    float acc[8][4];
    float asum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asum[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    float ss[8];
    if (live) {
#pragma unroll 4
      for (int s = 0; s < S; ++s) {
        const float4 a0 = *reinterpret_cast<const float4*>(A + s * KGP + k0);
        const float4 a1 = *reinterpret_cast<const float4*>(A + s * KGP + k0 + 4);
        const float4 x0 = *reinterpret_cast<const float4*>(X + (size_t)s * XS + d0);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float xv[4] = {x0.x, x0.y, x0.z, x0.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          asum[i] += av[i];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + i;
        ss[i] = 0.f;
        if (k >= K) continue;                               // warp-uniform
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(centers + (size_t)k * D + d0));
        const float cv[4] = {c0.x, c0.y, c0.z, c0.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] -= asum[i] * cv[j];
          ss[i] = fmaf(acc[i][j], acc[i][j], ss[i]);
        }
        ss[i] = warp_sum(ss[i]);
      }
      if (lane < 8) {
        float v = ss[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) v = lane == i ? ss[i] : v;
        nrm[((warp >> 1) * 8 + lane) * 2 + half] = v;
      }
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + i;
        if (k >= K) continue;
        const float* np = nrm + ((warp >> 1) * 8 + i) * 2;
        const float inv = 1.0f / sqrtf(fmaxf(np[0] + np[1], 1e-12f));      // fixed order: both halves get the same bits
        const float o[4] = {acc[i][0] * inv, acc[i][1] * inv, acc[i][2] * inv, acc[i][3] * inv};
        if (out) *reinterpret_cast<float4*>(out + ((size_t)b * K + k) * D + d0) = make_float4(o[0], o[1], o[2], o[3]);
        if (out_planes) {                 // fp16 hi/lo planes [2][B][K*D] for the tensor-core embedding GEMM
          __half* ph = out_planes + ((size_t)b * K + k) * D;
          __half* pl = ph + (size_t)gridDim.x * K * D;
          const __half2 h01 = __floats2half2_rn(o[0], o[1]), h23 = __floats2half2_rn(o[2], o[3]);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn((o[0] - f01.x) * 2048.f, (o[1] - f01.y) * 2048.f);
          const __half2 l23 = __floats2half2_rn((o[2] - f23.x) * 2048.f, (o[3] - f23.y) * 2048.f);
          *reinterpret_cast<uint2*>(ph + d0) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          *reinterpret_cast<uint2*>(pl + d0) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        }
      }
    }
    __syncthreads();                                         // nrm is reused by the next cluster groups
  }
}

static size_t vlad_smem_floats(int S, int D, int KG, bool with_w) {
  const size_t KGP = (KG + 7) & ~7, XS = D + 4, SP = (S + 3) & ~3;
  return SP * XS + 2 * SP * KGP + 128 + (with_w ? (size_t)D * KGP : 0);
}

}  // namespace sar

extern "C" int sar_vlad_fwd(const float* feat, const float* w_assign, const float* b_assign, const float* score,
                            const float* centers, float* out, int B, int S, int D, int K, int G, void* stream) {
  SAR_REQUIRE(out, SAR_ERR_BAD_ARG, "sar_vlad_fwd: null pointer");
  return sar_vlad_planes_fwd(feat, w_assign, b_assign, score, centers, out, nullptr, B, S, D, K, G, stream);
}

extern "C" int sar_vlad_planes_fwd(const float* feat, const float* w_assign, const float* b_assign, const float* score,
                                   const float* centers, float* out, void* out_planes, int B, int S, int D, int K, int G,
                                   void* stream) {
  using namespace sar;
  SAR_REQUIRE(feat && centers && (out || out_planes), SAR_ERR_BAD_ARG, "sar_vlad_fwd: null pointer");
  SAR_REQUIRE((score != nullptr) != (w_assign != nullptr && b_assign != nullptr), SAR_ERR_BAD_ARG,
              "sar_vlad_fwd: pass either (w_assign, b_assign) or score");
  SAR_REQUIRE(B > 0 && S > 0 && D > 0 && K > 0 && G >= 0, SAR_ERR_BAD_ARG, "sar_vlad_fwd: bad dimension");
  SAR_REQUIRE(D == 256 && K + G <= 128, SAR_ERR_UNSUPPORTED,
              "sar_vlad_fwd: this build supports D == 256 (hidden_dim) and K+G <= 128 (got D=%d K+G=%d)", D, K + G);
  SAR_REQUIRE(aligned16(feat) && (!out || aligned16(out)) && (!out_planes || aligned16(out_planes)) && aligned16(centers), SAR_ERR_ALIGN,
              "sar_vlad_fwd: unaligned pointer");
  const size_t limit = 227 * 1024;
  int wa_smem = (!score && vlad_smem_floats(S, D, K + G, true) * sizeof(float) <= limit) ? 1 : 0;
  size_t smem = vlad_smem_floats(S, D, K + G, wa_smem != 0) * sizeof(float);
  SAR_REQUIRE(smem <= limit, SAR_ERR_UNSUPPORTED, "sar_vlad_fwd: S*D too large for shared memory (%zu B)", smem);
  { const int arc = allow_max_smem(vlad_kernel, "sar_vlad_fwd"); if (arc) return arc; }
  launch_k(vlad_kernel, dim3(B), dim3(VLAD_THREADS), smem, (cudaStream_t)stream, feat, w_assign, b_assign, score, centers, out, reinterpret_cast<__half*>(out_planes), S, D, K, G, wa_smem);
  return check_launch("sar_vlad_fwd");
}
