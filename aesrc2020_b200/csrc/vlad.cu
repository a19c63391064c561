// vlad.cu -- NetVLAD / GhostVLAD in one pass over the utterance's descriptors.
//
// Replaces the 1x1 assignment Conv2D (model.py:89-95 / 99-105) + VladPooling.call
// (VLAD.py:26-49).  The reference materialises feat_res and weighted_res, two
// (B,1,S,K+G,D) fp32 tensors (VLAD.py:40-41; 14.2 MB per utterance at K+G=72, D=256, S=48).
// Here one CTA per utterance stages X (S x D) in shared memory ONCE and computes
//     score = X @ Wa + ba ;  A = softmax_k(score)                      VLAD.py:33-35
//     V[k] = sum_s A[s,k] * X[s] - (sum_s A[s,k]) * c[k]     k < K      VLAD.py:38-45
//     out[k] = V[k] / sqrt(max(|V[k]|^2, 1e-12))                        VLAD.py:47-48
// so HBM traffic is the algorithmic minimum: X read once, K*D written once (ghost
// clusters only take part in the softmax and are never accumulated).
#include "common.cuh"

namespace sar {

constexpr int VLAD_THREADS = 256;
constexpr int VLAD_KC = 8;     // clusters accumulated per register block

template <int DSL>   // columns per thread: D <= 256*DSL
__global__ void __launch_bounds__(VLAD_THREADS) vlad_kernel(const float* __restrict__ feat, const float* __restrict__ wa,
                                                             const float* __restrict__ ba, const float* __restrict__ score,
                                                             const float* __restrict__ centers,
                                                             float* __restrict__ out, int S, int D, int K, int G) {
  extern __shared__ __align__(16) float smem[];
  const int KG = K + G;
  const int KGP = (KG + 3) & ~3;                 // padded row of A for float4 reads
  float* X = smem;                                // [S][D]
  float* A = X + (size_t)S * D;                   // [S][KGP]
  float* asum = A + (size_t)S * KGP;              // [KGP]
  float* part = asum + KGP;                       // [VLAD_KC][8 warps]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;

  // ---- stage X (coalesced 16B loads)
  {
    const float4* src = reinterpret_cast<const float4*>(feat + (size_t)b * S * D);
    float4* dst = reinterpret_cast<float4*>(X);
    const int n4 = S * D / 4;
    for (int i = t; i < n4; i += VLAD_THREADS) dst[i] = __ldg(src + i);
  }
  for (int i = t; i < S * KGP; i += VLAD_THREADS) A[i] = 0.f;
  __syncthreads();

  // ---- scores: item (s,k); lanes run over k so Wa reads coalesce and X[s][d] broadcasts
  for (int i = t; i < S * KG; i += VLAD_THREADS) {
    const int s = i / KG, k = i - s * KG;
    float acc;
    if (score) {                                   // VladPooling called with external scores
      acc = __ldg(score + ((size_t)b * S + s) * KG + k);
    } else {                                       // fused 1x1 assignment conv
      const float* xr = X + (size_t)s * D;
      acc = __ldg(ba + k);
#pragma unroll 8
      for (int d = 0; d < D; ++d) acc = fmaf(xr[d], __ldg(wa + (size_t)d * KG + k), acc);
    }
    A[s * KGP + k] = acc;
  }
  __syncthreads();

  // ---- softmax over clusters, one warp per descriptor row
  for (int s = warp; s < S; s += VLAD_THREADS / 32) {
    float m = -INFINITY;
    for (int k = lane; k < KG; k += 32) m = fmaxf(m, A[s * KGP + k]);
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
  for (int k = t; k < KGP; k += VLAD_THREADS) {
    float a = 0.f;
    if (k < KG) for (int s = 0; s < S; ++s) a += A[s * KGP + k];
    asum[k] = a;
  }
  __syncthreads();

  // ---- residual accumulation, VLAD_KC clusters at a time; thread owns columns d = t (+256)
  for (int k0 = 0; k0 < K; k0 += VLAD_KC) {
    float acc[DSL][VLAD_KC];
#pragma unroll
    for (int j = 0; j < DSL; ++j)
#pragma unroll
      for (int q = 0; q < VLAD_KC; ++q) acc[j][q] = 0.f;
    for (int s = 0; s < S; ++s) {
      const float4 a0 = *reinterpret_cast<const float4*>(A + s * KGP + k0);
      const float4 a1 = *reinterpret_cast<const float4*>(A + s * KGP + k0 + 4);
      const float av[VLAD_KC] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int j = 0; j < DSL; ++j) {
        const int d = t + j * VLAD_THREADS;
        const float xv = (d < D) ? X[(size_t)s * D + d] : 0.f;
#pragma unroll
        for (int q = 0; q < VLAD_KC; ++q) acc[j][q] = fmaf(av[q], xv, acc[j][q]);
      }
    }
    float ss[VLAD_KC];
#pragma unroll
    for (int q = 0; q < VLAD_KC; ++q) {
      ss[q] = 0.f;
      const int k = k0 + q;
#pragma unroll
      for (int j = 0; j < DSL; ++j) {
        const int d = t + j * VLAD_THREADS;
        if (k < K && d < D) {
          acc[j][q] -= asum[k] * __ldg(centers + (size_t)k * D + d);
          ss[q] += acc[j][q] * acc[j][q];
        } else {
          acc[j][q] = 0.f;
        }
      }
      ss[q] = warp_sum(ss[q]);
    }
    __syncthreads();                     // previous chunk's `part` readers are done
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < VLAD_KC; ++q) part[q * 8 + warp] = ss[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < VLAD_KC; ++q) {
      const int k = k0 + q;
      if (k >= K) continue;
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += part[q * 8 + w];
      const float inv = 1.0f / sqrtf(fmaxf(tot, 1e-12f));
#pragma unroll
      for (int j = 0; j < DSL; ++j) {
        const int d = t + j * VLAD_THREADS;
        if (d < D) out[((size_t)b * K + k) * D + d] = acc[j][q] * inv;
      }
    }
  }
}

static size_t vlad_smem_bytes(int S, int D, int KG) {
  int KGP = (KG + 3) & ~3;
  return sizeof(float) * ((size_t)S * D + (size_t)S * KGP + KGP + VLAD_KC * 8);
}

}  // namespace sar

extern "C" int sar_vlad_fwd(const float* feat, const float* w_assign, const float* b_assign, const float* score,
                            const float* centers, float* out, int B, int S, int D, int K, int G, void* stream) {
  using namespace sar;
  SAR_REQUIRE(feat && centers && out, SAR_ERR_BAD_ARG, "sar_vlad_fwd: null pointer");
  SAR_REQUIRE((score != nullptr) != (w_assign != nullptr && b_assign != nullptr), SAR_ERR_BAD_ARG,
              "sar_vlad_fwd: pass either (w_assign, b_assign) or score");
  SAR_REQUIRE(B > 0 && S > 0 && D > 0 && K > 0 && G >= 0, SAR_ERR_BAD_ARG, "sar_vlad_fwd: bad dimension");
  SAR_REQUIRE(D % 32 == 0 && D <= 512 && K + G <= 128, SAR_ERR_UNSUPPORTED,
              "sar_vlad_fwd: needs D %% 32 == 0, D <= 512, K+G <= 128 (got D=%d K+G=%d)", D, K + G);
  SAR_REQUIRE(aligned16(feat) && aligned16(out), SAR_ERR_ALIGN, "sar_vlad_fwd: unaligned pointer");
  // A rows are read VLAD_KC at a time: pad the cluster count so reads past K+G stay in the row
  size_t smem = vlad_smem_bytes(S, D, K + G + VLAD_KC);
  SAR_REQUIRE(smem <= 227 * 1024, SAR_ERR_UNSUPPORTED, "sar_vlad_fwd: S*D too large for shared memory (%zu B)", smem);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (D <= 256) {
    e = cudaFuncSetAttribute(vlad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) vlad_kernel<1><<<B, VLAD_THREADS, smem, st>>>(feat, w_assign, b_assign, score, centers, out, S, D, K, G);
  } else {
    e = cudaFuncSetAttribute(vlad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) vlad_kernel<2><<<B, VLAD_THREADS, smem, st>>>(feat, w_assign, b_assign, score, centers, out, S, D, K, G);
  }
  if (e != cudaSuccess) { set_error("sar_vlad_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  return check_launch("sar_vlad_fwd");
}
