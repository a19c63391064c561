// bigru.cu -- recurrent step loop of Bidirectional(CuDNNGRU) (model.py:44-50).
//
// The input projections x*W + b_input of all S steps and both directions are one GEMM done
// by the caller (sar_conv2d_fwd against the concatenated [forward | backward] kernels); this
// kernel runs the S strictly sequential steps
//     hp = h @ U + b_rec                                   (u x 3u, gate order z|r|h)
//     z = sigmoid(xz + hpz)   r = sigmoid(xr + hpr)   hh = tanh(xh + r * hph)
//     h' = z*h + (1-z)*hh                                  (reset_after / cuDNN form)
// as ONE persistent launch: a CTA owns BG utterances of one direction for the whole sequence,
// h stays in shared memory, thread c owns recurrent column c (3u = 768 threads), the
// recurrent kernel is streamed from L2 each step (coalesced rows).  Both directions run
// concurrently (blockIdx.y).
#include "common.cuh"

namespace sar {

constexpr int GRU_U = 256;
constexpr int GRU_BG = 4;      // utterances per CTA

__global__ void __launch_bounds__(3 * GRU_U) bigru_kernel(const float* __restrict__ xp, const float* __restrict__ rec,
                                                           const float* __restrict__ rbias, float* __restrict__ out,
                                                           int B, int S, int seq) {
  constexpr int U = GRU_U, U3 = 3 * GRU_U, BG = GRU_BG;
  __shared__ __align__(16) float hs[BG][U];
  __shared__ float hp[BG][U3];

  const int c = threadIdx.x;            // recurrent column 0..767
  const int dir = blockIdx.y;           // 0 forward, 1 backward
  const int b0 = blockIdx.x * BG;
  const float* Ud = rec + (size_t)dir * U * U3;
  const float rb = __ldg(rbias + dir * U3 + c);

  for (int i = c; i < BG * U; i += U3) (&hs[0][0])[i] = 0.f;
  __syncthreads();

  for (int step = 0; step < S; ++step) {
    const int t = dir ? (S - 1 - step) : step;
    // ---- hp[b][c] = sum_k h[b][k] * U[k][c] + b_rec[c]
    float acc[BG];
#pragma unroll
    for (int b = 0; b < BG; ++b) acc[b] = 0.f;
#pragma unroll 8
    for (int k = 0; k < U; k += 4) {
      float u0 = __ldg(Ud + (size_t)(k + 0) * U3 + c);
      float u1 = __ldg(Ud + (size_t)(k + 1) * U3 + c);
      float u2 = __ldg(Ud + (size_t)(k + 2) * U3 + c);
      float u3 = __ldg(Ud + (size_t)(k + 3) * U3 + c);
#pragma unroll
      for (int b = 0; b < BG; ++b) {
        float4 h4 = *reinterpret_cast<const float4*>(&hs[b][k]);
        acc[b] = fmaf(h4.x, u0, acc[b]);
        acc[b] = fmaf(h4.y, u1, acc[b]);
        acc[b] = fmaf(h4.z, u2, acc[b]);
        acc[b] = fmaf(h4.w, u3, acc[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < BG; ++b) hp[b][c] = acc[b] + rb;
    __syncthreads();
    // ---- gates: BG*U items over 768 threads
    for (int i = c; i < BG * U; i += U3) {
      const int b = i / U, j = i - b * U;
      if (b0 + b < B) {
        const float* xrow = xp + (((size_t)(b0 + b) * S + t) * 2 + dir) * U3;
        float z = sigmoidf_(__ldg(xrow + j) + hp[b][j]);
        float r = sigmoidf_(__ldg(xrow + U + j) + hp[b][U + j]);
        float hh = tanhf(__ldg(xrow + 2 * U + j) + r * hp[b][2 * U + j]);
        float hn = z * hs[b][j] + (1.f - z) * hh;
        hs[b][j] = hn;
        if (seq) out[((size_t)(b0 + b) * S + t) * (2 * U) + dir * U + j] = hn;
        else if (step == S - 1) out[(size_t)(b0 + b) * (2 * U) + dir * U + j] = hn;
      }
    }
    __syncthreads();
  }
}

}  // namespace sar

extern "C" int sar_bigru_fwd(const float* xp, const float* rec, const float* rbias, float* out,
                             int B, int S, int u, int seq, void* stream) {
  using namespace sar;
  SAR_REQUIRE(xp && rec && rbias && out, SAR_ERR_BAD_ARG, "sar_bigru_fwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0, SAR_ERR_BAD_ARG, "sar_bigru_fwd: non-positive dimension");
  SAR_REQUIRE(u == GRU_U, SAR_ERR_UNSUPPORTED, "sar_bigru_fwd: hidden size %d unsupported (this build: %d)", u, GRU_U);
  dim3 grid((B + GRU_BG - 1) / GRU_BG, 2);
  bigru_kernel<<<grid, 3 * GRU_U, 0, (cudaStream_t)stream>>>(xp, rec, rbias, out, B, S, seq);
  return check_launch("sar_bigru_fwd");
}
