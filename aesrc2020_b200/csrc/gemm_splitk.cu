// gemm_splitk.cu -- AR_EMBEDDING: (B x K*D=16384) @ (16384 x 256), model.py:286-289.
//
// M = batch is small (64..512) and K is huge, so the weight read (16.8 MB fp32) dominates:
// HBM-bound.  Split K across CTAs so the whole chip streams the weight once; partial tiles go
// to a workspace and are summed in a FIXED order (deterministic, batch-size independent per
// row) by a second small kernel that also adds the (BN-folded) bias.
#include "common.cuh"

namespace sar {

constexpr int SK_BM = 64, SK_BN = 64, SK_BK = 16, SK_KSLAB = 512;

__global__ void __launch_bounds__(256) gemm_splitk_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                          float* __restrict__ ws, int M, int K, int N) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float As[SK_BK][SK_BM + 4];
  __shared__ __align__(16) float Bs[SK_BK][SK_BN];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * SK_BM, n0 = blockIdx.y * SK_BN;
  const int kbeg = blockIdx.z * SK_KSLAB;
  const int kend = min(K, kbeg + SK_KSLAB);
  const int a_m = t >> 2, a_k = (t & 3) * 4;         // 64 rows x 16 k
  const int b_k = t >> 4, b_n = (t & 15) * 4;        // 16 k x 64 n
  const int ty = t >> 4, tx = t & 15;                // 4x4 outputs per thread
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += SK_BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + a_m < M && k0 + a_k + 3 < kend)
      av = __ldg(reinterpret_cast<const float4*>(a + (size_t)(m0 + a_m) * K + k0 + a_k));
    else if (m0 + a_m < M) {
      float tmp[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < 4; ++i) if (k0 + a_k + i < kend) tmp[i] = __ldg(a + (size_t)(m0 + a_m) * K + k0 + a_k + i);
      av = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + b_k < kend) {
      if (n0 + b_n + 3 < N) bv = __ldg(reinterpret_cast<const float4*>(w + (size_t)(k0 + b_k) * N + n0 + b_n));
      else {
        float tmp[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j) if (n0 + b_n + j < N) tmp[j] = __ldg(w + (size_t)(k0 + b_k) * N + n0 + b_n + j);
        bv = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
      }
    }
    __syncthreads();
    As[a_k + 0][a_m] = av.x; As[a_k + 1][a_m] = av.y; As[a_k + 2][a_m] = av.z; As[a_k + 3][a_m] = av.w;
    *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK_BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  float* dst = ws + (size_t)blockIdx.z * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) dst[(size_t)m * N + n] = acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ bias,
                                     float* __restrict__ out, int M, int N, int splits) {
  pdl_wait();
  pdl_trigger();
  long long total = (long long)M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = bias ? __ldg(bias + (int)(i % N)) : 0.f;
    int z = 0;
    for (; z + 8 <= splits; z += 8) {           // 8 independent loads in flight, summed in the fixed order z = 0, 1, 2, ...
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ws[(size_t)(z + j) * total + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j];
    }
    for (; z < splits; ++z) acc += ws[(size_t)z * total + i];
    out[i] = acc;
  }
}

}  // namespace sar

extern "C" {

size_t sar_gemm_splitk_workspace_bytes(int M, int K, int N) {
  if (M <= 0 || K <= 0 || N <= 0) return 0;
  size_t splits = (size_t)(K + sar::SK_KSLAB - 1) / sar::SK_KSLAB;
  return splits * (size_t)M * N * sizeof(float);
}

int sar_gemm_splitk_fwd(const float* a, const float* w, const float* bias, float* out,
                        int M, int K, int N, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace sar;
  SAR_REQUIRE(a && w && out && workspace, SAR_ERR_BAD_ARG, "sar_gemm_splitk_fwd: null pointer");
  SAR_REQUIRE(M > 0 && K > 0 && N > 0, SAR_ERR_BAD_ARG, "sar_gemm_splitk_fwd: non-positive dimension");
  SAR_REQUIRE(K % 4 == 0 && N % 4 == 0, SAR_ERR_UNSUPPORTED, "sar_gemm_splitk_fwd: K and N must be multiples of 4");
  SAR_REQUIRE(aligned16(a) && aligned16(w) && aligned16(out) && aligned16(workspace), SAR_ERR_ALIGN,
              "sar_gemm_splitk_fwd: unaligned pointer");
  SAR_REQUIRE(workspace_bytes >= sar_gemm_splitk_workspace_bytes(M, K, N), SAR_ERR_WORKSPACE,
              "sar_gemm_splitk_fwd: workspace too small (%zu < %zu)", workspace_bytes,
              sar_gemm_splitk_workspace_bytes(M, K, N));
  int splits = (K + SK_KSLAB - 1) / SK_KSLAB;
  dim3 grid((M + SK_BM - 1) / SK_BM, (N + SK_BN - 1) / SK_BN, splits);
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(gemm_splitk_kernel, dim3(grid), dim3(256), 0, st, a, w, (float*)workspace, M, K, N);
  int rc = check_launch("sar_gemm_splitk_fwd(partial)");
  if (rc) return rc;
  long long total = (long long)M * N;
  unsigned rg = (unsigned)((total + 255) / 256);
  if (rg > 148 * 8) rg = 148 * 8;
  launch_k(splitk_reduce_kernel, dim3(rg), dim3(256), 0, st, (const float*)workspace, bias, out, M, N, splits);
  return check_launch("sar_gemm_splitk_fwd(reduce)");
}

int sar_splitk_reduce_fwd(const float* ws, const float* bias, float* out, int M, int N, int splits, void* stream) {
  using namespace sar;
  SAR_REQUIRE(ws && out && M > 0 && N > 0 && splits > 0, SAR_ERR_BAD_ARG, "sar_splitk_reduce_fwd: bad argument");
  long long total = (long long)M * N;
  unsigned rg = (unsigned)((total + 255) / 256);
  if (rg > 148 * 8) rg = 148 * 8;
  launch_k(splitk_reduce_kernel, dim3(rg), dim3(256), 0, (cudaStream_t)stream, ws, bias, out, M, N, splits);
  return check_launch("sar_splitk_reduce_fwd");
}

}  // extern "C"
