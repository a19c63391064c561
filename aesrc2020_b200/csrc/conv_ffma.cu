// conv_ffma.cu -- generic NHWC convolution / Dense as an fp32 FFMA implicit GEMM.
//
// Replaces the Conv2D call sites of resnet.py:39-42 (7x7/s2 stem, Cin=1) and serves as the
// Dense primitive of model.py:35-42 (kh=kw=1, W=1).  The 3x3 residual-block convolutions
// run on the tcgen05 kernel (conv_tc.cu); this kernel is the CUDA-core path for the layers
// that are not GEMM-shaped enough for the tensor pipe (K=49 stem) and for small GEMMs.
//
// GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin with Cin innermost -- the
// Keras HWIO kernel is then exactly the row-major [K][N] B operand.
// Tile 128x64x16, 256 threads, 8x4 outputs per thread, register-prefetched k-chunks.
#include "common.cuh"

namespace sar {

struct ConvP {
  const float* x; const float* w; const float* bias;
  const float* pre_scale; const float* pre_shift;
  const float* post_scale; const float* post_shift;
  const float* residual; float* out;
  int B, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad_t, pad_l, act;
  int M, K;
};

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;

template <bool VEC_A>
__global__ void __launch_bounds__(256) conv_ffma_kernel(ConvP p) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];

  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- A loader coordinates: row a_m of the tile, 8 consecutive k
  const int a_m = t >> 1, a_k = (t & 1) * 8;
  const int gm = m0 + a_m;
  const bool a_valid = gm < p.M;
  int an = 0, hi0 = 0, wi0 = 0;
  if (a_valid) {
    int hw = p.Ho * p.Wo;
    an = gm / hw;
    int rem = gm - an * hw;
    int ho = rem / p.Wo, wo = rem - ho * p.Wo;
    hi0 = ho * p.stride - p.pad_t;
    wi0 = wo * p.stride - p.pad_l;
  }
  // ---- B loader coordinates
  const int b_k = t >> 4, b_n = (t & 15) * 4;
  const bool vecB = (p.Cout & 3) == 0;

  float a_reg[8];
  float b_reg[4];

  auto load_a = [&](int k0) {
    if (VEC_A) {
      // Cin % 8 == 0: the 8 k's share one tap and are 8 contiguous channels
      int kg = k0 + a_k;
      bool ok = a_valid && kg < p.K;
      int tap = 0, ci = 0, hi = 0, wi = 0;
      if (ok) {
        tap = kg / p.Cin; ci = kg - tap * p.Cin;
        int r = tap / p.kw, s = tap - r * p.kw;
        hi = hi0 + r; wi = wi0 + s;
        ok = hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
      }
      if (ok) {
        const float4* src = reinterpret_cast<const float4*>(
            p.x + (((size_t)an * p.H + hi) * p.W + wi) * p.Cin + ci);
        float4 v0 = __ldg(src), v1 = __ldg(src + 1);
        a_reg[0] = v0.x; a_reg[1] = v0.y; a_reg[2] = v0.z; a_reg[3] = v0.w;
        a_reg[4] = v1.x; a_reg[5] = v1.y; a_reg[6] = v1.z; a_reg[7] = v1.w;
        if (p.pre_scale) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            a_reg[i] = relu_nan(fmaf(a_reg[i], __ldg(p.pre_scale + ci + i), __ldg(p.pre_shift + ci + i)));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) a_reg[i] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int kg = k0 + a_k + i;
        float v = 0.f;
        if (a_valid && kg < p.K) {
          int tap = kg / p.Cin, ci = kg - tap * p.Cin;
          int r = tap / p.kw, s = tap - r * p.kw;
          int hi = hi0 + r, wi = wi0 + s;
          if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
            v = __ldg(p.x + (((size_t)an * p.H + hi) * p.W + wi) * p.Cin + ci);
            if (p.pre_scale) v = relu_nan(fmaf(v, __ldg(p.pre_scale + ci), __ldg(p.pre_shift + ci)));
          }
        }
        a_reg[i] = v;
      }
    }
  };
  auto load_b = [&](int k0) {
    int kg = k0 + b_k, n = n0 + b_n;
    if (kg < p.K && vecB && n + 3 < p.Cout) {
      float4 v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)kg * p.Cout + n));
      b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b_reg[j] = (kg < p.K && n + j < p.Cout) ? __ldg(p.w + (size_t)kg * p.Cout + n + j) : 0.f;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_k + i][a_m] = a_reg[i];
    *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  const int ty = t >> 4, tx = t & 15;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_a(0);
  load_b(0);
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < p.K) { load_a(k0 + BK); load_b(k0 + BK); }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: +bias, +residual, post affine, activation
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + n);
      if (p.residual) v += __ldg(p.residual + (size_t)m * p.Cout + n);
      if (p.post_scale) v = fmaf(v, __ldg(p.post_scale + n), __ldg(p.post_shift + n));
      if (p.act == SAR_ACT_RELU) v = relu_nan(v);
      else if (p.act == SAR_ACT_TANH) v = tanhf(v);
      acc[i][j] = v;
    }
    int n = n0 + tx * TN;
    if (vecB && n + 3 < p.Cout) {
      *reinterpret_cast<float4*>(p.out + (size_t)m * p.Cout + n) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (n + j < p.Cout) p.out[(size_t)m * p.Cout + n + j] = acc[i][j];
    }
  }
}

}  // namespace sar

extern "C" int sar_conv2d_fwd(const float* x, const float* w_hwio, const float* bias,
                              const float* pre_scale, const float* pre_shift,
                              const float* post_scale, const float* post_shift,
                              const float* residual, float* out,
                              int B, int H, int W, int Cin, int Ho, int Wo, int Cout,
                              int kh, int kw, int stride, int pad_t, int pad_l, int act, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && w_hwio && out, SAR_ERR_BAD_ARG, "sar_conv2d_fwd: null x/w/out");
  SAR_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Ho > 0 && Wo > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0,
              SAR_ERR_BAD_ARG, "sar_conv2d_fwd: non-positive dimension");
  SAR_REQUIRE(pad_t >= 0 && pad_l >= 0 && pad_t < kh && pad_l < kw, SAR_ERR_BAD_ARG, "sar_conv2d_fwd: bad padding");
  SAR_REQUIRE((Ho - 1) * stride - pad_t < H && (Wo - 1) * stride - pad_l < W, SAR_ERR_BAD_ARG,
              "sar_conv2d_fwd: output extent reads entirely outside the input");
  SAR_REQUIRE((pre_scale == nullptr) == (pre_shift == nullptr) && (post_scale == nullptr) == (post_shift == nullptr),
              SAR_ERR_BAD_ARG, "sar_conv2d_fwd: scale/shift must come in pairs");
  SAR_REQUIRE(act >= SAR_ACT_NONE && act <= SAR_ACT_TANH, SAR_ERR_BAD_ARG, "sar_conv2d_fwd: bad act %d", act);
  SAR_REQUIRE(aligned16(x) && aligned16(w_hwio) && aligned16(out) && (!residual || aligned16(residual)),
              SAR_ERR_ALIGN, "sar_conv2d_fwd: pointers must be 16-byte aligned");
  long long M = (long long)B * Ho * Wo;
  SAR_REQUIRE(M < (1ll << 31) && (long long)kh * kw * Cin < (1ll << 31), SAR_ERR_UNSUPPORTED, "sar_conv2d_fwd: too large");
  ConvP p{x, w_hwio, bias, pre_scale, pre_shift, post_scale, post_shift, residual, out,
          B, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad_t, pad_l, act, (int)M, kh * kw * Cin};
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Cout + BN - 1) / BN));
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin % 8 == 0) launch_k(conv_ffma_kernel<true>, dim3(grid), dim3(256), 0, st, p);
  else launch_k(conv_ffma_kernel<false>, dim3(grid), dim3(256), 0, st, p);
  return check_launch("sar_conv2d_fwd");
}
