// stem_pool.cu -- ResNet stem in ONE kernel: Conv2D 7x7/s2 'same' (+bias) -> BatchNormalization ->
// ReLU -> MaxPooling2D 3x3/s2 'same'  (resnet.py:28-45 via :173/:191, and :174/:192), writing the
// pooled map straight into the flat-pad hi/lo planes the tensor-core blocks consume.
//
// Cin = 1, so the contraction is K = 49: not GEMM-shaped enough for the tensor pipe -- CUDA-core
// FFMA with register tiling (4 output pixels x 16 channels per thread, sliding 13-value input
// window per kernel row, weights broadcast from shared memory: ~11 FMA per shared-memory load).
// HBM: reads the fp32 input once (T*80*4 B / utterance), writes only the 4x smaller pooled map;
// the (T/2)x40xF0 conv map (2.56 MB / utterance at T=500) never leaves shared memory.
//
// One CTA = one utterance x PH=4 pooled rows (9 conv rows, 23 input rows).
#include <stdlib.h>
#include "common.cuh"

namespace sar {

constexpr int SP_PH = 4;                      // pooled rows per CTA
constexpr int SP_CR = 2 * SP_PH + 1;          // conv rows per CTA
constexpr int SP_IR = 2 * (SP_CR - 1) + 7;    // input rows per CTA (23)
constexpr int SP_THREADS = 192;
constexpr int SP_K = 7;

struct StemP {
  const float* x; const float* w; const float* bias; const float* scale; const float* shift;
  __half* planes;
  int B, T, D, F0;
  int Hc, Wc, pt, pl;        // conv output size and leading pads
  int Hp, Wp, ppt, ppl;      // pool output size and leading pads
  int XW;                    // padded input row width in smem
};

__global__ void __launch_bounds__(SP_THREADS, 2) stem_pool_kernel(StemP p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  float* w_s = sm;                               // [49][F0]
  float* x_s = w_s + 49 * p.F0;                  // [SP_IR][XW]
  float* c_s = x_s + SP_IR * p.XW;               // [SP_CR][Wc][F0]
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  const int hp0 = blockIdx.x * SP_PH;
  const int hc0 = 2 * hp0 - p.ppt;               // first conv row of this CTA (may be -1)
  const int hi0 = 2 * hc0 - p.pt;                // first input row
  const int F0 = p.F0;

  for (int i = t; i < 49 * F0; i += SP_THREADS) w_s[i] = __ldg(p.w + i);
  for (int i = t; i < SP_IR * p.XW; i += SP_THREADS) {
    const int r = i / p.XW, c = i - r * p.XW;
    const int hi = hi0 + r, wi = c - p.pl;
    float v = 0.f;                               // TF-SAME zero padding of the INPUT
    if (hi >= 0 && hi < p.T && wi >= 0 && wi < p.D) v = __ldg(p.x + ((size_t)n * p.T + hi) * p.D + wi);
    x_s[i] = v;
  }
  __syncthreads();

  // ---- conv + BN + ReLU into c_s: item = (conv row r, pixel quad q, 16-channel group cg)
  const int quads = p.Wc >> 2, cgs = F0 >> 4;
  const int items = SP_CR * quads * cgs;
  for (int item = t; item < items; item += SP_THREADS) {
    const int cg = item % cgs;
    const int rq = item / cgs;
    const int q = rq % quads, r = rq / quads;
    const int hc = hc0 + r;
    float* dst = c_s + ((size_t)r * p.Wc + 4 * q) * F0 + 16 * cg;
    if (hc < 0 || hc >= p.Hc) {                  // outside the conv map: never wins the max-pool
#pragma unroll
      for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int c = 0; c < 16; c += 4)
          *reinterpret_cast<float4*>(dst + px * F0 + c) = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      continue;
    }
    float acc[4][16];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[px][c] = 0.f;
#pragma unroll 1
    for (int kr = 0; kr < SP_K; ++kr) {
      const float* xr = x_s + (2 * r + kr) * p.XW + 8 * q;
      float xin[13];
#pragma unroll
      for (int j = 0; j < 13; ++j) xin[j] = xr[j];
#pragma unroll
      for (int kc = 0; kc < SP_K; ++kc) {
        const float4* wp4 = reinterpret_cast<const float4*>(w_s + (kr * SP_K + kc) * F0 + 16 * cg);
        const float4 w0 = wp4[0], w1 = wp4[1], w2 = wp4[2], w3 = wp4[3];
        const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float xv = xin[2 * px + kc];
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[px][c] = fmaf(xv, wv[c], acc[px][c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int ch = 16 * cg + c;
      const float b = __ldg(p.bias + ch), s = __ldg(p.scale + ch), sh = __ldg(p.shift + ch);
#pragma unroll
      for (int px = 0; px < 4; ++px) acc[px][c] = relu_nan(fmaf(acc[px][c] + b, s, sh));
    }
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int c = 0; c < 16; c += 4)
        *reinterpret_cast<float4*>(dst + px * F0 + c) = make_float4(acc[px][c], acc[px][c + 1], acc[px][c + 2], acc[px][c + 3]);
  }
  __syncthreads();

  // ---- 3x3/s2 max-pool from c_s, hi/lo split, store to planes (8 channels per thread-item)
  const int c8s = F0 >> 3;
  const int pitems = SP_PH * p.Wp * c8s;
  const int P = p.Wp + 1;
  const long long Rimg = (long long)(p.Hp + 1) * P, R = (long long)p.B * Rimg;
  for (int item = t; item < pitems; item += SP_THREADS) {
    const int c8 = item % c8s;
    const int rw = item / c8s;
    const int wp = rw % p.Wp, dh = rw / p.Wp;
    const int hp = hp0 + dh;
    if (hp >= p.Hp) continue;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int r = 2 * dh + a;                       // conv row index inside the CTA (hc = hc0 + r)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int wc = 2 * wp - p.ppl + b;
        if (wc < 0 || wc >= p.Wc) continue;
        const float4* src = reinterpret_cast<const float4*>(c_s + ((size_t)r * p.Wc + wc) * F0 + 8 * c8);
        const float4 v0 = src[0], v1 = src[1];
        m[0] = fmax_nan(m[0], v0.x); m[1] = fmax_nan(m[1], v0.y); m[2] = fmax_nan(m[2], v0.z); m[3] = fmax_nan(m[3], v0.w);
        m[4] = fmax_nan(m[4], v1.x); m[5] = fmax_nan(m[5], v1.y); m[6] = fmax_nan(m[6], v1.z); m[7] = fmax_nan(m[7], v1.w);
      }
    }
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half h0 = __float2half_rn(m[2 * e]), h1 = __float2half_rn(m[2 * e + 1]);
      hh[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      __half2 l = __floats2half2_rn((m[2 * e] - __half2float(h0)) * 2048.f, (m[2 * e + 1] - __half2float(h1)) * 2048.f);
      ll[e] = *reinterpret_cast<uint32_t*>(&l);
    }
    const long long row = (long long)n * Rimg + (long long)hp * P + wp;
    *reinterpret_cast<uint4*>(p.planes + (size_t)row * F0 + 8 * c8) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<uint4*>(p.planes + ((size_t)R + row) * F0 + 8 * c8) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
}

int stem_tc_launch(const float* x, const float* w, const float* bias, const float* scale, const float* shift,
                   void* planes, int B, int T, int D, int F0, int Hc, int pt, int Wc, int pl, int Hp, int ppt,
                   int Wp, int ppl, cudaStream_t stream);      // stem_tc.cu

static inline void same_pad_host(int n_in, int k, int s, int* n_out, int* pad_before) {
  *n_out = (n_in + s - 1) / s;
  int total = (*n_out - 1) * s + k - n_in;
  if (total < 0) total = 0;
  *pad_before = total / 2;
}

}  // namespace sar

extern "C" int sar_stem_pool_fwd(const float* x, const float* w, const float* bias, const float* scale,
                                 const float* shift, void* planes, int B, int T, int D, int F0, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && w && bias && scale && shift && planes, SAR_ERR_BAD_ARG, "sar_stem_pool_fwd: null pointer");
  SAR_REQUIRE(B > 0 && T > 0 && D > 0, SAR_ERR_BAD_ARG, "sar_stem_pool_fwd: non-positive dimension");
  SAR_REQUIRE(B <= 65535, SAR_ERR_UNSUPPORTED, "sar_stem_pool_fwd: B > 65535");
  SAR_REQUIRE(F0 % 16 == 0 && F0 <= 64, SAR_ERR_UNSUPPORTED, "sar_stem_pool_fwd: stem filters must be 16..64, multiple of 16");
  SAR_REQUIRE(aligned16(x) && aligned16(planes) && aligned16(w), SAR_ERR_ALIGN, "sar_stem_pool_fwd: unaligned pointer");
  StemP p{};
  p.x = x; p.w = w; p.bias = bias; p.scale = scale; p.shift = shift; p.planes = reinterpret_cast<__half*>(planes);
  p.B = B; p.T = T; p.D = D; p.F0 = F0;
  same_pad_host(T, 7, 2, &p.Hc, &p.pt);
  same_pad_host(D, 7, 2, &p.Wc, &p.pl);
  same_pad_host(p.Hc, 3, 2, &p.Hp, &p.ppt);
  same_pad_host(p.Wc, 3, 2, &p.Wp, &p.ppl);
  // D == 80, F0 == 64 (every res34 stem, default res18): tensor-core kernel (stem_tc.cu); else CUDA-core FFMA below
  if (!getenv("SAR_STEM_FFMA")) {
    const int rc = stem_tc_launch(x, w, bias, scale, shift, planes, B, T, D, F0, p.Hc, p.pt, p.Wc, p.pl, p.Hp, p.ppt,
                                  p.Wp, p.ppl, (cudaStream_t)stream);
    if (rc != SAR_ERR_UNSUPPORTED) return rc;
  }
  SAR_REQUIRE(p.Wc % 4 == 0, SAR_ERR_UNSUPPORTED, "sar_stem_pool_fwd: feature dim must give a conv width multiple of 4 (D=%d)", D);
  p.XW = ((8 * (p.Wc / 4 - 1) + 13) + 3) & ~3;      // widest column any thread touches, rounded to 4
  if (p.XW < D + p.pl + 4) p.XW = (D + p.pl + 7) & ~3;
  const size_t smem = sizeof(float) * ((size_t)49 * F0 + (size_t)SP_IR * p.XW + (size_t)SP_CR * p.Wc * F0);
  SAR_REQUIRE(smem <= 113 * 1024, SAR_ERR_UNSUPPORTED, "sar_stem_pool_fwd: feature dim too wide for shared memory (%zu B)", smem);
  { const int arc = allow_max_smem(stem_pool_kernel, "sar_stem_pool_fwd"); if (arc) return arc; }
  dim3 grid((p.Hp + SP_PH - 1) / SP_PH, B);
  launch_k(stem_pool_kernel, dim3(grid), dim3(SP_THREADS), smem, (cudaStream_t)stream, p);
  return check_launch("sar_stem_pool_fwd");
}
