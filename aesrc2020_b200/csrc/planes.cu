// planes.cu -- conversions between dense fp32 NHWC maps and the "flat-pad hi/lo planes" layout
// the tensor-core convolution consumes (see conv_tc.cu), plus MaxPooling2D(3x3,s2,'same')
// (resnet.py:174,192) writing straight into that layout.
//
// planes tensor: [nplanes][R][C] fp16, R = B*(H+1)*(W+1), row q = n*(H+1)*(W+1) + h*(W+1) + w;
// plane 0 = hi = fp16(x), plane 1 = lo = fp16((x-hi)*2^11).  Phase-split tensors hold 4 such
// pairs (plane = 2*((h&1)*2+(w&1)) + {0,1}) over the half-resolution geometry.
// Pad rows (w == W or h == H) are never written: buffers are zero-initialised once by the caller.
#include "common.cuh"

namespace sar {

struct PlaneGeom {
  int B, H, W, C;
  int split;
};

__device__ __forceinline__ void plane_addr(const PlaneGeom& g, int n, int h, int w, long long& row, long long& R, int& plane0) {
  if (g.split) {
    const int P2 = (g.W + 1) / 2 + 1, Rimg2 = ((g.H + 1) / 2 + 1) * P2;
    row = (long long)n * Rimg2 + (h >> 1) * P2 + (w >> 1);
    R = (long long)g.B * Rimg2;
    plane0 = 2 * ((h & 1) * 2 + (w & 1));
  } else {
    const int P = g.W + 1, Rimg = (g.H + 1) * P;
    row = (long long)n * Rimg + h * P + w;
    R = (long long)g.B * Rimg;
    plane0 = 0;
  }
}

__device__ __forceinline__ void store_hilo8(__half* planes, long long R, int plane0, long long row, int C, int c, const float* v) {
  uint32_t hh[4], ll[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __half h0 = __float2half_rn(v[2 * e]), h1 = __float2half_rn(v[2 * e + 1]);
    hh[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    __half2 l = __floats2half2_rn((v[2 * e] - __half2float(h0)) * 2048.f, (v[2 * e + 1] - __half2float(h1)) * 2048.f);
    ll[e] = *reinterpret_cast<uint32_t*>(&l);
  }
  *reinterpret_cast<uint4*>(planes + ((size_t)plane0 * R + row) * C + c) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
  *reinterpret_cast<uint4*>(planes + ((size_t)(plane0 + 1) * R + row) * C + c) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}

// dense fp32 NHWC -> planes, optional per-channel affine + ReLU; one thread per (pixel, 8 channels)
__global__ void planes_pack_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                   int relu, __half* __restrict__ planes, PlaneGeom g) {
  pdl_wait();
  pdl_trigger();
  const int C8 = g.C >> 3;
  const long long total = (long long)g.B * g.H * g.W * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    long long pix = i / C8;
    const int w = (int)(pix % g.W);
    pix /= g.W;
    const int h = (int)(pix % g.H);
    const int n = (int)(pix / g.H);
    const float4* src = reinterpret_cast<const float4*>(x + (((size_t)n * g.H + h) * g.W + w) * g.C + c);
    float4 a = __ldg(src), b = __ldg(src + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], __ldg(scale + c + j), __ldg(shift + c + j));
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = relu_nan(v[j]);
    }
    long long row, R; int plane0;
    plane_addr(g, n, h, w, row, R, plane0);
    store_hilo8(planes, R, plane0, row, g.C, c, v);
  }
}

__global__ void planes_unpack_kernel(const __half* __restrict__ planes, float* __restrict__ x, PlaneGeom g) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)g.B * g.H * g.W * g.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.C);
    long long pix = i / g.C;
    const int w = (int)(pix % g.W);
    pix /= g.W;
    const int h = (int)(pix % g.H);
    const int n = (int)(pix / g.H);
    long long row, R; int plane0;
    plane_addr(g, n, h, w, row, R, plane0);
    const float hi = __half2float(planes[((size_t)plane0 * R + row) * g.C + c]);
    const float lo = __half2float(planes[((size_t)(plane0 + 1) * R + row) * g.C + c]);
    x[i] = hi + lo * (1.f / 2048.f);
  }
}

// MaxPooling2D 'same' from dense fp32 NHWC into planes (pool output geometry in g)
__global__ void maxpool_planes_kernel(const float* __restrict__ x, __half* __restrict__ planes, int Hin, int Win,
                                      int k, int stride, int pad_t, int pad_l, PlaneGeom g) {
  pdl_wait();
  pdl_trigger();
  const int C8 = g.C >> 3;
  const long long total = (long long)g.B * g.H * g.W * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    long long pix = i / C8;
    const int wo = (int)(pix % g.W);
    pix /= g.W;
    const int ho = (int)(pix % g.H);
    const int n = (int)(pix / g.H);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = -INFINITY;
    for (int dh = 0; dh < k; ++dh) {
      const int hi = ho * stride - pad_t + dh;
      if (hi < 0 || hi >= Hin) continue;
      for (int dw = 0; dw < k; ++dw) {
        const int wi = wo * stride - pad_l + dw;
        if (wi < 0 || wi >= Win) continue;
        const float4* src = reinterpret_cast<const float4*>(x + (((size_t)n * Hin + hi) * Win + wi) * g.C + c);
        float4 a = __ldg(src), b = __ldg(src + 1);
        v[0] = fmax_nan(v[0], a.x); v[1] = fmax_nan(v[1], a.y); v[2] = fmax_nan(v[2], a.z); v[3] = fmax_nan(v[3], a.w);
        v[4] = fmax_nan(v[4], b.x); v[5] = fmax_nan(v[5], b.y); v[6] = fmax_nan(v[6], b.z); v[7] = fmax_nan(v[7], b.w);
      }
    }
    long long row, R; int plane0;
    plane_addr(g, n, ho, wo, row, R, plane0);
    store_hilo8(planes, R, plane0, row, g.C, c, v);
  }
}

static inline unsigned pgrid(long long total) {
  long long gsz = (total + 255) / 256;
  const long long cap = 148ll * 16;
  return (unsigned)(gsz < cap ? (gsz > 0 ? gsz : 1) : cap);
}

}  // namespace sar

extern "C" {

size_t sar_planes_bytes(int B, int H, int W, int C, int split) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
  if (split) {
    size_t P2 = (size_t)(W + 1) / 2 + 1, Rimg2 = ((size_t)(H + 1) / 2 + 1) * P2;
    return 8 * (size_t)B * Rimg2 * C * 2;
  }
  return 2 * (size_t)B * (H + 1) * (W + 1) * C * 2;
}

int sar_planes_pack_fwd(const float* x, const float* scale, const float* shift, int relu, void* planes,
                        int B, int H, int W, int C, int split, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && planes, SAR_ERR_BAD_ARG, "sar_planes_pack_fwd: null pointer");
  SAR_REQUIRE((scale == nullptr) == (shift == nullptr), SAR_ERR_BAD_ARG, "sar_planes_pack_fwd: scale/shift must pair");
  SAR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, SAR_ERR_BAD_ARG, "sar_planes_pack_fwd: non-positive dimension");
  SAR_REQUIRE(C % 8 == 0, SAR_ERR_UNSUPPORTED, "sar_planes_pack_fwd: C must be a multiple of 8");
  SAR_REQUIRE(aligned16(x) && aligned16(planes), SAR_ERR_ALIGN, "sar_planes_pack_fwd: unaligned pointer");
  PlaneGeom g{B, H, W, C, split ? 1 : 0};
  launch_k(planes_pack_kernel, dim3(pgrid((long long)B * H * W * (C / 8))), dim3(256), 0, (cudaStream_t)stream, x, scale, shift, relu, reinterpret_cast<__half*>(planes), g);
  return check_launch("sar_planes_pack_fwd");
}

int sar_planes_unpack_fwd(const void* planes, float* x, int B, int H, int W, int C, int split, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && planes, SAR_ERR_BAD_ARG, "sar_planes_unpack_fwd: null pointer");
  SAR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, SAR_ERR_BAD_ARG, "sar_planes_unpack_fwd: non-positive dimension");
  PlaneGeom g{B, H, W, C, split ? 1 : 0};
  launch_k(planes_unpack_kernel, dim3(pgrid((long long)B * H * W * C)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const __half*>(planes), x, g);
  return check_launch("sar_planes_unpack_fwd");
}

int sar_maxpool_planes_fwd(const float* x, void* planes, int B, int H, int W, int C, int Ho, int Wo,
                           int k, int stride, int pad_t, int pad_l, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && planes, SAR_ERR_BAD_ARG, "sar_maxpool_planes_fwd: null pointer");
  SAR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && k > 0 && stride > 0, SAR_ERR_BAD_ARG,
              "sar_maxpool_planes_fwd: non-positive dimension");
  SAR_REQUIRE(C % 8 == 0, SAR_ERR_UNSUPPORTED, "sar_maxpool_planes_fwd: C must be a multiple of 8");
  SAR_REQUIRE(aligned16(x) && aligned16(planes), SAR_ERR_ALIGN, "sar_maxpool_planes_fwd: unaligned pointer");
  PlaneGeom g{B, Ho, Wo, C, 0};
  launch_k(maxpool_planes_kernel, dim3(pgrid((long long)B * Ho * Wo * (C / 8))), dim3(256), 0, (cudaStream_t)stream, x, reinterpret_cast<__half*>(planes), H, W, k, stride, pad_t, pad_l, g);
  return check_launch("sar_maxpool_planes_fwd");
}

}  // extern "C"
