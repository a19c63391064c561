// elementwise.cu -- HBM-bound helper kernels: max-pool, BN-affine+ReLU, LayerNorm,
// global average pool, batch loss reduction.  All fp32, vectorised where shapes allow.
#include "common.cuh"

namespace sar {

// MaxPooling2D 'same' (resnet.py:174,192): one thread per (pixel, 4 channels)
__global__ void maxpool_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int H, int W, int C,
                               int Ho, int Wo, int k, int stride, int pad_t, int pad_l) {
  pdl_wait();
  pdl_trigger();
  const int C4 = C >> 2;
  long long total = (long long)B * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    long long pix = i / C4;
    int wo = (int)(pix % Wo);
    long long r = pix / Wo;
    int ho = (int)(r % Ho);
    int n = (int)(r / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dh = 0; dh < k; ++dh) {
      int hi = ho * stride - pad_t + dh;
      if (hi < 0 || hi >= H) continue;
      for (int dw = 0; dw < k; ++dw) {
        int wi = wo * stride - pad_l + dw;
        if (wi < 0 || wi >= W) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + hi) * W + wi) * C) + c4);
        m.x = fmax_nan(m.x, v.x); m.y = fmax_nan(m.y, v.y); m.z = fmax_nan(m.z, v.z); m.w = fmax_nan(m.w, v.w);
      }
    }
    reinterpret_cast<float4*>(out + (size_t)pix * C)[c4] = m;
  }
}

__global__ void affine_relu_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                   const float* __restrict__ shift, float* __restrict__ out,
                                   long long rows, int C, int relu) {
  pdl_wait();
  pdl_trigger();
  const int C4 = C >> 2;
  long long total = rows * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 s = __ldg(reinterpret_cast<const float4*>(scale) + c4);
    float4 b = __ldg(reinterpret_cast<const float4*>(shift) + c4);
    v.x = fmaf(v.x, s.x, b.x); v.y = fmaf(v.y, s.y, b.y); v.z = fmaf(v.z, s.z, b.z); v.w = fmaf(v.w, s.w, b.w);
    if (relu) { v.x = relu_nan(v.x); v.y = relu_nan(v.y); v.z = relu_nan(v.z); v.w = relu_nan(v.w); }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// LayerNormalization (model.py:32-33): one warp per row, values held in registers,
// two-pass mean / biased variance in fp32 (eps = 1e-14 amplifies a sloppy variance).
template <int MAXV>   // MAXV float4 per lane => C <= 128*MAXV
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ out,
                                 __half* __restrict__ planes, long long plane_rows, int seg,
                                 long long rows, int C, float eps) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int C4 = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  float4 v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c4 = lane + 32 * i;
    if (c4 < C4) {
      v[i] = __ldg(xr + c4);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c4 = lane + 32 * i;
    if (c4 < C4) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float var = warp_sum(sq) / (float)C;
  const float rstd = 1.0f / sqrtf(var + eps);
  float4* orow = out ? reinterpret_cast<float4*>(out + row * C) : nullptr;
  const long long prow = seg > 0 ? row + row / seg : row;
  __half* ph = planes ? planes + prow * C : nullptr;
  __half* pl = planes ? planes + (plane_rows + prow) * C : nullptr;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c4 = lane + 32 * i;
    if (c4 < C4) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (orow) orow[c4] = o;
      if (planes) {                       // x = hi + lo/2048 (conv_tc.cu operand format)
        const __half2 h01 = __floats2half2_rn(o.x, o.y), h23 = __floats2half2_rn(o.z, o.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((o.x - f01.x) * 2048.f, (o.y - f01.y) * 2048.f);
        const __half2 l23 = __floats2half2_rn((o.z - f23.x) * 2048.f, (o.w - f23.y) * 2048.f);
        *reinterpret_cast<uint2*>(ph + 4 * c4) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(pl + 4 * c4) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
    }
  }
}

// GlobalAveragePooling1D (model.py:125): block per utterance, thread per feature
__global__ void avgpool_kernel(const float* __restrict__ x, float* __restrict__ out, int S, int D) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += __ldg(x + ((size_t)b * S + s) * D + d);
    out[(size_t)b * D + d] = acc / (float)S;
  }
}

// one block, fixed reduction order => bitwise reproducible sums for a given B
__global__ void loss_reduce_kernel(const float* __restrict__ stats, const float* __restrict__ ctc,
                                   const float* __restrict__ bn, float* __restrict__ out8, int B) {
  pdl_wait();
  pdl_trigger();
  __shared__ float scratch[32];
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    if (stats) { a[0] += stats[4 * b]; a[1] += stats[4 * b + 1]; a[4] += stats[4 * b + 2]; a[5] += stats[4 * b + 3]; }
    if (ctc) a[2] += ctc[b];
    if (bn) { a[3] += bn[4 * b + 1]; a[7] += bn[4 * b + 3]; }
    a[6] += 1.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float r = block_sum(a[i], scratch);
    if (threadIdx.x == 0) out8[i] = r;
  }
}

static inline unsigned grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  long long cap = 148ll * 16;
  return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

// Row softmax (Dense(..., activation='softmax'), model.py:35-42,268; the ctc_pred posteriors K.ctc_decode reads,
// model.py:385-389): one warp per row, the row is read twice (max, then exp/sum with the values kept in registers for
// C <= 1024, re-read beyond), written once.  `ld` >= C is the row pitch of the input (tensor-core ctc_pred pads 1000
// classes to 1024 columns); the output is dense (rows, C).
__global__ void softmax_rows_kernel(const float* __restrict__ x, int ld, float* __restrict__ out, long long rows, int C) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + (size_t)row * ld;
  float* orow = out + (size_t)row * C;
  constexpr int KEEP = 32;                           // values per lane kept in registers (C <= 1024)
  float v[KEEP];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < KEEP; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < C ? __ldg(xr + c) : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  for (int c = lane + 32 * KEEP; c < C; c += 32) m = fmaxf(m, __ldg(xr + c));
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < KEEP; ++i) {
    v[i] = (lane + 32 * i < C) ? expf(v[i] - m) : 0.f;
    sum += v[i];
  }
  for (int c = lane + 32 * KEEP; c < C; c += 32) sum += expf(__ldg(xr + c) - m);
  sum = warp_sum(sum);
#pragma unroll
  for (int i = 0; i < KEEP; ++i) {
    const int c = lane + 32 * i;
    if (c < C) orow[c] = v[i] / sum;
  }
  for (int c = lane + 32 * KEEP; c < C; c += 32) orow[c] = expf(__ldg(xr + c) - m) / sum;
}

}  // namespace sar

extern "C" {

int sar_softmax_rows_fwd(const float* x, int ld, float* out, long long rows, int C, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && out, SAR_ERR_BAD_ARG, "sar_softmax_rows_fwd: null pointer");
  SAR_REQUIRE(rows > 0 && C > 0 && ld >= C, SAR_ERR_BAD_ARG, "sar_softmax_rows_fwd: need rows > 0, 0 < C <= ld");
  const int warps = 8;
  launch_k(softmax_rows_kernel, dim3((unsigned)((rows + warps - 1) / warps)), dim3(warps * 32), 0, (cudaStream_t)stream, x, ld, out, rows, C);
  return check_launch("sar_softmax_rows_fwd");
}

int sar_maxpool2d_fwd(const float* x, float* out, int B, int H, int W, int C, int Ho, int Wo,
                      int k, int stride, int pad_t, int pad_l, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && out, SAR_ERR_BAD_ARG, "sar_maxpool2d_fwd: null pointer");
  SAR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && k > 0 && stride > 0, SAR_ERR_BAD_ARG,
              "sar_maxpool2d_fwd: non-positive dimension");
  SAR_REQUIRE(C % 4 == 0, SAR_ERR_UNSUPPORTED, "sar_maxpool2d_fwd: C must be a multiple of 4");
  SAR_REQUIRE(aligned16(x) && aligned16(out), SAR_ERR_ALIGN, "sar_maxpool2d_fwd: unaligned pointer");
  long long total = (long long)B * Ho * Wo * (C / 4);
  launch_k(maxpool_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, out, B, H, W, C, Ho, Wo, k, stride, pad_t, pad_l);
  return check_launch("sar_maxpool2d_fwd");
}

int sar_affine_relu_fwd(const float* x, const float* scale, const float* shift, float* out,
                        long long rows, int C, int relu, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && scale && shift && out, SAR_ERR_BAD_ARG, "sar_affine_relu_fwd: null pointer");
  SAR_REQUIRE(rows > 0 && C > 0, SAR_ERR_BAD_ARG, "sar_affine_relu_fwd: non-positive dimension");
  SAR_REQUIRE(C % 4 == 0, SAR_ERR_UNSUPPORTED, "sar_affine_relu_fwd: C must be a multiple of 4");
  SAR_REQUIRE(aligned16(x) && aligned16(out) && aligned16(scale) && aligned16(shift), SAR_ERR_ALIGN,
              "sar_affine_relu_fwd: unaligned pointer");
  launch_k(affine_relu_kernel, dim3(grid_for(rows * (C / 4), 256)), dim3(256), 0, (cudaStream_t)stream, x, scale, shift, out, rows, C, relu);
  return check_launch("sar_affine_relu_fwd");
}

int sar_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* out,
                      long long rows, int C, float eps, void* stream) {
  SAR_REQUIRE(out, SAR_ERR_BAD_ARG, "sar_layernorm_fwd: null pointer");
  return sar_layernorm_planes_fwd(x, gamma, beta, out, nullptr, 0, 0, rows, C, eps, stream);
}

int sar_layernorm_planes_fwd(const float* x, const float* gamma, const float* beta, float* out, void* planes_v,
                             long long plane_rows, int seg, long long rows, int C, float eps, void* stream) {
  using namespace sar;
  __half* planes = reinterpret_cast<__half*>(planes_v);
  SAR_REQUIRE(x && gamma && beta && (out || planes), SAR_ERR_BAD_ARG, "sar_layernorm_fwd: null pointer");
  SAR_REQUIRE(!planes || (C % 8 == 0 && seg >= 0 && plane_rows >= rows + (seg > 0 ? (rows - 1) / seg : 0) && aligned16(planes)),
              SAR_ERR_BAD_ARG, "sar_layernorm_planes_fwd: bad planes geometry (rows=%lld seg=%d plane_rows=%lld C=%d)", rows, seg, plane_rows, C);
  SAR_REQUIRE(rows > 0 && C > 0, SAR_ERR_BAD_ARG, "sar_layernorm_fwd: non-positive dimension");
  SAR_REQUIRE(C % 4 == 0 && C <= 1024, SAR_ERR_UNSUPPORTED, "sar_layernorm_fwd: C must be a multiple of 4, <= 1024");
  SAR_REQUIRE(aligned16(x) && (!out || aligned16(out)) && aligned16(gamma) && aligned16(beta), SAR_ERR_ALIGN,
              "sar_layernorm_fwd: unaligned pointer");
  const int warps = 8;
  unsigned grid = (unsigned)((rows + warps - 1) / warps);
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 256) launch_k(layernorm_kernel<2>, dim3(grid), dim3(warps * 32), 0, st, x, gamma, beta, out, planes, plane_rows, seg, rows, C, eps);
  else if (C <= 512) launch_k(layernorm_kernel<4>, dim3(grid), dim3(warps * 32), 0, st, x, gamma, beta, out, planes, plane_rows, seg, rows, C, eps);
  else launch_k(layernorm_kernel<8>, dim3(grid), dim3(warps * 32), 0, st, x, gamma, beta, out, planes, plane_rows, seg, rows, C, eps);
  return check_launch("sar_layernorm_fwd");
}

int sar_avgpool_fwd(const float* x, float* out, int B, int S, int D, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && out, SAR_ERR_BAD_ARG, "sar_avgpool_fwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0 && D > 0, SAR_ERR_BAD_ARG, "sar_avgpool_fwd: non-positive dimension");
  launch_k(avgpool_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, x, out, S, D);
  return check_launch("sar_avgpool_fwd");
}

int sar_loss_reduce_fwd(const float* sample_stats, const float* ctc_loss, const float* bn_stats,
                        float* out8, int B, void* stream) {
  using namespace sar;
  SAR_REQUIRE(out8, SAR_ERR_BAD_ARG, "sar_loss_reduce_fwd: null out8");
  SAR_REQUIRE(B > 0, SAR_ERR_BAD_ARG, "sar_loss_reduce_fwd: B must be positive");
  launch_k(loss_reduce_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, sample_stats, ctc_loss, bn_stats, out8, B);
  return check_launch("sar_loss_reduce_fwd");
}

}  // extern "C"
