// stem_tc.cu -- the ResNet stem on tcgen05 tensor cores, one persistent kernel:
//   Conv2D 7x7/s2 'same' (+bias) -> BatchNormalization -> ReLU -> MaxPooling2D 3x3/s2 'same'
//   (resnet.py:28-45 via :173/:191, and :174/:192), writing the pooled map straight into the
//   flat-pad hi/lo planes the residual-block kernel consumes.
//
// Cin = 1, so the im2col matrix is built on chip: with the 7 kernel columns padded to 8,
//   A[pos, kr*8 + kc] = x[2*hc + kr - pt, 2*wc + kc - pl]
// and the 8 kc-values of one (pos, kr) are 8 CONSECUTIVE input samples -> one 16-byte chunk of the
// K-major, 128B-swizzled A row (K = 64 = one swizzle atom).  Four builder warps convert those
// chunks to fp16 hi/lo (x = hi + lo/2048) from a shared-memory copy of the item's 23 input rows;
// one elected thread issues  [acc0|acc1] (+)= A_hi x [W_hi;W_lo]  and  acc1 += A_lo x W_hi  per
// k16 step (the 2^-22 A_lo x W_lo term is dropped, as in conv_tc.cu); four epilogue warps drain
// TMEM: acc0 + acc1/2048 -> folded bias/BN -> ReLU -> a swizzled fp32 conv-row buffer in shared
// memory, then max-pool 3x3/s2 from that buffer and store hi/lo planes.  The (T/2)x40x64 conv map
// never leaves the SM.
//
// Work item = (utterance, 4 pooled rows) = 9 conv rows = 3 MMA tiles of 3 conv rows x 40 = 120
// positions (M = 128, 8 idle rows); CTAs are persistent, items are dealt round-robin.
// HBM per item: 23 x 80 fp32 in (7.4 KB), 4 x 20 x 64 x 2 x 2 B out (20 KB).
#include "tc_common.cuh"

namespace sar {

constexpr int ST_WC = 40;                 // conv width  (D = 80)
constexpr int ST_WP = 20;                 // pooled width
constexpr int ST_F0 = 64;
constexpr int ST_D = 80;
constexpr int ST_PL = 2;                  // TF-SAME left pad of the 7-wide kernel at D = 80, stride 2
constexpr int ST_PH = 4;                  // pooled rows per item
constexpr int ST_CR = 2 * ST_PH + 1;      // conv rows per item (9)
constexpr int ST_IR = 2 * (ST_CR - 1) + 7;   // input rows per item (23)
constexpr int ST_XW = 88;                 // x_s row pitch in floats (2 + 80 + 6 zero columns)
constexpr int ST_TROWS = 3;               // conv rows per MMA tile
constexpr int ST_TILES = ST_CR / ST_TROWS;   // 3
constexpr int ST_MROWS = ST_TROWS * ST_WC;   // 120 live rows of the M = 128 tile
constexpr int ST_THREADS = 288;           // warps 0-3 epilogue (TMEM quadrant = warp), 4-7 builders, 8 MMA
constexpr int ST_APLANE = 16384;          // 128 rows x 128 B
constexpr int ST_XBUF = 8192;             // >= 23 * 88 * 4
constexpr int ST_CS_BYTES = ST_CR * ST_WC * 256;

struct StemTcP {
  const float* x; const float* w; const float* bias; const float* scale; const float* shift;
  __half* planes;
  int B, T, Hc, pt, Hp, ppt, ppl, bands, n_items;
};

__global__ void __launch_bounds__(ST_THREADS, 1) stem_tc_kernel(const StemTcP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;                                   // [2 stages][hi, lo][16 KB]
  uint8_t* w_base = a_base + 4 * ST_APLANE;                 // [W_hi rows 0..63 ; W_lo rows 64..127] x 128 B
  uint8_t* c_base = w_base + ST_APLANE;                     // conv rows: [360 positions][64 fp32], 16 B chunks XOR (pos & 15)
  float* x_s = reinterpret_cast<float*>(c_base + ST_CS_BYTES);          // [2][23][88]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(x_s) + 2 * ST_XBUF);
  uint64_t* empty_bar = full_bar + 2;
  uint64_t* tfull_bar = empty_bar + 2;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_sc = reinterpret_cast<float*>(tmem_slot + 4);    // [64] BN scale
  float* s_sh = s_sc + ST_F0;                               // [64] bias*scale + shift

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1);
      mbar_init(&tfull_bar[s], 1);  mbar_init(&tempty_bar[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  if (threadIdx.x < ST_F0) {
    const float sc = __ldg(p.scale + threadIdx.x);
    s_sc[threadIdx.x] = sc;
    s_sh[threadIdx.x] = fmaf(__ldg(p.bias + threadIdx.x), sc, __ldg(p.shift + threadIdx.x));
  }
  // zero the operand buffers once: W's padded k slots, A's 8th chunk and idle rows, x_s's pad columns
  for (int i = threadIdx.x; i < (5 * ST_APLANE) / 16; i += ST_THREADS) reinterpret_cast<uint4*>(a_base)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < (2 * ST_XBUF) / 16; i += ST_THREADS) reinterpret_cast<uint4*>(x_s)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int Rimg = (p.Hp + 1) * (ST_WP + 1);
  const long long R = (long long)p.B * Rimg;

  if (warp >= 4 && warp < 8) {
    // ===================== builders: weights once, then x rows -> A tiles =====================
    const int b = threadIdx.x - 128;
    for (int i = b; i < 49 * ST_F0; i += 128) {
      const int f = i & 63, k = i >> 6;                     // HWIO (7,7,1,64): i = (kr*7 + kc)*64 + f
      const int kr = k / 7, kc = k - kr * 7;
      const float wv = __ldg(p.w + i);
      const __half h = __float2half_rn(wv);
      const __half l = __float2half_rn((wv - __half2float(h)) * 2048.f);
      const uint32_t off = (uint32_t)f * 128u + (uint32_t)((kr ^ (f & 7)) << 4) + (uint32_t)kc * 2u;
      *reinterpret_cast<__half*>(w_base + off) = h;
      *reinterpret_cast<__half*>(w_base + 8192 + off) = l;
    }
    float4 xr[4];
    auto load_x = [&](int item) {
      const int n = item / p.bands, band = item - n * p.bands;
      const int hi0 = 2 * (2 * band * ST_PH - p.ppt) - p.pt;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = b + 128 * j;
        const int r = idx / 20, c4 = idx - r * 20;
        const int hi = hi0 + r;
        xr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < ST_IR * 20 && hi >= 0 && hi < p.T)
          xr[j] = __ldg(reinterpret_cast<const float4*>(p.x + ((size_t)n * p.T + hi) * ST_D) + c4);
      }
    };
    auto store_x = [&](float* xb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = b + 128 * j;
        const int r = idx / 20, c4 = idx - r * 20;
        if (idx < ST_IR * 20) {
          float2* d = reinterpret_cast<float2*>(xb + r * ST_XW + ST_PL + 4 * c4);
          d[0] = make_float2(xr[j].x, xr[j].y);
          d[1] = make_float2(xr[j].z, xr[j].w);
        }
      }
    };
    const int m = b;
    const int mr = m / ST_WC, mw = m - mr * ST_WC;
    int g = 0, k = 0;
    if ((int)blockIdx.x < p.n_items) load_x(blockIdx.x);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++k) {
      float* xb = x_s + (k & 1) * (ST_XBUF / 4);
      store_x(xb);
      named_bar_sync(1, 128);
      if (item + (int)gridDim.x < p.n_items) load_x(item + gridDim.x);
      for (int tile = 0; tile < ST_TILES; ++tile, ++g) {
        const int s = g & 1;
        mbar_wait(&empty_bar[s], ((uint32_t)(g >> 1) & 1u) ^ 1u);
        if (m < ST_MROWS) {
          uint8_t* arow = a_base + s * 2 * ST_APLANE + m * 128;
          const float* src = xb + (2 * (ST_TROWS * tile + mr)) * ST_XW + 2 * mw;
#pragma unroll
          for (int kr = 0; kr < 7; ++kr) {
            const float2* s2 = reinterpret_cast<const float2*>(src + kr * ST_XW);
            uint32_t hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 v = s2[e];
              const __half2 h2 = __floats2half2_rn(v.x, v.y);
              const float2 hf = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn((v.x - hf.x) * 2048.f, (v.y - hf.y) * 2048.f);
              hh[e] = *reinterpret_cast<const uint32_t*>(&h2);
              ll[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            const uint32_t co = (uint32_t)((kr ^ (m & 7)) << 4);
            *reinterpret_cast<uint4*>(arow + co) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(arow + ST_APLANE + co) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dbh = make_desc(smem_u32(w_base), 128);
    int g = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int tile = 0; tile < ST_TILES; ++tile, ++g) {
        const int s = g & 1;
        const uint32_t ph = (uint32_t)(g >> 1) & 1u;
        mbar_wait(&tempty_bar[s], ph ^ 1u);
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dah = make_desc(smem_u32(a_base + s * 2 * ST_APLANE), 128);
          const uint64_t dal = make_desc(smem_u32(a_base + s * 2 * ST_APLANE + ST_APLANE), 128);
          const uint32_t acc0 = tmem_base + (uint32_t)(s * 128), acc1 = acc0 + 64u;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_f16(acc0, dah + 2u * kk, dbh + 2u * kk, idesc_128, kk ? 1u : 0u);   // [acc0|acc1] (+)= Ah x [Wh;Wl]
            umma_f16(acc1, dal + 2u * kk, dbh + 2u * kk, idesc_64, 1u);              // acc1 += Al x Wh
          }
          umma_commit(&empty_bar[s]);
          umma_commit(&tfull_bar[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps: TMEM -> conv rows in smem -> max-pool -> planes =====================
    const int m = threadIdx.x;                       // TMEM lane
    const int P = ST_WP + 1;
    int g = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int n = item / p.bands, band = item - n * p.bands;
      const int hp0 = band * ST_PH;
      const int hc0 = 2 * hp0 - p.ppt;
      for (int tile = 0; tile < ST_TILES; ++tile, ++g) {
        const int s = g & 1;
        mbar_wait(&tfull_bar[s], (uint32_t)(g >> 1) & 1u);
        tc_fence_after();
        const int hc = hc0 + ST_TROWS * tile + m / ST_WC;
        const bool live = m < ST_MROWS;
        const bool inside = hc >= 0 && hc < p.Hc;
        const int pos = tile * ST_MROWS + m;
        uint8_t* crow = c_base + (size_t)pos * 256;
        const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r0[32], r1[32];
          tmem_ld32(tb + (uint32_t)(32 * half), r0);
          tmem_ld32(tb + 64u + (uint32_t)(32 * half), r1);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int f = 32 * half + 4 * q + e;
                const float acc = fmaf(__uint_as_float(r1[4 * q + e]), 1.f / 2048.f, __uint_as_float(r0[4 * q + e]));
                o[e] = inside ? fmaxf(fmaf(acc, s_sc[f], s_sh[f]), 0.f) : -INFINITY;
              }
              const int j = 8 * half + q;
              *reinterpret_cast<float4*>(crow + ((j ^ (pos & 15)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[s]);
      }
      named_bar_sync(2, 128);
      // ---- 3x3/s2 max-pool of the 9 conv rows: (pooled position, 4-channel chunk) per thread-item
#pragma unroll 2
      for (int i = 0; i < (ST_PH * ST_WP * 16) / 128; ++i) {
        const int idx = m + 128 * i;
        const int j = idx & 15, pp = idx >> 4;
        const int dh = pp / ST_WP, wp = pp - dh * ST_WP;
        const int hp = hp0 + dh;
        if (hp >= p.Hp) continue;
        float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) {
            const int wc = 2 * wp - p.ppl + bb;
            if (wc < 0 || wc >= ST_WC) continue;
            const int pos = (2 * dh + a) * ST_WC + wc;
            const float4 v = *reinterpret_cast<const float4*>(c_base + (size_t)pos * 256 + ((j ^ (pos & 15)) << 4));
            mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
          }
        }
        const __half2 h01 = __floats2half2_rn(mx.x, mx.y), h23 = __floats2half2_rn(mx.z, mx.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((mx.x - f01.x) * 2048.f, (mx.y - f01.y) * 2048.f);
        const __half2 l23 = __floats2half2_rn((mx.z - f23.x) * 2048.f, (mx.w - f23.y) * 2048.f);
        const long long row = (long long)n * Rimg + (long long)hp * P + wp;
        *reinterpret_cast<uint2*>(p.planes + (size_t)row * ST_F0 + 4 * j) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(p.planes + ((size_t)R + row) * ST_F0 + 4 * j) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
      named_bar_sync(2, 128);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Launch for D == 80, F0 == 64 (every res34 stem and the default res18/64); returns SAR_ERR_UNSUPPORTED otherwise.
int stem_tc_launch(const float* x, const float* w, const float* bias, const float* scale, const float* shift,
                   void* planes, int B, int T, int D, int F0, int Hc, int pt, int Wc, int pl, int Hp, int ppt,
                   int Wp, int ppl, cudaStream_t stream) {
  if (D != ST_D || F0 != ST_F0 || Wc != ST_WC || pl != ST_PL || Wp != ST_WP) return SAR_ERR_UNSUPPORTED;
  StemTcP p{};
  p.x = x; p.w = w; p.bias = bias; p.scale = scale; p.shift = shift; p.planes = reinterpret_cast<__half*>(planes);
  p.B = B; p.T = T; p.Hc = Hc; p.pt = pt; p.Hp = Hp; p.ppt = ppt; p.ppl = ppl;
  p.bands = (Hp + ST_PH - 1) / ST_PH;
  p.n_items = B * p.bands;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 1024 + 5 * (size_t)ST_APLANE + ST_CS_BYTES + 2 * ST_XBUF + 128 + 2 * ST_F0 * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("sar_stem_pool_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  const int grid = p.n_items < sms ? p.n_items : sms;
  stem_tc_kernel<<<grid, ST_THREADS, smem, stream>>>(p);
  return check_launch("sar_stem_pool_fwd(tc)");
}

}  // namespace sar
