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

#ifdef SAR_STEM_PROFILE     // build with SAR_NVCC_EXTRA=-DSAR_STEM_PROFILE: CTA 0 prints per-role cycle counts
#define ST_PROF_DECL long long pr_wait = 0, pr_work = 0, pr_pool = 0, pr_t = clock64(), pr_t0 = pr_t;
#define ST_PROF(acc) { const long long _n = clock64(); acc += _n - pr_t; pr_t = _n; }
#define ST_PROF_PRINT(role) if (blockIdx.x == 0 && lane == 0 && (warp & 7) == 0) \
    printf("stem_tc %s: total %lld wait %lld work %lld pool %lld\n", role, clock64() - pr_t0, pr_wait, pr_work, pr_pool);
#else
#define ST_PROF_DECL
#define ST_PROF(acc)
#define ST_PROF_PRINT(role)
#endif

constexpr int ST_WC = 40;                 // conv width  (D = 80)
constexpr int ST_WP = 20;                 // pooled width
constexpr int ST_F0 = 64;
constexpr int ST_D = 80;
constexpr int ST_PL = 2;                  // TF-SAME left pad of the 7-wide kernel at D = 80, stride 2
constexpr int ST_PH = 4;                  // pooled rows per item
constexpr int ST_CR = 2 * ST_PH + 1;      // conv rows per item (9)
constexpr int ST_IR = 2 * (ST_CR - 1) + 7;   // input rows per item (23)
constexpr int ST_XW = 88;                 // x_s row pitch in halfs: 4 zero columns + 80 + 4 zero columns (input col c at c + 4)
constexpr int ST_XPLANE = 23 * 88;        // halfs per hi (or lo) plane of one x buffer
constexpr int ST_TROWS = 3;               // conv rows per MMA tile
constexpr int ST_TILES = ST_CR / ST_TROWS;   // 3
constexpr int ST_MROWS = ST_TROWS * ST_WC;   // 120 live rows of the M = 128 tile
// Warp roles.  The SM's schedulers favour the HIGHEST warp id of a sub-partition, and a warp polling an mbarrier
// still takes issue slots: the epilogue/pool warps (the critical path) get the high ids, the builders (which
// mostly wait for a free A stage) the low ids and a sleeping wait.
constexpr int ST_THREADS = 544;           // warps 0-7 builders, 8-15 epilogue (TMEM quadrant = warp & 3, column half = (warp >> 2) & 1), 16 MMA
constexpr int ST_WARP_MMA = 16;
constexpr int ST_APLANE = 16384;          // 128 rows x 128 B
constexpr int ST_XBUF = 8192;             // >= 23 * 88 * 4
constexpr int ST_CPITCH = 272;              // conv-row buffer: 64 fp32 + 16 B pad per position (conflict-free STS.128 by position / LDS.128 by chunk)
constexpr int ST_CS_BYTES = ST_CR * ST_WC * ST_CPITCH;

struct StemTcP {
  const float* x; const float* w; const float* bias; const float* scale; const float* shift;
  __half* planes;
  int B, T, Hc, pt, Hp, ppt, ppl, bands, n_items;
};

__global__ void __launch_bounds__(ST_THREADS, 1) stem_tc_kernel(const StemTcP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // offset arithmetic on the __shared__ array keeps the address space: LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_base = smem;                                   // [2 stages][hi, lo][16 KB]
  uint8_t* w_base = a_base + 4 * ST_APLANE;                 // [W_hi rows 0..63 ; W_lo rows 64..127] x 128 B
  uint8_t* c_base = w_base + ST_APLANE;                     // conv rows: [360 positions][pitch 272 B]
  __half* x_s = reinterpret_cast<__half*>(c_base + ST_CS_BYTES);        // [2 buffers][hi, lo][23][88] fp16
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
      mbar_init(&full_bar[s], 8); mbar_init(&empty_bar[s], 1);
      mbar_init(&tfull_bar[s], 1);  mbar_init(&tempty_bar[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == ST_WARP_MMA) tmem_alloc(tmem_slot, 256);
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
  pdl_wait();
  pdl_trigger();

  const int Rimg = (p.Hp + 1) * (ST_WP + 1);
  const long long R = (long long)p.B * Rimg;

  if (warp < 8) {
    // ===================== builders: weights once, then x rows -> A tiles =====================
    // 256 threads: thread (m = b & 127, half = b >> 7) converts kernel rows 0..3 / 4..6 of A row m
    const int b = threadIdx.x;
    for (int i = b; i < 49 * ST_F0; i += 256) {
      const int f = i & 63, k = i >> 6;                     // HWIO (7,7,1,64): i = (kr*7 + kc)*64 + f
      const int kr = k / 7, kc = k - kr * 7;
      const float wv = __ldg(p.w + i);
      const __half h = __float2half_rn(wv);
      const __half l = __float2half_rn((wv - __half2float(h)) * 2048.f);
      const uint32_t off = (uint32_t)f * 128u + (uint32_t)((kr ^ (f & 7)) << 4) + (uint32_t)kc * 2u;
      *reinterpret_cast<__half*>(w_base + off) = h;
      *reinterpret_cast<__half*>(w_base + 8192 + off) = l;
    }
    float4 xr[2];
    auto load_x = [&](int item) {
      const int n = item / p.bands, band = item - n * p.bands;
      const int hi0 = 2 * (2 * band * ST_PH - p.ppt) - p.pt;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = b + 256 * j;
        const int r = idx / 20, c4 = idx - r * 20;
        const int hi = hi0 + r;
        xr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < ST_IR * 20 && hi >= 0 && hi < p.T)
          xr[j] = __ldg(reinterpret_cast<const float4*>(p.x + ((size_t)n * p.T + hi) * ST_D) + c4);
      }
    };
    // every input sample is split into fp16 hi/lo ONCE here (a sample is used by up to 28 A elements)
    auto store_x = [&](__half* xb) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = b + 256 * j;
        const int r = idx / 20, c4 = idx - r * 20;
        if (idx < ST_IR * 20) {
          const __half2 h01 = __floats2half2_rn(xr[j].x, xr[j].y), h23 = __floats2half2_rn(xr[j].z, xr[j].w);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn((xr[j].x - f01.x) * 2048.f, (xr[j].y - f01.y) * 2048.f);
          const __half2 l23 = __floats2half2_rn((xr[j].z - f23.x) * 2048.f, (xr[j].w - f23.y) * 2048.f);
          __half* d = xb + r * ST_XW + 4 + 4 * c4;
          *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          *reinterpret_cast<uint2*>(d + ST_XPLANE) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        }
      }
    };
    const int m = b & 127;
    const int kr0 = (b >> 7) ? 4 : 0, kr1 = (b >> 7) ? 7 : 4;
    const int mr = m / ST_WC, mw = m - mr * ST_WC;
    int g = 0, k = 0;
    ST_PROF_DECL
    if ((int)blockIdx.x < p.n_items) load_x(blockIdx.x);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++k) {
      __half* xb = x_s + (k & 1) * (ST_XBUF / 2);
      store_x(xb);
      named_bar_sync(1, 256);
      if (item + (int)gridDim.x < p.n_items) load_x(item + gridDim.x);
      ST_PROF(pr_pool)
      for (int tile = 0; tile < ST_TILES; ++tile, ++g) {
        const int s = g & 1;
        while (!mbar_try_wait(&empty_bar[s], ((uint32_t)(g >> 1) & 1u) ^ 1u)) __nanosleep(64);
        ST_PROF(pr_wait)
        if (m < ST_MROWS) {
          uint8_t* arow = a_base + s * 2 * ST_APLANE + m * 128;
          // chunk (m, kr) = 8 consecutive samples starting at input column 2*mw - 2 = storage column 2*mw + 2
          const __half* src = xb + (2 * (ST_TROWS * tile + mr)) * ST_XW + 2 * mw + 2;
#pragma unroll 4
          for (int kr = kr0; kr < kr1; ++kr) {
            const uint32_t* sh = reinterpret_cast<const uint32_t*>(src + kr * ST_XW);
            const uint32_t* sl = reinterpret_cast<const uint32_t*>(src + kr * ST_XW + ST_XPLANE);
            const uint32_t co = (uint32_t)((kr ^ (m & 7)) << 4);
            *reinterpret_cast<uint4*>(arow + co) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
            *reinterpret_cast<uint4*>(arow + ST_APLANE + co) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        ST_PROF(pr_work)
      }
    }
    ST_PROF_PRINT("builder")
  } else if (warp == ST_WARP_MMA) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dbh = make_desc(smem_u32(w_base), 128);
    int g = 0;
    ST_PROF_DECL
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int tile = 0; tile < ST_TILES; ++tile, ++g) {
        const int s = g & 1;
        const uint32_t ph = (uint32_t)(g >> 1) & 1u;
        mbar_wait(&tempty_bar[s], ph ^ 1u);
        ST_PROF(pr_wait)
        mbar_wait(&full_bar[s], ph);
        ST_PROF(pr_pool)
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
        ST_PROF(pr_work)
      }
    }
    ST_PROF_PRINT("mma (wait=tempty pool=full)")
  } else {
    // ===================== epilogue warps: TMEM -> conv rows in smem -> max-pool -> planes =====================
    const int quad = warp & 3, chalf = (warp >> 2) & 1;
    const int et = threadIdx.x - 256;                // 0..255 within the epilogue group
    const int m = quad * 32 + lane;                  // TMEM lane = A row
    const int P = ST_WP + 1;
    const int pj = et & 15;                 // pool role: 4-channel chunk (constant per thread)
    const float4 sc4 = *reinterpret_cast<const float4*>(s_sc + 4 * pj);
    const float4 sh4 = *reinterpret_cast<const float4*>(s_sh + 4 * pj);
    int g = 0;
    ST_PROF_DECL
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int n = item / p.bands, band = item - n * p.bands;
      const int hp0 = band * ST_PH;
      const int hc0 = 2 * hp0 - p.ppt;
      const int rlo = max(0, -hc0), rhi = min(ST_CR - 1, p.Hc - 1 - hc0);     // conv rows of this item inside the map
      for (int tile = 0; tile < ST_TILES; ++tile, ++g) {
        const int s = g & 1;
        mbar_wait(&tfull_bar[s], (uint32_t)(g >> 1) & 1u);
        ST_PROF(pr_wait)
        tc_fence_after();
        // raw conv sums acc0 + acc1/2048 of my 32 channels -> swizzled conv-row buffer (BN/ReLU happen in the pool)
        const int pos = tile * ST_MROWS + m;
        uint8_t* crow = c_base + pos * ST_CPITCH;
        const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 128 + 32 * chalf);
        uint32_t r0[32], r1[32];
        tmem_ld32(tb, r0);
        tmem_ld32(tb + 64u, r1);
        tmem_ld_wait();
        if (m < ST_MROWS) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = fmaf(__uint_as_float(r1[4 * q + e]), 1.f / 2048.f, __uint_as_float(r0[4 * q + e]));
            const int j = 8 * chalf + q;
            *reinterpret_cast<float4*>(crow + (j << 4)) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[s]);
        ST_PROF(pr_work)
      }
      named_bar_sync(2, 256);
      // ---- BN + ReLU + 3x3/s2 max-pool of the 9 conv rows: (pooled position, 4-channel chunk) per thread-item.
      // relu(max(.)) = max(0, .): start from 0 and skip rows / columns outside the conv map.
#pragma unroll 1
      for (int i = 0; i < (ST_PH * ST_WP * 16) / 256; ++i) {
        const int pp = (et >> 4) + 16 * i;
        const int dh = pp / ST_WP, wp = pp - dh * ST_WP;
        const int hp = hp0 + dh;
        if (hp >= p.Hp) continue;
        // taps outside the conv map are CLAMPED onto a valid tap of the same window (a duplicate never changes
        // a max), so there is no masking: 9 loads at row/column offsets, BN, max with 0 (= ReLU)
        const uint8_t* cb = c_base + (pj << 4);
        int ro[3], co[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) ro[a] = min(max(2 * dh + a, rlo), rhi) * (ST_WC * ST_CPITCH);
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) co[bb] = min(max(2 * wp - p.ppl + bb, 0), ST_WC - 1) * ST_CPITCH;
        float4 v[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) v[3 * a + bb] = *reinterpret_cast<const float4*>(cb + ro[a] + co[bb]);
        float4 mx = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          mx.x = fmax_nan(mx.x, fmaf(v[k].x, sc4.x, sh4.x));
          mx.y = fmax_nan(mx.y, fmaf(v[k].y, sc4.y, sh4.y));
          mx.z = fmax_nan(mx.z, fmaf(v[k].z, sc4.z, sh4.z));
          mx.w = fmax_nan(mx.w, fmaf(v[k].w, sc4.w, sh4.w));
        }
        const __half2 h01 = __floats2half2_rn(mx.x, mx.y), h23 = __floats2half2_rn(mx.z, mx.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((mx.x - f01.x) * 2048.f, (mx.y - f01.y) * 2048.f);
        const __half2 l23 = __floats2half2_rn((mx.z - f23.x) * 2048.f, (mx.w - f23.y) * 2048.f);
        const long long row = (long long)n * Rimg + (long long)hp * P + wp;
        *reinterpret_cast<uint2*>(p.planes + (size_t)row * ST_F0 + 4 * pj) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(p.planes + ((size_t)R + row) * ST_F0 + 4 * pj) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
      named_bar_sync(2, 256);
      ST_PROF(pr_pool)
    }
    ST_PROF_PRINT("epilogue")
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ST_WARP_MMA) {
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
  { const int arc = allow_max_smem(stem_tc_kernel, "sar_stem_pool_fwd"); if (arc) return arc; }
  const int grid = p.n_items < sms ? p.n_items : sms;
  launch_k(stem_tc_kernel, dim3(grid), dim3(ST_THREADS), smem, stream, p);
  return check_launch("sar_stem_pool_fwd(tc)");
}

}  // namespace sar
