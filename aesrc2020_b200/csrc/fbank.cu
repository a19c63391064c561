// fbank.cu -- on-device feature front-end.
//
// Replaces local/make_fbank.py:24-28 (psf.fbank(y, 16000, nfilt=80)[0]: preemphasis 0.97,
// 400-sample rectangular frames every 160 samples, zero-padded tail, |rfft_512|^2 / 512,
// 80 triangular mel filters, LINEAR energies with zeros replaced by float64 eps) and
// utils.py:35-46 (per-utterance per-bin MinMaxScaler, truncate / zero-pad to T frames).
// The reference runs one python process per utterance for this (local/multi_jobs.sh:24-31).
//
// Kernel 1: one WARP per frame (persistent grid): preemphasised frame -> registers, 512-point real FFT as a 256-point
//           complex FFT in three register passes (radix 8, 8, 4) + untangling pass, power spectrum, SPARSE mel projection.
// Kernel 2: a cluster of 4 CTAs per utterance: per-bin min / max over ALL frames of the utterance
//           (before truncation, as the reference does), scale to [0,1], write (T,80) padded.
#include <cooperative_groups.h>
#include "common.cuh"

namespace sar {

constexpr int FB_NFFT = 512, FB_LEN = 400, FB_STEP = 160, FB_NFILT = 80, FB_NBIN = 257;
constexpr float FB_PREEMPH = 0.97f;

__device__ __forceinline__ int fb_num_frames(long long n) {
  if (n <= FB_LEN) return 1;
  return 1 + (int)((n - FB_LEN + FB_STEP - 1) / FB_STEP);
}

// sample -> float in [-1, 1): float waveforms as they are, 16-bit PCM divided by 32768 (what soundfile.read hands
// psf.fbank in make_fbank.py:26-27)
__device__ __forceinline__ float fb_sample(const float* y, long long j) { return y[j]; }
__device__ __forceinline__ float fb_sample(const short* y, long long j) { return (float)y[j] * (1.0f / 32768.0f); }
// two consecutive samples with one load when the address allows it
__device__ __forceinline__ bool fb_pair_aligned(const float* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
__device__ __forceinline__ bool fb_pair_aligned(const short* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }
__device__ __forceinline__ void fb_sample2(const float* p, float& a, float& b) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p)); a = v.x; b = v.y;
}
__device__ __forceinline__ void fb_sample2(const short* p, float& a, float& b) {
  const short2 v = __ldg(reinterpret_cast<const short2*>(p));
  a = (float)v.x * (1.0f / 32768.0f); b = (float)v.y * (1.0f / 32768.0f);
}

// 512-point REAL FFT of a frame as one 256-point complex FFT of z[n] = x[2n] + i x[2n+1] plus an untangling pass
//   X[k] = E[k] + W512^k O[k],  E[k] = (Z[k] + conj Z[256-k]) / 2,  O[k] = (Z[k] - conj Z[256-k]) / 2i,  k = 0..256
// (half the butterflies of the complex 512-point transform).
//
// ONE WARP PER FRAME, no CTA-wide barrier in the frame loop.  The 256-point transform is three register passes
// (radix 8, 8, 4) with two transposes through a per-warp shared-memory buffer:
//   n = 32a + 4b + c,  k = ka + 8kb + 64kc
//   X[k] = sum_c W4^{c kc} W256^{c(8kb+ka)} sum_b W8^{b kb} W64^{b ka} sum_a W8^{a ka} z[32a + 4b + c]
//   pass 1: lane = 4b + c     8-point DFT over a (loads z[32a + lane]: coalesced), twiddle W64^{b ka}
//   pass 2: lane = 8c + ka    8-point DFT over b, twiddle W256^{c(8kb+ka)}
//   pass 3: lane l, m = l, l + 32 (m = ka + 8kb): 4-point DFT over c -> Z[m + 64kc]
// so lane l ends up with Z[l + 32j], j = 0..7; Z[256 - k] then sits in lane (32 - l) & 31 (one shuffle per value).
// The first version (a CTA of 256 threads per pair of frames, radix-2 stages in shared memory with a CTA barrier per
// stage, 1024 CTAs each rebuilding the tables) took 164 us for 64 x 500 frames; see profiles/r2_fbank.md.
// Tables per CTA (built once, the grid is persistent: 2 CTAs per SM): W512^k for k < 512 and a COMPACT copy of the mel
// filterbank -- the triangles overlap only their neighbours, so a filter touches ~6 of the 257 bins.
constexpr int FB_WPC = 16;           // warps (frames in flight) per CTA
constexpr int FB_THREADS = FB_WPC * 32;
constexpr int FB_MAXNZ = 1024;       // non-zero filterbank weights (2 * 257 for triangular filters; 1024 = any sane bank)
constexpr int FB_T1S = 34;           // float2 stride between the ka rows of the first transpose (bank-conflict free)
constexpr int FB_T2S = 72;           // float2 stride between the c planes of the second transpose
constexpr int FB_WBUF = 4 * FB_T2S;  // float2 per warp: max(8 * FB_T1S, 4 * FB_T2S, 257 floats)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }          // a * (-i)

// forward 4- and 8-point DFTs (e^{-2 pi i nk/N}), natural order in and out
__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_mi(csub(v1, v3));
  v0 = cadd(a0, a2); v1 = cadd(a1, a3); v2 = csub(a0, a2); v3 = csub(a1, a3);
}
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  const float r = 0.70710678118654752f;
  const float2 w1 = make_float2((o1.x + o1.y) * r, (o1.y - o1.x) * r);        // o1 * (1 - i) / sqrt 2
  const float2 w2 = mul_mi(o2);
  const float2 w3 = make_float2((o3.y - o3.x) * r, -(o3.x + o3.y) * r);       // o3 * (-1 - i) / sqrt 2
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, w1); v[5] = csub(e1, w1);
  v[2] = cadd(e2, w2); v[6] = csub(e2, w2);
  v[3] = cadd(e3, w3); v[7] = csub(e3, w3);
}

template <typename SampleT>
__global__ void __launch_bounds__(FB_THREADS, 2) fbank_frame_kernel(const SampleT* __restrict__ wav, const long long* __restrict__ offsets,
                                                                    const float* __restrict__ melfb_t, float* __restrict__ feat,
                                                                    int B, int Fmax) {
  __shared__ float2 w512[512];                  // W512^k
  __shared__ float2 wbuf[FB_WPC][FB_WBUF];      // per-warp transposes, then the power spectrum
  __shared__ float mw[FB_MAXNZ];                // compact filter weights
  __shared__ int m_lo[FB_NFILT], m_len[FB_NFILT], m_off[FB_NFILT];
  __shared__ int s_lo[FB_NFILT], s_hi[FB_NFILT];
  __shared__ int dense_fallback;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // ---- tables (constants: before the PDL wait)
  {
    float sn, cs;
    sincospif(-(float)t / 256.0f, &sn, &cs);
    w512[t] = make_float2(cs, sn);
  }
  // support of every filter (first / last non-zero bin): the whole CTA scans the (257 x 80) matrix once, coalesced
  if (t < FB_NFILT) { s_lo[t] = FB_NBIN; s_hi[t] = -1; }
  __syncthreads();
  for (int i = t; i < FB_NBIN * FB_NFILT; i += FB_THREADS)
    if (__ldg(melfb_t + i) != 0.f) {
      const int k = i / FB_NFILT, j = i - k * FB_NFILT;
      atomicMin(&s_lo[j], k);
      atomicMax(&s_hi[j], k);
    }
  __syncthreads();
  if (t < FB_NFILT) {
    m_lo[t] = s_hi[t] >= 0 ? s_lo[t] : 0;
    m_len[t] = s_hi[t] >= 0 ? s_hi[t] - s_lo[t] + 1 : 0;
  }
  __syncthreads();
  if (warp == 0) {                              // exclusive prefix sum of the 80 lengths by one warp
    int run = 0;
    for (int j0 = 0; j0 < FB_NFILT; j0 += 32) {
      const int j = j0 + lane;
      const int len = j < FB_NFILT ? m_len[j] : 0;
      int inc = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
      }
      if (j < FB_NFILT) m_off[j] = run + inc - len;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) dense_fallback = run > FB_MAXNZ;
  }
  __syncthreads();
  if (!dense_fallback)                          // (filter, i-th bin of its support) pairs spread over the CTA
    for (int j = warp; j < FB_NFILT; j += FB_WPC)
      for (int i = lane; i < m_len[j]; i += 32) mw[m_off[j] + i] = __ldg(melfb_t + (m_lo[j] + i) * FB_NFILT + j);
  pdl_wait();
  pdl_trigger();
  __syncthreads();
  float2* buf = wbuf[warp];
  float* pw = reinterpret_cast<float*>(buf);
  const int total = B * Fmax;                   // (< 2^30, checked on the host)
  for (int idx = blockIdx.x * FB_WPC + warp; idx < total; idx += gridDim.x * FB_WPC) {
    const int b = idx / Fmax, f = idx - b * Fmax;
    const long long beg = __ldg(offsets + b), n = __ldg(offsets + b + 1) - beg;
    const int nf = n > 0 ? fb_num_frames(n) : 0;
    if (f >= nf) continue;                      // (warp-uniform)
    const SampleT* y = wav + beg;
    // ---- pass 1: preemphasised samples z[32a + lane] = (p[2n], p[2n+1]), p[j] = y[j] - 0.97 y[j-1], p[0] = y[0];
    //      zero beyond the 400-sample frame and beyond the end of the utterance
    float2 v[8];
    {
      const long long j00 = (long long)f * FB_STEP;
      const SampleT* yf = y + j00;                                    // first sample of the frame
      const long long left = n - j00;                                 // samples from the frame start to the end of the utterance
      const int rem = left < FB_LEN ? (int)left : FB_LEN;             // ... inside the frame (>= 1: the frame exists)
      float carry = (j00 > 0) ? fb_sample(y, j00 - 1) : 0.f;        // y[j - 1] of lane 0
      const bool pairs = fb_pair_aligned(yf);                         // (warp-uniform) my two samples are one aligned load
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int i2 = 2 * (32 * a + lane);                           // first real sample of my complex point, in the frame
        float2 z = make_float2(0.f, 0.f);
        if (a < 7) {                                                  // (a == 7: i2 >= 448 > 400 for every lane)
          float s0 = 0.f, s1 = 0.f;
          if (pairs && i2 + 1 < rem) fb_sample2(yf + i2, s0, s1);
          else {
            if (i2 < rem) s0 = fb_sample(yf, i2);
            if (i2 + 1 < rem) s1 = fb_sample(yf, i2 + 1);
          }
          float sm = __shfl_up_sync(0xffffffffu, s1, 1);
          if (lane == 0) sm = carry;
          carry = __shfl_sync(0xffffffffu, s1, 31);
          if (i2 < rem) z.x = (j00 + i2 == 0) ? s0 : fmaf(-FB_PREEMPH, sm, s0);
          if (i2 + 1 < rem) z.y = fmaf(-FB_PREEMPH, s0, s1);
        }
        v[a] = z;
      }
    }
    dft8(v);
    {
      const int bq = lane >> 2;                                       // twiddle W64^{b ka} = W512^{8 b ka}
#pragma unroll
      for (int ka = 0; ka < 8; ++ka) {
        const float2 r = ka ? cmul(v[ka], w512[8 * bq * ka]) : v[0];
        buf[ka * FB_T1S + lane] = r;
      }
    }
    __syncwarp();
    // ---- pass 2: lane = 8c + ka
    {
      const int c = lane >> 3, ka = lane & 7;
#pragma unroll
      for (int bq = 0; bq < 8; ++bq) v[bq] = buf[ka * FB_T1S + 4 * bq + c];
      __syncwarp();                                                   // every lane has its inputs: the buffer may be rewritten
      dft8(v);
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        const float2 r = cmul(v[kb], w512[2 * c * (8 * kb + ka)]);    // W256^{c(8kb+ka)}
        buf[c * FB_T2S + ka + 8 * kb] = r;
      }
    }
    __syncwarp();
    // ---- pass 3: lane l, m = l + 32h: 4-point DFT over c -> Z[l + 32(h + 2kc)]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      float2 t0 = buf[m], t1 = buf[FB_T2S + m], t2 = buf[2 * FB_T2S + m], t3 = buf[3 * FB_T2S + m];
      dft4(t0, t1, t2, t3);
      v[h] = t0; v[h + 2] = t1; v[h + 4] = t2; v[h + 6] = t3;
    }
    __syncwarp();                                                     // transposes read: the buffer becomes the power spectrum
    // ---- untangle -> power spectrum / 512.  My bins are k = lane + 32j; Z[256 - k] is register 7 - j of lane 32 - lane
    //      (lane 0: register (8 - j) & 7 of itself)
    const int src = (32 - lane) & 31;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 mine = (lane == 0) ? v[(8 - j) & 7] : v[7 - j];
      const float2 q = make_float2(__shfl_sync(0xffffffffu, mine.x, src), __shfl_sync(0xffffffffu, mine.y, src));   // Z[256 - k]
      const float2 A = v[j];
      const float er = 0.5f * (A.x + q.x), ei = 0.5f * (A.y - q.y);            // E = (A + conj q) / 2
      const float dr = A.x - q.x, di = A.y + q.y;                              // D = A - conj q
      const float orr = 0.5f * di, oi = -0.5f * dr;                            // O = D / 2i
      const float2 w = w512[lane + 32 * j];
      const float xr = er + (orr * w.x - oi * w.y), xi = ei + (orr * w.y + oi * w.x);
      pw[lane + 32 * j] = (xr * xr + xi * xi) * (1.0f / FB_NFFT);
      if (j == 0 && lane == 0) {                                               // Nyquist bin: X[256] = Re Z[0] - Im Z[0]
        const float xn = A.x - A.y;
        pw[256] = xn * xn * (1.0f / FB_NFFT);
      }
    }
    __syncwarp();
    // ---- sparse mel projection: filters lane, lane + 32, lane + 64
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int m = lane + 32 * g;
      if (m < FB_NFILT) {
        float acc = 0.f;
        if (!dense_fallback) {
          const int lo = m_lo[m], len = m_len[m], off = m_off[m];
          for (int i = 0; i < len; ++i) acc = fmaf(pw[lo + i], mw[off + i], acc);
        } else {
          for (int k = 0; k < FB_NBIN; ++k) acc = fmaf(pw[k], __ldg(melfb_t + k * FB_NFILT + m), acc);
        }
        if (acc == 0.f) acc = 2.220446049250313e-16f;
        feat[((size_t)b * Fmax + f) * FB_NFILT + m] = acc;
      }
    }
    __syncwarp();                                                     // the next frame's pass 1 rewrites the buffer
  }
}

// A cluster of FBN_CL CTAs per utterance, 1000 threads each = 50 frame phases x 20 float4 column groups: per-bin
// min / max over ALL frames of the utterance (before truncation, as the reference does) -- every CTA reduces its
// quarter of the frames, the partial results are exchanged through distributed shared memory -- then scale to [0,1] and
// write (T,80) zero padded.  (One CTA per utterance left 84 of the 148 SMs idle at B = 64: 17 us.)
constexpr int FBN_PH = 50, FBN_C4 = FB_NFILT / 4, FBN_THREADS = FBN_PH * FBN_C4, FBN_CL = 4;
__device__ __forceinline__ float4 f4min(float4 a, float4 b) { return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w)); }
__device__ __forceinline__ float4 f4max(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }
__global__ void __cluster_dims__(FBN_CL, 1, 1) __launch_bounds__(FBN_THREADS)
fbank_norm_kernel(const float* __restrict__ feat, const long long* __restrict__ offsets, float* __restrict__ x_data, int Fmax, int T) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  pdl_wait();
  pdl_trigger();
  __shared__ float4 smin[FBN_PH][FBN_C4], smax[FBN_PH][FBN_C4];
  __shared__ float4 fin[2][FBN_C4];             // this CTA's min | max, read by the whole cluster
  const int b = blockIdx.x / FBN_CL, rank = (int)cluster.block_rank();
  const int c4 = threadIdx.x % FBN_C4, q = threadIdx.x / FBN_C4;
  const long long n = offsets[b + 1] - offsets[b];
  const int nf = n > 0 ? min(fb_num_frames(n), Fmax) : 0;
  const float4* fr = reinterpret_cast<const float4*>(feat + (size_t)b * Fmax * FB_NFILT);
  float4 mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY), mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int f = rank * FBN_PH + q; f < nf; f += FBN_PH * FBN_CL) {
    const float4 v = fr[(size_t)f * FBN_C4 + c4];
    mn = f4min(mn, v); mx = f4max(mx, v);
  }
  smin[q][c4] = mn; smax[q][c4] = mx;
  __syncthreads();
  if (q < 5) {                                  // 50 phases -> 5
    for (int i = q + 5; i < FBN_PH; i += 5) { mn = f4min(mn, smin[i][c4]); mx = f4max(mx, smax[i][c4]); }
    smin[q][c4] = mn; smax[q][c4] = mx;
  }
  __syncthreads();
  if (q == 0) {
    for (int i = 1; i < 5; ++i) { mn = f4min(mn, smin[i][c4]); mx = f4max(mx, smax[i][c4]); }
    fin[0][c4] = mn; fin[1][c4] = mx;
  }
  cluster.sync();
  mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY); mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int r = 0; r < FBN_CL; ++r) {
    const float4* rf = cluster.map_shared_rank(&fin[0][0], r);
    mn = f4min(mn, rf[c4]); mx = f4max(mx, rf[FBN_C4 + c4]);
  }
  cluster.sync();                               // nobody leaves while its shared memory is still being read
  // sklearn MinMaxScaler: zero range -> scale 1 -> column of zeros
  auto sc = [](float lo, float hi) { const float r = hi - lo; return 1.f / ((r > 0.f) ? r : 1.f); };
  const float4 scale = make_float4(sc(mn.x, mx.x), sc(mn.y, mx.y), sc(mn.z, mx.z), sc(mn.w, mx.w));
  const float4 off = make_float4(-mn.x * scale.x, -mn.y * scale.y, -mn.z * scale.z, -mn.w * scale.w);
  float4* xo = reinterpret_cast<float4*>(x_data + (size_t)b * T * FB_NFILT);
  for (int f = rank * FBN_PH + q; f < T; f += FBN_PH * FBN_CL) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < nf) {
      const float4 w = fr[(size_t)f * FBN_C4 + c4];
      v = make_float4(fmaf(w.x, scale.x, off.x), fmaf(w.y, scale.y, off.y), fmaf(w.z, scale.z, off.z), fmaf(w.w, scale.w, off.w));
    }
    xo[(size_t)f * FBN_C4 + c4] = v;
  }
}

}  // namespace sar

namespace sar {
template <typename SampleT>
static int fbank_launch(const SampleT* wav, const long long* offsets, const float* melfb_t, float* feat_ws, float* x_data,
                        int B, int Fmax, int T, void* stream) {
  SAR_REQUIRE(wav && offsets && melfb_t && feat_ws && x_data, SAR_ERR_BAD_ARG, "sar_fbank_fwd: null pointer");
  SAR_REQUIRE(B > 0 && Fmax > 0 && T > 0, SAR_ERR_BAD_ARG, "sar_fbank_fwd: non-positive dimension");
  SAR_REQUIRE((long long)B * Fmax < (1ll << 30), SAR_ERR_UNSUPPORTED, "sar_fbank_fwd: B * Fmax must stay below 2^30 frames");
  SAR_REQUIRE(aligned16(feat_ws) && aligned16(x_data), SAR_ERR_ALIGN, "sar_fbank_fwd: feat_ws / x_data must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long frames = (long long)B * Fmax;
  long long grid = (frames + FB_WPC - 1) / FB_WPC;
  if (grid > 2LL * sms) grid = 2LL * sms;                 // persistent: every warp walks frames warp-id, warp-id + #warps, ...
  launch_k(fbank_frame_kernel<SampleT>, dim3((unsigned)grid), dim3(FB_THREADS), 0, st, wav, offsets, melfb_t, feat_ws, B, Fmax);
  int rc = check_launch("sar_fbank_fwd(frames)");
  if (rc) return rc;
  launch_k(fbank_norm_kernel, dim3(B * FBN_CL), dim3(FBN_THREADS), 0, st, feat_ws, offsets, x_data, Fmax, T);
  return check_launch("sar_fbank_fwd(norm)");
}
}  // namespace sar

extern "C" int sar_fbank_fwd(const float* wav, const long long* offsets, const float* melfb_t,
                             float* feat_ws, float* x_data, int B, int Fmax, int T, void* stream) {
  return sar::fbank_launch<float>(wav, offsets, melfb_t, feat_ws, x_data, B, Fmax, T, stream);
}

extern "C" int sar_fbank_pcm16_fwd(const int16_t* pcm, const long long* offsets, const float* melfb_t,
                                   float* feat_ws, float* x_data, int B, int Fmax, int T, void* stream) {
  return sar::fbank_launch<short>(reinterpret_cast<const short*>(pcm), offsets, melfb_t, feat_ws, x_data, B, Fmax, T, stream);
}
