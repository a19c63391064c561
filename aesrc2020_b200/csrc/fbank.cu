// fbank.cu -- on-device feature front-end.
//
// Replaces local/make_fbank.py:24-28 (psf.fbank(y, 16000, nfilt=80)[0]: preemphasis 0.97,
// 400-sample rectangular frames every 160 samples, zero-padded tail, |rfft_512|^2 / 512,
// 80 triangular mel filters, LINEAR energies with zeros replaced by float64 eps) and
// utils.py:35-46 (per-utterance per-bin MinMaxScaler, truncate / zero-pad to T frames).
// The reference runs one python process per utterance for this (local/multi_jobs.sh:24-31).
//
// Kernel 1: one CTA per 16 frames, two frames at a time: preemphasised frame -> shared memory, 512-point real FFT
//           as a 256-point complex radix-2 FFT + untangling pass, power spectrum, SPARSE mel projection.
// Kernel 2: one CTA per utterance: per-bin min / max over ALL frames of the utterance
//           (before truncation, as the reference does), scale to [0,1], write (T,80) padded.
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

// 512-point REAL FFT of a frame as one 256-point complex FFT of z[n] = x[2n] + i x[2n+1] plus an untangling pass
//   X[k] = E[k] + W512^k O[k],  E[k] = (Z[k] + conj Z[256-k]) / 2,  O[k] = (Z[k] - conj Z[256-k]) / 2i,  k = 0..256
// (half the butterflies of the complex 512-point transform, and no cross-talk between frames).  A CTA of 256 threads
// handles FB_FPC consecutive frames of one utterance, two at a time (128 threads per frame), so that its tables are
// built once: the 128 + 257 twiddles (sincospif, once per CTA instead of once per butterfly) and a COMPACT copy of the
// mel filterbank -- the triangles overlap only their neighbours, so a filter touches ~6 of the 257 bins (at most
// FB_MAXW); the dense (257 x 80) projection the first version read from L2 for every frame was 40x the work.
constexpr int FB_FPC = 32;           // frames per CTA
constexpr int FB_MAXNZ = 1024;       // non-zero filterbank weights (2 * 257 for triangular filters; 1024 = any sane bank)

template <typename SampleT>
__global__ void __launch_bounds__(256) fbank_frame_kernel(const SampleT* __restrict__ wav, const long long* __restrict__ offsets,
                                                           const float* __restrict__ melfb_t, float* __restrict__ feat,
                                                           int Fmax) {
  __shared__ float2 tw[128];                    // W256^j
  __shared__ float2 pt[FB_NBIN];                // W512^k
  __shared__ float2 z[2][256];
  __shared__ float pw[2][FB_NBIN + 3];
  __shared__ float mw[FB_MAXNZ];                // compact filter weights
  __shared__ int m_lo[FB_NFILT], m_len[FB_NFILT], m_off[FB_NFILT];
  __shared__ int dense_fallback;
  const int t = threadIdx.x, b = blockIdx.y;
  // ---- tables (constants: before the PDL wait)
  if (t < 128) {
    float sn, cs;
    sincospif(-2.0f * (float)t / 256.0f, &sn, &cs);
    tw[t] = make_float2(cs, sn);
  }
  for (int k = t; k < FB_NBIN; k += 256) {
    float sn, cs;
    sincospif(-(float)k / 256.0f, &sn, &cs);
    pt[k] = make_float2(cs, sn);
  }
  // support of every filter (first / last non-zero bin): the whole CTA scans the (257 x 80) matrix once, coalesced
  __shared__ int s_lo[FB_NFILT], s_hi[FB_NFILT];
  if (t < FB_NFILT) { s_lo[t] = FB_NBIN; s_hi[t] = -1; }
  __syncthreads();
  for (int i = t; i < FB_NBIN * FB_NFILT; i += 256)
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
  if (t == 0) {
    int off = 0;
    for (int j = 0; j < FB_NFILT; ++j) { m_off[j] = off; off += m_len[j]; }
    dense_fallback = off > FB_MAXNZ;
  }
  __syncthreads();
  if (t < FB_NFILT && !dense_fallback)
    for (int i = 0; i < m_len[t]; ++i) mw[m_off[t] + i] = __ldg(melfb_t + (m_lo[t] + i) * FB_NFILT + t);
  pdl_wait();
  pdl_trigger();
  const long long beg = offsets[b], n = offsets[b + 1] - beg;
  const int nf = n > 0 ? fb_num_frames(n) : 0;
  const SampleT* y = wav + beg;
  const int fr = t >> 7, h = t & 127;           // frame of the pair, thread within the frame
  __syncthreads();
  for (int f0 = blockIdx.x * FB_FPC; f0 < (blockIdx.x + 1) * FB_FPC && f0 < nf; f0 += 2) {
    const int f = f0 + fr;
    const bool live = f < nf;
    // ---- preemphasised frame -> z (bit-reversed order for the decimation-in-time butterflies)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = h + 128 * r;                // complex sample index 0..255 = real samples 2i, 2i+1
      float v0 = 0.f, v1 = 0.f;
      if (live) {
        const long long j0 = (long long)f * FB_STEP + 2 * i, j1 = j0 + 1;
        if (2 * i < FB_LEN && j0 < n) v0 = (j0 == 0) ? fb_sample(y, 0) : (fb_sample(y, j0) - FB_PREEMPH * fb_sample(y, j0 - 1));
        if (2 * i + 1 < FB_LEN && j1 < n) v1 = fb_sample(y, j1) - FB_PREEMPH * fb_sample(y, j1 - 1);
      }
      z[fr][__brev((unsigned)i) >> 24] = make_float2(v0, v1);
    }
    __syncthreads();
#pragma unroll
    for (int s = 1; s <= 8; ++s) {
      const int m = 1 << s, half = m >> 1;
      const int g = h / half, j = h - g * half;
      const int i0 = g * m + j, i1 = i0 + half;
      const float2 w = tw[j << (8 - s)];
      const float2 a = z[fr][i0], c = z[fr][i1];
      const float br = c.x * w.x - c.y * w.y, bi = c.x * w.y + c.y * w.x;
      z[fr][i0] = make_float2(a.x + br, a.y + bi);
      z[fr][i1] = make_float2(a.x - br, a.y - bi);
      __syncthreads();
    }
    // ---- untangle -> power spectrum / 512
    for (int k = h; k < FB_NBIN; k += 128) {
      const float2 A = z[fr][k & 255], Bq = z[fr][(256 - k) & 255];
      const float er = 0.5f * (A.x + Bq.x), ei = 0.5f * (A.y - Bq.y);          // E = (A + conj B) / 2
      const float dr = A.x - Bq.x, di = A.y + Bq.y;                            // D = A - conj B
      const float orr = 0.5f * di, oi = -0.5f * dr;                            // O = D / 2i
      const float2 w = pt[k];
      const float xr = er + (orr * w.x - oi * w.y), xi = ei + (orr * w.y + oi * w.x);
      pw[fr][k] = (xr * xr + xi * xi) * (1.0f / FB_NFFT);
    }
    __syncthreads();
    if (h < FB_NFILT && live) {
      float acc = 0.f;
      if (!dense_fallback) {
        const int lo = m_lo[h], len = m_len[h], off = m_off[h];
        for (int i = 0; i < len; ++i) acc = fmaf(pw[fr][lo + i], mw[off + i], acc);
      } else {
        for (int k = 0; k < FB_NBIN; ++k) acc = fmaf(pw[fr][k], __ldg(melfb_t + k * FB_NFILT + h), acc);
      }
      if (acc == 0.f) acc = 2.220446049250313e-16f;
      feat[((size_t)b * Fmax + f) * FB_NFILT + h] = acc;
    }
    // (the next pair's z / pw writes are ordered behind this pair's reads by the barriers of its own stages)
    __syncthreads();
  }
}

// One CTA per utterance, 1000 threads = 50 frame phases x 20 float4 column groups: per-bin min / max over ALL frames of
// the utterance (before truncation, as the reference does), then scale to [0,1] and write (T,80) zero padded.
constexpr int FBN_PH = 50, FBN_C4 = FB_NFILT / 4, FBN_THREADS = FBN_PH * FBN_C4;
__global__ void __launch_bounds__(FBN_THREADS) fbank_norm_kernel(const float* __restrict__ feat, const long long* __restrict__ offsets,
                                                                  float* __restrict__ x_data, int Fmax, int T) {
  pdl_wait();
  pdl_trigger();
  __shared__ float4 smin[FBN_PH][FBN_C4], smax[FBN_PH][FBN_C4];
  const int b = blockIdx.x, c4 = threadIdx.x % FBN_C4, q = threadIdx.x / FBN_C4;
  const long long n = offsets[b + 1] - offsets[b];
  const int nf = n > 0 ? min(fb_num_frames(n), Fmax) : 0;
  const float4* fr = reinterpret_cast<const float4*>(feat + (size_t)b * Fmax * FB_NFILT);
  float4 mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY), mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int f = q; f < nf; f += FBN_PH) {
    const float4 v = fr[(size_t)f * FBN_C4 + c4];
    mn.x = fminf(mn.x, v.x); mn.y = fminf(mn.y, v.y); mn.z = fminf(mn.z, v.z); mn.w = fminf(mn.w, v.w);
    mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
  }
  smin[q][c4] = mn; smax[q][c4] = mx;
  __syncthreads();
  for (int i = 0; i < FBN_PH; ++i) {
    const float4 a = smin[i][c4], c = smax[i][c4];
    mn.x = fminf(mn.x, a.x); mn.y = fminf(mn.y, a.y); mn.z = fminf(mn.z, a.z); mn.w = fminf(mn.w, a.w);
    mx.x = fmaxf(mx.x, c.x); mx.y = fmaxf(mx.y, c.y); mx.z = fmaxf(mx.z, c.z); mx.w = fmaxf(mx.w, c.w);
  }
  // sklearn MinMaxScaler: zero range -> scale 1 -> column of zeros
  auto sc = [](float lo, float hi) { const float r = hi - lo; return 1.f / ((r > 0.f) ? r : 1.f); };
  const float4 scale = make_float4(sc(mn.x, mx.x), sc(mn.y, mx.y), sc(mn.z, mx.z), sc(mn.w, mx.w));
  const float4 off = make_float4(-mn.x * scale.x, -mn.y * scale.y, -mn.z * scale.z, -mn.w * scale.w);
  float4* xo = reinterpret_cast<float4*>(x_data + (size_t)b * T * FB_NFILT);
  for (int f = q; f < T; f += FBN_PH) {
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
  SAR_REQUIRE(B <= 65535, SAR_ERR_UNSUPPORTED, "sar_fbank_fwd: B > 65535");
  SAR_REQUIRE(aligned16(feat_ws) && aligned16(x_data), SAR_ERR_ALIGN, "sar_fbank_fwd: feat_ws / x_data must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(fbank_frame_kernel<SampleT>, dim3(dim3((Fmax + FB_FPC - 1) / FB_FPC, B)), dim3(256), 0, st, wav, offsets, melfb_t, feat_ws, Fmax);
  int rc = check_launch("sar_fbank_fwd(frames)");
  if (rc) return rc;
  launch_k(fbank_norm_kernel, dim3(B), dim3(FBN_THREADS), 0, st, feat_ws, offsets, x_data, Fmax, T);
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
