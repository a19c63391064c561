// fbank.cu -- on-device feature front-end.
//
// Replaces local/make_fbank.py:24-28 (psf.fbank(y, 16000, nfilt=80)[0]: preemphasis 0.97,
// 400-sample rectangular frames every 160 samples, zero-padded tail, |rfft_512|^2 / 512,
// 80 triangular mel filters, LINEAR energies with zeros replaced by float64 eps) and
// utils.py:35-46 (per-utterance per-bin MinMaxScaler, truncate / zero-pad to T frames).
// The reference runs one python process per utterance for this (local/multi_jobs.sh:24-31).
//
// Kernel 1: one CTA per frame: preemphasised frame -> shared memory, 512-point radix-2 FFT
//           (9 butterfly stages, 256 threads), power spectrum, mel projection (bin-major
//           filterbank so the 80 filter threads read coalesced rows).
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

__global__ void __launch_bounds__(256) fbank_frame_kernel(const float* __restrict__ wav, const long long* __restrict__ offsets,
                                                           const float* __restrict__ melfb_t, float* __restrict__ feat,
                                                           int Fmax) {
  pdl_wait();
  pdl_trigger();
  __shared__ float re[FB_NFFT], im[FB_NFFT];
  __shared__ float pw[FB_NBIN + 3];
  const int f = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const long long beg = offsets[b], n = offsets[b + 1] - beg;
  if (n <= 0 || f >= fb_num_frames(n)) return;
  const float* y = wav + beg;
  for (int i = t; i < FB_NFFT; i += 256) {
    float v = 0.f;
    long long j = (long long)f * FB_STEP + i;
    if (i < FB_LEN && j < n) v = (j == 0) ? y[0] : (y[j] - FB_PREEMPH * y[j - 1]);
    int r = __brev((unsigned)i) >> (32 - 9);
    re[r] = v;
    im[r] = 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int s = 1; s <= 9; ++s) {
    const int m = 1 << s, half = m >> 1;
    const int g = t / half, j = t - g * half;
    const int i0 = g * m + j, i1 = i0 + half;
    float sn, cs;
    sincospif(-2.0f * (float)j / (float)m, &sn, &cs);
    float br = re[i1] * cs - im[i1] * sn;
    float bi = re[i1] * sn + im[i1] * cs;
    float ar = re[i0], ai = im[i0];
    re[i0] = ar + br; im[i0] = ai + bi;
    re[i1] = ar - br; im[i1] = ai - bi;
    __syncthreads();
  }
  for (int k = t; k < FB_NBIN; k += 256) pw[k] = (re[k] * re[k] + im[k] * im[k]) * (1.0f / FB_NFFT);
  __syncthreads();
  if (t < FB_NFILT) {
    float acc = 0.f;
    for (int k = 0; k < FB_NBIN; ++k) acc = fmaf(pw[k], __ldg(melfb_t + k * FB_NFILT + t), acc);
    if (acc == 0.f) acc = 2.220446049250313e-16f;
    feat[((size_t)b * Fmax + f) * FB_NFILT + t] = acc;
  }
}

__global__ void __launch_bounds__(320) fbank_norm_kernel(const float* __restrict__ feat, const long long* __restrict__ offsets,
                                                          float* __restrict__ x_data, int Fmax, int T) {
  pdl_wait();
  pdl_trigger();
  __shared__ float smin[4][FB_NFILT], smax[4][FB_NFILT];
  const int b = blockIdx.x, m = threadIdx.x % FB_NFILT, q = threadIdx.x / FB_NFILT;   // 4 frame-phases
  const long long n = offsets[b + 1] - offsets[b];
  const int nf = n > 0 ? min(fb_num_frames(n), Fmax) : 0;
  const float* fr = feat + (size_t)b * Fmax * FB_NFILT;
  float mn = INFINITY, mx = -INFINITY;
  for (int f = q; f < nf; f += 4) {
    float v = fr[(size_t)f * FB_NFILT + m];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  smin[q][m] = mn; smax[q][m] = mx;
  __syncthreads();
  mn = fminf(fminf(smin[0][m], smin[1][m]), fminf(smin[2][m], smin[3][m]));
  mx = fmaxf(fmaxf(smax[0][m], smax[1][m]), fmaxf(smax[2][m], smax[3][m]));
  float rng = mx - mn;
  if (!(rng > 0.f)) rng = 1.f;                 // sklearn: zero range -> scale 1 -> column of zeros
  const float scale = 1.f / rng;
  const float off = -mn * scale;
  float* xo = x_data + (size_t)b * T * FB_NFILT;
  for (int f = q; f < T; f += 4) {
    float v = 0.f;
    if (f < nf) v = fmaf(fr[(size_t)f * FB_NFILT + m], scale, off);
    xo[(size_t)f * FB_NFILT + m] = v;
  }
}

}  // namespace sar

extern "C" int sar_fbank_fwd(const float* wav, const long long* offsets, const float* melfb_t,
                             float* feat_ws, float* x_data, int B, int Fmax, int T, void* stream) {
  using namespace sar;
  SAR_REQUIRE(wav && offsets && melfb_t && feat_ws && x_data, SAR_ERR_BAD_ARG, "sar_fbank_fwd: null pointer");
  SAR_REQUIRE(B > 0 && Fmax > 0 && T > 0, SAR_ERR_BAD_ARG, "sar_fbank_fwd: non-positive dimension");
  SAR_REQUIRE(B <= 65535, SAR_ERR_UNSUPPORTED, "sar_fbank_fwd: B > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(fbank_frame_kernel, dim3(dim3(Fmax, B)), dim3(256), 0, st, wav, offsets, melfb_t, feat_ws, Fmax);
  int rc = check_launch("sar_fbank_fwd(frames)");
  if (rc) return rc;
  launch_k(fbank_norm_kernel, dim3(B), dim3(4 * FB_NFILT), 0, st, feat_ws, offsets, x_data, Fmax, T);
  return check_launch("sar_fbank_fwd(norm)");
}
