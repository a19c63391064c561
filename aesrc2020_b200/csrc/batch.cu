// batch.cu -- on-device batch assembly: what utils.data_loader (utils.py:71-117) does per batch on the
// host with numpy / sklearn, from ONE contiguous upload of the un-padded features:
//   x_data[b]      = feat_reshape(feat_norm(feat_b), T)     utils.py:35-46,91   per-utterance, per-bin MinMaxScaler over
//                    ALL frames of the utterance (zero range -> scale 1), then truncate / zero-pad to T frames
//   x_ctc_label[b] = text_ids_norm(trans_b, Lmax)            utils.py:57-63,95  truncate, pad with EOS_ID = 2 (float32)
//   x_ctc_out_len  = min(len(trans_b), Lmax), x_ctc_in_len = encoder_len        utils.py:96-97
//   x_accent[b]    = to_categorical(accent_b, n_classes)     utils.py:100
// The feature kernel is HBM bound: every feature value is read twice (min/max pass, scale pass; the second read hits
// L2 for utterances below ~100 MB) and written once; one CTA per utterance, thread = (bin, frame phase), so a warp reads
// whole 320-byte rows.
#include "common.cuh"

namespace sar {

constexpr int BA_THREADS = 512;
constexpr int BA_EOS_ID = 2;          // utils.py:55

__global__ void __launch_bounds__(BA_THREADS) feat_batch_kernel(const float* __restrict__ feats, const long long* __restrict__ offsets,
                                                                 float* __restrict__ x_data, int T, int D) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];                 // [phases][D] min | [phases][D] max
  const int b = blockIdx.x, t = threadIdx.x;
  const int phases = BA_THREADS / D;            // frame phases (D <= BA_THREADS)
  const int m = t % D, q = t / D;
  const long long f0 = offsets[b], nf = offsets[b + 1] - offsets[b];
  const float* fr = feats + (size_t)f0 * D;
  float* smin = sm;
  float* smax = sm + phases * D;
  float mn = INFINITY, mx = -INFINITY;
  if (q < phases) {
    for (long long f = q; f < nf; f += phases) {
      const float v = __ldg(fr + (size_t)f * D + m);
      mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    smin[q * D + m] = mn; smax[q * D + m] = mx;
  }
  __syncthreads();
  if (q < phases) {
    for (int k = 0; k < phases; ++k) { mn = fminf(mn, smin[k * D + m]); mx = fmaxf(mx, smax[k * D + m]); }
    float rng = mx - mn;
    if (!(rng > 0.f)) rng = 1.f;                // sklearn _handle_zeros_in_scale: zero range -> scale 1 -> zeros
    const float scale = 1.f / rng, off = -mn * scale;   // MinMaxScaler: X * scale_ + min_, scale_ = 1/range, min_ = -min*scale_
    float* xo = x_data + (size_t)b * T * D;
    for (int f = q; f < T; f += phases) {
      float v = 0.f;
      if (f < nf) v = fmaf(__ldg(fr + (size_t)f * D + m), scale, off);
      xo[(size_t)f * D + m] = v;
    }
  }
}

__global__ void labels_pack_kernel(const int* __restrict__ accent, int n_classes, float* __restrict__ onehot,
                                   const int* __restrict__ trans, const long long* __restrict__ trans_off, int Lmax, int encoder_len,
                                   float* __restrict__ ctc_label, int* __restrict__ ctc_out_len, int* __restrict__ ctc_in_len,
                                   int* __restrict__ status, int B) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x, t = threadIdx.x;
  if (b >= B) return;
  if (onehot) {
    const int a = accent[b];
    if (t == 0 && status && (a < 0 || a >= n_classes)) atomicOr(status, 1);      // to_categorical would raise IndexError
    for (int c = t; c < n_classes; c += blockDim.x) onehot[(size_t)b * n_classes + c] = (c == a) ? 1.f : 0.f;
  }
  if (ctc_label) {
    const long long o = trans_off[b], L = trans_off[b + 1] - trans_off[b];
    for (int i = t; i < Lmax; i += blockDim.x) ctc_label[(size_t)b * Lmax + i] = (i < L) ? (float)trans[o + i] : (float)BA_EOS_ID;
    if (t == 0) {
      ctc_out_len[b] = (int)(L < Lmax ? L : Lmax);
      ctc_in_len[b] = encoder_len;
    }
  }
}

}  // namespace sar

extern "C" int sar_feat_batch_fwd(const float* feats, const long long* frame_offsets, float* x_data, int B, int T, int D, void* stream) {
  using namespace sar;
  SAR_REQUIRE(feats && frame_offsets && x_data, SAR_ERR_BAD_ARG, "sar_feat_batch_fwd: null pointer");
  SAR_REQUIRE(B > 0 && T > 0 && D > 0 && D <= BA_THREADS, SAR_ERR_BAD_ARG, "sar_feat_batch_fwd: need B, T > 0 and 0 < D <= %d", BA_THREADS);
  const int phases = BA_THREADS / D;
  launch_k(feat_batch_kernel, dim3(B), dim3(BA_THREADS), (size_t)2 * phases * D * sizeof(float), (cudaStream_t)stream, feats, frame_offsets, x_data, T, D);
  return check_launch("sar_feat_batch_fwd");
}

extern "C" int sar_labels_pack_fwd(const int* accent, int n_classes, float* onehot,
                                   const int* trans, const long long* trans_offsets, int Lmax, int encoder_len,
                                   float* ctc_label, int* ctc_out_len, int* ctc_in_len, int* status, int B, void* stream) {
  using namespace sar;
  SAR_REQUIRE(B > 0, SAR_ERR_BAD_ARG, "sar_labels_pack_fwd: B <= 0");
  SAR_REQUIRE(onehot || ctc_label, SAR_ERR_BAD_ARG, "sar_labels_pack_fwd: nothing to pack");
  SAR_REQUIRE(!onehot || (accent && n_classes > 0), SAR_ERR_BAD_ARG, "sar_labels_pack_fwd: one-hot needs accent ids and n_classes");
  SAR_REQUIRE(!ctc_label || (trans && trans_offsets && Lmax > 0 && ctc_out_len && ctc_in_len), SAR_ERR_BAD_ARG,
              "sar_labels_pack_fwd: CTC labels need trans, trans_offsets, Lmax and the two length outputs");
  launch_k(labels_pack_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, accent, n_classes, onehot, trans, trans_offsets, Lmax,
           encoder_len, ctc_label, ctc_out_len, ctc_in_len, status, B);
  return check_launch("sar_labels_pack_fwd");
}
