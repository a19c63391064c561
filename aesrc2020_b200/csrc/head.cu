// head.cu -- classifier MLP + margin head + per-sample losses, one CTA per utterance.
//
// Replaces (all inference-mode):
//   AR_CF_DS1 / AR_CF_DS2 / y_accent   Dense 256->64 relu ->64 relu ->n softmax   model.py:294-296
//   disc_loss                          head selection                             model.py:142-167
//   SphereFace / CosFace / ArcFace     x^=x/|x|, W^=W/|W|col, cos, margin on the
//                                      target class, *s, softmax                  losses.py:28-47,74-91,121-144
//   circle-loss head + circle_loss     l2norm(x) @ W (raw), gamma=256             model.py:161-163, losses.py:157-172
//   categorical_crossentropy/accuracy  p/=sum p; clip(1e-7,1-1e-7); -sum y log p  model.py:344-357
// The reference spends ~25 launch-bound library kernels on this; here the embedding row is
// read once and every output of the row is produced by one small CTA.
#include "common.cuh"

namespace sar {

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_MAXC = 32;      // max classes (one warp lane per class in the final stage)
constexpr float K_EPS = 1e-7f;

struct HeadP {
  const float* emb; int D;
  const float* w1; const float* b1; int H1;
  const float* w2; const float* b2; int H2;
  const float* w3; const float* b3;
  const float* emb_d; int Dd; const float* wd;
  const float* onehot; int n; int head; float margin, s, gamma;
  float* y_accent; float* y_accent_logits; float* y_disc; float* y_disc_logits; float* stats;
};

// y[j] = sum_d x[d] * w[d*ldw + j] for j < nout, the d range cut into `parts` slices so that all HEAD_THREADS
// threads carry independent load streams (the op is pure L2 latency: a 256 x 64 weight block is read once per
// utterance); slice partials meet in shared memory and are summed in a fixed order (deterministic).
// With SQ, wsq[j] = sum_d w[d][j]^2 as well (the column norms of the Face heads).  Ends with a barrier.
template <bool SQ>
__device__ __forceinline__ void matvec_parts(const float* x, int D, const float* __restrict__ w, int nout, float xscale,
                                             float* pacc, float* psq, float* y, float* wsq, int t) {
  if (nout <= HEAD_THREADS) {
    const int parts = HEAD_THREADS / nout;
    if (t < parts * nout) {
      const int j = t % nout, part = t / nout;
      const int dlo = (int)(((long long)D * part) / parts), dhi = (int)(((long long)D * (part + 1)) / parts);
      float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
      int d = dlo;
#pragma unroll 8
      for (; d + 1 < dhi; d += 2) {
        const float w0 = __ldg(w + (size_t)d * nout + j), w1 = __ldg(w + (size_t)(d + 1) * nout + j);
        a0 = fmaf(x[d] * xscale, w0, a0);
        a1 = fmaf(x[d + 1] * xscale, w1, a1);
        if (SQ) { q0 = fmaf(w0, w0, q0); q1 = fmaf(w1, w1, q1); }
      }
      if (d < dhi) {
        const float w0 = __ldg(w + (size_t)d * nout + j);
        a0 = fmaf(x[d] * xscale, w0, a0);
        if (SQ) q0 = fmaf(w0, w0, q0);
      }
      pacc[part * nout + j] = a0 + a1;
      if (SQ) psq[part * nout + j] = q0 + q1;
    }
    __syncthreads();
    if (t < nout) {
      float a = 0.f, q = 0.f;
      for (int part = 0; part < parts; ++part) { a += pacc[part * nout + t]; if (SQ) q += psq[part * nout + t]; }
      y[t] = a;
      if (SQ) wsq[t] = q;
    }
  } else {                                   // wide layers: one output per thread per round
    for (int j = t; j < nout; j += HEAD_THREADS) {
      float a = 0.f, q = 0.f;
      for (int d = 0; d < D; ++d) {
        const float w0 = __ldg(w + (size_t)d * nout + j);
        a = fmaf(x[d] * xscale, w0, a);
        if (SQ) q = fmaf(w0, w0, q);
      }
      y[j] = a;
      if (SQ) wsq[j] = q;
    }
  }
  __syncthreads();
}

// lane c < n holds v; returns the index of the first maximum (np.argmax tie rule) on every lane
__device__ __forceinline__ int warp_argmax_first(float v, int lane, int n) {
  float bv = lane < n ? v : -INFINITY;
  int bi = lane < n ? lane : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  return bi;
}
// softmax over the n live lanes
__device__ __forceinline__ float warp_softmax(float logit, int lane, int n) {
  const float m = warp_max(lane < n ? logit : -INFINITY);
  const float e = lane < n ? expf(logit - m) : 0.f;
  return e / warp_sum(e);
}
// keras categorical_crossentropy on probabilities: p /= sum p; clip(1e-7, 1-1e-7); -sum y log p
__device__ __forceinline__ float warp_ce_on_probs(float pr, float y, int lane, int n) {
  const float sum = warp_sum(lane < n ? pr : 0.f);
  const float q = fminf(fmaxf(pr / sum, K_EPS), 1.f - K_EPS);
  return warp_sum(lane < n ? -y * logf(q) : 0.f);
}

__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(HeadP p) {
  extern __shared__ __align__(16) float sm[];
  const int Dmax = max(p.D, p.Dd), Hmax = max(max(p.H1, p.H2), HEAD_MAXC);
  float* x = sm;                         // [Dmax]
  float* h1 = x + Dmax;                  // [Hmax]
  float* h2 = h1 + Hmax;                 // [Hmax]
  float* wn = h2 + Hmax;                 // [HEAD_MAXC] column norms^2
  float* scratch = wn + HEAD_MAXC;       // [32]
  float* pacc = scratch + 32;            // [HEAD_THREADS]
  float* psq = pacc + HEAD_THREADS;      // [HEAD_THREADS]
  const int t = threadIdx.x, lane = t & 31, b = blockIdx.x, n = p.n;
  pdl_wait();
  pdl_trigger();
  const float* yrow = p.onehot ? p.onehot + (size_t)b * n : nullptr;
  const float yc = (yrow && t < n) ? __ldg(yrow + t) : 0.f;        // lane c of warp 0: one-hot entry of class c
  float loss_a = 0.f, loss_d = 0.f, corr_a = 0.f, corr_d = 0.f;

  if (p.w1) {
    for (int d = t; d < p.D; d += HEAD_THREADS) x[d] = __ldg(p.emb + (size_t)b * p.D + d);
    __syncthreads();
    matvec_parts<false>(x, p.D, p.w1, p.H1, 1.f, pacc, psq, h1, nullptr, t);
    for (int j = t; j < p.H1; j += HEAD_THREADS) h1[j] = relu_nan(h1[j] + __ldg(p.b1 + j));
    __syncthreads();
    matvec_parts<false>(h1, p.H1, p.w2, p.H2, 1.f, pacc, psq, h2, nullptr, t);
    for (int j = t; j < p.H2; j += HEAD_THREADS) h2[j] = relu_nan(h2[j] + __ldg(p.b2 + j));
    __syncthreads();
    matvec_parts<false>(h2, p.H2, p.w3, n, 1.f, pacc, psq, h1, nullptr, t);      // h1[0..n) = logits - bias
    if (t < 32) {
      const float la = lane < n ? h1[lane] + __ldg(p.b3 + lane) : 0.f;
      const float pr = warp_softmax(la, lane, n);
      if (lane < n) {
        if (p.y_accent) p.y_accent[(size_t)b * n + lane] = pr;
        if (p.y_accent_logits) p.y_accent_logits[(size_t)b * n + lane] = la;
      }
      if (yrow) {
        loss_a = warp_ce_on_probs(pr, yc, lane, n);
        corr_a = (warp_argmax_first(pr, lane, n) == warp_argmax_first(yc, lane, n)) ? 1.f : 0.f;
      }
    }
    __syncthreads();
  }

  if (p.head != SAR_HEAD_NONE) {
    const float* e = p.emb_d ? p.emb_d : p.emb;
    const int Dd = p.emb_d ? p.Dd : p.D;
    float ssq = 0.f;
    for (int d = t; d < Dd; d += HEAD_THREADS) {
      const float v = __ldg(e + (size_t)b * Dd + d);
      x[d] = v;
      ssq += v * v;
    }
    ssq = block_sum(ssq, scratch);      // contains the barriers that publish x[]
    const bool normalise_x = p.head != SAR_HEAD_SOFTMAX && p.head != SAR_HEAD_CIRCLE_RAW;
    const float xinv = normalise_x ? 1.0f / sqrtf(fmaxf(ssq, 1e-12f)) : 1.f;
    matvec_parts<true>(x, Dd, p.wd, n, xinv, pacc, psq, h2, wn, t);               // h2[c] = x^ . W[:, c]
    if (t < 32) {
      const bool face = p.head == SAR_HEAD_SPHEREFACE || p.head == SAR_HEAD_COSFACE || p.head == SAR_HEAD_ARCFACE;
      float v = lane < n ? h2[lane] : 0.f;
      if (face && lane < n) v *= 1.0f / sqrtf(fmaxf(wn[lane], 1e-12f));          // W normalised per column, losses.py:34,80,127
      if (p.head == SAR_HEAD_CIRCLE || p.head == SAR_HEAD_CIRCLE_RAW) {
        if (lane < n) {
          if (p.y_disc) p.y_disc[(size_t)b * n + lane] = v;
          if (p.y_disc_logits) p.y_disc_logits[(size_t)b * n + lane] = v;
        }
        if (yrow) {
          // circle_loss, losses.py:157-172
          const float m = p.margin;
          const float ap = fmaxf(1.f + m - v, 0.f), an = fmaxf(v + m, 0.f);
          const float lg = (yc * (ap * (v - (1.f - m))) + (1.f - yc) * (an * (v - m))) * p.gamma;
          const float mx = warp_max(lane < n ? lg : -INFINITY);
          const float lse = mx + logf(warp_sum(lane < n ? expf(lg - mx) : 0.f));
          loss_d = warp_sum(lane < n ? -yc * (lg - lse) : 0.f);
          corr_d = (warp_argmax_first(v, lane, n) == warp_argmax_first(yc, lane, n)) ? 1.f : 0.f;
        }
      } else {
        float lg = v;
        if (p.head != SAR_HEAD_SOFTMAX) {
          float target = v;
          if (p.head == SAR_HEAD_COSFACE) target = v - p.margin;
          else {
            const float th = acosf(fminf(fmaxf(v, -1.f + K_EPS), 1.f - K_EPS));
            target = (p.head == SAR_HEAD_SPHEREFACE) ? cosf(p.margin * th) : cosf(th + p.margin);
          }
          lg = (v * (1.f - yc) + target * yc) * p.s;
        }
        const float pr = warp_softmax(lg, lane, n);
        if (lane < n) {
          if (p.y_disc) p.y_disc[(size_t)b * n + lane] = pr;
          if (p.y_disc_logits) p.y_disc_logits[(size_t)b * n + lane] = lg;
        }
        if (yrow) {
          loss_d = warp_ce_on_probs(pr, yc, lane, n);
          corr_d = (warp_argmax_first(pr, lane, n) == warp_argmax_first(yc, lane, n)) ? 1.f : 0.f;
        }
      }
    }
  }
  if (t == 0 && p.stats) {
    p.stats[(size_t)b * 4 + 0] = loss_a;
    p.stats[(size_t)b * 4 + 1] = loss_d;
    p.stats[(size_t)b * 4 + 2] = corr_a;
    p.stats[(size_t)b * 4 + 3] = corr_d;
  }
}

}  // namespace sar

extern "C" int sar_head_fwd(const float* emb, int D,
                            const float* w1, const float* b1, int H1,
                            const float* w2, const float* b2, int H2,
                            const float* w3, const float* b3,
                            const float* emb_d, int Dd, const float* wd,
                            const float* onehot, int n_classes, int head, float margin, float s, float gamma,
                            float* y_accent, float* y_accent_logits, float* y_disc, float* y_disc_logits,
                            float* sample_stats, int B, void* stream) {
  using namespace sar;
  SAR_REQUIRE(B > 0 && n_classes > 0 && n_classes <= HEAD_MAXC, SAR_ERR_BAD_ARG,
              "sar_head_fwd: need 0 < n_classes <= %d, B > 0", HEAD_MAXC);
  SAR_REQUIRE(head >= SAR_HEAD_NONE && head <= SAR_HEAD_CIRCLE_RAW, SAR_ERR_BAD_ARG, "sar_head_fwd: bad head %d", head);
  const bool has_cls = w1 != nullptr;
  if (has_cls) {
    SAR_REQUIRE(emb && b1 && w2 && b2 && w3 && b3 && D > 0 && H1 > 0 && H2 > 0, SAR_ERR_BAD_ARG,
                "sar_head_fwd: incomplete classifier weights");
  } else {
    H1 = 0; H2 = 0;
  }
  if (head != SAR_HEAD_NONE) {
    SAR_REQUIRE(wd && (emb_d ? Dd > 0 : (emb && D > 0)), SAR_ERR_BAD_ARG, "sar_head_fwd: margin head needs wd and an embedding");
    const bool face = head == SAR_HEAD_SPHEREFACE || head == SAR_HEAD_COSFACE || head == SAR_HEAD_ARCFACE;
    SAR_REQUIRE(!face || onehot, SAR_ERR_BAD_ARG, "sar_head_fwd: Face heads need the one-hot label input (x_accent)");
  }
  SAR_REQUIRE(has_cls || head != SAR_HEAD_NONE, SAR_ERR_BAD_ARG, "sar_head_fwd: nothing to compute");
  if (!emb) D = 0;
  int dmax = D > 0 ? D : 0;
  if (emb_d && Dd > dmax) dmax = Dd;
  int hmax = H1 > H2 ? H1 : H2;
  if (hmax < HEAD_MAXC) hmax = HEAD_MAXC;
  size_t smem = sizeof(float) * ((size_t)dmax + 2 * (size_t)hmax + HEAD_MAXC + 32 + 2 * HEAD_THREADS);
  SAR_REQUIRE(smem <= 200 * 1024, SAR_ERR_UNSUPPORTED, "sar_head_fwd: embedding too wide");
  HeadP p{emb, D, w1, b1, H1, w2, b2, H2, w3, b3, emb_d, emb_d ? Dd : D, wd, onehot, n_classes, head,
          margin, s, gamma, y_accent, y_accent_logits, y_disc, y_disc_logits, sample_stats};
  { const int arc = allow_max_smem(head_kernel, "sar_head_fwd"); if (arc) return arc; }
  launch_k(head_kernel, dim3(B), dim3(HEAD_THREADS), smem, (cudaStream_t)stream, p);
  return check_launch("sar_head_fwd");
}
