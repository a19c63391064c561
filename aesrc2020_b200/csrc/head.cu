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

constexpr int HEAD_THREADS = 128;
constexpr int HEAD_MAXC = 32;      // max classes
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

__device__ __forceinline__ float ce_on_probs(const float* p, const float* y, int n) {
  // keras categorical_crossentropy on probabilities
  float sum = 0.f;
  for (int c = 0; c < n; ++c) sum += p[c];
  float l = 0.f;
  for (int c = 0; c < n; ++c) {
    float q = fminf(fmaxf(p[c] / sum, K_EPS), 1.f - K_EPS);
    l -= y[c] * logf(q);
  }
  return l;
}
__device__ __forceinline__ int argmax_first(const float* v, int n) {
  int a = 0;
  for (int c = 1; c < n; ++c) if (v[c] > v[a]) a = c;
  return a;
}
__device__ __forceinline__ void softmax_small(const float* logit, float* p, int n) {
  float m = logit[0];
  for (int c = 1; c < n; ++c) m = fmaxf(m, logit[c]);
  float sum = 0.f;
  for (int c = 0; c < n; ++c) { p[c] = expf(logit[c] - m); sum += p[c]; }
  for (int c = 0; c < n; ++c) p[c] /= sum;
}

__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(HeadP p) {
  extern __shared__ __align__(16) float sm[];
  float* x = sm;                         // [max(D,Dd)]
  float* h1 = x + max(p.D, p.Dd);        // [H1]
  float* h2 = h1 + p.H1;                 // [H2]   (h1|h2 region is at least HEAD_THREADS floats)
  float* la = h1 + max(p.H1 + p.H2, HEAD_THREADS);   // [n] accent logits
  float* ld = la + HEAD_MAXC;            // [n] disc cos / logits
  float* wn = ld + HEAD_MAXC;            // [n] column norms^2
  float* scratch = wn + HEAD_MAXC;       // [32]
  const int t = threadIdx.x, b = blockIdx.x, n = p.n;
  pdl_wait();
  pdl_trigger();
  const float* y = p.onehot ? p.onehot + (size_t)b * n : nullptr;
  float loss_a = 0.f, loss_d = 0.f, corr_a = 0.f, corr_d = 0.f;

  if (p.w1) {
    for (int d = t; d < p.D; d += HEAD_THREADS) x[d] = __ldg(p.emb + (size_t)b * p.D + d);
    __syncthreads();
    for (int j = t; j < p.H1; j += HEAD_THREADS) {
      float acc = __ldg(p.b1 + j), acc2 = 0.f;
      int d = 0;
#pragma unroll 8
      for (; d + 1 < p.D; d += 2) {          // two chains, loads batched by the unroll
        acc = fmaf(x[d], __ldg(p.w1 + (size_t)d * p.H1 + j), acc);
        acc2 = fmaf(x[d + 1], __ldg(p.w1 + (size_t)(d + 1) * p.H1 + j), acc2);
      }
      if (d < p.D) acc = fmaf(x[d], __ldg(p.w1 + (size_t)d * p.H1 + j), acc);
      h1[j] = fmaxf(acc + acc2, 0.f);
    }
    __syncthreads();
    for (int j = t; j < p.H2; j += HEAD_THREADS) {
      float acc = __ldg(p.b2 + j);
      for (int d = 0; d < p.H1; ++d) acc = fmaf(h1[d], __ldg(p.w2 + (size_t)d * p.H2 + j), acc);
      h2[j] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (t < n) {
      float acc = __ldg(p.b3 + t);
      for (int d = 0; d < p.H2; ++d) acc = fmaf(h2[d], __ldg(p.w3 + (size_t)d * n + t), acc);
      la[t] = acc;
    }
    __syncthreads();
    if (t == 0) {
      float pr[HEAD_MAXC];
      softmax_small(la, pr, n);
      for (int c = 0; c < n; ++c) {
        if (p.y_accent) p.y_accent[(size_t)b * n + c] = pr[c];
        if (p.y_accent_logits) p.y_accent_logits[(size_t)b * n + c] = la[c];
      }
      if (y) {
        loss_a = ce_on_probs(pr, y, n);
        corr_a = (argmax_first(pr, n) == argmax_first(y, n)) ? 1.f : 0.f;
      }
    }
    __syncthreads();
  }

  if (p.head != SAR_HEAD_NONE) {
    const float* e = p.emb_d ? p.emb_d : p.emb;
    const int Dd = p.emb_d ? p.Dd : p.D;
    float ssq = 0.f;
    for (int d = t; d < Dd; d += HEAD_THREADS) {
      float v = __ldg(e + (size_t)b * Dd + d);
      x[d] = v;
      ssq += v * v;
    }
    ssq = block_sum(ssq, scratch);      // contains the barriers that publish x[]
    const bool normalise_x = p.head != SAR_HEAD_SOFTMAX && p.head != SAR_HEAD_CIRCLE_RAW;
    const float xinv = normalise_x ? 1.0f / sqrtf(fmaxf(ssq, 1e-12f)) : 1.f;
    // x^ . W[:, c] and |W[:, c]|^2: classes x (HEAD_THREADS / n) d-slices, partials through smem
    const int parts = HEAD_THREADS / n;
    float* pacc = h1;                       // [parts][n]  (h1/h2 are free after the classifier)
    float* pwss = wn + HEAD_MAXC + 32;      // placed after scratch: [parts][n]
    if (t < parts * n) {
      const int c = t % n, part = t / n;
      const int dlo = (int)(((long long)Dd * part) / parts), dhi = (int)(((long long)Dd * (part + 1)) / parts);
      float a = 0.f, w2 = 0.f;
#pragma unroll 4
      for (int d = dlo; d < dhi; ++d) {
        const float wv = __ldg(p.wd + (size_t)d * n + c);
        a = fmaf(x[d] * xinv, wv, a);
        w2 = fmaf(wv, wv, w2);
      }
      pacc[part * n + c] = a;
      pwss[part * n + c] = w2;
    }
    __syncthreads();
    if (t < n) {
      float acc = 0.f, wss = 0.f;
      for (int part = 0; part < parts; ++part) { acc += pacc[part * n + t]; wss += pwss[part * n + t]; }
      const bool face = p.head == SAR_HEAD_SPHEREFACE || p.head == SAR_HEAD_COSFACE || p.head == SAR_HEAD_ARCFACE;
      if (face) acc *= 1.0f / sqrtf(fmaxf(wss, 1e-12f));      // W normalised per column, losses.py:34,80,127
      ld[t] = acc;
    }
    __syncthreads();
    if (t == 0) {
      float lg[HEAD_MAXC], pr[HEAD_MAXC];
      if (p.head == SAR_HEAD_CIRCLE || p.head == SAR_HEAD_CIRCLE_RAW) {
        for (int c = 0; c < n; ++c) {
          if (p.y_disc) p.y_disc[(size_t)b * n + c] = ld[c];
          if (p.y_disc_logits) p.y_disc_logits[(size_t)b * n + c] = ld[c];
        }
        if (y) {
          // circle_loss, losses.py:157-172
          const float m = p.margin;
          for (int c = 0; c < n; ++c) {
            float ap = fmaxf(1.f + m - ld[c], 0.f), an = fmaxf(ld[c] + m, 0.f);
            lg[c] = (y[c] * (ap * (ld[c] - (1.f - m))) + (1.f - y[c]) * (an * (ld[c] - m))) * p.gamma;
          }
          float mx = lg[0];
          for (int c = 1; c < n; ++c) mx = fmaxf(mx, lg[c]);
          float sum = 0.f;
          for (int c = 0; c < n; ++c) sum += expf(lg[c] - mx);
          float lse = mx + logf(sum);
          for (int c = 0; c < n; ++c) loss_d -= y[c] * (lg[c] - lse);
          corr_d = (argmax_first(ld, n) == argmax_first(y, n)) ? 1.f : 0.f;
        }
      } else {
        for (int c = 0; c < n; ++c) {
          float v = ld[c];
          if (p.head != SAR_HEAD_SOFTMAX) {
            float yc = y ? y[c] : 0.f;
            float target = v;
            if (p.head == SAR_HEAD_COSFACE) target = v - p.margin;
            else {
              float th = acosf(fminf(fmaxf(v, -1.f + K_EPS), 1.f - K_EPS));
              target = (p.head == SAR_HEAD_SPHEREFACE) ? cosf(p.margin * th) : cosf(th + p.margin);
            }
            v = (v * (1.f - yc) + target * yc) * p.s;
          }
          lg[c] = v;
        }
        softmax_small(lg, pr, n);
        for (int c = 0; c < n; ++c) {
          if (p.y_disc) p.y_disc[(size_t)b * n + c] = pr[c];
          if (p.y_disc_logits) p.y_disc_logits[(size_t)b * n + c] = lg[c];
        }
        if (y) {
          loss_d = ce_on_probs(pr, y, n);
          corr_d = (argmax_first(pr, n) == argmax_first(y, n)) ? 1.f : 0.f;
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
  size_t smem = sizeof(float) * ((size_t)dmax + (H1 + H2 > HEAD_THREADS ? H1 + H2 : HEAD_THREADS) + 3 * HEAD_MAXC + 32 + HEAD_THREADS);
  SAR_REQUIRE(smem <= 200 * 1024, SAR_ERR_UNSUPPORTED, "sar_head_fwd: embedding too wide");
  HeadP p{emb, D, w1, b1, H1, w2, b2, H2, w3, b3, emb_d, emb_d ? Dd : D, wd, onehot, n_classes, head,
          margin, s, gamma, y_accent, y_accent_logits, y_disc, y_disc_logits, sample_stats};
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("sar_head_fwd: %s", cudaGetErrorString(e)); return (int)e; }
  }
  launch_k(head_kernel, dim3(B), dim3(HEAD_THREADS), smem, (cudaStream_t)stream, p);
  return check_launch("sar_head_fwd");
}
