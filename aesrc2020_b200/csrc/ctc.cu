// ctc.cu -- ctc_pred softmax + K.ctc_batch_cost fused (model.py:268-269, 62-71).
//
// Reference data flow: Dense(softmax) on the GPU -> B*S*1000 fp32 probabilities copied to the
// HOST -> tf.nn.ctc_loss CPU kernel (TF 1.13 has no GPU CTC).  Here one CTA per utterance
//   1. per frame t (one warp per frame): row max / sum-exp of the 1000 logits, then
//      Z = sum_c (p_c + 1e-7)  (K.ctc_batch_cost takes log(p + eps) and tf.nn.ctc_loss
//      re-applies softmax, i.e. q = (p + 1e-7) / Z);
//   2. gathers only the <= 2L+1 log q[t, ext[s]] values the lattice needs into shared memory;
//   3. runs the alpha recursion in log space (blank = C-1, repeated labels need a blank),
//      one thread per lattice state, double-buffered, S sequential steps;
//   4. loss = -logsumexp(alpha[last], alpha[last-1]).
// The probabilities never leave the chip unless the caller asks for them (`probs`).
#include "common.cuh"

namespace sar {

constexpr int CTC_THREADS = 256;
constexpr float CTC_EPS = 1e-7f;

__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__global__ void __launch_bounds__(CTC_THREADS) ctc_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                           const int* __restrict__ in_len, const int* __restrict__ lab_len,
                                                           float* __restrict__ loss, float* __restrict__ probs,
                                                           int* __restrict__ status, int S, int C, int ld, int Lmax) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  const int NE = 2 * Lmax + 1;
  float* lq = sm;                               // [S][NE]
  float* alpha = lq + (size_t)S * NE;           // [2][NE]
  int* ext = reinterpret_cast<int*>(alpha + 2 * NE);   // [NE]
  __shared__ int bad;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  int T = in_len[b];
  int L = lab_len[b];
  if (t == 0) bad = 0;
  __syncthreads();
  if (T < 0 || T > S || L < 0 || L > Lmax) {   // malformed lengths
    if (t == 0) { loss[b] = INFINITY; if (status) status[b] = 2; }
    return;
  }
  const int blank = C - 1;
  const int n = 2 * L + 1;
  for (int s = t; s < n; s += CTC_THREADS) {
    int v = blank;
    if (s & 1) {
      v = (int)labels[(size_t)b * Lmax + (s >> 1)];   // K.ctc_batch_cost casts float labels to int32
      if (v < 0 || v >= blank) { bad = 1; v = 0; }
    }
    ext[s] = v;
  }
  __syncthreads();

  // ---- per-frame normalisers and gathered log q
  for (int f = warp; f < S; f += CTC_THREADS / 32) {
    const float* row = logits + ((size_t)b * S + f) * ld;          // rows may be padded (tensor-core ctc_pred: ld = 32-aligned C)
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, __ldg(row + c));
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += expf(__ldg(row + c) - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float z = 0.f;
    for (int c = lane; c < C; c += 32) {
      float pc = expf(__ldg(row + c) - m) * inv;
      if (probs) probs[((size_t)b * S + f) * C + c] = pc;
      z += pc + CTC_EPS;
    }
    z = warp_sum(z);
    const float logz = logf(z);
    if (f < T) {
      for (int s = lane; s < n; s += 32) {
        float pc = expf(__ldg(row + ext[s]) - m) * inv;
        lq[(size_t)f * NE + s] = logf(pc + CTC_EPS) - logz;
      }
    }
  }
  __syncthreads();

  // ---- alpha recursion
  float* a0 = alpha;
  float* a1 = alpha + NE;
  for (int s = t; s < n; s += CTC_THREADS) a0[s] = (s < 2 && T > 0) ? lq[s] : -INFINITY;
  __syncthreads();
  for (int f = 1; f < T; ++f) {
    for (int s = t; s < n; s += CTC_THREADS) {
      float x0 = a0[s];
      float x1 = (s >= 1) ? a0[s - 1] : -INFINITY;
      float x2 = (s >= 2 && (s & 1) && ext[s] != ext[s - 2]) ? a0[s - 2] : -INFINITY;
      a1[s] = lse3(x0, x1, x2) + lq[(size_t)f * NE + s];
    }
    __syncthreads();
    float* tmp = a0; a0 = a1; a1 = tmp;
  }
  if (t == 0) {
    float ll = (T > 0) ? ((n > 1) ? lse2(a0[n - 1], a0[n - 2]) : a0[0]) : -INFINITY;
    int st = bad ? 2 : ((ll == -INFINITY) ? 1 : 0);
    loss[b] = -ll;
    if (status) status[b] = st;
  }
}

// ------------------------------------------------------------------ training mode: d loss / d logits (fifth slice)
// ctc_kernel plus the beta recursion and the gradient of  loss_b = -log P(labels | q)  with respect to the PRE-softmax
// ctc_pred logits a, through K.ctc_batch_cost's double normalisation (model.py:62-71):
//   p = softmax(a),  u = log(p + 1e-7),  q = softmax(u) = (p + 1e-7) / Z;
//   d loss / d u[t,k] = q[t,k] - gamma[t,k],   gamma[t,k] = sum_{s: ext[s] = k} alpha_t(s) beta_t(s) / (P q[t,k])   (Graves 2006)
//   d loss / d p[t,j] = (q[t,j] - gamma[t,j]) / (p[t,j] + 1e-7),   d loss / d a[t,i] = p_i (dL/dp_i - sum_j p_j dL/dp_j).
// One CTA per utterance; alpha of every frame stays in shared memory, frames are finished from the last to the first as
// beta becomes available.  grad (B,S,C) receives scale * d loss_b / d a (frames >= in_len and rejected utterances: zeros).
__global__ void __launch_bounds__(CTC_THREADS) ctc_grad_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                                const int* __restrict__ in_len, const int* __restrict__ lab_len,
                                                                float* __restrict__ loss, float* __restrict__ grad,
                                                                int* __restrict__ status, int S, int C, int ld, int Lmax, float scale) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  const int NE = 2 * Lmax + 1;
  float* lq = sm;                               // [S][NE]
  float* al = lq + (size_t)S * NE;              // [S][NE]
  float* beta = al + (size_t)S * NE;            // [2][NE]
  float* stat = beta + 2 * NE;                  // [S][2]: row max, 1 / sum exp
  float* gam = stat + 2 * S;                    // [C]
  float* red = gam + C;                         // [32] block reduction scratch
  int* ext = reinterpret_cast<int*>(red + 32);  // [NE]
  __shared__ int bad;
  __shared__ float s_ll;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  const int T = in_len[b], L = lab_len[b];
  float* gb = grad + (size_t)b * S * C;
  if (t == 0) bad = 0;
  __syncthreads();
  const bool malformed = (T < 0 || T > S || L < 0 || L > Lmax);
  const int blank = C - 1;
  const int n = malformed ? 0 : 2 * L + 1;
  for (int s = t; s < n; s += CTC_THREADS) {
    int v = blank;
    if (s & 1) {
      v = (int)labels[(size_t)b * Lmax + (s >> 1)];
      if (v < 0 || v >= blank) { bad = 1; v = 0; }
    }
    ext[s] = v;
  }
  __syncthreads();
  for (int f = warp; f < S && !malformed; f += CTC_THREADS / 32) {
    const float* row = logits + ((size_t)b * S + f) * ld;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, __ldg(row + c));
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += expf(__ldg(row + c) - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float z = 0.f;
    for (int c = lane; c < C; c += 32) z += expf(__ldg(row + c) - m) * inv + CTC_EPS;
    z = warp_sum(z);
    if (lane == 0) { stat[2 * f] = m; stat[2 * f + 1] = inv; }
    const float logz = logf(z);
    if (f < T)
      for (int s = lane; s < n; s += 32) lq[(size_t)f * NE + s] = logf(expf(__ldg(row + ext[s]) - m) * inv + CTC_EPS) - logz;
  }
  __syncthreads();
  // ---- alpha, every frame kept
  for (int s = t; s < n; s += CTC_THREADS) al[s] = (s < 2 && T > 0) ? lq[s] : -INFINITY;
  __syncthreads();
  for (int f = 1; f < T && !malformed; ++f) {
    const float* a0 = al + (size_t)(f - 1) * NE;
    for (int s = t; s < n; s += CTC_THREADS) {
      const float x0 = a0[s];
      const float x1 = (s >= 1) ? a0[s - 1] : -INFINITY;
      const float x2 = (s >= 2 && (s & 1) && ext[s] != ext[s - 2]) ? a0[s - 2] : -INFINITY;
      al[(size_t)f * NE + s] = lse3(x0, x1, x2) + lq[(size_t)f * NE + s];
    }
    __syncthreads();
  }
  if (t == 0) {
    float ll = -INFINITY;
    if (!malformed && T > 0) {
      const float* aT = al + (size_t)(T - 1) * NE;
      ll = (n > 1) ? lse2(aT[n - 1], aT[n - 2]) : aT[0];
    }
    const int st = (malformed || bad) ? 2 : ((ll == -INFINITY) ? 1 : 0);
    loss[b] = malformed ? INFINITY : -ll;
    if (status) status[b] = st;
    s_ll = (st == 0) ? ll : -INFINITY;
  }
  __syncthreads();
  const float ll = s_ll;
  const bool ok = ll != -INFINITY;
  // ---- frames past the utterance (or all frames of a rejected one): zero gradient
  for (int f = ok ? T : 0; f < S; ++f)
    for (int c = t; c < C; c += CTC_THREADS) gb[(size_t)f * C + c] = 0.f;
  if (!ok) return;
  // ---- beta (emission included, like alpha) from the last frame down; each frame's gradient as soon as its beta is known
  const float Zc = 1.f + (float)C * CTC_EPS;    // sum_c (p_c + eps) up to rounding; the exact per-frame Z is recomputed below
  (void)Zc;
  float* b0 = beta;
  float* b1 = beta + NE;
  for (int f = T - 1; f >= 0; --f) {
    for (int s = t; s < n; s += CTC_THREADS) {
      float v;
      if (f == T - 1) v = (s >= n - 2) ? lq[(size_t)f * NE + s] : -INFINITY;
      else {
        const float x0 = b1[s];
        const float x1 = (s + 1 < n) ? b1[s + 1] : -INFINITY;
        const float x2 = (s + 2 < n && (s & 1) && ext[s + 2] != ext[s]) ? b1[s + 2] : -INFINITY;
        v = lse3(x0, x1, x2) + lq[(size_t)f * NE + s];
      }
      b0[s] = v;
    }
    for (int c = t; c < C; c += CTC_THREADS) gam[c] = 0.f;
    __syncthreads();
    for (int s = t; s < n; s += CTC_THREADS) {
      const float lg = al[(size_t)f * NE + s] + b0[s] - lq[(size_t)f * NE + s] - ll;
      if (lg > -80.f) atomicAdd(&gam[ext[s]], expf(lg));
    }
    __syncthreads();
    const float* row = logits + ((size_t)b * S + f) * ld;
    const float m = stat[2 * f], inv = stat[2 * f + 1];
    // pass 1: Z and sum_j p_j dL/dp_j need q, which needs Z: Z first
    float z = 0.f;
    for (int c = t; c < C; c += CTC_THREADS) z += expf(__ldg(row + c) - m) * inv + CTC_EPS;
    z = block_sum(z, red);
    const float invz = 1.f / z;
    float dot = 0.f;
    for (int c = t; c < C; c += CTC_THREADS) {
      const float pc = expf(__ldg(row + c) - m) * inv;
      dot += pc * ((pc + CTC_EPS) * invz - gam[c]) / (pc + CTC_EPS);
    }
    dot = block_sum(dot, red);
    for (int c = t; c < C; c += CTC_THREADS) {
      const float pc = expf(__ldg(row + c) - m) * inv;
      const float dp = ((pc + CTC_EPS) * invz - gam[c]) / (pc + CTC_EPS);
      gb[(size_t)f * C + c] = scale * pc * (dp - dot);
    }
    __syncthreads();
    float* tmp = b0; b0 = b1; b1 = tmp;
  }
}

// Greedy CTC decode (K.ctc_decode(..., greedy=True), model.py:385-389 -> tf.nn.ctc_greedy_decoder with
// merge_repeated=True): per frame the FIRST maximum over the C classes (softmax is monotone, so the pre-softmax
// logits give the same path), then collapse repeats and drop blanks (blank = C-1).  One CTA per utterance: a warp
// per frame for the argmax, one thread for the (<= S step) collapse.  dec (B, S) int32 padded with -1, len (B).
__global__ void __launch_bounds__(CTC_THREADS) ctc_greedy_kernel(const float* __restrict__ logits, const int* __restrict__ in_len,
                                                                  int fixed_len, int* __restrict__ dec, int* __restrict__ dec_len,
                                                                  int S, int C, int ld) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ int best[];                 // [S]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, b = blockIdx.x;
  int T = in_len ? in_len[b] : fixed_len;
  T = T < 0 ? 0 : (T > S ? S : T);
  for (int f = warp; f < T; f += CTC_THREADS / 32) {
    const float* row = logits + ((size_t)b * S + f) * ld;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = __ldg(row + c);
      if (v > bv) { bv = v; bi = c; }            // ascending c per lane: strict > keeps the first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) best[f] = bi;
  }
  __syncthreads();
  for (int f = t; f < S; f += CTC_THREADS) dec[(size_t)b * S + f] = -1;
  __syncthreads();
  if (t == 0) {
    const int blank = C - 1;
    int prev = -1, n = 0;
    for (int f = 0; f < T; ++f) {
      const int v = best[f];
      if (v != prev && v != blank) dec[(size_t)b * S + n++] = v;
      prev = v;
    }
    dec_len[b] = n;
  }
}

}  // namespace sar

extern "C" int sar_ctc_greedy_fwd(const float* logits, int ld, const int* in_len, int fixed_len, int* dec, int* dec_len,
                                  int B, int S, int C, void* stream) {
  using namespace sar;
  SAR_REQUIRE(logits && dec && dec_len, SAR_ERR_BAD_ARG, "sar_ctc_greedy_fwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0 && C > 1 && ld >= C, SAR_ERR_BAD_ARG, "sar_ctc_greedy_fwd: bad dimension");
  SAR_REQUIRE((size_t)S * sizeof(int) <= 48 * 1024, SAR_ERR_UNSUPPORTED, "sar_ctc_greedy_fwd: S too large");
  launch_k(ctc_greedy_kernel, dim3(B), dim3(CTC_THREADS), (size_t)S * sizeof(int), (cudaStream_t)stream, logits, in_len, fixed_len,
           dec, dec_len, S, C, ld);
  return check_launch("sar_ctc_greedy_fwd");
}

extern "C" int sar_ctc_fwd(const float* logits, const float* labels, const int* in_len, const int* lab_len,
                           float* loss, float* probs, int* status, int B, int S, int C, int Lmax, void* stream) {
  return sar_ctc_ld_fwd(logits, C, labels, in_len, lab_len, loss, probs, status, B, S, C, Lmax, stream);
}

extern "C" int sar_ctc_ld_fwd(const float* logits, int ld, const float* labels, const int* in_len, const int* lab_len,
                              float* loss, float* probs, int* status, int B, int S, int C, int Lmax, void* stream) {
  using namespace sar;
  SAR_REQUIRE(logits && labels && in_len && lab_len && loss, SAR_ERR_BAD_ARG, "sar_ctc_fwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0 && C > 1 && Lmax > 0 && ld >= C, SAR_ERR_BAD_ARG, "sar_ctc_fwd: bad dimension");
  const int NE = 2 * Lmax + 1;
  size_t smem = sizeof(float) * ((size_t)S * NE + 2 * NE) + sizeof(int) * NE;
  SAR_REQUIRE(smem <= 227 * 1024, SAR_ERR_UNSUPPORTED, "sar_ctc_fwd: S*(2*Lmax+1) too large for shared memory (%zu B)", smem);
  { const int arc = allow_max_smem(ctc_kernel, "sar_ctc_fwd"); if (arc) return arc; }
  launch_k(ctc_kernel, dim3(B), dim3(CTC_THREADS), smem, (cudaStream_t)stream, logits, labels, in_len, lab_len, loss, probs, status, S, C, ld, Lmax);
  return check_launch("sar_ctc_fwd");
}

extern "C" int sar_ctc_grad_fwd(const float* logits, int ld, const float* labels, const int* in_len, const int* lab_len,
                                float* loss, float* grad, int* status, int B, int S, int C, int Lmax, float scale, void* stream) {
  using namespace sar;
  SAR_REQUIRE(logits && labels && in_len && lab_len && loss && grad, SAR_ERR_BAD_ARG, "sar_ctc_grad_fwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0 && C > 1 && Lmax > 0 && ld >= C, SAR_ERR_BAD_ARG, "sar_ctc_grad_fwd: bad dimension");
  const int NE = 2 * Lmax + 1;
  const size_t smem = sizeof(float) * (2 * (size_t)S * NE + 2 * NE + 2 * (size_t)S + C + 32) + sizeof(int) * NE;
  SAR_REQUIRE(smem <= 227 * 1024, SAR_ERR_UNSUPPORTED, "sar_ctc_grad_fwd: S*(2*Lmax+1) too large for shared memory (%zu B)", smem);
  { const int arc = allow_max_smem(ctc_grad_kernel, "sar_ctc_grad_fwd"); if (arc) return arc; }
  launch_k(ctc_grad_kernel, dim3(B), dim3(CTC_THREADS), smem, (cudaStream_t)stream, logits, labels, in_len, lab_len, loss, grad, status,
           S, C, ld, Lmax, scale);
  return check_launch("sar_ctc_grad_fwd");
}
