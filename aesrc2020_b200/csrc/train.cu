// train.cu -- first slice of TRAINING mode: the building blocks of one optimisation step of the accent
// head (model.py:286-296, 142-167 in training mode; compile(), model.py:187-201: Adam(lr, decay=2e-4)).
//
//   sar_gemm_fwd          C = alpha * op(A) op(B) + beta * C    the Dense forward / backward contractions (x W, g W^T, x^T g)
//   sar_bn_train_fwd/bwd  BatchNormalization in training mode: batch statistics per replica (what multi_gpu_model does),
//                         eps 1e-3, moving averages with momentum 0.99; and its backward
//   sar_bias_act_fwd      y = act(x + b)                        Dense epilogue (relu / none)
//   sar_relu_bwd          g * (h > 0)
//   sar_l2norm_fwd/bwd    K.l2_normalize along rows or columns and its backward (Face heads, Circle-Loss head)
//   sar_head_grad_fwd     softmax + categorical cross-entropy of y_accent and the margin head (SphereFace / CosFace /
//                         ArcFace / Dense softmax / Circle-Loss): per-sample losses and d loss / d logits, d loss / d cos
//   sar_adam_fwd          Keras Adam update (+ l2 regulariser gradient, + unit_norm constraint helper)
//   sar_gru_gate_fwd/bwd  one time step of CuDNNGRU in training mode (gates kept) and its backward (fourth slice: CNN_LIN, CRNN)
//   sar_conv2d_bwd_data / _bwd_weight, sar_maxpool2d_bwd, sar_axpy_fwd   the ResNet's backward (sixth slice; correctness-first fp32)
//   sar_vlad_train_fwd/bwd  NetVLAD / GhostVLAD pooling in training mode: soft assignments kept, gradients of the centers and of
//                         the assignment scores (second slice: the pooling layer is trained together with the head)
//
// All fp32, CUDA cores, deterministic reduction orders.  This slice is about CORRECT training arithmetic behind the
// C ABI (checked against float64 autograd, tests/test_gpu_train.py), not yet about speed: the contractions of the head are
// <= 0.3 GFLOP per step.
#include "common.cuh"

namespace sar {

// ------------------------------------------------------------------ small fp32 GEMM, row-major, optional transposes
constexpr int TG = 64, TGK = 32;
// 64 x 64 tile per CTA, 32 k-values per round; the next round's global loads are issued (into registers) before the current
// round's FMAs, so the ~1 us round trip overlaps the math -- the GRU's per-step products (K = 256 / 768 on 4-12 CTAs) are
// round-trip bound: 16-value rounds without prefetch cost 2.3 us each.
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                    int M, int N, int K, int ta, int tb, float alpha, float beta) {
  pdl_wait();
  pdl_trigger();
  __shared__ float As[TGK][TG + 4], Bs[TGK][TG + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TG, n0 = blockIdx.x * TG;
  constexpr int PER = TG * TGK / 256;                       // elements of each operand tile per thread (8)
  float ra[PER], rb[PER];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = threadIdx.x + 256 * j;
      // A tile: element (m, k) = ta ? A[k][m] : A[m][k]; index so that consecutive threads read consecutive addresses
      int m, k;
      if (ta) { m = i % TG; k = i / TG; } else { k = i % TGK; m = i / TGK; }
      const int gm = m0 + m, gk = k0 + k;
      ra[j] = (gm < M && gk < K) ? (ta ? A[(size_t)gk * M + gm] : A[(size_t)gm * K + gk]) : 0.f;
      int n, kb;
      if (tb) { kb = i % TGK; n = i / TGK; } else { n = i % TG; kb = i / TG; }
      const int gn = n0 + n, gkb = k0 + kb;
      rb[j] = (gn < N && gkb < K) ? (tb ? B[(size_t)gn * K + gkb] : B[(size_t)gkb * N + gn]) : 0.f;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = threadIdx.x + 256 * j;
      int m, k;
      if (ta) { m = i % TG; k = i / TG; } else { k = i % TGK; m = i / TGK; }
      As[k][m] = ra[j];
      int n, kb;
      if (tb) { kb = i % TGK; n = i / TGK; } else { n = i % TG; kb = i / TG; }
      Bs[kb][n] = rb[j];
    }
  };
  float acc[4][4] = {};
  load_tile(0);
  for (int k0 = 0; k0 < K; k0 += TGK) {
    store_tile();
    __syncthreads();
    if (k0 + TGK < K) load_tile(k0 + TGK);
#pragma unroll
    for (int k = 0; k < TGK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) {
        float* c = C + (size_t)gm * N + gn;
        *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
      }
    }
}

// Skinny GEMM, M <= 64 rows, A not transposed: the per-time-step products of the GRU (h U, d_hu U^T with M = batch) and the
// embedding Dense at small batches.  The 64 x 64 tiles of gemm_kernel leave all but a few CTAs idle there and pay one
// global-memory round trip per 16 k-values (2-3 us each).  Here a CTA owns 32 output columns, lane = column, the 8 warps
// split K; every lane keeps its M accumulators in registers, A values are warp-broadcast loads, B is read coalesced
// (tb = 0: row k across the lanes) or as one contiguous row per lane (tb = 1).  The warps' partial sums are added in warp
// order: a fixed summation order.  (M <= 32: 8 warps; M <= 64: 4 warps, the partial sums must fit 48 KB of shared memory.)
template <int MR, int NW>
__global__ void __launch_bounds__(NW * 32) gemm_skinny_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                          int M, int N, int K, int tb, float alpha, float beta) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[NW][MR][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const int per = (K + NW - 1) / NW;
  const int k0 = warp * per, k1 = min(K, k0 + per);
  float acc[MR];
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;
  if (n < N) {
    const float* bp = tb ? B + (size_t)n * K : B + n;
    const size_t bstep = tb ? 1 : (size_t)N;
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
      const float bv = __ldg(bp + (size_t)k * bstep);
#pragma unroll
      for (int m = 0; m < MR; ++m)
        if (m < M) acc[m] = fmaf(__ldg(A + (size_t)m * K + k), bv, acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) red[warp][m][lane] = acc[m];
  __syncthreads();
  for (int i = threadIdx.x; i < MR * 32; i += NW * 32) {
    const int m = i >> 5, l = i & 31, gn = blockIdx.x * 32 + l;
    if (m < M && gn < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += red[w][m][l];
      float* c = C + (size_t)m * N + gn;
      *c = alpha * t + (beta != 0.f ? beta * *c : 0.f);
    }
  }
}

// ------------------------------------------------------------------ BatchNormalization, training mode, (rows, C)
// one thread per channel, rows walked in a fixed order (coalesced across channels)
__global__ void bn_train_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    float* __restrict__ mov_mean, float* __restrict__ mov_var, float* __restrict__ y,
                                    float* __restrict__ save_mean, float* __restrict__ save_invstd, int rows, int C, float eps,
                                    float momentum) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += x[(size_t)r * C + c];
  const float mean = s / rows;
  float q = 0.f;
  for (int r = 0; r < rows; ++r) { const float d = x[(size_t)r * C + c] - mean; q = fmaf(d, d, q); }
  const float var = q / rows;                                   // biased (Keras' non-fused path, also for the moving average)
  const float inv = 1.0f / sqrtf(var + eps);
  const float g = gamma[c], b = beta[c];
  for (int r = 0; r < rows; ++r) y[(size_t)r * C + c] = (x[(size_t)r * C + c] - mean) * inv * g + b;
  save_mean[c] = mean; save_invstd[c] = inv;
  if (mov_mean) mov_mean[c] = momentum * mov_mean[c] + (1.f - momentum) * mean;
  if (mov_var) mov_var[c] = momentum * mov_var[c] + (1.f - momentum) * var;
}
__global__ void bn_train_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                                    const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                                    float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int C) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = save_mean[c], inv = save_invstd[c], g = gamma[c];
  float sb = 0.f, sg = 0.f;
  for (int r = 0; r < rows; ++r) {
    const float d = dy[(size_t)r * C + c];
    sb += d;
    sg = fmaf(d, (x[(size_t)r * C + c] - mean) * inv, sg);
  }
  dbeta[c] = sb; dgamma[c] = sg;
  if (dx) {
    const float ib = sb / rows, ig = sg / rows;
    for (int r = 0; r < rows; ++r) {
      const float xh = (x[(size_t)r * C + c] - mean) * inv;
      dx[(size_t)r * C + c] = g * inv * (dy[(size_t)r * C + c] - ib - xh * ig);
    }
  }
}

__global__ void bias_act_kernel(const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ y, long long n, int C, int relu) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i] + (b ? b[i % C] : 0.f);
    y[i] = relu == 1 ? fmaxf(v, 0.f) : (relu == 2 ? tanhf(v) : v);
  }
}
__global__ void relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ h, float* __restrict__ out, long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = h[i] > 0.f ? g[i] : 0.f;
}
// column sums of g (rows, C): the bias gradients
__global__ void colsum_kernel(const float* __restrict__ g, float* __restrict__ out, int rows, int C) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += g[(size_t)r * C + c];
  out[c] = s;
}

// ------------------------------------------------------------------ row-parallel column reductions (maps with many rows)
// The one-thread-per-channel kernels above walk `rows` serially: fine for the head (rows = batch), 65 % of a whole-model
// training step once the ResNet's BatchNormalizations (rows = B*H*W up to 10^5..10^6) use them.  Here the rows are split into
// gridDim.y chunks; a block = 32 channels x 8 row lanes; partial sums go to part[chunk][c] and are added in chunk order by
// col_final_kernel -- still a fixed summation order, independent of scheduling.
//   mode 0: s1 = sum a            mode 1: s1 = sum (a - mean)^2            mode 2: s1 = sum b, s2 = sum b * (a - mean) * inv
__global__ void __launch_bounds__(256) col_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mean,
                                                          const float* __restrict__ inv, float* __restrict__ part1, float* __restrict__ part2,
                                                          int rows, int C, int mode) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh1[8][33], sh2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    const float m = mode ? mean[c] : 0.f, iv = mode == 2 ? inv[c] : 0.f;
    for (int r = r0 + ty; r < r1; r += 8) {
      const float v = a[(size_t)r * C + c];
      if (mode == 0) s1 += v;
      else if (mode == 1) { const float d = v - m; s1 = fmaf(d, d, s1); }
      else { const float d = b[(size_t)r * C + c]; s1 += d; s2 = fmaf(d, (v - m) * iv, s2); }
    }
  }
  sh1[ty][tx] = s1; sh2[ty][tx] = s2;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { t1 += sh1[k][tx]; t2 += sh2[k][tx]; }
    part1[(size_t)blockIdx.y * C + c] = t1;
    if (mode == 2) part2[(size_t)blockIdx.y * C + c] = t2;
  }
}
__global__ void col_final_kernel(const float* __restrict__ part, int nch, int C, float scale, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int k = 0; k < nch; ++k) s += part[(size_t)k * C + c];
  out[c] = s * scale;
}
__global__ void bn_stats_finish_kernel(const float* __restrict__ mean, const float* __restrict__ var, float* __restrict__ save_invstd,
                                       float* __restrict__ mov_mean, float* __restrict__ mov_var, int C, float eps, float momentum) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  save_invstd[c] = 1.0f / sqrtf(var[c] + eps);
  if (mov_mean) mov_mean[c] = momentum * mov_mean[c] + (1.f - momentum) * mean[c];
  if (mov_var) mov_var[c] = momentum * mov_var[c] + (1.f - momentum) * var[c];
}
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ inv,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y, long long n, int C) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    y[i] = (x[i] - mean[c]) * inv[c] * gamma[c] + beta[c];
  }
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                    const float* __restrict__ inv, const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, float* __restrict__ dx, long long n, int C, float inv_rows) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float xh = (x[i] - mean[c]) * inv[c];
    dx[i] = gamma[c] * inv[c] * (dy[i] - dbeta[c] * inv_rows - xh * dgamma[c] * inv_rows);
  }
}

// ------------------------------------------------------------------ K.l2_normalize (eps 1e-12 under the root) and backward
// axis 1: every row of v (rows, D); axis 0: every column.  One thread per vector, fixed order.
__global__ void l2norm_fwd_kernel(const float* __restrict__ v, float* __restrict__ out, float* __restrict__ inv_out, int rows, int D, int axis) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nvec = axis ? rows : D, len = axis ? D : rows;
  if (i >= nvec) return;
  const size_t base = axis ? (size_t)i * D : (size_t)i, stride = axis ? 1 : (size_t)D;
  float q = 0.f;
  for (int j = 0; j < len; ++j) { const float a = v[base + j * stride]; q = fmaf(a, a, q); }
  const float inv = 1.0f / sqrtf(fmaxf(q, 1e-12f));
  for (int j = 0; j < len; ++j) out[base + j * stride] = v[base + j * stride] * inv;
  inv_out[i] = inv;
}
// d loss / d v = (u - vhat (vhat . u)) * inv      (u = d loss / d vhat)
__global__ void l2norm_bwd_kernel(const float* __restrict__ vhat, const float* __restrict__ inv, const float* __restrict__ u,
                                  float* __restrict__ out, int rows, int D, int axis, float beta) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nvec = axis ? rows : D, len = axis ? D : rows;
  if (i >= nvec) return;
  const size_t base = axis ? (size_t)i * D : (size_t)i, stride = axis ? 1 : (size_t)D;
  float dot = 0.f;
  for (int j = 0; j < len; ++j) dot = fmaf(vhat[base + j * stride], u[base + j * stride], dot);
  const float s = inv[i];
  for (int j = 0; j < len; ++j) {
    const size_t o = base + j * stride;
    const float g = (u[o] - vhat[o] * dot) * s;
    out[o] = beta != 0.f ? beta * out[o] + g : g;
  }
}

// ------------------------------------------------------------------ losses and their gradients w.r.t. logits / cosines
// One thread per utterance (n <= 32 classes).  head: sar head selector (SAR_HEAD_*).
//   g_acc (B,n) = w_acc / B * d CE(softmax(z_acc)) / d z_acc
//   g_disc (B,n) = w_disc / B * d loss_disc / d (cos for the Face / Circle heads, logits for the softmax head)
//   losses (B,2) per-sample [CE_accent, loss_disc]
// Keras' categorical_crossentropy clips p to [1e-7, 1 - 1e-7]: outside that range the gradient is zero.
constexpr float TK_EPS = 1e-7f;
__global__ void head_grad_kernel(const float* __restrict__ z_acc, const float* __restrict__ c_disc, const float* __restrict__ onehot,
                                 int n, int head, float margin, float s, float gamma, float w_acc, float w_disc,
                                 float* __restrict__ g_acc, float* __restrict__ g_disc, float* __restrict__ losses, int B) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* y = onehot + (size_t)b * n;
  int lab = 0;
  for (int j = 1; j < n; ++j) if (y[j] > y[lab]) lab = j;
  float z[32], p[32];
  auto softmax_ce = [&](float* zz, float& loss, bool& clipped) {      // p <- softmax(zz)
    float m = -INFINITY;
    for (int j = 0; j < n; ++j) m = fmaxf(m, zz[j]);
    float sum = 0.f;
    for (int j = 0; j < n; ++j) { p[j] = expf(zz[j] - m); sum += p[j]; }
    for (int j = 0; j < n; ++j) p[j] /= sum;
    const float pl = p[lab];
    clipped = pl < TK_EPS || pl > 1.f - TK_EPS;
    loss = -logf(fminf(fmaxf(pl, TK_EPS), 1.f - TK_EPS));
  };
  if (z_acc) {
    for (int j = 0; j < n; ++j) z[j] = z_acc[(size_t)b * n + j];
    float loss; bool clipped;
    softmax_ce(z, loss, clipped);
    for (int j = 0; j < n; ++j) g_acc[(size_t)b * n + j] = clipped ? 0.f : (p[j] - (j == lab ? 1.f : 0.f)) * (w_acc / B);
    losses[2 * b] = loss;
  }
  if (c_disc && head != SAR_HEAD_NONE) {
    float c[32], dz[32];                                        // dz[j] = d logit_j / d c_j
    for (int j = 0; j < n; ++j) { c[j] = c_disc[(size_t)b * n + j]; z[j] = c[j]; dz[j] = 1.f; }
    if (head == SAR_HEAD_SPHEREFACE || head == SAR_HEAD_COSFACE || head == SAR_HEAD_ARCFACE) {
      for (int j = 0; j < n; ++j) { z[j] = s * c[j]; dz[j] = s; }
      const float cl = c[lab];
      if (head == SAR_HEAD_COSFACE) {
        z[lab] = s * (cl - margin);                                                  // losses.py:84
      } else {
        const float cc = fminf(fmaxf(cl, -1.f + TK_EPS), 1.f - TK_EPS);
        const bool inside = cl > -1.f + TK_EPS && cl < 1.f - TK_EPS;                 // K.clip: zero gradient outside
        const float th = acosf(cc), sn = sqrtf(fmaxf(1.f - cc * cc, 1e-30f));
        if (head == SAR_HEAD_ARCFACE) { z[lab] = s * cosf(th + margin); dz[lab] = inside ? s * sinf(th + margin) / sn : 0.f; }
        else { z[lab] = s * cosf(margin * th); dz[lab] = inside ? s * margin * sinf(margin * th) / sn : 0.f; }
      }
    } else if (head == SAR_HEAD_CIRCLE || head == SAR_HEAD_CIRCLE_RAW) {             // losses.py:157-172
      for (int j = 0; j < n; ++j) {
        if (j == lab) { const float ap = fmaxf(1.f + margin - c[j], 0.f); z[j] = gamma * ap * (c[j] - (1.f - margin)); dz[j] = ap > 0.f ? gamma * (2.f - 2.f * c[j]) : 0.f; }
        else { const float an = fmaxf(c[j] + margin, 0.f); z[j] = gamma * an * (c[j] - margin); dz[j] = an > 0.f ? gamma * 2.f * c[j] : 0.f; }
      }
    }
    float loss; bool clipped;
    softmax_ce(z, loss, clipped);
    if (head == SAR_HEAD_CIRCLE || head == SAR_HEAD_CIRCLE_RAW) {                    // -sum y log_softmax: no clip
      clipped = false;
      float m = -INFINITY;
      for (int j = 0; j < n; ++j) m = fmaxf(m, z[j]);
      float sum = 0.f;
      for (int j = 0; j < n; ++j) sum += expf(z[j] - m);
      loss = -(z[lab] - m - logf(sum));
    }
    for (int j = 0; j < n; ++j) g_disc[(size_t)b * n + j] = clipped ? 0.f : (p[j] - (j == lab ? 1.f : 0.f)) * dz[j] * (w_disc / B);
    losses[2 * b + 1] = loss;
  }
}

// ------------------------------------------------------------------ Adam (Keras 2.2.4) + l2 regulariser + unit_norm
// g_eff = g + 2 * l2 * p ; m, v updated in place ; p -= lr_t * m / (sqrt(v) + eps)   (lr_t computed by the caller)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_t, float b1, float b2, float eps, float l2) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float ge = fmaf(2.f * l2, p[i], g[i]);
    const float mi = b1 * m[i] + (1.f - b1) * ge;
    const float vi = b2 * v[i] + (1.f - b2) * ge * ge;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
// the same with lr_t read from device memory: a captured CUDA graph of the training step is replayed with a new step size
// (Adam's bias correction and Keras' decay change lr_t every iteration; a kernel ARGUMENT would be frozen into the graph)
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, const float* __restrict__ lr_t_p, float b1, float b2, float eps, float l2) {
  pdl_wait();
  pdl_trigger();
  const float lr_t = *lr_t_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float ge = fmaf(2.f * l2, p[i], g[i]);
    const float mi = b1 * m[i] + (1.f - b1) * ge;
    const float vi = b2 * v[i] + (1.f - b2) * ge * ge;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
// keras.constraints.unit_norm(axis=0): W[:, j] /= (1e-7 + ||W[:, j]||), W (D, n)
__global__ void unit_norm_kernel(float* __restrict__ w, int D, int n) {
  pdl_wait();
  pdl_trigger();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float q = 0.f;
  for (int d = 0; d < D; ++d) q = fmaf(w[(size_t)d * n + j], w[(size_t)d * n + j], q);
  const float inv = 1.f / (1e-7f + sqrtf(q));
  for (int d = 0; d < D; ++d) w[(size_t)d * n + j] *= inv;
}

static unsigned blocks_for(long long n, int bs, int cap = 4096) {
  long long b = (n + bs - 1) / bs;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------ NetVLAD / GhostVLAD in training mode (model.py:82-109, VLAD.py:26-49)
// One CTA per utterance, everything fp32 on CUDA cores in fixed reduction orders (S <= 128 positions, K+G <= 128).
//   forward : scores = x Wa + ba -> A = softmax over the K+G clusters (kept for the backward), asum[k] = sum_s A[s,k],
//             R[k,:] = sum_s A[s,k] x[s,:] - asum[k] c[k,:] for the K real clusters (ghost rows are dropped, VLAD.py:44-45).
//             The per-cluster K.l2_normalize (VLAD.py:47) is sar_l2norm_fwd on R viewed as (B*K, D).
//   backward: gR = d loss / d R (from sar_l2norm_bwd).  gA[s,k] = gR[k,:] . (x[s,:] - c[k,:]) (0 for ghosts),
//             g_scores = A * (gA - sum_k' A gA) (softmax), gc_part[b,k,:] = -asum[b,k] gR[b,k,:] (summed over b by
//             sar_colsum_fwd); g_Wa = x^T g_scores and g_ba = colsum(g_scores) are the caller's sar_gemm_fwd / sar_colsum_fwd.
__global__ void __launch_bounds__(256) vlad_train_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wa, const float* __restrict__ ba,
                                                              const float* __restrict__ centers, float* __restrict__ A_out,
                                                              float* __restrict__ R_out, float* __restrict__ asum_out, int S, int D, int K, int KG) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float vsm[];
  float* As = vsm;                 // [S][KG]
  float* asum = vsm + S * KG;      // [KG]
  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float* xb = x + (size_t)b * S * D;
  for (int i = t; i < S * KG; i += 256) {                  // scores: consecutive threads -> consecutive clusters (Wa rows coalesced)
    const int s_ = i / KG, k = i - s_ * KG;
    float acc = ba[k];
    for (int d = 0; d < D; ++d) acc = fmaf(xb[(size_t)s_ * D + d], wa[(size_t)d * KG + k], acc);
    As[i] = acc;
  }
  __syncthreads();
  for (int s_ = warp; s_ < S; s_ += 8) {                   // softmax of one row per warp (VLAD.py:33-35)
    float mx = -INFINITY;
    for (int k = lane; k < KG; k += 32) mx = fmaxf(mx, As[s_ * KG + k]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int k = lane; k < KG; k += 32) { const float e = expf(As[s_ * KG + k] - mx); As[s_ * KG + k] = e; sum += e; }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int k = lane; k < KG; k += 32) {
      const float a = As[s_ * KG + k] * inv;
      As[s_ * KG + k] = a;
      A_out[((size_t)b * S + s_) * KG + k] = a;
    }
  }
  __syncthreads();
  for (int k = t; k < KG; k += 256) {
    float a = 0.f;
    for (int s_ = 0; s_ < S; ++s_) a += As[s_ * KG + k];
    asum[k] = a;
    if (k < K) asum_out[(size_t)b * K + k] = a;
  }
  __syncthreads();
  for (int i = t; i < K * D; i += 256) {                   // residual sums of the K real clusters
    const int k = i / D, d = i - k * D;
    float acc = 0.f;
    for (int s_ = 0; s_ < S; ++s_) acc = fmaf(As[s_ * KG + k], xb[(size_t)s_ * D + d], acc);
    R_out[((size_t)b * K + k) * D + d] = acc - asum[k] * centers[(size_t)k * D + d];
  }
}

__global__ void __launch_bounds__(256) vlad_train_bwd_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ centers,
                                                              const float* __restrict__ gR, const float* __restrict__ asum,
                                                              float* __restrict__ g_scores, float* __restrict__ gc_part, float* __restrict__ g_x,
                                                              int S, int D, int K, int KG) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float vsm[];
  float* gA = vsm;                 // [S][K]
  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float* xb = x + (size_t)b * S * D;
  const float* gRb = gR + (size_t)b * K * D;
  for (int i = warp; i < S * K; i += 8) {                  // one (position, cluster) dot product per warp pass
    const int s_ = i / K, k = i - s_ * K;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(gRb[(size_t)k * D + d], xb[(size_t)s_ * D + d] - centers[(size_t)k * D + d], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) gA[i] = acc;
  }
  __syncthreads();
  for (int s_ = warp; s_ < S; s_ += 8) {                   // softmax backward of one row per warp
    const float* Ar = A + ((size_t)b * S + s_) * KG;
    float dot = 0.f;
    for (int k = lane; k < K; k += 32) dot = fmaf(Ar[k], gA[s_ * K + k], dot);
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int k = lane; k < KG; k += 32)
      g_scores[((size_t)b * S + s_) * KG + k] = Ar[k] * ((k < K ? gA[s_ * K + k] : 0.f) - dot);
  }
  for (int i = t; i < K * D; i += 256)
    gc_part[(size_t)b * K * D + i] = -asum[(size_t)b * K + i / D] * gRb[i];
  // the residual-sum path of d loss / d x: g_x[s,:] = sum_k A[s,k] gR[k,:]  (the score path, g_scores Wa^T, is the
  // caller's sar_gemm_fwd with beta = 1)
  if (g_x)
    for (int i = t; i < S * D; i += 256) {
      const int s_ = i / D, d = i - s_ * D;
      const float* Ar = A + ((size_t)b * S + s_) * KG;
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc = fmaf(Ar[k], gRb[(size_t)k * D + d], acc);
      g_x[((size_t)b * S + s_) * D + d] = acc;
    }
}

// ------------------------------------------------------------------ LayerNormalization (+ the tanh in front of it) backward
// Row per warp.  y (rows, C) is the LN INPUT (= tanh(pre) when `tanh_in`), g_z = d loss / d LN output.
//   xhat = (y - mean) * inv,  inv = 1 / sqrt(var + eps) (biased variance, eps = 1e-14: keras_layer_normalization)
//   g_y = inv * (gamma g_z - mean(gamma g_z) - xhat * mean(gamma g_z xhat));   g_pre = g_y * (1 - y^2) if tanh_in
// also writes gz_xhat = g_z * xhat (column sums = d loss / d gamma; d loss / d beta = column sums of g_z).
__global__ void __launch_bounds__(256) ln_train_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ g_z,
                                                            float* __restrict__ g_pre, float* __restrict__ gz_xhat, int rows, int C, float eps,
                                                            int tanh_in) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const float* yr = y + (size_t)r * C;
  const float* gr = g_z + (size_t)r * C;
  float s1 = 0.f;
  for (int c = lane; c < C; c += 32) s1 += yr[c];
  for (int o = 16; o; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  const float mean = s1 / C;
  float s2 = 0.f;
  for (int c = lane; c < C; c += 32) { const float dlt = yr[c] - mean; s2 = fmaf(dlt, dlt, s2); }
  for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  const float inv = rsqrtf(s2 / C + eps);
  float a1 = 0.f, a2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float xh = (yr[c] - mean) * inv, gg = gamma[c] * gr[c];
    a1 += gg; a2 = fmaf(gg, xh, a2);
  }
  for (int o = 16; o; o >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
  a1 /= C; a2 /= C;
  for (int c = lane; c < C; c += 32) {
    const float yv = yr[c], xh = (yv - mean) * inv;
    float gy = inv * (gamma[c] * gr[c] - a1 - xh * a2);
    if (tanh_in) gy *= 1.f - yv * yv;
    g_pre[(size_t)r * C + c] = gy;
    gz_xhat[(size_t)r * C + c] = gr[c] * xh;
  }
}

// ------------------------------------------------------------------ CuDNNGRU in training mode (fourth slice): one time step
// Gate arithmetic of one step of one direction (model.py:44-50; reset_after form, gate order z|r|h):
//   a = xp[b, t, :] (= x W + b_i) and hu = h_prev U (+ b_r added here);  z = s(a_z + hu_z), r = s(a_r + hu_r),
//   hh = tanh(a_h + r * hu_h),  h' = z h_prev + (1 - z) hh.  The step's z, r, hh and hu_h (+ bias) are kept for the backward.
__global__ void gru_gate_fwd_kernel(const float* __restrict__ xp, const float* __restrict__ hu, const float* __restrict__ b_r,
                                    const float* __restrict__ h_prev, float* __restrict__ z_s, float* __restrict__ r_s,
                                    float* __restrict__ hh_s, float* __restrict__ hph_s, float* __restrict__ h_new,
                                    float* __restrict__ out, int B, int S, int u, int t, int out_stride, int out_off) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)B * u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / u), j = (int)(i - (long long)b * u);
    const float* a = xp + ((size_t)b * S + t) * 3 * u;
    const float* hb = hu + (size_t)b * 3 * u;
    const float z = sigmoidf_(a[j] + hb[j] + b_r[j]);
    const float r = sigmoidf_(a[u + j] + hb[u + j] + b_r[u + j]);
    const float hph = hb[2 * u + j] + b_r[2 * u + j];
    const float hh = tanhf(a[2 * u + j] + r * hph);
    const float hp = h_prev[i];
    const float hn = z * hp + (1.f - z) * hh;
    z_s[i] = z; r_s[i] = r; hh_s[i] = hh; hph_s[i] = hph; h_new[i] = hn;
    if (out) out[((size_t)b * S + t) * out_stride + out_off + j] = hn;
  }
}
// ... and its backward: dh = g_out[b, t, off + j] + dh_rec[b, j] is d loss / d h'.  Writes d loss / d a into d_xp[b, t, :]
// (the input-projection pre-activations), d loss / d hu into d_hu (B, 3u), and dh * z into dh_prev (the caller adds
// d_hu U^T with one GEMM).
__global__ void gru_gate_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ dh_rec, const float* __restrict__ z_s,
                                    const float* __restrict__ r_s, const float* __restrict__ hh_s, const float* __restrict__ hph_s,
                                    const float* __restrict__ h_prev, float* __restrict__ d_xp, float* __restrict__ d_hu,
                                    float* __restrict__ dh_prev, int B, int S, int u, int t, int out_stride, int out_off) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)B * u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / u), j = (int)(i - (long long)b * u);
    float dh = dh_rec ? dh_rec[i] : 0.f;
    if (g_out) dh += g_out[((size_t)b * S + t) * out_stride + out_off + j];
    const float z = z_s[i], r = r_s[i], hh = hh_s[i], hph = hph_s[i], hp = h_prev[i];
    const float d_ah = dh * (1.f - z) * (1.f - hh * hh);          // through tanh
    const float d_az = dh * (hp - hh) * z * (1.f - z);            // through the update gate
    const float d_ar = d_ah * hph * r * (1.f - r);                // through the reset gate
    float* da = d_xp + ((size_t)b * S + t) * 3 * u;
    float* dhu = d_hu + (size_t)b * 3 * u;
    da[j] = d_az; da[u + j] = d_ar; da[2 * u + j] = d_ah;
    dhu[j] = d_az; dhu[u + j] = d_ar; dhu[2 * u + j] = d_ah * r;
    dh_prev[i] = dh * z;
  }
}

// ------------------------------------------------------------------ convolution / pooling backward (sixth slice: the ResNet)
// Correctness-first fp32 kernels for NHWC maps with HWIO kernels and TF-SAME leading pads (pad_t, pad_l), any stride.
// d loss / d x[b,h,w,ci] = sum_{kh,kw,co} dy[b,ho,wo,co] w[kh,kw,ci,co] over the (kh,kw) with h = ho*stride + kh - pad_t (same in w).
// One thread per dx element, the taps and output channels walked in a fixed order.
__global__ void conv2d_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int H, int W,
                                       int Cin, int Ho, int Wo, int Cout, int kh, int kw, int stride, int pad_t, int pad_l, float beta) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)B * H * W * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    long long r = i / Cin;
    const int x_ = (int)(r % W); r /= W;
    const int y_ = (int)(r % H);
    const int b = (int)(r / H);
    float acc = 0.f;
    for (int a = 0; a < kh; ++a) {
      const int hn = y_ + pad_t - a;
      if (hn < 0 || hn % stride) continue;
      const int ho = hn / stride;
      if (ho >= Ho) continue;
      for (int c = 0; c < kw; ++c) {
        const int wn = x_ + pad_l - c;
        if (wn < 0 || wn % stride) continue;
        const int wo = wn / stride;
        if (wo >= Wo) continue;
        const float* g = dy + (((size_t)b * Ho + ho) * Wo + wo) * Cout;
        const float* wr = w + ((size_t)(a * kw + c) * Cin + ci) * Cout;
        if ((Cout & 3) == 0) {                       // rows of dy and of w[kh,kw,ci,:] are 16-byte aligned
          const float4* g4 = reinterpret_cast<const float4*>(g);
          const float4* w4 = reinterpret_cast<const float4*>(wr);
          for (int co = 0; co < Cout / 4; ++co) {
            const float4 gv = __ldg(g4 + co), wv = __ldg(w4 + co);
            acc = fmaf(gv.x, wv.x, acc); acc = fmaf(gv.y, wv.y, acc); acc = fmaf(gv.z, wv.z, acc); acc = fmaf(gv.w, wv.w, acc);
          }
        } else {
          for (int co = 0; co < Cout; ++co) acc = fmaf(__ldg(g + co), __ldg(wr + co), acc);
        }
      }
    }
    dx[i] = acc + (beta != 0.f ? beta * dx[i] : 0.f);
  }
}
// d loss / d w[kh,kw,ci,co] = sum_{b,ho,wo} x[b, ho*stride+kh-pad_t, wo*stride+kw-pad_l, ci] dy[b,ho,wo,co]: a GEMM per tap,
// [Cin x positions] x [positions x Cout].  Block = one tap, a 32 x 32 (ci, co) tile and one of gridDim.y position chunks;
// 32 positions at a time are staged in shared memory (the shifted x rows, zero outside the map), thread (ty, tx) owns a
// 2 x 2 patch of the tile.  Chunk z writes its partial sums to part[z][kh*kw*Cin*Cout] (summed by sar_colsum_fwd: fixed order).
__global__ void __launch_bounds__(256) conv2d_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part,
                                                                int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int kh, int kw,
                                                                int stride, int pad_t, int pad_l) {
  pdl_wait();
  pdl_trigger();
  __shared__ float xs[32][33], ds[32][33];
  const int ct_o = (Cout + 31) / 32, ct_i = (Cin + 31) / 32;
  int t = blockIdx.x;
  const int co0 = (t % ct_o) * 32; t /= ct_o;
  const int ci0 = (t % ct_i) * 32; t /= ct_i;
  const int c = t % kw, a = t / kw;
  const long long nw = (long long)kh * kw * Cin * Cout;
  const long long npos = (long long)B * Ho * Wo;
  const long long per = (npos + gridDim.y - 1) / gridDim.y;
  const long long p0 = blockIdx.y * per, p1 = (p0 + per < npos) ? p0 + per : npos;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 5, lc = threadIdx.x & 31;       // loader role: row lr (+8k) of the 32-position slab, channel lc
  float acc[2][2] = {};
  for (long long q0 = p0; q0 < p1; q0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = lr + 8 * k;
      const long long q = q0 + r;
      float xv = 0.f, dv = 0.f;
      if (q < p1) {
        const int wo = (int)(q % Wo);
        const long long u = q / Wo;
        const int ho = (int)(u % Ho);
        const int b = (int)(u / Ho);
        const int hi = ho * stride + a - pad_t, wi = wo * stride + c - pad_l;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W && ci0 + lc < Cin) xv = __ldg(x + (((size_t)b * H + hi) * W + wi) * Cin + ci0 + lc);
        if (co0 + lc < Cout) dv = __ldg(dy + (size_t)q * Cout + co0 + lc);
      }
      xs[r][lc] = xv; ds[r][lc] = dv;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float x0 = xs[r][2 * ty], x1 = xs[r][2 * ty + 1], d0 = ds[r][2 * tx], d1 = ds[r][2 * tx + 1];
      acc[0][0] = fmaf(x0, d0, acc[0][0]); acc[0][1] = fmaf(x0, d1, acc[0][1]);
      acc[1][0] = fmaf(x1, d0, acc[1][0]); acc[1][1] = fmaf(x1, d1, acc[1][1]);
    }
    __syncthreads();
  }
  float* out = part + (size_t)blockIdx.y * nw + (size_t)(a * kw + c) * Cin * Cout;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ci = ci0 + 2 * ty + i, co = co0 + 2 * tx + j;
      if (ci < Cin && co < Cout) out[(size_t)ci * Cout + co] = acc[i][j];
    }
}
// The same for layers with >= 64 input and output channels: a 64 x 64 (ci, co) tile, 16 positions per slab, thread (ty, tx) owns
// a 4 x 4 patch -- 16 FMAs per 8 shared-memory loads instead of 4 per 4.
__global__ void __launch_bounds__(256) conv2d_bwd_weight64_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part,
                                                                  int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int kh, int kw,
                                                                  int stride, int pad_t, int pad_l) {
  pdl_wait();
  pdl_trigger();
  __shared__ float xs[16][68], ds[16][68];
  const int ct_o = (Cout + 63) / 64, ct_i = (Cin + 63) / 64;
  int t = blockIdx.x;
  const int co0 = (t % ct_o) * 64; t /= ct_o;
  const int ci0 = (t % ct_i) * 64; t /= ct_i;
  const int c = t % kw, a = t / kw;
  const long long nw = (long long)kh * kw * Cin * Cout;
  const long long npos = (long long)B * Ho * Wo;
  const long long per = (npos + gridDim.y - 1) / gridDim.y;
  const long long p0 = blockIdx.y * per, p1 = (p0 + per < npos) ? p0 + per : npos;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 6, lc = threadIdx.x & 63;       // loader role: row lr (+4k) of the 16-position slab, channel lc
  float acc[4][4] = {};
  for (long long q0 = p0; q0 < p1; q0 += 16) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = lr + 4 * k;
      const long long q = q0 + r;
      float xv = 0.f, dv = 0.f;
      if (q < p1) {
        const int wo = (int)(q % Wo);
        const long long u = q / Wo;
        const int ho = (int)(u % Ho);
        const int b = (int)(u / Ho);
        const int hi = ho * stride + a - pad_t, wi = wo * stride + c - pad_l;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W && ci0 + lc < Cin) xv = __ldg(x + (((size_t)b * H + hi) * W + wi) * Cin + ci0 + lc);
        if (co0 + lc < Cout) dv = __ldg(dy + (size_t)q * Cout + co0 + lc);
      }
      xs[r][lc] = xv; ds[r][lc] = dv;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 16; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(&xs[r][4 * ty]);
      const float4 dv = *reinterpret_cast<const float4*>(&ds[r][4 * tx]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], da[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + (size_t)blockIdx.y * nw + (size_t)(a * kw + c) * Cin * Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + 4 * ty + i, co = co0 + 4 * tx + j;
      if (ci < Cin && co < Cout) out[(size_t)ci * Cout + co] = acc[i][j];
    }
}
// MaxPooling2D backward: one thread per INPUT element; it collects dy of every window whose (first) maximum it is.
__global__ void maxpool2d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W,
                                     int C, int Ho, int Wo, int k, int stride, int pad_t, int pad_l) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)B * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C);
    long long r = i / C;
    const int x_ = (int)(r % W); r /= W;
    const int y_ = (int)(r % H);
    const int b = (int)(r / H);
    const float me = x[i];
    float acc = 0.f;
    // windows covering me: ho with ho*stride - pad_t <= y_ < ho*stride - pad_t + k (same along w)
    const int ho_lo = max(0, (y_ + pad_t - k + stride) / stride), ho_hi = min(Ho - 1, (y_ + pad_t) / stride);
    const int wo_lo = max(0, (x_ + pad_l - k + stride) / stride), wo_hi = min(Wo - 1, (x_ + pad_l) / stride);
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const int h0 = ho * stride - pad_t;
      if (y_ < h0 || y_ >= h0 + k) continue;
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const int w0 = wo * stride - pad_l;
        if (x_ < w0 || x_ >= w0 + k) continue;
        // am I the first maximum of this window (row-major scan, padded cells never win)?
        bool win = true;
        for (int a = 0; a < k && win; ++a) {
          const int hi = h0 + a;
          if (hi < 0 || hi >= H) continue;
          for (int c = 0; c < k; ++c) {
            const int wi = w0 + c;
            if (wi < 0 || wi >= W) continue;
            const float v = __ldg(x + (((size_t)b * H + hi) * W + wi) * C + ch);
            const bool before = (hi < y_) || (hi == y_ && wi < x_);
            if (v > me || (before && v == me)) { win = false; break; }
          }
        }
        if (win) acc += __ldg(dy + (((size_t)b * Ho + ho) * Wo + wo) * C + ch);
      }
    }
    dx[i] = acc;
  }
}
__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float alpha, long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = fmaf(alpha, x[i], y[i]);
}

}  // namespace sar

extern "C" {

int sar_gemm_fwd(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, int trans_b, float alpha, float beta,
                 void* stream) {
  using namespace sar;
  SAR_REQUIRE(A && B && C, SAR_ERR_BAD_ARG, "sar_gemm_fwd: null pointer");
  SAR_REQUIRE(M > 0 && N > 0 && K > 0, SAR_ERR_BAD_ARG, "sar_gemm_fwd: non-positive dimension");
  // skinny: the GRU's per-step products, the embedding Dense at small batches.  (M in (32, 64] has a 4-warp instance, measured
  // SLOWER than the tiled kernel at M = 64 -- 85 vs 49 us per h U product: 64 broadcast loads per k-value -- so it is not dispatched)
  if (!trans_a && M <= 32 && K >= 64) {
    const dim3 g((N + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
    const int tb_ = trans_b ? 1 : 0;
    if (M <= 16) launch_k(gemm_skinny_kernel<16, 8>, g, dim3(256), 0, st, A, B, C, M, N, K, tb_, alpha, beta);
    else if (M <= 32) launch_k(gemm_skinny_kernel<32, 8>, g, dim3(256), 0, st, A, B, C, M, N, K, tb_, alpha, beta);
    else launch_k(gemm_skinny_kernel<64, 4>, g, dim3(128), 0, st, A, B, C, M, N, K, tb_, alpha, beta);
    return check_launch("sar_gemm_fwd(skinny)");
  }
  launch_k(gemm_kernel, dim3((N + TG - 1) / TG, (M + TG - 1) / TG), dim3(256), 0, (cudaStream_t)stream, A, B, C, M, N, K, trans_a ? 1 : 0,
           trans_b ? 1 : 0, alpha, beta);
  return check_launch("sar_gemm_fwd");
}

int sar_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* y,
                     float* save_mean, float* save_invstd, int rows, int C, float eps, float momentum, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && gamma && beta && y && save_mean && save_invstd, SAR_ERR_BAD_ARG, "sar_bn_train_fwd: null pointer");
  SAR_REQUIRE(rows > 0 && C > 0, SAR_ERR_BAD_ARG, "sar_bn_train_fwd: non-positive dimension");
  launch_k(bn_train_fwd_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, x, gamma, beta, moving_mean, moving_var, y,
           save_mean, save_invstd, rows, C, eps, momentum);
  return check_launch("sar_bn_train_fwd");
}

int sar_bn_train_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd,
                     float* dx, float* dgamma, float* dbeta, int rows, int C, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && dy && gamma && save_mean && save_invstd && dgamma && dbeta, SAR_ERR_BAD_ARG, "sar_bn_train_bwd: null pointer");
  SAR_REQUIRE(rows > 0 && C > 0, SAR_ERR_BAD_ARG, "sar_bn_train_bwd: non-positive dimension");
  launch_k(bn_train_bwd_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, x, dy, gamma, save_mean, save_invstd, dx,
           dgamma, dbeta, rows, C);
  return check_launch("sar_bn_train_bwd");
}

// Row-parallel forms of sar_colsum_fwd / sar_bn_train_fwd / sar_bn_train_bwd for maps with many rows (the ResNet in training
// mode): `nch` row chunks, `ws` = caller-owned scratch of (2 * nch + 2) * C floats.  Same results up to fp32 summation order
// (fixed, chunk by chunk).
int sar_colsum_rows_fwd(const float* g, float* out, int rows, int C, int nch, float* ws, void* stream) {
  using namespace sar;
  SAR_REQUIRE(g && out && ws && rows > 0 && C > 0 && nch > 0 && nch <= 65535, SAR_ERR_BAD_ARG, "sar_colsum_rows_fwd: bad argument");
  launch_k(col_partial_kernel, dim3((C + 31) / 32, nch), dim3(256), 0, (cudaStream_t)stream, g, (const float*)nullptr, (const float*)nullptr,
           (const float*)nullptr, ws, (float*)nullptr, rows, C, 0);
  launch_k(col_final_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, (const float*)ws, nch, C, 1.0f, out);
  return check_launch("sar_colsum_rows_fwd");
}

int sar_bn_train_rows_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* y,
                          float* save_mean, float* save_invstd, int rows, int C, float eps, float momentum, int nch, float* ws,
                          void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && gamma && beta && y && save_mean && save_invstd && ws, SAR_ERR_BAD_ARG, "sar_bn_train_rows_fwd: null pointer");
  SAR_REQUIRE(rows > 0 && C > 0 && nch > 0 && nch <= 65535, SAR_ERR_BAD_ARG, "sar_bn_train_rows_fwd: bad dimension");
  cudaStream_t st = (cudaStream_t)stream;
  float* var = ws + (size_t)2 * nch * C;
  const dim3 gp((C + 31) / 32, nch), gc((C + 127) / 128);
  launch_k(col_partial_kernel, gp, dim3(256), 0, st, x, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, ws, (float*)nullptr, rows, C, 0);
  launch_k(col_final_kernel, gc, dim3(128), 0, st, (const float*)ws, nch, C, 1.0f / rows, save_mean);
  launch_k(col_partial_kernel, gp, dim3(256), 0, st, x, (const float*)nullptr, (const float*)save_mean, (const float*)nullptr, ws, (float*)nullptr, rows, C, 1);
  launch_k(col_final_kernel, gc, dim3(128), 0, st, (const float*)ws, nch, C, 1.0f / rows, var);
  launch_k(bn_stats_finish_kernel, gc, dim3(128), 0, st, (const float*)save_mean, (const float*)var, save_invstd, moving_mean, moving_var, C, eps, momentum);
  launch_k(bn_apply_kernel, dim3(blocks_for((long long)rows * C, 256, 1 << 16)), dim3(256), 0, st, x, (const float*)save_mean, (const float*)save_invstd,
           gamma, beta, y, (long long)rows * C, C);
  return check_launch("sar_bn_train_rows_fwd");
}

int sar_bn_train_rows_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd,
                          float* dx, float* dgamma, float* dbeta, int rows, int C, int nch, float* ws, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && dy && gamma && save_mean && save_invstd && dgamma && dbeta && ws, SAR_ERR_BAD_ARG, "sar_bn_train_rows_bwd: null pointer");
  SAR_REQUIRE(rows > 0 && C > 0 && nch > 0 && nch <= 65535, SAR_ERR_BAD_ARG, "sar_bn_train_rows_bwd: bad dimension");
  cudaStream_t st = (cudaStream_t)stream;
  float* p2 = ws + (size_t)nch * C;
  const dim3 gp((C + 31) / 32, nch), gc((C + 127) / 128);
  launch_k(col_partial_kernel, gp, dim3(256), 0, st, x, dy, save_mean, save_invstd, ws, p2, rows, C, 2);
  launch_k(col_final_kernel, gc, dim3(128), 0, st, (const float*)ws, nch, C, 1.0f, dbeta);
  launch_k(col_final_kernel, gc, dim3(128), 0, st, (const float*)p2, nch, C, 1.0f, dgamma);
  if (dx)
    launch_k(bn_bwd_apply_kernel, dim3(blocks_for((long long)rows * C, 256, 1 << 16)), dim3(256), 0, st, x, dy, save_mean, save_invstd, gamma,
             (const float*)dgamma, (const float*)dbeta, dx, (long long)rows * C, C, 1.0f / rows);
  return check_launch("sar_bn_train_rows_bwd");
}

int sar_bias_act_fwd(const float* x, const float* bias, float* y, long long rows, int C, int act, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && y, SAR_ERR_BAD_ARG, "sar_bias_act_fwd: null pointer");
  SAR_REQUIRE(rows > 0 && C > 0 && (act == SAR_ACT_NONE || act == SAR_ACT_RELU || act == SAR_ACT_TANH), SAR_ERR_BAD_ARG, "sar_bias_act_fwd: bad argument");
  launch_k(bias_act_kernel, dim3(blocks_for(rows * C, 256)), dim3(256), 0, (cudaStream_t)stream, x, bias, y, rows * C, C,
           act == SAR_ACT_RELU ? 1 : (act == SAR_ACT_TANH ? 2 : 0));
  return check_launch("sar_bias_act_fwd");
}

int sar_relu_bwd(const float* g, const float* h, float* out, long long n, void* stream) {
  using namespace sar;
  SAR_REQUIRE(g && h && out && n > 0, SAR_ERR_BAD_ARG, "sar_relu_bwd: bad argument");
  launch_k(relu_bwd_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, g, h, out, n);
  return check_launch("sar_relu_bwd");
}

int sar_gru_gate_fwd(const float* xp, const float* hu, const float* b_r, const float* h_prev, float* z, float* r, float* hh,
                     float* hph, float* h_new, float* out, int B, int S, int u, int t, int out_stride, int out_off, void* stream) {
  using namespace sar;
  SAR_REQUIRE(xp && hu && b_r && h_prev && z && r && hh && hph && h_new, SAR_ERR_BAD_ARG, "sar_gru_gate_fwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0 && u > 0 && t >= 0 && t < S && (!out || (out_off >= 0 && out_off + u <= out_stride)), SAR_ERR_BAD_ARG,
              "sar_gru_gate_fwd: bad argument");
  launch_k(gru_gate_fwd_kernel, dim3(blocks_for((long long)B * u, 256)), dim3(256), 0, (cudaStream_t)stream, xp, hu, b_r, h_prev, z, r,
           hh, hph, h_new, out, B, S, u, t, out_stride, out_off);
  return check_launch("sar_gru_gate_fwd");
}

int sar_gru_gate_bwd(const float* g_out, const float* dh_rec, const float* z, const float* r, const float* hh, const float* hph,
                     const float* h_prev, float* d_xp, float* d_hu, float* dh_prev, int B, int S, int u, int t, int out_stride,
                     int out_off, void* stream) {
  using namespace sar;
  SAR_REQUIRE(z && r && hh && hph && h_prev && d_xp && d_hu && dh_prev && (g_out || dh_rec), SAR_ERR_BAD_ARG, "sar_gru_gate_bwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0 && u > 0 && t >= 0 && t < S && (!g_out || (out_off >= 0 && out_off + u <= out_stride)), SAR_ERR_BAD_ARG,
              "sar_gru_gate_bwd: bad argument");
  launch_k(gru_gate_bwd_kernel, dim3(blocks_for((long long)B * u, 256)), dim3(256), 0, (cudaStream_t)stream, g_out, dh_rec, z, r, hh, hph,
           h_prev, d_xp, d_hu, dh_prev, B, S, u, t, out_stride, out_off);
  return check_launch("sar_gru_gate_bwd");
}

int sar_conv2d_bwd_data(const float* dy, const float* w_hwio, float* dx, int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int kh,
                        int kw, int stride, int pad_t, int pad_l, float beta, void* stream) {
  using namespace sar;
  SAR_REQUIRE(dy && w_hwio && dx, SAR_ERR_BAD_ARG, "sar_conv2d_bwd_data: null pointer");
  SAR_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Ho > 0 && Wo > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && pad_t >= 0 && pad_l >= 0,
              SAR_ERR_BAD_ARG, "sar_conv2d_bwd_data: bad dimension");
  launch_k(conv2d_bwd_data_kernel, dim3(blocks_for((long long)B * H * W * Cin, 256, 1 << 16)), dim3(256), 0, (cudaStream_t)stream, dy, w_hwio, dx,
           B, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad_t, pad_l, beta);
  return check_launch("sar_conv2d_bwd_data");
}

int sar_conv2d_bwd_weight(const float* x, const float* dy, float* partial, int chunks, int B, int H, int W, int Cin, int Ho, int Wo, int Cout,
                          int kh, int kw, int stride, int pad_t, int pad_l, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && dy && partial, SAR_ERR_BAD_ARG, "sar_conv2d_bwd_weight: null pointer");
  SAR_REQUIRE(chunks > 0 && chunks <= 65535 && B > 0 && H > 0 && W > 0 && Cin > 0 && Ho > 0 && Wo > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0,
              SAR_ERR_BAD_ARG, "sar_conv2d_bwd_weight: bad dimension");
  if (Cin >= 64 && Cout >= 64) {
    const long long tiles = (long long)kh * kw * ((Cin + 63) / 64) * ((Cout + 63) / 64);
    launch_k(conv2d_bwd_weight64_kernel, dim3((unsigned)tiles, chunks), dim3(256), 0, (cudaStream_t)stream, x, dy, partial, B, H, W, Cin, Ho, Wo,
             Cout, kh, kw, stride, pad_t, pad_l);
    return check_launch("sar_conv2d_bwd_weight");
  }
  const long long tiles = (long long)kh * kw * ((Cin + 31) / 32) * ((Cout + 31) / 32);
  SAR_REQUIRE(tiles < (1ll << 31), SAR_ERR_UNSUPPORTED, "sar_conv2d_bwd_weight: too many tiles");
  launch_k(conv2d_bwd_weight_kernel, dim3((unsigned)tiles, chunks), dim3(256), 0, (cudaStream_t)stream, x, dy, partial, B, H, W, Cin, Ho, Wo,
           Cout, kh, kw, stride, pad_t, pad_l);
  return check_launch("sar_conv2d_bwd_weight");
}

int sar_maxpool2d_bwd(const float* x, const float* dy, float* dx, int B, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad_t,
                      int pad_l, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && dy && dx && B > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && k > 0 && stride > 0, SAR_ERR_BAD_ARG,
              "sar_maxpool2d_bwd: bad argument");
  launch_k(maxpool2d_bwd_kernel, dim3(blocks_for((long long)B * H * W * C, 256, 1 << 16)), dim3(256), 0, (cudaStream_t)stream, x, dy, dx, B, H, W, C,
           Ho, Wo, k, stride, pad_t, pad_l);
  return check_launch("sar_maxpool2d_bwd");
}

int sar_axpy_fwd(const float* x, float* y, float alpha, long long n, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && y && n > 0, SAR_ERR_BAD_ARG, "sar_axpy_fwd: bad argument");
  launch_k(axpy_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, alpha, n);
  return check_launch("sar_axpy_fwd");
}

int sar_colsum_fwd(const float* g, float* out, int rows, int C, void* stream) {
  using namespace sar;
  SAR_REQUIRE(g && out && rows > 0 && C > 0, SAR_ERR_BAD_ARG, "sar_colsum_fwd: bad argument");
  launch_k(colsum_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, g, out, rows, C);
  return check_launch("sar_colsum_fwd");
}

int sar_l2norm_fwd(const float* v, float* out, float* inv_norm, int rows, int D, int axis, void* stream) {
  using namespace sar;
  SAR_REQUIRE(v && out && inv_norm && rows > 0 && D > 0 && (axis == 0 || axis == 1), SAR_ERR_BAD_ARG, "sar_l2norm_fwd: bad argument");
  const int nvec = axis ? rows : D;
  launch_k(l2norm_fwd_kernel, dim3((nvec + 63) / 64), dim3(64), 0, (cudaStream_t)stream, v, out, inv_norm, rows, D, axis);
  return check_launch("sar_l2norm_fwd");
}

int sar_l2norm_bwd(const float* vhat, const float* inv_norm, const float* u, float* out, int rows, int D, int axis, float beta,
                   void* stream) {
  using namespace sar;
  SAR_REQUIRE(vhat && inv_norm && u && out && rows > 0 && D > 0 && (axis == 0 || axis == 1), SAR_ERR_BAD_ARG, "sar_l2norm_bwd: bad argument");
  const int nvec = axis ? rows : D;
  launch_k(l2norm_bwd_kernel, dim3((nvec + 63) / 64), dim3(64), 0, (cudaStream_t)stream, vhat, inv_norm, u, out, rows, D, axis, beta);
  return check_launch("sar_l2norm_bwd");
}

int sar_head_grad_fwd(const float* z_accent, const float* c_disc, const float* onehot, int n_classes, int head, float margin, float s,
                      float gamma, float w_accent, float w_disc, float* g_accent, float* g_disc, float* losses, int B, void* stream) {
  using namespace sar;
  SAR_REQUIRE(onehot && losses && (z_accent || c_disc), SAR_ERR_BAD_ARG, "sar_head_grad_fwd: null pointer");
  SAR_REQUIRE((!z_accent || g_accent) && (!c_disc || g_disc), SAR_ERR_BAD_ARG, "sar_head_grad_fwd: missing gradient output");
  SAR_REQUIRE(B > 0 && n_classes > 0 && n_classes <= 32, SAR_ERR_UNSUPPORTED, "sar_head_grad_fwd: 1 <= n_classes <= 32");
  SAR_REQUIRE(head >= SAR_HEAD_NONE && head <= SAR_HEAD_CIRCLE_RAW, SAR_ERR_BAD_ARG, "sar_head_grad_fwd: bad head selector");
  launch_k(head_grad_kernel, dim3((B + 63) / 64), dim3(64), 0, (cudaStream_t)stream, z_accent, c_disc, onehot, n_classes, head, margin, s,
           gamma, w_accent, w_disc, g_accent, g_disc, losses, B);
  return check_launch("sar_head_grad_fwd");
}

int sar_adam_fwd(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1, float beta2, float eps, float l2,
                 void* stream) {
  using namespace sar;
  SAR_REQUIRE(p && g && m && v && n > 0, SAR_ERR_BAD_ARG, "sar_adam_fwd: bad argument");
  launch_k(adam_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, lr_t, beta1, beta2, eps, l2);
  return check_launch("sar_adam_fwd");
}

int sar_adam_dev_fwd(float* p, const float* g, float* m, float* v, long long n, const float* lr_t, float beta1, float beta2, float eps,
                     float l2, void* stream) {
  using namespace sar;
  SAR_REQUIRE(p && g && m && v && lr_t && n > 0, SAR_ERR_BAD_ARG, "sar_adam_dev_fwd: bad argument");
  launch_k(adam_dev_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, lr_t, beta1, beta2, eps, l2);
  return check_launch("sar_adam_dev_fwd");
}

int sar_unit_norm_fwd(float* w, int D, int n, void* stream) {
  using namespace sar;
  SAR_REQUIRE(w && D > 0 && n > 0, SAR_ERR_BAD_ARG, "sar_unit_norm_fwd: bad argument");
  launch_k(unit_norm_kernel, dim3((n + 63) / 64), dim3(64), 0, (cudaStream_t)stream, w, D, n);
  return check_launch("sar_unit_norm_fwd");
}

int sar_vlad_train_fwd(const float* x, const float* w_assign, const float* b_assign, const float* centers, float* A, float* R,
                       float* asum, int B, int S, int D, int K, int G, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && w_assign && b_assign && centers && A && R && asum, SAR_ERR_BAD_ARG, "sar_vlad_train_fwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0 && S <= 128 && D > 0 && K > 0 && G >= 0 && K + G <= 128, SAR_ERR_UNSUPPORTED,
              "sar_vlad_train_fwd: need S <= 128 and K + G <= 128 (S=%d K=%d G=%d)", S, K, G);
  const size_t smem = ((size_t)S * (K + G) + (K + G)) * sizeof(float);
  { const int arc = allow_max_smem(vlad_train_fwd_kernel, "sar_vlad_train_fwd"); if (arc) return arc; }
  launch_k(vlad_train_fwd_kernel, dim3(B), dim3(256), smem, (cudaStream_t)stream, x, w_assign, b_assign, centers, A, R, asum, S, D, K, K + G);
  return check_launch("sar_vlad_train_fwd");
}

int sar_ln_train_bwd(const float* y, const float* gamma, const float* g_z, float* g_pre, float* gz_xhat, int rows, int C, float eps,
                     int tanh_in, void* stream) {
  using namespace sar;
  SAR_REQUIRE(y && gamma && g_z && g_pre && gz_xhat && rows > 0 && C > 0, SAR_ERR_BAD_ARG, "sar_ln_train_bwd: bad argument");
  launch_k(ln_train_bwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, (cudaStream_t)stream, y, gamma, g_z, g_pre, gz_xhat, rows, C, eps,
           tanh_in ? 1 : 0);
  return check_launch("sar_ln_train_bwd");
}

int sar_vlad_train_bwd(const float* x, const float* A, const float* centers, const float* gR, const float* asum, float* g_scores,
                       float* gc_part, float* g_x, int B, int S, int D, int K, int G, void* stream) {
  using namespace sar;
  SAR_REQUIRE(x && A && centers && gR && asum && g_scores && gc_part, SAR_ERR_BAD_ARG, "sar_vlad_train_bwd: null pointer");
  SAR_REQUIRE(B > 0 && S > 0 && S <= 128 && D > 0 && K > 0 && G >= 0 && K + G <= 128, SAR_ERR_UNSUPPORTED,
              "sar_vlad_train_bwd: need S <= 128 and K + G <= 128 (S=%d K=%d G=%d)", S, K, G);
  const size_t smem = (size_t)S * K * sizeof(float);
  { const int arc = allow_max_smem(vlad_train_bwd_kernel, "sar_vlad_train_bwd"); if (arc) return arc; }
  launch_k(vlad_train_bwd_kernel, dim3(B), dim3(256), smem, (cudaStream_t)stream, x, A, centers, gR, asum, g_scores, gc_part, g_x, S, D, K, K + G);
  return check_launch("sar_vlad_train_bwd");
}

}  // extern "C"
