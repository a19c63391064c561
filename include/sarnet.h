/* sarnet.h -- C ABI of libsarnet_sm100.so, the B200 (sm_100a) engine for the
 * speech-accent-recognition network of pika-online/AESRC2020: the inference forward path on
 * tcgen05 tensor-core kernels, and the training-mode forward / backward kernels (fp32,
 * correctness-first; the "training mode" section near the end).
 *
 * The reference has NO native/FFI interface (it is Keras-on-TensorFlow Python); its
 * boundary for this path is the Python call surface of model.py / resnet.py / VLAD.py /
 * losses.py.  Each entry point below therefore cites the reference *function* whose
 * arithmetic it replaces (file:line under the reference tree); the Python mirror in
 * aesrc2020_b200/ binds them with ctypes (see INTEGRATION.md for the stub a reference
 * maintainer would add).
 *
 * Conventions
 *  - every pointer is CALLER-OWNED DEVICE memory unless the name ends in `_host`;
 *    the library never allocates, frees or synchronises (stream-ordered);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - tensors are dense, row-major, channels-last (NHWC) exactly like the reference
 *    (resnet.py:12-14); fp32 unless stated; 16-byte aligned base pointers;
 *  - return 0 on success, a negative sar_status for rejected arguments, the positive
 *    cudaError_t for a launch failure; sar_last_error() returns a thread-local message;
 *  - re-entrant across streams/devices: the caller sets the device; no global state.
 */
#ifndef SARNET_H_
#define SARNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SAR_OK = 0,
  SAR_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, inconsistent shape */
  SAR_ERR_UNSUPPORTED = -2,  /* shape/option combination this build has no kernel for */
  SAR_ERR_ALIGN = -3,        /* pointer or leading dimension not aligned as required */
  SAR_ERR_WORKSPACE = -4     /* workspace too small (see the *_workspace_bytes helper) */
} sar_status;

/* activation selector for the fused epilogues */
enum { SAR_ACT_NONE = 0, SAR_ACT_RELU = 1, SAR_ACT_TANH = 2 };

/* margin-head selector: model.py:142-167 (disc_loss) */
enum {
  SAR_HEAD_NONE = 0,
  SAR_HEAD_SOFTMAX = 1,     /* Dense(n, softmax, use_bias=False)            model.py:149-150 */
  SAR_HEAD_SPHEREFACE = 2,  /* losses.py:12-50  */
  SAR_HEAD_COSFACE = 3,     /* losses.py:58-94  */
  SAR_HEAD_ARCFACE = 4,     /* losses.py:105-147 */
  SAR_HEAD_CIRCLE = 5,      /* l2norm(x) @ W raw cosines + circle_loss      model.py:161-163, losses.py:157-172 */
  SAR_HEAD_CIRCLE_RAW = 6   /* circle_loss on x @ W WITHOUT normalising x: with W = I this evaluates
                               losses.circle_loss(y_true, y_pred) on given cosines (losses.py:157-172) */
};

int sar_version(void);
const char* sar_last_error(void);
/* compiled SM architecture of the embedded cubin (100 for sm_100a) */
int sar_compiled_arch(void);

/* ---- ResNet front-end -------------------------------------------------------------- */

/* Generic NHWC convolution on CUDA cores (fp32 FFMA implicit GEMM), used for the 7x7/s2
 * stem (Cin=1), and as the Dense/GEMM primitive (kh=kw=1, W=1).
 * Replaces: Conv2D call sites resnet.py:39-42, 60-63, 82-87, 113-117; Dense model.py:35-42.
 *   in' = pre_scale ? relu(pre_scale[ci]*x + pre_shift[ci]) : x     (_bn_relu on the conv INPUT,
 *         applied to in-bounds cells only: TF-SAME zero padding pads the ACTIVATED tensor)
 *   acc = conv(in', w_hwio) + bias + (residual ? residual : 0)     (Add(), resnet.py:89)
 *   out = act( post_scale ? post_scale[co]*acc + post_shift[co] : acc )
 * x (B,H,W,Cin); w (kh,kw,Cin,Cout) Keras HWIO; out/residual (B,Ho,Wo,Cout).
 * pad_t/pad_l are the TF-SAME leading pads (trailing pads are implied by Ho/Wo). */
int sar_conv2d_fwd(const float* x, const float* w_hwio, const float* bias,
                   const float* pre_scale, const float* pre_shift,
                   const float* post_scale, const float* post_shift,
                   const float* residual, float* out,
                   int B, int H, int W, int Cin, int Ho, int Wo, int Cout,
                   int kh, int kw, int stride, int pad_t, int pad_l, int act, void* stream);

/* ---- tensor-core path of the residual blocks --------------------------------------------
 * Activation layout "flat-pad hi/lo planes" (fp16): an (H,W,C) map of a batch of B is
 *   planes[plane][q][c],  q = n*(H+1)*(W+1) + h*(W+1) + w,  R = B*(H+1)*(W+1) rows per plane,
 * with ONE shared zero pad column (w == W) and pad row (h == H) per image; plane 0 = hi =
 * fp16(x), plane 1 = lo = fp16((x - hi) * 2^11) (x == hi + lo/2048 to ~2^-22).  A tensor that
 * feeds a stride-2 block is phase-split: 4 such pairs, plane = 2*((h&1)*2 + (w&1)) + {0,1},
 * each over the geometry H2 = ceil(H/2), W2 = ceil(W/2).  Buffers must be zero-initialised once
 * by the caller: kernels never write pad positions (that is what makes TF-SAME zero padding,
 * including its asymmetric stride-2 form, a pure row shift). */
size_t sar_planes_bytes(int B, int H, int W, int C, int split);

/* dense fp32 NHWC -> planes, optionally y = relu?(scale[c]*x + shift[c]) first (_bn_relu,
 * resnet.py:22-26).  C % 8 == 0. */
int sar_planes_pack_fwd(const float* x, const float* scale, const float* shift, int relu, void* planes,
                        int B, int H, int W, int C, int split, void* stream);
/* planes -> dense fp32 NHWC (debug / tests / API boundary). */
int sar_planes_unpack_fwd(const void* planes, float* x, int B, int H, int W, int C, int split, void* stream);
/* MaxPooling2D(3x3, strides 2, 'same') (resnet.py:174,192) from dense fp32 NHWC into planes. */
int sar_maxpool_planes_fwd(const float* x, void* planes, int B, int H, int W, int C, int Ho, int Wo,
                           int k, int stride, int pad_t, int pad_l, void* stream);

/* The whole ResNet stem in one kernel: Conv2D 7x7/s2 'same' (+bias) -> BatchNormalization -> ReLU ->
 * MaxPooling2D 3x3/s2 'same' (resnet.py:28-45 via :173/:191, and :174/:192).  x (B,T,D) fp32 (the
 * (B,T,D,1) x_data tensor), w (7,7,1,F0) Keras HWIO, scale/shift = folded inference BN; output: the
 * pooled (B, ceil(ceil(T/2)/2), ceil(ceil(D/2)/2), F0) map as flat-pad hi/lo planes (non-split).
 * TF-SAME pads are derived from T and D inside.  F0 in {16,32,48,64}; ceil(D/2) % 4 == 0. */
int sar_stem_pool_fwd(const float* x, const float* w, const float* bias, const float* scale,
                      const float* shift, void* planes, int B, int T, int D, int F0, void* stream);

/* One residual-block convolution on tcgen05 tensor cores (csrc/conv_tc.cu).
 * Replaces _bn_relu_conv / basic_block / _shortcut: resnet.py:47-65, 105-125, 67-89.
 *   acc = sum_taps A[q + tap_row_off[t], plane tap_plane[t]] @ W_t  (+ S[q, s_plane] @ W_s)
 *   v   = acc + bias (+ res[q])                         (Add(), resnet.py:89)
 *   out_raw = v ; out_act / out_dense = relu(act_scale*v + act_shift)   (next layer's _bn_relu)
 * a: planes of the ALREADY ACTIVATED conv input ([a_planes][a_rows][a_ch]); for a stride-2 conv
 *    it is phase-split and tap_plane/tap_row_off select phase and shift per tap.
 * s: planes of the RAW block input for the 1x1 projection shortcut (NULL = none).
 * w: [2][cout][ntaps*a_ch + s_ch] fp16 hi/lo, K-major, k = tap*a_ch + ci (shortcut rows last).
 * bias: conv bias (+ shortcut conv bias).  res: identity-shortcut planes [2][R][cout] or NULL.
 * Outputs: out_raw and/or out_act planes (phase-split when out_split), or out_raw and/or out_dense fp32
 * (B,H,W,cout) -- out_act and out_dense are the same activated values in two layouts, at most one of them.  All channel counts are multiples of 32. */
typedef struct sar_tc_conv {
  const void* a; long long a_rows; int a_ch; int a_planes;
  int ntaps; int tap_row_off[9]; int tap_plane[9];
  const void* s; long long s_rows; int s_ch; int s_planes; int s_plane;
  const void* w; int cout;
  const float* bias;
  const void* res;
  void* out_raw; void* out_act;
  const float* act_scale; const float* act_shift;
  float* out_dense;
  int out_split;
  int B, H, W;          /* output map geometry */
  void* dbg;            /* optional device buffer of >= 256 int64 clock64() stamps of CTA 0 (profiling aid); NULL */
  int act_kind;         /* out_act / out_dense activation: 0 relu(act_scale*v+act_shift) (the next layer's BN->ReLU),
                           1 identity, 2 tanh(v) -- the Dense layers of model.py:35-42 run as 1-tap "convolutions" */
  int nopad;            /* 1: the operand / output rows are plain row-major (B*H*W rows, no pad row or column): GEMM use */
  int ksplit;           /* > 1: split-K for a 1-tap GEMM with a long K (a_ch): K slice z writes its fp32 partial products to
                           out_dense + z*B*H*W*cout (identity activation, zero bias); sum them with sar_splitk_reduce_fwd */
  const float* res_f32; /* the identity shortcut as ONE fp32 plane [R][cout] (flat-pad rows) instead of `res` hi/lo planes */
  float* out_raw_f32;   /* the raw sum as one fp32 plane [R][cout] instead of `out_raw`: the residual stream between the
                           blocks of a stage is only ever ADDED in an epilogue (resnet.py:123,89), never an MMA operand, so
                           it needs no hi/lo split; pad rows are left unwritten (their sums are discarded downstream).
                           Needs unsplit outputs (the TMA-store epilogue). */
} sar_tc_conv;
int sar_conv_tc_fwd(const sar_tc_conv* d, void* stream);

/* Up to 12 stride-1 3x3 layers of ONE ResNet stage (same B, H, W, a_ch == cout; plane outputs; only descs[0] may
 * carry a projection shortcut `s`) as a single persistent launch: layer i+1 reads what layer i wrote (its `a` is
 * layer i's out_act), tiles synchronise through per-M-tile counters in `workspace` (device memory, zero-filled once
 * by the caller, sar_conv_tc_chain_workspace_bytes(); the kernel leaves it zeroed) instead of kernel boundaries.
 * Results are bitwise those of n sar_conv_tc_fwd calls. */
size_t sar_conv_tc_chain_workspace_bytes(const sar_tc_conv* first, int n);
int sar_conv_tc_chain_fwd(const sar_tc_conv* descs, int n, void* workspace, size_t workspace_bytes, void* stream);
/* Same with the grid capped at `max_ctas` CTAs (0: one per SM): the launch is cooperative (all CTAs co-resident), so two
 * capped chains of different streams share the GPU -- every CTA then walks several tiles per layer back to back, with the
 * epilogue of one under the mainloop of the next (less SM-time per layer than one launch per layer). */
int sar_conv_tc_chain_grid_fwd(const sar_tc_conv* descs, int n, void* workspace, size_t workspace_bytes, int max_ctas,
                               void* stream);

/* MaxPooling2D(3x3, strides 2, 'same') -- resnet.py:174,192.  Padded cells never win. */
int sar_maxpool2d_fwd(const float* x, float* out, int B, int H, int W, int C, int Ho, int Wo,
                      int k, int stride, int pad_t, int pad_l, void* stream);

/* y = relu(scale[c]*x + shift[c]) over the last axis: _bn_relu, resnet.py:22-26 (inference BN
 * folded to an affine by the caller: scale=gamma/sqrt(var+1e-3), shift=beta-mean*scale). */
int sar_affine_relu_fwd(const float* x, const float* scale, const float* shift, float* out,
                        long long rows, int C, int relu, void* stream);

/* ---- encoder tail ------------------------------------------------------------------ */

/* LayerNormalization over the last axis (keras_layer_normalization, model.py:32-33):
 * y = gamma*(x-mean)/sqrt(var+eps)+beta, biased variance, two-pass fp32.  C <= 1024. */
int sar_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* out,
                      long long rows, int C, float eps, void* stream);
/* Same, writing the normalised rows as fp16 hi/lo planes [2][plane_rows][C] for a following tensor-core Dense
 * (sar_conv_tc_fwd with one tap): input row r lands on plane row r + r/seg (one zero pad row after every `seg`
 * rows: the flat-pad layout of a (1, rows/seg, seg) map; seg = 0: no pad rows).  `out` (fp32, optional) also
 * receives the rows in the plain layout.  C % 8 == 0. */
int sar_layernorm_planes_fwd(const float* x, const float* gamma, const float* beta, float* out, void* planes,
                             long long plane_rows, int seg, long long rows, int C, float eps, void* stream);

/* Recurrent part of Bidirectional(CuDNNGRU) -- model.py:44-50.
 * xp (B,S,2,3u): input projections x*W + b_input for [forward | backward], gate order z|r|h
 *   (computed by sar_conv2d_fwd as one GEMM against the concatenated kernels);
 * rec (2,u,3u) recurrent kernels, rbias (2,3u) recurrent biases.
 * seq!=0: out (B,S,2u) = concat(fwd[t], bwd[t]) in time order;
 * seq==0: out (B,2u)   = concat(fwd final state, bwd final state).  u must be 256. */
int sar_bigru_fwd(const float* xp, const float* rec, const float* rbias, float* out,
                  int B, int S, int u, int seq, void* stream);
/* Same, choosing the utterances per 8-CTA cluster: nb = 16 (shortest step: half the SM-to-SM bytes per CTA),
 * 32 (half the SMs for the same batch: the better choice when other work shares the GPU), 0 = automatic
 * (16 while both directions of the batch fit in one wave of 15 resident clusters). */
int sar_bigru_nb_fwd(const float* xp, const float* rec, const float* rbias, float* out,
                     int B, int S, int u, int seq, int nb, void* stream);

/* ---- many-to-one integration ------------------------------------------------------- */

/* NetVLAD / GhostVLAD: the 1x1 assignment conv (model.py:89-95 / 99-105) fused with
 * VladPooling.call (VLAD.py:26-49).  feat (B,S,D); centers (K+G,D), ghosts are the LAST G
 * rows; out (B,K*D), per-cluster L2-normalised.  Cluster scores come from EITHER the fused
 * assignment conv (w_assign (D,K+G), b_assign (K+G); score = NULL) -- the model.vlad() path --
 * OR a precomputed `score` (B,S,K+G) tensor (w_assign = b_assign = NULL) -- the bare
 * VladPooling([feat, cluster_score]) call surface of VLAD.py:26-28.
 * Requires D == 256 (hidden_dim of this build) and K+G <= 128. */
int sar_vlad_fwd(const float* feat, const float* w_assign, const float* b_assign, const float* score,
                 const float* centers, float* out, int B, int S, int D, int K, int G, void* stream);
/* Same, with the descriptor ALSO / INSTEAD written as fp16 hi/lo planes [2][B][K*D] (x = hi + lo/2048) for the
 * tensor-core AR_EMBEDDING GEMM (sar_conv_tc_fwd with nopad + ksplit).  `out` and `out_planes` may each be NULL. */
int sar_vlad_planes_fwd(const float* feat, const float* w_assign, const float* b_assign, const float* score,
                        const float* centers, float* out, void* out_planes, int B, int S, int D, int K, int G, void* stream);

/* Tensor-core NetVLAD / GhostVLAD (csrc/vlad_tc.cu): the same result as sar_vlad_planes_fwd with the fused assignment
 * conv, but both contractions (scores = X @ Wa, V = A^T @ X) run as tcgen05.mma (fp16 hi/lo operands, fp32 accumulate).
 *   x_planes   fp16 hi/lo planes [2][x_rows][256] of the (B*S, 256) descriptors (sar_layernorm_planes_fwd, seg = 0)
 *   wa_packed  fp16 hi/lo [2][KGP][256], KGP = K+G rounded up to 16: row kg = column kg of the assignment kernel
 *              (w_assign[:, kg]), zero rows for the padding
 * `out` (B, K*256) fp32 and/or `out_planes` [2][B][K*256] fp16 hi/lo.  Shapes: D == 256, S <= 128, K+G <= 128 and the
 * tiles must fit shared memory: sar_vlad_tc_supported() returns 1 when they do (0: call sar_vlad_fwd instead).
 * Replaces: model.py:82-109 (vlad()), VLAD.py:26-49 (VladPooling.call). */
int sar_vlad_tc_supported(int B, int S, int D, int K, int G);
int sar_vlad_tc_fwd(const void* x_planes, long long x_rows, const void* wa_packed, const float* b_assign,
                    const float* centers, float* out, void* out_planes, int B, int S, int D, int K, int G, void* stream);

/* Row softmax: out (rows, C) = softmax over the first C columns of x (rows, ld), ld >= C.
 * Replaces: the 'softmax' activation of Dense (model.py:35-42; ctc_pred, model.py:268) -- the posteriors
 * K.ctc_decode reads in ctc_pred() (model.py:385-389).  ld > C: rows padded by the tensor-core ctc_pred GEMM. */
int sar_softmax_rows_fwd(const float* x, int ld, float* out, long long rows, int C, void* stream);

/* GlobalAveragePooling1D (model.py:125): out (B,D) = mean over S of x (B,S,D). */
int sar_avgpool_fwd(const float* x, float* out, int B, int S, int D, void* stream);

/* Split-K GEMM for the AR_EMBEDDING layer (model.py:286-289; AR_BN1/AR_BN2 folded into
 * w/bias by the caller): out (M,N) = a (M,K) @ w (K,N) + bias.  Deterministic two-pass
 * reduction through `workspace` (sar_gemm_splitk_workspace_bytes). */
size_t sar_gemm_splitk_workspace_bytes(int M, int K, int N);
int sar_gemm_splitk_fwd(const float* a, const float* w, const float* bias, float* out,
                        int M, int K, int N, void* workspace, size_t workspace_bytes, void* stream);

/* out (M,N) = bias + sum over `splits` partial (M,N) fp32 tiles of `ws`, in a fixed order (deterministic). */
int sar_splitk_reduce_fwd(const float* ws, const float* bias, float* out, int M, int N, int splits, void* stream);

/* ---- heads and losses -------------------------------------------------------------- */

/* Classifier MLP + margin head + per-sample losses in one pass over the embedding row.
 * Replaces: AR_CF_DS1/AR_CF_DS2/y_accent (model.py:294-296), disc_loss (model.py:142-167),
 * SphereFace/CosFace/ArcFace.call (losses.py:28-47,74-91,121-144), circle_loss
 * (losses.py:157-172), categorical_crossentropy + accuracy wiring (model.py:344-367).
 * emb (B,D).  Classifier weights may be NULL (then y_accent outputs are skipped).
 * wd (Dd,n) is the head weight applied to `emb_d` (B,Dd) (NULL -> emb, Dd=D).
 * onehot (B,n) may be NULL when head is NONE/SOFTMAX/CIRCLE and no losses are wanted.
 * Outputs (any may be NULL): y_accent (B,n) probs, y_accent_logits (B,n),
 * y_disc (B,n) probs (raw cosines for CIRCLE), y_disc_logits (B,n) pre-softmax,
 * sample_stats (B,4) = [loss_accent, loss_disc, correct_accent, correct_disc]. */
int sar_head_fwd(const float* emb, int D,
                 const float* w1, const float* b1, int H1,
                 const float* w2, const float* b2, int H2,
                 const float* w3, const float* b3,
                 const float* emb_d, int Dd, const float* wd,
                 const float* onehot, int n_classes, int head, float margin, float s, float gamma,
                 float* y_accent, float* y_accent_logits, float* y_disc, float* y_disc_logits,
                 float* sample_stats, int B, void* stream);

/* CTC negative log-likelihood from PRE-softmax ctc_pred logits.
 * Replaces: Dense(softmax) 'ctc_pred' activation + K.ctc_batch_cost (model.py:268-269,62-71):
 * p = softmax(logits); q = (p+1e-7)/sum(p+1e-7); blank = C-1; merge_repeated.
 * logits (B,S,C); labels (B,Lmax) FLOAT ids (utils.py:107); in_len/lab_len (B) int32.
 * loss (B).  probs (B,S,C) optional output of the softmax (model.predict of ctc_pred).
 * status (B) int32 optional: 0 ok, 1 infeasible label sequence (TF raises), 2 bad label. */
int sar_ctc_fwd(const float* logits, const float* labels, const int* in_len, const int* lab_len,
                float* loss, float* probs, int* status, int B, int S, int C, int Lmax, void* stream);
/* Same, with `ld` >= C floats between consecutive (b, s) rows of `logits`: the tensor-core ctc_pred Dense pads
 * its output columns to a multiple of 32 (1000 -> 1024); only the first C columns of a row are classes. */
int sar_ctc_ld_fwd(const float* logits, int ld, const float* labels, const int* in_len, const int* lab_len,
                   float* loss, float* probs, int* status, int B, int S, int C, int Lmax, void* stream);
/* Training mode (model.py:62-71 under compile(), model.py:187-201): the same loss plus grad (B,S,C) = scale * d loss_b / d logits
 * (alpha-beta recursion; through q = softmax(log(softmax(logits) + 1e-7))); frames >= in_len and utterances with status != 0
 * get zero gradient. */
int sar_ctc_grad_fwd(const float* logits, int ld, const float* labels, const int* in_len, const int* lab_len,
                     float* loss, float* grad, int* status, int B, int S, int C, int Lmax, float scale, void* stream);

/* Greedy CTC decode of the ctc_pred posteriors from PRE-softmax logits (rows `ld` >= C floats apart).
 * Replaces: ctc_pred() = K.ctc_decode(pred, input_len, greedy=True) (model.py:385-389; tf.nn.ctc_greedy_decoder with
 * merge_repeated=True): per frame the first maximum over the C classes, repeats merged, blank = C-1 dropped.
 * in_len (B) int32 or null (then every utterance uses fixed_len, as the reference's constant input_len does).
 * dec (B,S) int32, padded with -1 like the dense tensor K.ctc_decode returns; dec_len (B) int32. */
int sar_ctc_greedy_fwd(const float* logits, int ld, const int* in_len, int fixed_len, int* dec, int* dec_len,
                       int B, int S, int C, void* stream);

/* Deterministic batch reduction of per-sample statistics into the 8-float vector that is
 * all-reduced across GPUs (one ncclAllReduce(SUM), replaces multi_gpu_model, model.py:193-194):
 * out8 = [sum loss_accent, sum loss_disc, sum loss_ctc, sum loss_disc_bn,
 *         #correct_accent, #correct_disc, count, #correct_disc_bn].
 * sample_stats (B,4) or NULL, ctc_loss (B) or NULL, bn_stats (B,4) or NULL. */
int sar_loss_reduce_fwd(const float* sample_stats, const float* ctc_loss, const float* bn_stats,
                        float* out8, int B, void* stream);

/* ---- training mode, first slice ------------------------------------------
 * Building blocks of one optimisation step of the accent head -- the layers after integration() of model.py:286-296 and
 * disc_loss (model.py:142-167) in TRAINING mode, the loss wiring of model.py:344-367 and compile()'s Adam(lr, decay=2e-4)
 * (model.py:187-201).  fp32, deterministic; aesrc2020_b200/training.py composes them (HeadTrainer). */

/* C (M,N) = alpha * op(A) op(B) + beta * C, row-major; trans_a: A is stored (K,M); trans_b: B is stored (N,K).
 * Replaces: the Dense forward / backward contractions (x W, g W^T, x^T g). */
int sar_gemm_fwd(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, int trans_b, float alpha, float beta,
                 void* stream);
/* BatchNormalization in training mode on (rows, C): batch mean / BIASED variance per replica (what multi_gpu_model's towers
 * do), y = gamma (x - mean) / sqrt(var + eps) + beta, moving statistics updated in place with `momentum` (NULL: not
 * updated).  save_mean / save_invstd (C) feed the backward.  Replaces: BN(name=...) of model.py:29-30 with learning phase 1. */
int sar_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* y,
                     float* save_mean, float* save_invstd, int rows, int C, float eps, float momentum, void* stream);
int sar_bn_train_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd,
                     float* dx /* may be NULL */, float* dgamma, float* dbeta, int rows, int C, void* stream);
/* Row-parallel forms for maps with many rows (the ResNet's BatchNormalizations in training mode, conv bias gradients): the rows
 * are reduced in `nch` chunks with a fixed summation order; `ws` = caller-owned scratch of (2 * nch + 2) * C floats. */
int sar_colsum_rows_fwd(const float* g, float* out, int rows, int C, int nch, float* ws, void* stream);
int sar_bn_train_rows_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* y,
                          float* save_mean, float* save_invstd, int rows, int C, float eps, float momentum, int nch, float* ws,
                          void* stream);
int sar_bn_train_rows_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd,
                          float* dx /* may be NULL */, float* dgamma, float* dbeta, int rows, int C, int nch, float* ws, void* stream);
/* y = act(x + bias) on (rows, C), act in {SAR_ACT_NONE, SAR_ACT_RELU, SAR_ACT_TANH}; out = g * (h > 0); out (C) = column sums of g (rows, C). */
int sar_bias_act_fwd(const float* x, const float* bias, float* y, long long rows, int C, int act, void* stream);
int sar_relu_bwd(const float* g, const float* h, float* out, long long n, void* stream);
int sar_colsum_fwd(const float* g, float* out, int rows, int C, void* stream);
/* Backward of Conv2D / MaxPooling2D on NHWC fp32 maps (resnet.py in training mode; HWIO kernels, TF-SAME leading pads, any stride):
 *   sar_conv2d_bwd_data:   dx (B,H,W,Cin) = beta * dx + d loss / d x from dy (B,Ho,Wo,Cout)
 *   sar_conv2d_bwd_weight: partial (chunks, kh*kw*Cin*Cout): the output positions split into `chunks` ranges, one partial
 *                          d loss / d w per range (sum them with sar_colsum_fwd: deterministic order)
 *   sar_maxpool2d_bwd:     dx = dy routed to the first maximum of every window (padded cells never win)
 *   sar_axpy_fwd:          y += alpha * x  (gradient accumulation where two paths meet: Add(), resnet.py:89)
 * Correctness-first CUDA-core kernels (fixed summation orders), not tensor-core code. */
int sar_conv2d_bwd_data(const float* dy, const float* w_hwio, float* dx, int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int kh,
                        int kw, int stride, int pad_t, int pad_l, float beta, void* stream);
int sar_conv2d_bwd_weight(const float* x, const float* dy, float* partial, int chunks, int B, int H, int W, int Cin, int Ho, int Wo, int Cout,
                          int kh, int kw, int stride, int pad_t, int pad_l, void* stream);
int sar_maxpool2d_bwd(const float* x, const float* dy, float* dx, int B, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad_t,
                      int pad_l, void* stream);
int sar_axpy_fwd(const float* x, float* y, float alpha, long long n, void* stream);
/* One time step of one direction of CuDNNGRU in TRAINING mode (model.py:44-50; reset_after, gates z|r|h) and its backward.
 * xp (B,S,3u) = x W + b_i (batch-major), hu (B,3u) = h_prev U, b_r (3u); z, r, hh, hph (B,u) are this step's saved gates
 * (hph = hu_h + b_r_h), h_new (B,u), out (B,S,out_stride) receives h_new at [b, t, out_off + j] (NULL: not stored).
 * Backward: dh = g_out[b,t,out_off+j] + dh_rec[b,j] (either may be NULL); d_xp[b,t,:] = d loss / d (x W + b_i),
 * d_hu (B,3u) = d loss / d (h_prev U + b_r), dh_prev = dh * z (the caller adds d_hu U^T).  fp32, CUDA cores. */
int sar_gru_gate_fwd(const float* xp, const float* hu, const float* b_r, const float* h_prev, float* z, float* r, float* hh,
                     float* hph, float* h_new, float* out, int B, int S, int u, int t, int out_stride, int out_off, void* stream);
int sar_gru_gate_bwd(const float* g_out, const float* dh_rec, const float* z, const float* r, const float* hh, const float* hph,
                     const float* h_prev, float* d_xp, float* d_hu, float* dh_prev, int B, int S, int u, int t, int out_stride,
                     int out_off, void* stream);
/* K.l2_normalize of every row (axis 1) or column (axis 0) of v (rows, D): out = v / sqrt(max(|v|^2, 1e-12)), inv_norm per
 * vector; backward: out = beta * out + (u - vhat (vhat . u)) * inv_norm, u = d loss / d vhat.  (losses.py:30-33, model.py:162) */
int sar_l2norm_fwd(const float* v, float* out, float* inv_norm, int rows, int D, int axis, void* stream);
int sar_l2norm_bwd(const float* vhat, const float* inv_norm, const float* u, float* out, int rows, int D, int axis, float beta,
                   void* stream);
/* Losses of the two accent outputs and their gradients (model.py:344-357, losses.py): z_accent (B,n) pre-softmax logits of
 * y_accent; c_disc (B,n) the margin head's cosines (SAR_HEAD_SPHEREFACE / COSFACE / ARCFACE / CIRCLE) or logits
 * (SAR_HEAD_SOFTMAX).  g_accent = w_accent / B * dCE/dz, g_disc = w_disc / B * d loss_disc / d c_disc (margin, scale s and
 * Circle-Loss weighting chained in), losses (B,2) = per-sample [CE_accent, loss_disc]. */
int sar_head_grad_fwd(const float* z_accent, const float* c_disc, const float* onehot, int n_classes, int head, float margin, float s,
                      float gamma, float w_accent, float w_disc, float* g_accent, float* g_disc, float* losses, int B, void* stream);
/* Keras 2.2.4 Adam: g' = g + 2 l2 p (the l2 regulariser of DS, model.py:35-42); m, v, p updated in place;
 * p -= lr_t m / (sqrt(v) + eps).  lr_t = lr / (1 + decay * iterations) * sqrt(1 - beta2^t) / (1 - beta1^t) is the caller's. */
int sar_adam_fwd(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1, float beta2, float eps, float l2,
                 void* stream);
/* The same update with the step size read from DEVICE memory (`lr_t`, one float): lets a captured CUDA graph of the whole
 * training step be replayed while lr_t changes every iteration (bias correction, decay). */
int sar_adam_dev_fwd(float* p, const float* g, float* m, float* v, long long n, const float* lr_t, float beta1, float beta2, float eps,
                     float l2, void* stream);
/* keras.constraints.unit_norm(axis=0) on W (D, n): the Circle-Loss head's kernel constraint (model.py:163). */
int sar_unit_norm_fwd(float* w, int D, int n, void* stream);
/* NetVLAD / GhostVLAD pooling in TRAINING mode (model.py:82-109: the 1x1 assignment Conv2D with l2(1e-4) kernel / bias
 * regularisers; VLAD.py:26-49: softmax over the K+G clusters, residual sums, ghost rows dropped).  x (B,S,D) the frozen
 * descriptors (AR_DS_LN output), w_assign (D,K+G), b_assign (K+G), centers (K+G,D).  Forward: A (B,S,K+G) soft assignments
 * (kept for the backward), asum (B,K) = sum_s A, R (B,K,D) = sum_s A[s,k] x[s,:] - asum[k] c[k,:]; the per-cluster
 * K.l2_normalize (VLAD.py:47) is sar_l2norm_fwd on R viewed as (B*K, D).  Backward, from gR = d loss / d R (sar_l2norm_bwd):
 * g_scores (B,S,K+G) = d loss / d scores (softmax chained in; g_w_assign = x^T g_scores via sar_gemm_fwd, g_b_assign =
 * sar_colsum_fwd(g_scores)) and gc_part (B,K,D) = -asum[b,k] gR[b,k,:], whose sum over b (sar_colsum_fwd on (B, K*D)) is the
 * gradient of the K real centers (ghost centers have none).  S <= 128, K+G <= 128. */
int sar_vlad_train_fwd(const float* x, const float* w_assign, const float* b_assign, const float* centers, float* A, float* R,
                       float* asum, int B, int S, int D, int K, int G, void* stream);
int sar_vlad_train_bwd(const float* x, const float* A, const float* centers, const float* gR, const float* asum, float* g_scores,
                       float* gc_part, float* g_x /* may be NULL: (B,S,D) = sum_k A[s,k] gR[k,:], the residual-sum path of d loss / d x;
                       the score path g_scores w_assign^T is one sar_gemm_fwd with beta = 1 */,
                       int B, int S, int D, int K, int G, void* stream);
/* Backward of LN(tanh(.)) = DS(..., 'tanh') -> LayerNormalization (model.py:32-42, the AR_DS / AR_DS_LN pair): y (rows, C) is the
 * LN input (= tanh(pre) when tanh_in), g_z = d loss / d LN output; g_pre = d loss / d pre (or d loss / d y when !tanh_in);
 * gz_xhat = g_z * xhat, whose column sums are d loss / d gamma (d loss / d beta = column sums of g_z).  eps = 1e-14. */
int sar_ln_train_bwd(const float* y, const float* gamma, const float* g_z, float* g_pre, float* gz_xhat, int rows, int C, float eps,
                     int tanh_in, void* stream);

/* ---- feature front-end ------------------------------------------------------------- */

/* On-device restatement of psf.fbank(y, 16000, nfilt=80)[0] + feat_norm + feat_reshape
 * (local/make_fbank.py:24-28, utils.py:35-46): preemphasis 0.97, 400/160 rectangular
 * frames, 512-point power spectrum / 512, 80 triangular mel filters (linear energies, no
 * log), per-utterance per-bin min-max to [0,1], truncate / zero-pad to T frames.
 * wav (total samples) fp32 concatenated utterances; offsets (B+1) int64 sample offsets;
 * melfb_t (257,80) BIN-MAJOR filterbank matrix (built by the caller,
 * aesrc2020_b200.fbank.mel_filterbank(...).T); feat_ws (B,Fmax,80) workspace for the raw
 * energies, Fmax >= the longest utterance's frame count 1+ceil((n-400)/160);
 * x_data (B,T,80) output. */
int sar_fbank_fwd(const float* wav, const long long* offsets, const float* melfb_t,
                  float* feat_ws, float* x_data, int B, int Fmax, int T, void* stream);
/* Same from 16-bit PCM (sample / 32768, what soundfile.read hands psf.fbank, make_fbank.py:26-27): the host ships
 * raw audio (2 B/sample) and the whole front-end runs on the device. */
int sar_fbank_pcm16_fwd(const int16_t* pcm, const long long* offsets, const float* melfb_t,
                        float* feat_ws, float* x_data, int B, int Fmax, int T, void* stream);

/* ---- on-device batch assembly (the callers' side of the path: utils.data_loader, utils.py:71-117) ---------- */

/* x_data[b] = feat_reshape(feat_norm(feat_b), T) (utils.py:35-46,91): per-utterance, per-bin sklearn MinMaxScaler
 * over ALL frames of the utterance (zero range -> scale 1), then truncate / zero-pad to T frames.
 * feats (total frames, D) fp32: the un-padded feature matrices of the batch, concatenated (ONE upload);
 * frame_offsets (B+1) int64; x_data (B,T,D) output.  D <= 512. */
int sar_feat_batch_fwd(const float* feats, const long long* frame_offsets, float* x_data, int B, int T, int D, void* stream);

/* Label packing of data_loader: onehot (B,n_classes) = to_categorical(accent) (utils.py:100; null to skip);
 * ctc_label (B,Lmax) fp32 = text_ids_norm(trans_b, Lmax) -- truncate, pad with EOS_ID = 2 (utils.py:57-63,95; null to
 * skip), ctc_out_len (B) = min(len, Lmax), ctc_in_len (B) = encoder_len (utils.py:96-97).  trans: concatenated int32
 * token ids, trans_offsets (B+1) int64.  status (1) int32 optional: bit 0 set when an accent id is outside
 * [0, n_classes) (to_categorical raises). */
int sar_labels_pack_fwd(const int* accent, int n_classes, float* onehot,
                        const int* trans, const long long* trans_offsets, int Lmax, int encoder_len,
                        float* ctc_label, int* ctc_out_len, int* ctc_in_len, int* status, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SARNET_H_ */
