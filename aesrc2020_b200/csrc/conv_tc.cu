// conv_tc.cu -- residual-block convolutions of resnet.py on the 5th-gen tensor cores.
//
// Replaces _bn_relu_conv / basic_block / _shortcut (resnet.py:47-65, 105-125, 67-89): every 3x3
// convolution of the body (stride 1 and 2), with the 1x1 projection shortcut folded into the
// same accumulator and bias / identity-shortcut add / next-layer BN->ReLU fused in the epilogue.
//
// Formulation (implicit GEMM without im2col):
//  * Activations live in HBM as "flat-pad planes": an (H,W,C) map is stored as rows
//    q = n*(H+1)*(W+1) + h*(W+1) + w of C fp16 channels with ONE shared zero pad column / pad row
//    (the right pad of a row is the left pad of the next, the bottom pad row of an image is the top
//    pad row of the next).  A 3x3/s1 tap (kh,kw) of output row q then reads input row
//    q + (kh-1)*(W+1) + (kw-1): the A tile of tap t for 128 consecutive output rows is ONE 2-D TMA
//    box at a shifted row coordinate; rows before the tensor start are TMA zero fill.  Outputs
//    computed at pad positions are discarded by the epilogue (pads stay zero).
//  * A tensor that feeds a stride-2 block is stored phase-split (4 planes by (h&1, w&1), each a
//    flat-pad map with the OUTPUT geometry), which turns the stride-2 taps into stride-1 row
//    shifts on one phase plane each -- TF-SAME's asymmetric padding becomes a choice of phase and
//    shift per tap (computed on the host).
//  * fp32 accuracy on fp16 tensor cores: every tensor is a pair of fp16 planes, hi = fp16(x) and
//    lo = fp16((x - hi) * 2^11).  D = Ah*Wh accumulates in one TMEM accumulator and the cross
//    terms Al*Wh + Ah*Wl in a second one; the epilogue combines acc0 + 2^-11 * acc1 (the dropped
//    Al*Wl term is 2^-22 relative).  3 tcgen05.mma per k-step instead of 1.
//
// Kernel: persistent, one CTA per SM, 192 threads = TMA producer warp, MMA issuer warp (one
// elected lane issues tcgen05.mma, accumulators in TMEM, double buffered across tiles), and four
// epilogue warps (tcgen05.ld -> registers -> bias/residual/BN/ReLU -> hi/lo planes in HBM).
// 3-stage TMA->SMEM ring with mbarrier full/empty pairs, 128B/64B hardware swizzle.
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"

namespace sar {

constexpr int TC_BM = 128;            // output rows per tile (TMEM lanes)
constexpr int TC_STAGES = 8;          // generic kernel: most ring stages (barrier slots); the launch picks p.stages <= this
// -DSAR_BIASFOLD: the MODE 0 epilogue (conv1 of a block: only relu(bn(.)) leaves the item) finds the conv bias folded into
// the BN shift, relu(s acc + (s b + t)), and skips the bias add + its two shared-memory loads per 8 columns.  Same-box A/B
// (scripts/ab_so.sh, 3 runs each): conv segment 0.4865 vs 0.4878 ms at B=64, 2.75 vs 2.73 ms at B=512 -- neutral, so off.
#ifdef SAR_BIASFOLD
constexpr bool TC_BIASFOLD = true;
#else
constexpr bool TC_BIASFOLD = false;
#endif
constexpr int TC_THREADS = 320;            // 8 epilogue warps + TMA producer + MMA issuer
constexpr int TC_THREADS_MAX = 576;        // 16 epilogue warps: one warp per (quadrant, 32-column chunk) of a single 128-wide tile
// Warp roles.  The SM's schedulers favour the HIGHEST warp id of a sub-partition (wid % 4), so the two
// single-lane issuing warps get ids 4 and 5: they share sub-partitions 0 and 1 with epilogue warps 0 and 1
// and win arbitration while those spin on an mbarrier (with ids 0/1 the MMA warp crawled at ~600 cycles/tap).
constexpr int TC_WARP_TMA = 8, TC_WARP_MMA = 9;   // warps 0..7 = epilogue: quadrant = warp & 3, group = warp >> 2
// Two epilogue groups: group g drains TMEM accumulator stage g (tiles with (it & 1) == g), so one tile's
// conversion/stores may take up to two MMA tile times before the tensor pipe has to wait.
constexpr float TC_LO_SCALE = 2048.f;
constexpr float TC_LO_INV = 1.f / 2048.f;

struct TcParams {
  // K loop
  int ntaps, chunks_main, kc_main;       // main operand: ntaps * chunks_main k-steps of kc_main channels
  int chunks_sc, kc_sc, sc_plane;        // projection-shortcut operand: chunks_sc k-steps of kc_sc channels
  int tap_row_off[9], tap_plane[9];
  // tiles
  long long R;                           // flat output rows (incl. pads)
  int m_tiles, n_tiles, BN, Cout;
  // output geometry
  int B, H, W, P, Rimg;                  // P = W+1, Rimg = (H+1)*P
  int split, P2, Rimg2;                  // phase-split output geometry (consumer is a stride-2 block)
  long long R2;
  // epilogue
  const float* bias; const float* act_scale; const float* act_shift;
  const __half* res;                     // identity shortcut planes [2][R][Cout] or null
  __half* out_raw; __half* out_act; float* out_dense;
  long long* dbg;                        // optional: clock64 timestamps of CTA 0 (profiling aid)
  int dbg_it0;                           // first CTA-local item the epilogue stamps cover (chain: a later phase)
  int mma_mask;                          // experiment: which of the 3 hi/lo products to issue (7 = all)
  int act_kind;                          // activated outputs: 0 relu(scale*v+shift), 1 identity, 2 tanh(v)
  unsigned div_rimg_m, div_p_m;          // floor(x / Rimg), floor(x / P) for x < 2^31 as __umulhi(x, m) >> sh (0: divisor is 1)
  int div_rimg_sh, div_p_sh;
  int epi_alias;                         // every CTA owns ONE tile: the epilogue staging lives on top of the (then idle) operand region
  int nacc_log2;                         // TMEM accumulator stages = 1 << nacc_log2 (2; 4 in the slab kernel's 16-warp thin-tile form)
  int epi_warps;                         // slab kernel: epilogue warps the launch carries (8 or 16) -> staging bytes
  int stages;                            // generic kernel: depth of the TMA ring (2..TC_STAGES: whatever shared memory holds)
  int stage_bytes, a_plane_bytes;        // generic kernel: one ring stage = [A_hi | A_lo | B_hi ; B_lo], A planes a_plane_bytes apart
  int ksplit, ksteps_split, mn_tiles;    // generic kernel, split-K (1-tap GEMMs with a long K): tile = z * mn_tiles + (mt, nt)
  long long dense_zstride;               // elements between the fp32 partial outputs of consecutive K slices
  // chain kernel (several layers of one stage in ONE persistent launch, see conv_tc_chain_kernel)
  int* flag_done;                        // [m_tiles] epilogue items finished per M tile of THIS layer (or null)
  const int* flag_dep;                   // [m_tiles] the same counters of the layer this one reads (or null)
  int flag_need;                         // items per M tile = 4 * Cout / 32
  int epi_mode;                          // epilogue mode of this layer (-1, 0, 2, 3, 4)
  int tma_out;                           // plane outputs of an unsplit map leave through TMA stores (see epilogue_item)
  int pair;                              // conv_tc_pair_kernel: the accumulator-drained arrivals go to the LEADER CTA's barrier
  // The residual stream between the blocks of a stage as ONE fp32 plane [R][Cout] (flat-pad rows) instead of hi/lo
  // planes: it is only ever added in an epilogue (never an MMA operand), so the hi/lo split on write and the re-join
  // on read -- 16 of the ~18 instructions per element a conv2 epilogue spends on it -- are dropped.
  const float* res32;                    // identity shortcut, fp32 (instead of `res`)
  float* out_raw32;                      // raw sum, fp32 (instead of `out_raw`)
};
// tensor maps of the plane outputs, box = (32 channels, 32 rows, 1 plane), 64B swizzle
struct OutMaps { CUtensorMap raw, act; };

// ------------------------------------------------------------------ epilogue (shared by both kernels)
// Work item = (tile, 32-column chunk) handled by ONE warp.  The eight epilogue warps are two groups of four (one
// warp per TMEM lane quadrant); group g takes the items with (it * nchunks + chunk) % 2 == g, so the chunks of a
// single-tile CTA (late stages) drain in parallel and thin tiles (BN = 32) alternate between the groups.
//
// A row-per-thread global store touches 32 different cache lines per instruction (a row chunk is 64 B of a
// Cout*2-byte pitch; measured 2-4k cycles per chunk), so the warp transposes through shared memory: lane = row on
// the TMEM side (tcgen05.ld hands a lane one accumulator row), lane = (row, 16-byte piece) on the global side, so
// every LDG/STG instruction covers whole 64-byte (planes) or 128-byte (fp32) row chunks.  Staging tiles are
// XOR-swizzled by row, which makes both access patterns bank-conflict free.  The identity-shortcut chunk is
// requested BEFORE waiting for the accumulator (it does not depend on the MMAs) and sits in registers until then.
// Non-split plane outputs store zeros at pad positions (pads must stay zero); split planes and the dense output
// skip them.  The code is kept compact on purpose (rolled loops over 8-column groups): a CTA of the late stages
// runs it once or twice per warp, and the unrolled version was instruction-fetch bound (ncu: stall_no_inst).
constexpr int EPI_PLANE_BYTES = 32 * 64;           // 32 rows x 32 fp16 channels
constexpr int EPI_BUF_BYTES = 2 * EPI_PLANE_BYTES; // hi | lo planes, or 32 rows x 32 fp32
constexpr int EPI_WARP_BYTES = 2 * EPI_BUF_BYTES;  // R (shortcut in, raw out in place), A (activated planes or dense fp32)
constexpr int EPI_BYTES = 8 * EPI_WARP_BYTES;      // 64 KB

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t hh[4], ll[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 h2 = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((v[2 * e] - hf.x) * TC_LO_SCALE, (v[2 * e + 1] - hf.y) * TC_LO_SCALE);
    hh[e] = *reinterpret_cast<const uint32_t*>(&h2);
    ll[e] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  hi = make_uint4(hh[0], hh[1], hh[2], hh[3]);
  lo = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __noinline__ float tanh_precise(float x) { return tanhf(x); }

// MODE < 0: every output option is a run-time flag.  MODE >= 0 fixes the combination at compile time (the hot
// residual-block cases; the per-round flag tests were ~25 % of the executed instructions):
//   bit 0 = identity shortcut, bit 1 = raw planes out, the activated planes (BN->ReLU) are always written.
//   MODE 4 = the stage-ending conv2: fp32 shortcut stream in, raw hi/lo PLANES out (the next stage's projection operand)
//   + activated planes, both phase-split for the strided consumers (generic stores, shortcut row held in registers).
// CHAIN (conv_tc_chain_kernel): the bias / BN vectors of the item's 32 columns are staged from global memory into a
// per-warp area (s_bias_u, 32 floats apart), shortcut rows are read with ld.global.cg (another CTA wrote them
// during this very kernel), and the finished item is published in the layer's per-M-tile counter.
template <int MODE, bool CHAIN = false>
__device__ __forceinline__ void epilogue_item(const TcParams& p, const OutMaps& om, int tile, int c, int it, int quad, int lane,
                                           uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                           uint32_t s_bias_u, uint32_t stage_u, uint64_t* pub_bar = nullptr) {
  const int BN = p.BN;
  const int as = it & ((1 << p.nacc_log2) - 1);
  const uint32_t aphase = (uint32_t)(it >> p.nacc_log2) & 1u;
  const int dit = it - p.dbg_it0;
  if (p.dbg && blockIdx.x == 0 && lane == 0 && dit >= 0 && dit < 2 && c < 4) p.dbg[64 + ((dit * 4 + c) * 4 + quad) * 8] = clock64();
  const int z = tile / p.mn_tiles, tmn = tile - z * p.mn_tiles;         // K slice (split-K), tile within the M x N grid
  const int mt = tmn / p.n_tiles, nt = tmn - mt * p.n_tiles;
  const int n0 = nt * BN, c0 = c * 32;
  const long long row0 = (long long)mt * TC_BM + quad * 32;
  const bool res32 = MODE == 4 ? true : p.res32 != nullptr, raw32 = MODE == 4 ? false : p.out_raw32 != nullptr;
  const bool has_res = MODE < 0 ? (p.res != nullptr || res32) : (MODE == 4 || (MODE & 1) != 0);
  const bool out_raw = MODE < 0 ? (p.out_raw != nullptr || raw32) : (MODE == 4 || (MODE & 2) != 0);
  const bool out_act = MODE < 0 ? (p.out_act != nullptr) : true;
  const bool out_dense = MODE < 0 ? (p.out_dense != nullptr) : false;
  const int act_kind = MODE < 0 ? p.act_kind : 0;
  const uint32_t rb = stage_u, ab = stage_u + EPI_BUF_BYTES;
  // global side of the plane tiles: lane -> (row 8i + lane/4, 16-byte piece lane%4); slot = piece ^ ((row >> 1) & 3)
  const int g_row = lane >> 2, g_piece = lane & 3;
  const size_t g_col = (size_t)(n0 + c0 + 8 * g_piece);
  uint4 res_h[4], res_l[4];                          // planes: 4 hi + 4 lo pieces; fp32: pieces of rows 4i + lane/8
  if (has_res) {
    if (res32) {                                     // lane -> (row 4i + lane/8, 16-byte piece lane%8), 128-byte rows
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long q = row0 + 4 * i + (lane >> 3);
        uint4 r = make_uint4(0, 0, 0, 0);
        if (q < p.R) {
          const uint4* src = reinterpret_cast<const uint4*>(p.res32 + (size_t)q * p.Cout + n0 + c0 + 4 * (lane & 7));
          r = CHAIN ? __ldcg(src) : __ldg(src);
        }
        if (i < 4) res_h[i] = r; else res_l[i - 4] = r;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long q = row0 + 8 * i + g_row;
        res_h[i] = make_uint4(0, 0, 0, 0); res_l[i] = make_uint4(0, 0, 0, 0);
        if (q < p.R) {
          if (CHAIN) {
            res_h[i] = __ldcg(reinterpret_cast<const uint4*>(p.res + (size_t)q * p.Cout + g_col));
            res_l[i] = __ldcg(reinterpret_cast<const uint4*>(p.res + ((size_t)p.R + q) * p.Cout + g_col));
          } else {
            res_h[i] = __ldg(reinterpret_cast<const uint4*>(p.res + (size_t)q * p.Cout + g_col));
            res_l[i] = __ldg(reinterpret_cast<const uint4*>(p.res + ((size_t)p.R + q) * p.Cout + g_col));
          }
        }
      }
    }
  }
  if (CHAIN) {                                       // bias | scale | shift of columns n0+c0 .. +31 -> [3][32] floats
    const int col = n0 + c0 + lane;
    const float bv = p.bias ? __ldg(p.bias + col) : 0.f;
    const float sv = p.act_scale ? __ldg(p.act_scale + col) : 1.f;
    float tv = p.act_shift ? __ldg(p.act_shift + col) : 0.f;
    if (MODE == 0 && TC_BIASFOLD) tv = fmaf(bv, sv, tv);            // MODE 0: only relu(bn(.)) leaves the item -- the bias rides in the shift
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_bias_u + (uint32_t)lane * 4u), "f"(bv) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_bias_u + 128u + (uint32_t)lane * 4u), "f"(sv) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_bias_u + 256u + (uint32_t)lane * 4u), "f"(tv) : "memory");
    __syncwarp();
  }
  const uint32_t vstride = CHAIN ? 128u : (uint32_t)p.Cout * 4u;
  // TMEM side: lane -> row.  drow_p / drow_d: destination row of the plane / dense outputs; -1: not stored;
  // -2 - q: pad position of a non-split plane, row q is stored as zeros (pads must stay zero)
  const long long q = row0 + lane;
  int drow_p = -1, drow_d = -1;
  if (q < p.R) {
    const uint32_t qu = (uint32_t)q;
    const int n = p.div_rimg_m ? (int)(__umulhi(qu, p.div_rimg_m) >> p.div_rimg_sh) : (int)qu;
    const uint32_t rem = qu - (uint32_t)n * (uint32_t)p.Rimg;
    const int h = p.div_p_m ? (int)(__umulhi(rem, p.div_p_m) >> p.div_p_sh) : (int)rem;
    const int w = (int)rem - h * p.P;
    const bool valid = (h < p.H) && (w < p.W);
    if (p.split) drow_p = valid ? (int)((long long)(2 * ((h & 1) * 2 + (w & 1))) * p.R2 + (long long)n * p.Rimg2 + (h >> 1) * p.P2 + (w >> 1)) : -1;
    else drow_p = valid ? (int)q : -2 - (int)q;
    if (valid) drow_d = (n * p.H + h) * p.W + w;
  }
  const size_t lo_off = (size_t)(p.split ? p.R2 : p.R) * p.Cout;       // hi plane -> lo plane, in elements
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);
  const uint32_t rowoff = (uint32_t)lane * 64u;
  // TMA-store path (plane outputs of an unsplit map): the staging tiles ARE the 64B-swizzled boxes of the output
  // tensor maps (slot = piece ^ ((row >> 1) & 3) is CU_TENSOR_MAP_SWIZZLE_64B for 64-byte rows), so the write-out
  // is one cp.async.bulk.tensor store per plane instead of an LDS -> STG loop; pad rows are staged as zeros.
  const bool tma_out = MODE == 4 ? false : p.tma_out != 0;
  const bool zrow = tma_out && drow_p < 0;
  mbar_wait(&tfull_bar[as], aphase);
  tc_fence_after();
  if (tma_out) {                                     // the previous item's stores have finished reading the staging
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
  }
  // profiling aid (sar_tc_conv.dbg, >= 384 int64): CTA 0, per (quadrant, chunk) item of the first two tiles:
  // [item entered | accumulator ready | math + staging done | item finished]
#define EPI_STAMP(j) if (p.dbg && blockIdx.x == 0 && lane == 0 && dit >= 0 && dit < 2 && c < 4) p.dbg[64 + ((dit * 4 + c) * 4 + quad) * 8 + (j)] = clock64();
  EPI_STAMP(1)
  if (has_res) {                                     // shortcut pieces -> staging
    if (res32) {                                     // 32 rows x 128 B, piece j of a row in slot j ^ (row & 7)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t r_ = (uint32_t)(4 * i + (lane >> 3));
        sts128(rb + r_ * 128u + ((((uint32_t)lane & 7u) ^ (r_ & 7u)) << 4), i < 4 ? res_h[i] : res_l[i - 4]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = (uint32_t)(8 * i + g_row) * 64u + (((uint32_t)g_piece ^ (uint32_t)(((8 * i + g_row) >> 1) & 3)) << 4);
        sts128(rb + off, res_h[i]);
        sts128(rb + EPI_PLANE_BYTES + off, res_l[i]);
      }
    }
    __syncwarp();
  }
  // fp32 tile of the residual stream (res32 in, raw32 out, in place): my row, 16-byte pieces 2g and 2g + 1
  const uint32_t frow = rb + (uint32_t)lane * 128u;
  const uint32_t fx7 = (uint32_t)(lane & 7);
  // fp32 shortcut in, hi/lo PLANES out (a stage-ending conv2: its raw sum is the next stage's projection operand): the
  // plane tiles are staged on top of the fp32 tile, whose rows do not coincide with plane rows -- every lane first
  // takes its whole shortcut row into registers
  // ... and the mirror case, hi/lo plane shortcut in, fp32 raw sum out (the first block of a stage with an identity
  // shortcut on the stem's planes: res18 with 64 filters).  (The host routes both combinations to MODE -1.)
  const bool res_to_regs = MODE == 4 || (MODE < 0 && has_res && out_raw && (res32 != raw32));
  uint4 rrow[8];
  if (res_to_regs) {
    if (res32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) rrow[j] = lds128(frow + ((((uint32_t)j) ^ fx7) << 4));
    } else {                                         // planes: pieces 0..3 of my hi row, then of my lo row
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rrow[j] = lds128(rb + rowoff + ((((uint32_t)j) ^ sw) << 4));
        rrow[4 + j] = lds128(rb + EPI_PLANE_BYTES + rowoff + ((((uint32_t)j) ^ sw) << 4));
      }
    }
    __syncwarp();
  }
  const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 2 * BN + c0);
#pragma unroll(MODE == 4 ? 4 : 2)
  for (int g = 0; g < 4; ++g) {                      // 8 columns per round
    uint32_t r0[8], r1[8];
    tmem_ld8(tbase + (uint32_t)(8 * g), r0);
    tmem_ld8(tbase + (uint32_t)(BN + 8 * g), r1);
    tmem_ld_wait();
    if (g == 3) {                                    // the chunk has left TMEM: hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (p.pair) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));   // the pair's MMA issuer is CTA 0
        else mbar_arrive(&tempty_bar[as]);
      }
    }
    const uint32_t vec = s_bias_u + (uint32_t)((CHAIN ? 0 : n0 + c0) + 8 * g) * 4u;   // bias | scale | shift, vstride bytes apart
    float v[8];
    if (MODE == 0 && TC_BIASFOLD) {                  // relu(s (acc + b) + t) = relu(s acc + (s b + t)): the staging code folded b
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(__uint_as_float(r1[e]), TC_LO_INV, __uint_as_float(r0[e]));
    } else {
      const uint4 b0 = lds128(vec), b1 = lds128(vec + 16);
      const float bb[8] = {__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z), __uint_as_float(b0.w),
                           __uint_as_float(b1.x), __uint_as_float(b1.y), __uint_as_float(b1.z), __uint_as_float(b1.w)};
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(__uint_as_float(r1[e]), TC_LO_INV, __uint_as_float(r0[e])) + bb[e];
    }
    const uint32_t off = rowoff + (((uint32_t)g ^ sw) << 4);
    const uint32_t f0 = frow + ((((uint32_t)(2 * g)) ^ fx7) << 4), f1 = frow + ((((uint32_t)(2 * g + 1)) ^ fx7) << 4);
    if (has_res && res32) {
      uint4 a, b;
      if (res_to_regs) {                             // (g is a compile-time index only when the loop is unrolled)
        a = g == 0 ? rrow[0] : g == 1 ? rrow[2] : g == 2 ? rrow[4] : rrow[6];
        b = g == 0 ? rrow[1] : g == 1 ? rrow[3] : g == 2 ? rrow[5] : rrow[7];
      } else {
        a = lds128(f0); b = lds128(f1);
      }
      v[0] += __uint_as_float(a.x); v[1] += __uint_as_float(a.y); v[2] += __uint_as_float(a.z); v[3] += __uint_as_float(a.w);
      v[4] += __uint_as_float(b.x); v[5] += __uint_as_float(b.y); v[6] += __uint_as_float(b.z); v[7] += __uint_as_float(b.w);
    } else if (has_res) {
      uint4 rh, rl;
      if (res_to_regs) {
        rh = g == 0 ? rrow[0] : g == 1 ? rrow[1] : g == 2 ? rrow[2] : rrow[3];
        rl = g == 0 ? rrow[4] : g == 1 ? rrow[5] : g == 2 ? rrow[6] : rrow[7];
      } else {
        rh = lds128(rb + off); rl = lds128(rb + EPI_PLANE_BYTES + off);
      }
      const __half2* ah = reinterpret_cast<const __half2*>(&rh);
      const __half2* bl = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fh = __half22float2(ah[e]), fl = __half22float2(bl[e]);
        v[e * 2] += fmaf(fl.x, TC_LO_INV, fh.x);
        v[e * 2 + 1] += fmaf(fl.y, TC_LO_INV, fh.y);
      }
    }
    if (out_raw && raw32) {                          // in place (a lane only touches its own row); pads need no zeros:
      sts128(f0, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));   // the
      sts128(f1, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));   // stream is never an MMA operand
    } else if (out_raw) {                            // in place: a lane only reads and writes its own row here
      uint4 hi, lo;
      split8(v, hi, lo);
      if (zrow) { hi = make_uint4(0, 0, 0, 0); lo = hi; }
      sts128(rb + off, hi);
      sts128(rb + EPI_PLANE_BYTES + off, lo);
    }
    if (out_act || out_dense) {
      if (act_kind == 0) {
        const uint4 s0 = lds128(vec + vstride), s1 = lds128(vec + vstride + 16);
        const uint4 t0 = lds128(vec + 2u * vstride), t1 = lds128(vec + 2u * vstride + 16);
        const float ss[8] = {__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s0.z), __uint_as_float(s0.w),
                             __uint_as_float(s1.x), __uint_as_float(s1.y), __uint_as_float(s1.z), __uint_as_float(s1.w)};
        const float tt[8] = {__uint_as_float(t0.x), __uint_as_float(t0.y), __uint_as_float(t0.z), __uint_as_float(t0.w),
                             __uint_as_float(t1.x), __uint_as_float(t1.y), __uint_as_float(t1.z), __uint_as_float(t1.w)};
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = relu_nan(fmaf(v[e], ss[e], tt[e]));
      } else if (act_kind == 2) {                  // Dense(..., activation='tanh'), model.py:35-42
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = tanh_precise(v[e]);
      }
      if (out_act) {
        uint4 hi, lo;
        split8(v, hi, lo);
        if (zrow) { hi = make_uint4(0, 0, 0, 0); lo = hi; }
        sts128(ab + off, hi);
        sts128(ab + EPI_PLANE_BYTES + off, lo);
      } else {                                       // fp32 tile: 32 rows x 128 B, 16-byte piece j in slot j ^ (row & 7)
        const uint32_t drow = ab + (uint32_t)lane * 128u;
        const uint32_t x7 = (uint32_t)(lane & 7);
        sts128(drow + ((((uint32_t)(2 * g)) ^ x7) << 4), make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
        sts128(drow + ((((uint32_t)(2 * g + 1)) ^ x7) << 4), make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
      }
    }
  }
  EPI_STAMP(2)
  if (tma_out) fence_proxy_async();                  // my generic-proxy staging writes -> visible to the TMA engine
  __syncwarp();
  if (tma_out) {
    if (lane == 0) {
      const int cc = n0 + c0, cr = (int)row0;
      if (out_raw && raw32) tma_store_3d(&om.raw, rb, cc, cr, 0);          // one fp32 box, 128B swizzle
      else if (out_raw) { tma_store_3d(&om.raw, rb, cc, cr, 0); tma_store_3d(&om.raw, rb + EPI_PLANE_BYTES, cc, cr, 1); }
      if (out_act) { tma_store_3d(&om.act, ab, cc, cr, 0); tma_store_3d(&om.act, ab + EPI_PLANE_BYTES, cc, cr, 1); }
      bulk_commit();
      EPI_STAMP(4)
    }
  } else
  // write-out of the plane tiles: every store instruction covers 8 rows x 64 B
  if (out_raw || out_act) {
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const int rl_ = 8 * i + g_row;
      const int dr = __shfl_sync(0xffffffffu, drow_p, rl_);
      const uint32_t off = (uint32_t)rl_ * 64u + (((uint32_t)g_piece ^ (uint32_t)((rl_ >> 1) & 3)) << 4);
      if (dr != -1) {
        const bool zero = dr < 0;                       // pad position: the staged values are garbage, store zeros
        const size_t e0 = (size_t)(zero ? -2 - dr : dr) * p.Cout + g_col;
        const uint4 z4 = make_uint4(0, 0, 0, 0);
        if (out_raw && !raw32) {
          *reinterpret_cast<uint4*>(p.out_raw + e0) = zero ? z4 : lds128(rb + off);
          *reinterpret_cast<uint4*>(p.out_raw + e0 + lo_off) = zero ? z4 : lds128(rb + EPI_PLANE_BYTES + off);
        }
        if (out_act) {
          *reinterpret_cast<uint4*>(p.out_act + e0) = zero ? z4 : lds128(ab + off);
          *reinterpret_cast<uint4*>(p.out_act + e0 + lo_off) = zero ? z4 : lds128(ab + EPI_PLANE_BYTES + off);
        }
      }
    }
  }
  if (!tma_out && out_raw && raw32) {                // fp32 residual stream, generic stores: 4 rows x 128 B per instruction
    const int d_row = lane >> 3, d_piece = lane & 7;   // (pad rows carry whatever the MMA produced: never an MMA operand)
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const int rl_ = 4 * i + d_row;
      const long long qq = row0 + rl_;
      if (qq < p.R)
        *reinterpret_cast<uint4*>(p.out_raw32 + (size_t)qq * p.Cout + n0 + c0 + 4 * d_piece) =
            lds128(rb + (uint32_t)rl_ * 128u + ((((uint32_t)d_piece) ^ (uint32_t)(rl_ & 7)) << 4));
    }
  }
  if (out_dense) {                                   // 4 rows x 128 B per store instruction
    const int d_row = lane >> 3, d_piece = lane & 7;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const int rl_ = 4 * i + d_row;
      const int dr = __shfl_sync(0xffffffffu, drow_d, rl_);
      if (dr >= 0)
        *reinterpret_cast<uint4*>(p.out_dense + (size_t)z * p.dense_zstride + (size_t)dr * p.Cout + n0 + c0 + 4 * d_piece) =
            lds128(ab + (uint32_t)rl_ * 128u + ((((uint32_t)d_piece) ^ (uint32_t)(rl_ & 7)) << 4));
    }
  }
  __syncwarp();                                      // staging is reused by this warp's next item
  if (CHAIN && p.flag_done) {                        // publish: the rows this item wrote are visible GPU-wide
    if (lane == 0) {
      if (tma_out) { bulk_wait0(); asm volatile("fence.proxy.async;" ::: "memory"); }   // the bulk stores have landed
      EPI_STAMP(5)
      if (pub_bar) {
        // hand the tile to the CTA's publisher thread (conv_tc_chain_kernel): this warp's stores are ordered before
        // the arrive (release.cta; the lanes' stores by the __syncwarp above), the publisher's wait acquires them and
        // its ONE fence.acq_rel.gpu + counter bump per tile is cumulative over all eight warps' rows.  The warp does
        // not sit out the ~1.5k-cycle fence.  pub_bar[0..1]: stored (one arrival per item), pub_bar[2..3]: taken by
        // the publisher -- waited for before slot it & 1 is used again, so neither barrier can run two phases ahead.
        if (it >= 2) mbar_wait(&pub_bar[2 + as], (uint32_t)((it - 2) >> 1) & 1u);   // publisher took tile it - 2 off this slot
        mbar_arrive(&pub_bar[as]);
      } else {
        __threadfence();
        EPI_STAMP(6)
        atomicAdd(p.flag_done + mt, 1);
      }
    }
  }
  EPI_STAMP(3)
#undef EPI_STAMP
}

// all items of this CTA for one epilogue warp
template <int MODE>
__device__ __forceinline__ void epilogue_warp(const TcParams& p, const OutMaps& om, int total_tiles, int warp, int lane, int ngroups,
                                              uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar, const float* s_bias,
                                              uint8_t* epi_base) {
  const int quad = warp & 3, group = warp >> 2;
  const uint32_t stage_u = smem_u32(epi_base + (size_t)warp * EPI_WARP_BYTES);
  const uint32_t s_bias_u = smem_u32(s_bias);
  const int nchunks = p.BN >> 5;
  int it = 0, item = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it)
    for (int c = 0; c < nchunks; ++c, ++item)
      if (item % ngroups == group)
        epilogue_item<MODE>(p, om, tile, c, it, quad, lane, tmem_base, tfull_bar, tempty_bar, s_bias_u, stage_u);
  if (p.tma_out && lane == 0) bulk_wait0();          // my bulk stores are complete before the CTA may exit
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapS,
               const __grid_constant__ CUtensorMap mapWm, const __grid_constant__ CUtensorMap mapWs,
               const __grid_constant__ OutMaps om, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [stages x 64 KB ring][epilogue staging 64 KB unless aliased onto the ring][barriers][bias | scale | shift]
  const int nepi = (int)(blockDim.x >> 5) - 2;          // epilogue warps (8, or 16 for a single wide tile per CTA)
  const int WARP_TMA = nepi, WARP_MMA = nepi + 1;
  uint8_t* epi_own = smem + (size_t)p.stages * p.stage_bytes;
  uint8_t* epi_base = p.epi_alias ? smem : epi_own;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_own + (p.epi_alias ? 0 : EPI_BYTES));
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tfull_bar = empty_bar + TC_STAGES;      // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_base_slot + 2) + 15) & ~uintptr_t(15));   // [Cout] x 3
  float* s_scale = s_bias + p.Cout;
  float* s_shift = s_scale + p.Cout;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const uint32_t tmem_cols = (4 * BN <= 128) ? 128u : (4 * BN <= 256 ? 256u : 512u);   // 2 stages x (acc0, acc1)

  // short prologue: see conv_tc_slab_kernel
  if (warp == WARP_TMA) {
    if (lane < 2 * TC_STAGES + 4) {
      const bool is_tempty = lane >= 2 * TC_STAGES + 2;
      mbar_init(&full_bar[lane], is_tempty ? 4u * (uint32_t)(BN >> 5) : 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (lane == 0) {
      prefetch_tmap(&mapA); prefetch_tmap(&mapWm);
      if (p.chunks_sc) { prefetch_tmap(&mapS); prefetch_tmap(&mapWs); }
    }
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (warp < nepi) {
    for (int i = threadIdx.x; i < p.Cout; i += nepi * 32) {
      s_bias[i] = p.bias ? p.bias[i] : 0.f;
      s_scale[i] = p.act_scale ? p.act_scale[i] : 1.f;
      s_shift[i] = p.act_shift ? p.act_shift[i] : 0.f;
    }
    named_bar_sync(1, nepi * 32);
  }

  const int n_main = p.ksplit > 1 ? p.ksteps_split : p.ntaps * p.chunks_main;      // split-K: one tap, a slice of the chunks
  const int n_ksteps = n_main + p.chunks_sc;
  const int total_tiles = p.mn_tiles * p.ksplit;
  pdl_wait();          // everything above touched only constants (bias / BN affines) and on-chip state
  pdl_trigger();

  if (warp == WARP_TMA) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int z = tile / p.mn_tiles, tmn = tile - z * p.mn_tiles;
      const int mt = tmn / p.n_tiles, nt = tmn - mt * p.n_tiles;
      const long long q0 = (long long)mt * TC_BM;
      const int n0 = nt * BN;
      const int ch0 = z * p.ksteps_split;                  // first channel chunk of this K slice (0 without split-K)
      for (int ks = 0; ks < n_ksteps; ++ks) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          // stage = [A_hi | A_lo | B_hi ; B_lo]: the lo weight tile sits right behind the hi tile (one 2*BN-row operand)
          uint8_t* st = smem + (size_t)stage * p.stage_bytes;
          const uint32_t a_hi = smem_u32(st), a_lo = a_hi + (uint32_t)p.a_plane_bytes, b_hi = a_lo + (uint32_t)p.a_plane_bytes;
          const uint32_t b_lo = b_hi + (uint32_t)(BN * (ks < n_main ? p.kc_main : p.kc_sc) * 2);
          if (ks < n_main) {
            const int tap = (ch0 + ks) / p.chunks_main, ch = (ch0 + ks) - tap * p.chunks_main;
            const int kc = p.kc_main;
            mbar_expect_tx(&full_bar[stage], (uint32_t)(2 * (TC_BM + BN) * kc * 2));
            const int row = (int)(q0 + p.tap_row_off[tap]);
            tma_load_3d(&mapA, a_hi, &full_bar[stage], ch * kc, row, p.tap_plane[tap]);
            tma_load_3d(&mapA, a_lo, &full_bar[stage], ch * kc, row, p.tap_plane[tap] + 1);
            const int kofs = (ch0 + ks) * kc;
            tma_load_3d(&mapWm, b_hi, &full_bar[stage], kofs, n0, 0);
            tma_load_3d(&mapWm, b_lo, &full_bar[stage], kofs, n0, 1);
          } else {
            const int ch = ks - n_main;
            const int kc = p.kc_sc;
            mbar_expect_tx(&full_bar[stage], (uint32_t)(2 * (TC_BM + BN) * kc * 2));
            tma_load_3d(&mapS, a_hi, &full_bar[stage], ch * kc, (int)q0, p.sc_plane);
            tma_load_3d(&mapS, a_lo, &full_bar[stage], ch * kc, (int)q0, p.sc_plane + 1);
            const int kofs = n_main * p.kc_main + ch * kc;
            tma_load_3d(&mapWs, b_hi, &full_bar[stage], kofs, n0, 0);
            tma_load_3d(&mapWs, b_lo, &full_bar[stage], kofs, n0, 1);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    // instruction descriptor: D=f32, A=B=f16, K-major both, N=BN, M=128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc_2n = (1u << 4) | ((uint32_t)(BN >> 2) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tempty_bar[as], aphase ^ 1);            // epilogue drained this accumulator pair
      tc_fence_after();
      const uint32_t acc0 = tmem_base + (uint32_t)(as * 2 * BN);
      const uint32_t acc1 = acc0 + (uint32_t)BN;
      for (int ks = 0; ks < n_ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const int kc = (ks < n_main) ? p.kc_main : p.kc_sc;
          const int row_bytes = kc * 2;
          uint8_t* st = smem + (size_t)stage * p.stage_bytes;
          const uint32_t a_hi = smem_u32(st), a_lo = a_hi + (uint32_t)p.a_plane_bytes, b_hi = a_lo + (uint32_t)p.a_plane_bytes;
          const uint64_t dah = make_desc(a_hi, row_bytes), dal = make_desc(a_lo, row_bytes);
          const uint64_t dbh = make_desc(b_hi, row_bytes);
          for (int kk = 0; kk < kc / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);       // 32 bytes per UMMA_K=16 halves, in 16-byte units
            const uint32_t first = (ks | kk) ? 1u : 0u;
            umma_f16(acc0, dah + adv, dbh + adv, idesc_2n, first);     // [acc0 | acc1] (+)= A_hi x [B_hi ; B_lo]  (N = 2 BN)
            umma_f16(acc1, dal + adv, dbh + adv, idesc, 1u);           // acc1 += A_lo x B_hi
          }
          umma_commit(&empty_bar[stage]);                  // frees the smem stage when these MMAs retire
          if (ks == n_ksteps - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quadrant = warp % 4) =====================
    epilogue_warp<-1>(p, om, total_tiles, warp, lane, nepi >> 2, tmem_base, tfull_bar, tempty_bar, s_bias, epi_base);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------ the slab kernel (3x3 stride 1)
// Same roles and epilogue as conv_tc_kernel, but
//  * the A operand of a tile is loaded ONCE per 64-channel chunk as a halo'd slab of
//    128 + 2(W+1) + 2 consecutive flat rows; the 9 taps are 9 shared-memory matrix descriptors into
//    that slab, shifted by (kh-1)(W+1)+(kw-1) rows (the swizzle is a function of the absolute smem
//    address, so a row-shifted start is just a different address);
//  * weights are RESIDENT in shared memory for the whole kernel when the layer's packed kernel fits
//    (thin stages), otherwise they stream through a ring whose depth is whatever shared memory is
//    left (the k-step loop is TMA-latency bound, not bandwidth bound: depth is what matters);
//  * the 1x1 projection shortcut rides along as extra chunks with one tap.
constexpr int SL_MAX_RING = 16;
constexpr int SL_MAX_SLABS = 4;
constexpr int SL_THREADS = TC_THREADS;             // warps 0..7 epilogue (quadrant = warp & 3, group = warp >> 2), 8 TMA, 9 MMA
constexpr int SL_WARP_TMA = TC_WARP_TMA, SL_WARP_MMA = TC_WARP_MMA;

struct SlabParams {
  int slab_rows, lead;          // rows per main slab (multiple of 8), rows in front of q0 (= W + 2)
  int slab_bytes;               // bytes per plane of one slab buffer (1024-aligned)
  int nslab;                    // slab ring depth (2..4)
  int resident;                 // 1: all weight tiles stay in smem; 0: ring of `nring` stages
  int nring;
  int bplane_bytes;             // bytes of one weight tile plane (BN x kc_max x 2, 1024-aligned)
};

// Start address on ANY row boundary of a swizzled slab: measured on B200, the tensor core applies the
// 128B/64B XOR swizzle to the absolute shared-memory address bits (exactly like TMA wrote them), so a
// row-shifted start needs no base-offset field (setting (addr>>7)&7 there gives wrong results).
__device__ __forceinline__ uint64_t make_desc_shifted(uint32_t saddr, int row_bytes) { return make_desc(saddr, row_bytes); }

// (the generic epilogue, EPI_MODE -1, needs more registers than 576 threads leave: it keeps the 320-thread bound)
template <int KC, bool RESIDENT, int EPI_MODE>
__global__ void __launch_bounds__(EPI_MODE >= 0 ? TC_THREADS_MAX : TC_THREADS, 1)
conv_tc_slab_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapS,
                    const __grid_constant__ CUtensorMap mapWm, const __grid_constant__ CUtensorMap mapWs,
                    const __grid_constant__ OutMaps om, const TcParams p, const SlabParams sp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const long long t_entry = p.dbg ? clock64() : 0;
  const int nepi = (int)(blockDim.x >> 5) - 2;          // epilogue warps (8, or 16 for a single wide tile per CTA)
  const int WARP_TMA = nepi, WARP_MMA = nepi + 1;
  const int n_main = p.ntaps * p.chunks_main;
  const int n_ksteps = n_main + p.chunks_sc;
  const int nb = sp.resident ? n_ksteps : sp.nring;                    // weight slots in smem
  uint8_t* slab_base = smem;                                           // [nslab][2 planes][slab_bytes]
  uint8_t* b_base = smem + (size_t)sp.nslab * 2 * sp.slab_bytes;       // [nb][2 planes][bplane_bytes]
  // epilogue staging [8 warps][EPI_WARP_BYTES]: its own region, or (epi_alias: one tile per CTA) on top of the
  // operand region, which is idle once the tile's last MMA has retired
  uint8_t* epi_own = b_base + (size_t)nb * 2 * sp.bplane_bytes;
  uint8_t* epi_base = p.epi_alias ? smem : epi_own;
  uint64_t* sfull_bar = reinterpret_cast<uint64_t*>(epi_own + (p.epi_alias ? 0 : (size_t)p.epi_warps * EPI_WARP_BYTES));
  uint64_t* sempty_bar = sfull_bar + SL_MAX_SLABS;
  uint64_t* bfull_bar = sempty_bar + SL_MAX_SLABS;
  uint64_t* bempty_bar = bfull_bar + SL_MAX_RING;
  uint64_t* tfull_bar = bempty_bar + SL_MAX_RING;
  uint64_t* tempty_bar = tfull_bar + 4;                                // [4] (2 or 4 accumulator stages in use)
  uint64_t* rbar = tempty_bar + 4;                                     // [8 warps] shortcut chunk landed
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(rbar + 8);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_base_slot + 2) + 15) & ~uintptr_t(15));   // float4 reads
  float* s_scale = s_bias + p.Cout;
  float* s_shift = s_scale + p.Cout;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const int nacc = 1 << p.nacc_log2;
  const uint32_t tmem_cols = (2 * nacc * BN <= 128) ? 128u : (2 * nacc * BN <= 256 ? 256u : 512u);

  // Short prologue (it is on the critical path of every layer: the previous kernel's CTA must leave the SM before
  // this one starts): the 52 mbarriers are initialised one per lane, TMEM is allocated by the MMA warp, and only
  // the epilogue warps wait for the bias / BN vectors (they are idle until the first accumulator is ready).
  if (warp == WARP_TMA) {
    constexpr int NBAR = 2 * SL_MAX_SLABS + 2 * SL_MAX_RING + 4 + 4 + 8;       // contiguous from sfull_bar
    for (int i = lane; i < NBAR; i += 32) {
      const bool is_tempty = i >= 2 * SL_MAX_SLABS + 2 * SL_MAX_RING + 4 && i < 2 * SL_MAX_SLABS + 2 * SL_MAX_RING + 8;
      mbar_init(&sfull_bar[i], is_tempty ? 4u * (uint32_t)(p.BN >> 5) : 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (lane == 0) {
      prefetch_tmap(&mapA); prefetch_tmap(&mapWm);
      if (p.chunks_sc) { prefetch_tmap(&mapS); prefetch_tmap(&mapWs); }
    }
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (warp < nepi) {
    for (int i = threadIdx.x; i < p.Cout; i += nepi * 32) {
      s_bias[i] = p.bias ? p.bias[i] : 0.f;
      s_scale[i] = p.act_scale ? p.act_scale[i] : 1.f;
      s_shift[i] = p.act_shift ? p.act_shift[i] : 0.f;
      if (EPI_MODE == 0 && TC_BIASFOLD) s_shift[i] = fmaf(s_bias[i], s_scale[i], s_shift[i]);      // MODE 0 epilogue: bias folded into the BN shift
    }
    named_bar_sync(1, nepi * 32);
  }
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[61] = clock64();     // prologue done

  const int n_chunks = p.chunks_main + p.chunks_sc;       // slabs per tile
  const int total_tiles = p.m_tiles * p.n_tiles;
  // weight k-step index of (chunk c, tap): main steps are tap-major in K (k = tap*Cin + ci)
  auto kstep_of = [&](int c, int tap) { return c < p.chunks_main ? tap * p.chunks_main + c : n_main + (c - p.chunks_main); };
  auto kofs_of = [&](int c, int tap) {
    return c < p.chunks_main ? (tap * p.chunks_main + c) * p.kc_main : n_main * p.kc_main + (c - p.chunks_main) * p.kc_sc;
  };

  if (warp == WARP_TMA) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    {
      if (sp.resident) {
        // all weight tiles of this CTA's N tile (n_tiles == 1 in resident mode) under one barrier
        if (elect_one()) {
          uint32_t bytes = 0;
          for (int c = 0; c < n_chunks; ++c) bytes += (uint32_t)((c < p.chunks_main ? p.ntaps * p.kc_main : p.kc_sc) * 2 * BN * 2);
          mbar_expect_tx(&bfull_bar[0], bytes);
          for (int c = 0; c < n_chunks; ++c) {
            const bool main = c < p.chunks_main;
            const CUtensorMap* wm = main ? &mapWm : &mapWs;
            for (int tap = 0; tap < (main ? p.ntaps : 1); ++tap) {
              const uint32_t b_hi = smem_u32(b_base + (size_t)kstep_of(c, tap) * 2 * sp.bplane_bytes);
              const uint32_t lo_off = (uint32_t)(BN * (main ? p.kc_main : p.kc_sc) * 2);   // lo tile right behind the hi tile
              tma_load_3d(wm, b_hi, &bfull_bar[0], kofs_of(c, tap), 0, 0);
              tma_load_3d(wm, b_hi + lo_off, &bfull_bar[0], kofs_of(c, tap), 0, 1);
            }
          }
        }
        __syncwarp();
      }
      pdl_wait();                             // the resident weights (constants) are already in flight
      pdl_trigger();
      int sb = 0; uint32_t sphase = 0;        // slab ring
      int bs = 0; uint32_t bphase = 0;        // weight ring
      int exp_s = 0, exp_w = 0;               // (SAR_TC_MMAMASK bits 3/4: traffic experiments, wrong results)
      // One "step" = one slab (a 64/32-channel chunk of a tile) and the weight tiles of its taps.  The slab of step k+1
      // is requested after the first three weight tiles of step k, not after all nine (the weight ring lets this warp
      // run only nring taps ahead of the MMAs, so a slab requested behind the last tap is requested late).  Measured
      // at B=512: neutral to -1 % on the streamed-weight layers (stage 2: 683 -> 678 us) -- the slab was not what the
      // ~0.9k cycles between two tiles wait for.
      auto issue_slab = [&](int tile, int c) {
        const int mt = tile / p.n_tiles;
        const long long q0 = (long long)mt * TC_BM;
        const bool main = c < p.chunks_main;
        const int kc = main ? p.kc_main : p.kc_sc;
        const int rows = main ? sp.slab_rows : TC_BM;
        mbar_wait(&sempty_bar[sb], sphase ^ 1);
        if ((p.mma_mask & 16) && exp_s >= sp.nslab) {        // experiment: no slab traffic after the first ring fill
          if (elect_one()) mbar_arrive(&sfull_bar[sb]);
        } else
        if (elect_one()) {
          const uint32_t dst = smem_u32(slab_base + (size_t)sb * 2 * sp.slab_bytes);
          mbar_expect_tx(&sfull_bar[sb], (uint32_t)(2 * rows * kc * 2));
          if (main) {
            const int row0 = (int)(q0 - sp.lead);
            tma_load_3d(&mapA, dst, &sfull_bar[sb], c * kc, row0, 0);
            tma_load_3d(&mapA, dst + sp.slab_bytes, &sfull_bar[sb], c * kc, row0, 1);
          } else {
            const int ch = c - p.chunks_main;
            tma_load_3d(&mapS, dst, &sfull_bar[sb], ch * kc, (int)q0, p.sc_plane);
            tma_load_3d(&mapS, dst + sp.slab_bytes, &sfull_bar[sb], ch * kc, (int)q0, p.sc_plane + 1);
          }
        }
        __syncwarp();
        ++exp_s;
        if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
      };
      auto issue_weight = [&](int tile, int c, int tap) {
        const int nt = tile - (tile / p.n_tiles) * p.n_tiles;
        const int n0 = nt * BN;
        const bool main = c < p.chunks_main;
        const int kc = main ? p.kc_main : p.kc_sc;
        const CUtensorMap* wm = main ? &mapWm : &mapWs;
        mbar_wait(&bempty_bar[bs], bphase ^ 1);
        if ((p.mma_mask & 8) && exp_w++ >= sp.nring) {       // experiment: no weight traffic after the first ring fill
          if (elect_one()) mbar_arrive(&bfull_bar[bs]);
        } else
        if (elect_one()) {
          const uint32_t b_hi = smem_u32(b_base + (size_t)bs * 2 * sp.bplane_bytes);
          mbar_expect_tx(&bfull_bar[bs], (uint32_t)(2 * BN * kc * 2));
          tma_load_3d(wm, b_hi, &bfull_bar[bs], kofs_of(c, tap), n0, 0);
          tma_load_3d(wm, b_hi + (uint32_t)(BN * kc * 2), &bfull_bar[bs], kofs_of(c, tap), n0, 1);   // [B_hi ; B_lo] adjacent
        }
        __syncwarp();
        if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
      };
      int tile = blockIdx.x, c = 0;
      if (tile < total_tiles) issue_slab(tile, 0);
      while (tile < total_tiles) {
        int ntile = tile, nc = c + 1;
        if (nc == n_chunks) { nc = 0; ntile = tile + (int)gridDim.x; }
        const int ntaps = sp.resident ? 0 : (c < p.chunks_main ? p.ntaps : 1);
        const int early = ntaps < 3 ? ntaps : 3;
        for (int tap = 0; tap < early; ++tap) issue_weight(tile, c, tap);
        if (ntile < total_tiles) issue_slab(ntile, nc);
        for (int tap = early; tap < ntaps; ++tap) issue_weight(tile, c, tap);
        tile = ntile; c = nc;
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer: ONE elected lane runs the whole loop =====================
    // The loop is issue-bound (one thread, dependent uniform-datapath ops at ~8 cycles each), so the
    // instruction count per tcgen05.mma is what sets the speed of the thin layers:
    //  * taps and k16 steps fully unrolled, descriptors = constant hi word + ONE 32-bit add per operand;
    //  * the hi/lo products need only TWO instructions per k16 step: B_hi and B_lo tiles are adjacent in
    //    shared memory, so  [acc0 | acc1] (+)= A_hi x [B_hi ; B_lo]  is a single N = 2*BN MMA, followed by
    //    acc1 += A_lo x B_hi (N = BN).
    if (elect_one()) {
      const uint32_t idesc_n = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t idesc_2n = (1u << 4) | ((uint32_t)(BN >> 2) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      constexpr uint32_t ROWB = KC * 2, ROW16 = ROWB >> 4;
      constexpr uint32_t DHI = ((8 * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
      const uint32_t a_base0 = ((smem_u32(slab_base) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_slab16 = (uint32_t)(2 * sp.slab_bytes) >> 4, a_lo16 = (uint32_t)sp.slab_bytes >> 4;
      const uint32_t b_base0 = ((smem_u32(b_base) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_slot16 = (uint32_t)(2 * sp.bplane_bytes) >> 4;
      const uint32_t p_row16 = (uint32_t)p.P * ROW16;
      const uint32_t tap0_16 = (uint32_t)(sp.lead - p.P - 1) * ROW16;      // slab row of tap (0,0), in 16 B units
      int sb = 0; uint32_t sphase = 0;
      int bs = 0; uint32_t bphase = 0;
      int it = 0;
      if (RESIDENT) { mbar_wait(&bfull_bar[0], 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & (nacc - 1);
        const uint32_t aphase = (uint32_t)(it >> p.nacc_log2) & 1u;
        if (p.dbg && blockIdx.x == 0 && it < 8) p.dbg[it * 4] = clock64();
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && it < 8) p.dbg[it * 4 + 1] = clock64();
        const uint32_t acc0 = tmem_base + (uint32_t)(as * 2 * BN);
        const uint32_t acc1 = acc0 + (uint32_t)BN;
        uint32_t acc_on = 0;
        for (int c = 0; c < p.chunks_main; ++c) {
          mbar_wait(&sfull_bar[sb], sphase);
          tc_fence_after();
          if (p.dbg && blockIdx.x == 0 && it < 8 && c == 0) p.dbg[it * 4 + 2] = clock64();
          uint32_t a_row = a_base0 + (uint32_t)sb * a_slab16 + tap0_16;
          uint32_t b_res = b_base0 + (uint32_t)c * b_slot16;               // resident: slot of (tap 0, chunk c)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              uint32_t b0;
              if (RESIDENT) {
                b0 = b_res;
                b_res += (uint32_t)p.chunks_main * b_slot16;
              } else {
                mbar_wait(&bfull_bar[bs], bphase);
                tc_fence_after();
                b0 = b_base0 + (uint32_t)bs * b_slot16;
              }
              const uint32_t a0 = a_row + (uint32_t)kw * ROW16;
#pragma unroll
              for (int kk = 0; kk < KC / 16; ++kk) {
                const uint64_t dah = ((uint64_t)DHI << 32) | (a0 + 2u * kk), dal = ((uint64_t)DHI << 32) | (a0 + a_lo16 + 2u * kk);
                const uint64_t dbh = ((uint64_t)DHI << 32) | (b0 + 2u * kk);
                umma_f16(acc0, dah, dbh, idesc_2n, acc_on);                  // [acc0|acc1] (+)= Ah x [Bh;Bl]
                umma_f16(acc1, dal, dbh, idesc_n, 1u);                       // acc1 += Al x Bh
                acc_on = 1u;
              }
              if (!RESIDENT) {
                umma_commit(&bempty_bar[bs]);
                if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
              }
            }
            a_row += p_row16;
          }
          umma_commit(&sempty_bar[sb]);                     // slab free once its 9 taps have retired
          if (p.chunks_sc == 0 && c == p.chunks_main - 1) umma_commit(&tfull_bar[as]);
          if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
        }
        // 1x1 projection shortcut: chunks of kc_sc channels, one tap, no shift
        for (int c = 0; c < p.chunks_sc; ++c) {
          const uint32_t rowb = (uint32_t)p.kc_sc * 2;
          const uint32_t dhi_s = ((8 * rowb) >> 4) | (1u << 14) | ((rowb == 128 ? 2u : 4u) << 29);
          mbar_wait(&sfull_bar[sb], sphase);
          tc_fence_after();
          const uint32_t a0 = a_base0 + (uint32_t)sb * a_slab16;
          uint32_t b0;
          if (RESIDENT) b0 = b_base0 + (uint32_t)(n_main + c) * b_slot16;
          else { mbar_wait(&bfull_bar[bs], bphase); tc_fence_after(); b0 = b_base0 + (uint32_t)bs * b_slot16; }
          for (int kk = 0; kk < p.kc_sc / 16; ++kk) {
            const uint64_t dah = ((uint64_t)dhi_s << 32) | (a0 + 2u * kk), dal = ((uint64_t)dhi_s << 32) | (a0 + a_lo16 + 2u * kk);
            const uint64_t dbh = ((uint64_t)dhi_s << 32) | (b0 + 2u * kk);
            umma_f16(acc0, dah, dbh, idesc_2n, 1u);
            umma_f16(acc1, dal, dbh, idesc_n, 1u);
          }
          if (!RESIDENT) {
            umma_commit(&bempty_bar[bs]);
            if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
          }
          umma_commit(&sempty_bar[sb]);
          if (c == p.chunks_sc - 1) umma_commit(&tfull_bar[as]);
          if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
        }
        if (p.dbg && blockIdx.x == 0 && it < 8) p.dbg[it * 4 + 3] = clock64();
      }
    }
    __syncwarp();
  } else {
    pdl_wait();                               // identity-shortcut rows and the output planes
    epilogue_warp<EPI_MODE>(p, om, total_tiles, warp, lane, nepi >> 2, tmem_base, tfull_bar, tempty_bar, s_bias, epi_base);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) { p.dbg[62] = t_entry; p.dbg[63] = clock64(); }
}

// ------------------------------------------------------------------ the pair kernel (3x3 stride 1, cta_group::2)
// conv_tc_slab_kernel's non-resident form on CTA PAIRS: a cluster of two CTAs (one TPC) owns two vertically adjacent
// M tiles (256 output rows) of one N tile.  Each CTA loads the halo'd A slab of ITS 128 rows and only HALF of every
// weight tile (rows [BN/2 * rank, BN/2 * rank + BN/2) of B_hi and of B_lo); the leader issues M = 256
// tcgen05.mma.cta_group::2 instructions that read both CTAs' shared memory and write each CTA's 128 accumulator rows
// into its own tensor memory.  Per SM the weight stream -- the larger part of the L2 -> shared-memory traffic of the
// stage 2-4 layers (590 KB of a 662 KB tile in stage 3) -- is halved.
// Per k16 step the hi/lo products are three N = BN instructions (the [B_hi ; B_lo] 2N trick of the slab kernel would
// put all of B_hi in one CTA and all of B_lo in the other, and the A_lo x B_hi product needs B_hi from both):
//   acc0 (+)= A_hi x B_hi,  acc1 (+)= A_hi x B_lo,  acc1 += A_lo x B_hi.
// Protocol (CUTLASS' 2-SM pipeline): "full" barriers live in the leader -- its producer posts the bytes of BOTH CTAs'
// loads, both CTAs' TMA loads complete on it; "empty" / "accumulator ready" arrive in both CTAs through multicast
// tcgen05.commit; "accumulator drained" collects the epilogue warps of both CTAs on the leader's barrier.
template <int EPI_MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EPI_MODE >= 0 ? TC_THREADS_MAX : TC_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapS,
                    const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ OutMaps om, const TcParams p,
                    const SlabParams sp) {
  constexpr int KC = 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nepi = (int)(blockDim.x >> 5) - 2;
  const int WARP_TMA = nepi, WARP_MMA = nepi + 1;
  const int BN = p.BN;
  const uint32_t bhalf = (uint32_t)(BN / 2) * KC * 2;                  // bytes of my half of one weight plane tile
  uint8_t* slab_base = smem;                                           // [nslab][2 planes][slab_bytes]
  uint8_t* b_base = smem + (size_t)sp.nslab * 2 * sp.slab_bytes;       // [nring][B_hi half | B_lo half]
  uint8_t* epi_own = b_base + (size_t)sp.nring * 2 * bhalf;
  uint8_t* epi_base = p.epi_alias ? smem : epi_own;
  uint64_t* sfull_bar = reinterpret_cast<uint64_t*>(epi_own + (p.epi_alias ? 0 : (size_t)p.epi_warps * EPI_WARP_BYTES));
  uint64_t* sempty_bar = sfull_bar + SL_MAX_SLABS;
  uint64_t* bfull_bar = sempty_bar + SL_MAX_SLABS;
  uint64_t* bempty_bar = bfull_bar + SL_MAX_RING;
  uint64_t* tfull_bar = bempty_bar + SL_MAX_RING;
  uint64_t* tempty_bar = tfull_bar + 4;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty_bar + 4 + 8);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_base_slot + 2) + 15) & ~uintptr_t(15));
  float* s_scale = s_bias + p.Cout;
  float* s_shift = s_scale + p.Cout;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t tmem_cols = (4 * BN <= 128) ? 128u : (4 * BN <= 256 ? 256u : 512u);   // 2 stages x (acc0, acc1)

  if (warp == WARP_TMA) {
    constexpr int NBAR = 2 * SL_MAX_SLABS + 2 * SL_MAX_RING + 4 + 4;
    for (int i = lane; i < NBAR; i += 32) {
      const bool is_tempty = i >= 2 * SL_MAX_SLABS + 2 * SL_MAX_RING + 4;
      mbar_init(&sfull_bar[i], is_tempty ? 2u * 4u * (uint32_t)(BN >> 5) : 1u);      // drained: both CTAs' epilogue items
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (lane == 0) { prefetch_tmap(&mapA); prefetch_tmap(&mapWh); if (p.chunks_sc) prefetch_tmap(&mapS); }
  }
  if (warp == WARP_MMA) {                       // the same warp id in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                           // both CTAs' barriers are initialised before anything crosses over
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (warp < nepi) {
    for (int i = threadIdx.x; i < p.Cout; i += nepi * 32) {
      s_bias[i] = p.bias ? p.bias[i] : 0.f;
      s_scale[i] = p.act_scale ? p.act_scale[i] : 1.f;
      s_shift[i] = p.act_shift ? p.act_shift[i] : 0.f;
      if (EPI_MODE == 0 && TC_BIASFOLD) s_shift[i] = fmaf(s_bias[i], s_scale[i], s_shift[i]);      // MODE 0 epilogue: bias folded into the BN shift
    }
    named_bar_sync(1, nepi * 32);
  }

  const int pairs_m = (p.m_tiles + 1) >> 1;
  const int total_pairs = pairs_m * p.n_tiles;
  const int ncl = (int)(gridDim.x >> 1), cl = (int)(blockIdx.x >> 1);

  if (warp == WARP_TMA) {
    // ===================== TMA producer (both CTAs) =====================
    pdl_wait();
    pdl_trigger();
    int sb = 0; uint32_t sphase = 0;
    int bs = 0; uint32_t bphase = 0;
    const int n_chunks = p.chunks_main + p.chunks_sc;
    // (step order as in conv_tc_slab_kernel: the next step's slab goes out after three weight tiles of this one)
    auto issue_slab = [&](int pr, int c) {
      const int pm = pr / p.n_tiles;
      const int mt = 2 * pm + (int)rank;                     // (an odd tail's second tile lies past the tensor: zero fill)
      const long long q0 = (long long)mt * TC_BM;
      const bool main = c < p.chunks_main;
      mbar_wait(&sempty_bar[sb], sphase ^ 1);
      if (elect_one()) {
        const uint32_t dst = smem_u32(slab_base + (size_t)sb * 2 * sp.slab_bytes);
        if (leader) mbar_expect_tx(&sfull_bar[sb], 2u * (uint32_t)(2 * (main ? sp.slab_rows : TC_BM) * KC * 2));
        const uint32_t bar = mapa_u32(smem_u32(&sfull_bar[sb]), 0);
        if (main) {
          const int row0 = (int)(q0 - sp.lead);
          tma_load_3d_2sm(&mapA, dst, bar, c * KC, row0, 0);
          tma_load_3d_2sm(&mapA, dst + sp.slab_bytes, bar, c * KC, row0, 1);
        } else {                                             // 1x1 projection shortcut (kc_sc == 64): my 128 rows, no halo
          const int ch = c - p.chunks_main;
          tma_load_3d_2sm(&mapS, dst, bar, ch * KC, (int)q0, p.sc_plane);
          tma_load_3d_2sm(&mapS, dst + sp.slab_bytes, bar, ch * KC, (int)q0, p.sc_plane + 1);
        }
      }
      __syncwarp();
      if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
    };
    auto issue_weight = [&](int pr, int c, int tap) {
      const int nt = pr - (pr / p.n_tiles) * p.n_tiles;
      const int n0 = nt * BN + (int)rank * (BN / 2);         // my half of the weight rows
      const int kofs = (c < p.chunks_main ? tap * p.chunks_main + c : 9 * p.chunks_main + (c - p.chunks_main)) * KC;
      mbar_wait(&bempty_bar[bs], bphase ^ 1);
      if (elect_one()) {
        const uint32_t b_hi = smem_u32(b_base + (size_t)bs * 2 * bhalf);
        if (leader) mbar_expect_tx(&bfull_bar[bs], 2u * 2u * bhalf);
        const uint32_t bar = mapa_u32(smem_u32(&bfull_bar[bs]), 0);
        tma_load_3d_2sm(&mapWh, b_hi, bar, kofs, n0, 0);
        tma_load_3d_2sm(&mapWh, b_hi + bhalf, bar, kofs, n0, 1);
      }
      __syncwarp();
      if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
    };
    int pr = cl, c = 0;
    if (pr < total_pairs) issue_slab(pr, 0);
    while (pr < total_pairs) {
      int npr = pr, nc = c + 1;
      if (nc == n_chunks) { nc = 0; npr = pr + ncl; }
      const int ntaps = c < p.chunks_main ? 9 : 1;
      const int early = ntaps < 3 ? ntaps : 3;
      for (int tap = 0; tap < early; ++tap) issue_weight(pr, c, tap);
      if (npr < total_pairs) issue_slab(npr, nc);
      for (int tap = early; tap < ntaps; ++tap) issue_weight(pr, c, tap);
      pr = npr; c = nc;
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer: one elected lane of the LEADER =====================
    if (leader && elect_one()) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256 over the pair
      constexpr uint32_t ROWB = KC * 2, ROW16 = ROWB >> 4;
      constexpr uint32_t DHI = ((8 * ROWB) >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_base0 = ((smem_u32(slab_base) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_slab16 = (uint32_t)(2 * sp.slab_bytes) >> 4, a_lo16 = (uint32_t)sp.slab_bytes >> 4;
      const uint32_t b_base0 = ((smem_u32(b_base) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_slot16 = (2u * bhalf) >> 4, b_lo16 = bhalf >> 4;
      const uint32_t p_row16 = (uint32_t)p.P * ROW16;
      const uint32_t tap0_16 = (uint32_t)(sp.lead - p.P - 1) * ROW16;
      int sb = 0; uint32_t sphase = 0;
      int bs = 0; uint32_t bphase = 0;
      int it = 0;
      for (int pr = cl; pr < total_pairs; pr += ncl, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait_cluster(&tempty_bar[as], aphase ^ 1);   // drained by the epilogue warps of BOTH CTAs
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(as * 2 * BN);
        const uint32_t acc1 = acc0 + (uint32_t)BN;
        uint32_t acc_on = 0;
        for (int c = 0; c < p.chunks_main; ++c) {
          mbar_wait_cluster(&sfull_bar[sb], sphase);
          tc_fence_after();
          uint32_t a_row = a_base0 + (uint32_t)sb * a_slab16 + tap0_16;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              mbar_wait_cluster(&bfull_bar[bs], bphase);
              tc_fence_after();
              const uint32_t b0 = b_base0 + (uint32_t)bs * b_slot16;
              const uint32_t a0 = a_row + (uint32_t)kw * ROW16;
#pragma unroll
              for (int kk = 0; kk < KC / 16; ++kk) {
                const uint64_t dah = ((uint64_t)DHI << 32) | (a0 + 2u * kk), dal = ((uint64_t)DHI << 32) | (a0 + a_lo16 + 2u * kk);
                const uint64_t dbh = ((uint64_t)DHI << 32) | (b0 + 2u * kk), dbl = ((uint64_t)DHI << 32) | (b0 + b_lo16 + 2u * kk);
                umma_f16_2sm(acc0, dah, dbh, idesc, acc_on);          // acc0 (+)= A_hi x B_hi
                umma_f16_2sm(acc1, dah, dbl, idesc, acc_on);          // acc1 (+)= A_hi x B_lo
                umma_f16_2sm(acc1, dal, dbh, idesc, 1u);              // acc1  += A_lo x B_hi
                acc_on = 1u;
              }
              umma_commit_2sm(&bempty_bar[bs]);
              if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
            }
            a_row += p_row16;
          }
          umma_commit_2sm(&sempty_bar[sb]);
          if (p.chunks_sc == 0 && c == p.chunks_main - 1) umma_commit_2sm(&tfull_bar[as]);
          if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
        }
        for (int c = 0; c < p.chunks_sc; ++c) {           // 1x1 projection shortcut: one tap, rows from slab row 0
          mbar_wait_cluster(&sfull_bar[sb], sphase);
          tc_fence_after();
          mbar_wait_cluster(&bfull_bar[bs], bphase);
          tc_fence_after();
          const uint32_t a0 = a_base0 + (uint32_t)sb * a_slab16;
          const uint32_t b0 = b_base0 + (uint32_t)bs * b_slot16;
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            const uint64_t dah = ((uint64_t)DHI << 32) | (a0 + 2u * kk), dal = ((uint64_t)DHI << 32) | (a0 + a_lo16 + 2u * kk);
            const uint64_t dbh = ((uint64_t)DHI << 32) | (b0 + 2u * kk), dbl = ((uint64_t)DHI << 32) | (b0 + b_lo16 + 2u * kk);
            umma_f16_2sm(acc0, dah, dbh, idesc, 1u);
            umma_f16_2sm(acc1, dah, dbl, idesc, 1u);
            umma_f16_2sm(acc1, dal, dbh, idesc, 1u);
          }
          umma_commit_2sm(&bempty_bar[bs]);
          if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
          umma_commit_2sm(&sempty_bar[sb]);
          if (c == p.chunks_sc - 1) umma_commit_2sm(&tfull_bar[as]);
          if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (both CTAs, each drains its own 128 accumulator rows) =====================
    pdl_wait();
    const int quad = warp & 3, group = warp >> 2, ngroups = nepi >> 2;
    const uint32_t stage_u = smem_u32(epi_base + (size_t)warp * EPI_WARP_BYTES);
    const uint32_t s_bias_u = smem_u32(s_bias);
    const int nchunks = BN >> 5;
    int it = 0, item = 0;
    for (int pr = cl; pr < total_pairs; pr += ncl, ++it) {
      const int pm = pr / p.n_tiles, nt = pr - pm * p.n_tiles;
      const int mt = 2 * pm + (int)rank;
      for (int c = 0; c < nchunks; ++c, ++item) {
        if (item % ngroups != group) continue;
        if (mt < p.m_tiles) {
          epilogue_item<EPI_MODE>(p, om, mt * p.n_tiles + nt, c, it, quad, lane, tmem_base, tfull_bar, tempty_bar, s_bias_u, stage_u);
        } else {                                  // the tile past an odd tail: nothing to store, the barrier protocol still runs
          const int as = it & 1;
          mbar_wait(&tfull_bar[as], (uint32_t)(it >> 1) & 1u);
          tc_fence_after();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
        }
      }
    }
    if (p.tma_out && lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                           // the peer's tensor memory / shared memory are in use until both are done
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// ------------------------------------------------------------------ the chain kernel
// All stride-1 3x3 convolutions of one ResNet stage (conv2 of the first block, then conv1/conv2 of every further
// block) as ONE persistent launch.  At small batches a stage-3/4 layer has fewer tiles than the chip has SMs and
// pays its own launch, prologue and epilogue tail; here a CTA walks the layers ("phases") back to back, the
// epilogue of one phase overlaps the mainloop of the next, and the M x N tiles of all phases fill all SMs.
// There is no grid-wide barrier: a 3x3 tile of layer l reads the rows of M tiles m-1, m, m+1 of layer l-1, so the
// TMA producer waits for those three per-M-tile counters (bumped by the epilogue warps with a release pattern:
// stores, __syncwarp, __threadfence, atomicAdd; read with ld.acquire + fence.proxy.async before the TMA load).
// The (phase, tile) items of the whole chain are numbered phase-major and dealt round-robin (item = blockIdx.x +
// k * gridDim.x), so a phase whose tile count is not a multiple of the grid (160 or 315 tiles on 148 SMs) costs
// tiles / grid rounds instead of ceil(tiles / grid) -- with a per-phase deal the CTAs holding the extra tile set
// the pace of the whole chain.  Every CTA walks its items in increasing order, all CTAs are resident (cooperative
// launch, grid <= #SMs) and an item only waits for items with a smaller number: no deadlock.  An item of phase
// l + 1 starts after the three M tiles of phase l that read its rows have finished, so the read-after-write
// flags also order the write-after-read reuse of the two-slot activation buffers.
// Weights stream through the TMA ring (non-resident form of conv_tc_slab_kernel), epilogue staging has its own
// 64 KB, the per-item bias / BN vectors a 3 KB per-warp area.
constexpr int CH_MAX = 12;
constexpr int CH_DBG_ITEMS = 32, CH_DBG_STRIDE = 8 + CH_DBG_ITEMS * 8;
// profiling aid: [cta][0] clock64 at start, [1] globaltimer at start, [2] clock64 at exit; per item it < 32 at
// 8 + it * 8: [0] item number + 1, [1] producer reached the item, [2] dependencies met, [3] first slab landed (MMA warp),
// [4] all MMAs issued, [5] epilogue warp 0/4 entered chunk 0, [6] chunk 0 published, [7] last chunk published
#define CH_STAMP(itv, j) if (cp.dbg && (itv) < CH_DBG_ITEMS) cp.dbg[(size_t)blockIdx.x * CH_DBG_STRIDE + 8 + (itv) * 8 + (j)] = clock64();
struct ChainPhase { CUtensorMap mapA, mapS, mapWm, mapWs; OutMaps om; TcParams p; };
constexpr int CH_WARP_PUB = SL_WARP_MMA + 1;        // publisher: one fence + counter bump per finished tile
constexpr int CH_THREADS = SL_THREADS + 32;
struct ChainParams {
  int n_phases;
  int publisher;                 // 1: epilogue warps hand finished tiles to the publisher thread; 0: every warp fences itself
  int* flags;                    // [n_phases][m_tiles] counters + [1] CTA exit counter (self-cleaning)
  long long* dbg;                // optional (descs[0].dbg): per CTA CH_DBG_STRIDE int64 of clock64 stamps, see scripts/chain_dbg.py
  SlabParams sp;
  ChainPhase ph[CH_MAX];
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int KC>
__global__ void __launch_bounds__(CH_THREADS, 1) conv_tc_chain_kernel(const __grid_constant__ ChainParams cp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const SlabParams& sp = cp.sp;
  const TcParams& g0 = cp.ph[0].p;                                     // geometry / tiling: identical in every phase
  uint8_t* slab_base = smem;                                           // [nslab][2 planes][slab_bytes]
  uint8_t* b_base = smem + (size_t)sp.nslab * 2 * sp.slab_bytes;       // [nring][2 planes][bplane_bytes]
  uint8_t* epi_base = b_base + (size_t)sp.nring * 2 * sp.bplane_bytes; // [8 warps][EPI_WARP_BYTES]
  uint8_t* vec_base = epi_base + EPI_BYTES;                            // [8 warps][3][32] floats
  uint64_t* sfull_bar = reinterpret_cast<uint64_t*>(vec_base + 8 * 384);
  uint64_t* sempty_bar = sfull_bar + SL_MAX_SLABS;
  uint64_t* bfull_bar = sempty_bar + SL_MAX_SLABS;
  uint64_t* bempty_bar = bfull_bar + SL_MAX_RING;
  uint64_t* tfull_bar = bempty_bar + SL_MAX_RING;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* pub_bar = tempty_bar + 2;                                  // [2] tile it (slot it & 1) stored by all its items, [2] taken
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(pub_bar + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g0.BN;
  const uint32_t tmem_cols = (4 * BN <= 128) ? 128u : (4 * BN <= 256 ? 256u : 512u);
  const int total_tiles = g0.m_tiles * g0.n_tiles;
  const int n_items = cp.n_phases * total_tiles;                       // item = phase * total_tiles + tile, dealt round-robin

  if (warp == SL_WARP_TMA) {
    constexpr int NBAR = 2 * SL_MAX_SLABS + 2 * SL_MAX_RING + 2 + 2 + 4;
    for (int i = lane; i < NBAR; i += 32) {
      const bool per_item = i >= NBAR - 6 && i < NBAR - 2;             // tempty, pub stored: one arrival per epilogue item of a tile
      mbar_init(&sfull_bar[i], per_item ? 4u * (uint32_t)(BN >> 5) : 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (lane == 0) { prefetch_tmap(&cp.ph[0].mapA); prefetch_tmap(&cp.ph[0].mapWm); }
  }
  if (warp == SL_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();
  pdl_trigger();
  if (cp.dbg && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    cp.dbg[(size_t)blockIdx.x * CH_DBG_STRIDE] = clock64();
    cp.dbg[(size_t)blockIdx.x * CH_DBG_STRIDE + 1] = (long long)gt;
  }

  if (warp == SL_WARP_TMA) {
    // ===================== TMA producer =====================
    int sb = 0; uint32_t sphase = 0;        // slab ring
    int bs = 0; uint32_t bphase = 0;        // weight ring
    int ph_seen = -1, pit = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int ph = item / total_tiles, tile = item - ph * total_tiles;
      const ChainPhase& P = cp.ph[ph];
      const TcParams& p = P.p;
      const int n_main = p.ntaps * p.chunks_main;
      const int n_chunks = p.chunks_main + p.chunks_sc;
      if (ph != ph_seen) {
        ph_seen = ph;
        if (lane == 0 && ph + 1 < cp.n_phases) { prefetch_tmap(&cp.ph[ph + 1].mapA); prefetch_tmap(&cp.ph[ph + 1].mapWm); }
      }
      {
        const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
        const long long q0 = (long long)mt * TC_BM;
        const int n0 = nt * BN;
        if (lane == 0) {
          if (cp.dbg && pit < CH_DBG_ITEMS) cp.dbg[(size_t)blockIdx.x * CH_DBG_STRIDE + 8 + pit * 8] = item + 1;
          CH_STAMP(pit, 1)
        }
        if (p.flag_dep) {                    // rows of M tiles mt-1 .. mt+1 of the previous layer must be complete
          const int lo = mt > 0 ? mt - 1 : 0, hi = mt + 1 < p.m_tiles ? mt + 1 : p.m_tiles - 1;
          const long long t0 = clock64();
          for (int mm = lo; mm <= hi; ++mm)
            while (ld_acquire_gpu(p.flag_dep + mm) < p.flag_need)
              if (clock64() - t0 > 4000000000LL) __trap();     // a protocol bug traps instead of hanging the GPU
          asm volatile("fence.proxy.async;" ::: "memory");
          __syncwarp();
        }
        if (lane == 0) { CH_STAMP(pit, 2) }
        ++pit;
        for (int c = 0; c < n_chunks; ++c) {
          const bool main = c < p.chunks_main;
          const int kc = main ? p.kc_main : p.kc_sc;
          const int rows = main ? sp.slab_rows : TC_BM;
          mbar_wait(&sempty_bar[sb], sphase ^ 1);
          if (elect_one()) {
            const uint32_t dst = smem_u32(slab_base + (size_t)sb * 2 * sp.slab_bytes);
            mbar_expect_tx(&sfull_bar[sb], (uint32_t)(2 * rows * kc * 2));
            if (main) {
              const int row0 = (int)(q0 - sp.lead);
              tma_load_3d(&P.mapA, dst, &sfull_bar[sb], c * kc, row0, 0);
              tma_load_3d(&P.mapA, dst + sp.slab_bytes, &sfull_bar[sb], c * kc, row0, 1);
            } else {
              const int ch = c - p.chunks_main;
              tma_load_3d(&P.mapS, dst, &sfull_bar[sb], ch * kc, (int)q0, p.sc_plane);
              tma_load_3d(&P.mapS, dst + sp.slab_bytes, &sfull_bar[sb], ch * kc, (int)q0, p.sc_plane + 1);
            }
          }
          __syncwarp();
          if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
          const CUtensorMap* wm = main ? &P.mapWm : &P.mapWs;
          for (int tap = 0; tap < (main ? p.ntaps : 1); ++tap) {
            const int kofs = main ? (tap * p.chunks_main + c) * p.kc_main : n_main * p.kc_main + (c - p.chunks_main) * p.kc_sc;
            mbar_wait(&bempty_bar[bs], bphase ^ 1);
            if (elect_one()) {
              const uint32_t b_hi = smem_u32(b_base + (size_t)bs * 2 * sp.bplane_bytes);
              mbar_expect_tx(&bfull_bar[bs], (uint32_t)(2 * BN * kc * 2));
              tma_load_3d(wm, b_hi, &bfull_bar[bs], kofs, n0, 0);
              tma_load_3d(wm, b_hi + (uint32_t)(BN * kc * 2), &bfull_bar[bs], kofs, n0, 1);
            }
            __syncwarp();
            if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == SL_WARP_MMA) {
    // ===================== MMA issuer (see conv_tc_slab_kernel) =====================
    if (elect_one()) {
      const uint32_t idesc_n = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t idesc_2n = (1u << 4) | ((uint32_t)(BN >> 2) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      constexpr uint32_t ROWB = KC * 2, ROW16 = ROWB >> 4;
      constexpr uint32_t DHI = ((8 * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
      const uint32_t a_base0 = ((smem_u32(slab_base) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_slab16 = (uint32_t)(2 * sp.slab_bytes) >> 4, a_lo16 = (uint32_t)sp.slab_bytes >> 4;
      const uint32_t b_base0 = ((smem_u32(b_base) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_slot16 = (uint32_t)(2 * sp.bplane_bytes) >> 4;
      const uint32_t p_row16 = (uint32_t)g0.P * ROW16;
      const uint32_t tap0_16 = (uint32_t)(sp.lead - g0.P - 1) * ROW16;
      int sb = 0; uint32_t sphase = 0;
      int bs = 0; uint32_t bphase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const TcParams& p = cp.ph[item / total_tiles].p;
        {
          const int as = it & 1;
          const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t acc0 = tmem_base + (uint32_t)(as * 2 * BN);
          const uint32_t acc1 = acc0 + (uint32_t)BN;
          uint32_t acc_on = 0;
          for (int c = 0; c < p.chunks_main; ++c) {
            mbar_wait(&sfull_bar[sb], sphase);
            tc_fence_after();
            if (c == 0) { CH_STAMP(it, 3) }
            uint32_t a_row = a_base0 + (uint32_t)sb * a_slab16 + tap0_16;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&bfull_bar[bs], bphase);
                tc_fence_after();
                const uint32_t b0 = b_base0 + (uint32_t)bs * b_slot16;
                const uint32_t a0 = a_row + (uint32_t)kw * ROW16;
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint64_t dah = ((uint64_t)DHI << 32) | (a0 + 2u * kk), dal = ((uint64_t)DHI << 32) | (a0 + a_lo16 + 2u * kk);
                  const uint64_t dbh = ((uint64_t)DHI << 32) | (b0 + 2u * kk);
                  umma_f16(acc0, dah, dbh, idesc_2n, acc_on);
                  umma_f16(acc1, dal, dbh, idesc_n, 1u);
                  acc_on = 1u;
                }
                umma_commit(&bempty_bar[bs]);
                if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
              }
              a_row += p_row16;
            }
            umma_commit(&sempty_bar[sb]);
            if (p.chunks_sc == 0 && c == p.chunks_main - 1) umma_commit(&tfull_bar[as]);
            if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
          }
          for (int c = 0; c < p.chunks_sc; ++c) {
            const uint32_t rowb = (uint32_t)p.kc_sc * 2;
            const uint32_t dhi_s = ((8 * rowb) >> 4) | (1u << 14) | ((rowb == 128 ? 2u : 4u) << 29);
            mbar_wait(&sfull_bar[sb], sphase);
            tc_fence_after();
            const uint32_t a0 = a_base0 + (uint32_t)sb * a_slab16;
            mbar_wait(&bfull_bar[bs], bphase);
            tc_fence_after();
            const uint32_t b0 = b_base0 + (uint32_t)bs * b_slot16;
            for (int kk = 0; kk < p.kc_sc / 16; ++kk) {
              const uint64_t dah = ((uint64_t)dhi_s << 32) | (a0 + 2u * kk), dal = ((uint64_t)dhi_s << 32) | (a0 + a_lo16 + 2u * kk);
              const uint64_t dbh = ((uint64_t)dhi_s << 32) | (b0 + 2u * kk);
              umma_f16(acc0, dah, dbh, idesc_2n, 1u);
              umma_f16(acc1, dal, dbh, idesc_n, 1u);
            }
            umma_commit(&bempty_bar[bs]);
            if (++bs == sp.nring) { bs = 0; bphase ^= 1; }
            umma_commit(&sempty_bar[sb]);
            if (c == p.chunks_sc - 1) umma_commit(&tfull_bar[as]);
            if (++sb == sp.nslab) { sb = 0; sphase ^= 1; }
          }
          CH_STAMP(it, 4)
        }
      }
    }
    __syncwarp();
  } else if (warp == CH_WARP_PUB) {
    // ===================== publisher: tile stored by all its epilogue warps -> fence -> counter =====================
    if (cp.publisher && lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int ph = item / total_tiles, tile = item - ph * total_tiles;
        const TcParams& p = cp.ph[ph].p;
        mbar_wait(&pub_bar[it & 1], (uint32_t)(it >> 1) & 1u);
        mbar_arrive(&pub_bar[2 + (it & 1)]);
        __threadfence();
        atomicAdd(p.flag_done + tile / p.n_tiles, 4 * (BN >> 5));
        CH_STAMP(it, 7)
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3, group = warp >> 2;
    uint64_t* const pubp = cp.publisher ? pub_bar : nullptr;
    const uint32_t stage_u = smem_u32(epi_base + (size_t)warp * EPI_WARP_BYTES);
    const uint32_t vec_u = smem_u32(vec_base + (size_t)warp * 384);
    const int nchunks = BN >> 5;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int ph = item / total_tiles, tile = item - ph * total_tiles;
      const TcParams& p = cp.ph[ph].p;
      const OutMaps& om = cp.ph[ph].om;
      for (int c = 0; c < nchunks; ++c)
          if (((it * nchunks + c) & 1) == group) {
            if (quad == 0 && lane == 0 && c == 0) { CH_STAMP(it, 5) }
            switch (p.epi_mode) {
              case 0: epilogue_item<0, true>(p, om, tile, c, it, quad, lane, tmem_base, tfull_bar, tempty_bar, vec_u, stage_u, pubp); break;
              case 3: epilogue_item<3, true>(p, om, tile, c, it, quad, lane, tmem_base, tfull_bar, tempty_bar, vec_u, stage_u, pubp); break;
              case 4: epilogue_item<4, true>(p, om, tile, c, it, quad, lane, tmem_base, tfull_bar, tempty_bar, vec_u, stage_u, pubp); break;
              default: epilogue_item<-1, true>(p, om, tile, c, it, quad, lane, tmem_base, tfull_bar, tempty_bar, vec_u, stage_u, pubp); break;
            }
            if (quad == 0 && lane == 0 && c == 0) { CH_STAMP(it, 6) }
            if (!cp.publisher && quad == 0 && lane == 0 && c == nchunks - 1) { CH_STAMP(it, 7) }
          }
    }
    if (lane == 0) bulk_wait0();                     // (items of TMA-store layers already waited before publishing)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SL_WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
  if (cp.dbg && threadIdx.x == 0) cp.dbg[(size_t)blockIdx.x * CH_DBG_STRIDE + 2] = clock64();
  // self-cleaning counters: the last CTA to leave zeroes them for the next launch
  __shared__ int s_last;
  const int nflags = cp.n_phases * g0.m_tiles;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(cp.flags + nflags, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last)
    for (int i = threadIdx.x; i <= nflags; i += CH_THREADS) cp.flags[i] = 0;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// planes tensor [planes][rows][ch] fp16, box = (kc, box_rows, 1)
int tc_make_map(CUtensorMap* map, const void* base, long long rows, int ch, int planes, int kc, int box_rows);
static int make_map(CUtensorMap* map, const void* base, long long rows, int ch, int planes, int kc, int box_rows) {
  return tc_make_map(map, base, rows, ch, planes, kc, box_rows);
}
static int make_map_t(CUtensorMap* map, const void* base, long long rows, int ch, int planes, int kc, int box_rows, int esize);
int tc_make_map(CUtensorMap* map, const void* base, long long rows, int ch, int planes, int kc, int box_rows) {
  return make_map_t(map, base, rows, ch, planes, kc, box_rows, 2);
}
// esize 2: fp16 planes; esize 4: one fp32 plane (the residual stream)
static int make_map_t(CUtensorMap* map, const void* base, long long rows, int ch, int planes, int kc, int box_rows, int esize) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return SAR_ERR_UNSUPPORTED; }
  cuuint64_t dims[3] = {(cuuint64_t)ch, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)ch * esize, (cuuint64_t)rows * ch * esize};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = (kc * esize == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(map, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld ch=%d planes=%d kc=%d box_rows=%d", (int)r, rows, ch, planes, kc, box_rows); return SAR_ERR_BAD_ARG; }
  return SAR_OK;
}

static int pick_kc(int ch) { return (ch % 64 == 0) ? 64 : 32; }

// floor(x / d) == __umulhi(x, m) >> sh for every x < 2^31 (d >= 2); m = 0 flags d == 1
static void fast_div(unsigned d, unsigned* m, int* sh) {
  if (d <= 1) { *m = 0; *sh = 0; return; }
  int s = 0;
  while ((1ull << s) < d) ++s;                       // s = ceil(log2 d) >= 1
  *m = (unsigned)(((1ull << (31 + s)) / d) + 1);
  *sh = s - 1;
}

}  // namespace sar

namespace sar {
// argument checks + the layer description shared by sar_conv_tc_fwd and sar_conv_tc_chain_fwd
static int fill_params(const sar_tc_conv* d, TcParams& p) {
  SAR_REQUIRE(d, SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: null descriptor");
  SAR_REQUIRE(d->a && d->w && d->bias, SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: null a/w/bias");
  SAR_REQUIRE(d->ntaps >= 1 && d->ntaps <= 9, SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: ntaps must be 1..9");
  SAR_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->cout > 0 && d->a_ch > 0 && d->a_rows > 0, SAR_ERR_BAD_ARG,
              "sar_conv_tc_fwd: non-positive dimension");
  SAR_REQUIRE(d->a_ch % 32 == 0 && d->cout % 32 == 0 && (!d->s || d->s_ch % 32 == 0), SAR_ERR_UNSUPPORTED,
              "sar_conv_tc_fwd: channel counts must be multiples of 32 (a_ch=%d s_ch=%d cout=%d)", d->a_ch, d->s_ch, d->cout);
  SAR_REQUIRE(d->out_raw || d->out_raw_f32 || d->out_act || d->out_dense, SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: no output requested");
  SAR_REQUIRE(d->act_kind >= 0 && d->act_kind <= 2, SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: act_kind must be 0, 1 or 2");
  SAR_REQUIRE(!(d->out_act || d->out_dense) || d->act_kind != 0 || (d->act_scale && d->act_shift), SAR_ERR_BAD_ARG,
              "sar_conv_tc_fwd: act_kind 0 (BN->ReLU) outputs need act_scale/act_shift");
  SAR_REQUIRE(aligned16(d->a) && aligned16(d->w) && (!d->s || aligned16(d->s)) && (!d->res || aligned16(d->res)) &&
                  (!d->out_raw || aligned16(d->out_raw)) && (!d->out_act || aligned16(d->out_act)) &&
                  (!d->out_dense || aligned16(d->out_dense)),
              SAR_ERR_ALIGN, "sar_conv_tc_fwd: pointers must be 16-byte aligned");

  p.ntaps = d->ntaps;
  p.kc_main = pick_kc(d->a_ch);
  p.chunks_main = d->a_ch / p.kc_main;
  // shortcut chunks never wider than the main chunks: the slab / weight-slot sizes follow the widest chunk, and a
  // 64-wide shortcut chunk on a 32-channel layer (stage 1, block 1: 64-channel stem -> 32) doubled both, which pushed
  // the layer off the resident-weights form (207 -> ~140 us at B=512)
  p.kc_sc = d->s ? pick_kc(d->s_ch) : 32;
  if (p.kc_sc > p.kc_main) p.kc_sc = p.kc_main;
  p.chunks_sc = d->s ? d->s_ch / p.kc_sc : 0;
  p.sc_plane = d->s_plane;
  for (int t = 0; t < d->ntaps; ++t) {
    p.tap_row_off[t] = d->tap_row_off[t];
    p.tap_plane[t] = d->tap_plane[t];
    SAR_REQUIRE(d->tap_plane[t] >= 0 && d->tap_plane[t] + 1 < d->a_planes, SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: tap plane out of range");
  }
  p.B = d->B; p.H = d->H; p.W = d->W;
  p.P = d->W + 1; p.Rimg = (d->H + 1) * p.P;
  if (d->nopad) { p.P = d->W; p.Rimg = d->H * d->W; }      // plain row-major operand (GEMM use): no pad row / column
  p.R = (long long)d->B * p.Rimg;
  SAR_REQUIRE(p.R < (1ll << 31) - 4096, SAR_ERR_UNSUPPORTED, "sar_conv_tc_fwd: too many rows");
  p.split = d->out_split ? 1 : 0;
  p.P2 = (d->W + 1) / 2 + 1;
  p.Rimg2 = ((d->H + 1) / 2 + 1) * p.P2;
  p.R2 = (long long)d->B * p.Rimg2;
  p.Cout = d->cout;
  p.BN = (d->cout % 128 == 0) ? 128 : (d->cout % 64 == 0 ? 64 : 32);
  p.m_tiles = (int)((p.R + TC_BM - 1) / TC_BM);
  p.n_tiles = d->cout / p.BN;
  p.bias = d->bias; p.act_scale = d->act_scale; p.act_shift = d->act_shift;
  p.res = reinterpret_cast<const __half*>(d->res);
  p.res32 = d->res_f32;
  p.out_raw32 = d->out_raw_f32;
  SAR_REQUIRE(!(d->res && d->res_f32) && !(d->out_raw && d->out_raw_f32), SAR_ERR_BAD_ARG,
              "sar_conv_tc_fwd: the residual stream is either hi/lo planes or one fp32 plane, not both");
  SAR_REQUIRE((!d->res_f32 || aligned16(d->res_f32)) && (!d->out_raw_f32 || aligned16(d->out_raw_f32)), SAR_ERR_ALIGN,
              "sar_conv_tc_fwd: fp32 residual pointers must be 16-byte aligned");
  p.out_raw = reinterpret_cast<__half*>(d->out_raw);
  p.out_act = reinterpret_cast<__half*>(d->out_act);
  p.out_dense = d->out_dense;
  p.dbg = reinterpret_cast<long long*>(d->dbg);
  p.act_kind = d->act_kind;
  p.mma_mask = getenv("SAR_TC_MMAMASK") ? atoi(getenv("SAR_TC_MMAMASK")) : 7;
  SAR_REQUIRE(!(p.split && p.out_dense), SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: dense output cannot be phase-split");
  SAR_REQUIRE(!(p.out_act && p.out_dense), SAR_ERR_BAD_ARG,
              "sar_conv_tc_fwd: out_act and out_dense are the same activated values in two layouts -- request one of them");

  p.nacc_log2 = 1; p.epi_warps = 8;
  p.ksplit = d->ksplit > 1 ? d->ksplit : 1;
  if (p.ksplit > 1) {
    SAR_REQUIRE(d->ntaps == 1 && !d->s && !d->res && !d->out_raw && !d->out_act && d->out_dense && d->act_kind == 1,
                SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: split-K needs a 1-tap GEMM with only the (identity) dense output");
    SAR_REQUIRE(p.chunks_main % p.ksplit == 0, SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: ksplit must divide a_ch / %d", p.kc_main);
  }
  p.ksteps_split = p.ksplit > 1 ? p.chunks_main / p.ksplit : 0;
  // plane outputs of an unsplit map leave through TMA stores (SAR_TC_TMA_OUT=0: the LDS -> STG write-out, an A/B aid)
  static const bool tma_ok = !(getenv("SAR_TC_TMA_OUT") && getenv("SAR_TC_TMA_OUT")[0] == '0');
  p.tma_out = (tma_ok && !p.split && !d->out_dense && (d->out_raw || d->out_raw_f32 || d->out_act)) ? 1 : 0;
  SAR_REQUIRE(!d->out_raw_f32 || (!p.split && !d->out_dense), SAR_ERR_UNSUPPORTED,
              "sar_conv_tc_fwd: out_raw_f32 needs unsplit plane outputs");
  SAR_REQUIRE(!(p.ksplit > 1) || (!d->res_f32 && !d->out_raw_f32), SAR_ERR_BAD_ARG, "sar_conv_tc_fwd: split-K has no residual stream");
  return SAR_OK;
}

// tensor maps of the plane outputs for the TMA-store epilogue (dummies when the layer does not use it)
static int fill_out_maps(const sar_tc_conv* d, const TcParams& p, const CUtensorMap& dummy, OutMaps& om) {
  om.raw = dummy; om.act = dummy;
  if (!p.tma_out) return SAR_OK;
  int rc;
  if (d->out_raw && (rc = tc_make_map(&om.raw, d->out_raw, p.R, d->cout, 2, 32, 32))) return rc;
  if (d->out_raw_f32 && (rc = make_map_t(&om.raw, d->out_raw_f32, p.R, d->cout, 1, 32, 32, 4))) return rc;
  if (d->out_act && (rc = tc_make_map(&om.act, d->out_act, p.R, d->cout, 2, 32, 32))) return rc;
  return SAR_OK;
}
}  // namespace sar

extern "C" int sar_conv_tc_fwd(const sar_tc_conv* d, void* stream) {
  using namespace sar;
  TcParams p{};
  { const int frc = fill_params(d, p); if (frc) return frc; }
  const int ktot = d->ntaps * d->a_ch + (d->s ? d->s_ch : 0);
  // slab path: plain 3x3 stride-1 taps on a non-split tensor whose halo'd slab fits shared memory
  bool slab = d->ntaps == 9 && d->a_planes == 2;
  for (int t = 0; slab && t < 9; ++t)
    slab = d->tap_plane[t] == 0 && d->tap_row_off[t] == (t / 3 - 1) * p.P + (t % 3 - 1);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // N tile: minimise (rounds over the SMs) x (per-tile cost ~ BN + fixed) -- late, small maps have fewer
  // M tiles than SMs and want a narrower N tile so more SMs share the layer
  {
    int best = p.BN; long long best_cost = -1;
    for (int bn = p.BN; bn >= 32; bn >>= 1) {
      if (d->cout % bn) continue;
      const long long tiles = (long long)p.m_tiles * (d->cout / bn) * p.ksplit;
      const long long cost = ((tiles + sms - 1) / sms) * (bn + 32);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
    }
    p.BN = best; p.n_tiles = d->cout / p.BN;
  }
  SlabParams sp{};
  sp.lead = p.P + 1;
  sp.slab_rows = (TC_BM + 2 * p.P + 2 + 7) & ~7;
  const int kc_max = (d->s && p.kc_sc > p.kc_main) ? p.kc_sc : p.kc_main;
  sp.slab_bytes = (sp.slab_rows * kc_max * 2 + 1023) & ~1023;
  if (sp.slab_bytes < TC_BM * kc_max * 2) sp.slab_bytes = TC_BM * kc_max * 2;
  sp.bplane_bytes = p.BN * kc_max * 2;       // hi and lo tiles adjacent: [B_hi ; B_lo] is one 2*BN-row operand
  if (sp.slab_rows > 192) slab = false;
  // TMA epilogue: plane outputs of an unsplit map (every conv1 and all but the three stage-ending conv2's)
  // epilogue staging (64 KB): aliased onto the operand region when every CTA owns a single tile
  fast_div((unsigned)p.Rimg, &p.div_rimg_m, &p.div_rimg_sh);
  fast_div((unsigned)p.P, &p.div_p_m, &p.div_p_sh);
  p.mn_tiles = p.m_tiles * p.n_tiles;
  p.dense_zstride = (long long)d->B * d->H * d->W * d->cout;
  p.epi_alias = ((long long)p.mn_tiles * p.ksplit <= sms) ? 1 : 0;
  // epilogue mode: the residual-block combinations get compile-time flags (bit 0 identity shortcut, bit 1 raw out)
  int mode = -1;
  if (d->out_act && !d->out_dense && d->act_kind == 0 && !p.split && !(d->res_f32 && d->out_raw) && !(d->res && d->out_raw_f32))
    mode = ((d->res || d->res_f32) ? 1 : 0) | ((d->out_raw || d->out_raw_f32) ? 2 : 0);
  if (mode == 1) mode = -1;                                            // (no compile-time instance of shortcut-without-raw)
  static const bool mode4_ok = !(getenv("SAR_TC_MODE4") && getenv("SAR_TC_MODE4")[0] == '0');
  if (mode4_ok && p.split && d->out_act && d->out_raw && d->res_f32 && !d->res && !d->out_raw_f32 && !d->out_dense && d->act_kind == 0)
    mode = 4;                                                          // the stage-ending conv2 (see epilogue_item)
  static const bool epi16_ok = !(getenv("SAR_TC_EPI16") && getenv("SAR_TC_EPI16")[0] == '0');
  static const bool epi16_thin_ok = !(getenv("SAR_TC_EPI16_THIN") && getenv("SAR_TC_EPI16_THIN")[0] == '0');
  size_t fixed = 0, budget = 0;
  auto plan_slab = [&](int epi_warps) -> bool {          // shared-memory plan of the slab kernel with this many epilogue warps
    fixed = 1024 + 1024 + 640 + 3 * (size_t)d->cout * sizeof(float) + (p.epi_alias ? 0 : (size_t)epi_warps * EPI_WARP_BYTES);   // align slack (x2) + barriers + epilogue vectors (+ staging)
    if (fixed + 4096 > 227 * 1024) return false;
    budget = 227 * 1024 - fixed;
    const int n_ksteps = p.ntaps * p.chunks_main + p.chunks_sc;
    const size_t wres = (size_t)n_ksteps * 2 * sp.bplane_bytes;
    const size_t slab1 = 2 * (size_t)sp.slab_bytes;
    if (p.n_tiles == 1 && wres + 2 * slab1 <= budget) {
      sp.resident = 1;
      sp.nring = 0;
      sp.nslab = (int)((budget - wres) / slab1);
    } else {
      if (budget < 2 * slab1 + 2 * 2 * (size_t)sp.bplane_bytes) return false;
      sp.resident = 0;
      sp.nslab = 2;
      sp.nring = (int)((budget - 2 * slab1) / (2 * (size_t)sp.bplane_bytes));
      if (sp.nring > SL_MAX_RING) sp.nring = SL_MAX_RING;
      if (sp.nring < 2) return false;
    }
    if (sp.nslab > SL_MAX_SLABS) sp.nslab = SL_MAX_SLABS;
    return true;
  };
  if (slab) {
    // Thin tiles (BN <= 64) of a launch with several tiles per CTA: an epilogue item (tcgen05.ld, convert, split, stage,
    // store) is ~3.5k cycles of mostly latency per warp while the tile's MMAs take 1-2k, so eight warps (two tiles in
    // flight) leave the tensor pipe waiting for a free accumulator.  Sixteen warps drain four 32-wide tiles (out of
    // four TMEM accumulator stages) or two 64-wide ones at a time, when the 128 KB of staging still leaves room for
    // the slabs and the weights (SAR_TC_EPI16_THIN=0 disables).
    bool thin16 = epi16_ok && epi16_thin_ok && mode >= 0 && !p.epi_alias && p.BN <= 64 && p.mn_tiles >= 2 * sms;
    // ... and the slab ring still runs a tile ahead: a tile of a projection layer takes chunks_main + chunks_sc slabs,
    // and with fewer buffers than that + 1 the next tile's first slab waits for this tile's MMAs to retire (TMA
    // latency exposed on every tile: 5.5k instead of 2.9k cycles per tile on stage 1's projection layer)
    const int slabs_per_tile = p.chunks_main + p.chunks_sc;
    if (thin16 && plan_slab(16) && (sp.resident || sp.nring >= 4) && sp.nslab >= (slabs_per_tile > 1 ? slabs_per_tile + 1 : 2)) {
      p.epi_warps = 16;
      p.nacc_log2 = (p.BN == 32) ? 2 : 1;
    } else {
      thin16 = false;
      slab = plan_slab(8);
    }
  }
  (void)budget;
  size_t smem = 0;                                         // generic kernel: dynamic shared memory
  if (!slab) {
    // ring stage sized for this layer's chunk width and N tile (a 32-channel chunk needs a quarter of the 64 KB the
    // first version reserved), as many stages as shared memory holds: the k-step loop of a strided conv / Dense is
    // TMA-latency bound (2 stages: ~1.2k cycles per k-step), depth is what hides it.  Wide tiles (64-channel chunks,
    // BN = 128: 64 KB per stage) would still get only 2 stages next to the epilogue staging: they run 32-channel
    // chunks instead (twice the k-steps, 4-5 stages in flight).
    const size_t other = 1024 + (p.epi_alias ? 0 : EPI_BYTES) + 256 + 3 * (size_t)d->cout * sizeof(float);
    auto stages_for = [&](int kc) { return (int)((227 * 1024 - other) / (size_t)(2 * TC_BM * kc * 2 + 2 * p.BN * kc * 2)); };
    if (p.kc_main == 64 && stages_for(64) < 4) {
      p.kc_main = 32; p.chunks_main *= 2;
      if (d->s) { p.kc_sc = 32; p.chunks_sc = d->s_ch / 32; }
      if (p.ksplit > 1) p.ksteps_split = p.chunks_main / p.ksplit;
    }
    const int kcm = p.kc_main;                             // (kc_sc <= kc_main)
    p.a_plane_bytes = TC_BM * kcm * 2;
    p.stage_bytes = 2 * p.a_plane_bytes + 2 * p.BN * kcm * 2;
    p.stages = stages_for(kcm);
    if (p.stages > TC_STAGES) p.stages = TC_STAGES;
    const int n_ksteps_tile = (p.ksplit > 1 ? p.ksteps_split : p.ntaps * p.chunks_main) + p.chunks_sc;
    if (p.stages > n_ksteps_tile * 2) p.stages = n_ksteps_tile * 2;      // never more than two tiles' worth
    if (p.epi_alias)                                       // aliased staging (64 KB) lives on the ring
      while ((size_t)p.stages * p.stage_bytes < (size_t)EPI_BYTES) ++p.stages;
    SAR_REQUIRE(p.stages >= 2 && p.stages <= TC_STAGES, SAR_ERR_UNSUPPORTED, "sar_conv_tc_fwd: shared-memory plan (stages=%d)", p.stages);
    smem = other + (size_t)p.stages * p.stage_bytes;
  }

  CUtensorMap mapA, mapS, mapWm, mapWs;
  int rc;
  if ((rc = make_map(&mapA, d->a, d->a_rows, d->a_ch, d->a_planes, p.kc_main, slab ? sp.slab_rows : TC_BM))) return rc;
  if ((rc = make_map(&mapWm, d->w, d->cout, ktot, 2, p.kc_main, p.BN))) return rc;
  if (d->s) {
    SAR_REQUIRE(d->s_rows > 0 && d->s_planes >= 2 && d->s_plane >= 0 && d->s_plane + 1 < d->s_planes, SAR_ERR_BAD_ARG,
                "sar_conv_tc_fwd: bad shortcut operand");
    if ((rc = make_map(&mapS, d->s, d->s_rows, d->s_ch, d->s_planes, p.kc_sc, TC_BM))) return rc;
    if ((rc = make_map(&mapWs, d->w, d->cout, ktot, 2, p.kc_sc, p.BN))) return rc;
  } else {
    mapS = mapA; mapWs = mapWm;
  }
  OutMaps om;
  if ((rc = fill_out_maps(d, p, mapA, om))) return rc;
  const int tiles = p.mn_tiles * p.ksplit;
  const int grid = tiles < sms ? tiles : sms;
  // 16 epilogue warps when every CTA owns ONE 128-wide tile: its 16 (quadrant, 32-column) items then drain in one
  // round instead of two -- the epilogue of such a launch is a tail nothing overlaps (SAR_TC_EPI16=0 disables)
  auto threads_for = [&](size_t operand_bytes, int mode_) {
    if (p.epi_warps == 16) return TC_THREADS_MAX;
    return (epi16_ok && mode_ >= 0 && p.epi_alias && p.BN >= 128 && operand_bytes >= 2 * (size_t)EPI_BYTES) ? TC_THREADS_MAX : TC_THREADS;
  };
  // CTA pairs (conv_tc_pair_kernel): the multi-round layers whose weights stream through the ring (stages 2-4 at large
  // batches) -- every SM then fetches half of each weight tile.  SAR_TC_PAIR=0 disables, =2 forces it for every eligible layer.
  static const int pair_env = getenv("SAR_TC_PAIR") ? atoi(getenv("SAR_TC_PAIR")) : 1;
  // (64-channel layers -- stage 2 -- are epilogue-bound at this tile size: coupling two CTAs' accumulator hand-back
  // made their shortcut layers 13 % slower, so pairs start at 128 input channels)
  if (slab && pair_env && !sp.resident && p.kc_main == 64 && (p.chunks_sc == 0 || p.kc_sc == 64) && p.epi_warps == 8 && p.m_tiles >= 2 &&
      (sms & 1) == 0 && (pair_env == 2 || ((long long)p.mn_tiles >= 2LL * sms && p.chunks_main >= 2))) {
    const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int ncl = pairs < sms / 2 ? pairs : sms / 2;
    TcParams q = p;
    SlabParams sq = sp;
    q.pair = 1; q.nacc_log2 = 1;
    q.epi_alias = (pairs <= ncl) ? 1 : 0;
    const size_t fixed_p = 1024 + 1024 + 640 + 3 * (size_t)d->cout * sizeof(float) + (q.epi_alias ? 0 : (size_t)EPI_BYTES);
    const size_t slab1 = 2 * (size_t)sp.slab_bytes, bstage = (size_t)p.BN * 64 * 2;      // [B_hi half | B_lo half] of one k-step
    sq.resident = 0;
    for (sq.nslab = (p.chunks_main >= 2) ? 3 : 2; sq.nslab >= 2; --sq.nslab) {       // a third slab only if >= 6 weight stages remain
      const long long left = (long long)227 * 1024 - (long long)fixed_p - (long long)(sq.nslab * slab1);
      sq.nring = left > 0 ? (int)(left / (long long)bstage) : 0;
      if (sq.nring > SL_MAX_RING) sq.nring = SL_MAX_RING;
      if (sq.nring >= 6 || sq.nslab == 2) break;
    }
    if (sq.nring >= 4) {
      CUtensorMap mapWh;
      if ((rc = make_map(&mapWh, d->w, d->cout, ktot, 2, 64, p.BN / 2))) return rc;
      size_t smem_p = fixed_p + sq.nslab * slab1 + (size_t)sq.nring * bstage;
      if (q.epi_alias && smem_p - fixed_p < (size_t)EPI_BYTES) smem_p = fixed_p + EPI_BYTES;
      // 16 epilogue warps when every CTA owns ONE 128-wide tile (its 16 items then drain in one round), as in the slab kernel
      const bool wide = epi16_ok && mode >= 0 && q.epi_alias && p.BN >= 128 && smem_p - fixed_p >= 2 * (size_t)EPI_BYTES;
      if (wide) q.epi_warps = 16;
      auto launch_p = [&](auto kern) -> int {
        { const int arc = allow_max_smem(kern, "sar_conv_tc_fwd(pair)"); if (arc) return arc; }
        launch_k(kern, dim3(2 * ncl), dim3(wide ? TC_THREADS_MAX : TC_THREADS), smem_p, (cudaStream_t)stream, mapA, mapS, mapWh, om, q, sq);
        return 0;
      };
      int lrc;
      switch (mode) {
        case 0: lrc = launch_p(conv_tc_pair_kernel<0>); break;
        case 2: lrc = launch_p(conv_tc_pair_kernel<2>); break;
        case 3: lrc = launch_p(conv_tc_pair_kernel<3>); break;
        case 4: lrc = launch_p(conv_tc_pair_kernel<4>); break;
        default: lrc = launch_p(conv_tc_pair_kernel<-1>); break;
      }
      if (lrc) return lrc;
      return check_launch("sar_conv_tc_fwd(pair)");
    }
  }
  if (slab) {
    const int nb = sp.resident ? (p.ntaps * p.chunks_main + p.chunks_sc) : sp.nring;
    size_t smem = fixed + (size_t)sp.nslab * 2 * sp.slab_bytes + (size_t)nb * 2 * sp.bplane_bytes;
    if (p.epi_alias && smem - fixed < (size_t)EPI_BYTES) smem = fixed + EPI_BYTES;    // aliased staging needs 64 KB of operand region
    auto launch = [&](auto kern) -> int {
      { const int arc = allow_max_smem(kern, "sar_conv_tc_fwd"); if (arc) return arc; }
      const int m_eff = (mode == 0 || mode == 2 || mode == 3 || mode == 4) ? mode : -1;
      launch_k(kern, dim3(grid), dim3(threads_for(smem - fixed, m_eff)), smem, (cudaStream_t)stream, mapA, mapS, mapWm, mapWs, om, p, sp);
      return 0;
    };
    auto pick = [&](auto kc_tag, auto res_tag) -> int {
      constexpr int KCv = decltype(kc_tag)::value;
      constexpr bool RSv = decltype(res_tag)::value;
      switch (mode) {
        case 0: return launch(conv_tc_slab_kernel<KCv, RSv, 0>);
        case 2: return launch(conv_tc_slab_kernel<KCv, RSv, 2>);
        case 3: return launch(conv_tc_slab_kernel<KCv, RSv, 3>);
        case 4: return launch(conv_tc_slab_kernel<KCv, RSv, 4>);
        default: return launch(conv_tc_slab_kernel<KCv, RSv, -1>);
      }
    };
    int lrc;
    if (p.kc_main == 64) lrc = sp.resident ? pick(std::integral_constant<int, 64>{}, std::true_type{}) : pick(std::integral_constant<int, 64>{}, std::false_type{});
    else lrc = sp.resident ? pick(std::integral_constant<int, 32>{}, std::true_type{}) : pick(std::integral_constant<int, 32>{}, std::false_type{});
    if (lrc) return lrc;
  } else {
    { const int arc = allow_max_smem(conv_tc_kernel, "sar_conv_tc_fwd"); if (arc) return arc; }
    launch_k(conv_tc_kernel, dim3(grid), dim3(TC_THREADS), smem, (cudaStream_t)stream, mapA, mapS, mapWm, mapWs, om, p);
  }
  return check_launch("sar_conv_tc_fwd");
}

extern "C" size_t sar_conv_tc_chain_workspace_bytes(const sar_tc_conv* first, int n) {
  if (!first || n <= 0) return 0;
  const long long R = (long long)first->B * (first->H + 1) * (first->W + 1);
  const long long m_tiles = (R + sar::TC_BM - 1) / sar::TC_BM;
  return (size_t)(n * m_tiles + 1) * sizeof(int);
}

extern "C" int sar_conv_tc_chain_fwd(const sar_tc_conv* descs, int n, void* workspace, size_t workspace_bytes, void* stream) {
  return sar_conv_tc_chain_grid_fwd(descs, n, workspace, workspace_bytes, 0, stream);
}

extern "C" int sar_conv_tc_chain_grid_fwd(const sar_tc_conv* descs, int n, void* workspace, size_t workspace_bytes, int max_ctas,
                                          void* stream) {
  using namespace sar;
  SAR_REQUIRE(descs && n >= 1 && n <= CH_MAX, SAR_ERR_BAD_ARG, "sar_conv_tc_chain_fwd: need 1..%d layers", CH_MAX);
  SAR_REQUIRE(workspace && workspace_bytes >= sar_conv_tc_chain_workspace_bytes(descs, n), SAR_ERR_WORKSPACE,
              "sar_conv_tc_chain_fwd: workspace too small");
  static thread_local ChainParams cp;          // ~10 KB: kept off the stack
  static_assert(sizeof(ChainParams) < 32000, "ChainParams must fit the kernel parameter space");
  cp.n_phases = n;
  cp.flags = reinterpret_cast<int*>(workspace);
  // SAR_CHAIN_PUB=0: the round-2a protocol (TMA stores, every epilogue warp waits for its stores and fences) -- A/B aid
  static const bool use_pub = !(getenv("SAR_CHAIN_PUB") && getenv("SAR_CHAIN_PUB")[0] == '0');
  cp.publisher = use_pub ? 1 : 0;
  cp.dbg = reinterpret_cast<long long*>(descs->dbg);      // chain profiling: >= 148 * (8 + 32 * 8) int64 (scripts/chain_dbg.py)
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int kc_max = 32;
  for (int i = 0; i < n; ++i) {
    const sar_tc_conv* d = descs + i;
    TcParams& p = cp.ph[i].p;
    p = TcParams{};
    { const int frc = fill_params(d, p); if (frc) return frc; }
    SAR_REQUIRE(d->ntaps == 9 && d->a_planes == 2 && !d->nopad && p.ksplit == 1 && !d->out_dense, SAR_ERR_UNSUPPORTED,
                "sar_conv_tc_chain_fwd: layer %d is not a stride-1 3x3 convolution with plane outputs", i);
    for (int t = 0; t < 9; ++t)
      SAR_REQUIRE(d->tap_plane[t] == 0 && d->tap_row_off[t] == (t / 3 - 1) * p.P + (t % 3 - 1), SAR_ERR_UNSUPPORTED,
                  "sar_conv_tc_chain_fwd: layer %d: not the stride-1 tap pattern", i);
    SAR_REQUIRE(d->B == descs->B && d->H == descs->H && d->W == descs->W && d->cout == descs->cout && d->a_ch == descs->a_ch,
                SAR_ERR_BAD_ARG, "sar_conv_tc_chain_fwd: layer %d: all layers of a chain share geometry and channel counts", i);
    SAR_REQUIRE(!d->s || i == 0, SAR_ERR_UNSUPPORTED, "sar_conv_tc_chain_fwd: only the first layer may carry a projection shortcut");
    if (p.kc_main > kc_max) kc_max = p.kc_main;
    if (d->s && p.kc_sc > kc_max) kc_max = p.kc_sc;
  }
  TcParams& g0 = cp.ph[0].p;
  const int cout = descs->cout;
  // N tile: 64 columns when possible -- the weight ring shares shared memory with the epilogue staging, and the
  // M x N tiles of consecutive phases pipeline, so narrow tiles fill the SMs without a round-quantisation loss
  const int BN = (cout % 64 == 0) ? 64 : 32;
  SlabParams sp{};
  sp.lead = g0.P + 1;
  sp.slab_rows = (TC_BM + 2 * g0.P + 2 + 7) & ~7;
  SAR_REQUIRE(sp.slab_rows <= 192, SAR_ERR_UNSUPPORTED, "sar_conv_tc_chain_fwd: map too wide for the slab path (W=%d)", descs->W);
  sp.slab_bytes = (sp.slab_rows * kc_max * 2 + 1023) & ~1023;
  if (sp.slab_bytes < TC_BM * kc_max * 2) sp.slab_bytes = TC_BM * kc_max * 2;
  sp.bplane_bytes = BN * kc_max * 2;
  sp.resident = 0;
  sp.nslab = 2;
  const size_t fixed = 1024 + 1024 + 640 + EPI_BYTES + 8 * 384;
  const size_t budget = 227 * 1024 - fixed;
  const size_t slab1 = 2 * (size_t)sp.slab_bytes;
  SAR_REQUIRE(budget > 2 * slab1 + 4 * (size_t)sp.bplane_bytes, SAR_ERR_UNSUPPORTED, "sar_conv_tc_chain_fwd: shared memory budget");
  size_t left = budget - 2 * slab1;
  if (g0.chunks_main >= 2 && left >= slab1 + 6 * 2 * (size_t)sp.bplane_bytes) { sp.nslab = 3; left -= slab1; }
  sp.nring = (int)(left / (2 * (size_t)sp.bplane_bytes));
  if (sp.nring > SL_MAX_RING) sp.nring = SL_MAX_RING;
  cp.sp = sp;
  const int m_tiles = g0.m_tiles;
  for (int i = 0; i < n; ++i) {
    const sar_tc_conv* d = descs + i;
    ChainPhase& P = cp.ph[i];
    TcParams& p = P.p;
    p.BN = BN; p.n_tiles = cout / BN; p.mn_tiles = p.m_tiles * p.n_tiles;
    fast_div((unsigned)p.Rimg, &p.div_rimg_m, &p.div_rimg_sh);
    fast_div((unsigned)p.P, &p.div_p_m, &p.div_p_sh);
    p.epi_alias = 0;
    p.dbg = nullptr;
    {   // SAR_CHAIN_DBG_PHASE=<l>: epilogue_item stamps of CTA 0's first two items of layer l, behind the per-CTA table
      static const int dbg_ph = getenv("SAR_CHAIN_DBG_PHASE") ? atoi(getenv("SAR_CHAIN_DBG_PHASE")) : -1;
      if (cp.dbg && i == dbg_ph) {
        const int tiles_l = m_tiles * (cout / BN), g = tiles_l < sms ? tiles_l : sms;
        p.dbg = cp.dbg + (size_t)148 * CH_DBG_STRIDE;
        p.dbg_it0 = (i * tiles_l + g - 1) / g;
      }
    }
    p.flag_done = cp.flags + (size_t)i * m_tiles;
    p.flag_dep = i > 0 ? cp.flags + (size_t)(i - 1) * m_tiles : nullptr;
    p.flag_need = 4 * (cout / 32);
    p.epi_mode = -1;
    if (d->out_act && d->act_kind == 0 && !p.split) {
      const int m = ((d->res || d->res_f32) ? 1 : 0) | ((d->out_raw || d->out_raw_f32) ? 2 : 0);
      if ((m == 0 || m == 3) && !(d->res_f32 && d->out_raw) && !(d->res && d->out_raw_f32)) p.epi_mode = m;
    }
    if (p.split && d->out_act && d->out_raw && d->res_f32 && !d->res && !d->out_raw_f32 && !d->out_dense && d->act_kind == 0 &&
        !(getenv("SAR_TC_MODE4") && getenv("SAR_TC_MODE4")[0] == '0'))
      p.epi_mode = 4;
    const int ktot = 9 * d->a_ch + (d->s ? d->s_ch : 0);
    int rc;
    if ((rc = make_map(&P.mapA, d->a, d->a_rows, d->a_ch, d->a_planes, p.kc_main, sp.slab_rows))) return rc;
    if ((rc = make_map(&P.mapWm, d->w, cout, ktot, 2, p.kc_main, BN))) return rc;
    if (d->s) {
      SAR_REQUIRE(d->s_rows > 0 && d->s_planes >= 2 && d->s_plane >= 0 && d->s_plane + 1 < d->s_planes, SAR_ERR_BAD_ARG,
                  "sar_conv_tc_chain_fwd: bad shortcut operand");
      if ((rc = make_map(&P.mapS, d->s, d->s_rows, d->s_ch, d->s_planes, p.kc_sc, TC_BM))) return rc;
      if ((rc = make_map(&P.mapWs, d->w, cout, ktot, 2, p.kc_sc, BN))) return rc;
    } else {
      P.mapS = P.mapA; P.mapWs = P.mapWm;
    }
    // publisher protocol: generic stores (a bulk store's completion is only visible to the thread that issued it)
    if (cp.publisher) p.tma_out = 0;
    SAR_REQUIRE(!d->out_raw_f32 || !p.split, SAR_ERR_UNSUPPORTED, "sar_conv_tc_chain_fwd: out_raw_f32 needs unsplit plane outputs");
    if ((rc = fill_out_maps(d, p, P.mapA, P.om))) return rc;
  }
  const int tiles = g0.m_tiles * (cout / BN);
  int grid = tiles < sms ? tiles : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;          // leave SMs to a chain of another stream (cooperative launches)
  const size_t smem = fixed + (size_t)sp.nslab * slab1 + (size_t)sp.nring * 2 * sp.bplane_bytes;
  auto launch = [&](auto kern) -> int {
    { const int arc = allow_max_smem(kern, "sar_conv_tc_chain_fwd"); if (arc) return arc; }
    // every CTA spin-waits on counters other CTAs of this launch publish: the whole grid must be co-resident.
    // cooperative launch makes the driver guarantee it (MPS / green contexts / a concurrent kernel could otherwise
    // leave waiting CTAs resident and their producers unscheduled).  SAR_CHAIN_COOP=0: plain PDL launch (A/B aid).
    static const bool coop = !(getenv("SAR_CHAIN_COOP") && getenv("SAR_CHAIN_COOP")[0] == '0');
    int nblk = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, kern, CH_THREADS, smem);
    if (nblk * sms < grid) { set_error("sar_conv_tc_chain_fwd: grid of %d CTAs cannot be co-resident (%d per SM x %d SMs)", grid, nblk, sms); return SAR_ERR_UNSUPPORTED; }
    cudaError_t le = coop ? launch_k_coop(kern, dim3(grid), dim3(CH_THREADS), smem, (cudaStream_t)stream, cp)
                          : launch_k(kern, dim3(grid), dim3(CH_THREADS), smem, (cudaStream_t)stream, cp);
    if (le != cudaSuccess) { set_error("sar_conv_tc_chain_fwd: launch: %s", cudaGetErrorString(le)); return (int)le; }
    return 0;
  };
  const int lrc = g0.kc_main == 64 ? launch(conv_tc_chain_kernel<64>) : launch(conv_tc_chain_kernel<32>);
  if (lrc) return lrc;
  return check_launch("sar_conv_tc_chain_fwd");
}
