// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), split-fp16 operands.
//
// Replaces (reference, per image, fp32): im2col_gpu (src/caffe/util/im2col.cu:8-62) +
// caffe_gpu_gemm/cublasSgemm (math_functions.cu:13-27) in
// BaseConvolutionLayer::forward_gpu_gemm (base_conv_layer.cpp:325-341), plus the
// BatchNorm/Scale/ReLU/Eltwise layers that follow it (batch_norm_layer.cu:10-90,
// scale_layer.cu:30-56, relu_layer.cu:17-32, eltwise_layer.cu:47-53) as a fused epilogue.
//
// GEMM view: D[128 output pixels][BN out-channels] += A[pixels][64 ch of tap t] * W[BN][64],
// looping over taps and 64-channel chunks.  Activations live in HBM as NHWC "split fp16":
// plane 0 = hi = fp16(x), plane 1 = lo = fp16(x - hi); weights likewise.  Each K-step issues
// three kind::f16 MMAs (hi*hi + hi*lo + lo*hi): ~22 mantissa bits per operand, fp32-level
// parity, at 1/3 of the f16 tensor rate.  The tensor core adds into its fp32 accumulator with
// round-toward-zero, a bias that grows with the number of accumulate steps (measured 1.3e-4 abs
// at K=4608 with one accumulator), so the two small cross terms go to a SECOND TMEM accumulator
// ("cross") and are added to the main one once, with round-to-nearest, in the epilogue: the main
// chain is K/16 steps instead of 3K/16.
//
// Pipeline (persistent CTAs, static round-robin tile schedule):
//   warp 0   TMA producer: 4 bulk-tensor loads / stage (A_hi, A_lo: 5-D NHWC boxes with
//            negative/OOB coordinates zero-filled = padding & dilation; B_hi, B_lo)
//   warp 1   TMEM allocator + tcgen05.mma issuer; tcgen05.commit frees stages
//            (both warps run their loops converged; one elect.sync lane issues -- see the producer)
//   warps 2-9 epilogue (two per TMEM lane quarter): tcgen05.ld main + cross accumulators ->
//            *scale[c] + shift[c] (+ residual) (ReLU) -> split fp16 NHWC stores; or fp32 rows for
//            the head GEMMs, which run with A and B swapped (weights on the 128 accumulator
//            lanes, pixels on the columns) so their output is channel-major like Caffe's col buffer
//   TMEM holds two (main, cross) accumulator pairs so the epilogue of tile i overlaps the MMAs
//   of tile i+1: 4 * BN columns, hence BN <= 128.
#pragma once
#include "dc_ptx.cuh"

namespace dc {

// The lane of the producer / MMA warp that issues TMA loads and MMAs: elected with elect.sync inside converged code (see the producer).
// -DDC_ISSUE_LANE0 builds the round-1 form (`lane == 0`, every issue wrapped in ptxas's per-lane waterfall) for A/B runs.
#ifdef DC_ISSUE_LANE0
#define DC_ISSUER_LANE() (lane == 0)
#else
#define DC_ISSUER_LANE() elect_one()
#endif

constexpr int kBM = 128;        // output pixels per tile (TMEM lanes)
constexpr int kBK = 64;         // fp16 channels per K-chunk = one 128-byte swizzle row
constexpr int kMaxTaps = 9;
constexpr int kConvThreads = 320;    // TMA warp, MMA warp, 8 epilogue warps (EW = 8)
constexpr int conv_threads(int ew) { return 64 + 32 * ew; }

enum OutMode : int { kOutSplitNHWC = 0, kOutF32Rows = 1, kOutF32RowsT = 2 };

// Division by a launch constant as multiply-high + shift (host: fastdiv_make).  The tile decode -- three div/mod pairs per work
// unit, done by every epilogue warp twice per tile -- compiled to ~20 SASS instructions per division (I2F / MUFU.RCP / fix-up) and
// made up a third of the lean epilogue's instruction stream (profiles/r2_ncu_summary.md); exact for 0 <= x < 2^31.
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv fastdiv_make(int d) {
  FastDiv f;
  f.d = static_cast<uint32_t>(d > 0 ? d : 1);
  if (f.d == 1) { f.mul = 0; f.shr = 0; return f; }
  uint32_t lg = 0;
  while ((1u << lg) < f.d) ++lg;                    // ceil(log2 d)
  const uint32_t p = 31 + lg;
  f.mul = static_cast<uint32_t>(((1ull << p) + f.d - 1) / f.d);
  f.shr = p - 32;
  return f;
}
__device__ __forceinline__ void fastdivmod(const FastDiv& f, int x, int& q, int& r) {
  q = f.d == 1 ? x : static_cast<int>(__umulhi(static_cast<uint32_t>(x), f.mul) >> f.shr);
  r = x - q * static_cast<int>(f.d);
}

struct ConvParams {
  int H, W;                 // input spatial dims as seen by the A tensor map
  int Ho, Wo, Cout;         // output geometry (Cout = real channel count)
  int Cin;                  // multiple of 64
  int ntaps;
  int tap_dy[kMaxTaps];     // input row/col offset of each tap relative to the output pixel
  int tap_dx[kMaxTaps];
  int tap_kblk[kMaxTaps];   // which Cin-wide block of the packed K axis holds tap t's weights.  Identity except for the 64 -> 64 channel
                            // 3x3 convs, whose taps are visited column offset by column offset -- the order TALL mode needs -- so that the
                            // plain kernel (small launches) and the TALL kernel (throughput launches) add every output element's products
                            // in the same order: the result stays bitwise independent of the batch size
  int TH, TW;               // tile rectangle, TH * TW == 128 (TW a power of two)
  int log2_tw;
  int tiles_x, tiles_y;     // per image
  int n_tiles_m;            // N * tiles_y * tiles_x
  int n_tiles_n;            // ceil(Cout / BN)
  FastDiv div_ntn, div_tx, div_ty;   // fast division by n_tiles_n, tiles_x, tiles_y
  const float* scale;       // [n_tiles_n * BN] per-channel multiplier (folded BN*Scale*weight pow2)
  const float* shift;       // [n_tiles_n * BN]
  const __half* res;        // residual (split NHWC, same geometry as the output) or nullptr
  long long res_plane;      // elements between the hi and lo planes of res
  void* out;
  long long out_plane;      // elements between hi and lo planes (split mode)
  int ldc;                  // row stride in floats (fp32-rows mode)
  int relu;
  int out_mode;
  int swap_ab;              // BN == 128 only: D[weight row][pixel] instead of D[pixel][channel]
  int in_stride;            // spatial stride of a 1x1 conv (the A map traverses W and H with this element stride); >= 1
  int early_weights;        // request the first stages' weight tiles before the grid dependency resolves (DC_EARLY_WEIGHTS)
  int debug_skip;           // MICROBENCHMARK ONLY (wrong results): after the first pipeline fill the producer stops loading the weight
                            // tiles (bit 0) / the activation tiles (bit 1); the MMAs run on whatever the stages hold.  Gives the time a
                            // weight-resident / activation-resident variant of a layer could reach before building it (DC_DEBUG_SKIP)
  int w_evict_last;         // weight tiles are loaded with the L2 evict_last priority (every CTA re-reads them for each of its pixel tiles
                            // while the activations stream through L2: -2 % of the 16x720p step, profiles/r2_chunk_sweep.md)
  // TALL mode (conv_igemm_kernel<64, 2, 8, 0, 1>: 64 -> 64 channel convs whose taps differ mostly in dy -- res2's 3x3 convs, the stem's
  // four vertical taps): the taps are grouped by column offset; per group ONE box of tall_rows = TH + (taps_per_group - 1) * row_step
  // input rows is loaded and every tap of the group is the same shared-memory tile read from a row offset (a multiple of 1024 bytes,
  // so the 128-byte-swizzle phase is unchanged); all weight tiles stay resident in shared memory.
  int tall_groups;          // tap groups (distinct dx): 3 for a 3x3, 1 for the stem
  int tall_taps_per_group;  // the taps are visited group by group: visiting position g * taps_per_group + i (tap_dy / tap_dx / tap_kblk are in that order)
  int tall_gdx[3];          // input column offset of each group
  int tall_top;             // input row of the box's first row relative to the tile's first output row
  int tall_plane_bytes;     // one plane of a box: tall_rows * TW * 128
  int tall_tap_bytes;       // shared-memory distance between consecutive taps of a group: row_step * TW * 128
  int tall_stages;          // boxes in flight (2..4)
  float* sk_ws;             // split-K scratch: [unit][peer - 1][BN columns][128 rows] fp32 partial tiles (global memory, L2-resident)
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2 MMA with M = 256)
// computes two vertically adjacent 128-pixel tiles against the same BN channels; each CTA stages its
// own 128 A rows but only BN/2 rows of B, so the bytes every SM pulls from L2 per MMA drop by 25 %
// (64 KB -> 48 KB per K-chunk) and a fourth pipeline stage fits in shared memory.
// EW = epilogue warps: 8 (two per TMEM lane quarter, two 32-channel chunks each per tile; r/rx/prefetch in
// registers) or 16 ("lean" epilogue for the epilogue-bound 1x1 expand convs: one chunk per warp, 16-column
// TMEM loads, residual brought in by cp.async, ~100 registers/thread; costs one pipeline stage of smem).
template <int BN, int CG = 1, int EW = 8>
struct ConvCfg {
  static constexpr int kABytes = kBM * kBK * 2;          // one plane of A per stage
  static constexpr int kBRows = BN / CG;                 // B rows resident in THIS CTA
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static_assert(BN == 64 || BN == 128 || (BN == 256 && CG == 2 && EW == 8), "BN = 256: CTA pairs with the 8-warp epilogue only");
  // TMEM holds kAccBufs (main, cross) accumulator pairs of BN columns each: two for BN <= 128 (the epilogue of tile i overlaps the
  // MMAs of tile i+1), ONE for BN = 256 (all 512 columns).  BN = 256 exists for the long-K 1x1 reduce convs: one A tile against 256
  // channels (64 KB per 2 x 768 MMA cycles instead of 48 KB per 768, half as many MMA issues per FLOP) measured ~7 % faster per unit
  // of work than two 128-channel tiles while every MMA issue cost a ~14-instruction waterfall; with the elect.sync issue path the
  // 128-channel pairs are faster everywhere it had been chosen, so dc_conv_forward no longer takes it unless DC_CONV_BN256 asks
  // (profiles/r2_round2_sweeps.md).  Kept: it is bitwise-neutral, tested, and the tile a longer-K layer would want.
  static constexpr int kAccBufs = BN == 256 ? 1 : 2;
  static_assert(EW == 8 || (EW == 16 && BN == 128), "the 16-warp epilogue owns one 32-channel chunk per warp: BN = 128");
  static constexpr int kStages = BN == 256 ? 3 : (CG == 2 ? 4 : (BN >= 128 ? 3 : 4)) - (EW == 16 ? 1 : 0);
  static constexpr int kTmemCols = 2 * BN * kAccBufs;              // kAccBufs x (main, cross)
  static constexpr int kStagingBytes = EW * 4096;        // per epilogue warp: [32 px][32 ch] fp16 x {hi, lo}
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  // TALL mode: box ring + resident weights share what one CTA can have (same launch size as the four-stage pair kernel)
  static constexpr int kTallOperandBytes = 192 * 1024;
  static constexpr int kTallSmemBytes = kTallOperandBytes + kStagingBytes + 1024 + 256;
};

// SK = 1 ("split-K", the latency regime: fewer work units than SMs): a cluster of S = 2 or 4 CTAs shares ONE unit,
// CTA r accumulating K-steps [ksteps * r / S, ksteps * (r + 1) / S) into its own TMEM.  Each peer (r > 0) then adds
// main + cross and writes the fp32 partial tile to a global scratch slot (it stays in L2: an SM moves ~64 B/clk to and
// from L2 but only ~17 B/clk over distributed shared memory -- a version that exchanged the tiles by bulk DSMEM copies paid
// ~7.5 us per launch, this one ~4.7 us, profiles/r1_microbench_latency.txt), fences, and arrives (release, cluster scope) on an
// mbarrier in the leader's shared memory; the leader's epilogue waits on it (acquire, cluster scope), reads the slots
// with L1-bypassing loads and sums own + slot 0 + slot 1 + ... in that fixed order (no atomics: repeated runs are bitwise
// identical).  The grid is exactly units * S CTAs, so nothing is persistent in this mode.
// work unit -> (n-tile, tile x, tile y, image); a phantom tile (m-tile == n_tiles_m, the odd CTA of the last pair) decodes to
// image == N: all of its TMA boxes are out of bounds
template <int CG>
__device__ __forceinline__ void decode_unit(const ConvParams& p, int unit, int cta_rank, int& nt, int& mt, int& tx, int& ty, int& img) {
  int mg;
  fastdivmod(p.div_ntn, unit, mg, nt);
  mt = mg * CG + cta_rank;
  int rest;
  fastdivmod(p.div_tx, mt, rest, tx);
  fastdivmod(p.div_ty, rest, img, ty);
}

template <int BN, int CG, int EW, int SK = 0, int TALL = 0>
__global__ void __launch_bounds__(conv_threads(EW), 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const ConvParams p) {
  using Cfg = ConvCfg<BN, CG, EW>;
  static_assert(SK == 0 || (CG == 1 && EW == 8), "split-K runs on single CTAs with the 8-warp epilogue");
  static_assert(TALL == 0 || (BN == 64 && CG == 2 && EW == 8 && SK == 0), "TALL: 64-channel tiles, CTA pairs, 8-warp epilogue");
  const int cta_rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int ksplit = SK ? static_cast<int>(cluster_nctarank()) : 1;
  const int krank = SK ? static_cast<int>(cluster_ctarank()) : 0;
  // work units: (m-tile group of CG tiles, n-tile); unit u -> n-tile u % n_tiles_n, m-tile CG * (u / n_tiles_n) + rank
  const int unit_first = SK ? static_cast<int>(blockIdx.x) / ksplit : (CG == 2 ? (blockIdx.x >> 1) : blockIdx.x);
  const int unit_stride = SK ? static_cast<int>(gridDim.x) / ksplit : (CG == 2 ? (gridDim.x >> 1) : gridDim.x);
  const int total_units = ((p.n_tiles_m + CG - 1) / CG) * p.n_tiles_n;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* staging_all = smem + (TALL ? Cfg::kTallOperandBytes : kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging_all + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  [[maybe_unused]] uint64_t* red_full = bars + 2 * kStages + 5;    // split-K leader: every peer has published its partial tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kchunks = p.Cin / kBK;
  const int ksteps = p.ntaps * kchunks;
  const int ks_begin = SK ? ksteps * krank / ksplit : 0;            // this CTA's share of the K loop
  const int ks_end = SK ? ksteps * (krank + 1) / ksplit : ksteps;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.out_mode == kOutSplitNHWC) prefetch_tmap(&tmO);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EW * CG);     // every epilogue warp of every CTA of the group
    }
    if (SK) {
      mbar_init(red_full, static_cast<uint32_t>(ksplit - 1));
    }
    if (TALL) mbar_init(red_full, 1);          // TALL: "resident weights have landed" (the slot split-K uses otherwise)
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2cta<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (CG == 2 || SK) cluster_sync_all();       // peer barriers are initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the
  // previous kernel; from here on we read what it wrote.
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp runs the loop converged and ONE ELECTED lane issues (elect.sync).  Under `if (lane == 0)` ptxas cannot prove
    // that the operands of the uniform-datapath instructions (UTMALDG here, UTCHMMA / UTCBAR in the MMA warp) are warp-uniform and
    // wraps every one of them in a per-lane waterfall (ELECT, four or five R2UR.BROADCAST, two PLOP3, a back branch: ~14 SASS
    // instructions per instruction issued); inside an elect.sync region of converged code they are plain uniform-register operands.
    if constexpr (TALL != 0) {
      const uint64_t pol_w = l2_policy(p.w_evict_last ? 2 : 0);
      const uint32_t plane_bytes = static_cast<uint32_t>(p.tall_plane_bytes), stage_bytes = 2u * plane_bytes;
      const int nst = p.tall_stages;
      uint8_t* bres = smem + nst * stage_bytes;                   // [tap][plane][32 rows][128 B]
      // every weight tile of the layer, once per CTA, before griddepcontrol.wait (weights do not depend on the predecessor)
      if (DC_ISSUER_LANE()) {
        if (cta_rank == 0) mbar_expect_tx(red_full, 2u * static_cast<uint32_t>(p.ntaps) * 2u * Cfg::kBBytes);
        for (int t = 0; t < p.ntaps; ++t) {
          tma_load_3d_2cta(bres + (2 * t) * Cfg::kBBytes, &tmB, red_full, t * kBK, cta_rank * Cfg::kBRows, 0, pol_w);
          tma_load_3d_2cta(bres + (2 * t + 1) * Cfg::kBBytes, &tmB, red_full, t * kBK, cta_rank * Cfg::kBRows, 1, pol_w);
        }
      }
      __syncwarp();
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit_first; unit < total_units; unit += unit_stride) {
        int nt, mt, tx, ty, img;
        decode_unit<CG>(p, unit, cta_rank, nt, mt, tx, ty, img);
        const int x0 = tx * p.TW, y0 = ty * p.TH + p.tall_top;
        for (int g = 0; g < p.tall_groups; ++g) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (DC_ISSUER_LANE()) {
            uint8_t* sa = smem + stage * stage_bytes;
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2u * stage_bytes);      // both CTAs' boxes, both planes
            tma_load_5d_2cta(sa, &tmA, &full_bar[stage], 0, x0 + p.tall_gdx[g], y0, img, 0);
            tma_load_5d_2cta(sa + plane_bytes, &tmA, &full_bar[stage], 0, x0 + p.tall_gdx[g], y0, img, 1);
          }
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      }
    } else {
      const uint64_t pol_w = l2_policy(p.w_evict_last ? 2 : 0);
      // Weights do not depend on the predecessor kernel: the weight tiles of this CTA's first K-steps (one per pipeline
      // stage) are requested BEFORE griddepcontrol.wait, so their HBM latency (a single image re-reads all 251 MB of
      // packed weights from HBM every forward) overlaps the predecessor's tail; the activation tiles follow after the wait.
      int npre = 0;
      if (p.early_weights && unit_first < total_units) {
        int q0, nt0;
        fastdivmod(p.div_ntn, unit_first, q0, nt0);
        const int n0 = nt0 * BN + cta_rank * Cfg::kBRows;
        for (int ks = ks_begin; ks < ks_end && npre < kStages; ++ks, ++npre) {
          uint8_t* sa = smem + npre * Cfg::kStageBytes;
          const int kcoord0 = (p.tap_kblk[ks / kchunks] * kchunks + ks % kchunks) * kBK;
          if (DC_ISSUER_LANE()) {
            if (CG == 2) {
              if (cta_rank == 0) mbar_expect_tx(&full_bar[npre], 2 * Cfg::kStageBytes);
              tma_load_3d_2cta(sa + 2 * Cfg::kABytes, &tmB, &full_bar[npre], kcoord0, n0, 0, pol_w);
              tma_load_3d_2cta(sa + 2 * Cfg::kABytes + Cfg::kBBytes, &tmB, &full_bar[npre], kcoord0, n0, 1, pol_w);
            } else {
              mbar_expect_tx(&full_bar[npre], Cfg::kStageBytes);
              tma_load_3d(sa + 2 * Cfg::kABytes, &tmB, &full_bar[npre], kcoord0, n0, 0, pol_w);
              tma_load_3d(sa + 2 * Cfg::kABytes + Cfg::kBBytes, &tmB, &full_bar[npre], kcoord0, n0, 1, pol_w);
            }
          }
          __syncwarp();
        }
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      int issued = 0;
      for (int unit = unit_first; unit < total_units; unit += unit_stride) {
        int nt, mt, tx, ty, img;
        decode_unit<CG>(p, unit, cta_rank, nt, mt, tx, ty, img);
        const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = nt * BN + cta_rank * Cfg::kBRows;
        for (int t = 0; t < p.ntaps; ++t) {
          const int ix = x0 * p.in_stride + p.tap_dx[t], iy = y0 * p.in_stride + p.tap_dy[t];
          for (int kc = 0; kc < kchunks; ++kc) {
            if (SK && (t * kchunks + kc < ks_begin || t * kchunks + kc >= ks_end)) continue;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            const int kcoord = (p.tap_kblk[t] * kchunks + kc) * kBK;
            const bool fresh = issued >= npre;      // else: this stage's barrier is armed and its weight tiles are on their way
            if (SK == 0 && p.debug_skip && issued >= kStages) {
              // microbenchmark mode: arm the barrier for exactly what is still loaded
              const bool ld_a = !(p.debug_skip & 2), ld_b = !(p.debug_skip & 1);
              const uint32_t bytes = (ld_a ? 2u * Cfg::kABytes : 0u) + (ld_b ? 2u * Cfg::kBBytes : 0u);
              ++issued;
              if (!DC_ISSUER_LANE()) {
              } else if (CG == 2) {
                if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * bytes);
                if (ld_a) {
                  tma_load_5d_2cta(sa, &tmA, &full_bar[stage], kc * kBK, ix, iy, img, 0);
                  tma_load_5d_2cta(sa + Cfg::kABytes, &tmA, &full_bar[stage], kc * kBK, ix, iy, img, 1);
                }
                if (ld_b) {
                  tma_load_3d_2cta(sa + 2 * Cfg::kABytes, &tmB, &full_bar[stage], kcoord, n0, 0, pol_w);
                  tma_load_3d_2cta(sa + 2 * Cfg::kABytes + Cfg::kBBytes, &tmB, &full_bar[stage], kcoord, n0, 1, pol_w);
                }
              } else {
                mbar_expect_tx(&full_bar[stage], bytes);
                if (ld_a) {
                  tma_load_5d(sa, &tmA, &full_bar[stage], kc * kBK, ix, iy, img, 0);
                  tma_load_5d(sa + Cfg::kABytes, &tmA, &full_bar[stage], kc * kBK, ix, iy, img, 1);
                }
                if (ld_b) {
                  tma_load_3d(sa + 2 * Cfg::kABytes, &tmB, &full_bar[stage], kcoord, n0, 0, pol_w);
                  tma_load_3d(sa + 2 * Cfg::kABytes + Cfg::kBBytes, &tmB, &full_bar[stage], kcoord, n0, 1, pol_w);
                }
              }
              __syncwarp();
              if (++stage == kStages) { stage = 0; phase ^= 1; }
              continue;
            }
            ++issued;
            if (!DC_ISSUER_LANE()) {
            } else if (CG == 2) {
              // both CTAs' loads count on the leader's barrier; only the leader arms it (for both halves)
              if (cta_rank == 0 && fresh) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              tma_load_5d_2cta(sa, &tmA, &full_bar[stage], kc * kBK, ix, iy, img, 0);
              tma_load_5d_2cta(sa + Cfg::kABytes, &tmA, &full_bar[stage], kc * kBK, ix, iy, img, 1);
              if (fresh) {
                tma_load_3d_2cta(sa + 2 * Cfg::kABytes, &tmB, &full_bar[stage], kcoord, n0, 0, pol_w);
                tma_load_3d_2cta(sa + 2 * Cfg::kABytes + Cfg::kBBytes, &tmB, &full_bar[stage], kcoord, n0, 1, pol_w);
              }
            } else {
              if (fresh) mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              tma_load_5d(sa, &tmA, &full_bar[stage], kc * kBK, ix, iy, img, 0);
              tma_load_5d(sa + Cfg::kABytes, &tmA, &full_bar[stage], kc * kBK, ix, iy, img, 1);
              if (fresh) {
                tma_load_3d(sa + 2 * Cfg::kABytes, &tmB, &full_bar[stage], kcoord, n0, 0, pol_w);
                tma_load_3d(sa + 2 * Cfg::kABytes + Cfg::kBBytes, &tmB, &full_bar[stage], kcoord, n0, 1, pol_w);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (cta_rank == 0) {             // the leader CTA issues for the whole group: converged warp, one elected lane per K-step (see the producer)
      constexpr uint32_t idesc = umma_idesc_f16(kBM * CG, BN);
      [[maybe_unused]] constexpr uint32_t idesc_wide = umma_idesc_f16(kBM, 2 * BN);     // CG = 1: fused hi*hi | hi*lo
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if constexpr (TALL != 0) {
        const uint32_t plane_bytes = static_cast<uint32_t>(p.tall_plane_bytes), stage_bytes = 2u * plane_bytes;
        const int nst = p.tall_stages;
        const uint32_t bres = smem_u32(smem + nst * stage_bytes);
        mbar_wait(red_full, 0);                       // the resident weight tiles have landed (both CTAs')
        tc_fence_after();
        for (int unit = unit_first; unit < total_units; unit += unit_stride) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + static_cast<uint32_t>(acc * 2 * BN);
          const uint32_t dx = d + BN;
          for (int g = 0; g < p.tall_groups; ++g) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * stage_bytes);
            if (DC_ISSUER_LANE()) {
              for (int i = 0; i < p.tall_taps_per_group; ++i) {
                const int tap = p.tap_kblk[g * p.tall_taps_per_group + i];        // the packed K block of the group's i-th tap
                const uint32_t aoff = static_cast<uint32_t>(i * p.tall_tap_bytes);      // a multiple of 1024: same swizzle phase
                const uint64_t a_hi = umma_desc_k_sw128(sa + aoff);
                const uint64_t a_lo = umma_desc_k_sw128(sa + plane_bytes + aoff);
                const uint64_t b_hi = umma_desc_k_sw128(bres + static_cast<uint32_t>(2 * tap) * Cfg::kBBytes);
                const uint64_t b_lo = umma_desc_k_sw128(bres + static_cast<uint32_t>(2 * tap + 1) * Cfg::kBBytes);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                  const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);
                  const uint32_t accum = (g | i | k) != 0 ? 1u : 0u;          // the unit's first MMAs overwrite the accumulators
                  umma_f16_2cta(d, a_hi + adv, b_hi + adv, idesc, accum);
                  umma_f16_2cta(dx, a_hi + adv, b_lo + adv, idesc, accum);
                  umma_f16_2cta(dx, a_lo + adv, b_hi + adv, idesc, 1);
                }
              }
              umma_commit_2cta(&empty_bar[stage]);
              if (g + 1 == p.tall_groups) umma_commit_2cta(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++stage == nst) { stage = 0; phase ^= 1; }
          }
          if (++acc == Cfg::kAccBufs) { acc = 0; acc_phase ^= 1; }
        }
      } else
      for (int unit = unit_first; unit < total_units; unit += unit_stride) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>(acc * 2 * BN);   // main
        const uint32_t dx = d + BN;                                            // cross terms
        for (int ks = ks_begin; ks < ks_end; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t a_hi = umma_desc_k_sw128(sa);
          const uint64_t a_lo = umma_desc_k_sw128(sa + Cfg::kABytes);
          const uint64_t b_hi = umma_desc_k_sw128(sa + 2 * Cfg::kABytes);
          const uint64_t b_lo = umma_desc_k_sw128(sa + 2 * Cfg::kABytes + Cfg::kBBytes);
          const uint32_t later = ks != ks_begin;      // 0 on the unit's first K-step: its first MMAs overwrite the accumulators
          if (DC_ISSUER_LANE()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t adv = static_cast<uint64_t>((k * 32) >> 4);   // 16 fp16 = 32 B along K
              const uint32_t accum = k > 0 ? 1u : later;
              // hi*hi, hi*lo and lo*hi are symmetric in (A, B): swapping the operands only transposes D
              if (CG == 2) {
                umma_f16_2cta(d, a_hi + adv, b_hi + adv, idesc, accum);
                umma_f16_2cta(dx, a_hi + adv, b_lo + adv, idesc, accum);
                umma_f16_2cta(dx, a_lo + adv, b_hi + adv, idesc, 1);
              } else if (p.swap_ab) {
                // hi*hi and hi*lo share their M-side operand: one MMA with N = 2*BN over the stacked [a_hi; a_lo] rows
                // (contiguous in the stage) writes main | cross side by side and reads the shared operand once
                umma_f16(d, b_hi + adv, a_hi + adv, idesc_wide, accum);
                umma_f16(dx, b_lo + adv, a_hi + adv, idesc, 1);
              } else {
                umma_f16(d, a_hi + adv, b_hi + adv, idesc_wide, accum);      // [b_hi; b_lo] rows are contiguous too
                umma_f16(dx, a_lo + adv, b_hi + adv, idesc, 1);
              }
            }
            if (CG == 2) umma_commit_2cta(&empty_bar[stage]);
            else umma_commit(&empty_bar[stage]);
            if (ks + 1 == ks_end) {                 // the unit's last K-step: publish the accumulators
              if (CG == 2) umma_commit_2cta(&tfull_bar[acc]);
              else umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == Cfg::kAccBufs) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (EW == 16) {
    // ------------------------------------------------------------ lean epilogue (warps 2..17), split-NHWC only
    // (Round 2 tried two staging tiles per warp -- the residual of tile i+1 streaming in while tile i is converted -- paid for
    // with one operand stage (227 KB of shared memory hold 3 x 48 KB stages + 64 KB of staging, or 2 + 128): 112.7 -> 132.6 us
    // on res4's 2c: with two stages the mainloop starves and the faster epilogue only waits longer on `tmem_full`.  For this shape
    // shared memory holds three operand stages or a double-buffered epilogue, not both; profiles/r2_ncu_summary.md.)
    // One 32-pixel x 32-channel chunk per warp per tile.  The residual chunk is copied global -> staging by
    // cp.async as soon as the previous tile's TMA stores have drained the staging tile, i.e. while the
    // next accumulator is still being computed; accumulators are read 16 columns at a time.
    const int q = warp & 3;
    const int c0 = ((warp - 2) >> 2) * 32;
    uint8_t* stg = staging_all + (warp - 2) * 4096;
    const int piece = lane & 3;
    const int own_sw = (lane >> 1) & 3;
    const bool has_res = p.res != nullptr;
    auto issue_residual = [&](int u) {
      if (u >= total_units) return;
      int t_nt, t_mt, t_tx, t_ty, t_img;
      decode_unit<CG>(p, u, cta_rank, t_nt, t_mt, t_tx, t_ty, t_img);
      if (t_mt >= p.n_tiles_m || t_nt * BN + c0 >= p.Cout) return;
      const int yy0 = t_ty * p.TH + ((q * 32) >> p.log2_tw), xx0 = t_tx * p.TW + ((q * 32) & (p.TW - 1));
      const long long pix0 = (static_cast<long long>(t_img) * p.Ho + yy0) * p.Wo + xx0;
      const __half* base = p.res + t_nt * BN + c0 + piece * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = (lane >> 2) + 8 * i;
        const int dy = rr >> p.log2_tw, dx = rr & (p.TW - 1);
        if (yy0 + dy < p.Ho && xx0 + dx < p.Wo) {      // rows outside the image are clipped by the TMA store: leave garbage
          const __half* src = base + (pix0 + dy * p.Wo + dx) * p.Cout;
          uint8_t* dst = stg + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4);
          cp_async_16(dst, src);
          cp_async_16(dst + 2048, src + p.res_plane);
        }
      }
      cp_async_commit();
    };
    if (has_res) issue_residual(unit_first);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = unit_first; unit < total_units; unit += unit_stride) {
      int nt, mt, tx, ty, img;
      decode_unit<CG>(p, unit, cta_rank, nt, mt, tx, ty, img);
      const int n0 = nt * BN;
      const bool chunk_ok = n0 + c0 < p.Cout;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 2 * BN) + c0;
      uint8_t* my_hi = stg + lane * 64;
      uint8_t* my_lo = my_hi + 2048;
      constexpr uint16_t kOneH = 0x3C00, kMinusOneH = 0xBC00;
      auto process16 = [&](const uint32_t (&a)[16], const uint32_t (&b)[16], int g0) {
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const int g = g0 + gg;
          const int slot = (g ^ own_sw) << 4;
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + c0) + 2 * g);
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + c0) + 2 * g + 1);
          const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.shift + n0 + c0) + 2 * g);
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(p.shift + n0 + c0) + 2 * g + 1);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            v[e] = fmaf(__uint_as_float(a[gg * 8 + e]) + __uint_as_float(b[gg * 8 + e]), sc[e], sh[e]);
          if (has_res) {
            const uint4 h4 = *reinterpret_cast<const uint4*>(my_hi + slot);
            const uint4 l4 = *reinterpret_cast<const uint4*>(my_lo + slot);
            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint16_t h0, h1, l0, l1;
              unpack_h2(hw[e], h0, h1);
              unpack_h2(lw[e], l0, l1);
              v[e * 2 + 0] = fma_hhf(l0, kOneH, fma_hhf(h0, kOneH, v[e * 2 + 0]));
              v[e * 2 + 1] = fma_hhf(l1, kOneH, fma_hhf(h1, kOneH, v[e * 2 + 1]));
            }
          }
          uint32_t ho[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x0 = v[e * 2], x1 = v[e * 2 + 1];
            if (p.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
            ho[e] = pack_f2h2_rn(x0, x1);
            uint16_t ha, hb;
            unpack_h2(ho[e], ha, hb);
            lo[e] = pack_f2h2_rn(fma_hhf(ha, kMinusOneH, x0), fma_hhf(hb, kMinusOneH, x1));
          }
          *reinterpret_cast<uint4*>(my_hi + slot) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
          *reinterpret_cast<uint4*>(my_lo + slot) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      };
      if (chunk_ok) {
        uint32_t a[16], b[16];
        tmem_ld_32x16(taddr, a);
        tmem_ld_32x16(taddr + BN, b);
        if (has_res) { cp_async_wait_all(); __syncwarp(); }      // this tile's residual chunk is in the staging tile
        tmem_ld_wait();
        process16(a, b, 0);
        tmem_ld_32x16(taddr + 16, a);
        tmem_ld_32x16(taddr + BN + 16, b);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                          // the accumulator is free for tile i+2
          if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);
          else mbar_arrive(&tempty_bar[acc]);
        }
        process16(a, b, 2);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int r0 = q * 32;
          const int box_x = tx * p.TW + (r0 & (p.TW - 1)), box_y = ty * p.TH + (r0 >> p.log2_tw);
          tma_store_5d(&tmO, stg, n0 + c0, box_x, box_y, img, 0);
          tma_store_5d(&tmO, stg + 2048, n0 + c0, box_x, box_y, img, 1);
          tma_store_commit();
          tma_store_wait_read();                                  // staging tile drained: safe to refill
        }
        __syncwarp();
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);
          else mbar_arrive(&tempty_bar[acc]);
        }
      }
      if (has_res) issue_residual(unit + unit_stride);
      if (++acc == Cfg::kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quarter; they split the tile's 32-column chunks between them.
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;       // which of the two warps of this quarter
    const int row = q * 32 + lane;          // accumulator row (TMEM lane) this thread owns
    // residual prefetch registers (split-NHWC mode): 4 rows x {hi, lo} x 16 B of the NEXT chunk
    uint4 res_h[4], res_l[4];
    auto prefetch_residual = [&](int u, int c0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { res_h[i] = make_uint4(0, 0, 0, 0); res_l[i] = make_uint4(0, 0, 0, 0); }
      if (u >= total_units) return;
      int t_nt, t_mt, t_tx, t_ty, t_img;
      decode_unit<CG>(p, u, cta_rank, t_nt, t_mt, t_tx, t_ty, t_img);
      if (t_mt >= p.n_tiles_m) return;
      if (t_nt * BN + c0 >= p.Cout) return;
      // rows (lane >> 2) + 8 i of the warp's 32-row sub-rectangle (TW is a power of two: shifts, no divides)
      const int yy0 = t_ty * p.TH + ((q * 32) >> p.log2_tw), xx0 = t_tx * p.TW + ((q * 32) & (p.TW - 1));
      const long long pix0 = (static_cast<long long>(t_img) * p.Ho + yy0) * p.Wo + xx0;
      const __half* base = p.res + t_nt * BN + c0 + (lane & 3) * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = (lane >> 2) + 8 * i;
        const int dy = rr >> p.log2_tw, dx = rr & (p.TW - 1);
        if (yy0 + dy < p.Ho && xx0 + dx < p.Wo) {
          const __half* src = base + (pix0 + dy * p.Wo + dx) * p.Cout;
          // plain (coherent) loads: the output may alias the residual (in-place block output, dc_engine.cpp), which the
          // read-only path must not see
          res_h[i] = *reinterpret_cast<const uint4*>(src);
          res_l[i] = *reinterpret_cast<const uint4*>(src + p.res_plane);
        }
      }
    };
    if (p.out_mode == kOutSplitNHWC && p.res != nullptr && krank == 0) prefetch_residual(unit_first, half * 32);
    // split-K: fp32 partial tiles of the peers, slot s = CTA s + 1, [BN columns][128 rows] floats, in the stage memory
    [[maybe_unused]] const float* slots = p.sk_ws + static_cast<long long>(unit_first) * (ksplit - 1) * (BN * kBM) + row;
    [[maybe_unused]] auto add_partials = [&](uint32_t (&r)[32], uint32_t (&rx)[32], int c0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rx[j]));
        rx[j] = 0u;
      }
      // one slot at a time, all 32 loads of a slot in flight together (they are L2 round trips: issued one by one
      // behind their adds they cost ~19 us per launch, profiles/r1_microbench_latency.txt); rank order keeps the sum deterministic
      for (int s_ = 0; s_ < ksplit - 1; ++s_) {
        float pv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = __ldcg(slots + (s_ * BN + c0 + j) * kBM);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + pv[j]);
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = unit_first; unit < total_units; unit += unit_stride) {
      int nt, mt, tx, ty, img;
      decode_unit<CG>(p, unit, cta_rank, nt, mt, tx, ty, img);
      const bool tile_ok = mt < p.n_tiles_m;       // the odd CTA of the last pair may own a phantom tile
      const int n0 = nt * BN;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 2 * BN);
      if constexpr (SK != 0) {
        if (krank == 0) {
          mbar_wait_cluster(red_full, 0);              // every peer's tile is in L2 and visible
        } else {
          float* mine = p.sk_ws + (static_cast<long long>(unit) * (ksplit - 1) + (krank - 1)) * (BN * kBM) + row;
#pragma unroll 1
          for (int c0 = half * 32; c0 < BN; c0 += 64) {
            uint32_t r[32], rx[32];
            tmem_ld_32x32(taddr + c0, r);
            tmem_ld_32x32(taddr + BN + c0, rx);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcg(mine + (c0 + j) * kBM, __uint_as_float(r[j]) + __uint_as_float(rx[j]));   // 128 B per warp store
          }
          // publish: the CTA barrier orders every epilogue thread's stores before the elected thread's gpu-scope release
          // fence (cumulative), which the remote arrive (release, cluster scope) follows
          named_bar_sync(1, EW * 32);
          if (warp == 2 && lane == 0) {
            fence_acq_rel_gpu();
            mbar_arrive_cluster(mapa_shared(smem_u32(red_full), 0));
          }
          continue;                                    // single unit per cluster: the peers are done
        }
      }
      if (p.out_mode == kOutF32RowsT) {
        // swapped operands: row = weight row (output channel), columns = the tile's 128 pixels
        // (flat 1x1 geometry: pixel = tx * 128 + column).  out[(n0 + row) * ldc + pixel], fp32.
        const float sc = __ldg(p.scale + n0 + row), sh = __ldg(p.shift + n0 + row);
        float* orow = reinterpret_cast<float*>(p.out) + static_cast<long long>(n0 + row) * p.ldc;
#pragma unroll 1
        for (int c0 = half * 32; c0 < kBM; c0 += 64) {
          uint32_t r[32], rx[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_32x32(taddr + BN + c0, rx);
          tmem_ld_wait();
          const int pix0 = tx * p.TW + c0;
          if (pix0 + 32 <= p.Wo) {
            float4* o = reinterpret_cast<float4*>(orow + pix0);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float4 f;
              f.x = fmaf(__uint_as_float(r[4 * g + 0]) + __uint_as_float(rx[4 * g + 0]), sc, sh);
              f.y = fmaf(__uint_as_float(r[4 * g + 1]) + __uint_as_float(rx[4 * g + 1]), sc, sh);
              f.z = fmaf(__uint_as_float(r[4 * g + 2]) + __uint_as_float(rx[4 * g + 2]), sc, sh);
              f.w = fmaf(__uint_as_float(r[4 * g + 3]) + __uint_as_float(rx[4 * g + 3]), sc, sh);
              o[g] = f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (pix0 + j < p.Wo) orow[pix0 + j] = fmaf(__uint_as_float(r[j]) + __uint_as_float(rx[j]), sc, sh);
          }
        }
      } else if (p.out_mode == kOutSplitNHWC) {
        // Coalesced path: the warp's 32 pixels x 32 channels chunk is staged in shared memory in the
        // TMA SWIZZLE_64B layout ([row][64 B], 16-byte piece index ^= (row >> 1) & 3, which is also
        // bank-conflict free for both access patterns below).  The residual is fetched cooperatively
        // (lane -> 16 B piece (lane & 3) of rows (lane >> 2) + 8 i: full 32 B sectors) ONE CHUNK AHEAD
        // (registers res_h/res_l, loaded while the previous chunk is converted and stored and while the
        // next accumulator is still being computed); the result leaves through two TMA bulk-tensor
        // stores (hi and lo plane) that clip ragged tile edges.
        uint8_t* stg = staging_all + (warp - 2) * 4096;
        const int r0 = q * 32;
        const int box_x = tx * p.TW + (r0 & (p.TW - 1)), box_y = ty * p.TH + (r0 >> p.log2_tw);
        const int piece = lane & 3;
        const int own_sw = (lane >> 1) & 3;                // swizzle of this thread's own row (row index == lane)
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          if (n0 + c0 >= p.Cout) {
            // ragged last channel tile: nothing to do here, but if this was the warp's first chunk the
            // prefetch registers still hold the zeros meant for it -- refill them for the next tile
            if (p.res != nullptr && c0 == half * 32) prefetch_residual(unit + unit_stride, half * 32);
            break;
          }
          uint32_t r[32], rx[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_32x32(taddr + BN + c0, rx);
          if (lane == 0) tma_store_wait_read();            // previous chunk's stores have drained the staging tile
          __syncwarp();
          if (p.res != nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = (lane >> 2) + 8 * i;
              const int sw = (rr >> 1) & 3;
              *reinterpret_cast<uint4*>(stg + rr * 64 + ((piece ^ sw) << 4)) = res_h[i];
              *reinterpret_cast<uint4*>(stg + 2048 + rr * 64 + ((piece ^ sw) << 4)) = res_l[i];
            }
            __syncwarp();
            // prefetch the residual of the chunk this warp handles next (same tile, or the next tile's first)
            int nunit = unit, nc0 = c0 + 64;
            if (nc0 >= BN || nt * BN + nc0 >= p.Cout) { nunit = unit + unit_stride; nc0 = half * 32; }
            prefetch_residual(nunit, nc0);
          }
          float sc[32], sh[32];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + c0) + g);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.shift + n0 + c0) + g);
            sc[4 * g] = a4.x; sc[4 * g + 1] = a4.y; sc[4 * g + 2] = a4.z; sc[4 * g + 3] = a4.w;
            sh[4 * g] = b4.x; sh[4 * g + 1] = b4.y; sh[4 * g + 2] = b4.z; sh[4 * g + 3] = b4.w;
          }
          tmem_ld_wait();
          if constexpr (SK != 0) add_partials(r, rx, c0);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]) + __uint_as_float(rx[j]), sc[j], sh[j]);
          uint8_t* my_hi = stg + lane * 64;
          uint8_t* my_lo = my_hi + 2048;
          constexpr uint16_t kOneH = 0x3C00, kMinusOneH = 0xBC00;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int slot = (g ^ own_sw) << 4;
            if (p.res != nullptr) {
              const uint4 h4 = *reinterpret_cast<const uint4*>(my_hi + slot);
              const uint4 l4 = *reinterpret_cast<const uint4*>(my_lo + slot);
              const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                uint16_t h0, h1, l0, l1;
                unpack_h2(hw[e], h0, h1);
                unpack_h2(lw[e], l0, l1);
                // v += hi + lo, one FHFMA each (fp16 * 1 + fp32)
                v[g * 8 + e * 2 + 0] = fma_hhf(l0, kOneH, fma_hhf(h0, kOneH, v[g * 8 + e * 2 + 0]));
                v[g * 8 + e * 2 + 1] = fma_hhf(l1, kOneH, fma_hhf(h1, kOneH, v[g * 8 + e * 2 + 1]));
              }
            }
            uint32_t ho[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = v[g * 8 + e * 2], b = v[g * 8 + e * 2 + 1];
              if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
              ho[e] = pack_f2h2_rn(a, b);
              uint16_t ha, hb;
              unpack_h2(ho[e], ha, hb);
              lo[e] = pack_f2h2_rn(fma_hhf(ha, kMinusOneH, a), fma_hhf(hb, kMinusOneH, b));   // x - hi, exact in fp32
            }
            *reinterpret_cast<uint4*>(my_hi + slot) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
            *reinterpret_cast<uint4*>(my_lo + slot) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async_smem();                         // generic-proxy writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tmO, stg, n0 + c0, box_x, box_y, img, 0);
            tma_store_5d(&tmO, stg + 2048, n0 + c0, box_x, box_y, img, 1);
            tma_store_commit();
          }
        }
      } else {   // kOutF32Rows
        const int oy = ty * p.TH + (row >> p.log2_tw);
        const int ox = tx * p.TW + (row & (p.TW - 1));
        const bool valid = tile_ok && (oy < p.Ho) && (ox < p.Wo);
        const long long pix = (static_cast<long long>(img) * p.Ho + oy) * p.Wo + ox;
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          uint32_t r[32], rx[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_32x32(taddr + BN + c0, rx);
          tmem_ld_wait();
          if constexpr (SK != 0) add_partials(r, rx, c0);
          if (!valid) continue;
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.ldc + n0 + c0);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 f;
            f.x = fmaf(__uint_as_float(r[4 * g + 0]) + __uint_as_float(rx[4 * g + 0]), __ldg(p.scale + n0 + c0 + 4 * g + 0), __ldg(p.shift + n0 + c0 + 4 * g + 0));
            f.y = fmaf(__uint_as_float(r[4 * g + 1]) + __uint_as_float(rx[4 * g + 1]), __ldg(p.scale + n0 + c0 + 4 * g + 1), __ldg(p.shift + n0 + c0 + 4 * g + 1));
            f.z = fmaf(__uint_as_float(r[4 * g + 2]) + __uint_as_float(rx[4 * g + 2]), __ldg(p.scale + n0 + c0 + 4 * g + 2), __ldg(p.shift + n0 + c0 + 4 * g + 2));
            f.w = fmaf(__uint_as_float(r[4 * g + 3]) + __uint_as_float(rx[4 * g + 3]), __ldg(p.scale + n0 + c0 + 4 * g + 3), __ldg(p.shift + n0 + c0 + 4 * g + 3));
            if (p.relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
            o[g] = f;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == Cfg::kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();     // bulk stores must complete before the CTA's shared memory goes away
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();             // nobody touches the peer's barriers / TMEM after this
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace dc
