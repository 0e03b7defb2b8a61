// conv3d_tc.cu -- tcgen05 (5th-gen tensor core) implicit-GEMM 3x3x3 convolution family for sm_100a.
//
// Replaces the cuDNN Conv3d / ConvTranspose3d + BatchNorm3d + ReLU (+ add) chains of
// disprcnn/modeling/psmnet/stackhourglass.py:11-30,63-88 (applied :130-144): stride-1 convs,
// stride-2 convs, the stride-2 transposed convs and the 32->1 classifier convs -- all 28 layers.
// No im2col, no materialised patches.
//
// GEMM view (all modes).  Activations are bf16, channel-blocked-8: [N][C/8][D][H][W][8]
// (common.cuh); one voxel of one channel block is exactly the 16-byte row of a NO-SWIZZLE K-major
// UMMA core matrix.  M = 128 rows = an 8 (w) x 16 (h) tile of ONE plane of the "row grid"
// (output grid for convs, input grid for the transposed conv), K = Cin per filter tap (K=16 per
// MMA), N = 32 output channels per stacked block.
// A operand.  TMA (cp.async.bulk.tensor over (8ch*W, H, D, N*Cin/8): 144/160-byte rows, zero OOB
//   fill == conv padding) lands a haloed plane tile in shared memory as [Cin/8][rows][cols][8]:
//   GEMM rows (w) 16 B apart, 8-row groups (h) one tile row apart (SBO), K core matrices one
//   channel-block plane apart (LBO).  Every in-plane tap is the SAME bytes read through a
//   descriptor whose start address is shifted by (dh*cols+dw)*16 B: the halo is loaded once.
// Depth streaming + stacking.  A CTA walks one (n, h-tile, w-tile) column through all planes.
//   stride 1: input plane z feeds output planes z-1,z,z+1 (kd=2,1,0): ONE MMA with
//             N = 96 = [W(kd=2)|W(kd=1)|W(kd=0)] updates three neighbouring accumulators.
//   stride 2: the input is first re-laid into its 8 parity sub-volumes (space_to_depth kernel) so
//             every tap is again a unit-stride tile; odd input planes feed two output planes (N=64).
//   transposed (stride 2): rows are INPUT positions; each input plane feeds output planes
//             2z-1,2z,2z+1; per output plane the 4 in-plane output parity classes are 4 x 32
//             accumulator columns, and taps that share an input shift are stacked (N = 128/64/32).
//   Accumulators live in a TMEM ring (512 columns = 16 slots x 32 or 4 slots x 128).  The epilogue
//   zeroes a slot right after reading it, so every MMA accumulates and the issue loop carries no
//   first-touch logic.  (Measured: the single issuing thread is the scarce resource -- a loop with
//   per-MMA address arithmetic ran at ~300 cycles/MMA; descriptors are now built once per plane and
//   advanced by compile-time constants.)
// Warp roles (256 threads, 1 CTA/SM, persistent over columns): warp 0 = TMA producer, warp 1 = MMA
//   issuer (one thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld 32x32b.x32 ->
//   +bias (+residual) (ReLU) -> bf16 -> 16-byte stores).
// Pipelines: full/empty mbarriers over the input ring (TMA <-> MMA), acc_full/acc_empty over the TMEM
//   ring (MMA <-> epilogue); the layer's weights (55/110 KB) stay resident in shared memory.
// Roofline: tensor (dense bf16/fp16); algorithmic FLOPs = 2*27*Cin*Cout per output voxel (stride 1).
//
// Scheduling forms added on top of that (each documented at its Cfg flag / Params field): per-step accumulators with merged
// x_hi * [w_hi | w_lo] MMAs (Cfg::TRI / MRG, stride 1), per-step pairs (Cfg::S2T, stride 2 32 -> 64) and class-major kd-stacked
// buffers (Cfg::DTR, transposed); Cin = 64 layers split by input channels into two such launches (tc_conv3d_split); work items that
// split the last round's columns -- or, for small batches, all columns -- in depth (Items); CTA pairs that fetch their common input
// once with multicast TMA (Params::cluster); the parity-layout copy of a launch's input written by its two idle warps
// (Params::x_split); residual boxes prefetched to L2 by the MMA warp (Params::res_map); weights loaded by bulk TMA copies.
//
// Storage formats (template parameter FMT): bf16 or IEEE-half words (one word per value), or SPLIT PRECISION -- every activation
// and weight is two IEEE-half words (hi + lo, block groups [hi | lo] per sample), a product is x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// accumulated in fp32: the tensor-core mode that meets the 1e-3 px parity bar (Cfg::XP: the three terms as extra k-steps of ONE
// launch where both weight words fit in shared memory; Cfg::TRI / accumulator banks: tcgen05.mma adds into the fp32 accumulator
// by truncation, so no accumulator column may collect a long chain of full-magnitude adds -- see Cfg and DESIGN.md 4b).
#include "conv3d_tc.cuh"
#include "sm100_ptx.cuh"

#include <cstring>
#include <type_traits>
#include <vector>

namespace idisp {

#ifndef IDISP_TRI
#define IDISP_TRI 1  // per-step accumulator triples in the stride-1 split-precision kernels (see Cfg::TRI); 0: banked plane ring
#endif
#ifndef IDISP_MRG
#define IDISP_MRG 1  // TRI kernels with both weight words resident: x_hi feeds [w_hi | w_lo] in ONE N=192 MMA (see Cfg::MRG)
#endif
#ifndef IDISP_TRI_EG
#define IDISP_TRI_EG 2  // epilogue groups (of 4 warps) of the per-step-triple kernels with 32-wide blocks: 4 x 8 channels or 2 x 16
#endif
#ifndef IDISP_S2T
#define IDISP_S2T 1  // stride-2 32->64 split-precision conv with per-step accumulator pairs (see Cfg::S2T); 0: banked plane ring
#endif
#ifndef IDISP_DTR
#define IDISP_DTR 1  // transposed split-precision conv with per-step accumulators and kd-stacked, class-major MMAs (see Cfg::DTR)
#endif
#ifndef IDISP_NMAIN
#define IDISP_NMAIN 3  // accumulator banks of the main term in the split-precision kernels (see Cfg)
#endif

namespace tc {
enum { M_S1 = 0, M_S2 = 1, M_DEC = 2 };
constexpr int TW = 8, TH = 16;  // row tile (w x h) = 128 GEMM rows

// ACC_BLOCKS = accumulator blocks (of NT columns) one output plane owns
template <int MODE> struct ModeCfg;
template <> struct ModeCfg<M_S1> { static constexpr int SUB_W = TW + 2, SUB_H = TH + 2, SUBS = 1, ACC_BLOCKS = 1; };
// (SUB_H is 18 rather than the 17 rows these modes read so that Cin/8 * SUB_W*SUB_H*16 B stays a multiple of 128 B,
//  the alignment TMA needs for the second sub-tile of a stage)
template <> struct ModeCfg<M_S2> { static constexpr int SUB_W = TW + 1, SUB_H = TH + 2, SUBS = 2, ACC_BLOCKS = 1; };
// (transposed conv: Cin = 64 only, so the 17 rows it reads already make 128 B-multiple sub-tiles; the smaller stage buys the
//  two-word split-precision form a third pipeline stage next to its 110 KB of weights)
template <> struct ModeCfg<M_DEC> { static constexpr int SUB_W = TW + 1, SUB_H = TH + 1, SUBS = 1, ACC_BLOCKS = 4; };

// OCC = CTAs per SM.  OCC 2 (stride-1, Cin 32 only: two 55 KB weight copies fit) gives the tensor pipe a second,
// independent MMA stream that fills the bubbles one issuing warp leaves at plane boundaries (barrier round trips,
// commits, descriptor set-up); each CTA then owns 256 TMEM columns and one epilogue group.
// NT = output channels per stacked block: 32 by default; 16 for the 32->1 heads (their zero-padded kernel then costs
// N=48 instead of N=96 of shared-memory B traffic per MMA); 64 for the stride-2 32->64 conv (both Cout halves in one MMA,
// N=128/64 instead of two CTAs re-reading A with N=64/32).
// XP = split-precision K-concatenation done INSIDE one launch (activations hold hi|lo block groups):
//   1: A = [x_hi | x_lo | x_hi], B = [w_hi | w_hi | w_lo]  (both weight words resident: Cin = 32 layers)
//   2: A = [x_hi | x_lo],        B = [w_hi | w_hi]          (Cin = 64 layers: w_lo does not fit next to w_hi; its
//                                                            x_hi*w_lo term is a second, plain launch chained through a partial)
// XM = split-precision mode of the launch: 0 none, 1 / 2 = XP above, 3 = a plain-K pass of a multi-launch split layer.
// Accumulator banks (XM != 0).  The tensor core adds each MMA's K=16 dot products into the fp32 accumulator by TRUNCATION
// (measured: the disparity error grew ~3x when the two correction terms were accumulated in the same TMEM columns as the
// main term, i.e. with 3x as many full-magnitude adds per output).  So the main term's chain is split over NMAIN
// column banks (by kw) and the small correction terms get a bank of their own, where
// their truncation is 2^-11 smaller; the epilogue sums the banks in fp32 round-to-nearest.
template <int CIN, int MODE, int OCC, int NT, int XM = 0> struct Cfg {
  using MC = ModeCfg<MODE>;
  static constexpr int XP = XM == 3 ? 0 : XM;
  // TRI (stride-1, 32-wide blocks, split precision): accumulators are not per output plane but per STEP.  Input plane z adds its
  // three kd contributions into a fresh, aligned triple of column blocks [plane z-1 | z | z+1] (one N=96 MMA, never split by a
  // ring wrap -- with four banks the 4-slot plane ring split half of them into N=64 + N=32, 88 instead of 56 port cycles); the
  // epilogue drains the triple after every step and carries the two open planes' sums in registers (fp32 round-to-nearest).
  // Each column then sees only 9*KS truncating adds, and the correction terms get a triple of their own.
  static constexpr bool TRI = IDISP_TRI && XM != 0 && MODE == M_S1 && (NT == 32 || NT == 16) && OCC == 1;
  // MRG (TRI, both weight words resident, 32-wide blocks): the two terms that share x_hi are ONE MMA with B = [w_hi | w_lo]
  // (N = 192: 96 tensor cycles, where two N = 96 MMAs cost 2 x 56 cycles of the 128 B/clk shared-memory port -- the A tile
  // is read once instead of twice); its D is the step's main triple followed by its correction triple, and the x_lo * w_hi
  // MMA (N = 96, the hi rows of the same B chunk) adds into the correction triple.  Edge planes compute all three kd blocks:
  // the epilogue never reads the block of a plane outside the volume.
  static constexpr bool MRG = IDISP_MRG && TRI && XM == 1 && NT == 32;
  // S2T (stride 2, Cin 32, both weight words resident, 32-wide blocks): the stride-2 twin of TRI + MRG.  A step = one PAIR of
  // input planes (2p, 2p+1); its accumulators are a fresh pair of blocks [plane p | plane p+1] in three column groups,
  // laid out [C0 | M0 | M1 | C1 | X0 | X1] (M = x_hi*w_hi, C = x_hi*w_lo, X = x_lo*w_hi) so that every MMA's D is contiguous:
  //   even plane (kd 1 -> block 0):  x_hi * [w_lo | w_hi](kd1)                     N =  64 -> [C0 | M0];  x_lo * w_hi(kd1)  N = 32 -> X0
  //   odd plane (kd 2 -> block 0, kd 0 -> block 1): x_hi * [lo2 | hi2 | hi0 | lo0]  N = 128 -> [C0 M0 M1 C1];  x_lo * [hi2 | hi0]  N = 64 -> [X0 X1]
  // The epilogue drains the pair after every step: plane p = carry (block 1 of the previous step) + block 0; carry = block 1.
  // Against the banked plane ring (N = 32 / 64 MMAs, inline barrier waits): 3 600 instead of 4 752 shared-memory-port cycles
  // per output plane, waits of the next stage issued inside the current stage's MMA stream.
  static constexpr bool S2T = IDISP_S2T && IDISP_TRI && XM == 1 && MODE == M_S2 && CIN == 32 && NT == 32 && OCC == 1;
  // DTR (transposed conv, both weight words resident, 16-wide blocks): the transposed twin of TRI.  A step = one INPUT plane z; its
  // accumulators are a fresh 12-block buffer laid out CLASS-major, [c00 | c10 | c11 | c01] (class = (ph, pw) of the output voxel
  // (2h+ph, 2w+pw)), each class holding the three output planes the input plane feeds, [2z-1 (kd 0) | 2z (kd 1) | 2z+1 (kd 2)].
  // In that order the classes an input shift feeds are adjacent, and the three kd are stacked in N:
  //   shift (0,0) -> all four classes, N = 12*NT;  (0,1) -> [c11 | c01], N = 6*NT;  (1,0) -> [c10 | c11], N = 6*NT;  (1,1) -> c11, N = 3*NT
  // i.e. 4 MMAs per k-step (252 shared-memory-port cycles) where the plane-ring form issues 3 x 5 of N = 64 / 32 / 16 (588 cycles):
  // with 16-wide blocks every MMA paid the 4 KB A read for 16-64 columns of work.  The epilogue drains a buffer after every step:
  // plane 2z-1 = carry + block 0, plane 2z = block 1, carry = block 2 (fp32 round-to-nearest, carried in registers).
  static constexpr bool DTR = IDISP_DTR && XM == 1 && MODE == M_DEC && NT == 16 && OCC == 1;
  static constexpr int DTR_STRIDE = 12 * NT;                      // DTR: TMEM columns of one step buffer
  static constexpr int NMAIN = (!TRI && !S2T && XM != 0 && MODE != M_DEC && NT <= 32 && OCC == 1) ? IDISP_NMAIN : 1;
  static constexpr int NB = NMAIN + ((NMAIN > 1 && XP) ? 1 : 0);
  static constexpr int AW = XP ? 2 : 1;        // activation words per stage
  static constexpr int BW = XP == 1 ? 2 : 1;   // weight words resident in shared memory
  static constexpr int ACC_COLS = MC::ACC_BLOCKS * NT;   // TMEM columns of one output plane
  static constexpr int WCHUNK = 2 * 3 * NT * 16;          // B operand of one (kh,kw,kstep): [2 kcores][3 blocks x NT rows][8] bf16
  static_assert(MODE != M_DEC || NT == 32 || NT == 16, "the transposed-conv stacking table scales from 32-wide blocks");
  // epilogue groups of 4 warps.  Plane-ring kernels: the groups take alternate output planes.  Per-step-triple kernels: every
  // group drains every step and owns 32/EGROUPS of the block's channels -- four groups, because the epilogue of a step is a
  // dependent instruction chain per warp (TMEM loads -> sums -> pack -> stores).  Measured per 32->32 layer: 2 groups x 16
  // channels 1.87 ms, 4 groups x 8 channels 2.07 ms (MMA stream alone: 1.50 ms) -- more epilogue warps take issue slots from
  // the MMA warp's scheduler, so two groups it is
  static constexpr int EGROUPS = OCC == 2 ? 1 : (((TRI && NT == 32) || S2T) ? IDISP_TRI_EG : 2);
  static constexpr int NTHREADS = 128 + 128 * EGROUPS;    // warps 0-3: TMA producer / MMA issuer / TMEM allocator / idle
  static constexpr int TCOLS = 512 / OCC;                 // TMEM columns of this CTA
  static constexpr int KS = CIN / 16;    // K=16 MMAs per tap
  static constexpr int CBLK = CIN / 8;   // channel blocks
  // (S2T: 17 rows instead of the 18 ModeCfg pads to -- with two activation words the sub-tile stays a multiple of 128 B -- which buys
  //  the third pipeline stage next to the 110 KB of weights: with two stages the layer was bound by TMA latency)
  static constexpr int SUB_H = S2T ? TH + 1 : MC::SUB_H;
  static constexpr int PLANE_BYTES = MC::SUB_W * SUB_H * 16;  // one channel block of one sub-tile (LBO of A)
  static constexpr int ROW_BYTES = MC::SUB_W * 16;                // one tile row (SBO of A)
  static constexpr int SUB_BYTES = AW * CBLK * PLANE_BYTES;
  static constexpr int STAGE_BYTES = MC::SUBS * SUB_BYTES;
  static constexpr int KSW = BW * KS;                              // weight k-steps per tap
  static constexpr int KSM = XP == 1 ? 3 * KS : (XP == 2 ? 2 * KS : KS);  // MMAs per tap
  static constexpr int WBYTES = 27 * KSW * NT * 32;               // 27 taps x Cin (x words) x NT couts x 16 bit
  static constexpr int NSLOT = (TRI || S2T || DTR) ? 2 : TCOLS / (ACC_COLS * NB);   // TRI: two step-triples (MMA fills one while the other drains)
  static constexpr int TRI_STRIDE = MRG ? 6 * NT : 3 * NT;        // TRI: TMEM columns between the two step buffers
  static constexpr int TRI_SMALL = MRG ? 3 * NT : 2 * 3 * NT;     // TRI: column offset of a step's correction triple from its main triple
  static constexpr int BANK_COLS = NSLOT * ACC_COLS;              // TMEM column distance between accumulator banks
  static constexpr int S2T_STRIDE = 6 * NT;                       // S2T: TMEM columns of one step buffer
  static_assert(TRI || S2T || DTR || NSLOT >= 4, "the accumulator ring needs four slots");
  static constexpr int STAGES_FIT = (228 * 1024 / OCC - 1024 - WBYTES - 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int BAR_OFF = WBYTES + STAGES * STAGE_BYTES;
  static constexpr int SMEM = BAR_OFF + (2 * STAGES + 2 * NSLOT) * 8 + 16 + NT * 4;
  static_assert(STAGES >= 2, "input ring needs at least two stages");
  static_assert(STAGE_BYTES % 128 == 0 && WBYTES % 128 == 0 && SUB_BYTES % 128 == 0, "TMA destinations must stay 128 B aligned");
};

struct Params {
  const __nv_bfloat16 *w;         // packed per mode, [nh][...] (see tc_weights_prepare)
  const float *bias;              // [Cout] or nullptr
  const __nv_bfloat16 *residual;  // blocked, output shape, or nullptr
  __nv_bfloat16 *y;               // blocked [B][Cout/8][Do][Ho][Wo][8]
  const float *res1;              // 32->1 head: running sum [B][D][H][W] f32 or nullptr
  float *y1;                      // 32->1 head: output [B][D][H][W] f32 (non-null selects this epilogue)
  int y1_cols;                    // 32->1 head: accumulator columns summed into the logit (2: column 1 = the w_lo products)
  __nv_bfloat16 *y_split;         // optional second copy of y in the 8-parity-sub-volume layout a stride-2 consumer reads
  int residual_is_split;          // transposed conv only: `residual` is stored in that parity layout (of the OUTPUT grid)
  int skip_y;                     // write only y_split (the natural copy has no reader)
  __nv_bfloat16 *x_split;         // per-step-triple 32 -> 32 kernel: warps 2 and 3 (otherwise idle) copy every input plane tile from its shared-memory
                                  // stage into the parity layout here (the stage is released by the MMA commit AND these two warps' arrivals).
                                  // Why: the transposed conv that produces this tensor is memory-bound and wrote both layouts (natural for
                                  // this launch, parity for the next hourglass's stride-2 conv); this launch is MMA-bound with DRAM at 30 %.
  int cluster;                    // Cfg::S2T: CTAs (2k, 2k+1) -- the two output-channel slices of one column walk -- are a cluster of two and fetch
                                  // each input stage ONCE between them: CTA r loads sub-tile pw = r with a multicast TMA that lands in both
                                  // CTAs' shared memory; a stage is free when BOTH CTAs' MMAs have released it (multicast commit)
  int res_map;                    // Cfg::DTR: the second tensor map covers `residual` (boxes of one output plane tile x 2 channel blocks)
  // split-precision ("x2") passes: a product of (hi+lo) operands is three launches whose accumulators are chained through an
  // fp32 partial; the last pass applies the epilogue and stores the result as two 16-bit words (hi blocks, then lo blocks)
  const float *part_in;           // fp32 partial sums of the earlier pass(es), blocked [B][Cout/8][V][8], or nullptr
  float *part_out;                // non-null: store this pass's accumulator (+ part_in) there and do nothing else
  int x2;                         // final pass: y / residual / y_split carry 2*Cout/8 blocks per sample (hi | lo)
  int *range_flag;                // optional device int: set to 1 when a value to be stored as IEEE half exceeds its range
  int in_blk_stride, in_blk_off;  // channel blocks per input sample in memory, and which block this launch's channels start at
  int in_lo_off;                  // != 0: the lo group of this launch's channels starts that many blocks after its hi group (two TMA boxes per stage)
  int B, Din, Dout, Ho, Wo, Hr, Wr, Cout, relu;  // (Hr,Wr): row grid the 8x16 tiles cover
  int tiles_h, tiles_w, nh;
  int cv_shift0;                  // mindisp/4: plane k <-> right-view shift i = k + cv_shift0
  int cv_view;                    // CV == 2: which view this launch reads (0 left, 1 right)
  int dbg;                        // timing experiments only (IDISP_TC_DBG): 1 no MMAs, 2 no TMA loads, 4 no global stores, 8 no tcgen05.ld, 16 no tcgen05.st, 32 TRI epilogue = handshake only, 64 TRI epilogue without the per-voxel work, 128 TRI epilogue without TMEM reads, 256 DTR without the residual L2 prefetch, 512 DTR without residual loads, 1024 DTR plain instead of streaming stores, 2048 no depth-split work items, 4096 DTR prefetch one step further ahead
};

// DECONV stacking table: per kd, five MMAs (entries) that share an input shift
//   e0 shift(0,0) -> classes 0..3 (N=128), e1 shift(0,1) -> classes 2,3 (N=64),
//   e2 shift(1,0) -> class 1, e3 shift(1,0) -> class 3, e4 shift(1,1) -> class 3   (class = pw*2+ph)
__host__ __device__ constexpr int dec_rows(int e) { return e == 0 ? 128 : (e == 1 ? 64 : 32); }
__host__ __device__ constexpr int dec_row_off(int e) { return e == 0 ? 0 : (e == 1 ? 128 : (e == 2 ? 192 : (e == 3 ? 224 : 256))); }
__host__ __device__ constexpr int dec_dcol(int e) { return e == 0 ? 0 : (e == 1 ? 64 : (e == 2 ? 32 : 96)); }
__host__ __device__ constexpr int dec_shift_h(int e) { return e >= 2 ? 1 : 0; }
__host__ __device__ constexpr int dec_shift_w(int e) { return (e == 1 || e == 4) ? 1 : 0; }

// Per-plane tensor maps of the fused cost volume travel as a __grid_constant__ kernel parameter (the documented way to
// hand TMA a descriptor; CUDA >= 12.1 allows 32 KB of parameters).  Non-cost-volume launches pass an empty struct.
constexpr int CV_MAX_PLANES = 64;
template <bool CV> struct CvMaps { CUtensorMap m[CV_MAX_PLANES]; };
template <> struct CvMaps<false> { char unused; };

__device__ __forceinline__ F8 unpack8(const uint4 &a)
{
  F8 r;
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}

// Work items of the per-step-triple kernels.  A CTA takes items cta, cta + ncta, ...  The rounds every CTA takes part in are whole
// columns; the columns of the last, PARTIAL round (3 136 full-resolution columns over 148 CTAs = 21.2 rounds: 28 CTAs would walk
// 48 planes while 120 idle) are split in depth into `split` chunks so that all CTAs share them.  A chunk of output planes [e0, e1)
// streams the input planes [e0-1, e1+1) -- one extra step per cut -- and emits only its own planes.
struct Items { int nfull, split, total; };
__device__ __forceinline__ Items make_items(int ncols, int ncta, int D, bool enable)
{
  Items it;
  it.nfull = enable ? (ncols / ncta) * ncta : ncols;
  const int tail = ncols - it.nfull;
  int sp = tail > 0 ? ncta / tail : 1;
  if (sp > D / 6) sp = D / 6;   // chunks of at least six planes (two extra steps per chunk)
  if (sp < 1) sp = 1;
  it.split = sp;
  it.total = it.nfull + tail * sp;
  return it;
}
__device__ __forceinline__ void get_item(const Items &it, int i, int D, int &col, int &e0, int &e1)
{
  if (i < it.nfull) { col = i; e0 = 0; e1 = D; return; }
  const int j = i - it.nfull, c = j % it.split;
  col = it.nfull + j / it.split;
  e0 = c * D / it.split;
  e1 = (c + 1) * D / it.split;
}

// K-concatenation modes (Cfg::XP): MMA k-step kk of a tap -> activation word / weight k-step / byte offset (inside a
// sub-tile) of the first of its two channel blocks.  A stage holds [hi blocks | lo blocks]; with the fused cost volume the
// TMA box lands [left: hi, lo][right: hi, lo], so the logical block order (left, right) is permuted.
template <int XP, int KS> __host__ __device__ constexpr int xp_a_word(int kk) { return XP == 0 ? 0 : (XP == 1 ? ((kk >= KS && kk < 2 * KS) ? 1 : 0) : kk / KS); }
template <int XP, int KS> __host__ __device__ constexpr int xp_b_step(int kk) { return XP == 0 ? kk : (XP == 1 ? (kk < KS ? kk : kk - KS) : kk % KS); }
template <int XP, int KS, int CV, int PLANE> __host__ __device__ constexpr uint32_t xp_a_off(int kk)
{
  const int word = xp_a_word<XP, KS>(kk), blk = 2 * (kk % KS), cblk = 2 * KS, h = cblk / 2;
  const int phys = (CV == 1 && XP) ? ((blk < h ? blk : blk + h) + word * h) : word * cblk + blk;
  return (uint32_t)(phys * PLANE);
}

// operands a split-precision epilogue adds to one voxel block: fp32 partial of the earlier pass, residual hi / lo words
struct XPre { float4 p0, p1; uint4 rh, rl; };

__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t byte_off) { return d + (uint64_t)(byte_off >> 4); }

// FMT: 0 bf16, 1 IEEE half, 2 IEEE half split-precision pass (plain K), 3 / 4 the same with in-launch K concatenation XP = 1 / 2
// CV: 0 plain input tensor; 1 fused cost volume, both views (Cin = 2C); 2 fused cost volume, ONE view (Cin = C: the box lands
//     that view's [hi blocks | lo blocks], i.e. the plain stage layout -- Params::cv_view selects left / right)
template <int CIN, int MODE, int OCC, int CV, int NT, int FMT>
__global__ void __launch_bounds__((Cfg<CIN, MODE, OCC, NT, (FMT < 2 ? 0 : (FMT == 2 ? 3 : FMT - 2))>::NTHREADS), OCC)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap rmap,
                 const __grid_constant__ CvMaps<(CV != 0)> lmaps, const Params p)
{
  constexpr int XM = FMT < 2 ? 0 : (FMT == 2 ? 3 : FMT - 2);
  using C = Cfg<CIN, MODE, OCC, NT, XM>;
  constexpr int XP = C::XP;
  constexpr bool F16 = FMT != 0, X2 = FMT >= 2;
#define A_KOFF(kk) (xp_a_off<XP, C::KS, CV, C::PLANE_BYTES>(kk))
#define B_KS(kk) (xp_b_step<XP, C::KS>(kk))
// TMEM column offset of the accumulator bank the MMA of (kw, k-step) adds into
#define BANK(kw, kk) ((uint32_t)((C::NB > C::NMAIN && (kk) >= C::KS) ? C::NMAIN * C::BANK_COLS : ((kw) % C::NMAIN) * C::BANK_COLS))
  constexpr int NTHREADS = C::NTHREADS;
  using MC = ModeCfg<MODE>;
  constexpr int NSLOT = C::NSLOT;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t w_addr = smem_base;
  const uint32_t stage_addr0 = smem_base + C::WBYTES;
  const uint32_t bar0 = smem_base + C::BAR_OFF;
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (C::STAGES + s); };
  auto accf_bar = [&](uint32_t r) { return bar0 + 8u * (2 * C::STAGES + r); };
  auto acce_bar = [&](uint32_t r) { return bar0 + 8u * (2 * C::STAGES + NSLOT + r); };
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + C::BAR_OFF + (2 * C::STAGES + 2 * NSLOT) * 8);
  float *bias_s = reinterpret_cast<float *>(smem + C::BAR_OFF + (2 * C::STAGES + 2 * NSLOT) * 8 + 16);

  // warp index made PROVABLY warp-uniform (shfl from lane 0) so the role branches are uniform branches and the
  // producer / MMA warps can keep their addresses and descriptors in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();       // (the next kernel's CTAs may take this SM as soon as this CTA is gone: their prologue overlaps our tail)
  const int nh = blockIdx.x % p.nh;  // this CTA's fixed 32-wide output-channel slice
  const int cta = blockIdx.x / p.nh, ncta = gridDim.x / p.nh;
  const int ncols = p.B * p.tiles_h * p.tiles_w;
  const int Din = p.Din, Dout = p.Dout;

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&xmap);
    ptx::prefetch_tensormap(&rmap);
    for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), (C::MRG && p.x_split) ? 3 : ((C::S2T && p.cluster) ? 2 : 1)); }
    for (int r = 0; r < NSLOT; ++r) { ptx::mbar_init(accf_bar(r), 1); ptx::mbar_init(acce_bar(r), (C::TRI || C::S2T || C::DTR) ? 4 * C::EGROUPS : 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<C::TCOLS>(ptx::smem_u32(tmem_ptr_smem));
  // This slice's weights (28-110 KB) -> shared memory with bulk TMA copies tracked by their own mbarrier; only the MMA warp waits for
  // it, right before its first MMA, so the copy overlaps the rest of the prologue (TMEM allocation, accumulator clearing, the first
  // input stages).  (A loop of 16-byte loads and stores by all threads took ~10 us per launch: 18 dependent round trips to L2 -- a
  // quarter of a launch at the live shape, where a forward is ~50 launches of 20-30 us.)
  const uint32_t wbar = bar0 + 8u * (2 * C::STAGES + 2 * NSLOT) + 8u;   // (second half of the 16-byte slot that holds the TMEM pointer)
  if (warp == 0 && lane == 0) {
    ptx::mbar_init(wbar, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(wbar, C::WBYTES);
    const char *src = reinterpret_cast<const char *>(p.w) + (size_t)nh * C::WBYTES;
    for (uint32_t off = 0; off < (uint32_t)C::WBYTES; off += 16384u)
      ptx::bulk_g2s(w_addr + off, src + off, (uint32_t)C::WBYTES - off < 16384u ? (uint32_t)C::WBYTES - off : 16384u, wbar);
  }
  if (threadIdx.x < NT) bias_s[threadIdx.x] = (p.bias && nh * NT + (int)threadIdx.x < p.Cout) ? p.bias[nh * NT + threadIdx.x] : 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp >= 4 && warp < 8) {  // accumulators start at zero; afterwards the epilogue re-zeroes each slot it drains
    uint32_t zero[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) zero[i] = 0u;
    for (int c = 0; c < C::TCOLS; c += 32) ptx::tmem_st_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + c, zero);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const bool paired = C::S2T && p.cluster != 0;
  const uint32_t crank = paired ? ptx::cluster_ctarank() : 0u;
  if (paired) ptx::cluster_sync();   // the peer's barriers are initialised before anything of ours can reach them
  // Programmatic dependent launch: everything above (barriers, TMEM, weight copies, bias) touches nothing the previous kernel in the
  // stream produces, so it ran while that kernel drained; from here on its activations are read.
  ptx::pdl_wait();

  if (warp == 0) {
    // ================= TMA producer (whole warp converged; one elected lane issues) =================
    {
      const bool lead = ptx::elect_one();
      uint32_t q = 0;
      // (items count OUTPUT planes: Dout = Din for the stride-1 kernels; the stride-2 form (S2T) walks a chunk [e0, e1) as the steps
      //  [e0-1, e1) -- the extra first step only produces the odd input plane's carry into plane e0 -- i.e. input planes [2(e0-1), 2 e1))
      //  the transposed form (DTR) counts INPUT planes: a chunk [e0, e1) runs the steps [e0-1, e1), the extra first step only producing
      //  the carry into output plane 2*e0-1)
      const int DI = C::DTR ? Din : Dout;
      const Items items = make_items(ncols, ncta, DI, (C::MRG || C::S2T || C::DTR) && !(p.dbg & 2048));
      for (int item = cta; item < items.total; item += ncta) {
        int col, e0, e1;
        get_item(items, item, DI, col, e0, e1);
        const int zb = C::S2T ? 2 * (e0 > 0 ? e0 - 1 : 0) : (((C::MRG || C::DTR) && e0 > 0) ? e0 - 1 : 0);
        const int ze = C::S2T ? 2 * e1 : (C::DTR ? e1 : ((C::MRG && e1 < Din) ? e1 + 1 : Din));
        const int tw = col % p.tiles_w, th = (col / p.tiles_w) % p.tiles_h, n = col / (p.tiles_w * p.tiles_h);
        for (int z = zb; z < ze; ++z) {
          if (MODE == M_S2) {
            // two pipeline steps per input plane: even-row (ph=0) and odd-row (ph=1) sub-grids, each {pw=0, pw=1}
            for (int ph = 0; ph < 2; ++ph, ++q) {
              const uint32_t s = q % C::STAGES;
              ptx::mbar_wait(empty_bar(s), ((q / C::STAGES) & 1) ^ 1);
              if (lead && (p.dbg & 2)) ptx::mbar_arrive(full_bar(s));
              else if (lead && paired) {
                // this CTA fetches sub-tile pw = its cluster rank for both CTAs of the pair; the peer's box completes the stage
                ptx::mbar_arrive_expect_tx(full_bar(s), C::STAGE_BYTES);
                const int pw = (int)crank, cls = (z & 1) * 4 + ph * 2 + pw;
                const bool merged = XP != 0 && p.in_lo_off;
                ptx::tma_load_5d_multicast(stage_addr0 + s * C::STAGE_BYTES + pw * C::SUB_BYTES, &xmap, full_bar(s), (tw * TW - 1) * 8, th * TH - 1,
                                           merged ? cls * (Din >> 1) + (z >> 1) : (z >> 1), merged ? 0 : cls,
                                           merged ? n * 2 : n * p.in_blk_stride + p.in_blk_off, (uint16_t)3);
              } else if (lead) {
                ptx::mbar_arrive_expect_tx(full_bar(s), C::STAGE_BYTES);
                for (int pw = 0; pw < 2; ++pw) {
                  if (XP != 0 && p.in_lo_off)
                    // channel subset of a wider tensor (its lo blocks are not adjacent to its hi blocks; half a sub-tile would not be a
                    // 128-byte-aligned TMA destination): the map merges (class, depth) into one dimension and splits the blocks into
                    // (block of the group, group = sample x word), so ONE box still lands [hi blocks | lo blocks]
                    ptx::tma_load_5d(stage_addr0 + s * C::STAGE_BYTES + pw * C::SUB_BYTES, &xmap, full_bar(s), (tw * TW - 1) * 8, th * TH - 1,
                                     ((z & 1) * 4 + ph * 2 + pw) * (Din >> 1) + (z >> 1), 0, n * 2);
                  else
                    ptx::tma_load_5d(stage_addr0 + s * C::STAGE_BYTES + pw * C::SUB_BYTES, &xmap, full_bar(s), (tw * TW - 1) * 8,
                                     th * TH - 1, z >> 1, (z & 1) * 4 + ph * 2 + pw, n * p.in_blk_stride + p.in_blk_off);
                }
              }
            }
          } else {
            const uint32_t s = q % C::STAGES;
            ptx::mbar_wait(empty_bar(s), ((q / C::STAGES) & 1) ^ 1);
            const int halo = MODE == M_S1 ? 1 : 0;
            if (lead && (p.dbg & 2)) ptx::mbar_arrive(full_bar(s));
            else if (CV) {
              if (lead) {
              // Fused cost volume (stackhourglass.py:115-128).  Plane k holds, for x in [max(i,0), W+min(i,0)), LEFT[x] in
              // channel blocks [0, CBLK/2) and RIGHT[x-i] in [CBLK/2, CBLK), zero elsewhere (and zero outside [0,W): conv
              // padding).  The plane's own 4-D map (origin x = max(i,0), width W-|i|, outermost dim = {left, right} with the
              // right base advanced by max(-i,0)) makes every masked column plain TMA zero fill: ONE load per plane.
              const int i = z + p.cv_shift0;
              const bool dead = i >= p.Wo || -i >= p.Wo;  // fully masked plane: read far out of range (all zero fill)
              ptx::mbar_arrive_expect_tx(full_bar(s), C::STAGE_BYTES);
              if constexpr (CV != 0)
                ptx::tma_load_4d(stage_addr0 + s * C::STAGE_BYTES, &lmaps.m[z], full_bar(s), dead ? -(1 << 20) : (tw * TW - 1 - (i > 0 ? i : 0)) * 8,
                                 th * TH - 1, n * p.in_blk_stride + p.in_blk_off, CV == 2 ? p.cv_view : 0);
              }
            } else if (lead) {
              ptx::mbar_arrive_expect_tx(full_bar(s), C::STAGE_BYTES);
              ptx::tma_load_4d(stage_addr0 + s * C::STAGE_BYTES, &xmap, full_bar(s), (tw * TW - halo) * 8, th * TH - halo, z, n * p.in_blk_stride + p.in_blk_off);
              if (XP != 0 && p.in_lo_off)   // channel subset of a wider tensor: its lo blocks are not adjacent to its hi blocks
                ptx::tma_load_4d(stage_addr0 + s * C::STAGE_BYTES + C::CBLK * C::PLANE_BYTES, &xmap, full_bar(s), (tw * TW - halo) * 8, th * TH - halo, z,
                                 n * p.in_blk_stride + p.in_blk_off + p.in_lo_off);
            }
            ++q;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs this loop converged and ONE elected lane issues: tcgen05.mma takes its descriptors from
    // uniform registers, and inside an `if (lane == 0)` region the compiler cannot prove uniformity -- it wrapped
    // every MMA in an ELECT / 6x R2UR.BROADCAST / BRA.U.ANY waterfall (~190 cycles per MMA, profiles/r01_notes.md).
    {
      const bool lead = ptx::elect_one();
      const bool do_mma = lead && !(p.dbg & 1);
      ptx::mbar_wait(wbar, 0);   // the weights have landed (async proxy -> visible to the MMAs issued after this wait)
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc = 1u) { if (do_mma) ptx::umma_bf16_ss(d, a, b, id, acc); };
      auto commit = [&](uint32_t bar) { if (lead) ptx::umma_commit(bar); };
      const uint64_t a_desc0 = ptx::make_smem_desc(stage_addr0, C::PLANE_BYTES, C::ROW_BYTES);
      auto wait_acc_empty = [&](uint32_t g) { ptx::mbar_wait(acce_bar(g % NSLOT), ((g / NSLOT) & 1) ^ 1); };
      uint32_t q = 0, g0 = 0;
      if constexpr (C::TRI) {
        // one stage and one accumulator triple per step; the waits of step q+1 are issued in the middle of step q's stream
        auto waits = [&](bool more, uint32_t q_) {
          if (!more) return;
          ptx::mbar_wait(acce_bar(q_ % NSLOT), ((q_ / NSLOT) & 1) ^ 1);
          ptx::mbar_wait(full_bar(q_ % C::STAGES), (q_ / C::STAGES) & 1);
        };
        // (col, z) walk the work items (see Items): whole columns, or depth chunks of the last round's columns
        const Items items = make_items(ncols, ncta, Din, C::MRG && !(p.dbg & 2048));
        auto bounds = [&](int item_, int &zb_, int &ze_) {
          int c_, e0_, e1_;
          get_item(items, item_, Din, c_, e0_, e1_);
          zb_ = e0_ > 0 ? e0_ - 1 : 0; ze_ = e1_ < Din ? e1_ + 1 : Din;
        };
        int col = cta, z = 0, zend = Din;   // `col` counts items here
        if (col < items.total) bounds(col, z, zend);
        waits(col < items.total, 0);
        while (col < items.total) {
          ptx::tc_fence_after();
          const uint32_t s = q % C::STAGES, t = q % NSLOT;
          const int plo = z > 0 ? z - 1 : 0, phi = z + 1 < Dout ? z + 1 : Dout - 1;  // output planes this input plane feeds
          const int j0 = C::MRG ? 0 : plo - (z - 1), nblk = C::MRG ? 3 : phi - plo + 1;   // first block / blocks of [kd=2|kd=1|kd=0]
          const uint32_t dm = tmem_base + t * C::TRI_STRIDE + j0 * NT, ds = dm + C::TRI_SMALL;
          const uint32_t id1 = ptx::make_idesc_h<F16>(128, NT * nblk);
          const uint64_t a0 = desc_add(a_desc0, s * C::STAGE_BYTES);
          int ncol = col, nz = z + 1, nzend = zend;
          if (nz == zend) { ncol += ncta; if (ncol < items.total) bounds(ncol, nz, nzend); }
          if constexpr (C::MRG) {
            // B chunk of (tap, k-step): [2 kcores][hi: 3 x NT rows | lo: 3 x NT rows][8]
            const uint32_t id2 = ptx::make_idesc_h<F16>(128, 2 * 3 * NT);
            const uint64_t bm = ptx::make_smem_desc(w_addr, 6 * NT * 16, 128);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              if (tap == 5) waits(ncol < items.total, q + 1);
#pragma unroll
              for (int ks = 0; ks < C::KS; ++ks) {
                const uint32_t a_tap = ((tap / 3) * MC::SUB_W + (tap % 3)) * 16;
                const uint32_t boff = (tap * C::KS + ks) * 2 * C::WCHUNK;
                // (the step's triples are fresh: the very first MMA overwrites them, nothing is zeroed on drain)
                mma(dm, desc_add(a0, a_tap + A_KOFF(ks)), desc_add(bm, boff), id2, (tap | ks) ? 1u : 0u);   // x_hi * [w_hi | w_lo] -> [main | corr]
                mma(ds, desc_add(a0, a_tap + A_KOFF(C::KS + ks)), desc_add(bm, boff), id1);                 // x_lo * w_hi -> corr
              }
            }
          } else {
          const uint64_t b1 = ptx::make_smem_desc(w_addr + j0 * NT * 16, 3 * NT * 16, 128);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (tap == 5) waits(ncol < items.total, q + 1);
#pragma unroll
            for (int ks = 0; ks < C::KSM; ++ks) {
              const uint32_t aoff = ((tap / 3) * MC::SUB_W + (tap % 3)) * 16 + A_KOFF(ks);
              // first MMA into the main triple (tap 0, k-step 0) and into the correction triple (tap 0, k-step KS) overwrite
              mma((XP != 0 && ks >= C::KS) ? ds : dm, desc_add(a0, aoff), desc_add(b1, (tap * C::KSW + B_KS(ks)) * C::WCHUNK), id1,
                  (tap == 0 && (ks == 0 || ks == C::KS)) ? 0u : 1u);
            }
          }
          }
          commit(empty_bar(s));
          commit(accf_bar(t));
          col = ncol; z = nz; zend = nzend; ++q;
        }
      } else if (MODE == M_S1) {
        // Software-pipelined: the mbarrier waits of step q+1 (TMA data landed, fresh accumulator slot drained) are
        // issued in the MIDDLE of step q's MMA stream, so their ~100-cycle round trips hide behind queued MMAs.
        auto waits = [&](int col_, int z_, uint32_t q_, uint32_t g0_) {
          if (col_ >= ncols) return;
          if (z_ == 0) { wait_acc_empty(g0_); if (Dout > 1) wait_acc_empty(g0_ + 1); }
          else if (z_ + 1 < Dout) wait_acc_empty(g0_ + z_ + 1);
          ptx::mbar_wait(full_bar(q_ % C::STAGES), (q_ / C::STAGES) & 1);
        };
        int col = cta, z = 0;
        waits(col, 0, 0, 0);
        while (col < ncols) {
          ptx::tc_fence_after();
          const uint32_t s = q % C::STAGES;
          // output planes [plo, phi] <- B blocks [plo-(z-1), ...] (block j <-> kd = 2-j); split where the ring wraps
          const int plo = z > 0 ? z - 1 : 0, phi = z + 1 < Dout ? z + 1 : Dout - 1;
          const uint32_t slot0 = (g0 + plo) % NSLOT;
          const int nblk = phi - plo + 1;
          const int len1 = nblk < (int)(NSLOT - slot0) ? nblk : (int)(NSLOT - slot0), len2 = nblk - len1;
          const uint32_t d1 = tmem_base + slot0 * NT, d2 = tmem_base;
          const uint32_t id1 = ptx::make_idesc_h<F16>(128, NT * len1), id2 = ptx::make_idesc_h<F16>(128, NT * (len2 > 0 ? len2 : 1));
          const uint64_t a0 = desc_add(a_desc0, s * C::STAGE_BYTES);
          const uint64_t b1 = ptx::make_smem_desc(w_addr + (plo - (z - 1)) * NT * 16, 3 * NT * 16, 128);
          const uint64_t b2 = desc_add(b1, len1 * NT * 16);
          int ncol = col, nz = z + 1;
          uint32_t ng0 = g0;
          if (nz == Din) { nz = 0; ncol += ncta; ng0 += Dout; }
          // (two separate streams: a predicated-off second MMA still costs ~28 issue cycles, measured)
          if (len2 == 0) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              if (tap == 5) waits(ncol, nz, q + 1, ng0);
#pragma unroll
              for (int ks = 0; ks < C::KSM; ++ks) {
                const uint32_t aoff = ((tap / 3) * MC::SUB_W + (tap % 3)) * 16 + A_KOFF(ks);
                mma(d1 + BANK(tap % 3, ks), desc_add(a0, aoff), desc_add(b1, (tap * C::KSW + B_KS(ks)) * C::WCHUNK), id1);
              }
            }
          } else {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              if (tap == 5) waits(ncol, nz, q + 1, ng0);
#pragma unroll
              for (int ks = 0; ks < C::KSM; ++ks) {
                const uint32_t aoff = ((tap / 3) * MC::SUB_W + (tap % 3)) * 16 + A_KOFF(ks);
                const uint32_t boff = (tap * C::KSW + B_KS(ks)) * C::WCHUNK;
                mma(d1 + BANK(tap % 3, ks), desc_add(a0, aoff), desc_add(b1, boff), id1);
                mma(d2 + BANK(tap % 3, ks), desc_add(a0, aoff), desc_add(b2, boff), id2);
              }
            }
          }
          commit(empty_bar(s));
          if (z >= 1) commit(accf_bar((g0 + z - 1) % NSLOT));
          if (z == Din - 1) commit(accf_bar((g0 + z) % NSLOT));
          col = ncol; z = nz; g0 = ng0; ++q;
        }
      }
      if constexpr (C::S2T) {
        // B chunk of (kh, kw, k-step): [2 kcores][192 rows][8]: rows [lo kd1 | hi kd1 | lo kd2 | hi kd2 | hi kd0 | lo kd0]
        const uint32_t id_e = ptx::make_idesc_h<F16>(128, 2 * NT), id_o = ptx::make_idesc_h<F16>(128, 4 * NT), id_x = ptx::make_idesc_h<F16>(128, NT);
        const uint64_t b0 = ptx::make_smem_desc(w_addr, 6 * NT * 16, 128);
        uint32_t sq = 0;  // stage counter (two stages per input plane), tq: step counter
        uint32_t tq2 = 0;
        bool first_wait = true;
        const Items items = make_items(ncols, ncta, Dout, !(p.dbg & 2048));
        for (int item = cta; item < items.total; item += ncta) {
          int col_, e0, e1;
          get_item(items, item, Dout, col_, e0, e1);
          for (int pz = (e0 > 0 ? e0 - 1 : 0); pz < e1; ++pz, ++tq2) {
            const uint32_t t = tq2 % NSLOT;
            const uint32_t dbase = tmem_base + t * C::S2T_STRIDE;   // [C0 | M0 | M1 | C1 | X0 | X1]
            ptx::mbar_wait(acce_bar(t), ((tq2 / NSLOT) & 1) ^ 1);
#pragma unroll
            for (int odd = 0; odd < 2; ++odd) {
#pragma unroll
              for (int ph = 0; ph < 2; ++ph, ++sq) {
                const uint32_t s = sq % C::STAGES;
                if (first_wait) { ptx::mbar_wait(full_bar(s), (sq / C::STAGES) & 1); first_wait = false; }
                ptx::tc_fence_after();
                const uint64_t a0 = desc_add(a_desc0, s * C::STAGE_BYTES);
                constexpr int NTAPS_MAX = 6;
#pragma unroll
                for (int tt = 0; tt < NTAPS_MAX; ++tt) {
                  if (ph == 0 && tt >= 3) break;
                  // ph=0 (even input rows): kh=1, kw=tt.  ph=1 (odd rows): kh = 0 (tt<3) or 2 (tt>=3), kw = tt%3
                  const int kh = ph == 0 ? 1 : (tt < 3 ? 0 : 2), kw = tt % 3;
                  const int rh = kh == 0 ? 0 : 1, rw = kw == 0 ? 0 : 1, sub_t = kw != 1 ? 1 : 0;
                  const uint32_t aoff0 = sub_t * C::SUB_BYTES + (rh * MC::SUB_W + rw) * 16;
                  if (tt == (ph == 0 ? 1 : 3)) {   // the next stage's TMA data: waited for in the middle of this stage's stream
                    const uint32_t ns = (sq + 1) % C::STAGES;
                    const bool more = !(odd == 1 && ph == 1 && pz == e1 - 1 && item + ncta >= items.total);
                    if (more) ptx::mbar_wait(full_bar(ns), ((sq + 1) / C::STAGES) & 1);
                  }
#pragma unroll
                  for (int ks = 0; ks < C::KS; ++ks) {
                    const uint32_t boff = ((kh * 3 + kw) * C::KS + ks) * 2 * C::WCHUNK;
                    const uint64_t a_hi = desc_add(a0, aoff0 + A_KOFF(ks)), a_lo = desc_add(a0, aoff0 + A_KOFF(C::KS + ks));
                    const bool first = ph == 0 && tt == 0 && ks == 0;   // first MMA of this input plane
                    if (odd == 0) {
                      mma(dbase, a_hi, desc_add(b0, boff), id_e, first ? 0u : 1u);                       // -> [C0 | M0]
                      mma(dbase + 4 * NT, a_lo, desc_add(b0, boff + NT * 16), id_x, first ? 0u : 1u);     // hi kd1 -> X0
                    } else if (first) {
                      // block 0 already holds the even plane's sums, block 1 is fresh: the first MMAs of the odd plane are split
                      mma(dbase, a_hi, desc_add(b0, boff + 2 * NT * 16), id_e, 1u);                       // [lo2 | hi2] -> [C0 | M0]
                      mma(dbase + 2 * NT, a_hi, desc_add(b0, boff + 4 * NT * 16), id_e, 0u);              // [hi0 | lo0] -> [M1 | C1]
                      mma(dbase + 4 * NT, a_lo, desc_add(b0, boff + 3 * NT * 16), id_x, 1u);              // hi2 -> X0
                      mma(dbase + 5 * NT, a_lo, desc_add(b0, boff + 4 * NT * 16), id_x, 0u);              // hi0 -> X1
                    } else {
                      mma(dbase, a_hi, desc_add(b0, boff + 2 * NT * 16), id_o);                           // -> [C0 | M0 | M1 | C1]
                      mma(dbase + 4 * NT, a_lo, desc_add(b0, boff + 3 * NT * 16), id_e);                  // [hi2 | hi0] -> [X0 | X1]
                    }
                  }
                }
                if (paired) { if (lead) ptx::umma_commit_multicast(empty_bar(s), (uint16_t)3); }   // frees the stage in BOTH CTAs
                else commit(empty_bar(s));
              }
            }
            commit(accf_bar(t));
          }
        }
      }
      if constexpr (C::DTR) {
        // weights: [k-step][2 kcores][27*NT rows][8]; rows = shift (0,0): 4 classes x 3 kd x NT | (0,1): 2 x 3 x NT | (1,0): 2 x 3 x NT | (1,1): 3 x NT
        constexpr uint32_t KCH = 2 * 27 * NT * 16;   // bytes of one weight k-step
        const uint64_t b0 = ptx::make_smem_desc(w_addr, 27 * NT * 16, 128);
        const uint32_t id12 = ptx::make_idesc_h<F16>(128, 12 * NT), id6 = ptx::make_idesc_h<F16>(128, 6 * NT), id3 = ptx::make_idesc_h<F16>(128, 3 * NT);
        auto waits = [&](bool more, uint32_t q_) {
          if (!more) return;
          ptx::mbar_wait(acce_bar(q_ % NSLOT), ((q_ / NSLOT) & 1) ^ 1);
          ptx::mbar_wait(full_bar(q_ % C::STAGES), (q_ / C::STAGES) & 1);
        };
        // work items (see Items), here in INPUT planes: a chunk [e0, e1) runs the steps [e0-1, e1)
        const Items items = make_items(ncols, ncta, Din, !(p.dbg & 2048));
        int it = cta, col = 0, z = 0, zend = Din;
        auto bounds = [&](int item_, int &col_, int &zb_, int &ze_) {
          int e0_, e1_;
          get_item(items, item_, Din, col_, e0_, e1_);
          zb_ = e0_ > 0 ? e0_ - 1 : 0; ze_ = e1_;
        };
        if (it < items.total) bounds(it, col, z, zend);
        waits(it < items.total, 0);
        while (it < items.total) {
          ptx::tc_fence_after();
          const uint32_t s = q % C::STAGES, t = q % NSLOT;
          const uint32_t d = tmem_base + t * C::DTR_STRIDE;   // [c00 | c10 | c11 | c01] x [2z-1 | 2z | 2z+1] x NT
          const uint64_t a0 = desc_add(a_desc0, s * C::STAGE_BYTES);
          int nit = it, ncol = col, nz = z + 1, nzend = zend;
          if (nz == zend) { nit += ncta; if (nit < items.total) bounds(nit, ncol, nz, nzend); }
          // The epilogue adds a residual tensor of the output's size; its loads are the layer's critical path.  Ask L2 for the
          // residual boxes of output planes 2z and 2z+1 now -- one tiled TMA prefetch per plane and precision word (16 KB each) --
          // one MMA step (plus the epilogue's lag) before the epilogue reads them.  Measured: issued from the producer warp (2-4
          // steps = 12-25 us earlier) the lines were evicted again before their use (the layer streams ~5 TB/s through the 126 MB
          // L2: DRAM reads rose from 3.5 to 5.8 GB, no speed-up); per-row bulk prefetches (512 per step) made the layer 2x slower.
          if (p.res_map && lead && !(p.dbg & 256)) {
            const int tw = col % p.tiles_w, th = (col / p.tiles_w) % p.tiles_h, n = col / (p.tiles_w * p.tiles_h);
            const int cblk_out = p.Cout / 8, out_blocks = p.x2 ? 2 * cblk_out : cblk_out;
            const int ahead = (p.dbg & 4096) ? 1 : 0;   // (timing experiment: one more step of lead -- measured WORSE, conv6 1.30 -> 1.51 ms)
            for (int zz = (z == 0 ? 0 : z + ahead); zz <= z + ahead && zz < Din; ++zz)
              for (int i = 0; i < (p.x2 ? 4 : 2); ++i) {
                const int qo = 2 * zz + (i & 1), blk = n * out_blocks + nh * 2 + (i >> 1) * cblk_out;
                if (p.residual_is_split) ptx::tma_prefetch_5d(&rmap, tw * TW * 8, th * TH, qo >> 1, (qo & 1) * 4, blk);
                else ptx::tma_prefetch_4d(&rmap, 2 * tw * TW * 8, 2 * th * TH, qo, blk);
              }
          }
          __syncwarp();
#pragma unroll
          for (int ks = 0; ks < C::KSM; ++ks) {
            if (ks == C::KSM / 2) waits(nit < items.total, q + 1);
            const uint32_t ak = A_KOFF(ks), bk = B_KS(ks) * KCH;
            // (the step's buffer is fresh: the first MMA, which covers all of it, overwrites)
            mma(d, desc_add(a0, ak), desc_add(b0, bk), id12, ks ? 1u : 0u);                                                  // shift (0,0)
            mma(d + 6 * NT, desc_add(a0, ak + 16), desc_add(b0, bk + 12 * NT * 16), id6);                                    // shift (0,1) -> [c11 | c01]
            mma(d + 3 * NT, desc_add(a0, ak + MC::SUB_W * 16), desc_add(b0, bk + 18 * NT * 16), id6);                         // shift (1,0) -> [c10 | c11]
            mma(d + 6 * NT, desc_add(a0, ak + (MC::SUB_W + 1) * 16), desc_add(b0, bk + 24 * NT * 16), id3);                   // shift (1,1) -> c11
          }
          commit(empty_bar(s));
          commit(accf_bar(t));
          it = nit; col = ncol; z = nz; zend = nzend; ++q;
        }
      }
      for (int col = cta; MODE != M_S1 && !C::TRI && !C::S2T && !C::DTR && col < ncols; col += ncta, g0 += Dout) {
        for (int z = 0; z < Din; ++z) {
          if (MODE == M_S1) {
          } else if (MODE == M_S2) {
            const int pz = z >> 1;  // output plane of kd=1 (even z) / kd=2 (odd z); odd z also feeds pz+1 with kd=0
            const bool odd = z & 1;
            for (int ph = 0; ph < 2; ++ph, ++q) {
              if (ph == 0) {
                if (z == 0) wait_acc_empty(g0);
                else if (odd && pz + 1 < Dout) wait_acc_empty(g0 + pz + 1);
              }
              const uint32_t s = q % C::STAGES;
              ptx::mbar_wait(full_bar(s), (q / C::STAGES) & 1);
              ptx::tc_fence_after();
              // B chunk rows: [kd=2 | kd=0 | kd=1]; odd z -> rows 0.. (planes pz, pz+1), even z -> rows 64.. (plane pz)
              const uint32_t slot0 = (g0 + pz) % NSLOT;
              const int nblk = odd ? (pz + 1 < Dout ? 2 : 1) : 1;
              const int len1 = nblk < (int)(NSLOT - slot0) ? nblk : (int)(NSLOT - slot0), len2 = nblk - len1;
              const uint32_t d1 = tmem_base + slot0 * NT, d2 = tmem_base;
              const uint32_t id1 = ptx::make_idesc_h<F16>(128, NT * len1), id2 = ptx::make_idesc_h<F16>(128, NT);
              const uint64_t a0 = desc_add(a_desc0, s * C::STAGE_BYTES);
              const uint64_t b1 = ptx::make_smem_desc(w_addr + (odd ? 0 : 2 * NT * 16), 3 * NT * 16, 128);
              const uint64_t b2 = desc_add(b1, len1 * NT * 16);
              auto s2_taps = [&](auto PH) {
                constexpr int ph_ = decltype(PH)::value;
#pragma unroll
                for (int t = 0; t < (ph_ == 0 ? 3 : 6); ++t) {
                  // ph=0 (even input rows): kh=1, kw=t.  ph=1 (odd rows): kh = 0 (t<3) or 2 (t>=3), kw = t%3
                  const int kh = ph_ == 0 ? 1 : (t < 3 ? 0 : 2), kw = t % 3;
                  const int rh = kh == 0 ? 0 : 1, rw = kw == 0 ? 0 : 1, sub = kw != 1 ? 1 : 0;
                  const uint32_t aoff0 = sub * C::SUB_BYTES + (rh * MC::SUB_W + rw) * 16;
                  const uint32_t boff0 = (kh * 3 + kw) * C::KSW * C::WCHUNK;
                  if (len2 == 0) {
#pragma unroll
                    for (int ks = 0; ks < C::KSM; ++ks)
                      mma(d1 + BANK(kw, ks), desc_add(a0, aoff0 + A_KOFF(ks)), desc_add(b1, boff0 + B_KS(ks) * C::WCHUNK), id1);
                  } else {
#pragma unroll
                    for (int ks = 0; ks < C::KSM; ++ks) {
                      mma(d1 + BANK(kw, ks), desc_add(a0, aoff0 + A_KOFF(ks)), desc_add(b1, boff0 + B_KS(ks) * C::WCHUNK), id1);
                      mma(d2 + BANK(kw, ks), desc_add(a0, aoff0 + A_KOFF(ks)), desc_add(b2, boff0 + B_KS(ks) * C::WCHUNK), id2);
                    }
                  }
                }
              };
              if (ph == 0) s2_taps(std::integral_constant<int, 0>{}); else s2_taps(std::integral_constant<int, 1>{});
              commit(empty_bar(s));
            }
            if (odd || z == Din - 1) commit(accf_bar((g0 + pz) % NSLOT));
          } else {  // M_DEC: rows are input positions; input plane z -> output planes 2z-1 (kd=0), 2z (kd=1), 2z+1 (kd=2)
            wait_acc_empty(g0 + 2 * z);
            wait_acc_empty(g0 + 2 * z + 1);
            const uint32_t s = q % C::STAGES;
            ptx::mbar_wait(full_bar(s), (q / C::STAGES) & 1);
            ptx::tc_fence_after();
            const uint64_t a0 = desc_add(a_desc0, s * C::STAGE_BYTES);
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
              const int qo = 2 * z - 1 + kd;
              if (qo < 0 || qo >= Dout) continue;
              const uint32_t dbase = tmem_base + ((g0 + qo) % NSLOT) * C::ACC_COLS;
#pragma unroll
              for (int e = 0; e < 5; ++e) {
#pragma unroll
                for (int ks = 0; ks < C::KSM; ++ks) {
                  const uint32_t aoff = (dec_shift_h(e) * MC::SUB_W + dec_shift_w(e)) * 16 + A_KOFF(ks);
                  // weights: [kd][ks][kcore][9*NT rows][8]; entry e owns rows [row_off, row_off+rows) (tables written for NT = 32)
                  const uint32_t boff = ((kd * C::KSW + B_KS(ks)) * 2 * 9 * NT + dec_row_off(e) * NT / 32) * 16;
                  const uint64_t bd = ptx::make_smem_desc(w_addr + boff, 9 * NT * 16, 128);
                  mma(dbase + dec_dcol(e) * NT / 32, desc_add(a0, aoff), bd, ptx::make_idesc_h<F16>(128, dec_rows(e) * NT / 32));
                }
              }
            }
            commit(empty_bar(s));
            ++q;
            if (z >= 1) commit(accf_bar((g0 + 2 * z - 1) % NSLOT));
            commit(accf_bar((g0 + 2 * z) % NSLOT));
            if (z == Din - 1) commit(accf_bar((g0 + 2 * z + 1) % NSLOT));
          }
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ================= input side copy (Params::x_split; the two otherwise idle warps) =================
    // Every input plane tile that passes through a stage is also written to global memory in the parity layout: consumers of the
    // stage are then the MMA warp (commit) AND these two warps (one arrival each).  64 threads move the 128 voxels x 8 channel
    // blocks (hi 4 | lo 4) of a tile, 16 x 16 B each; a chunked item's halo planes are written twice (same bytes).
    if constexpr (C::MRG) {
      if (p.x_split) {
        const int ct = (warp - 2) * 32 + lane;
        const int64_t Vo = (int64_t)Dout * p.Ho * p.Wo, sub = Vo / 8, plane_spl = (int64_t)(p.Ho / 2) * (p.Wo / 2) * 8;
        const Items items = make_items(ncols, ncta, Dout, !(p.dbg & 2048));
        uint32_t q = 0;
        for (int item = cta; item < items.total; item += ncta) {
          int col, e0, e1;
          get_item(items, item, Dout, col, e0, e1);
          const int tw = col % p.tiles_w, th = (col / p.tiles_w) % p.tiles_h, n = col / (p.tiles_w * p.tiles_h);
          for (int z = (e0 > 0 ? e0 - 1 : 0); z < (e1 < Din ? e1 + 1 : Din); ++z, ++q) {
            const uint32_t sq = q % C::STAGES;
            ptx::mbar_wait(full_bar(sq), (q / C::STAGES) & 1);
#pragma unroll 4
            for (int j = 0; j < 16; ++j) {
              const int u = ct + 64 * j, v = u & 127, b8 = u >> 7;        // voxel of the tile, block of [hi 4 | lo 4]
              const int hl = v >> 3, wl = v & 7, hr = th * TH + hl, wr = tw * TW + wl;
              if (hr < p.Hr && wr < p.Wr) {
                const uint4 val = ptx::lds_v4(stage_addr0 + sq * C::STAGE_BYTES + b8 * C::PLANE_BYTES + ((hl + 1) * MC::SUB_W + (wl + 1)) * 16);
                const int64_t off = ((int64_t)n * 8 + b8) * Vo * 8 +
                                    ((int64_t)((z & 1) * 4 + (hr & 1) * 2 + (wr & 1)) * sub + (int64_t)(hr >> 1) * (p.Wo / 2) + (wr >> 1)) * 8 + (int64_t)(z >> 1) * plane_spl;
                ptx::stg_cs_v4(p.x_split + off, val);
              }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty_bar(sq));
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (4 warps = 128 TMEM lanes) =================
    const int quarter = warp & 3;
    const int egroup = (warp >> 2) - 1;  // 0 or 1: the two epilogue groups take alternate output planes, so two planes'
                                         // TMEM-drain / residual-load / store latency chains are in flight per CTA
    const int m = quarter * 32 + lane;  // GEMM row = TMEM lane
    const int wl = m & 7, hl = m >> 3;  // position inside the 8 x 16 row tile
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int64_t Vo = (int64_t)Dout * p.Ho * p.Wo;
    const int cblk_out = p.Cout / 8;
    const int out_blocks = (X2 && p.x2) ? 2 * cblk_out : cblk_out;  // channel blocks per sample in y / residual / y_split
    const int64_t sub = Vo / 8;                              // voxels of one parity sub-volume
    uint32_t zero[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) zero[i] = 0u;
    uint32_t g0 = 0, tq = 0;  // tq: step counter of the TRI mode (one accumulator triple per input plane)
    const int DIe = C::DTR ? Din : Dout;   // (Cfg::DTR counts input planes, the others output planes; Cfg::MRG: Dout == Din)
    const Items items = make_items(ncols, ncta, DIe, (C::MRG || C::S2T || C::DTR) && !(p.dbg & 2048));
    for (int item = cta; item < items.total; item += ncta, g0 += Dout) {
      int col, e0, e1;   // this item's column and the planes [e0, e1) it owns
      get_item(items, item, DIe, col, e0, e1);
      const int tw = col % p.tiles_w, th = (col / p.tiles_w) % p.tiles_h, n = col / (p.tiles_w * p.tiles_h);
      const int hr = th * TH + hl, wr = tw * TW + wl;
      const bool valid = hr < p.Hr && wr < p.Wr;
      // Split-precision passes: the operands the epilogue adds (fp32 partial of the earlier pass, hi and lo words of the
      // residual) are independent of the accumulator -> requested as a batch before it is needed (xload), consumed in finish.
      auto xload = [&](XPre &q, int cb, int64_t pos, int cls, int64_t sidx) {
        const int cbg = nh * (NT / 8) + cb;
        if (p.part_in) {
          const float4 *pp = reinterpret_cast<const float4 *>(p.part_in + (((int64_t)n * cblk_out + cbg) * Vo + pos) * 8);
          q.p0 = __ldg(pp); q.p1 = __ldg(pp + 1);
        }
        if (p.residual) {
          const int64_t ro = p.residual_is_split ? ((((int64_t)n * out_blocks + cbg) * 8 + cls) * sub + sidx) * 8
                                                 : (((int64_t)n * out_blocks + cbg) * Vo + pos) * 8;
          q.rh = __ldg(reinterpret_cast<const uint4 *>(p.residual + ro));
          if (p.x2) q.rl = __ldg(reinterpret_cast<const uint4 *>(p.residual + ro + (int64_t)cblk_out * (p.residual_is_split ? 8 * sub : Vo) * 8));
          else q.rl = make_uint4(0u, 0u, 0u, 0u);
        }
      };
      // Everything after the accumulator: (+ fp32 partial of an earlier pass) -> either the fp32 partial of this pass, or
      // + bias (+ residual) (ReLU) -> 16-bit store(s).  `cb` = channel block inside this CTA's slice, `pos` = natural
      // voxel index, (`cls`, `sidx`) = parity class and index inside the parity sub-volume, `pre` = prefetched residual
      // (one-word modes), `xq` = prefetched operands of a split-precision pass.
      auto finish = [&](F8 a, int cb, int64_t pos, int cls, int64_t sidx, const uint4 &pre, const XPre &xq) {
        const int cbg = nh * (NT / 8) + cb;
        const int64_t onat = (((int64_t)n * out_blocks + cbg) * Vo + pos) * 8;
        const int64_t ospl = ((((int64_t)n * out_blocks + cbg) * 8 + cls) * sub + sidx) * 8;
        if (X2 && (p.part_in || p.part_out)) {
          if (p.part_in) {
            a.v[0] += xq.p0.x; a.v[1] += xq.p0.y; a.v[2] += xq.p0.z; a.v[3] += xq.p0.w;
            a.v[4] += xq.p1.x; a.v[5] += xq.p1.y; a.v[6] += xq.p1.z; a.v[7] += xq.p1.w;
          }
          if (p.part_out) {
            float4 *po = reinterpret_cast<float4 *>(p.part_out + (((int64_t)n * cblk_out + cbg) * Vo + pos) * 8);
            po[0] = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
            po[1] = make_float4(a.v[4], a.v[5], a.v[6], a.v[7]);
            return;
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) a.v[c] += bias_s[cb * 8 + c];
        if (p.residual) {
          const F8 q = unpack8h<F16>(X2 ? xq.rh : pre);
#pragma unroll
          for (int c = 0; c < 8; ++c) a.v[c] += q.v[c];
          if (X2) {  // low word of the residual
            const F8 ql = unpack8h<F16>(xq.rl);
#pragma unroll
            for (int c = 0; c < 8; ++c) a.v[c] += ql.v[c];
          }
        }
        if (p.relu) {
#pragma unroll
          for (int c = 0; c < 8; ++c) a.v[c] = fmaxf(a.v[c], 0.f);
        }
        if (F16 && p.range_flag) {  // fp16 words: beyond 65504 (or NaN) the stored activation is no longer the computed one
          bool bad = false;
#pragma unroll
          for (int c = 0; c < 8; ++c) bad |= !(fabsf(a.v[c]) <= 65504.f);
          if (bad) *p.range_flag = 1;
        }
        const uint4 hi = pack8h<F16>(a);
        if (!p.skip_y) *reinterpret_cast<uint4 *>(p.y + onat) = hi;
        if (p.y_split) *reinterpret_cast<uint4 *>(p.y_split + ospl) = hi;
        if (X2 && p.x2) {  // split-precision output: second 16-bit word holds what the first one rounded away
          const F8 h = unpack8h<F16>(hi);
#pragma unroll
          for (int c = 0; c < 8; ++c) a.v[c] -= h.v[c];
          const uint4 lo = pack8h<F16>(a);
          if (!p.skip_y) *reinterpret_cast<uint4 *>(p.y + onat + (int64_t)cblk_out * Vo * 8) = lo;
          if (p.y_split) *reinterpret_cast<uint4 *>(p.y_split + ospl + (int64_t)cblk_out * 8 * sub * 8) = lo;
        }
      };
      if constexpr ((C::TRI && NT == 32) || C::S2T) {
        // ---- per-step triples, 32-wide blocks (every stride-1 split-precision layer; S2T: per-step pairs of the stride-2 conv) ----
        // All EGROUPS epilogue groups drain EVERY step's triple; group `egroup` owns CPG = 32/EGROUPS output channels = NCBG channel
        // blocks starting at cbg0.  Measured (IDISP_TC_DBG=32/28): with the epilogue reduced to its barrier handshake a 32->32 layer
        // runs at the MMA stream's 1.54 ms, with it at 2.2 ms, and neither the TMEM traffic nor the global stores matter -- the
        // epilogue's INSTRUCTION count is the limiter.  So: all per-column addressing is hoisted out of the step loop (plane
        // strides are added to running offsets), nothing is zeroed on drain (the first MMA of a step overwrites), the half-range
        // test works on the packed words.
        constexpr int CPG = 32 / C::EGROUPS, NCBG = CPG / 8;   // channels / channel blocks per epilogue group
        const int cbg0 = nh * 4 + egroup * NCBG;
        const int64_t blk_elems = Vo * 8;                                   // one channel block, either layout (8 * sub = Vo)
        const int64_t lo_off = (int64_t)cblk_out * blk_elems;               // hi word -> lo word
        const int64_t col_blk = ((int64_t)n * out_blocks + cbg0) * blk_elems;
        const int64_t plane_nat = (int64_t)p.Ho * p.Wo * 8, plane_spl = (int64_t)(p.Ho / 2) * (p.Wo / 2) * 8;
        const int64_t nat0 = col_blk + ((int64_t)hr * p.Wo + wr) * 8;       // natural layout, plane 0
        const int64_t spl0 = col_blk + ((int64_t)((hr & 1) * 2 + (wr & 1)) * sub + (int64_t)(hr >> 1) * (p.Wo / 2) + (wr >> 1)) * 8;  // parity layout, plane 0
        const int64_t part0 = (((int64_t)n * cblk_out + cbg0) * Vo + (int64_t)hr * p.Wo + wr) * 8;   // fp32 partial (natural), plane 0
        const bool has_res = p.residual != nullptr, has_part = p.part_in != nullptr, out_x2 = p.x2 != 0;
        float P0[CPG], P1[CPG];
#pragma unroll
        for (int i = 0; i < CPG; ++i) { P0[i] = 0.f; P1[i] = 0.f; }
        bool bad = false;
        // element offset of plane q inside a channel block
        auto nat_of = [&](int q) { return nat0 + (int64_t)q * plane_nat; };
        auto spl_of = [&](int q) { return spl0 + (int64_t)(q & 1) * 4 * sub * 8 + (int64_t)(q >> 1) * plane_spl; };
        auto emit = [&](int q, const float (&sum)[CPG], const XPre (&xq)[NCBG]) {
          const int64_t onat = nat_of(q), ospl = spl_of(q);
#pragma unroll
          for (int i = 0; i < NCBG; ++i) {
            float a[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] = sum[i * 8 + c];
            if (has_part) {
              a[0] += xq[i].p0.x; a[1] += xq[i].p0.y; a[2] += xq[i].p0.z; a[3] += xq[i].p0.w;
              a[4] += xq[i].p1.x; a[5] += xq[i].p1.y; a[6] += xq[i].p1.z; a[7] += xq[i].p1.w;
            }
            if (p.part_out) {
              float4 *po = reinterpret_cast<float4 *>(p.part_out + part0 + (int64_t)q * plane_nat + (int64_t)i * blk_elems);
              po[0] = make_float4(a[0], a[1], a[2], a[3]);
              po[1] = make_float4(a[4], a[5], a[6], a[7]);
              continue;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] += bias_s[egroup * CPG + i * 8 + c];
            if (has_res) {
              const F8 rh = unpack8h<F16>(xq[i].rh);
#pragma unroll
              for (int c = 0; c < 8; ++c) a[c] += rh.v[c];
              if (out_x2) {
                const F8 rl = unpack8h<F16>(xq[i].rl);
#pragma unroll
                for (int c = 0; c < 8; ++c) a[c] += rl.v[c];
              }
            }
            if (p.relu) {
#pragma unroll
              for (int c = 0; c < 8; ++c) a[c] = fmaxf(a[c], 0.f);
            }
            F8 f;
#pragma unroll
            for (int c = 0; c < 8; ++c) f.v[c] = a[c];
            const uint4 hi = pack8h<F16>(f);
            if (F16) {  // a half whose exponent field is all ones: the value left the IEEE-half range (or was NaN)
              const uint32_t m = ((hi.x & 0x7fff7fffu) + 0x04000400u) | ((hi.y & 0x7fff7fffu) + 0x04000400u) |
                                 ((hi.z & 0x7fff7fffu) + 0x04000400u) | ((hi.w & 0x7fff7fffu) + 0x04000400u);
              bad |= (m & 0x80008000u) != 0;
            }
            const int64_t bo = (int64_t)i * blk_elems;
            if (!p.skip_y) *reinterpret_cast<uint4 *>(p.y + onat + bo) = hi;
            if (p.y_split) *reinterpret_cast<uint4 *>(p.y_split + ospl + bo) = hi;
            if (out_x2) {  // second word: what the first one rounded away
              const F8 h = unpack8h<F16>(hi);
#pragma unroll
              for (int c = 0; c < 8; ++c) f.v[c] -= h.v[c];
              const uint4 lo = pack8h<F16>(f);
              if (!p.skip_y) *reinterpret_cast<uint4 *>(p.y + onat + bo + lo_off) = lo;
              if (p.y_split) *reinterpret_cast<uint4 *>(p.y_split + ospl + bo + lo_off) = lo;
            }
          }
        };
        auto xload2 = [&](XPre (&xq)[NCBG], int q) {
          if (has_part) {
#pragma unroll
            for (int i = 0; i < NCBG; ++i) {
              const float4 *pp = reinterpret_cast<const float4 *>(p.part_in + part0 + (int64_t)q * plane_nat + (int64_t)i * blk_elems);
              xq[i].p0 = __ldg(pp); xq[i].p1 = __ldg(pp + 1);
            }
          }
          if (has_res) {
            const int64_t ro = p.residual_is_split ? spl_of(q) : nat_of(q);
#pragma unroll
            for (int i = 0; i < NCBG; ++i) {
              xq[i].rh = __ldg(reinterpret_cast<const uint4 *>(p.residual + ro + (int64_t)i * blk_elems));
              xq[i].rl = out_x2 ? __ldg(reinterpret_cast<const uint4 *>(p.residual + ro + (int64_t)i * blk_elems + lo_off)) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
        };
        if constexpr (C::S2T) {
          float CAR[CPG];   // block 1 of the previous step: what input plane 2p-1 (kd = 0) contributed to output plane p
#pragma unroll
          for (int i = 0; i < CPG; ++i) CAR[i] = 0.f;
          for (int pz = (e0 > 0 ? e0 - 1 : 0); pz < e1; ++pz, ++tq) {   // (a chunk's first step only builds the carry into plane e0)
            const uint32_t t = tq % NSLOT;
            XPre xq[NCBG];
            if (valid && pz >= e0) xload2(xq, pz);
            ptx::mbar_wait(accf_bar(t), (tq / NSLOT) & 1);
            ptx::tc_fence_after();
            const uint32_t tb = tmem_base + lane_addr + t * C::S2T_STRIDE + egroup * CPG;   // [C0 | M0 | M1 | C1 | X0 | X1]
            {
              uint32_t c0[CPG], m0[CPG], x0[CPG];
              ptx::tmem_ld_cols(tb, c0);
              ptx::tmem_ld_cols(tb + NT, m0);
              ptx::tmem_ld_cols(tb + 4 * NT, x0);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < CPG; ++i) P0[i] = CAR[i] + (__uint_as_float(m0[i]) + (__uint_as_float(c0[i]) + __uint_as_float(x0[i])));
            }
            {
              uint32_t m1[CPG], c1[CPG], x1[CPG];
              ptx::tmem_ld_cols(tb + 2 * NT, m1);
              ptx::tmem_ld_cols(tb + 3 * NT, c1);
              ptx::tmem_ld_cols(tb + 5 * NT, x1);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < CPG; ++i) CAR[i] = __uint_as_float(m1[i]) + (__uint_as_float(c1[i]) + __uint_as_float(x1[i]));
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acce_bar(t));
            if (valid && pz >= e0 && !(p.dbg & 4)) emit(pz, P0, xq);
            if (pz == Dout - 1) {
#pragma unroll
              for (int i = 0; i < CPG; ++i) CAR[i] = 0.f;
            }
          }
        } else
        for (int z = (e0 > 0 ? e0 - 1 : 0); z < (e1 < Dout ? e1 + 1 : Dout); ++z, ++tq) {
          const uint32_t t = tq % NSLOT;
          XPre xq[NCBG];
          const bool mine = z - 1 >= e0;   // plane z-1 belongs to this item (a chunk's first two steps only build up its first plane)
          if (valid && mine) xload2(xq, z - 1);      // operands of the plane this step completes: requested before the wait
          ptx::mbar_wait(accf_bar(t), (tq / NSLOT) & 1);
          ptx::tc_fence_after();
          if (p.dbg & 32) {  // timing experiment: handshake only (no TMEM traffic, no arithmetic, no global memory)
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acce_bar(t));
            continue;
          }
          const uint32_t tb = tmem_base + lane_addr + t * C::TRI_STRIDE + egroup * CPG;
          if (p.dbg & 128) {  // timing experiment: no TMEM reads, the arithmetic and the stores run on whatever the registers hold
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acce_bar(t));
            if (mine && valid) emit(z - 1, P0, xq);
            continue;
          }
          // plane z-1 = P0 + block 0 (complete);  plane z: P1 + block 1;  plane z+1: block 2 (first contribution).
          // Two TMEM round trips (blocks 0+1, then block 2) keep the live registers under the 168-register cap; the
          // correction triple (x_lo*w_hi + x_hi*w_lo) is summed with the main one in fp32 round-to-nearest.
          float N0[CPG];   // next step's P0 = plane z so far
          {
            uint32_t b0[CPG], b1[CPG];
            ptx::tmem_ld_cols(tb, b0);
            ptx::tmem_ld_cols(tb + NT, b1);
            if (XP != 0) {
              uint32_t u0[CPG], u1[CPG];
              ptx::tmem_ld_cols(tb + C::TRI_SMALL, u0);
              ptx::tmem_ld_cols(tb + C::TRI_SMALL + NT, u1);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < CPG; ++i) {
                P0[i] += __uint_as_float(b0[i]) + __uint_as_float(u0[i]);
                N0[i] = P1[i] + (__uint_as_float(b1[i]) + __uint_as_float(u1[i]));
              }
            } else {
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < CPG; ++i) { P0[i] += __uint_as_float(b0[i]); N0[i] = P1[i] + __uint_as_float(b1[i]); }
            }
          }
          {
            uint32_t b2[CPG];
            ptx::tmem_ld_cols(tb + 2 * NT, b2);
            if (XP != 0) {
              uint32_t u2[CPG];
              ptx::tmem_ld_cols(tb + C::TRI_SMALL + 2 * NT, u2);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < CPG; ++i) P1[i] = __uint_as_float(b2[i]) + __uint_as_float(u2[i]);
            } else {
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < CPG; ++i) P1[i] = __uint_as_float(b2[i]);
            }
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acce_bar(t));   // the MMA warp may overwrite this buffer (first MMA of a step: accumulate = 0)
          if (mine && valid && !(p.dbg & 64)) emit(z - 1, P0, xq);   // (64: timing experiment without the per-voxel work)
#pragma unroll
          for (int i = 0; i < CPG; ++i) P0[i] = N0[i];
          if (z == Dout - 1) {  // no step z+1: plane z is complete as well
            if (valid) {
              xload2(xq, z);
              emit(z, P0, xq);
            }
#pragma unroll
            for (int i = 0; i < CPG; ++i) { P0[i] = 0.f; P1[i] = 0.f; }
          }
        }
        if (bad && p.range_flag) *p.range_flag = 1;
      } else if constexpr (C::TRI) {
        // (16-wide blocks: the 1-channel head in its generic-kernel form, kept as a cross-check of head_tc.cu -- IDISP_OLD_HEAD=1)
        // Both epilogue groups drain EVERY step's triple; group `egroup` owns output channels [16*egroup, 16*egroup+16) of the
        // CTA's 32 (16-wide blocks, i.e. the 1-channel head: group 0 owns all of them, group 1 only keeps the barrier count).
        // P0 / P1: running sums of the two open planes (z and z+1 after step z).
        const bool owner = NT == 32 || egroup == 0;
        const int cbase = NT == 32 ? egroup * 16 : 0;   // first accumulator column (of a block) this group drains
        float P0[16], P1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { P0[i] = 0.f; P1[i] = 0.f; }
        for (int z = 0; z < Dout; ++z, ++tq) {
          const uint32_t t = tq % NSLOT;
          const int qf = z - 1;  // the plane this step completes (also plane z itself at the last step)
          XPre xq[2];
          if (valid && qf >= 0 && !p.y1) {
            const int64_t pos = ((int64_t)qf * p.Ho + hr) * p.Wo + wr;
            const int64_t sidx = ((int64_t)(qf >> 1) * (p.Ho / 2) + (hr >> 1)) * (p.Wo / 2) + (wr >> 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) xload(xq[i], egroup * 2 + i, pos, (qf & 1) * 4 + (hr & 1) * 2 + (wr & 1), sidx);
          }
          ptx::mbar_wait(accf_bar(t), (tq / NSLOT) & 1);
          ptx::tc_fence_after();
          if (p.dbg & 32) {  // timing experiment: handshake only (no TMEM traffic, no arithmetic, no global memory)
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acce_bar(t));
            continue;
          }
          uint32_t b0[16], b1[16], b2[16];
          if (owner) {
            const uint32_t tb = tmem_base + lane_addr + t * C::TRI_STRIDE + cbase;
            ptx::tmem_ld_32x16(tb, b0);
            ptx::tmem_ld_32x16(tb + NT, b1);
            ptx::tmem_ld_32x16(tb + 2 * NT, b2);
            ptx::tmem_ld_wait();
            if (XP != 0) {  // correction terms: their own triple, summed in fp32 (round to nearest)
              uint32_t u0[16], u1[16], u2[16];
              ptx::tmem_ld_32x16(tb + C::TRI_SMALL, u0);
              ptx::tmem_ld_32x16(tb + C::TRI_SMALL + NT, u1);
              ptx::tmem_ld_32x16(tb + C::TRI_SMALL + 2 * NT, u2);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                b0[i] = __float_as_uint(__uint_as_float(b0[i]) + __uint_as_float(u0[i]));
                b1[i] = __float_as_uint(__uint_as_float(b1[i]) + __uint_as_float(u1[i]));
                b2[i] = __float_as_uint(__uint_as_float(b2[i]) + __uint_as_float(u2[i]));
              }
            }
            ptx::tmem_st_32x16(tb, zero); ptx::tmem_st_32x16(tb + NT, zero); ptx::tmem_st_32x16(tb + 2 * NT, zero);
            if (XP != 0) {
              ptx::tmem_st_32x16(tb + C::TRI_SMALL, zero); ptx::tmem_st_32x16(tb + C::TRI_SMALL + NT, zero);
              ptx::tmem_st_32x16(tb + C::TRI_SMALL + 2 * NT, zero);
            }
            ptx::tmem_st_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) { b0[i] = 0u; b1[i] = 0u; b2[i] = 0u; }
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acce_bar(t));
          // plane z-1 = P0 + block 0 (complete);  plane z: P1 + block 1;  plane z+1: block 2 (first contribution)
          auto emit = [&](int q_, const float *sum, const uint32_t *blk, const XPre *pre2) {
            const int64_t pos = ((int64_t)q_ * p.Ho + hr) * p.Wo + wr;
            if (p.y1) {  // 32->1 classifier head: channel 0 (+ column 1 = the w_lo products), f32, running sum fused
              if (owner) {
                const int64_t o1 = (int64_t)n * Vo + pos;
                float v1 = sum[0] + (blk ? __uint_as_float(blk[0]) : 0.f);
                if (p.y1_cols == 2) v1 += sum[1] + (blk ? __uint_as_float(blk[1]) : 0.f);
                p.y1[o1] = v1 + (p.res1 ? p.res1[o1] : 0.f);
              }
              return;
            }
            const int64_t sidx = ((int64_t)(q_ >> 1) * (p.Ho / 2) + (hr >> 1)) * (p.Wo / 2) + (wr >> 1);
            const int cls = (q_ & 1) * 4 + (hr & 1) * 2 + (wr & 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              XPre one;
              if (!pre2) xload(one, egroup * 2 + i, pos, cls, sidx);
              F8 a8;
#pragma unroll
              for (int c = 0; c < 8; ++c) a8.v[c] = sum[i * 8 + c] + (blk ? __uint_as_float(blk[i * 8 + c]) : 0.f);
              finish(a8, egroup * 2 + i, pos, cls, sidx, one.rh, pre2 ? pre2[i] : one);
            }
          };
          if (valid && qf >= 0) emit(qf, P0, b0, xq);
#pragma unroll
          for (int i = 0; i < 16; ++i) { P0[i] = P1[i] + __uint_as_float(b1[i]); P1[i] = __uint_as_float(b2[i]); }
          if (z == Dout - 1) {  // no step z+1: plane z is complete as well
            if (valid) emit(z, P0, nullptr, nullptr);
#pragma unroll
            for (int i = 0; i < 16; ++i) { P0[i] = 0.f; P1[i] = 0.f; }
          }
        }
      } else if constexpr (C::DTR) {
        // ---- per-step buffers of the transposed conv (see Cfg::DTR) ----
        // Rows are INPUT positions (hr, wr).  Both epilogue groups drain every step: group `egroup` owns the output rows of parity
        // ph = egroup, i.e. the classes (ph, pw = 0) and (ph, pw = 1) -- the neighbouring output voxels (2wr, 2wr+1) -- and handles
        // them TOGETHER plane by plane, so every natural-layout access of a thread is one 32-byte sector (256-bit LDG / STG; with one
        // class at a time each 16-byte store half-filled a sector and the pw = 1 store came 16 stores later).
        //   plane 2z-1 = carry + block 0, plane 2z = block 1, carry = block 2   (per class; fp32 round-to-nearest, in registers)
        // The residual words of a plane are requested before its TMEM read (the producer warp prefetched them into L2 earlier).
        const int ph = egroup;
        const int ci0 = ph == 0 ? 0 : 1, ci1 = ph == 0 ? 3 : 2;   // column groups of (ph, pw=0), (ph, pw=1): [c00 | c10 | c11 | c01]
        const bool live_col = valid && !(p.dbg & 4);
        const int64_t blk_elems = Vo * 8, lo_off = (int64_t)cblk_out * blk_elems;
        const int64_t col_blk = ((int64_t)n * out_blocks + nh * 2) * blk_elems;
        const int64_t plane_nat = (int64_t)p.Ho * p.Wo * 8, plane_spl = (int64_t)(p.Ho / 2) * (p.Wo / 2) * 8;
        const int64_t nat0 = col_blk + ((int64_t)(2 * hr + ph) * p.Wo + 2 * wr) * 8;                                      // plane 0, pw 0
        const int64_t spl0 = col_blk + ((int64_t)(ph * 2) * sub + (int64_t)hr * (p.Wo / 2) + wr) * 8;                      // plane 0, pw 0
        auto nat_of = [&](int q) { return nat0 + (int64_t)q * plane_nat; };                                               // (pw 1: + 8 elements)
        auto spl_of = [&](int q) { return spl0 + (int64_t)(q & 1) * 4 * sub * 8 + (int64_t)(q >> 1) * plane_spl; };      // (pw 1: + sub * 8)
        const bool has_res = p.residual != nullptr && !(p.dbg & 512), out_x2 = p.x2 != 0, res_split = p.residual_is_split != 0;
        struct Res { uint4 h[2][2], l[2][2]; };   // residual words of the voxel pair: [pw][channel block] x (hi, lo)
        auto rload = [&](Res &r, int q) {
          if (res_split) {
            const __nv_bfloat16 *rp = p.residual + spl_of(q);
#pragma unroll
            for (int pw = 0; pw < 2; ++pw)
#pragma unroll
              for (int cb = 0; cb < 2; ++cb) {
                r.h[pw][cb] = __ldg(reinterpret_cast<const uint4 *>(rp + (int64_t)pw * sub * 8 + (int64_t)cb * blk_elems));
                r.l[pw][cb] = out_x2 ? __ldg(reinterpret_cast<const uint4 *>(rp + (int64_t)pw * sub * 8 + (int64_t)cb * blk_elems + lo_off)) : make_uint4(0u, 0u, 0u, 0u);
              }
          } else {
            const __nv_bfloat16 *rp = p.residual + nat_of(q);
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
              const ptx::U8 a = ptx::ldg_v8(rp + (int64_t)cb * blk_elems);
              r.h[0][cb] = a.a; r.h[1][cb] = a.b;
              if (out_x2) {
                const ptx::U8 b = ptx::ldg_v8(rp + (int64_t)cb * blk_elems + lo_off);
                r.l[0][cb] = b.a; r.l[1][cb] = b.b;
              } else {
                r.l[0][cb] = make_uint4(0u, 0u, 0u, 0u); r.l[1][cb] = make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
        };
        bool bad = false;
        // one voxel's 8 channels of one channel block: + bias (+ residual) (ReLU) -> hi / lo words
        auto cook = [&](const float *v8, int cb, const uint4 &rh, const uint4 &rl, uint4 &hi, uint4 &lo) {
          F8 f;
#pragma unroll
          for (int c = 0; c < 8; ++c) f.v[c] = v8[c] + bias_s[cb * 8 + c];
          if (has_res) {
            const F8 r0 = unpack8h<F16>(rh);
#pragma unroll
            for (int c = 0; c < 8; ++c) f.v[c] += r0.v[c];
            if (out_x2) {
              const F8 r1 = unpack8h<F16>(rl);
#pragma unroll
              for (int c = 0; c < 8; ++c) f.v[c] += r1.v[c];
            }
          }
          if (p.relu) {
#pragma unroll
            for (int c = 0; c < 8; ++c) f.v[c] = fmaxf(f.v[c], 0.f);
          }
          hi = pack8h<F16>(f);
          {  // a half whose exponent field is all ones: the value left the IEEE-half range (or was NaN)
            const uint32_t mm = ((hi.x & 0x7fff7fffu) + 0x04000400u) | ((hi.y & 0x7fff7fffu) + 0x04000400u) |
                                ((hi.z & 0x7fff7fffu) + 0x04000400u) | ((hi.w & 0x7fff7fffu) + 0x04000400u);
            bad |= (mm & 0x80008000u) != 0;
          }
          const F8 h = unpack8h<F16>(hi);
#pragma unroll
          for (int c = 0; c < 8; ++c) f.v[c] -= h.v[c];
          lo = pack8h<F16>(f);
        };
        auto emit = [&](int q, const float (&v0)[16], const float (&v1)[16], const Res &r) {
          const int64_t onat = nat_of(q), ospl = spl_of(q);
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            uint4 h0, l0, h1, l1;
            cook(v0 + cb * 8, cb, r.h[0][cb], r.l[0][cb], h0, l0);
            cook(v1 + cb * 8, cb, r.h[1][cb], r.l[1][cb], h1, l1);
            const int64_t bo = (int64_t)cb * blk_elems;
            if (p.dbg & 1024) {  // A/B: default cache policy
              if (!p.skip_y) {
                ptx::stg_v8(p.y + onat + bo, h0, h1);
                if (out_x2) ptx::stg_v8(p.y + onat + bo + lo_off, l0, l1);
              }
              if (p.y_split) {
                *reinterpret_cast<uint4 *>(p.y_split + ospl + bo) = h0;
                *reinterpret_cast<uint4 *>(p.y_split + ospl + bo + sub * 8) = h1;
                if (out_x2) {
                  *reinterpret_cast<uint4 *>(p.y_split + ospl + bo + lo_off) = l0;
                  *reinterpret_cast<uint4 *>(p.y_split + ospl + bo + lo_off + sub * 8) = l1;
                }
              }
              continue;
            }
            if (!p.skip_y) {
              ptx::stg_cs_v8(p.y + onat + bo, h0, h1);
              if (out_x2) ptx::stg_cs_v8(p.y + onat + bo + lo_off, l0, l1);
            }
            if (p.y_split) {
              ptx::stg_cs_v4(p.y_split + ospl + bo, h0);
              ptx::stg_cs_v4(p.y_split + ospl + bo + sub * 8, h1);
              if (out_x2) {
                ptx::stg_cs_v4(p.y_split + ospl + bo + lo_off, l0);
                ptx::stg_cs_v4(p.y_split + ospl + bo + lo_off + sub * 8, l1);
              }
            }
          }
        };
        float CAR0[16], CAR1[16];   // block 2 of the previous step: what input plane z-1 (kd = 2) contributed to output plane 2z-1
#pragma unroll
        for (int i = 0; i < 16; ++i) { CAR0[i] = 0.f; CAR1[i] = 0.f; }
        for (int z = (e0 > 0 ? e0 - 1 : 0); z < e1; ++z, ++tq) {
          const uint32_t t = tq % NSLOT;
          const uint32_t tb0 = tmem_base + lane_addr + t * C::DTR_STRIDE + ci0 * 3 * NT, tb1 = tmem_base + lane_addr + t * C::DTR_STRIDE + ci1 * 3 * NT;
          const bool live = live_col && z >= e0;   // (a chunk's extra first step only produces the carry into plane 2*e0-1)
          Res r, r2;
          if (live && has_res && z >= 1) rload(r, 2 * z - 1);
          ptx::mbar_wait(accf_bar(t), (tq / NSLOT) & 1);
          ptx::tc_fence_after();
          {
            uint32_t a0[16], a1[16];
            ptx::tmem_ld_32x16(tb0, a0);
            ptx::tmem_ld_32x16(tb1, a1);
            if (live && has_res) rload(r2, 2 * z);   // (in flight during the per-voxel work of plane 2z-1)
            ptx::tmem_ld_wait();
            if (live && z >= 1) {
              float v0[16], v1[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) { v0[i] = CAR0[i] + __uint_as_float(a0[i]); v1[i] = CAR1[i] + __uint_as_float(a1[i]); }
              emit(2 * z - 1, v0, v1, r);
            }
          }
          {
            uint32_t b0[16], b1[16], c0[16], c1[16];
            ptx::tmem_ld_32x16(tb0 + NT, b0);
            ptx::tmem_ld_32x16(tb1 + NT, b1);
            ptx::tmem_ld_32x16(tb0 + 2 * NT, c0);
            ptx::tmem_ld_32x16(tb1 + 2 * NT, c1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acce_bar(t));   // the MMA warp may overwrite this buffer
#pragma unroll
            for (int i = 0; i < 16; ++i) { CAR0[i] = __uint_as_float(c0[i]); CAR1[i] = __uint_as_float(c1[i]); }
            if (live) {
              float v0[16], v1[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) { v0[i] = __uint_as_float(b0[i]); v1[i] = __uint_as_float(b1[i]); }
              emit(2 * z, v0, v1, r2);
            }
          }
          if (z == Din - 1 && live) {  // no step z+1: plane 2z+1 is complete with its kd = 2 contribution
            if (has_res) rload(r, 2 * z + 1);
            emit(2 * z + 1, CAR0, CAR1, r);
          }
        }
        if (bad && p.range_flag) *p.range_flag = 1;
      }
      for (int qo = egroup; !C::TRI && !C::S2T && !C::DTR && qo < Dout; qo += C::EGROUPS) {
        const uint32_t g = g0 + qo, r = g % NSLOT;
        // residual operands do not depend on the accumulator: request them BEFORE waiting for it (single-precision-word
        // modes only; the split-precision passes load at use)
        constexpr int NRES = MODE == M_DEC ? 16 : (NT >= 8 ? NT / 8 : 1);
        uint4 resv[NRES];
        const bool prefetch = !X2 && p.residual && valid;
        constexpr int NXQ = (X2 && MODE != M_DEC && NT >= 8 && NT <= 32) ? NT / 8 : 1;
        XPre xq[NXQ];
        if (X2 && MODE != M_DEC && NT <= 32 && valid && !p.y1) {
          const int64_t pos = ((int64_t)qo * p.Ho + hr) * p.Wo + wr;
          const int64_t sidx = ((int64_t)(qo >> 1) * (p.Ho / 2) + (hr >> 1)) * (p.Wo / 2) + (wr >> 1);
#pragma unroll
          for (int cb = 0; cb < NXQ; ++cb) xload(xq[cb], cb, pos, (qo & 1) * 4 + (hr & 1) * 2 + (wr & 1), sidx);
        }
        constexpr bool DEC_LEAN = MODE == M_DEC && X2 && NT == 16;   // (this form loads its operands itself, before the wait: below)
        if (MODE == M_DEC && X2 && !DEC_LEAN && valid && (p.part_in || p.residual)) {
          // the plane's 16 voxel blocks per thread are consumed class by class below; ask L2 for all of them now
#pragma unroll
          for (int i = 0; i < 4 * (NT / 8); ++i) {
            const int cb = i % (NT / 8), ph = (i / (NT / 8)) >> 1, pw = (i / (NT / 8)) & 1, cbg = nh * (NT / 8) + cb;
            const int64_t pos = ((int64_t)qo * p.Ho + 2 * hr + ph) * p.Wo + 2 * wr + pw;
            if (p.part_in) ptx::prefetch_l2(p.part_in + (((int64_t)n * cblk_out + cbg) * Vo + pos) * 8);
            if (p.residual) {
              const int64_t ro = p.residual_is_split ? ((((int64_t)n * out_blocks + cbg) * 8 + (qo & 1) * 4 + ph * 2 + pw) * sub +
                                                         ((int64_t)(qo >> 1) * (p.Ho / 2) + hr) * (p.Wo / 2) + wr) * 8
                                                      : (((int64_t)n * out_blocks + cbg) * Vo + pos) * 8;
              ptx::prefetch_l2(p.residual + ro);
              if (p.x2) ptx::prefetch_l2(p.residual + ro + (int64_t)cblk_out * (p.residual_is_split ? 8 * sub : Vo) * 8);
            }
          }
        }
        if (prefetch) {
#pragma unroll
          for (int i = 0; i < NRES; ++i) {
            // DEC: i = ph*8 + cb*2 + pw -> voxel (2hr+ph, 2wr+pw); conv: i = cb
            const int cb = MODE == M_DEC ? (i >> 1) & 3 : i;
            const int64_t pos = MODE == M_DEC ? ((int64_t)qo * p.Ho + 2 * hr + (i >> 3)) * p.Wo + 2 * wr + (i & 1)
                                              : ((int64_t)qo * p.Ho + hr) * p.Wo + wr;
            int64_t ro = (((int64_t)n * out_blocks + nh * (NT / 8) + cb) * Vo + pos) * 8;
            if (MODE == M_DEC && p.residual_is_split)  // class (qo&1, ph, pw) at (qo>>1, hr, wr): 128 B contiguous per 8 rows
              ro = ((((int64_t)n * out_blocks + nh * 4 + cb) * 8 + (qo & 1) * 4 + (i >> 3) * 2 + (i & 1)) * sub +
                    ((int64_t)(qo >> 1) * (p.Ho / 2) + hr) * (p.Wo / 2) + wr) * 8;
            resv[i] = __ldg(reinterpret_cast<const uint4 *>(p.residual + ro));
          }
        }
        if constexpr (!DEC_LEAN) {
          ptx::mbar_wait(accf_bar(r), (g / NSLOT) & 1);
          ptx::tc_fence_after();
        }
        if constexpr (DEC_LEAN) {
          // Split-precision transposed conv (16-wide blocks).  The epilogue, not the MMA stream, bounds this layer (2.1 ms without
          // any MMA against 1.5 ms for the MMA stream alone, IDISP_TC_DBG ablations), and inside it the chain "load the residual
          // words of a parity class -> wait -> finish -> next class" was serialised on L2 / DRAM latency.  So: the operands of ALL
          // four classes (16 x 16 B per thread) are requested before the accumulator is even waited for (see below the wait),
          // addressing is hoisted, the per-voxel work is the lean form.
          {
            // (tcgen05.ld / .st are warp-collective: every lane runs them; only the global-memory work is predicated)
            const bool live = valid && !(p.dbg & 4);
            const int64_t blk_elems = Vo * 8, lo_off = (int64_t)cblk_out * blk_elems;
            const int64_t col_blk = ((int64_t)n * out_blocks + nh * 2) * blk_elems;
            const int64_t natq = col_blk + (((int64_t)qo * p.Ho + 2 * hr) * p.Wo + 2 * wr) * 8;
            const int64_t splq = col_blk + ((int64_t)(qo & 1) * 4 * sub + ((int64_t)(qo >> 1) * (p.Ho / 2) + hr) * (p.Wo / 2) + wr) * 8;
            const int64_t partq = (((int64_t)n * cblk_out + nh * 2) * Vo + ((int64_t)qo * p.Ho + 2 * hr) * p.Wo + 2 * wr) * 8;
            const bool has_res = p.residual != nullptr, has_part = p.part_in != nullptr, out_x2 = p.x2 != 0;
            bool bad = false;
            uint4 rh[4][2], rl[4][2];   // residual words of the 4 classes x 2 channel blocks
            if (has_res && live) {
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const int ph = c4 >> 1, pw = c4 & 1;
                const int64_t ro = p.residual_is_split ? splq + (int64_t)(ph * 2 + pw) * sub * 8 : natq + ((int64_t)ph * p.Wo + pw) * 8;
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                  rh[c4][cb] = __ldg(reinterpret_cast<const uint4 *>(p.residual + ro + (int64_t)cb * blk_elems));
                  rl[c4][cb] = out_x2 ? __ldg(reinterpret_cast<const uint4 *>(p.residual + ro + (int64_t)cb * blk_elems + lo_off)) : make_uint4(0u, 0u, 0u, 0u);
                }
              }
            }
            ptx::mbar_wait(accf_bar(r), (g / NSLOT) & 1);
            ptx::tc_fence_after();
            const uint32_t t0 = tmem_base + lane_addr + r * C::ACC_COLS;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int ph = c4 >> 1, pw = c4 & 1, cblock = pw * 2 + ph;   // accumulator column block of class (ph, pw)
              uint32_t v[16];
              ptx::tmem_ld_32x16(t0 + cblock * NT, v);
              ptx::tmem_ld_wait();
              ptx::tmem_st_32x16(t0 + cblock * NT, zero);
              if (c4 == 3) {
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(acce_bar(r));
              }
              if (!live) continue;
              const int64_t onat = natq + ((int64_t)ph * p.Wo + pw) * 8, ospl = splq + (int64_t)(ph * 2 + pw) * sub * 8;
              const int64_t opart = partq + ((int64_t)ph * p.Wo + pw) * 8;
#pragma unroll
              for (int cb = 0; cb < 2; ++cb) {
                float a[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) a[c] = __uint_as_float(v[cb * 8 + c]);
                if (has_part) {
                  const float4 *pp = reinterpret_cast<const float4 *>(p.part_in + opart + (int64_t)cb * blk_elems);
                  const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
                  a[0] += p0.x; a[1] += p0.y; a[2] += p0.z; a[3] += p0.w;
                  a[4] += p1.x; a[5] += p1.y; a[6] += p1.z; a[7] += p1.w;
                }
                if (p.part_out) {
                  float4 *po = reinterpret_cast<float4 *>(p.part_out + opart + (int64_t)cb * blk_elems);
                  po[0] = make_float4(a[0], a[1], a[2], a[3]);
                  po[1] = make_float4(a[4], a[5], a[6], a[7]);
                  continue;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) a[c] += bias_s[cb * 8 + c];
                if (has_res) {
                  const F8 r0 = unpack8h<F16>(rh[c4][cb]);
#pragma unroll
                  for (int c = 0; c < 8; ++c) a[c] += r0.v[c];
                  if (out_x2) {
                    const F8 r1 = unpack8h<F16>(rl[c4][cb]);
#pragma unroll
                    for (int c = 0; c < 8; ++c) a[c] += r1.v[c];
                  }
                }
                if (p.relu) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) a[c] = fmaxf(a[c], 0.f);
                }
                F8 f;
#pragma unroll
                for (int c = 0; c < 8; ++c) f.v[c] = a[c];
                const uint4 hi = pack8h<F16>(f);
                {  // a half whose exponent field is all ones: the value left the IEEE-half range (or was NaN)
                  const uint32_t mm = ((hi.x & 0x7fff7fffu) + 0x04000400u) | ((hi.y & 0x7fff7fffu) + 0x04000400u) |
                                      ((hi.z & 0x7fff7fffu) + 0x04000400u) | ((hi.w & 0x7fff7fffu) + 0x04000400u);
                  bad |= (mm & 0x80008000u) != 0;
                }
                const int64_t bo = (int64_t)cb * blk_elems;
                if (!p.skip_y) *reinterpret_cast<uint4 *>(p.y + onat + bo) = hi;
                if (p.y_split) *reinterpret_cast<uint4 *>(p.y_split + ospl + bo) = hi;
                if (out_x2) {
                  const F8 h = unpack8h<F16>(hi);
#pragma unroll
                  for (int c = 0; c < 8; ++c) f.v[c] -= h.v[c];
                  const uint4 lo = pack8h<F16>(f);
                  if (!p.skip_y) *reinterpret_cast<uint4 *>(p.y + onat + bo + lo_off) = lo;
                  if (p.y_split) *reinterpret_cast<uint4 *>(p.y_split + ospl + bo + lo_off) = lo;
                }
              }
            }
            if (bad && p.range_flag) *p.range_flag = 1;
          }
        } else if (MODE == M_DEC && X2) {
          // split-precision pass: one parity class (32 accumulator columns, column block = pw*2 + ph) at a time; its
          // partial / residual operands are requested as one batch before the TMEM read
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int ph = c4 >> 1, pw = c4 & 1;
            constexpr int NCB = NT / 8;  // channel blocks of this CTA's slice
            const int64_t pos = ((int64_t)qo * p.Ho + 2 * hr + ph) * p.Wo + 2 * wr + pw;
            const int64_t sidx = ((int64_t)(qo >> 1) * (p.Ho / 2) + hr) * (p.Wo / 2) + wr;
            const int cls = (qo & 1) * 4 + ph * 2 + pw;
            XPre dq[NCB];
            if (valid) {
#pragma unroll
              for (int cb = 0; cb < NCB; ++cb) xload(dq[cb], cb, pos, cls, sidx);
            }
            uint32_t v[32];
            const uint32_t t0 = tmem_base + lane_addr + r * C::ACC_COLS + (pw * 2 + ph) * NT;
            if (!(p.dbg & 8)) {
              if (NT == 16) {
                uint32_t v16[16];
                ptx::tmem_ld_32x16(t0, v16);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = v16[i];
              } else {
                ptx::tmem_ld_32x32(t0, v);
                ptx::tmem_ld_wait();
              }
            }
            if (!(p.dbg & 16)) { if (NT == 16) ptx::tmem_st_32x16(t0, zero); else ptx::tmem_st_32x32(t0, zero); }
            if (c4 == 3) {
              ptx::tmem_st_wait();
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(acce_bar(r));
            }
            if (!valid || (p.dbg & 4)) continue;
#pragma unroll
            for (int cb = 0; cb < NCB; ++cb) {
              F8 a8;
#pragma unroll
              for (int c = 0; c < 8; ++c) a8.v[c] = __uint_as_float(v[cb * 8 + c]);
              finish(a8, cb, pos, cls, sidx, resv[0], dq[cb]);
            }
          }
        } else if (MODE == M_DEC) {
          // class = pw*2 + ph.  The two pw classes of one ph are neighbouring output voxels (2w, 2w+1): drain both and
          // store 32 contiguous bytes per thread and channel block (a lone 16-byte store half-fills its 32 B sector).
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            uint32_t v0[32], v1[32];
            const uint32_t t0 = tmem_base + lane_addr + r * C::ACC_COLS + ph * NT, t1 = t0 + 2 * NT;
            if (!(p.dbg & 8)) { ptx::tmem_ld_32x32(t0, v0); ptx::tmem_ld_32x32(t1, v1); ptx::tmem_ld_wait(); }
            if (!(p.dbg & 16)) { ptx::tmem_st_32x32(t0, zero); ptx::tmem_st_32x32(t1, zero); }
            if (ph == 1) {
              ptx::tmem_st_wait();
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(acce_bar(r));
            }
            if (!valid || (p.dbg & 4)) continue;
            const int64_t pos = ((int64_t)qo * p.Ho + 2 * hr + ph) * p.Wo + 2 * wr;
            const int64_t sidx = ((int64_t)(qo >> 1) * (p.Ho / 2) + hr) * (p.Wo / 2) + wr;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
              F8 a8, b8;
#pragma unroll
              for (int c = 0; c < 8; ++c) { a8.v[c] = __uint_as_float(v0[cb * 8 + c]); b8.v[c] = __uint_as_float(v1[cb * 8 + c]); }
              finish(a8, cb, pos, (qo & 1) * 4 + ph * 2, sidx, resv[ph * 8 + cb * 2], xq[0]);
              finish(b8, cb, pos + 1, (qo & 1) * 4 + ph * 2 + 1, sidx, resv[ph * 8 + cb * 2 + 1], xq[0]);
            }
          }
        } else {
          constexpr int NV = NT < 32 ? NT : 32;      // columns per TMEM load
          constexpr int NLD = NT / NV;               // loads per plane (NT = 64 -> 2)
          const uint32_t taddr = tmem_base + lane_addr + r * C::ACC_COLS;
          const int64_t pos = ((int64_t)qo * p.Ho + hr) * p.Wo + wr;
          const int64_t sidx = ((int64_t)(qo >> 1) * (p.Ho / 2) + (hr >> 1)) * (p.Wo / 2) + (wr >> 1);
#pragma unroll
          for (int part = 0; part < NLD; ++part) {
            uint32_t v[32];
            if (p.dbg & 8) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0u;
            } else if (NT == 16) {
              uint32_t v16[16];
              ptx::tmem_ld_32x16(taddr, v16);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v16[i];
#pragma unroll
              for (int bk = 1; bk < C::NB; ++bk) {  // sum the accumulator banks (fp32, round to nearest)
                ptx::tmem_ld_32x16(taddr + bk * C::BANK_COLS, v16);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v16[i]));
              }
            } else {
              ptx::tmem_ld_32x32(taddr + part * 32, v);
              if (C::NB > 1) {  // banks 0 and 1 with ONE round trip (the drain latency bounds how few ring slots suffice)
                uint32_t u[32];
                ptx::tmem_ld_32x32(taddr + C::BANK_COLS + part * 32, u);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
              } else {
                ptx::tmem_ld_wait();
              }
#pragma unroll
              for (int bk = 2; bk < C::NB; ++bk) {
                uint32_t u[32];
                ptx::tmem_ld_32x32(taddr + bk * C::BANK_COLS + part * 32, u);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
              }
            }
            if (!(p.dbg & 16)) {  // leave the slot zeroed for its next output plane
#pragma unroll
              for (int bk = 0; bk < C::NB; ++bk) {
                if (NT == 16) ptx::tmem_st_32x16(taddr + bk * C::BANK_COLS, zero); else ptx::tmem_st_32x32(taddr + bk * C::BANK_COLS + part * 32, zero);
              }
            }
            if (part == NLD - 1) {
              ptx::tmem_st_wait();
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(acce_bar(r));
            }
            if (!valid || (p.dbg & 4)) continue;
            if (p.y1) {  // 32->1 classifier head: channel 0 only, f32, running sum fused (stackhourglass.py:142-144)
              const int64_t o1 = (int64_t)n * Vo + pos;
              p.y1[o1] = __uint_as_float(v[0]) + (p.y1_cols == 2 ? __uint_as_float(v[1]) : 0.f) + (p.res1 ? p.res1[o1] : 0.f);
              continue;
            }
#pragma unroll
            for (int cbl = 0; cbl < NV / 8; ++cbl) {
              const int cb = part * 4 + cbl;  // channel block inside this CTA's NT-wide slice
              F8 r8;
#pragma unroll
              for (int c = 0; c < 8; ++c) r8.v[c] = __uint_as_float(v[cbl * 8 + c]);
              if (X2 && NXQ == 1 && NT > 8) {  // (64-wide blocks: no batch, operands requested at use)
                XPre one;
                xload(one, cb, pos, (qo & 1) * 4 + (hr & 1) * 2 + (wr & 1), sidx);
                finish(r8, cb, pos, (qo & 1) * 4 + (hr & 1) * 2 + (wr & 1), sidx, resv[cb], one);
              } else {
                finish(r8, cb, pos, (qo & 1) * 4 + (hr & 1) * 2 + (wr & 1), sidx, resv[cb], xq[cb < NXQ ? cb : 0]);
              }
            }
          }
        }
      }
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (C::S2T && p.cluster) ptx::cluster_sync();   // (a CTA must not exit while its peer's commits / boxes may still be addressed to it)
  if (warp == 2) ptx::tmem_dealloc<C::TCOLS>(tmem_base);
}

// natural blocked layout -> 8 parity sub-volumes: [N*C/8][D][H][W][8] -> [N*C/8][pd*4+ph*2+pw][D/2][H/2][W/2][8].
// One thread moves the voxel pair (2w2, 2w2+1): a fully used 32-byte read, two 16-byte writes that are contiguous
// across the warp inside their (pw) sub-volume rows.
__global__ void __launch_bounds__(256)
space_to_depth_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int64_t nblk, int D, int H, int W)
{
  const int D2 = D / 2, H2 = H / 2, W2 = W / 2;
  const int64_t sub = (int64_t)D2 * H2 * W2;
  const int64_t total = nblk * D * H * W2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int w2 = (int)(idx % W2);
    const int h = (int)((idx / W2) % H);
    const int d = (int)((idx / ((int64_t)W2 * H)) % D);
    const int64_t b = idx / ((int64_t)W2 * H * D);
    const uint4 *s2 = src + ((b * D + d) * H + h) * W + 2 * w2;
    const uint4 v0 = __ldg(s2), v1 = __ldg(s2 + 1);
    const int cls = (d & 1) * 4 + (h & 1) * 2;
    const int64_t o = (b * 8 + cls) * sub + ((int64_t)(d >> 1) * H2 + (h >> 1)) * W2 + w2;
    dst[o] = v0;
    dst[o + sub] = v1;
  }
}

#undef A_KOFF
#undef B_KS
#undef BANK

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

static int env_flag(const char *name)
{
  const char *e = getenv(name);
  return (e && e[0] == '1') ? 1 : 0;
}

static int mode_of(int kind) { return kind == IDISP_CONV_S1 ? M_S1 : (kind == IDISP_CONV_S2 ? M_S2 : M_DEC); }

// output channels per stacked block for a layer (see Cfg): 16 for the 32->1 heads, 64 for the stride-2 32->64 conv
static int nt_of(int kind, int cin, int cout)
{
  if (env_flag("IDISP_TC_NT32")) return 32;  // A/B switch
  if (kind == IDISP_CONV_S1 && cout == 1) return 16;
  if (kind == IDISP_CONV_S2 && cin == 32 && cout == 64) return 64;
  return 32;
}

}  // namespace tc

// Pack [27][cin][cout] f32 (tap = (kd*3+kh)*3+kw, BN scale folded) into the per-mode UMMA B layout, bf16:
// rows of 8 input channels (16 B), 8-row core matrices contiguous (SBO 128 B), K cores LBO apart.
int tc_weights_prepare(const float *w_tap, int kind, int cin, int cout, int f16, TcWeights &out, cudaStream_t s, int words, int nt, int ncat)
{
  tc_weights_free(out);
  out.kind = kind; out.cin = cin; out.cout = cout; out.f16 = f16; out.words = words; out.ncat = ncat;
  if (ncat && !(cout == 1 && f16 && words == 1)) { set_error("tc_weights_prepare: the lo word goes to output column 1 only for the 1-channel head"); return IDISP_ERR_INVALID; }
  if (words != 1 && !(words == 2 && f16)) { set_error("tc_weights_prepare: two-word weights are IEEE half only"); return IDISP_ERR_INVALID; }
  if (!tc_supported(kind, cin, cout, 4, 16, 16)) return IDISP_OK;  // layer stays on the SIMT kernel
  const int NT = nt > 0 ? nt : tc::nt_of(kind, cin, cout);
  out.nt = NT;
  // Cfg::MRG layout (two-word stride-1 weights, 32-wide blocks): per (tap, k-step) ONE chunk [2 kcores][hi 3*NT | lo 3*NT][8]
  const bool mrg = IDISP_MRG && IDISP_TRI && words == 2 && kind == IDISP_CONV_S1 && NT == 32 && cin == 32;
  // Cfg::S2T layout (two-word stride-2 weights, Cin 32, 32-wide blocks): per (kh, kw, k-step) ONE chunk of 192 rows
  // [lo kd1 | hi kd1 | lo kd2 | hi kd2 | hi kd0 | lo kd0]
  const bool s2t = IDISP_S2T && IDISP_TRI && words == 2 && kind == IDISP_CONV_S2 && NT == 32 && cin == 32;
  const bool dtr = IDISP_DTR && words == 2 && kind == IDISP_DECONV_S2 && NT == 16;   // Cfg::DTR layout (two-word transposed conv, 16-wide blocks)
  const int KS = words * cin / 16, NH = (cout + NT - 1) / NT;  // k-steps per tap: the lo word's follow the hi word's
  const size_t per_nh = (size_t)27 * KS * NT * 16;  // bf16 elements
  // 16-bit storage words (bf16 or IEEE half, same size): convert through cvt()
  auto cvt = [f16](float v) -> __nv_bfloat16 {
    if (!f16) return __float2bfloat16_rn(v);
    const __half hh = __float2half_rn(v);
    __nv_bfloat16 o;
    memcpy(&o, &hh, 2);
    return o;
  };
  std::vector<__nv_bfloat16> h(NH * per_nh, cvt(0.f));
  auto wv = [&](int kd, int kh, int kw, int ci, int co) -> float {
    if (ncat && co == 1) {  // 1-channel head, split precision: column 1 = what column 0's half rounded away
      const float v = w_tap[((size_t)((kd * 3 + kh) * 3 + kw) * cin + ci) * cout];
      return v - __half2float(__float2half_rn(v));
    }
    if (co >= cout) return 0.f;
    const float v = w_tap[((size_t)((kd * 3 + kh) * 3 + kw) * cin + ci % cin) * cout + co];
    return ci < cin ? v : v - __half2float(__float2half_rn(v));  // second word: what the first one rounded away
  };
  for (int nh = 0; nh < NH; ++nh) {
    __nv_bfloat16 *base = h.data() + nh * per_nh;
    if (kind == IDISP_CONV_S1 || kind == IDISP_CONV_S2) {
      // chunk (kh,kw,ks): [2 kcores][3 blocks x NT couts][8]; block j <-> kd = S1 {2,1,0}, S2 {2,0,1}
      static const int kd_s1[3] = {2, 1, 0}, kd_s2[3] = {2, 0, 1};
      const int *kdj = kind == IDISP_CONV_S1 ? kd_s1 : kd_s2;
      for (int t2 = 0; t2 < 9; ++t2)
        for (int ks = 0; ks < KS; ++ks)
          for (int kc = 0; kc < 2; ++kc)
            for (int n = 0; n < 3 * NT; ++n)
              for (int e = 0; e < 8; ++e) {
                // (ks >= cin/16: the lo word's k-steps; wv() turns channel index cin + c into the lo word of channel c)
                const int ksw = (mrg || s2t) ? ks % (cin / 16) : ks, word = (mrg || s2t) ? ks / (cin / 16) : 0;
                size_t dst = mrg ? ((((size_t)t2 * (cin / 16) + ksw) * 2 + kc) * 6 * NT + word * 3 * NT + n) * 8 + e
                                 : ((((size_t)t2 * KS + ks) * 2 + kc) * 3 * NT + n) * 8 + e;
                if (s2t) {
                  const int kd = kdj[n / NT];   // row group of (kd, word): kd1 -> {lo 0, hi 1}; kd2 -> {lo 2, hi 3}; kd0 -> {hi 4, lo 5}
                  const int grp = kd == 1 ? (word ? 0 : 1) : (kd == 2 ? (word ? 2 : 3) : (word ? 5 : 4));
                  dst = ((((size_t)t2 * (cin / 16) + ksw) * 2 + kc) * 6 * NT + grp * NT + n % NT) * 8 + e;
                }
                base[dst] = cvt(wv(kdj[n / NT], t2 / 3, t2 % 3, ks * 16 + kc * 8 + e, nh * NT + n % NT));
              }
    } else {
      // DECONV: [kd][ks][2 kcores][9*NT rows][8]; rows = entries e0..e4 (tc::dec_*), blocks = output classes
      //   class = pw*2+ph; per axis: p=0 -> k=1 (shift 0); p=1 -> k=2 (shift 0), k=0 (shift 1)
      struct Blk { int kh, kw; };
      static const Blk blocks[9] = {
          {1, 1}, {2, 1}, {1, 2}, {2, 2},  // e0 shift(0,0): classes 0 (ph0,pw0), 1 (ph1,pw0), 2 (ph0,pw1), 3 (ph1,pw1)
          {1, 0}, {2, 0},                  // e1 shift(0,1): classes 2, 3  (pw=1 -> kw=0)
          {0, 1},                          // e2 shift(1,0): class 1       (ph=1 -> kh=0, pw=0 -> kw=1)
          {0, 2},                          // e3 shift(1,0): class 3       (kh=0, pw=1 with shift_w 0 -> kw=2)
          {0, 0}};                         // e4 shift(1,1): class 3
      if (dtr) {
        // Cfg::DTR layout: [ks][2 kcores][27*NT rows][8]; rows = (input shift, class, kd, cout) with the classes in the order
        // c00 | c10 | c11 | c01 ((ph,pw) of the output voxel) and, inside a class, the three kd (output planes 2z-1, 2z, 2z+1)
        static const Blk dblocks[9] = {
            {1, 1}, {2, 1}, {2, 2}, {1, 2},  // shift (0,0): c00, c10, c11, c01
            {2, 0}, {1, 0},                  // shift (0,1): c11, c01  (pw = 1 -> kw = 0 reads in[w+1])
            {0, 1}, {0, 2},                  // shift (1,0): c10, c11  (ph = 1 -> kh = 0 reads in[h+1])
            {0, 0}};                         // shift (1,1): c11
        for (int ks = 0; ks < KS; ++ks)
          for (int kc = 0; kc < 2; ++kc)
            for (int row = 0; row < 27 * NT; ++row)
              for (int e = 0; e < 8; ++e) {
                const Blk b = dblocks[row / (3 * NT)];
                const int kd = (row / NT) % 3;
                base[(((size_t)ks * 2 + kc) * 27 * NT + row) * 8 + e] = cvt(wv(kd, b.kh, b.kw, ks * 16 + kc * 8 + e, nh * NT + row % NT));
              }
      } else
      for (int kd = 0; kd < 3; ++kd)
        for (int ks = 0; ks < KS; ++ks)
          for (int kc = 0; kc < 2; ++kc)
            for (int row = 0; row < 9 * NT; ++row)
              for (int e = 0; e < 8; ++e) {
                const Blk b = blocks[row / NT];
                base[((((size_t)kd * KS + ks) * 2 + kc) * 9 * NT + row) * 8 + e] =
                    cvt(wv(kd, b.kh, b.kw, ks * 16 + kc * 8 + e, nh * NT + row % NT));
              }
    }
  }
  out.bytes = h.size() * sizeof(__nv_bfloat16);
  IDISP_CUDA(cudaMalloc(&out.dev, out.bytes));
  IDISP_CUDA(cudaMemcpyAsync(out.dev, h.data(), out.bytes, cudaMemcpyHostToDevice, s));
  IDISP_CUDA(cudaStreamSynchronize(s));
  return IDISP_OK;
}

void tc_weights_free(TcWeights &w)
{
  if (w.dev) cudaFree(w.dev);
  w.dev = nullptr; w.bytes = 0;
}

bool tc_supported(int kind, int cin, int cout, int D, int H, int W)
{
  static int disabled = -1, only_s1 = -1;
  if (disabled < 0) { disabled = tc::env_flag("IDISP_TC_DISABLE"); only_s1 = tc::env_flag("IDISP_TC_ONLY_S1"); }
  if (disabled || D < 1 || H < 1 || W < 1) return false;
  if (kind == IDISP_CONV_S1) return (cin == 32 || cin == 64) && (cout == 32 || cout == 64 || (cout == 1 && !only_s1));
  if (only_s1) return false;
  if (kind == IDISP_CONV_S2) return (cin == 32 || cin == 64) && (cout == 32 || cout == 64) && D % 2 == 0 && H % 2 == 0 && W % 2 == 0;
  if (kind == IDISP_DECONV_S2) return cin == 64 && (cout == 32 || cout == 64);
  return false;
}

size_t tc_scratch_bytes(int kind, int B, int cin, int D, int H, int W)
{
  return kind == IDISP_CONV_S2 ? (size_t)B * cin * D * H * W * sizeof(__nv_bfloat16) : 0;
}

template <int CIN, int MODE, int OCC, int NT>
static int tc_launch(const TcWeights &w, const __nv_bfloat16 *x, int B, int D, int H, int W, int Cout, const float *bias,
                     const __nv_bfloat16 *residual, int relu, __nv_bfloat16 *y, const float *res1, float *y1, void *scratch,
                     int x_is_split, __nv_bfloat16 *y_split, const TcCostVolume *cv, const TcOpts &opts, cudaStream_t s)
{
  using C = tc::Cfg<CIN, MODE, OCC, NT>;
  const int aw = opts.xp ? 2 : 1;  // activation words one TMA box fetches (hi|lo block groups are adjacent in memory)
  const int blk_stride = opts.in_blk_stride > 0 ? opts.in_blk_stride : (cv && cv->view < 0 ? C::CBLK / 2 : C::CBLK);  // blocks per input sample (and view)
  using MC = tc::ModeCfg<MODE>;
  tc::EncodeTiledFn enc = tc::get_encode();
  if (!enc) { set_error("tc_conv3d: cuTensorMapEncodeTiled not available from the driver"); return IDISP_ERR_CUDA; }
  CUtensorMap map, rmap;
  const int fmt = !w.f16 ? 0 : (opts.xp ? 2 + opts.xp : ((opts.x2 || opts.part_in || opts.part_out) ? 2 : 1));
  // rows of a sub-tile box: Cfg::SUB_H of the kernel variant this launch selects (the S2T form reads 17 rows, see Cfg)
  const int sub_h = (IDISP_S2T && IDISP_TRI && MODE == tc::M_S2 && CIN == 32 && NT == 32 && OCC == 1 && fmt == 3) ? tc::TH + 1 : MC::SUB_H;
  static thread_local tc::CvMaps<true> cvmaps;  // ~8 KB of kernel parameters, encoded per launch (host-only work)
  CUresult r;
  const void *src = x;
  if (MODE == tc::M_S2) {
    if (x_is_split & 1) {
      src = x;  // the producer's epilogue already wrote the parity sub-volumes (Params::y_split)
    } else {
      if (!scratch) { set_error("tc_conv3d: stride-2 layer needs the space-to-depth scratch buffer"); return IDISP_ERR_INVALID; }
      const int64_t nblk = (int64_t)B * blk_stride, total = nblk * D * H * (W / 2);
      const int64_t want = ceil_div64(total, 256);
      tc::space_to_depth_kernel<<<(int)(want < 148 * 32 ? want : 148 * 32), 256, 0, s>>>((const uint4 *)x, (uint4 *)scratch, nblk, D, H, W);
      IDISP_LAUNCH_CHECK();
      src = scratch;
    }
    const int D2 = D / 2, H2 = H / 2, W2 = W / 2;
    const cuuint64_t dims[5] = {(cuuint64_t)W2 * 8, (cuuint64_t)H2, (cuuint64_t)D2, 8, (cuuint64_t)B * blk_stride};
    const cuuint64_t strides[4] = {(cuuint64_t)W2 * 16, (cuuint64_t)H2 * W2 * 16, (cuuint64_t)D2 * H2 * W2 * 16, (cuuint64_t)8 * D2 * H2 * W2 * 16};
    const cuuint32_t box[5] = {8 * MC::SUB_W, (cuuint32_t)sub_h, 1, 1, (cuuint32_t)(aw * C::CBLK)};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (opts.in_lo_off) {
      // K-split launch (channels [in_blk_off*8, +32) of a 64-channel tensor, hi group and lo group in_lo_off blocks apart): dimensions
      // (8ch*W2, H2, class*D2 + depth, block of the group, sample*2 + word); the channel half is the base pointer
      if (blk_stride != 2 * opts.in_lo_off) { set_error("tc_conv3d: K-split stride-2 input must be [hi blocks | lo blocks] per sample"); return IDISP_ERR_INVALID; }
      const cuuint64_t sub16 = (cuuint64_t)D2 * H2 * W2 * 16;
      const cuuint64_t kdims[5] = {(cuuint64_t)W2 * 8, (cuuint64_t)H2, (cuuint64_t)8 * D2, (cuuint64_t)C::CBLK, (cuuint64_t)B * 2};
      const cuuint64_t kstrides[4] = {(cuuint64_t)W2 * 16, (cuuint64_t)H2 * W2 * 16, 8 * sub16, (cuuint64_t)opts.in_lo_off * 8 * sub16};
      const cuuint32_t kbox[5] = {8 * MC::SUB_W, (cuuint32_t)sub_h, 1, (cuuint32_t)C::CBLK, 2};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (char *)const_cast<void *>(src) + (size_t)opts.in_blk_off * 8 * sub16, kdims, kstrides, kbox, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (cv) {
    // fused cost volume: x is unused; one 4-D map per plane over {left, right} x [B*C/8] x [Hf] x [(Wf-|i|)*8]
    const bool one_view = cv->view >= 0;
    const int halfblk = one_view ? C::CBLK : C::CBLK / 2;   // channel blocks of one view (per precision word)
    if (D > tc::CV_MAX_PLANES) { set_error("tc_conv3d: fused cost volume supports at most %d planes", tc::CV_MAX_PLANES); return IDISP_ERR_INVALID; }
    const int64_t lr_bytes = (const char *)cv->right - (const char *)cv->left;
    if (lr_bytes <= 0 || lr_bytes % 16) { set_error("tc_conv3d: right features must follow the left ones in memory (16 B aligned)"); return IDISP_ERR_INVALID; }
    const cuuint32_t box[4] = {8 * MC::SUB_W, MC::SUB_H, (cuuint32_t)(aw * halfblk), one_view ? 1u : 2u};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    r = CUDA_SUCCESS;
    for (int k = 0; r == CUDA_SUCCESS && k < D; ++k) {
      const int i = k + cv->shift0, ai = i < 0 ? -i : i;
      const int wk = ai >= W ? 1 : W - ai;  // fully masked planes get a 1-voxel map that the kernel reads far out of range
      const int lo = ai >= W ? 0 : (i > 0 ? i : 0), ro = ai >= W ? 0 : (i < 0 ? -i : 0);
      const cuuint64_t dims[4] = {(cuuint64_t)wk * 8, (cuuint64_t)H, (cuuint64_t)B * blk_stride, 2};
      const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)(lr_bytes + (int64_t)(ro - lo) * 16)};
      r = enc(&cvmaps.m[k], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16 *>(cv->left) + (size_t)lo * 8, dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r == CUDA_SUCCESS) rmap = cvmaps.m[0];
    map = rmap;
  } else {
    // (8 ch, W) are contiguous in the blocked layout -> ONE tensor dimension of 8*W elements: a box row is SUB_W voxels x 16 B
    const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * blk_stride};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
    const cuuint32_t box[4] = {8 * MC::SUB_W, MC::SUB_H, 1, (cuuint32_t)((opts.in_lo_off ? 1 : aw) * C::CBLK)};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { set_error("tc_conv3d: cuTensorMapEncodeTiled failed (%d) for dims W=%d H=%d D=%d mode=%d", (int)r, W, H, D, MODE); return IDISP_ERR_CUDA; }
  tc::Params p;
  p.w = (const __nv_bfloat16 *)w.dev; p.bias = bias; p.residual = residual; p.y = y; p.res1 = res1; p.y1 = y1; p.y_split = y_split; p.residual_is_split = (x_is_split >> 1) & 1; p.skip_y = (x_is_split >> 2) & 1;
  p.B = B; p.Din = D; p.Cout = Cout; p.relu = relu; p.y1_cols = w.ncat ? 2 : 1;
  p.cv_shift0 = cv ? cv->shift0 : 0; p.cv_view = cv && cv->view > 0 ? 1 : 0;
  p.part_in = opts.part_in; p.part_out = opts.part_out; p.x2 = opts.x2; p.in_blk_stride = blk_stride; p.in_blk_off = opts.in_blk_off; p.in_lo_off = opts.in_lo_off;
  p.range_flag = opts.range_flag;
  if (!cv) rmap = map;
  p.cluster = 0;
  p.x_split = nullptr;
  if (opts.x_copy_split) {
    if (!(MODE == tc::M_S1 && CIN == 32 && NT == 32 && OCC == 1 && fmt == 3 && IDISP_MRG && IDISP_TRI && Cout == 32 && !cv && opts.x2 && !opts.in_lo_off && blk_stride == 8)) {
      set_error("tc_conv3d: the input side copy exists in the stride-1 32 -> 32 split-precision kernel only");
      return IDISP_ERR_INVALID;
    }
    if (D % 2 || H % 2 || W % 2) { set_error("tc_conv3d: parity-split copy needs even dims"); return IDISP_ERR_INVALID; }
    p.x_split = opts.x_copy_split;
  }
  p.res_map = 0;
  if (MODE == tc::M_DEC && NT == 16 && fmt == 3 && IDISP_DTR && residual) {
    // Cfg::DTR: a second map over the residual tensor (output-sized, 2 * Cout/8 blocks per sample) for the producer's L2 prefetches:
    // one box = the output-plane tile of one CTA column (16 x 32 voxels) x this CTA's two channel blocks of one precision word
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    const cuuint64_t nblk = (cuuint64_t)B * (opts.x2 ? 2 : 1) * (Cout / 8);
    const cuuint32_t estr5[5] = {1, 1, 1, 1, 1};
    CUresult rr;
    if ((x_is_split >> 1) & 1) {  // parity layout [blk][8 classes][Do/2][Ho/2][Wo/2][8]
      const cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, 8, nblk};
      const cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16, (cuuint64_t)8 * D * H * W * 16};
      const cuuint32_t box[5] = {8 * tc::TW, tc::TH, 1, 4, 2};
      rr = enc(&rmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16 *>(residual), dims, strides, box, estr5, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t dims[4] = {(cuuint64_t)Wo * 8, (cuuint64_t)Ho, (cuuint64_t)Do, nblk};
      const cuuint64_t strides[3] = {(cuuint64_t)Wo * 16, (cuuint64_t)Ho * Wo * 16, (cuuint64_t)Do * Ho * Wo * 16};
      const cuuint32_t box[4] = {16 * tc::TW, 2 * tc::TH, 1, 2};
      rr = enc(&rmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16 *>(residual), dims, strides, box, estr5, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (rr == CUDA_SUCCESS) p.res_map = 1; else rmap = map;   // (prefetching is an optimisation: a shape the map cannot express just goes without)
  }
  { static int dbg = -1; if (dbg < 0) { const char *e = getenv("IDISP_TC_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  if (MODE == tc::M_S1) { p.Dout = D; p.Ho = H; p.Wo = W; p.Hr = H; p.Wr = W; }
  if (MODE == tc::M_S2) { p.Dout = D / 2; p.Ho = H / 2; p.Wo = W / 2; p.Hr = H / 2; p.Wr = W / 2; }
  if (MODE == tc::M_DEC) { p.Dout = 2 * D; p.Ho = 2 * H; p.Wo = 2 * W; p.Hr = H; p.Wr = W; }
  p.tiles_h = ceil_div(p.Hr, tc::TH); p.tiles_w = ceil_div(p.Wr, tc::TW); p.nh = (Cout + NT - 1) / NT;
  const int ncols = B * p.tiles_h * p.tiles_w;
  int dev = 0;
  cudaGetDevice(&dev);
  static int sm_count[64];  // per device, queried once
  if (dev < 0 || dev >= 64) { set_error("tc_conv3d: device ordinal %d out of range", dev); return IDISP_ERR_INVALID; }
  if (!sm_count[dev]) cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
  const int sms = sm_count[dev] > 0 ? sm_count[dev] : 148;
  int per_slice = sms * OCC / p.nh;
  if (per_slice > ncols) per_slice = ncols;
  const int grid_cols = per_slice * p.nh;
  auto go = [&](auto fmt_c, auto cv_c) -> int {
    constexpr int FMT = decltype(fmt_c)::value;
    constexpr int CVK = decltype(cv_c)::value;
    constexpr int XP = FMT == 3 ? 1 : (FMT == 4 ? 2 : 0);
    // in-launch K concatenation exists where its shared-memory budget closes (see Cfg)
    constexpr bool ok = (XP == 0 && (MODE != tc::M_DEC || NT == 32)) || (XP == 1 && CIN == 32 && OCC == 1 && NT <= 32 && MODE != tc::M_DEC) ||
                        (XP == 1 && MODE == tc::M_DEC && NT == 16) || (XP == 2 && (CIN == 64 || NT == 16) && OCC == 1 && MODE != tc::M_S2);
    // (the one-view cost-volume loader exists for the launches that use it: the two halves of the K-split first layer)
    constexpr bool ok1 = CVK != 2 || (CIN == 32 && FMT == 3 && NT == 32 && OCC == 1 && MODE == tc::M_S1);
    if constexpr (!ok || !ok1) {
      set_error("tc_conv3d: K-concatenation mode %d not built for Cin=%d mode=%d", XP, CIN, MODE);
      return IDISP_ERR_UNSUPPORTED;
    } else {
      using CX = tc::Cfg<CIN, MODE, OCC, NT, (FMT < 2 ? 0 : (FMT == 2 ? 3 : FMT - 2))>;
      auto kern = tc::conv3d_tc_kernel<CIN, MODE, OCC, CVK, NT, FMT>;
      static bool smem_opt_in[64];  // per kernel instantiation and device: the opt-in is sticky, set it once
      if (!smem_opt_in[dev]) {
        IDISP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CX::SMEM));
        smem_opt_in[dev] = true;
      }
      // Kernels that walk work items (Items) get one CTA per SM even when there are fewer columns than that: the columns are then
      // split in depth so that a small batch (the live call: 28 columns per ROI pair) still fills the machine.
      static const int no_fill = tc::env_flag("IDISP_NO_DEPTH_FILL");   // A/B switch
      const int grid = ((CX::MRG || CX::S2T || CX::DTR) && !no_fill && !(p.dbg & 2048)) ? (sms * OCC / p.nh) * p.nh : grid_cols;
      const bool pdl = (long long)ncols * p.nh <= 4ll * sms;   // short launches only (see launch_ex): at most four rounds of columns
      if constexpr (CVK != 0) IDISP_CUDA(launch_ex(kern, grid, CX::NTHREADS, CX::SMEM, s, false, pdl, map, rmap, cvmaps, p));
      else if constexpr (CX::S2T) {
        static const int no_cluster = tc::env_flag("IDISP_NO_CLUSTER");   // A/B switch
        tc::Params pc = p;
        pc.cluster = (p.nh == 2 && grid % 2 == 0 && !no_cluster) ? 1 : 0;
        IDISP_CUDA(launch_ex(kern, grid, CX::NTHREADS, CX::SMEM, s, pc.cluster != 0, pdl, map, rmap, tc::CvMaps<false>{}, pc));
      }
      else IDISP_CUDA(launch_ex(kern, grid, CX::NTHREADS, CX::SMEM, s, false, pdl, map, rmap, tc::CvMaps<false>{}, p));
      return IDISP_OK;
    }
  };
  auto by_fmt = [&](auto cv_c) -> int {
    using std::integral_constant;
    switch (fmt) {
      case 0: return go(integral_constant<int, 0>{}, cv_c);
      case 1: return go(integral_constant<int, 1>{}, cv_c);
      case 2: return go(integral_constant<int, 2>{}, cv_c);
      case 3: return go(integral_constant<int, 3>{}, cv_c);
      default: return go(integral_constant<int, 4>{}, cv_c);
    }
  };
  int lrc = IDISP_OK;
  bool launched = false;
  if constexpr (MODE == tc::M_S1) {
    if (cv) { lrc = cv->view >= 0 ? by_fmt(std::integral_constant<int, 2>{}) : by_fmt(std::integral_constant<int, 1>{}); launched = true; }
  }
  if (!launched) lrc = by_fmt(std::integral_constant<int, 0>{});
  if (lrc) return lrc;
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

int tc_conv3d(const TcWeights &w, const __nv_bfloat16 *x, int B, int Cin, int D, int H, int W, int Cout, int kind,
              const float *bias, const __nv_bfloat16 *residual, int relu, __nv_bfloat16 *y, const float *res1, float *y1,
              void *scratch, int x_is_split, __nv_bfloat16 *y_split, const TcCostVolume *cv, cudaStream_t s, const TcOpts *optsp)
{
  const TcOpts opts = optsp ? *optsp : TcOpts();
  if (cv && (kind != IDISP_CONV_S1 || !cv->left || !cv->right)) { set_error("tc_conv3d: bad fused cost-volume arguments"); return IDISP_ERR_INVALID; }
  if (opts.in_lo_off && (kind == IDISP_DECONV_S2 || cv || !opts.xp)) { set_error("tc_conv3d: a separate lo-group offset exists for (strided) convolution K-concatenation launches only"); return IDISP_ERR_INVALID; }
  if (opts.xp < 0 || opts.xp > 2 || w.words != (opts.xp == 1 ? 2 : 1) || (opts.xp && !w.f16)) { set_error("tc_conv3d: weights do not match K-concatenation mode %d", opts.xp); return IDISP_ERR_INVALID; }
  if (!tc_supported(kind, Cin, Cout, D, H, W) || !w.dev || w.cin != Cin || w.cout != Cout || w.kind != kind) {
    set_error("tc_conv3d: layer (kind=%d, %d->%d) not prepared for the tensor-core path", kind, Cin, Cout);
    return IDISP_ERR_INVALID;
  }
  if (y_split) {  // output dims must be even for the parity layout
    const int Do = kind == IDISP_DECONV_S2 ? 2 * D : (kind == IDISP_CONV_S2 ? D / 2 : D), Ho = kind == IDISP_DECONV_S2 ? 2 * H : (kind == IDISP_CONV_S2 ? H / 2 : H),
              Wo = kind == IDISP_DECONV_S2 ? 2 * W : (kind == IDISP_CONV_S2 ? W / 2 : W);
    if (Do % 2 || Ho % 2 || Wo % 2 || Cout == 1) { set_error("tc_conv3d: parity-split output needs even output dims"); return IDISP_ERR_INVALID; }
  }
  if ((Cout == 1) != (y1 != nullptr)) { set_error("tc_conv3d: the 1-channel head needs the f32 output (and only it)"); return IDISP_ERR_INVALID; }
  if (B == 0) return IDISP_OK;
#define IDISP_TC(CI, MD, OC, NTT) return tc_launch<CI, MD, OC, NTT>(w, x, B, D, H, W, Cout, bias, residual, relu, y, res1, y1, scratch, x_is_split, y_split, cv, opts, s)
  const int mode = tc::mode_of(kind);
  static const int occ1 = tc::env_flag("IDISP_TC_OCC1");  // A/B switch for the 2-CTA/SM stride-1 variant
  if (mode == tc::M_S1) {
    if (Cin == 64) IDISP_TC(64, tc::M_S1, 1, 32);
    const bool one = occ1 || opts.xp || opts.x2 || opts.part_in || opts.part_out;  // split-precision launches: one CTA per SM (two weight words, banked accumulators)
    if (w.nt == 16) { if (one) IDISP_TC(32, tc::M_S1, 1, 16); else IDISP_TC(32, tc::M_S1, 2, 16); }
    if (one) IDISP_TC(32, tc::M_S1, 1, 32); else IDISP_TC(32, tc::M_S1, 2, 32);
  }
  if (mode == tc::M_S2) {
    if (Cin == 64) IDISP_TC(64, tc::M_S2, 1, 32);
    if (w.nt == 64) IDISP_TC(32, tc::M_S2, 1, 64);
    IDISP_TC(32, tc::M_S2, 1, 32);
  }
  if (w.nt == 16) IDISP_TC(64, tc::M_DEC, 1, 16);
  IDISP_TC(64, tc::M_DEC, 1, 32);
#undef IDISP_TC
}

// ---------------------------------------------------------------------------------------
// split precision: x = x_hi + x_lo, w = w_hi + w_lo (IEEE half words), product = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo in fp32
// ---------------------------------------------------------------------------------------
// launches per layer: 1 where both weight words fit in shared memory next to a two-word input ring (Cin 32), 2 where only
// one does (Cin 64: the x_hi*w_lo term is a second launch chained through an fp32 partial), 3 otherwise (stride-2, Cin 64)
static int split_launches(int kind, int cin)
{
  static const int three = tc::env_flag("IDISP_X2_THREE_PASS");  // A/B switch
  if (three) return 3;
  if (cin == 32 && kind != IDISP_DECONV_S2) return 1;
  if (cin == 64 && kind == IDISP_DECONV_S2) return 1;  // 16-wide blocks: both weight words of a 16-channel slice fit
  if (cin == 64 && kind != IDISP_CONV_S2) return 2;
  return 3;
}

void tc_split_weights_free(TcSplitWeights &w) { tc_weights_free(w.hi); tc_weights_free(w.lo); tc_weights_free(w.both); tc_weights_free(w.k0); tc_weights_free(w.k1); }

// Stride-1 Cin = 64 layers, K split (default; IDISP_NO_KSPLIT=1 restores the term split).  Both weight words of all 64 input channels
// do not fit in shared memory (2 x 110 KB), so the layer was two launches split by TERM: x_hi*w_lo -> fp32 partial, then
// (x_hi, x_lo)*w_hi + partial (168 port cycles per tap and 16 input channels, two- or three-deep input ring).  Split by INPUT CHANNELS
// instead, each launch is a complete 32 -> Cout split-precision convolution of one channel half (both words resident, the merged
// x_hi*[w_hi | w_lo] MMA of Cfg::MRG: 152 port cycles, five-deep ring), chained through the same fp32 partial.
static bool ksplit_enabled()
{
  static const int off = tc::env_flag("IDISP_NO_KSPLIT");
  return !off;
}
static bool ksplit_layer(int kind, int cin, int cout)
{
  if (!IDISP_TRI || !ksplit_enabled() || cin != 64 || cout % 32) return false;
  // stride 2 (64 -> 64 at the hourglass bottom, three term-split launches before): two launches of the stride-2 32 -> 64 form (Cfg::S2T)
  return (kind == IDISP_CONV_S1 && IDISP_MRG) || (kind == IDISP_CONV_S2 && IDISP_S2T && cout == 64);
}

int tc_split_weights_prepare(const float *w_tap, int kind, int cin, int cout, TcSplitWeights &out, cudaStream_t s)
{
  tc_split_weights_free(out);
  const size_t n = (size_t)27 * cin * cout;
  std::vector<float> lo(n);
  for (size_t j = 0; j < n; ++j) lo[j] = w_tap[j] - __half2float(__float2half_rn(w_tap[j]));
  int rc;
  if ((rc = tc_weights_prepare(w_tap, kind, cin, cout, 1, out.hi, s))) return rc;
  if ((rc = tc_weights_prepare(lo.data(), kind, cin, cout, 1, out.lo, s))) return rc;
  if (ksplit_layer(kind, cin, cout)) {
    std::vector<float> half((size_t)27 * 32 * cout);
    for (int h = 0; h < 2; ++h) {
      for (int t = 0; t < 27; ++t)
        for (int c = 0; c < 32; ++c)
          memcpy(&half[((size_t)t * 32 + c) * cout], &w_tap[((size_t)t * cin + 32 * h + c) * cout], sizeof(float) * cout);
      if ((rc = tc_weights_prepare(half.data(), kind, 32, cout, 1, h ? out.k1 : out.k0, s, 2, kind == IDISP_CONV_S2 ? 32 : 0))) return rc;
    }
  }
  if (cout == 1 && split_launches(kind, cin) == 1) {
    // 1-channel head: w_lo rides in the (otherwise zero) output column 1, so x_hi feeds both weight words in ONE MMA
    if ((rc = tc_weights_prepare(w_tap, kind, cin, cout, 1, out.both, s, 1, 0, 1))) return rc;
  } else if (split_launches(kind, cin) == 1) {
    // two-word packing; the stride-2 32->64 conv keeps 32-wide blocks here (2 x 110 KB of 64-wide weights would not fit)
    const int nt = (kind == IDISP_CONV_S2 && cout == 64) ? 32 : (kind == IDISP_DECONV_S2 ? 16 : 0);
    if ((rc = tc_weights_prepare(w_tap, kind, cin, cout, 1, out.both, s, 2, nt))) return rc;
  }
  return IDISP_OK;
}

int tc_conv3d_split(const TcSplitWeights &w, const __nv_bfloat16 *x, int B, int Cin, int D, int H, int W, int Cout, int kind,
                    const float *bias, const __nv_bfloat16 *residual, int relu, __nv_bfloat16 *y, const float *res1, float *y1,
                    void *scratch, int flags, __nv_bfloat16 *y_split, const TcCostVolume *cv, float *part, cudaStream_t s, int *launches,
                    int *range_flag, __nv_bfloat16 *x_copy_split)
{
  const int per_view = cv ? Cin / 16 : Cin / 8;  // channel blocks of one precision word (per view for the fused cost volume)
  const int nl = split_launches(kind, Cin);
  if (launches) *launches = nl;
  TcOpts o;
  o.in_blk_stride = 2 * per_view;
  o.range_flag = range_flag;
  int rc;
  if (nl == 1) {
    o.xp = w.both.ncat ? 2 : 1;
    if (y1) return tc_conv3d(w.both, x, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, res1, y1, scratch, flags & 1, nullptr, cv, s, &o);
    o.x2 = 1;
    o.x_copy_split = x_copy_split;
    return tc_conv3d(w.both, x, B, Cin, D, H, W, Cout, kind, bias, residual, relu, y, nullptr, nullptr, scratch, flags, y_split, cv, s, &o);
  }
  if (x_copy_split) { set_error("tc_conv3d_split: the input side copy needs a single-launch layer"); return IDISP_ERR_INVALID; }
  if (y1) {  // 1-channel head: the passes accumulate straight into the f32 logits
    if (nl == 2) o.xp = 2;
    if ((rc = tc_conv3d(w.hi, x, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, res1, y1, scratch, flags & 1, nullptr, cv, s, &o))) return rc;
    if (nl == 3) {
      o.in_blk_off = per_view;
      if ((rc = tc_conv3d(w.hi, x, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, y1, y1, scratch, flags & 1, nullptr, cv, s, &o))) return rc;
    }
    o.xp = 0; o.in_blk_off = 0;
    return tc_conv3d(w.lo, x, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, y1, y1, scratch, flags & 1, nullptr, cv, s, &o);
  }
  if (!part) { set_error("tc_conv3d_split: this layer needs the fp32 partial buffer"); return IDISP_ERR_INVALID; }
  if (cv && w.k0.dev && w.k1.dev && ksplit_layer(kind, Cin, Cout)) {
    // fused cost volume: the two channel halves ARE the two views (left features masked, right features shifted)
    if (launches) *launches = 2;
    TcCostVolume half = *cv;
    o.xp = 1; o.part_out = part;
    half.view = 0;
    if ((rc = tc_conv3d(w.k0, x, B, 32, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, nullptr, nullptr, scratch, flags & 1, nullptr, &half, s, &o))) return rc;
    half.view = 1;
    o.part_in = part; o.part_out = nullptr; o.x2 = 1;
    return tc_conv3d(w.k1, x, B, 32, D, H, W, Cout, kind, bias, residual, relu, y, nullptr, nullptr, scratch, flags, y_split, &half, s, &o);
  }
  if (!cv && w.k0.dev && w.k1.dev && ksplit_layer(kind, Cin, Cout)) {
    if (launches) *launches = 2;
    o.xp = 1; o.in_lo_off = per_view; o.part_out = part;
    if ((rc = tc_conv3d(w.k0, x, B, 32, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, nullptr, nullptr, scratch, flags & 1, nullptr, nullptr, s, &o))) return rc;
    o.in_blk_off = 4; o.part_in = part; o.part_out = nullptr; o.x2 = 1;
    return tc_conv3d(w.k1, x, B, 32, D, H, W, Cout, kind, bias, residual, relu, y, nullptr, nullptr, scratch, flags, y_split, nullptr, s, &o);
  }
  static const int heavy_first = tc::env_flag("IDISP_X2_HEAVY_FIRST");  // A/B switch: the order before this scheduling
  if (nl == 2 && !heavy_first) {
    // light launch first: x_hi*w_lo (one term, its epilogue only stores the fp32 partial); then the two-term launch, whose
    // twice-as-long MMA phase hides the real epilogue (partial + residual reads, hi|lo stores).  The other order left the
    // epilogue-heavy work to the MMA-light launch (dres0.0: 2.96 + 1.74 ms).
    o.part_out = part;
    if ((rc = tc_conv3d(w.lo, x, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, nullptr, nullptr, scratch, flags & 1, nullptr, cv, s, &o))) return rc;
    o.xp = 2; o.part_in = part; o.part_out = nullptr; o.x2 = 1;
    return tc_conv3d(w.hi, x, B, Cin, D, H, W, Cout, kind, bias, residual, relu, y, nullptr, nullptr, scratch, flags, y_split, cv, s, &o);
  }
  o.part_out = part;
  if (nl == 2) o.xp = 2;
  if ((rc = tc_conv3d(w.hi, x, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, nullptr, nullptr, scratch, flags & 1, nullptr, cv, s, &o))) return rc;
  if (nl == 3) {
    o.part_in = part; o.in_blk_off = per_view;
    if ((rc = tc_conv3d(w.hi, x, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, nullptr, nullptr, scratch, flags & 1, nullptr, cv, s, &o))) return rc;
  }
  o.xp = 0; o.part_in = part; o.part_out = nullptr; o.in_blk_off = 0; o.x2 = 1;
  return tc_conv3d(w.lo, x, B, Cin, D, H, W, Cout, kind, bias, residual, relu, y, nullptr, nullptr, scratch, flags, y_split, cv, s, &o);
}

}  // namespace idisp
