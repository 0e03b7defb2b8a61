// conv3d_tc.cuh -- interface of the tcgen05 (5th-gen tensor core) 3x3x3 convolution path.
#pragma once
#include "common.cuh"

namespace idisp {

// bf16 weights re-laid for the UMMA B operand (see conv3d_tc.cu), device memory
struct TcWeights {
  void *dev = nullptr;
  size_t bytes = 0;
  int kind = -1, cin = 0, cout = 0;
  int nt = 32;  // output channels per stacked block the packing was made for
  int f16 = 0;  // 16-bit storage format of weights AND activations: 0 = bfloat16, 1 = IEEE half
  int ncat = 0;   // 1-channel head, split precision: output column 1 holds half(w - half(w)) (see tc_split_weights_prepare)
  int words = 1;  // 2: split precision, the lo word's k-steps follow the hi word's inside every tap (TcOpts::xp == 1)
};

// w_tap: HOST pointer, [27][cin][cout] f32 with the BN scale already folded in
// words = 2 packs half(w) and half(w - half(w)); nt > 0 overrides the layer's default block width
int tc_weights_prepare(const float *w_tap, int kind, int cin, int cout, int f16, TcWeights &out, cudaStream_t s, int words = 1, int nt = 0,
                       int ncat = 0);
void tc_weights_free(TcWeights &w);
// Fused cost volume for the first layer (stackhourglass.py:115-128 folded into dres0.0's loader): the per-view features
// in blocked bf16 [B][C/8][Hf][Wf][8]; D <= 64 planes (the per-plane tensor maps travel as kernel parameters).
struct TcCostVolume {
  const __nv_bfloat16 *left = nullptr, *right = nullptr;
  int shift0 = 0;  // mindisp / 4
  int view = -1;   // -1: both views (the layer's Cin = 2C channels); 0 / 1: only the left / right half of the volume (Cin = C channels:
                   // one launch of the K-split form of the first layer, see tc_conv3d_split)
};

// Split-precision ("x2") pass options.  Activations are then stored as 2*C/8 channel blocks per sample (hi words, then lo
// words); a layer is three launches -- (x_hi, w_hi), (x_lo, w_hi), (x_hi, w_lo) -- chained through an fp32 partial buffer.
struct TcOpts {
  const float *part_in = nullptr;  // fp32 partial of the earlier pass(es), blocked [B][Cout/8][V][8]
  float *part_out = nullptr;       // non-null: this launch only stores its accumulator (+ part_in) there
  int x2 = 0;                      // final pass: y / residual / y_split hold hi|lo block groups
  int in_blk_stride = 0;           // channel blocks per input sample in memory (0: Cin/8)
  int in_blk_off = 0;              // first channel block this launch reads
  int in_lo_off = 0;               // xp launches over a channel SUBSET: block distance from the launch's hi group to its lo group
                                   // (0: adjacent, i.e. Cin/8) -- the stage is then filled by two TMA boxes
  int *range_flag = nullptr;       // device int set to 1 if an activation leaves the IEEE-half range (fp16 modes)
  __nv_bfloat16 *x_copy_split = nullptr;  // stride-1 32 -> 32 split-precision launch only: ALSO write this launch's INPUT tensor, as it passes through
                                   // shared memory, in the 8-parity-sub-volume layout a stride-2 consumer reads (see Params::x_split)
  int xp = 0;                      // K concatenation inside the launch: 1 = (x_hi,w_hi)+(x_lo,w_hi)+(x_hi,w_lo) with two-word
                                   // weights (Cin 32), 2 = (x_hi,w_hi)+(x_lo,w_hi) with one-word weights (Cin 64)
};

// Split-precision layer = the fewest launches its shared-memory budget allows (see conv3d_tc.cu, Cfg::XP).
struct TcSplitWeights {
  TcWeights hi, lo;  // one-word packings of half(w) and half(w - half(w))
  TcWeights both;    // two-word packing (Cin = 32 layers), empty otherwise
  TcWeights k0, k1;  // stride-1 Cin = 64 layers, K split: two-word packings of input channels [0,32) and [32,64) (see tc_conv3d_split)
};
int tc_split_weights_prepare(const float *w_tap, int kind, int cin, int cout, TcSplitWeights &out, cudaStream_t s);
void tc_split_weights_free(TcSplitWeights &w);
// x / residual / y / y_split hold hi|lo block groups (2*C/8 blocks per sample); `part` = fp32 scratch [B][Cout][Vout]
// (used by the layers that need more than one launch).  Returns the number of kernel launches in *launches.
int tc_conv3d_split(const TcSplitWeights &w, const __nv_bfloat16 *x, int B, int Cin, int D, int H, int W, int Cout, int kind,
                    const float *bias, const __nv_bfloat16 *residual, int relu, __nv_bfloat16 *y, const float *res1, float *y1,
                    void *scratch, int flags, __nv_bfloat16 *y_split, const TcCostVolume *cv, float *part, cudaStream_t s,
                    int *launches = nullptr, int *range_flag = nullptr, __nv_bfloat16 *x_copy_split = nullptr);

// The 32 -> 1 classifier convolutions as "all 27 taps in N" GEMM + shifted sum (head_tc.cu); f16 / x2 as above.
struct TcHeadWeights {
  void *dev = nullptr;
  int f16 = 0, x2 = 0;
};
int tc_head_weights_prepare(const float *w_tap, int cin, int f16, int x2, TcHeadWeights &out, cudaStream_t s);
void tc_head_weights_free(TcHeadWeights &w);
bool tc_head_supported(const TcHeadWeights &w, int D, int H, int W);
// x: blocked 16-bit [B][(x2 ? 2 : 1) * 4][D][H][W][8]; y1 = conv(x) (+ res1), both [B][D][H][W] f32
int tc_head_conv(const TcHeadWeights &w, const __nv_bfloat16 *x, int B, int D, int H, int W, const float *res1, float *y1, cudaStream_t s);

// whether the tensor-core kernel covers this layer shape (otherwise the SIMT kernel runs it)
bool tc_supported(int kind, int cin, int cout, int D, int H, int W);
// device scratch the layer needs (stride-2 convs re-lay their input into parity sub-volumes)
size_t tc_scratch_bytes(int kind, int B, int cin, int D, int H, int W);
// Cout == 1 (the classifier head): y1/res1 are [B][D][H][W] f32 and y/residual/bias are unused.
int tc_conv3d(const TcWeights &w, const __nv_bfloat16 *x, int B, int Cin, int D, int H, int W, int Cout, int kind,
              const float *bias, const __nv_bfloat16 *residual, int relu, __nv_bfloat16 *y, const float *res1, float *y1,
              void *scratch, int x_is_split, __nv_bfloat16 *y_split, const TcCostVolume *cv, cudaStream_t s,
              const TcOpts *opts = nullptr);
// cv != nullptr: the layer's input IS the cost volume of (cv->left, cv->right); x is ignored and nothing is materialised.
// x_is_split is a bit set: 1 = a stride-2 layer's input pointer already holds the 8 parity sub-volumes (written by its producer's
// epilogue through y_split), so the space-to-depth pass is skipped; 2 = (transposed conv) `residual` is stored in the parity
// layout of the output grid; 4 = write only y_split, not y.

}  // namespace idisp
