// conv2d_tc.cuh -- interface of the tcgen05 2-D convolution path of the feature extractor (conv2d_tc.cu).
#pragma once
#include "common.cuh"

namespace idisp {

// packed B operand of one 3x3 conv layer (see c2d_weights_prepare)
struct C2dWeights {
  void *dev = nullptr;
  int cin = 0, cout = 0, taps = 9;
};

// a blocked split-precision activation tensor [N][blocks][H][W][8] (IEEE-half words), or a channel range inside one:
// the channels in question are blocks [blk0, blk0 + C/8) (hi words) and the same range `lo` blocks further on (lo words)
struct C2dTensor {
  __nv_bfloat16 *p = nullptr;
  int blocks = 0;   // channel blocks per sample in memory (= 2 * C_total / 8)
  int blk0 = 0;     // first block of the channel range
  int lo = 0;       // hi -> lo block distance (= C_total / 8)
};

int c2d_weights_prepare(const float *w /* HOST [Cin][taps][Cout] f32, taps = 9 (3x3) or 1 (1x1) */, int cin, int cout, int taps, C2dWeights &out,
                        cudaStream_t s);
void c2d_weights_free(C2dWeights &w);
int c2d_nchw_to_x2(const float *src, long long src_bs, __nv_bfloat16 *dst, int blocks, int blk0, int lo_off, int B, int C, long long HW, int *range_flag,
                   cudaStream_t s);
int c2d_x2_to_nchw(const __nv_bfloat16 *src, int blocks, int blk0, int lo_off, float *dst, long long dst_bs, int B, int C, long long HW, cudaStream_t s);
int c2d_conv(const C2dWeights &w, int dil, const C2dTensor &x, int chunk0, int nchunks, int B, int H, int W, const float *bias, const C2dTensor *res,
             int relu, const C2dTensor &y, const float *part_in, float *part_out, int *range_flag, cudaStream_t s);

}  // namespace idisp
