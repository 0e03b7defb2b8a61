// roi_align.cu -- ROIAlign forward for sm_100a.
//
// Replaces disprcnn/csrc/cuda/ROIAlign_cuda.cu:65-122 (RoIAlignForward) + :16-62
// (bilinear_interpolate).  HBM/L2-bound gather: one thread owns one output position
// (n, ph, pw) for a group of CG channels, so the sample coordinates, the four corner
// indices and the four weights are computed once and reused across the channel group
// (the reference recomputes them per channel).  pw is the fastest thread index, so the
// stores are fully coalesced and the four corner gathers of neighbouring lanes fall in the
// same or adjacent 32-byte sectors.
//
// Bit-exactness: all coordinate/weight arithmetic uses the round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn), which nvcc never contracts into FMAs, in the
// operation order of the reference's float instantiation
// (disprcnn/csrc/cpu/ROIAlign_cpu.cpp:36-43,64-93,196-206).  The result is therefore
// bit-identical to the reference's CPU kernel, indices and values.
#include "common.cuh"

namespace idisp {

constexpr int ROI_CG = 4;  // channels per thread

__global__ void __launch_bounds__(256)
roi_align_fwd_kernel(const float *__restrict__ in, const float *__restrict__ rois, int C, int H, int W,
                     int R, float scale, int ph_n, int pw_n, int sr, const float *__restrict__ mean,
                     const float *__restrict__ stdv, float *__restrict__ out, int cgroups)
{
  const int64_t total = (int64_t)R * cgroups * ph_n * pw_n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % pw_n);
    const int ph = (int)((idx / pw_n) % ph_n);
    const int cg = (int)((idx / ((int64_t)pw_n * ph_n)) % cgroups);
    const int n = (int)(idx / ((int64_t)pw_n * ph_n * cgroups));
    const float *roi = rois + (int64_t)n * 5;
    const int b = (int)roi[0];
    const float rsw = __fmul_rn(roi[1], scale), rsh = __fmul_rn(roi[2], scale);
    const float rew = __fmul_rn(roi[3], scale), reh = __fmul_rn(roi[4], scale);
    const float roi_w = fmaxf(__fsub_rn(rew, rsw), 1.0f);
    const float roi_h = fmaxf(__fsub_rn(reh, rsh), 1.0f);
    const float bin_h = __fdiv_rn(roi_h, (float)ph_n);
    const float bin_w = __fdiv_rn(roi_w, (float)pw_n);
    const int gh = sr > 0 ? sr : (int)ceilf(__fdiv_rn(roi_h, (float)ph_n));
    const int gw = sr > 0 ? sr : (int)ceilf(__fdiv_rn(roi_w, (float)pw_n));
    const float count = (float)(gh * gw);
    const int c0 = cg * ROI_CG;
    const int nc = min(ROI_CG, C - c0);
    const float *img = in + ((int64_t)b * C + c0) * H * W;
    const int64_t plane = (int64_t)H * W;
    float acc[ROI_CG];
#pragma unroll
    for (int c = 0; c < ROI_CG; ++c) acc[c] = 0.f;
    const float ybase = __fadd_rn(rsh, __fmul_rn((float)ph, bin_h));
    const float xbase = __fadd_rn(rsw, __fmul_rn((float)pw, bin_w));
    for (int iy = 0; iy < gh; ++iy) {
      float y = __fadd_rn(ybase, __fdiv_rn(__fmul_rn((float)iy + .5f, bin_h), (float)gh));
      const bool y_oob = (y < -1.0f) || (y > (float)H);
      if (y <= 0) y = 0;
      int y_low = (int)y, y_high;
      if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
      const float ly = __fsub_rn(y, (float)y_low), hy = __fsub_rn(1.f, ly);
      for (int ix = 0; ix < gw; ++ix) {
        float x = __fadd_rn(xbase, __fdiv_rn(__fmul_rn((float)ix + .5f, bin_w), (float)gw));
        if (y_oob || x < -1.0f || x > (float)W) continue;  // contributes exactly 0
        if (x <= 0) x = 0;
        int x_low = (int)x, x_high;
        if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
        const float lx = __fsub_rn(x, (float)x_low), hx = __fsub_rn(1.f, lx);
        const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
        const int p1 = y_low * W + x_low, p2 = y_low * W + x_high, p3 = y_high * W + x_low, p4 = y_high * W + x_high;
#pragma unroll
        for (int c = 0; c < ROI_CG; ++c) {
          if (c < nc) {
            const float *pl = img + c * plane;
            float v = __fadd_rn(__fmul_rn(w1, __ldg(pl + p1)), __fmul_rn(w2, __ldg(pl + p2)));
            v = __fadd_rn(v, __fmul_rn(w3, __ldg(pl + p3)));
            v = __fadd_rn(v, __fmul_rn(w4, __ldg(pl + p4)));
            acc[c] = __fadd_rn(acc[c], v);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < ROI_CG; ++c) {
      if (c < nc) {
        float v = __fdiv_rn(acc[c], count);
        if (mean != nullptr) v = __fdiv_rn(__fsub_rn(v, mean[c0 + c]), stdv[c0 + c]);
        out[(((int64_t)n * C + c0 + c) * ph_n + ph) * pw_n + pw] = v;
      }
    }
  }
}

}  // namespace idisp

extern "C" int idisp_roi_align_forward(const float *input, int N, int C, int H, int W, const float *rois,
                                       int R, float spatial_scale, int pooled_h, int pooled_w,
                                       int sampling_ratio, const float *mean, const float *stdv, float *out,
                                       void *stream)
{
  using namespace idisp;
  IDISP_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && pooled_h > 0 && pooled_w > 0,
                "roi_align: bad shape N=%d C=%d H=%d W=%d R=%d pooled=%dx%d", N, C, H, W, R, pooled_h, pooled_w);
  IDISP_REQUIRE((mean == nullptr) == (stdv == nullptr), "roi_align: mean and std must both be given or both NULL");
  if (R == 0) return IDISP_OK;  // reference early return, ROIAlign_cuda.cu:278-281
  IDISP_REQUIRE(input && rois && out, "roi_align: NULL pointer");
  const int cgroups = ceil_div(C, ROI_CG);
  const int64_t total = (int64_t)R * cgroups * pooled_h * pooled_w;
  const int threads = 256;
  const int64_t want = ceil_div64(total, threads);
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);  // persistent-ish: 16 CTAs/SM, grid-stride
  roi_align_fwd_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(input, rois, C, H, W, R, spatial_scale,
                                                                   pooled_h, pooled_w, sampling_ratio, mean,
                                                                   stdv, out, cgroups);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

extern "C" int idisp_roi_align_backward(const float *, const float *, int, float, int, int, int, int, int, int,
                                        int, float *, void *)
{
  // Training-only (disprcnn/layers/roi_align.py:28-46).  The reference itself raises for the
  // non-CUDA build (csrc/ROIAlign.h:44); this inference path mirrors that.
  idisp::set_error("roi_align_backward: not implemented on the B200 inference path");
  return IDISP_ERR_UNSUPPORTED;
}
