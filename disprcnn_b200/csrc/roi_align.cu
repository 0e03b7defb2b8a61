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
    int pw, ph, cg, n;
    if (total < (1ll << 31)) {  // (32-bit index arithmetic: the 64-bit divisions were a third of this kernel's instructions)
      const uint32_t i = (uint32_t)idx, a = i / (uint32_t)pw_n, b2 = a / (uint32_t)ph_n;
      pw = (int)(i - a * (uint32_t)pw_n); ph = (int)(a - b2 * (uint32_t)ph_n);
      n = (int)(b2 / (uint32_t)cgroups); cg = (int)(b2 - (uint32_t)n * (uint32_t)cgroups);
    } else {
      pw = (int)(idx % pw_n);
      ph = (int)((idx / pw_n) % ph_n);
      cg = (int)((idx / ((int64_t)pw_n * ph_n)) % cgroups);
      n = (int)(idx / ((int64_t)pw_n * ph_n * cgroups));
    }
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
      // (x / 1.0f == x exactly in IEEE arithmetic: the one-sample-per-bin case -- ROIs up to the crop size -- skips those divisions)
      float y = __fadd_rn(ybase, gh == 1 ? __fmul_rn((float)iy + .5f, bin_h) : __fdiv_rn(__fmul_rn((float)iy + .5f, bin_h), (float)gh));
      const bool y_oob = (y < -1.0f) || (y > (float)H);
      if (y <= 0) y = 0;
      int y_low = (int)y, y_high;
      if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
      const float ly = __fsub_rn(y, (float)y_low), hy = __fsub_rn(1.f, ly);
      for (int ix = 0; ix < gw; ++ix) {
        float x = __fadd_rn(xbase, gw == 1 ? __fmul_rn((float)ix + .5f, bin_w) : __fdiv_rn(__fmul_rn((float)ix + .5f, bin_w), (float)gw));
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
        float v = count == 1.f ? acc[c] : __fdiv_rn(acc[c], count);
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

// ---------------------------------------------------------------------------------------
// Stereo ROI preparation on the device (disprcnn3d.py:126-146 + utils/stereo_utils.py:219-229): the reference walks the
// boxes in Python (.tolist() = one device->host sync per image) to build the aligned left/right crop rectangles.  Same
// integer arithmetic here, one thread per ROI, no host round trip: floor/ceil to integers, clamp to the image, common width
// max_width = min(max(x2-x1, x2p-x1p), W-x1, W-x1p); left ROI (i, x1, y1, x1+mw, y2), right ROI (i, x1p, y1, x1p+mw, y2).
// ---------------------------------------------------------------------------------------
namespace idisp {
__global__ void stereo_rois_kernel(const float *__restrict__ lb, const float *__restrict__ rb, const int *__restrict__ img, int R,
                                   int width, int height, const int *__restrict__ image_wh, int n_images,
                                   float *__restrict__ rois_l, float *__restrict__ rois_r, long long *__restrict__ xs)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (image_wh) {  // per-image (unpadded BoxList) size: left_result[i].width / .height, disprcnn3d.py:136-141
    const int i = min(max(img[r], 0), n_images - 1);
    width = image_wh[2 * i];
    height = image_wh[2 * i + 1];
  }
  int x1 = (int)floorf(lb[r * 4 + 0]), y1 = (int)floorf(lb[r * 4 + 1]), x2 = (int)ceilf(lb[r * 4 + 2]), y2 = (int)ceilf(lb[r * 4 + 3]);
  int x1p = (int)floorf(rb[r * 4 + 0]), x2p = (int)ceilf(rb[r * 4 + 2]);
  x1 = max(0, x1); x1p = max(0, x1p); y1 = max(0, y1);
  y2 = min(y2, height - 1); x2 = min(x2, width - 1); x2p = min(x2p, width - 1);
  int mw = max(x2 - x1, x2p - x1p);
  mw = min(mw, min(width - x1, width - x1p));
  const float fi = (float)img[r];
  rois_l[r * 5 + 0] = fi; rois_l[r * 5 + 1] = (float)x1;  rois_l[r * 5 + 2] = (float)y1; rois_l[r * 5 + 3] = (float)(x1 + mw);  rois_l[r * 5 + 4] = (float)y2;
  rois_r[r * 5 + 0] = fi; rois_r[r * 5 + 1] = (float)x1p; rois_r[r * 5 + 2] = (float)y1; rois_r[r * 5 + 3] = (float)(x1p + mw); rois_r[r * 5 + 4] = (float)y2;
  if (xs) { xs[r] = x1; xs[R + r] = x1p; xs[2 * R + r] = x1 + mw; xs[3 * R + r] = x1p + mw; }
}
}  // namespace idisp

extern "C" int idisp_stereo_rois(const float *left_boxes, const float *right_boxes, const int *image_index, int R, int width,
                                 int height, const int *image_wh, int n_images, float *rois_left, float *rois_right,
                                 long long *x1_x1p_x2_x2p, void *stream)
{
  using namespace idisp;
  IDISP_REQUIRE(R >= 0 && ((image_wh && n_images > 0) || (width > 0 && height > 0)),
                "stereo_rois: bad arguments R=%d width=%d height=%d n_images=%d", R, width, height, n_images);
  if (R == 0) return IDISP_OK;
  IDISP_REQUIRE(left_boxes && right_boxes && image_index && rois_left && rois_right, "stereo_rois: NULL pointer");
  stereo_rois_kernel<<<ceil_div(R, 128), 128, 0, (cudaStream_t)stream>>>(left_boxes, right_boxes, image_index, R, width, height, image_wh,
                                                                         n_images, rois_left, rois_right, x1_x1p_x2_x2p);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
