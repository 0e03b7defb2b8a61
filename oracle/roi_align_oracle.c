/* CPU oracle for ROIAlign forward -- TEST INFRASTRUCTURE ONLY (see idispnet_oracle.py).
 *
 * Plain-C restatement of the reference's CPU kernel
 *   disprcnn/csrc/cpu/ROIAlign_cpu.cpp:18-111  (pre_calc_for_bilinear_interpolate)
 *   disprcnn/csrc/cpu/ROIAlign_cpu.cpp:114-219 (ROIAlignForward_cpu_kernel, T = float)
 * written as one scalar loop nest (no pre-calc table); the float operation
 * order per sample and the (iy, ix) accumulation order are those of the
 * reference, so results are bit-identical.  Built by oracle/Makefile with
 * -ffp-contract=off so gcc cannot fuse the multiply-adds.
 */
#include <math.h>
#include <stdint.h>

int oracle_roi_align_forward(const float *input, int N, int C, int H, int W,
                             const float *rois, int R, float spatial_scale,
                             int pooled_h, int pooled_w, int sampling_ratio, float *out)
{
    (void)N;
    for (int n = 0; n < R; ++n) {
        const float *roi = rois + (int64_t)n * 5;
        int b = (int)roi[0];
        float rsw = roi[1] * spatial_scale;
        float rsh = roi[2] * spatial_scale;
        float rew = roi[3] * spatial_scale;
        float reh = roi[4] * spatial_scale;
        float roi_w = fmaxf(rew - rsw, 1.0f);
        float roi_h = fmaxf(reh - rsh, 1.0f);
        float bin_h = roi_h / (float)pooled_h;
        float bin_w = roi_w / (float)pooled_w;
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / pooled_h);
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / pooled_w);
        const float count = (float)(gh * gw);
        for (int c = 0; c < C; ++c) {
            const float *img = input + ((int64_t)b * C + c) * H * W;
            float *o = out + ((int64_t)n * C + c) * pooled_h * pooled_w;
            for (int ph = 0; ph < pooled_h; ++ph)
                for (int pw = 0; pw < pooled_w; ++pw) {
                    float acc = 0.f;
                    for (int iy = 0; iy < gh; ++iy) {
                        float y = rsh + ph * bin_h + (float)(iy + .5f) * bin_h / (float)gh;
                        for (int ix = 0; ix < gw; ++ix) {
                            float x = rsw + pw * bin_w + (float)(ix + .5f) * bin_w / (float)gw;
                            float yy = y;
                            if (yy < -1.0f || yy > H || x < -1.0f || x > W) continue;
                            if (yy <= 0) yy = 0;
                            if (x <= 0) x = 0;
                            int y_low = (int)yy, x_low = (int)x, y_high, x_high;
                            if (y_low >= H - 1) { y_high = y_low = H - 1; yy = (float)y_low; }
                            else y_high = y_low + 1;
                            if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; }
                            else x_high = x_low + 1;
                            float ly = yy - y_low, lx = x - x_low;
                            float hy = 1.f - ly, hx = 1.f - lx;
                            float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                            acc += w1 * img[y_low * W + x_low] + w2 * img[y_low * W + x_high] +
                                   w3 * img[y_high * W + x_low] + w4 * img[y_high * W + x_high];
                        }
                    }
                    o[ph * pooled_w + pw] = acc / count;
                }
        }
    }
    return 0;
}
