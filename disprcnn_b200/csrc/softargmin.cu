// softargmin.cu -- fused trilinear upsample (align_corners) + softmax over disparity + expectation.
//
// Replaces disprcnn/modeling/psmnet/stackhourglass.py:169-172 (F.interpolate to
// [maxdisp-mindisp, H, W], squeeze, F.softmax(dim=1)) and submodule.py:51-57
// (disparityregression).  The reference materialises the upsampled volume, the softmax and a
// repeated disparity ramp -- 3 x 154 MB per ROI at BASELINE config 2; here the logits
// [B,D,Hf,Wf] (2.4 MB/ROI) are read through L1/L2 and only the [B,H,W] map (0.8 MB/ROI) is
// written, so algorithmic bytes = 4*(V + H*W) per ROI and the work is 4D exp per pixel (SFU).
//
// Interpolation follows ATen's upsample_trilinear3d exactly: per axis
// scale=(in-1)/(out-1) in f32, src=scale*o, i0=(int)src, i1=min(i0+1,in-1), l1=src-i0, l0=1-l1,
// blend nested x -> y -> d.  One thread owns one output pixel: it bilinearly samples the D
// low-resolution planes at (y,x) (warp-coalesced: 4 neighbouring pixels share their corners),
// takes their maximum as the softmax stabiliser (an upper bound of every interpolated logit),
// then streams the 4D output disparities keeping two adjacent plane samples in registers.
#include "common.cuh"

namespace idisp {

__global__ void __launch_bounds__(256)
softargmin_kernel(const float *__restrict__ logits, int B, int D, int Hf, int Wf, int mindisp, int Dfull, int H,
                  int W, float *__restrict__ out)
{
  const int64_t total = (int64_t)B * H * W;
  const float sd = Dfull > 1 ? (float)(D - 1) / (float)(Dfull - 1) : 0.f;
  const float sh = H > 1 ? (float)(Hf - 1) / (float)(H - 1) : 0.f;
  const float sw = W > 1 ? (float)(Wf - 1) / (float)(W - 1) : 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int b = (int)(idx / ((int64_t)W * H));
    const float fy = sh * (float)y, fx = sw * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hf - 1 ? 1 : 0), x1 = x0 + (x0 < Wf - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
    const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
    const float *base = logits + (int64_t)b * D * Hf * Wf;
    const int o00 = y0 * Wf + x0, o01 = y0 * Wf + x1, o10 = y1 * Wf + x0, o11 = y1 * Wf + x1;
    const int plane = Hf * Wf;
    auto sample = [&](int k) {
      const float *p = base + (int64_t)k * plane;
      return ly0 * (lx0 * __ldg(p + o00) + lx1 * __ldg(p + o01)) + ly1 * (lx0 * __ldg(p + o10) + lx1 * __ldg(p + o11));
    };
    // softmax stabiliser = the maximum of the INTERPOLATED logits themselves (same expression as below): the maximum of the
    // plane samples is only an upper bound, and with sharply peaked logits every exp(v - bound) could underflow -> 0/0
    float m = -INFINITY;
    {
      int k0 = 0;
      float p0 = sample(0), p1 = sample(D > 1 ? 1 : 0);
      for (int d = 0; d < Dfull; ++d) {
        const float fd = sd * (float)d;
        const int d0 = (int)fd;
        if (d0 != k0) { k0 = d0; p0 = p1; p1 = sample(d0 + (d0 < D - 1 ? 1 : 0)); }
        const float l1 = fd - (float)d0, l0 = 1.f - l1;
        m = fmaxf(m, l0 * p0 + l1 * p1);
      }
    }
    float s = 0.f, t = 0.f;
    int k0 = 0;
    float p0 = sample(0), p1 = sample(D > 1 ? 1 : 0);
    for (int d = 0; d < Dfull; ++d) {
      const float fd = sd * (float)d;
      const int d0 = (int)fd;
      if (d0 != k0) {  // advances by at most one plane per step (scale < 1)
        k0 = d0;
        p0 = p1;
        p1 = sample(d0 + (d0 < D - 1 ? 1 : 0));
      }
      const float l1 = fd - (float)d0, l0 = 1.f - l1;
      const float v = l0 * p0 + l1 * p1;
      const float e = __expf(v - m);
      s += e;
      t = fmaf(e, (float)(mindisp + d), t);
    }
    out[idx] = t / s;
  }
}

// ---- pieces of the Dfull = 4*D specialisation (below) ----
struct SampleCtx {  // bilinear sampling of one low-resolution plane at this thread's output pixel
  const float *p00;
  int plane, d01, d10;
  float lx0, lx1, ly0, ly1;
  __device__ __forceinline__ float operator()(int k) const
  {
    const float *q = p00 + k * plane;
    return ly0 * (lx0 * __ldg(q) + lx1 * __ldg(q + d01)) + ly1 * (lx0 * __ldg(q + d10) + lx1 * __ldg(q + d10 + d01));
  }
};

// sum and disparity-weighted sum of exp(v(d) - m) over the 4D interpolated logits; rolling window of two plane samples.
// RECUR: between two planes the interpolated logit is linear in d, so its exponentials form a geometric progression -- the first one
// of a segment and the ratio g = 2^(sd * (P[k+1] - P[k]) * log2 e) cost two MUFU.EX2, the other three or four are one multiply each
// (96 instead of 192 exponentials and 3 instead of 6 instructions per output disparity).  The progression starts from the segment's
// first term; if that one underflows while a later one would not (logits more than ~44 apart between adjacent planes) the terms in
// between are lost, so such a pixel reports `steep` and the caller redoes it with one exponential per disparity (RECUR = false).
template <int D, bool RECUR>
__device__ __forceinline__ bool softargmin_accumulate(const SampleCtx &sample, float mneg, int mindisp, float &s_out, float &t_out)
{
  constexpr int Dfull = 4 * D;
  constexpr float sd = (float)(D - 1) / (float)(Dfull - 1);
  constexpr float LOG2E = 1.4426950408889634f;
  // four independent (sum, weighted-sum) chains: with one chain the add latency of 2 x 192 dependent accumulations sets the pace
  float s4[4] = {0.f, 0.f, 0.f, 0.f}, t4[4] = {0.f, 0.f, 0.f, 0.f};
  float Pk = sample(0), Pk1 = sample(D > 1 ? 1 : 0);
  float e = 0.f, g = 1.f;
  bool steep = false;
#pragma unroll
  for (int d = 0; d < Dfull; ++d) {
    const float fd = sd * (float)d;      // compile-time per unrolled iteration
    const int d0 = (int)fd;
    const bool first = d == 0 || d0 != (int)(sd * (float)(d - 1));   // (compile-time) first output disparity of a segment
    if (d > 0 && first) {  // the window advances by one plane
      Pk = Pk1;
      Pk1 = sample(d0 + (d0 < D - 1 ? 1 : 0));
    }
    const float l1 = fd - (float)d0, l0 = 1.f - l1;
    if (RECUR) {
      if (first) {
        const float delta2 = (Pk1 - Pk) * LOG2E;
        steep |= fabsf(delta2) > 64.f;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(sd * delta2));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(l1, delta2, fmaf(Pk, LOG2E, mneg))));
      } else {
        e *= g;
      }
    } else {
      const float v = l0 * Pk + l1 * Pk1;
      // one MUFU.EX2 (arguments are <= 0; flushing the far tail to zero is what softmax does to it anyway)
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v, LOG2E, mneg)));
    }
    s4[d & 3] += e;
    t4[d & 3] = fmaf(e, (float)(mindisp + d), t4[d & 3]);
  }
  s_out = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  t_out = (t4[0] + t4[1]) + (t4[2] + t4[3]);
  return steep;
}

// rare path (kept out of line so that it does not weigh on the fast path's registers): exact maximum of the interpolated
// logits -- the SAME expression the accumulation evaluates, so the largest term is exp(0) -- then the accumulation again
template <int D>
__device__ __noinline__ float softargmin_exact(SampleCtx sample, int mindisp)
{
  constexpr int Dfull = 4 * D;
  constexpr float sd = (float)(D - 1) / (float)(Dfull - 1);
  float m = -INFINITY;
  float Pk = sample(0), Pk1 = sample(D > 1 ? 1 : 0);
#pragma unroll
  for (int d = 0; d < Dfull; ++d) {
    const float fd = sd * (float)d;
    const int d0 = (int)fd;
    if (d > 0 && d0 != (int)(sd * (float)(d - 1))) {
      Pk = Pk1;
      Pk1 = sample(d0 + (d0 < D - 1 ? 1 : 0));
    }
    // between two planes the interpolant is monotone in d: only the first and the last d of a segment can hold the maximum
    const bool first = d == 0 || d0 != (int)(sd * (float)(d - 1)), last = d == Dfull - 1 || d0 != (int)(sd * (float)(d + 1));
    if (first || last) {
      const float l1 = fd - (float)d0, l0 = 1.f - l1;
      m = fmaxf(m, l0 * Pk + l1 * Pk1);
    }
  }
  asm volatile("" : "+l"(sample.p00));
  float s, t;
  softargmin_accumulate<D, false>(sample, -m * 1.4426950408889634f, mindisp, s, t);
  return t / s;
}

// Specialisation for the reference's geometry Dfull = 4*D (D = 24 or 48).  The depth interpolation indices/weights d0(d), l1(d) are
// compile-time constants of the fully unrolled 4D loop.  Two passes over the D plane samples (each a bilinear blend of 4 L1/L2-
// resident logits): pass 1 takes their maximum (upper bound of the interpolated logits), pass 2 streams the 4D output disparities
// through a rolling window of TWO samples and accumulates the exponentials (with an exact-maximum fallback, see below).  Keeping all D samples in registers instead (the
// first version) cost 168 registers = 3 warps per scheduler, and the kernel ran at a third of its issue rate; re-sampling is
// 48 x 11 instructions per pixel against 192 x 6 for the exponentials.
// per low-resolution cell: maximum of its D logits (one thread per cell, coalesced over x)
__global__ void __launch_bounds__(256) cell_max_kernel(const float *__restrict__ logits, int64_t cells, int D, int plane, float *__restrict__ cm)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
    const float *p = logits + (i / plane) * (int64_t)D * plane + i % plane;
    float m = __ldg(p);
    for (int k = 1; k < D; ++k) m = fmaxf(m, __ldg(p + (int64_t)k * plane));
    cm[i] = m;
  }
}

// PRE: the softmax stabiliser comes from `cellmax` (maximum over depth of each low-resolution cell, computed once by cell_max_kernel):
// the maximum over a pixel's four corner cells bounds every interpolated logit from above, like the maximum of the D blended plane
// samples did, but costs 4 loads instead of a first sampling pass (48 x 11 of the ~2 300 instructions per pixel).
template <int D, bool PRE>
__global__ void __launch_bounds__(256, 3)  // <= 85 registers: the unrolled loops must not hoist all 4*D*... loads at once
softargmin_x4_kernel(const float *__restrict__ logits, int B, int Hf, int Wf, int mindisp, int H, int W, float *__restrict__ out,
                     const float *__restrict__ cellmax)
{
  constexpr int Dfull = 4 * D;
  constexpr float sd = (float)(D - 1) / (float)(Dfull - 1);
  const int64_t total = (int64_t)B * H * W;
  const float sh = H > 1 ? (float)(Hf - 1) / (float)(H - 1) : 0.f;
  const float sw = W > 1 ? (float)(Wf - 1) / (float)(W - 1) : 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int b = (int)(idx / ((int64_t)W * H));
    const float fy = sh * (float)y, fx = sw * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hf - 1 ? 1 : 0), x1 = x0 + (x0 < Wf - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
    const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
    const int plane = Hf * Wf;
    const float *p00 = logits + (int64_t)b * D * plane + y0 * Wf + x0;
    const int d01 = x1 - x0, d10 = (y1 - y0) * Wf;
    SampleCtx sample{p00, plane, d01, d10, lx0, lx1, ly0, ly1};
    // softmax stabiliser.  Fast path: the maximum M of the D plane samples -- an upper bound of every interpolated logit, one
    // independent sample per plane.  It is NOT safe on its own: with sharply peaked logits every exp(v - M) can underflow
    // (sum = 0 -> 0/0), so a vanishing sum falls back to softargmin_exact, as F.softmax on the upsampled volume
    // (stackhourglass.py:169-172) would behave.
    float M;
    if constexpr (PRE) {
      const float *c = cellmax + (int64_t)b * plane + y0 * Wf + x0;
      M = fmaxf(fmaxf(__ldg(c), __ldg(c + d01)), fmaxf(__ldg(c + d10), __ldg(c + d10 + d01)));
    } else {
      M = sample(0);
#pragma unroll
      for (int k = 1; k < D; ++k) M = fmaxf(M, sample(k));
      asm volatile("" : "+l"(sample.p00));  // the accumulation RE-samples (L1 hits): keeping the D samples alive would cost the occupancy
    }
    float s, t;
    const bool steep = softargmin_accumulate<D, true>(sample, -M * 1.4426950408889634f, mindisp, s, t);
    out[idx] = (s >= 1e-30f && !steep) ? t / s : softargmin_exact<D>(sample, mindisp);
  }
}

int launch_softargmin(const float *logits, int B, int D, int Hf, int Wf, int mindisp, int maxdisp, int H, int W,
                      float *out, cudaStream_t s, float *cellmax_scratch)
{
  const int64_t total = (int64_t)B * H * W;
  if (total == 0) return IDISP_OK;
  static const bool generic = getenv("IDISP_SOFTARGMIN_GENERIC") != nullptr, no_pre = getenv("IDISP_SOFTARGMIN_NO_CELLMAX") != nullptr;
  if (maxdisp - mindisp == 4 * D && (D == 24 || D == 48) && !generic) {
    const int64_t want256 = ceil_div64(total, 256);
    const int grid = (int)(want256 < 148 * 64 ? want256 : 148 * 64);
    if (cellmax_scratch && !no_pre) {   // (B * Hf * Wf floats of caller scratch)
      const int64_t cells = (int64_t)B * Hf * Wf;
      cell_max_kernel<<<(int)ceil_div64(cells, 256), 256, 0, s>>>(logits, cells, D, Hf * Wf, cellmax_scratch);
      if (D == 24) softargmin_x4_kernel<24, true><<<grid, 256, 0, s>>>(logits, B, Hf, Wf, mindisp, H, W, out, cellmax_scratch);
      else softargmin_x4_kernel<48, true><<<grid, 256, 0, s>>>(logits, B, Hf, Wf, mindisp, H, W, out, cellmax_scratch);
    } else if (D == 24) softargmin_x4_kernel<24, false><<<grid, 256, 0, s>>>(logits, B, Hf, Wf, mindisp, H, W, out, nullptr);
    else softargmin_x4_kernel<48, false><<<grid, 256, 0, s>>>(logits, B, Hf, Wf, mindisp, H, W, out, nullptr);
    IDISP_LAUNCH_CHECK();
    return IDISP_OK;
  }
  const int64_t want = ceil_div64(total, 256);
  softargmin_kernel<<<(int)(want < 148 * 64 ? want : 148 * 64), 256, 0, s>>>(logits, B, D, Hf, Wf, mindisp,
                                                                              maxdisp - mindisp, H, W, out);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

}  // namespace idisp

extern "C" int idisp_softargmin(const float *logits, int B, int D, int Hf, int Wf, int mindisp, int maxdisp, int H,
                                int W, float *out, void *stream)
{
  using namespace idisp;
  IDISP_REQUIRE(B >= 0 && D > 0 && Hf > 0 && Wf > 0 && H > 0 && W > 0 && maxdisp > mindisp,
                "softargmin: bad shape B=%d D=%d Hf=%d Wf=%d H=%d W=%d disp=[%d,%d)", B, D, Hf, Wf, H, W, mindisp, maxdisp);
  // trilinear UP-sampling only (scale <= 1 per axis), which is all the reference does (:169)
  IDISP_REQUIRE(maxdisp - mindisp >= D && H >= Hf && W >= Wf, "softargmin: output must not be smaller than the logits");
  if (B == 0) return IDISP_OK;
  IDISP_REQUIRE(logits && out, "softargmin: NULL pointer");
  return launch_softargmin(logits, B, D, Hf, Wf, mindisp, maxdisp, H, W, out, (cudaStream_t)stream);
}
