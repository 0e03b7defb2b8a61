// conv3d_simt.cu -- CUDA-core 3x3x3 convolution family (fp32 accumulate), the parity path.
//
// Replaces the cuDNN Conv3d / ConvTranspose3d + BatchNorm3d + ReLU + residual-add chains of
// disprcnn/modeling/psmnet/stackhourglass.py:11-30,63-88 (modules) as applied at :130-144.
// One kernel covers stride-1, stride-2 and the stride-2 transposed conv (gather form, by
// output-parity class) with the folded-BN bias, one optional
// residual tensor and an optional ReLU fused into the epilogue.
//
// Mapping.  Activations are channel-blocked-8 (common.cuh).  A CTA owns 128 consecutive
// output positions of one (n, d) plane (flattened h*W+w, so no tile overhang) and ALL Cout:
// warp q computes output channel block q for the 128 positions, lane l owns positions
// l, l+32, l+64, l+96 -> a 4x8 fp32 register tile per thread.  Per tap the [Cin][Cout] weight
// slice is staged in shared memory (cp.async, double buffered) and read as warp-wide
// broadcasts; activations are read straight from global memory as one 16/32-byte
// channel-block voxel per lane (coalesced across the warp; the 27-tap reuse is served by L1).
// Roofline: fp32 FFMA (148 SMs x 128 lanes x 2 flop x clock); this is the accuracy path, the
// throughput path is the tcgen05 kernel in conv3d_tc.cu.
#include "common.cuh"

namespace idisp {

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// KIND: IDISP_CONV_S1 / S2 / DECONV_S2.  For DECONV the "positions" are INPUT-resolution
// positions and blockIdx.z also enumerates the 8 output parity classes.
template <typename T, int CIN, int COUT, int KIND>
__global__ void __launch_bounds__(COUT / 8 * 32)
conv3d_simt_kernel(const T *__restrict__ x, const float *__restrict__ w_tap, const float *__restrict__ bias,
                   const T *__restrict__ residual, int relu, T *__restrict__ y, int Di, int Hi, int Wi, int Do,
                   int Ho, int Wo)
{
  constexpr int NW = COUT / 8;
  constexpr int NT = NW * 32;
  constexpr int WSLICE = CIN * COUT;  // floats per tap
  extern __shared__ __align__(16) float wsm[];  // [2][CIN][COUT]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int n = blockIdx.z, pd = 0, ph = 0, pw = 0;
  if (KIND == IDISP_DECONV_S2) { const int cls = n & 7; n >>= 3; pd = cls >> 2; ph = (cls >> 1) & 1; pw = cls & 1; }
  const int d = blockIdx.y;  // output plane (conv) / input-resolution plane (deconv)
  // position grid this CTA tiles: output grid for conv, input grid for deconv
  const int Hp = KIND == IDISP_DECONV_S2 ? Hi : Ho, Wp = KIND == IDISP_DECONV_S2 ? Wi : Wo;
  const int npos = Hp * Wp;
  const int p0 = blockIdx.x * 128 + lane;

  int hh[4], ww[4];
  bool pv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = p0 + 32 * j;
    pv[j] = p < npos;
    const int pc = pv[j] ? p : 0;
    hh[j] = pc / Wp;
    ww[j] = pc - hh[j] * Wp;
  }

  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

  const int ntd = KIND == IDISP_DECONV_S2 ? 1 + pd : 3;
  const int nth = KIND == IDISP_DECONV_S2 ? 1 + ph : 3;
  const int ntw = KIND == IDISP_DECONV_S2 ? 1 + pw : 3;
  const int ntaps = ntd * nth * ntw;

  // tap t -> kernel index (kd,kh,kw) and input offset (od,oh,ow) relative to the base position
  auto decode = [&](int t, int &kd, int &kh, int &kw, int &od, int &oh, int &ow) {
    const int td = t / (nth * ntw), th = (t / ntw) % nth, tw = t % ntw;
    if (KIND == IDISP_DECONV_S2) {
      // even output: k=1,in[n]; odd output: (k=2,in[n]), (k=0,in[n+1])
      kd = pd ? (td == 0 ? 2 : 0) : 1; od = pd ? td : 0;
      kh = ph ? (th == 0 ? 2 : 0) : 1; oh = ph ? th : 0;
      kw = pw ? (tw == 0 ? 2 : 0) : 1; ow = pw ? tw : 0;
    } else {
      kd = td; kh = th; kw = tw; od = td - 1; oh = th - 1; ow = tw - 1;
    }
  };
  auto stage = [&](int t, int buf) {
    int kd, kh, kw, od, oh, ow;
    decode(t, kd, kh, kw, od, oh, ow);
    const float *src = w_tap + (int64_t)((kd * 3 + kh) * 3 + kw) * WSLICE;
    float *dst = wsm + buf * WSLICE;
    for (int i = threadIdx.x; i < WSLICE / 4; i += NT) cp_async16(dst + 4 * i, src + 4 * i);
    cp_async_commit();
  };

  const int64_t Vi = (int64_t)Di * Hi * Wi;
  const T *xn = x + (int64_t)n * (CIN / 8) * Vi * 8;

  stage(0, 0);
  for (int t = 0; t < ntaps; ++t) {
    cp_async_wait<0>();
    __syncthreads();  // tap t landed for everyone; everyone finished reading tap t-1's buffer
    if (t + 1 < ntaps) stage(t + 1, (t + 1) & 1);
    int kd, kh, kw, od, oh, ow;
    decode(t, kd, kh, kw, od, oh, ow);
    const int di = KIND == IDISP_CONV_S2 ? 2 * d + od : d + od;
    if (di < 0 || di >= Di) continue;  // uniform across the CTA
    int64_t off[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int hi = KIND == IDISP_CONV_S2 ? 2 * hh[j] + oh : hh[j] + oh;
      const int wi = KIND == IDISP_CONV_S2 ? 2 * ww[j] + ow : ww[j] + ow;
      ok[j] = pv[j] && hi >= 0 && hi < Hi && wi >= 0 && wi < Wi;
      off[j] = (((int64_t)di * Hi + hi) * Wi + wi) * 8;
    }
    const float *wt = wsm + (t & 1) * WSLICE + warp * 8;
#pragma unroll 1
    for (int cb = 0; cb < CIN / 8; ++cb) {
      F8 a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ok[j]) a[j] = load8<T>(xn + (int64_t)cb * Vi * 8 + off[j]);
        else {
#pragma unroll
          for (int c = 0; c < 8; ++c) a[j].v[c] = 0.f;
        }
      }
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const float4 w0 = *reinterpret_cast<const float4 *>(wt + (cb * 8 + ci) * COUT);
        const float4 w1 = *reinterpret_cast<const float4 *>(wt + (cb * 8 + ci) * COUT + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int co = 0; co < 8; ++co) acc[j][co] = fmaf(a[j].v[ci], wv[co], acc[j][co]);
      }
    }
  }

  // epilogue: + bias (+ residual) (ReLU) -> blocked store of channel block `warp`
  float bv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) bv[c] = bias ? __ldg(bias + warp * 8 + c) : 0.f;
  const int64_t Vo = (int64_t)Do * Ho * Wo;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!pv[j]) continue;
    int dd = d, ho = hh[j], wo = ww[j];
    if (KIND == IDISP_DECONV_S2) { dd = 2 * d + pd; ho = 2 * hh[j] + ph; wo = 2 * ww[j] + pw; }
    const int64_t o = (((int64_t)n * NW + warp) * Vo + ((int64_t)dd * Ho + ho) * Wo + wo) * 8;
    F8 r;
#pragma unroll
    for (int c = 0; c < 8; ++c) r.v[c] = acc[j][c] + bv[c];
    if (residual) {
      const F8 q = load8<T>(residual + o);
#pragma unroll
      for (int c = 0; c < 8; ++c) r.v[c] += q.v[c];
    }
    if (relu) {
#pragma unroll
      for (int c = 0; c < 8; ++c) r.v[c] = fmaxf(r.v[c], 0.f);
    }
    store8<T>(y + o, r);
  }
}

template <typename T, int CIN, int COUT, int KIND>
static int launch_one(const T *x, int B, int D, int H, int W, const float *w_tap, const float *bias,
                      const T *residual, int relu, T *y, cudaStream_t s)
{
  int Do = D, Ho = H, Wo = W;
  if (KIND == IDISP_CONV_S2) { Do = (D + 1) / 2; Ho = (H + 1) / 2; Wo = (W + 1) / 2; }
  if (KIND == IDISP_DECONV_S2) { Do = 2 * D; Ho = 2 * H; Wo = 2 * W; }
  const int npos = KIND == IDISP_DECONV_S2 ? H * W : Ho * Wo;
  const int planes = KIND == IDISP_DECONV_S2 ? D : Do;
  dim3 grid(ceil_div(npos, 128), planes, KIND == IDISP_DECONV_S2 ? B * 8 : B);
  const size_t smem = 2 * CIN * COUT * sizeof(float);
  auto kern = conv3d_simt_kernel<T, CIN, COUT, KIND>;
  IDISP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, COUT / 8 * 32, smem, s>>>(x, w_tap, bias, residual, relu, y, D, H, W, Do, Ho, Wo);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

template <typename T>
int launch_conv3d_simt(const T *x, int B, int Cin, int D, int H, int W, const float *w_tap, int Cout, int kind,
                       const float *bias, const T *residual, int relu, T *y, cudaStream_t s)
{
  if (B == 0) return IDISP_OK;
#define IDISP_CASE(CI, CO, K)                                                             \
  if (Cin == CI && Cout == CO && kind == K)                                               \
    return launch_one<T, CI, CO, K>(x, B, D, H, W, w_tap, bias, residual, relu, y, s);
  IDISP_CASE(32, 32, IDISP_CONV_S1)
  IDISP_CASE(64, 32, IDISP_CONV_S1)
  IDISP_CASE(64, 64, IDISP_CONV_S1)
  IDISP_CASE(32, 64, IDISP_CONV_S2)
  IDISP_CASE(64, 64, IDISP_CONV_S2)
  IDISP_CASE(64, 64, IDISP_DECONV_S2)
  IDISP_CASE(64, 32, IDISP_DECONV_S2)
  // feature widths other than C=32 (BASELINE config 1 uses C=16 -> 2C=32; also 2C in {16,48,...})
  IDISP_CASE(16, 32, IDISP_CONV_S1)
  IDISP_CASE(48, 32, IDISP_CONV_S1)
  IDISP_CASE(128, 32, IDISP_CONV_S1)
#undef IDISP_CASE
  set_error("conv3d: unsupported (Cin=%d, Cout=%d, kind=%d) on the SIMT path", Cin, Cout, kind);
  return IDISP_ERR_INVALID;
}
template int launch_conv3d_simt<float>(const float *, int, int, int, int, int, const float *, int, int, const float *, const float *, int, float *, cudaStream_t);
template int launch_conv3d_simt<__nv_bfloat16>(const __nv_bfloat16 *, int, int, int, int, int, const float *, int, int, const float *, const __nv_bfloat16 *, int, __nv_bfloat16 *, cudaStream_t);

// ---------------------------------------------------------------------------------------
// 32 -> 1 classifier conv (stackhourglass.py:80,84,88: Conv3d(32,1,3,pad 1,bias=False)) with the
// running sum of the previous head fused (:143-144).  1.04 GFLOP/ROI at config 2: CUDA cores,
// one thread per output voxel, w [27][Cin] in shared memory, output/residual plain [B][D][H][W] f32.
// ---------------------------------------------------------------------------------------
// X2: x holds split-precision activations (IEEE-half hi blocks, then lo blocks: [N][2*CIN/8][V][8]); the value is hi + lo.
template <typename T, int CIN, bool X2 = false>
__global__ void __launch_bounds__(128)
conv3d_to1_kernel(const T *__restrict__ x, const float *__restrict__ w_tap, const float *__restrict__ residual,
                  float *__restrict__ y, int D, int H, int W)
{
  __shared__ float ws[27 * CIN];
  for (int i = threadIdx.x; i < 27 * CIN; i += blockDim.x) ws[i] = w_tap[i];
  __syncthreads();
  const int n = blockIdx.z, d = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int h = p / W, w = p - h * W;
  const int64_t V = (int64_t)D * H * W;
  const T *xn = x + (int64_t)n * ((X2 ? 2 : 1) * CIN / 8) * V * 8;
  float acc = 0.f;
  for (int kd = 0; kd < 3; ++kd) {
    const int di = d + kd - 1;
    if (di < 0 || di >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = h + kh - 1;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = w + kw - 1;
        if (wi < 0 || wi >= W) continue;
        const float *wt = ws + ((kd * 3 + kh) * 3 + kw) * CIN;
        const int64_t off = (((int64_t)di * H + hi) * W + wi) * 8;
#pragma unroll
        for (int cb = 0; cb < CIN / 8; ++cb) {
          F8 a;
          if (X2) {
            a = unpack8h<true>(__ldg(reinterpret_cast<const uint4 *>(xn + (int64_t)cb * V * 8 + off)));
            const F8 lo = unpack8h<true>(__ldg(reinterpret_cast<const uint4 *>(xn + (int64_t)(cb + CIN / 8) * V * 8 + off)));
#pragma unroll
            for (int c = 0; c < 8; ++c) a.v[c] += lo.v[c];
          } else {
            a = load8<T>(xn + (int64_t)cb * V * 8 + off);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) acc = fmaf(a.v[c], wt[cb * 8 + c], acc);
        }
      }
    }
  }
  const int64_t o = (int64_t)n * V + ((int64_t)d * H + h) * W + w;
  if (residual) acc += residual[o];
  y[o] = acc;
}

template <typename T>
int launch_conv3d_to1(const T *x, int B, int Cin, int D, int H, int W, const float *w_tap, const float *residual,
                      float *y, cudaStream_t s)
{
  if (B == 0) return IDISP_OK;
  if (Cin != 32) { set_error("conv3d_to1: Cin=%d unsupported (32 only)", Cin); return IDISP_ERR_INVALID; }
  dim3 grid(ceil_div(H * W, 128), D, B);
  conv3d_to1_kernel<T, 32><<<grid, 128, 0, s>>>(x, w_tap, residual, y, D, H, W);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
int launch_conv3d_to1_x2(const __nv_bfloat16 *x, int B, int Cin, int D, int H, int W, const float *w_tap, const float *residual,
                         float *y, cudaStream_t s)
{
  if (B == 0) return IDISP_OK;
  if (Cin != 32) { set_error("conv3d_to1: Cin=%d unsupported (32 only)", Cin); return IDISP_ERR_INVALID; }
  dim3 grid(ceil_div(H * W, 128), D, B);
  conv3d_to1_kernel<__nv_bfloat16, 32, true><<<grid, 128, 0, s>>>(x, w_tap, residual, y, D, H, W);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
template int launch_conv3d_to1<float>(const float *, int, int, int, int, int, const float *, const float *, float *, cudaStream_t);
template int launch_conv3d_to1<__nv_bfloat16>(const __nv_bfloat16 *, int, int, int, int, int, const float *, const float *, float *, cudaStream_t);

}  // namespace idisp
