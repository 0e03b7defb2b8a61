// common.cuh -- shared helpers for libidisp (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>

#include "../../include/idisp.h"

namespace idisp {

// Internal activation layout ("channel-blocked-8"): [N][C/8][D][H][W][8] with element type T
// (float in IDISP_PREC_FP32, __nv_bfloat16 in IDISP_PREC_BF16).  One voxel of one channel
// block is 32 B (f32) / 16 B (bf16): the unit every kernel loads, stores and -- on the tensor
// path -- the 16-byte row of a no-swizzle K-major UMMA core matrix.
constexpr int CB = 8;

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define IDISP_CUDA(expr)                                                     \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) return ::idisp::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define IDISP_LAUNCH_CHECK() IDISP_CUDA(cudaGetLastError())

#define IDISP_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::idisp::set_error(__VA_ARGS__);      \
      return IDISP_ERR_INVALID;             \
    }                                       \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- element conversion --------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 channels of one voxel (one channel block), as floats in registers
struct F8 { float v[8]; };

template <typename T> __device__ __forceinline__ F8 load8(const T *p);
template <> __device__ __forceinline__ F8 load8<float>(const float *p) {
  const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
  const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
  F8 r; r.v[0]=a.x; r.v[1]=a.y; r.v[2]=a.z; r.v[3]=a.w; r.v[4]=b.x; r.v[5]=b.y; r.v[6]=b.z; r.v[7]=b.w;
  return r;
}
template <> __device__ __forceinline__ F8 load8<__nv_bfloat16>(const __nv_bfloat16 *p) {
  const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
  F8 r;
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}
template <typename T> __device__ __forceinline__ void store8(T *p, const F8 &r);
template <> __device__ __forceinline__ void store8<float>(float *p, const F8 &r) {
  reinterpret_cast<float4 *>(p)[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  reinterpret_cast<float4 *>(p)[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16 *p, const F8 &r) {
  uint4 a;
  uint32_t *w = reinterpret_cast<uint32_t *>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t *>(&h);
  }
  *reinterpret_cast<uint4 *>(p) = a;
}

// 16-bit storage in either format behind the same (opaque) pointer type: the tensor-core path stores bf16 or fp16
// (IDISP_PREC_BF16 / IDISP_PREC_FP16); F16 selects the conversion.  Buffers are typed __nv_bfloat16* in both cases.
template <bool F16> __device__ __forceinline__ F8 unpack8h(const uint4 &a)
{
  F8 r;
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (F16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
      r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
    } else {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  return r;
}
template <bool F16> __device__ __forceinline__ uint4 pack8h(const F8 &r)
{
  uint4 a;
  uint32_t *w = reinterpret_cast<uint32_t *>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (F16) {
      const __half2 h = __floats2half2_rn(r.v[2 * i], r.v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t *>(&h);
    } else {
      const __nv_bfloat162 h = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t *>(&h);
    }
  }
  return a;
}

// ---- kernels' host launchers (defined in the .cu files) ------------------------------
// NCDHW f32 <-> blocked 16-bit storage (f16 = 1: IEEE half, 0: bfloat16)
int launch_ncdhw_to_blocked_h(const float *src, __nv_bfloat16 *dst, int B, int C, int64_t V, int f16, cudaStream_t s);
int launch_blocked_to_ncdhw_h(const __nv_bfloat16 *src, float *dst, int B, int C, int64_t V, int f16, cudaStream_t s);
// split precision (two IEEE-half words per value): [B][2*C/8][V][8], hi blocks then lo blocks
// range_flag (device int, optional): set to 1 if a value does not fit the IEEE-half range
int launch_ncdhw_to_blocked_x2(const float *src, __nv_bfloat16 *dst, int B, int C, int64_t V, cudaStream_t s, int *range_flag = nullptr);
int launch_blocked_x2_to_ncdhw(const __nv_bfloat16 *src, float *dst, int B, int C, int64_t V, cudaStream_t s);
// layout converters (test hooks + plan I/O)
template <typename T>
int launch_ncdhw_to_blocked(const float *src, T *dst, int B, int C, int64_t V, cudaStream_t s);
template <typename T>
int launch_blocked_to_ncdhw(const T *src, float *dst, int B, int C, int64_t V, cudaStream_t s);

template <typename T>
int launch_cost_volume_blocked(const float *L, const float *R, int B, int C, int Hf, int Wf, int mindisp,
                               int D, T *cost, cudaStream_t s);

// SIMT conv (conv3d_simt.cu).  w: [27][Cin][Cout] f32 (tap-major, cout innermost), scale folded in.
template <typename T>
int launch_conv3d_simt(const T *x, int B, int Cin, int D, int H, int W, const float *w_tap, int Cout,
                       int kind, const float *bias, const T *residual, int relu, T *y, cudaStream_t s);
// 32->1 classifier conv; w: [27][32] f32; in blocked T, out/residual [B][D][H][W] f32
template <typename T>
int launch_conv3d_to1(const T *x, int B, int Cin, int D, int H, int W, const float *w_tap,
                      const float *residual, float *y, cudaStream_t s);

// the same on split-precision activations (hi|lo IEEE-half block groups), f32 weights, f32 FMA
int launch_conv3d_to1_x2(const __nv_bfloat16 *x, int B, int Cin, int D, int H, int W, const float *w_tap, const float *residual,
                         float *y, cudaStream_t s);

// cellmax_scratch: optional B*Hf*Wf floats of device scratch (enables the one-pass form, see softargmin.cu)
int launch_softargmin(const float *logits, int B, int D, int Hf, int Wf, int mindisp, int maxdisp, int H,
                      int W, float *out, cudaStream_t s, float *cellmax_scratch = nullptr);

// Launch with the optional attributes the tensor-core kernels use: a cluster of two CTAs, and programmatic dependent launch (the
// kernel may start its prologue -- barriers, TMEM allocation, weight copies -- while the previous kernel in the stream drains; it
// blocks at griddepcontrol.wait before touching anything that kernel produced).  Callers ask for it for SMALL launches only (at most
// two rounds of work per SM): measured at the KITTI shape it shortens a forward by 8 % (R = 1) to 2 % (R = 8) -- a launch there is
// 20-30 us, a third of it prologue -- while at 32 ROI pairs of the benchmark shape it costs 2 % (the early CTAs of the next kernel
// compete with the running one).  IDISP_NO_PDL=1 switches it off altogether.
inline bool pdl_enabled()
{
  static const int off = [] { const char *e = getenv("IDISP_NO_PDL"); return (e && e[0] == '1') ? 1 : 0; }();
  return !off;
}
template <typename K, typename... Args>
inline cudaError_t launch_ex(K kern, int grid, int block, size_t smem, cudaStream_t s, bool cluster2, bool pdl, Args... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster2) { at[n].id = cudaLaunchAttributeClusterDimension; at[n].val.clusterDim.x = 2; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n; }
  if (pdl && pdl_enabled()) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

}  // namespace idisp
