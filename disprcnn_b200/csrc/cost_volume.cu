// cost_volume.cu -- concatenation cost volume + layout converters (HBM-bound, exact copies).
//
// Replaces disprcnn/modeling/psmnet/stackhourglass.py:115-128: a CPU-side torch.zeros, an H2D
// copy of the whole zero volume and 2*D strided slice-copy launches become ONE kernel.
// Closed form, plane k <-> shift i = k + mindisp/4:
//   valid(k,x) = x >= max(i,0) && x < W + min(i,0)
//   cost[b,   c, k,y,x] = valid ? L[b,c,y,x]   : 0
//   cost[b, C+c, k,y,x] = valid ? R[b,c,y,x-i] : 0
// Algorithmic bytes: 2C*V*sizeof(T) written per ROI (V = D*Hf*Wf); reads are L2 hits after
// the first plane.  Every thread writes one 16/32-byte channel-block voxel (blocked layout)
// or one float4 of a row (NCDHW test hook) -> 128-bit coalesced stores.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace idisp {

// ---------------- blocked output: [B][2C/8][D][H][W][8] -------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
cost_volume_blocked_kernel(const float *__restrict__ L, const float *__restrict__ R, int B, int C, int Hf,
                           int Wf, int shift0, int D, T *__restrict__ cost)
{
  const int nblk = 2 * C / CB;
  const int64_t HW = (int64_t)Hf * Wf;
  const int64_t total = (int64_t)B * nblk * D * HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % Wf);
    const int y = (int)((idx / Wf) % Hf);
    const int k = (int)((idx / HW) % D);
    const int cb = (int)((idx / (HW * D)) % nblk);
    const int b = (int)(idx / (HW * D * nblk));
    const int i = k + shift0;
    const bool valid = (x >= max(i, 0)) && (x < Wf + min(i, 0));
    F8 v;
    if (valid) {
      const bool right = cb >= C / CB;
      const int c0 = (right ? cb - C / CB : cb) * CB;
      const float *src = (right ? R : L) + ((int64_t)b * C + c0) * HW + (int64_t)y * Wf + (right ? x - i : x);
#pragma unroll
      for (int c = 0; c < CB; ++c) v.v[c] = __ldg(src + c * HW);
    } else {
#pragma unroll
      for (int c = 0; c < CB; ++c) v.v[c] = 0.f;
    }
    store8<T>(cost + idx * CB, v);
  }
}

template <typename T>
int launch_cost_volume_blocked(const float *L, const float *R, int B, int C, int Hf, int Wf, int mindisp, int D,
                               T *cost, cudaStream_t s)
{
  const int64_t total = (int64_t)B * (2 * C / CB) * D * Hf * Wf;
  if (total == 0) return IDISP_OK;
  const int64_t want = ceil_div64(total, 256);
  const int grid = (int)(want < 148 * 32 ? want : 148 * 32);
  // Python floor division of a (possibly negative) multiple of 4
  const int shift0 = mindisp >= 0 ? mindisp / 4 : -((-mindisp + 3) / 4);
  cost_volume_blocked_kernel<T><<<grid, 256, 0, s>>>(L, R, B, C, Hf, Wf, shift0, D, cost);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
template int launch_cost_volume_blocked<float>(const float *, const float *, int, int, int, int, int, int, float *, cudaStream_t);
template int launch_cost_volume_blocked<__nv_bfloat16>(const float *, const float *, int, int, int, int, int, int, __nv_bfloat16 *, cudaStream_t);

// ---------------- NCDHW f32 output (the reference's layout; C-ABI entry idisp_cost_volume) -----------------
// One CTA owns RY feature rows of one (b, c): the rows of BOTH views are contiguous in NCHW, so two linear TMA copies
// (cp.async.bulk, one mbarrier) stage them in shared memory once; the CTA then writes those rows into all D planes of
// channels c (left, unshifted) and C+c (right, shifted by the plane's disparity, read from shared memory) with 128-bit
// stores -- RY*Wf*4 contiguous bytes per plane and channel.  Algorithmic bytes: 2C*V*4 written per ROI (154 MB at config 2).
constexpr int CV_RY = 8;

__global__ void __launch_bounds__(256)
cost_volume_ncdhw_tma_kernel(const float *__restrict__ L, const float *__restrict__ R, int C, int Hf, int Wf, int shift0, int D,
                             float *__restrict__ cost)
{
  extern __shared__ __align__(128) float cv_smem[];  // [2][RY*Wf] + mbarrier
  const int tiles_y = (Hf + CV_RY - 1) / CV_RY;
  const int ty = blockIdx.x % tiles_y, c = (blockIdx.x / tiles_y) % C, b = blockIdx.x / (tiles_y * C);
  const int y0 = ty * CV_RY, rows = min(CV_RY, Hf - y0);
  const int n = rows * Wf;  // floats per view
  float *sl = cv_smem, *sr = cv_smem + CV_RY * Wf;
  const uint32_t bar = ptx::smem_u32(cv_smem + 2 * CV_RY * Wf);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t src = (((int64_t)b * C + c) * Hf + y0) * Wf;
    ptx::mbar_arrive_expect_tx(bar, 2u * n * 4u);
    ptx::bulk_g2s(ptx::smem_u32(sl), L + src, n * 4u, bar);
    ptx::bulk_g2s(ptx::smem_u32(sr), R + src, n * 4u, bar);
  }
  ptx::mbar_wait(bar, 0);
  const int w4 = Wf / 4, per_plane = rows * w4;
  const int64_t HW = (int64_t)Hf * Wf;
  float *outl = cost + (((int64_t)b * 2 * C + c) * D) * HW + (int64_t)y0 * Wf;
  float *outr = cost + (((int64_t)b * 2 * C + C + c) * D) * HW + (int64_t)y0 * Wf;
  for (int it = threadIdx.x; it < D * per_plane; it += 256) {
    const int k = it / per_plane, rem = it - k * per_plane;
    const int yr = rem / w4, x = (rem - yr * w4) * 4;
    const int i = k + shift0, lo = max(i, 0), hi = Wf + min(i, 0);  // valid x in [lo, hi)
    const float *pl = sl + yr * Wf + x, *pr = sr + yr * Wf + x - i;
    float4 vl, vr;
    vl.x = (x + 0 >= lo && x + 0 < hi) ? pl[0] : 0.f; vr.x = (x + 0 >= lo && x + 0 < hi) ? pr[0] : 0.f;
    vl.y = (x + 1 >= lo && x + 1 < hi) ? pl[1] : 0.f; vr.y = (x + 1 >= lo && x + 1 < hi) ? pr[1] : 0.f;
    vl.z = (x + 2 >= lo && x + 2 < hi) ? pl[2] : 0.f; vr.z = (x + 2 >= lo && x + 2 < hi) ? pr[2] : 0.f;
    vl.w = (x + 3 >= lo && x + 3 < hi) ? pl[3] : 0.f; vr.w = (x + 3 >= lo && x + 3 < hi) ? pr[3] : 0.f;
    const int64_t o = (int64_t)k * HW + (int64_t)yr * Wf + x;
    __stcs(reinterpret_cast<float4 *>(outl + o), vl);  // streaming: written once, read by the next layer from HBM
    __stcs(reinterpret_cast<float4 *>(outr + o), vr);
  }
}

// any width (Wf % 4 != 0 breaks the 16-byte alignment of rows): one element per thread
__global__ void __launch_bounds__(256)
cost_volume_ncdhw_kernel(const float *__restrict__ L, const float *__restrict__ R, int B, int C, int Hf,
                         int Wf, int shift0, int D, float *__restrict__ cost)
{
  const int64_t HW = (int64_t)Hf * Wf;
  const int64_t total = (int64_t)B * 2 * C * D * HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % Wf);
    const int y = (int)((idx / Wf) % Hf);
    const int k = (int)((idx / HW) % D);
    const int c = (int)((idx / (HW * D)) % (2 * C));
    const int b = (int)(idx / (HW * D * 2 * C));
    const int i = k + shift0;
    const bool valid = (x >= max(i, 0)) && (x < Wf + min(i, 0));
    float v = 0.f;
    if (valid) {
      v = c < C ? __ldg(L + ((int64_t)b * C + c) * HW + (int64_t)y * Wf + x)
                : __ldg(R + ((int64_t)b * C + (c - C)) * HW + (int64_t)y * Wf + (x - i));
    }
    cost[idx] = v;
  }
}

// ---------------- layout converters NCDHW f32 <-> blocked T ----------------------------
template <typename T>
__global__ void __launch_bounds__(256)
ncdhw_to_blocked_kernel(const float *__restrict__ src, T *__restrict__ dst, int B, int C, int64_t V)
{
  const int nblk = C / CB;
  const int64_t total = (int64_t)B * nblk * V;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx % V;
    const int cb = (int)((idx / V) % nblk);
    const int b = (int)(idx / (V * nblk));
    const float *s = src + ((int64_t)b * C + cb * CB) * V + v;
    F8 r;
#pragma unroll
    for (int c = 0; c < CB; ++c) r.v[c] = __ldg(s + c * V);
    store8<T>(dst + idx * CB, r);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
blocked_to_ncdhw_kernel(const T *__restrict__ src, float *__restrict__ dst, int B, int C, int64_t V)
{
  const int nblk = C / CB;
  const int64_t total = (int64_t)B * nblk * V;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx % V;
    const int cb = (int)((idx / V) % nblk);
    const int b = (int)(idx / (V * nblk));
    const F8 r = load8<T>(src + idx * CB);
    float *d = dst + ((int64_t)b * C + cb * CB) * V + v;
#pragma unroll
    for (int c = 0; c < CB; ++c) d[c * V] = r.v[c];
  }
}

template <bool F16>
__global__ void __launch_bounds__(256)
ncdhw_to_blocked_h_kernel(const float *__restrict__ src, uint4 *__restrict__ dst, int B, int C, int64_t V)
{
  const int nblk = C / CB;
  const int64_t total = (int64_t)B * nblk * V;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx % V;
    const int cb = (int)((idx / V) % nblk);
    const int b = (int)(idx / (V * nblk));
    const float *s = src + ((int64_t)b * C + cb * CB) * V + v;
    F8 r;
#pragma unroll
    for (int c = 0; c < CB; ++c) r.v[c] = __ldg(s + c * V);
    dst[idx] = pack8h<F16>(r);
  }
}
template <bool F16>
__global__ void __launch_bounds__(256)
blocked_to_ncdhw_h_kernel(const uint4 *__restrict__ src, float *__restrict__ dst, int B, int C, int64_t V)
{
  const int nblk = C / CB;
  const int64_t total = (int64_t)B * nblk * V;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx % V;
    const int cb = (int)((idx / V) % nblk);
    const int b = (int)(idx / (V * nblk));
    const F8 r = unpack8h<F16>(__ldg(src + idx));
    float *d = dst + ((int64_t)b * C + cb * CB) * V + v;
#pragma unroll
    for (int c = 0; c < CB; ++c) d[c * V] = r.v[c];
  }
}
// split precision: [B][2*C/8][V][8] 16-bit words -- blocks [0,C/8) hold hi = half(x), blocks [C/8, 2C/8) hold lo = half(x - hi)
__global__ void __launch_bounds__(256)
ncdhw_to_blocked_x2_kernel(const float *__restrict__ src, uint4 *__restrict__ dst, int B, int C, int64_t V, int *__restrict__ range_flag)
{
  const int nblk = C / CB;
  const int64_t total = (int64_t)B * nblk * V;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx % V;
    const int cb = (int)((idx / V) % nblk);
    const int b = (int)(idx / (V * nblk));
    const float *s = src + ((int64_t)b * C + cb * CB) * V + v;
    F8 r;
#pragma unroll
    for (int c = 0; c < CB; ++c) r.v[c] = __ldg(s + c * V);
    if (range_flag) {  // a value outside the IEEE-half range (or non-finite) cannot be split into hi + lo words
      bool bad = false;
#pragma unroll
      for (int c = 0; c < CB; ++c) bad |= !(fabsf(r.v[c]) <= 65504.f);
      if (bad) *range_flag = 1;
    }
    const uint4 hi = pack8h<true>(r);
    const F8 h = unpack8h<true>(hi);
#pragma unroll
    for (int c = 0; c < CB; ++c) r.v[c] -= h.v[c];
    const int64_t o = ((int64_t)b * 2 * nblk + cb) * V + v;
    dst[o] = hi;
    dst[o + (int64_t)nblk * V] = pack8h<true>(r);
  }
}
__global__ void __launch_bounds__(256)
blocked_x2_to_ncdhw_kernel(const uint4 *__restrict__ src, float *__restrict__ dst, int B, int C, int64_t V)
{
  const int nblk = C / CB;
  const int64_t total = (int64_t)B * nblk * V;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx % V;
    const int cb = (int)((idx / V) % nblk);
    const int b = (int)(idx / (V * nblk));
    const int64_t o = ((int64_t)b * 2 * nblk + cb) * V + v;
    const F8 h = unpack8h<true>(__ldg(src + o)), l = unpack8h<true>(__ldg(src + o + (int64_t)nblk * V));
    float *d = dst + ((int64_t)b * C + cb * CB) * V + v;
#pragma unroll
    for (int c = 0; c < CB; ++c) d[c * V] = h.v[c] + l.v[c];
  }
}
int launch_ncdhw_to_blocked_x2(const float *src, __nv_bfloat16 *dst, int B, int C, int64_t V, cudaStream_t s, int *range_flag)
{
  const int64_t total = (int64_t)B * (C / CB) * V;
  if (total == 0) return IDISP_OK;
  const int64_t want = ceil_div64(total, 256);
  ncdhw_to_blocked_x2_kernel<<<(int)(want < 148 * 32 ? want : 148 * 32), 256, 0, s>>>(src, (uint4 *)dst, B, C, V, range_flag);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
int launch_blocked_x2_to_ncdhw(const __nv_bfloat16 *src, float *dst, int B, int C, int64_t V, cudaStream_t s)
{
  const int64_t total = (int64_t)B * (C / CB) * V;
  if (total == 0) return IDISP_OK;
  const int64_t want = ceil_div64(total, 256);
  blocked_x2_to_ncdhw_kernel<<<(int)(want < 148 * 32 ? want : 148 * 32), 256, 0, s>>>((const uint4 *)src, dst, B, C, V);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

int launch_ncdhw_to_blocked_h(const float *src, __nv_bfloat16 *dst, int B, int C, int64_t V, int f16, cudaStream_t s)
{
  const int64_t total = (int64_t)B * (C / CB) * V;
  if (total == 0) return IDISP_OK;
  const int64_t want = ceil_div64(total, 256);
  const int grid = (int)(want < 148 * 32 ? want : 148 * 32);
  if (f16) ncdhw_to_blocked_h_kernel<true><<<grid, 256, 0, s>>>(src, (uint4 *)dst, B, C, V);
  else ncdhw_to_blocked_h_kernel<false><<<grid, 256, 0, s>>>(src, (uint4 *)dst, B, C, V);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
int launch_blocked_to_ncdhw_h(const __nv_bfloat16 *src, float *dst, int B, int C, int64_t V, int f16, cudaStream_t s)
{
  const int64_t total = (int64_t)B * (C / CB) * V;
  if (total == 0) return IDISP_OK;
  const int64_t want = ceil_div64(total, 256);
  const int grid = (int)(want < 148 * 32 ? want : 148 * 32);
  if (f16) blocked_to_ncdhw_h_kernel<true><<<grid, 256, 0, s>>>((const uint4 *)src, dst, B, C, V);
  else blocked_to_ncdhw_h_kernel<false><<<grid, 256, 0, s>>>((const uint4 *)src, dst, B, C, V);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

template <typename T>
int launch_ncdhw_to_blocked(const float *src, T *dst, int B, int C, int64_t V, cudaStream_t s)
{
  const int64_t total = (int64_t)B * (C / CB) * V;
  if (total == 0) return IDISP_OK;
  const int64_t want = ceil_div64(total, 256);
  ncdhw_to_blocked_kernel<T><<<(int)(want < 148 * 32 ? want : 148 * 32), 256, 0, s>>>(src, dst, B, C, V);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
template <typename T>
int launch_blocked_to_ncdhw(const T *src, float *dst, int B, int C, int64_t V, cudaStream_t s)
{
  const int64_t total = (int64_t)B * (C / CB) * V;
  if (total == 0) return IDISP_OK;
  const int64_t want = ceil_div64(total, 256);
  blocked_to_ncdhw_kernel<T><<<(int)(want < 148 * 32 ? want : 148 * 32), 256, 0, s>>>(src, dst, B, C, V);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
template int launch_ncdhw_to_blocked<float>(const float *, float *, int, int, int64_t, cudaStream_t);
template int launch_ncdhw_to_blocked<__nv_bfloat16>(const float *, __nv_bfloat16 *, int, int, int64_t, cudaStream_t);
template int launch_blocked_to_ncdhw<float>(const float *, float *, int, int, int64_t, cudaStream_t);
template int launch_blocked_to_ncdhw<__nv_bfloat16>(const __nv_bfloat16 *, float *, int, int, int64_t, cudaStream_t);

}  // namespace idisp

extern "C" int idisp_cost_volume(const float *left, const float *right, int B, int C, int Hf, int Wf,
                                 int mindisp, int maxdisp, float *cost, void *stream)
{
  using namespace idisp;
  IDISP_REQUIRE(B >= 0 && C > 0 && Hf > 0 && Wf > 0, "cost_volume: bad shape B=%d C=%d Hf=%d Wf=%d", B, C, Hf, Wf);
  IDISP_REQUIRE(maxdisp > mindisp && mindisp % 4 == 0 && maxdisp % 4 == 0,
                "cost_volume: mindisp=%d maxdisp=%d must be multiples of 4 with maxdisp>mindisp", mindisp, maxdisp);
  const int D = (maxdisp - mindisp) / 4;
  const int shift0 = mindisp >= 0 ? mindisp / 4 : -((-mindisp + 3) / 4);
  IDISP_REQUIRE(-shift0 < Wf + 1 && shift0 + D - 1 < Wf + 1, "cost_volume: |shift| exceeds feature width %d", Wf);
  if (B == 0) return IDISP_OK;
  IDISP_REQUIRE(left && right && cost, "cost_volume: NULL pointer");
  const bool aligned = Wf % 4 == 0 && (((uintptr_t)left | (uintptr_t)right | (uintptr_t)cost) & 15) == 0;
  if (aligned && !getenv("IDISP_CV_SCALAR")) {
    const int tiles_y = (Hf + CV_RY - 1) / CV_RY;
    const size_t smem = (size_t)2 * CV_RY * Wf * 4 + 16;
    if (smem <= 200 * 1024) {
      IDISP_CUDA(cudaFuncSetAttribute(cost_volume_ncdhw_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cost_volume_ncdhw_tma_kernel<<<B * C * tiles_y, 256, smem, (cudaStream_t)stream>>>(left, right, C, Hf, Wf, shift0, D, cost);
      IDISP_LAUNCH_CHECK();
      return IDISP_OK;
    }
  }
  const int64_t total = (int64_t)B * 2 * C * D * Hf * Wf;
  const int64_t want = ceil_div64(total, 256);
  cost_volume_ncdhw_kernel<<<(int)(want < 148 * 32 ? want : 148 * 32), 256, 0, (cudaStream_t)stream>>>(
      left, right, B, C, Hf, Wf, shift0, D, cost);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}
