// feature2d.cu -- the 2-D feature extractor of iDispNet (disprcnn/modeling/psmnet/submodule.py:60-139: firstconv, layer1-4 of
// BasicBlocks (:25-48), four SPP branches (avg-pool -> 1x1 convbn -> ReLU -> bilinear align_corners upsample), concat,
// lastconv) on the CUDA cores in fp32 -- SURVEY.md section 8(f) row 1.  In the live call (stackhourglass.py:112-113) it runs once
// per view before the cost volume; with it the whole PSMNet.forward executes inside libidisp (no cuDNN / ATen kernels).
//
// Layout: plain NCHW f32 (the layout of the image crops that come in and of the [B,32,H/4,W/4] features the 3-D stack's
// entry point takes).  BatchNorm2d (eval, eps 1e-5) is folded in float64 into the kernel (per-Cout scale) and a bias, like the
// 3-D layers (plan.cu).  One generic direct-convolution kernel covers every conv of the extractor:
//   k in {1, 3}, stride in {1, 2}, dilation in {1, 2} (padding = dilation for k = 3, 0 for k = 1: submodule.py:13-16,
//   downsample :103-105), fused bias (+ residual) (+ ReLU), input / output addressed with their own batch strides so that
//   `raw` and `skip` live INSIDE the 320-channel concat tensor (no copy for torch.cat, submodule.py:134-135).
// Tiling: a CTA computes 32 x 8 output pixels x COB output channels; a thread owns 4 neighbouring pixels x COB/4 channels
// (64 or 32 accumulators); input patch and weight slice of 8 input channels are staged in shared memory per step.
// Roofline: fp32 FFMA (CUDA cores; ~75 TFLOP/s peak): 22.19 GFLOP per 224 x 224 crop.
#include <map>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "conv2d_tc.cuh"

namespace idisp {
namespace f2d {

constexpr int TW = 32, TH = 8, CI = 16, PXT = 4;   // output tile, input channels per step, pixels per thread

struct ConvParams {
  const float *x, *w, *bias, *res;
  float *y;
  int Cin, Cout, H, W, Ho, Wo, stride, dil, pad, relu;
  long long xbs, ybs, rbs;   // batch strides (elements) of input, output, residual
  int tiles_w, tiles_h;
};

// w: [Cin][K*K][Cout] f32 (BN scale folded in)
template <int K, int COB>
__global__ void __launch_bounds__(256, 2) conv2d_kernel(const ConvParams p)
{
  constexpr int CPT = COB / 4;                       // output channels per thread
  extern __shared__ float smem[];
  const int PH = (TH - 1) * p.stride + (K - 1) * p.dil + 1, PW = (TW - 1) * p.stride + (K - 1) * p.dil + 1;
  const int PWp = PW | 1;                            // odd row pitch: fewer bank conflicts for the stride-2 reads
  float *s_in = smem;                                // [CI][PH][PWp]
  float *s_w = smem + CI * PH * PWp;                 // [CI][K*K][COB]
  const int tid = threadIdx.x;
  const int pg = tid & 63, cg = tid >> 6;
  const int px0 = (pg & 7) * PXT, py = pg >> 3;
  const int tile = blockIdx.x, tw = tile % p.tiles_w, th = tile / p.tiles_w;
  const int cb = blockIdx.y * COB;                   // first output channel of this CTA
  const int n = blockIdx.z;
  const int ox0 = tw * TW, oy0 = th * TH;
  const int ix0 = ox0 * p.stride - p.pad, iy0 = oy0 * p.stride - p.pad;
  const float *xn = p.x + (long long)n * p.xbs;
  float acc[PXT][CPT];
#pragma unroll
  for (int i = 0; i < PXT; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;

  for (int c0 = 0; c0 < p.Cin; c0 += CI) {
    const int nci = min(CI, p.Cin - c0);
    __syncthreads();
    // input patch (zero outside the image = conv padding; zero for the channels beyond Cin): one warp per (channel, row),
    // lanes along the row -- coalesced, and no integer division per element
    for (int rowi = tid >> 5; rowi < nci * PH; rowi += 8) {   // (only the channels that exist: firstconv.0 has 3 of the 16)
      const int ci = rowi / PH, yy = rowi - ci * PH;
      const int gy = iy0 + yy;
      const bool rok = ci < nci && gy >= 0 && gy < p.H;
      const float *src = xn + ((long long)(c0 + ci) * p.H + gy) * p.W;
      float *dst = s_in + rowi * PWp;
      for (int xx = tid & 31; xx < PW; xx += 32) {
        const int gx = ix0 + xx;
        dst[xx] = (rok && gx >= 0 && gx < p.W) ? __ldg(src + gx) : 0.f;
      }
    }
    for (int i = tid; i < nci * K * K * COB; i += 256) {
      const int co = i % COB, t = (i / COB) % (K * K), ci = i / (COB * K * K);
      float v = 0.f;
      if (ci < nci && cb + co < p.Cout) v = __ldg(p.w + ((long long)(c0 + ci) * K * K + t) * p.Cout + cb + co);
      s_w[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < nci; ++ci) {
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const float *row = s_in + (ci * PH + py * p.stride + kh * p.dil) * PWp + px0 * p.stride;
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          float xv[PXT];
#pragma unroll
          for (int i = 0; i < PXT; ++i) xv[i] = row[i * p.stride + kw * p.dil];
          const float4 *wp = reinterpret_cast<const float4 *>(s_w + (ci * K * K + kh * K + kw) * COB + cg * CPT);
#pragma unroll
          for (int c4 = 0; c4 < CPT / 4; ++c4) {
            const float4 wv = wp[c4];
#pragma unroll
            for (int i = 0; i < PXT; ++i) {
              acc[i][c4 * 4 + 0] = fmaf(xv[i], wv.x, acc[i][c4 * 4 + 0]);
              acc[i][c4 * 4 + 1] = fmaf(xv[i], wv.y, acc[i][c4 * 4 + 1]);
              acc[i][c4 * 4 + 2] = fmaf(xv[i], wv.z, acc[i][c4 * 4 + 2]);
              acc[i][c4 * 4 + 3] = fmaf(xv[i], wv.w, acc[i][c4 * 4 + 3]);
            }
          }
        }
      }
    }
  }
  const int oy = oy0 + py;
  if (oy >= p.Ho) return;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int co = cb + cg * CPT + c;
    if (co >= p.Cout) continue;
    const float b = p.bias ? __ldg(p.bias + co) : 0.f;
#pragma unroll
    for (int i = 0; i < PXT; ++i) {
      const int ox = ox0 + px0 + i;
      if (ox >= p.Wo) continue;
      const long long o = ((long long)co * p.Ho + oy) * p.Wo + ox;
      float v = acc[i][c] + b;
      if (p.res) v += __ldg(p.res + (long long)n * p.rbs + o);
      if (p.relu) v = fmaxf(v, 0.f);
      p.y[(long long)n * p.ybs + o] = v;
    }
  }
}

// AvgPool2d(k, stride k) (submodule.py:78-92: kernel = stride, no padding, floor): one WARP per output element (the 56 x 56
// window of branch1 is 3136 elements), lanes stride over the window, shuffle reduction
__global__ void avgpool_kernel(const float *__restrict__ x, long long xbs, int C, int H, int W, int k, int Ho, int Wo, float *__restrict__ y)
{
  const int n = blockIdx.y, lane = threadIdx.x & 31;
  const int nwarp = gridDim.x * (blockDim.x >> 5);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < C * Ho * Wo; i += nwarp) {
    const int ox = i % Wo, oy = (i / Wo) % Ho, c = i / (Wo * Ho);
    const float *src = x + (long long)n * xbs + ((long long)c * H + oy * k) * W + ox * k;
    float s = 0.f;
    for (int e = lane; e < k * k; e += 32) s += __ldg(src + (long long)(e / k) * W + e % k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[((long long)n * C + c) * Ho * Wo + oy * Wo + ox] = s / (float)(k * k);
  }
}

// 1x1 conv (any stride) + bias (+ residual) (+ ReLU) straight from global memory: one thread per output pixel and group of 16 output
// channels, x read once per input channel (coalesced along the row; stride 2 uses every other element of its sectors), the 16 weights
// of a step are one address for the whole warp.  The tiled kernel above stages the full-resolution patch of a stride-2 1x1 conv in
// shared memory (four times the pixels it uses): 138 us for the 0.2 GFLOP of layer2.0.downsample at 16 images, this one ~15 us.
__global__ void __launch_bounds__(256) pointwise_kernel(const float *__restrict__ x, long long xbs, const float *__restrict__ w, const float *__restrict__ bias,
                                                          const float *__restrict__ res, long long rbs, float *__restrict__ y, long long ybs, int Cin, int Cout,
                                                          int H, int W, int Ho, int Wo, int stride, int relu)
{
  const int n = blockIdx.z, g = blockIdx.y;   // image, group of 16 output channels
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= Ho * Wo) return;
  const int oy = pix / Wo, ox = pix - oy * Wo;
  const float *xp = x + (long long)n * xbs + (long long)(oy * stride) * W + ox * stride;
  const float4 *wp = reinterpret_cast<const float4 *>(w + g * 16);
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll 4
  for (int ci = 0; ci < Cin; ++ci) {
    const float xv = __ldg(xp + (long long)ci * H * W);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 wv = __ldg(wp + (long long)ci * (Cout / 4) + q);
      acc[q * 4 + 0] = fmaf(xv, wv.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(xv, wv.y, acc[q * 4 + 1]);
      acc[q * 4 + 2] = fmaf(xv, wv.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(xv, wv.w, acc[q * 4 + 3]);
    }
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int co = g * 16 + c;
    const long long o = ((long long)co * Ho + oy) * Wo + ox;
    float v = acc[c] + (bias ? __ldg(bias + co) : 0.f);
    if (res) v += __ldg(res + (long long)n * rbs + o);
    if (relu) v = fmaxf(v, 0.f);
    y[(long long)n * ybs + o] = v;
  }
}

// 1x1 conv + bias (+ ReLU) on a handful of pixels (the SPP branches: 128 -> 32 channels on 1 .. 49 pixels per image): one thread
// per output element, weights [Cin][Cout] read coalesced across the output channels
__global__ void pointwise_small_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, int Cin, int Cout,
                                       int HW, int relu, float *__restrict__ y)
{
  // one WARP per output element, the lanes split the input channels (a serial 128-long chain of dependent loads per thread took
  // 47 us for a few hundred outputs), shuffle reduction
  const int n = blockIdx.y, lane = threadIdx.x & 31;
  const int nwarp = gridDim.x * (blockDim.x >> 5);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < Cout * HW; i += nwarp) {
    const int co = i % Cout, pos = i / Cout;
    const float *xp = x + (long long)n * Cin * HW + pos;
    float a = 0.f;
    for (int ci = lane; ci < Cin; ci += 32) a = fmaf(__ldg(xp + (long long)ci * HW), __ldg(w + (long long)ci * Cout + co), a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      a += bias ? __ldg(bias + co) : 0.f;
      y[((long long)n * Cout + co) * HW + pos] = relu ? fmaxf(a, 0.f) : a;
    }
  }
}

// F.interpolate(mode='bilinear', align_corners=True) (submodule.py:115-132) of [B,C,Hi,Wi] into channels [c_off, c_off+C) of
// the concat tensor [B,Ctot,Ho,Wo].  ATen's index math: scale = (in-1)/(out-1) (0 when out == 1), src = scale*o,
// i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0.
__global__ void upsample_bilinear_kernel(const float *__restrict__ x, int C, int Hi, int Wi, int Ho, int Wo, float *__restrict__ y, long long ybs,
                                         int c_off)
{
  const int n = blockIdx.y;
  const float sh = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f, sw = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * Ho * Wo; i += gridDim.x * blockDim.x) {
    const int ox = i % Wo, oy = (i / Wo) % Ho, c = i / (Wo * Ho);
    const float fy = sh * oy, fx = sw * ox;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float *src = x + ((long long)n * C + c) * Hi * Wi;
    const float v = hy * (hx * __ldg(src + y0 * Wi + x0) + lx * __ldg(src + y0 * Wi + x1)) +
                    ly * (hx * __ldg(src + y1 * Wi + x0) + lx * __ldg(src + y1 * Wi + x1));
    y[(long long)n * ybs + ((long long)(c_off + c) * Ho + oy) * Wo + ox] = v;
  }
}

}  // namespace f2d

// ---------------------------------------------------------------------------------------
// host side: the layer table, BN folding, the forward schedule
// ---------------------------------------------------------------------------------------
struct F2dLayer {
  std::string prefix;   // state_dict prefix relative to feature_extraction ("" prefix handled by the caller)
  int cin, cout, k, stride, dil;
  bool bn;              // prefix.0.weight + prefix.1.* (convbn) / downsample: prefix.0.weight + prefix.1.* too / bare conv: prefix.weight
  float *w = nullptr, *bias = nullptr;   // device: [Cin][k*k][Cout], [Cout]
  C2dWeights tc;                         // tensor-core packing (3x3 stride-1 layers with Cin a multiple of 32, precision fp16x2)
};

}  // namespace idisp

using namespace idisp;

struct idisp_extractor {
  std::vector<F2dLayer> layers;
  std::map<std::string, int> index;                  // prefix -> layer
  std::map<std::string, std::vector<float>> host;    // reference-keyed tensors
  float *blob = nullptr;
  bool finalized = false;
  int launches = 0;
  // IDISP_PREC_FP16X2 (default): the 53 stride-1 3x3 convs run on tcgen05 in split precision (conv2d_tc.cu), the rest on the
  // fp32 FFMA kernels below; IDISP_PREC_FP32: everything on the FFMA kernels
  int precision = IDISP_PREC_FP16X2;
  int *range_flag = nullptr;   // device int: a value left the IEEE-half range in the last fp16x2 forward
  // CUDA-graph replay of the tensor-core path's middle section (everything between the first conv, which reads the caller's
  // images, and the last converter, which writes the caller's features, touches only the workspace): ~90 launches, each with a
  // host-side tensor-map encode, become one cudaGraphLaunch per (B, H, W, workspace)
  struct GraphEntry { int B, H, W; void *ws; cudaGraphExec_t exec; int launches; unsigned long long stamp; };
  std::vector<GraphEntry> graphs;
  unsigned long long graph_clock = 0;
  cudaStream_t cap_stream = nullptr;
  bool no_graph = false;
};

static void f2d_add(idisp_extractor *e, const std::string &prefix, int cin, int cout, int k, int stride, int dil, bool bn)
{
  F2dLayer L;
  L.prefix = prefix; L.cin = cin; L.cout = cout; L.k = k; L.stride = stride; L.dil = dil; L.bn = bn;
  e->index[prefix] = (int)e->layers.size();
  e->layers.push_back(L);
}

extern "C" int idisp_extractor_create(idisp_extractor_t **out)
{
  IDISP_REQUIRE(out != nullptr, "extractor_create: NULL out pointer");
  idisp_extractor *e = new idisp_extractor();
  // submodule.py:63-68
  f2d_add(e, "firstconv.0", 3, 32, 3, 2, 1, true);
  f2d_add(e, "firstconv.2", 32, 32, 3, 1, 1, true);
  f2d_add(e, "firstconv.4", 32, 32, 3, 1, 1, true);
  // :70-73 _make_layer(BasicBlock, planes, blocks, stride, pad, dilation)
  struct Stage { const char *name; int planes, blocks, stride, dil; };
  const Stage stages[4] = {{"layer1", 32, 3, 1, 1}, {"layer2", 64, 16, 2, 1}, {"layer3", 128, 3, 1, 1}, {"layer4", 128, 3, 1, 2}};
  int inpl = 32;
  for (const Stage &s : stages) {
    for (int b = 0; b < s.blocks; ++b) {
      const std::string p = std::string(s.name) + "." + std::to_string(b);
      const int cin = b == 0 ? inpl : s.planes, st = b == 0 ? s.stride : 1;
      f2d_add(e, p + ".conv1.0", cin, s.planes, 3, st, s.dil, true);
      f2d_add(e, p + ".conv2", s.planes, s.planes, 3, 1, s.dil, true);
      if (b == 0 && (s.stride != 1 || inpl != s.planes)) f2d_add(e, p + ".downsample", inpl, s.planes, 1, s.stride, 1, true);
    }
    inpl = s.planes;
  }
  for (const char *br : {"branch1", "branch2", "branch3", "branch4"}) f2d_add(e, std::string(br) + ".1", 128, 32, 1, 1, 1, true);
  f2d_add(e, "lastconv.0", 320, 128, 3, 1, 1, true);
  f2d_add(e, "lastconv.2", 128, 32, 1, 1, 1, false);
  e->no_graph = getenv("IDISP_NO_GRAPH") != nullptr;
  *out = e;
  return IDISP_OK;
}

extern "C" void idisp_extractor_destroy(idisp_extractor_t *e)
{
  if (!e) return;
  for (auto &L : e->layers) c2d_weights_free(L.tc);
  for (auto &g : e->graphs) cudaGraphExecDestroy(g.exec);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->blob) cudaFree(e->blob);
  if (e->range_flag) cudaFree(e->range_flag);
  delete e;
}

extern "C" int idisp_extractor_set_precision(idisp_extractor_t *e, int precision)
{
  IDISP_REQUIRE(e != nullptr, "extractor_set_precision: NULL extractor");
  IDISP_REQUIRE(precision == IDISP_PREC_FP32 || precision == IDISP_PREC_FP16X2, "extractor_set_precision: IDISP_PREC_FP32 or IDISP_PREC_FP16X2 expected, got %d",
                precision);
  if (precision != e->precision) e->finalized = false;   // the tensor-core packing is made at finalize
  e->precision = precision;
  return IDISP_OK;
}

extern "C" int idisp_extractor_range_exceeded(idisp_extractor_t *e, int *exceeded, void *stream)
{
  IDISP_REQUIRE(e != nullptr && exceeded != nullptr, "extractor_range_exceeded: NULL argument");
  *exceeded = 0;
  if (!e->range_flag) return IDISP_OK;
  IDISP_CUDA(cudaMemcpyAsync(exceeded, e->range_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  IDISP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return IDISP_OK;
}

extern "C" int idisp_extractor_set_tensor(idisp_extractor_t *e, const char *key, const float *data, size_t numel)
{
  IDISP_REQUIRE(e && key && (data || numel == 0), "extractor_set_tensor: NULL argument");
  const std::string k = key;
  if (k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return IDISP_OK;
  e->host[k].assign(data, data + numel);
  e->finalized = false;
  return IDISP_OK;
}

static int f2d_need(idisp_extractor *e, const std::string &key, size_t numel, const float **out)
{
  auto it = e->host.find(key);
  if (it == e->host.end()) { set_error("extractor_finalize: missing state_dict entry '%s'", key.c_str()); return IDISP_ERR_STATE; }
  if (it->second.size() != numel) {
    set_error("extractor_finalize: '%s' has %zu elements, expected %zu", key.c_str(), it->second.size(), numel);
    return IDISP_ERR_STATE;
  }
  *out = it->second.data();
  return IDISP_OK;
}

extern "C" int idisp_extractor_finalize(idisp_extractor_t *e, void *stream)
{
  IDISP_REQUIRE(e != nullptr, "extractor_finalize: NULL extractor");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<std::vector<float>> wt(e->layers.size()), bs(e->layers.size());
  size_t total = 0;
  for (size_t i = 0; i < e->layers.size(); ++i) {
    F2dLayer &L = e->layers[i];
    const int kk = L.k * L.k;
    const float *w = nullptr;
    int rc;
    std::vector<double> scale(L.cout, 1.0);
    if (L.bn) {
      const float *g, *b, *m, *v;
      if ((rc = f2d_need(e, L.prefix + ".0.weight", (size_t)L.cout * L.cin * kk, &w))) return rc;
      if ((rc = f2d_need(e, L.prefix + ".1.weight", L.cout, &g))) return rc;
      if ((rc = f2d_need(e, L.prefix + ".1.bias", L.cout, &b))) return rc;
      if ((rc = f2d_need(e, L.prefix + ".1.running_mean", L.cout, &m))) return rc;
      if ((rc = f2d_need(e, L.prefix + ".1.running_var", L.cout, &v))) return rc;
      bs[i].resize(L.cout);
      for (int c = 0; c < L.cout; ++c) {
        scale[c] = (double)g[c] / std::sqrt((double)v[c] + 1e-5);
        bs[i][c] = (float)((double)b[c] - (double)m[c] * scale[c]);
      }
    } else {
      if ((rc = f2d_need(e, L.prefix + ".weight", (size_t)L.cout * L.cin * kk, &w))) return rc;
    }
    wt[i].assign((size_t)L.cin * kk * L.cout, 0.f);   // [Cout][Cin][k][k] -> [Cin][k*k][Cout]
    for (int co = 0; co < L.cout; ++co)
      for (int ci = 0; ci < L.cin; ++ci)
        for (int t = 0; t < kk; ++t)
          wt[i][((size_t)ci * kk + t) * L.cout + co] = (float)((double)w[((size_t)co * L.cin + ci) * kk + t] * scale[co]);
    total += (wt[i].size() + 63) / 64 * 64 + (bs[i].size() + 63) / 64 * 64;
  }
  if (e->blob) { cudaFree(e->blob); e->blob = nullptr; }
  IDISP_CUDA(cudaMalloc(&e->blob, total * sizeof(float)));
  size_t off = 0;
  for (size_t i = 0; i < e->layers.size(); ++i) {
    F2dLayer &L = e->layers[i];
    L.w = e->blob + off;
    IDISP_CUDA(cudaMemcpyAsync(L.w, wt[i].data(), wt[i].size() * sizeof(float), cudaMemcpyHostToDevice, s));
    off += (wt[i].size() + 63) / 64 * 64;
    L.bias = nullptr;
    if (!bs[i].empty()) {
      L.bias = e->blob + off;
      IDISP_CUDA(cudaMemcpyAsync(L.bias, bs[i].data(), bs[i].size() * sizeof(float), cudaMemcpyHostToDevice, s));
      off += (bs[i].size() + 63) / 64 * 64;
    }
  }
  for (auto &g : e->graphs) cudaGraphExecDestroy(g.exec);   // captured launches hold the old weight pointers
  e->graphs.clear();
  for (size_t i = 0; i < e->layers.size(); ++i) {
    F2dLayer &L = e->layers[i];
    c2d_weights_free(L.tc);
    if (e->precision == IDISP_PREC_FP16X2 && L.stride == 1 && L.cin % 32 == 0 && L.cout % 32 == 0 && L.prefix.compare(0, 6, "branch") != 0) {
      const int rc = c2d_weights_prepare(wt[i].data(), L.cin, L.cout, L.k * L.k, L.tc, s);
      if (rc) return rc;
    }
  }
  IDISP_CUDA(cudaStreamSynchronize(s));
  e->finalized = true;
  return IDISP_OK;
}

namespace {
struct F2dDims { int H2, W2, H4, W4; };
inline F2dDims f2d_dims(int H, int W)
{
  F2dDims d;
  d.H2 = (H - 1) / 2 + 1; d.W2 = (W - 1) / 2 + 1;      // k3 s2 p1
  d.H4 = (d.H2 - 1) / 2 + 1; d.W4 = (d.W2 - 1) / 2 + 1;
  return d;
}
inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
}  // namespace

// buffers: two half-resolution 32-channel tensors + a temp (firstconv / layer1), three quarter-resolution 128-channel tensors
// (layer2-4 ping-pong + temp), the 320-channel concat tensor, pooled / branch temporaries, lastconv.0's 128-channel output
extern "C" size_t idisp_extractor_workspace_bytes(const idisp_extractor_t *e, int B, int H, int W)
{
  if (!e || B <= 0 || H <= 0 || W <= 0) return 0;
  const F2dDims d = f2d_dims(H, W);
  const size_t half = up256((size_t)B * 32 * d.H2 * d.W2 * 4), quart = up256((size_t)B * 128 * d.H4 * d.W4 * 4);
  const size_t cat = up256((size_t)B * 320 * d.H4 * d.W4 * 4), pool = up256((size_t)B * 128 * d.H4 * d.W4 / 64 * 4 + 4096);
  // + tensor-core path: three half-resolution and six quarter-resolution blocked split-precision tensors (same bytes as their
  // f32 twins), the blocked concat tensor, one more NCHW staging tensor at each resolution, the fp32 partial of lastconv.0
  return 3 * half + 4 * quart + cat + 2 * pool + (3 * half + 6 * quart + cat) + (half + 2 * quart) + quart;
}

static int f2d_conv(idisp_extractor *e, const std::string &prefix, const float *x, long long xbs, int B, int H, int W, const float *res, long long rbs,
                    int relu, float *y, long long ybs, int *Ho_out, int *Wo_out, cudaStream_t s)
{
  auto it = e->index.find(prefix);
  if (it == e->index.end()) { set_error("extractor: unknown layer '%s'", prefix.c_str()); return IDISP_ERR_INVALID; }
  const F2dLayer &L = e->layers[it->second];
  f2d::ConvParams p;
  p.x = x; p.w = L.w; p.bias = L.bias; p.res = res; p.y = y;
  p.Cin = L.cin; p.Cout = L.cout; p.H = H; p.W = W; p.stride = L.stride; p.dil = L.dil; p.relu = relu;
  p.pad = L.k == 3 ? L.dil : 0;
  p.Ho = (H + 2 * p.pad - L.dil * (L.k - 1) - 1) / L.stride + 1;
  p.Wo = (W + 2 * p.pad - L.dil * (L.k - 1) - 1) / L.stride + 1;
  p.xbs = xbs; p.ybs = ybs > 0 ? ybs : (long long)L.cout * p.Ho * p.Wo; p.rbs = rbs > 0 ? rbs : (long long)L.cout * p.Ho * p.Wo;
  p.tiles_w = ceil_div(p.Wo, f2d::TW); p.tiles_h = ceil_div(p.Ho, f2d::TH);
  if (Ho_out) *Ho_out = p.Ho;
  if (Wo_out) *Wo_out = p.Wo;
  // 64 output channels per CTA when that still fills the machine (two CTAs per SM), else 32: small ROI batches (R = 1..4 on
  // the live path) otherwise leave most SMs idle
  int cob = L.cout >= 64 ? 64 : 32;
  if (cob == 64 && (long long)p.tiles_w * p.tiles_h * ceil_div(L.cout, 64) * B < 2 * 148) cob = 32;
  const int PH = (f2d::TH - 1) * L.stride + (L.k - 1) * L.dil + 1, PW = ((f2d::TW - 1) * L.stride + (L.k - 1) * L.dil + 1) | 1;
  const size_t smem = (size_t)(f2d::CI * PH * PW + f2d::CI * L.k * L.k * cob) * sizeof(float);
  dim3 grid(p.tiles_w * p.tiles_h, ceil_div(L.cout, cob), B);
  auto go = [&](auto kern, bool *opted) -> int {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) { set_error("extractor: device ordinal %d out of range", dev); return IDISP_ERR_INVALID; }
    if (!opted[dev]) { IDISP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); opted[dev] = true; }
    kern<<<grid, 256, smem, s>>>(p);
    IDISP_LAUNCH_CHECK();
    return IDISP_OK;
  };
  static bool o0[64], o1[64], o2[64], o3[64];
  ++e->launches;
  if (L.k == 1 && L.cout % 16 == 0) {
    f2d::pointwise_kernel<<<dim3(ceil_div(p.Ho * p.Wo, 256), L.cout / 16, B), 256, 0, s>>>(x, xbs, L.w, L.bias, res, p.rbs, y, p.ybs, L.cin, L.cout, H, W, p.Ho, p.Wo,
                                                                                            L.stride, relu);
    IDISP_LAUNCH_CHECK();
    return IDISP_OK;
  }
  if (L.k == 3) return cob == 64 ? go(f2d::conv2d_kernel<3, 64>, o0) : go(f2d::conv2d_kernel<3, 32>, o1);
  return cob == 64 ? go(f2d::conv2d_kernel<1, 64>, o2) : go(f2d::conv2d_kernel<1, 32>, o3);
}

static int f2d_pointwise_small(idisp_extractor *e, const std::string &prefix, const float *x, int B, int HW, int relu, float *y, cudaStream_t s)
{
  auto it = e->index.find(prefix);
  if (it == e->index.end()) { set_error("extractor: unknown layer '%s'", prefix.c_str()); return IDISP_ERR_INVALID; }
  const F2dLayer &L = e->layers[it->second];
  f2d::pointwise_small_kernel<<<dim3(ceil_div(L.cout * HW, 8), B), 256, 0, s>>>(x, L.w, L.bias, L.cin, L.cout, HW, relu, y);   // 8 warps = 8 outputs per CTA
  IDISP_LAUNCH_CHECK();
  ++e->launches;
  return IDISP_OK;
}

// images [B,3,H,W] f32 NCHW -> features [B,32,H/4,W/4] f32 NCHW (submodule.py:112-139)
extern "C" int idisp_extractor_forward(idisp_extractor_t *e, const float *images, int B, int H, int W, void *workspace, size_t workspace_bytes,
                                       float *features, void *stream)
{
  IDISP_REQUIRE(e != nullptr, "extractor_forward: NULL extractor");
  if (!e->finalized) { set_error("extractor_forward: not finalised (call idisp_extractor_finalize after loading weights)"); return IDISP_ERR_STATE; }
  IDISP_REQUIRE(B >= 0 && H > 0 && W > 0, "extractor_forward: bad shape B=%d H=%d W=%d", B, H, W);
  if (B == 0) return IDISP_OK;
  const F2dDims d = f2d_dims(H, W);
  IDISP_REQUIRE(d.H4 >= 56 && d.W4 >= 56, "extractor_forward: %dx%d input gives a %dx%d feature map, smaller than branch1's 56x56 average pool "
                "(submodule.py:78)", H, W, d.H4, d.W4);
  IDISP_REQUIRE(images && features && workspace, "extractor_forward: NULL pointer");
  IDISP_REQUIRE(workspace_bytes >= idisp_extractor_workspace_bytes(e, B, H, W), "extractor_forward: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  e->launches = 0;
  char *base = (char *)workspace;
  const size_t half = up256((size_t)B * 32 * d.H2 * d.W2 * 4), quart = up256((size_t)B * 128 * d.H4 * d.W4 * 4);
  const size_t catb = up256((size_t)B * 320 * d.H4 * d.W4 * 4), poolb = up256((size_t)B * 128 * d.H4 * d.W4 / 64 * 4 + 4096);
  float *h0 = (float *)base, *h1 = (float *)(base + half), *h2 = (float *)(base + 2 * half);
  float *q0 = (float *)(base + 3 * half), *q1 = (float *)(base + 3 * half + quart), *q2 = (float *)(base + 3 * half + 2 * quart),
        *q3 = (float *)(base + 3 * half + 3 * quart);
  float *cat = (float *)(base + 3 * half + 4 * quart);
  float *pool = (float *)(base + 3 * half + 4 * quart + catb), *brt = (float *)(base + 3 * half + 4 * quart + catb + poolb);
  int rc, Ho, Wo;
#define FR(expr) do { if ((rc = (expr)) != IDISP_OK) return rc; } while (0)
  const long long hw4 = (long long)d.H4 * d.W4, cat_bs = 320 * hw4;
  if (e->precision == IDISP_PREC_FP16X2) {
    // ---------------- tensor-core path: the 53 stride-1 3x3 convs on tcgen05 in split precision (conv2d_tc.cu) ----------------
    // Blocked split-precision tensors ("x2": [N][hi C/8 | lo C/8][H][W][8] halves) carry the activations between those layers;
    // the few layers that stay on the FFMA kernels (the two stride-2 3x3 convs, the 1x1 convs, pooling, upsampling) read and
    // write NCHW f32 through the converters.  `raw`, `skip` and the four SPP branches live inside the blocked 320-channel concat
    // tensor (torch.cat of submodule.py:134-135 is never executed).
    char *tb = base + 3 * half + 4 * quart + catb + 2 * poolb;
    __nv_bfloat16 *X[3], *Y[6];
    for (int i = 0; i < 3; ++i) X[i] = (__nv_bfloat16 *)(tb + i * half);
    for (int i = 0; i < 6; ++i) Y[i] = (__nv_bfloat16 *)(tb + 3 * half + i * quart);
    __nv_bfloat16 *CATX = (__nv_bfloat16 *)(tb + 3 * half + 6 * quart);
    float *nh1 = (float *)(tb + 3 * half + 6 * quart + catb);                 // NCHW staging, half resolution (32 ch)
    float *nq4 = (float *)(tb + 3 * half + 6 * quart + catb + half);           // NCHW staging, quarter resolution (128 ch) x 2
    float *nq5 = (float *)(tb + 3 * half + 6 * quart + catb + half + quart);
    float *part = (float *)(tb + 3 * half + 6 * quart + catb + half + 2 * quart);
    if (!e->range_flag) IDISP_CUDA(cudaMalloc(&e->range_flag, sizeof(int)));
    IDISP_CUDA(cudaMemsetAsync(e->range_flag, 0, sizeof(int), s));
    auto T = [](__nv_bfloat16 *p, int C) { C2dTensor t; t.p = p; t.blocks = 2 * C / 8; t.blk0 = 0; t.lo = C / 8; return t; };
    auto sub = [](C2dTensor t, int c_off) { t.blk0 += c_off / 8; return t; };
    auto layer = [&](const std::string &prefix) -> const F2dLayer * {
      auto it = e->index.find(prefix);
      return it == e->index.end() ? nullptr : &e->layers[it->second];
    };
    // one tensor-core conv layer; Cin > 128 runs as chunk groups of <= 4 chained through the fp32 partial
    auto tc = [&](const std::string &prefix, const C2dTensor &x, int Hc, int Wc, const C2dTensor *res, int relu, const C2dTensor &y) -> int {
      const F2dLayer *L = layer(prefix);
      if (!L || !L->tc.dev) { set_error("extractor: layer '%s' has no tensor-core packing", prefix.c_str()); return IDISP_ERR_STATE; }
      const int nch = L->cin / 32;
      for (int c0 = 0; c0 < nch; c0 += 4) {
        const int n = nch - c0 < 4 ? nch - c0 : 4;
        const bool first = c0 == 0, last = c0 + n == nch;
        const int r = c2d_conv(L->tc, L->k == 1 ? 0 : L->dil, x, c0, n, B, Hc, Wc, last ? L->bias : nullptr, last ? res : nullptr, last ? relu : 0, y,
                               first ? nullptr : part, last ? nullptr : part, e->range_flag, s);
        if (r) return r;
        ++e->launches;
      }
      return IDISP_OK;
    };
    auto to_x2 = [&](const float *src, int C, long long HWc, const C2dTensor &dst) -> int {
      ++e->launches;
      return c2d_nchw_to_x2(src, (long long)C * HWc, dst.p, dst.blocks, dst.blk0, dst.lo, B, C, HWc, e->range_flag, s);
    };
    auto to_nchw = [&](const C2dTensor &src, int C, long long HWc, float *dst) -> int {
      ++e->launches;
      return c2d_x2_to_nchw(src.p, src.blocks, src.blk0, src.lo, dst, (long long)C * HWc, B, C, HWc, s);
    };
    // firstconv (:63-68): the stride-2 3 -> 32 conv on the FFMA kernel, then two tensor-core convs
    FR(f2d_conv(e, "firstconv.0", images, (long long)3 * H * W, B, H, W, nullptr, 0, 1, h0, 0, &Ho, &Wo, s));
    const int H2 = Ho, W2 = Wo;
    const long long hw2 = (long long)H2 * W2;
    Ho = (H2 - 1) / 2 + 1; Wo = (W2 - 1) / 2 + 1;   // quarter resolution (layer2.0's stride-2 conv)
    auto middle = [&]() -> int {
    FR(to_x2(h0, 32, hw2, T(X[0], 32)));
    FR(tc("firstconv.2", T(X[0], 32), H2, W2, nullptr, 1, T(X[1], 32)));
    FR(tc("firstconv.4", T(X[1], 32), H2, W2, nullptr, 1, T(X[0], 32)));
    int cur = 0, tmp = 1, nxt = 2;
    for (int b = 0; b < 3; ++b) {   // layer1 (:25-48: conv1 + ReLU, conv2, += x)
      const std::string p = "layer1." + std::to_string(b);
      const C2dTensor xc = T(X[cur], 32), xt = T(X[tmp], 32), xn = T(X[nxt], 32);
      FR(tc(p + ".conv1.0", xc, H2, W2, nullptr, 1, xt));
      FR(tc(p + ".conv2", xt, H2, W2, &xc, 0, xn));
      const int t = cur; cur = nxt; nxt = t;
    }
    // layer2.0: stride-2 conv1 and the 1x1 stride-2 downsample on the FFMA kernels (NCHW), conv2 on the tensor cores
    FR(to_nchw(T(X[cur], 32), 32, hw2, nh1));
    FR(f2d_conv(e, "layer2.0.conv1.0", nh1, 32 * hw2, B, H2, W2, nullptr, 0, 1, nq4, 0, nullptr, nullptr, s));
    FR(f2d_conv(e, "layer2.0.downsample", nh1, 32 * hw2, B, H2, W2, nullptr, 0, 0, nq5, 0, nullptr, nullptr, s));
    FR(to_x2(nq4, 64, hw4, T(Y[0], 64)));
    FR(to_x2(nq5, 64, hw4, T(Y[1], 64)));
    {
      const C2dTensor a = T(Y[0], 64), dsm = T(Y[1], 64);
      FR(tc("layer2.0.conv2", a, Ho, Wo, &dsm, 0, T(Y[2], 64)));
    }
    const C2dTensor CAT = T(CATX, 320);
    int qc = 2, qt = 0, qn = 1;
    for (int b = 1; b < 16; ++b) {   // layer2.1 .. 15; the last block writes `raw` = channels [0, 64) of the concat tensor
      const std::string p = "layer2." + std::to_string(b);
      const C2dTensor xc = T(Y[qc], 64), xt = T(Y[qt], 64);
      const C2dTensor xn = b == 15 ? sub(CAT, 0) : T(Y[qn], 64);
      FR(tc(p + ".conv1.0", xc, Ho, Wo, nullptr, 1, xt));
      FR(tc(p + ".conv2", xt, Ho, Wo, &xc, 0, xn));
      const int t = qc; qc = qn; qn = t;
    }
    // layer3.0: conv1 64 -> 128 and the 1x1 downsample on the tensor cores (both read `raw` in place)
    const C2dTensor raw = sub(CAT, 0);
    FR(tc("layer3.0.conv1.0", raw, Ho, Wo, nullptr, 1, T(Y[0], 128)));
    FR(tc("layer3.0.downsample", raw, Ho, Wo, nullptr, 0, T(Y[1], 128)));   // 1x1 (stride 1) 64 -> 128 + BN: one-tap tensor-core conv
    {
      const C2dTensor a = T(Y[0], 128), dsm = T(Y[1], 128);
      FR(tc("layer3.0.conv2", a, Ho, Wo, &dsm, 0, T(Y[2], 128)));
    }
    int sc = 2, st = 0, sn = 1;
    for (int li = 3; li <= 4; ++li)
      for (int b = (li == 3 ? 1 : 0); b < 3; ++b) {   // layer3.1, .2, layer4.0 .. 2 (dilation 2); the last one writes `skip`
        const std::string p = "layer" + std::to_string(li) + "." + std::to_string(b);
        const bool lastb = li == 4 && b == 2;
        const C2dTensor xc = T(Y[sc], 128), xt = T(Y[st], 128);
        const C2dTensor xn = lastb ? sub(CAT, 64) : T(Y[sn], 128);
        FR(tc(p + ".conv1.0", xc, Ho, Wo, nullptr, 1, xt));
        FR(tc(p + ".conv2", xt, Ho, Wo, &xc, 0, xn));
        const int t = sc; sc = sn; sn = t;
      }
    // SPP branches (:78-92, :115-132) on the FFMA kernels from an NCHW copy of `skip`; their upsampled outputs are collected in
    // concat order (branch4, branch3, branch2, branch1) and converted into channels [192, 320) of the concat tensor in one go
    FR(to_nchw(sub(CAT, 64), 128, hw4, nq4));
    {
      const int ks[4] = {56, 32, 16, 8};
      const int c_off[4] = {96, 64, 32, 0};  // branch1 .. branch4 inside the 128-channel staging tensor
      for (int bi = 0; bi < 4; ++bi) {
        const int k = ks[bi], Hp = (Ho - k) / k + 1, Wp = (Wo - k) / k + 1;
        f2d::avgpool_kernel<<<dim3(ceil_div(128 * Hp * Wp, 8), B), 256, 0, s>>>(nq4, 128 * hw4, 128, Ho, Wo, k, Hp, Wp, pool);
        IDISP_LAUNCH_CHECK();
        FR(f2d_pointwise_small(e, "branch" + std::to_string(bi + 1) + ".1", pool, B, Hp * Wp, 1, brt, s));
        f2d::upsample_bilinear_kernel<<<dim3(ceil_div(32 * Ho * Wo, 256), B), 256, 0, s>>>(brt, 32, Hp, Wp, Ho, Wo, nq5, 128 * hw4, c_off[bi]);
        IDISP_LAUNCH_CHECK();
        e->launches += 2;
      }
    }
    FR(to_x2(nq5, 128, hw4, sub(CAT, 192)));
    // lastconv (:94-96): 320 -> 128 (three chunk groups) and the final 1x1 on the tensor cores, then NCHW f32 for the 3-D stack
    FR(tc("lastconv.0", CAT, Ho, Wo, nullptr, 1, T(Y[0], 128)));
    FR(tc("lastconv.2", T(Y[0], 128), Ho, Wo, nullptr, 0, T(Y[1], 32)));    // 1x1 128 -> 32, no BN, no bias
    return IDISP_OK;
    };  // middle
    {
      cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
      const bool graph_ok = !e->no_graph && cudaStreamIsCapturing(s, &cst) == cudaSuccess && cst == cudaStreamCaptureStatusNone;
      idisp_extractor::GraphEntry *hit = nullptr;
      if (graph_ok)
        for (auto &g : e->graphs)
          if (g.B == B && g.H == H && g.W == W && g.ws == workspace) { hit = &g; break; }
      if (graph_ok && !hit) {
        const int before = e->launches;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool ok = e->cap_stream || cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          cudaStream_t user = s;
          s = e->cap_stream;
          const int crc = middle();
          s = user;
          const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
          if (crc != IDISP_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return crc; }
          ok = ce == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
          if (graph) cudaGraphDestroy(graph);
        }
        if (!ok) { cudaGetLastError(); e->no_graph = true; e->launches = before; }
        else {
          if (e->graphs.size() >= 32) {
            size_t lru = 0;
            for (size_t i = 1; i < e->graphs.size(); ++i) if (e->graphs[i].stamp < e->graphs[lru].stamp) lru = i;
            cudaGraphExecDestroy(e->graphs[lru].exec);
            e->graphs.erase(e->graphs.begin() + lru);
          }
          e->graphs.push_back({B, H, W, workspace, exec, e->launches - before, 0ull});
          hit = &e->graphs.back();
          e->launches = before;
        }
      }
      if (hit) {
        hit->stamp = ++e->graph_clock;
        IDISP_CUDA(cudaGraphLaunch(hit->exec, s));
        e->launches += hit->launches;
      } else {
        FR(middle());
      }
    }
    FR(to_nchw(T(Y[1], 32), 32, hw4, features));
    return IDISP_OK;
  }
  // firstconv (:63-68)
  FR(f2d_conv(e, "firstconv.0", images, (long long)3 * H * W, B, H, W, nullptr, 0, 1, h0, 0, &Ho, &Wo, s));
  FR(f2d_conv(e, "firstconv.2", h0, (long long)32 * Ho * Wo, B, Ho, Wo, nullptr, 0, 1, h1, 0, nullptr, nullptr, s));
  FR(f2d_conv(e, "firstconv.4", h1, (long long)32 * Ho * Wo, B, Ho, Wo, nullptr, 0, 1, h0, 0, nullptr, nullptr, s));
  // layer1: three BasicBlocks at half resolution (:25-48: conv1+ReLU, conv2, += x, no ReLU after the add)
  float *cur = h0, *tmp = h1, *nxt = h2;
  const long long hbs = (long long)32 * Ho * Wo;
  for (int b = 0; b < 3; ++b) {
    const std::string p = "layer1." + std::to_string(b);
    FR(f2d_conv(e, p + ".conv1.0", cur, hbs, B, Ho, Wo, nullptr, 0, 1, tmp, 0, nullptr, nullptr, s));
    FR(f2d_conv(e, p + ".conv2", tmp, hbs, B, Ho, Wo, cur, hbs, 0, nxt, 0, nullptr, nullptr, s));
    float *t = cur; cur = nxt; nxt = t;
  }
  // layer2: 16 blocks, first one stride 2 with a 1x1 stride-2 downsample of the input; the LAST block writes `raw` into
  // channels [0,64) of the concat tensor
  const int H2 = Ho, W2 = Wo;
  float *qc = q0, *qt = q1, *qn = q2;
  long long qc_bs = 64 * hw4;
  for (int b = 0; b < 16; ++b) {
    const std::string p = "layer2." + std::to_string(b);
    if (b == 0) {
      FR(f2d_conv(e, p + ".conv1.0", cur, hbs, B, H2, W2, nullptr, 0, 1, qt, 0, &Ho, &Wo, s));
      FR(f2d_conv(e, p + ".downsample", cur, hbs, B, H2, W2, nullptr, 0, 0, q3, 0, nullptr, nullptr, s));
      FR(f2d_conv(e, p + ".conv2", qt, 64 * hw4, B, Ho, Wo, q3, 64 * hw4, 0, qc, 0, nullptr, nullptr, s));
    } else {
      float *dst = b == 15 ? cat : qn;
      const long long dbs = b == 15 ? cat_bs : 64 * hw4;
      FR(f2d_conv(e, p + ".conv1.0", qc, qc_bs, B, Ho, Wo, nullptr, 0, 1, qt, 0, nullptr, nullptr, s));
      FR(f2d_conv(e, p + ".conv2", qt, 64 * hw4, B, Ho, Wo, qc, qc_bs, 0, dst, dbs, nullptr, nullptr, s));
      if (b < 15) { float *t = qc; qc = qn; qn = t; }
      else { qc = cat; qc_bs = cat_bs; }
    }
  }
  // layer3 (64 -> 128, 1x1 downsample of `raw` in block 0), layer4 (dilation 2); the last block writes `skip` into channels [64,192)
  float *raw = cat;
  float *sc = q0, *st = q1, *sn = q2;
  long long sc_bs = 128 * hw4;
  for (int li = 3; li <= 4; ++li)
    for (int b = 0; b < 3; ++b) {
      const std::string p = "layer" + std::to_string(li) + "." + std::to_string(b);
      const bool last = li == 4 && b == 2;
      float *dst = last ? cat + 64 * hw4 : sn;
      const long long dbs = last ? cat_bs : 128 * hw4;
      if (li == 3 && b == 0) {
        FR(f2d_conv(e, p + ".conv1.0", raw, cat_bs, B, Ho, Wo, nullptr, 0, 1, st, 0, nullptr, nullptr, s));
        FR(f2d_conv(e, p + ".downsample", raw, cat_bs, B, Ho, Wo, nullptr, 0, 0, q3, 0, nullptr, nullptr, s));
        FR(f2d_conv(e, p + ".conv2", st, 128 * hw4, B, Ho, Wo, q3, 128 * hw4, 0, sc, 0, nullptr, nullptr, s));
        continue;
      }
      FR(f2d_conv(e, p + ".conv1.0", sc, sc_bs, B, Ho, Wo, nullptr, 0, 1, st, 0, nullptr, nullptr, s));
      FR(f2d_conv(e, p + ".conv2", st, 128 * hw4, B, Ho, Wo, sc, sc_bs, 0, dst, dbs, nullptr, nullptr, s));
      if (!last) { float *t = sc; sc = sn; sn = t; }
    }
  const float *skip = cat + 64 * hw4;
  // SPP branches (:78-92, :115-132); concat order raw, skip, branch4, branch3, branch2, branch1 (:134-135)
  const int ks[4] = {56, 32, 16, 8};
  const int c_off[4] = {288, 256, 224, 192};  // branch1 .. branch4
  for (int bi = 0; bi < 4; ++bi) {
    const int k = ks[bi], Hp = (Ho - k) / k + 1, Wp = (Wo - k) / k + 1;
    const int n_el = 128 * Hp * Wp;
    f2d::avgpool_kernel<<<dim3(ceil_div(n_el, 8), B), 256, 0, s>>>(skip, cat_bs, 128, Ho, Wo, k, Hp, Wp, pool);
    IDISP_LAUNCH_CHECK();
    FR(f2d_pointwise_small(e, "branch" + std::to_string(bi + 1) + ".1", pool, B, Hp * Wp, 1, brt, s));
    f2d::upsample_bilinear_kernel<<<dim3(ceil_div(32 * Ho * Wo, 256), B), 256, 0, s>>>(brt, 32, Hp, Wp, Ho, Wo, cat, cat_bs, c_off[bi]);
    IDISP_LAUNCH_CHECK();
    e->launches += 2;
  }
  // lastconv (:94-96)
  FR(f2d_conv(e, "lastconv.0", cat, cat_bs, B, Ho, Wo, nullptr, 0, 1, q0, 0, nullptr, nullptr, s));
  FR(f2d_conv(e, "lastconv.2", q0, 128 * hw4, B, Ho, Wo, nullptr, 0, 0, features, 0, nullptr, nullptr, s));
#undef FR
  return IDISP_OK;
}

extern "C" int idisp_extractor_launches_per_forward(const idisp_extractor_t *e) { return e ? e->launches : 0; }
