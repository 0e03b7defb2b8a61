// plan.cu -- host side of libidisp: error state, weight folding, the 28-layer forward schedule.
//
// Schedule and wiring follow disprcnn/modeling/psmnet/stackhourglass.py:130-144 (and the
// hourglass at :32-51); layer inventory = SURVEY.md Appendix A.  Weights arrive under the
// reference's own state_dict keys (stackhourglass.py:63-88), BatchNorm3d (eval, eps 1e-5,
// submodule.py:22) is folded into a per-output-channel scale (multiplied into the kernel) and
// bias.
#include <map>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "conv3d_tc.cuh"

namespace idisp {

static thread_local std::string g_err;

void set_error(const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return IDISP_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------
struct LayerSpec {
  std::string prefix;
  int kind, cin, cout;
  bool bn;   // convbn_3d / deconv+BN (prefix.0.weight + prefix.1.*) vs bare Conv3d (prefix.weight)
};

struct LayerDev {
  float *w_tap = nullptr;  // [27][cin][cout] f32, BN scale folded in
  float *bias = nullptr;   // [cout] f32 (nullptr for the bare 32->1 convs)
  TcWeights tc;            // 16-bit re-lay for the tcgen05 kernel (tensor-core precisions)
  TcSplitWeights sp;       // split precision: hi / lo / two-word packings
  TcHeadWeights head;      // 32 -> 1 layers: the taps-in-N packing of head_tc.cu
};

static std::vector<LayerSpec> make_layers(int C)
{
  std::vector<LayerSpec> L;
  L.push_back({"dres0.0", IDISP_CONV_S1, 2 * C, 32, true});
  L.push_back({"dres0.2", IDISP_CONV_S1, 32, 32, true});
  L.push_back({"dres1.0", IDISP_CONV_S1, 32, 32, true});
  L.push_back({"dres1.2", IDISP_CONV_S1, 32, 32, true});
  for (const char *h : {"dres2", "dres3", "dres4"}) {
    const std::string p = h;
    L.push_back({p + ".conv1.0", IDISP_CONV_S2, 32, 64, true});
    L.push_back({p + ".conv2", IDISP_CONV_S1, 64, 64, true});
    L.push_back({p + ".conv3.0", IDISP_CONV_S2, 64, 64, true});
    L.push_back({p + ".conv4.0", IDISP_CONV_S1, 64, 64, true});
    L.push_back({p + ".conv5", IDISP_DECONV_S2, 64, 64, true});
    L.push_back({p + ".conv6", IDISP_DECONV_S2, 64, 32, true});
  }
  for (const char *c : {"classif1", "classif2", "classif3"}) L.push_back({std::string(c) + ".0", IDISP_CONV_S1, 32, 32, true});
  for (const char *c : {"classif1", "classif2", "classif3"}) L.push_back({std::string(c) + ".2", IDISP_CONV_S1, 32, 1, false});
  return L;
}

// re-lay a PyTorch conv kernel into tap-major [27][cin][cout] with an optional per-cout scale
static void relayout_taps(const float *w, int kind, int cin, int cout, const double *scale, std::vector<float> &out)
{
  out.assign((size_t)27 * cin * cout, 0.f);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < 27; ++t) {
        const size_t src = kind == IDISP_DECONV_S2 ? ((size_t)ci * cout + co) * 27 + t   // [Cin][Cout][27]
                                                   : ((size_t)co * cin + ci) * 27 + t;  // [Cout][Cin][27]
        const double v = (double)w[src] * (scale ? scale[co] : 1.0);
        out[((size_t)t * cin + ci) * cout + co] = (float)v;
      }
}

}  // namespace idisp

using namespace idisp;

struct idisp_plan {
  int C, mindisp, maxdisp, precision, D;
  int f16 = 0;  // 16-bit storage format of the tensor-core path: 0 bf16, 1 IEEE half
  int x2 = 0;   // split precision: activations/weights are hi+lo pairs of IEEE-half words, three MMA passes per layer
  std::vector<LayerSpec> layers;
  std::map<std::string, std::vector<float>> host;  // reference-keyed tensors
  std::vector<LayerDev> dev;
  float *blob = nullptr;  // one allocation behind all LayerDev f32 pointers
  bool finalized = false;
  // state of the last forward (for get_logits)
  const float *last_logits = nullptr;
  int last_B = 0, last_Hf = 0, last_Wf = 0;
  int launches = 0;
  // host-buffer entry point staging
  void *stage = nullptr;
  size_t stage_bytes = 0;
  // pipelined host-buffer entry point (idisp_plan_forward_host_async): copy streams, per-slot events, double-buffered staging
  void *pstage = nullptr;
  size_t pstage_bytes = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  unsigned long long n_async = 0, async_base = 0;   // calls issued; first call that used the current staging allocation
  // optional per-launch CUDA-event timing (bench.py's roofline leg)
  int *range_flag = nullptr;  // device int: an fp16-mode forward saw a value outside the IEEE-half range
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_layer;  // launch slot -> layer index (-1 cost volume, -2 soft-argmin, 25..27 the 32->1 convs)
  // switches read ONCE at plan creation (tests flip them between plans): IDISP_NO_FUSED_SPLIT / IDISP_NO_FUSED_CV /
  // IDISP_X2_SIMT_HEADS / IDISP_NO_GRAPH
  bool no_fused_split = false, no_fused_cv = false, x2_simt_heads = false, no_graph = false, no_side_copy = false;
  __nv_bfloat16 *x_copy_next = nullptr;   // set right before a tc_layer call: that launch also writes its input in the parity layout (fp16x2)
  // CUDA-graph replay of the conv section (everything between the input conversion and the soft-argmin touches only the
  // workspace, so its ~45 launches -- each with a host-side tensor-map encode -- are captured once per
  // (B, Hf, Wf, workspace) and replayed with ONE cudaGraphLaunch)
  struct GraphEntry { int B, Hf, Wf; void *ws; cudaGraphExec_t exec; int launches; unsigned long long stamp; };
  std::vector<GraphEntry> graphs;
  unsigned long long graph_clock = 0;
  int graph_hits = 0, graph_captures = 0;
  cudaStream_t cap_stream = nullptr;  // capture happens here: the caller's stream may be the legacy default stream, which cannot capture
};

extern "C" int idisp_version(void) { return IDISP_VERSION; }
extern "C" const char *idisp_last_error(void) { return g_err.c_str(); }

extern "C" int idisp_plan_create(int C, int mindisp, int maxdisp, int precision, idisp_plan_t **plan)
{
  IDISP_REQUIRE(plan != nullptr, "plan_create: NULL out pointer");
  *plan = nullptr;
  IDISP_REQUIRE(C > 0 && (2 * C) % 8 == 0, "plan_create: C=%d must make 2C a multiple of 8", C);
  IDISP_REQUIRE(maxdisp > mindisp && mindisp % 4 == 0 && maxdisp % 4 == 0,
                "plan_create: mindisp=%d maxdisp=%d must be multiples of 4, maxdisp>mindisp", mindisp, maxdisp);
  IDISP_REQUIRE(((maxdisp - mindisp) / 4) % 4 == 0,
                "plan_create: D=(maxdisp-mindisp)/4=%d must be a multiple of 4 (two stride-2 stages)", (maxdisp - mindisp) / 4);
  IDISP_REQUIRE(precision == IDISP_PREC_FP32 || precision == IDISP_PREC_BF16 || precision == IDISP_PREC_FP16 || precision == IDISP_PREC_FP16X2, "plan_create: unknown precision %d", precision);
  idisp_plan *p = new idisp_plan();
  p->C = C; p->mindisp = mindisp; p->maxdisp = maxdisp; p->precision = precision; p->f16 = precision == IDISP_PREC_FP16 || precision == IDISP_PREC_FP16X2; p->x2 = precision == IDISP_PREC_FP16X2;
  p->D = (maxdisp - mindisp) / 4;
  p->layers = make_layers(C);
  p->no_fused_split = getenv("IDISP_NO_FUSED_SPLIT") != nullptr;
  p->no_fused_cv = getenv("IDISP_NO_FUSED_CV") != nullptr;
  p->x2_simt_heads = getenv("IDISP_X2_SIMT_HEADS") != nullptr;
  p->no_graph = getenv("IDISP_NO_GRAPH") != nullptr;
  p->no_side_copy = getenv("IDISP_NO_SIDE_COPY") != nullptr;
  *plan = p;
  return IDISP_OK;
}

extern "C" void idisp_plan_destroy(idisp_plan_t *p)
{
  if (!p) return;
  for (auto &d : p->dev) { tc_weights_free(d.tc); tc_split_weights_free(d.sp); tc_head_weights_free(d.head); }
  for (auto e : p->ev) cudaEventDestroy(e);
  for (auto &g : p->graphs) cudaGraphExecDestroy(g.exec);
  if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
  if (p->blob) cudaFree(p->blob);
  if (p->range_flag) cudaFree(p->range_flag);
  if (p->stage) cudaFree(p->stage);
  if (p->pstage) cudaFree(p->pstage);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  for (int i = 0; i < 2; ++i) {
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]);
    if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
  }
  delete p;
}

extern "C" int idisp_plan_set_tensor(idisp_plan_t *p, const char *key, const float *data, size_t numel)
{
  IDISP_REQUIRE(p && key && (data || numel == 0), "plan_set_tensor: NULL argument");
  const std::string k = key;
  bool wanted = false;
  for (const auto &L : p->layers) {
    if (k.compare(0, L.prefix.size(), L.prefix) == 0 && k.size() > L.prefix.size() && k[L.prefix.size()] == '.') {
      wanted = true;
      break;
    }
  }
  if (!wanted || (k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0)) return IDISP_OK;
  p->host[k].assign(data, data + numel);
  p->finalized = false;
  return IDISP_OK;
}

static int need(idisp_plan *p, const std::string &key, size_t numel, const float **out)
{
  auto it = p->host.find(key);
  if (it == p->host.end()) {
    set_error("plan_finalize: missing state_dict entry '%s'", key.c_str());
    return IDISP_ERR_STATE;
  }
  if (it->second.size() != numel) {
    set_error("plan_finalize: '%s' has %zu elements, expected %zu", key.c_str(), it->second.size(), numel);
    return IDISP_ERR_STATE;
  }
  *out = it->second.data();
  return IDISP_OK;
}

extern "C" int idisp_plan_finalize(idisp_plan_t *p, void *stream)
{
  IDISP_REQUIRE(p != nullptr, "plan_finalize: NULL plan");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nl = p->layers.size();
  std::vector<std::vector<float>> wt(nl), bs(nl);
  size_t total = 0;
  for (size_t i = 0; i < nl; ++i) {
    const LayerSpec &L = p->layers[i];
    const float *w = nullptr;
    int rc;
    std::vector<double> scale(L.cout, 1.0);
    if (L.bn) {
      const float *g, *b, *m, *v;
      if ((rc = need(p, L.prefix + ".0.weight", (size_t)27 * L.cin * L.cout, &w))) return rc;
      if ((rc = need(p, L.prefix + ".1.weight", L.cout, &g))) return rc;
      if ((rc = need(p, L.prefix + ".1.bias", L.cout, &b))) return rc;
      if ((rc = need(p, L.prefix + ".1.running_mean", L.cout, &m))) return rc;
      if ((rc = need(p, L.prefix + ".1.running_var", L.cout, &v))) return rc;
      bs[i].resize(L.cout);
      for (int c = 0; c < L.cout; ++c) {
        scale[c] = (double)g[c] / std::sqrt((double)v[c] + 1e-5);
        bs[i][c] = (float)((double)b[c] - (double)m[c] * scale[c]);
      }
    } else {
      if ((rc = need(p, L.prefix + ".weight", (size_t)27 * L.cin * L.cout, &w))) return rc;
    }
    relayout_taps(w, L.kind, L.cin, L.cout, L.bn ? scale.data() : nullptr, wt[i]);
    total += (wt[i].size() + 63) / 64 * 64 + (bs[i].size() + 63) / 64 * 64;
  }
  for (auto &d : p->dev) { tc_weights_free(d.tc); tc_split_weights_free(d.sp); tc_head_weights_free(d.head); }
  for (auto &g : p->graphs) cudaGraphExecDestroy(g.exec);  // captured launches hold the old weight pointers
  p->graphs.clear();
  if (p->blob) { cudaFree(p->blob); p->blob = nullptr; }
  IDISP_CUDA(cudaMalloc(&p->blob, total * sizeof(float)));
  p->dev.assign(nl, LayerDev());
  size_t off = 0;
  for (size_t i = 0; i < nl; ++i) {
    p->dev[i].w_tap = p->blob + off;
    IDISP_CUDA(cudaMemcpyAsync(p->dev[i].w_tap, wt[i].data(), wt[i].size() * sizeof(float), cudaMemcpyHostToDevice, s));
    off += (wt[i].size() + 63) / 64 * 64;
    if (!bs[i].empty()) {
      p->dev[i].bias = p->blob + off;
      IDISP_CUDA(cudaMemcpyAsync(p->dev[i].bias, bs[i].data(), bs[i].size() * sizeof(float), cudaMemcpyHostToDevice, s));
      off += (bs[i].size() + 63) / 64 * 64;
    }
    if (p->precision != IDISP_PREC_FP32) {
      const LayerSpec &L = p->layers[i];
      int rc = p->x2 ? tc_split_weights_prepare(wt[i].data(), L.kind, L.cin, L.cout, p->dev[i].sp, s)
                     : tc_weights_prepare(wt[i].data(), L.kind, L.cin, L.cout, p->f16, p->dev[i].tc, s);
      if (rc) return rc;
      if (L.cout == 1 && (rc = tc_head_weights_prepare(wt[i].data(), L.cin, p->f16, p->x2, p->dev[i].head, s))) return rc;
    }
  }
  IDISP_CUDA(cudaStreamSynchronize(s));  // host vectors go out of scope
  p->finalized = true;
  return IDISP_OK;
}

// ---------------------------------------------------------------------------------------
// workspace arena (same code computes the size and hands out the pointers)
// ---------------------------------------------------------------------------------------
namespace {
struct Arena {
  char *base;
  size_t off = 0;
  explicit Arena(void *b) : base((char *)b) {}
  void *take(size_t bytes)
  {
    void *p = base ? base + off : nullptr;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  }
};

struct Buffers {
  void *cv, *a, *t0, *cost0, *out, *c;       // full resolution
  void *h1, *pre1, *prek, *postA, *postB;    // half resolution, 64 ch
  void *q1, *q2;                             // quarter resolution, 64 ch
  float *costX, *costY;                      // 1-channel f32 logits
  void *split;                               // parity re-lay scratch of the stride-2 tensor-core convs
  void *feaL, *feaR;                         // fused cost volume: blocked 16-bit per-view features
  float *part;                               // split precision: fp32 partial sums between the three passes of a layer
};

Buffers carve(Arena &A, int C, int B, int D, int Hf, int Wf, size_t esz)
{
  const size_t V = (size_t)D * Hf * Wf;
  const size_t full32 = (size_t)B * 32 * V * esz, half64 = (size_t)B * 64 * (V / 8) * esz, quart64 = (size_t)B * 64 * (V / 64) * esz;
  Buffers b;
  b.cv = A.take((size_t)B * 2 * C * V * esz);
  b.a = A.take(full32); b.t0 = A.take(full32); b.cost0 = A.take(full32); b.out = A.take(full32); b.c = A.take(full32);
  b.h1 = A.take(half64); b.pre1 = A.take(half64); b.prek = A.take(half64); b.postA = A.take(half64); b.postB = A.take(half64);
  b.q1 = A.take(quart64); b.q2 = A.take(quart64);
  b.costX = (float *)A.take((size_t)B * V * 4); b.costY = (float *)A.take((size_t)B * V * 4);
  b.split = A.take(full32);
  b.feaL = A.take((size_t)B * C * Hf * Wf * esz); b.feaR = A.take((size_t)B * C * Hf * Wf * esz);  // esz 2 (one word) or 4 (hi|lo)
  b.part = (float *)A.take((size_t)B * 32 * V * 4);
  return b;
}
}  // namespace

extern "C" size_t idisp_plan_workspace_bytes(const idisp_plan_t *p, int B, int Hf, int Wf)
{
  if (!p || B <= 0 || Hf <= 0 || Wf <= 0) return 0;
  Arena A(nullptr);
  carve(A, p->C, B, p->D, Hf, Wf, (p->precision == IDISP_PREC_FP32 || p->x2) ? 4 : 2);
  return A.off;
}

// One layer on the tensor-core path: a single launch, or -- split precision -- the 1-3 launches of tc_conv3d_split
// (chained through the fp32 partial buffer; through y1 itself for the 1-channel heads).
//   xflags bit0: the input pointer holds the parity sub-volume layout;  eflags: 2 = residual in parity layout, 4 = skip y
static int tc_layer(idisp_plan *p, int li, const __nv_bfloat16 *xin, int xflags, const TcCostVolume *cv, int B, int d, int h, int w,
                    const __nv_bfloat16 *res, int relu, __nv_bfloat16 *y, __nv_bfloat16 *ysp, int eflags, const float *res1, float *y1,
                    void *scratch, float *part, cudaStream_t s, int &launches)
{
  const LayerSpec &L = p->layers[li];
  const float *bias = p->dev[li].bias;
  if (!p->x2) {
    TcOpts o;
    o.range_flag = p->f16 ? p->range_flag : nullptr;
    return tc_conv3d(p->dev[li].tc, xin, B, L.cin, d, h, w, L.cout, L.kind, bias, res, relu, y, res1, y1, scratch, xflags | eflags, ysp, cv, s, &o);
  }
  int nl = 1;
  __nv_bfloat16 *xcopy = p->x_copy_next;
  p->x_copy_next = nullptr;
  const int rc = tc_conv3d_split(p->dev[li].sp, xin, B, L.cin, d, h, w, L.cout, L.kind, bias, res, relu, y, res1, y1, scratch, xflags | eflags,
                                 ysp, cv, part, s, &nl, p->range_flag, xcopy);
  launches += nl - 1;
  return rc;
}

template <typename T>
static int forward_impl(idisp_plan *p, const float *left, const float *right, int B, int Hf, int Wf, int H, int W,
                        void *workspace, float *out, cudaStream_t s)
{
  const int D = p->D, C = p->C;
  Arena A(workspace);
  Buffers b = carve(A, C, B, D, Hf, Wf, p->x2 ? 4 : sizeof(T));
  int launches = 0;
  int rc;
  if (p->f16) {  // fp16-word modes: a device flag collects "value left the IEEE-half range" over this forward
    if (!p->range_flag) IDISP_CUDA(cudaMalloc(&p->range_flag, sizeof(int)));
    IDISP_CUDA(cudaMemsetAsync(p->range_flag, 0, sizeof(int), s));
  }
  if (p->timing) p->ev_layer.clear();
  int nmark = 0;
  auto mark = [&](int layer) {  // record an event BEFORE the launch(es) of `layer`
    if (!p->timing) return;
    while ((int)p->ev.size() <= nmark) { cudaEvent_t e; cudaEventCreate(&e); p->ev.push_back(e); }
    cudaEventRecord(p->ev[nmark++], s);
    if (layer != -99) p->ev_layer.push_back(layer);
  };
  // split_out: also emit the output as 8 parity sub-volumes into b.split (feeds the next stride-2 tensor-core conv);
  // x_split: the input IS b.split (written that way by its producer), so no re-lay pass is needed.
  auto conv = [&](int li, const void *x, int d, int h, int w, const void *res, int relu, void *y, bool split_out = false,
                  bool x_split = false, int extra_flags = 0, void *split_dst = nullptr) -> int {
    const LayerSpec &L = p->layers[li];
    mark(li);
    ++launches;
    if (std::is_same<T, __nv_bfloat16>::value && tc_supported(L.kind, L.cin, L.cout, d, h, w)) {
      if (L.kind == IDISP_CONV_S2 && !x_split) ++launches;  // + the space-to-depth re-lay
      const __nv_bfloat16 *xin = (const __nv_bfloat16 *)(x_split ? b.split : x);
      __nv_bfloat16 *ysp = split_out ? (__nv_bfloat16 *)(split_dst ? split_dst : b.split) : nullptr;
      return tc_layer(p, li, xin, x_split ? 1 : 0, nullptr, B, d, h, w, (const __nv_bfloat16 *)res, relu, (__nv_bfloat16 *)y, ysp, extra_flags,
                      nullptr, nullptr, b.split, b.part, s, launches);
    }
    if (p->f16) { set_error("plan_forward: fp16 mode needs tensor-core-supported layer shapes (layer %s)", L.prefix.c_str()); return IDISP_ERR_UNSUPPORTED; }
    return launch_conv3d_simt<T>((const T *)x, B, L.cin, d, h, w, p->dev[li].w_tap, L.cout, L.kind, p->dev[li].bias,
                                 (const T *)res, relu, (T *)y, s);
  };
  // the fused parity-split hand-off needs every layer of the chain on the tensor-core path
  const bool fuse_split = std::is_same<T, __nv_bfloat16>::value && D % 4 == 0 && Hf % 4 == 0 && Wf % 4 == 0 &&
                          tc_supported(IDISP_CONV_S2, 32, 64, D, Hf, Wf) && tc_supported(IDISP_DECONV_S2, 64, 32, D / 2, Hf / 2, Wf / 2) &&
                          tc_supported(IDISP_CONV_S1, 32, 32, D, Hf, Wf) && !p->no_fused_split;
#define RUN(expr) do { if ((rc = (expr)) != IDISP_OK) return rc; } while (0)
  // cost volume (stackhourglass.py:115-128)
  const bool fuse_cv = std::is_same<T, __nv_bfloat16>::value && tc_supported(IDISP_CONV_S1, 2 * C, 32, D, Hf, Wf) && C % 8 == 0 && D <= 64 &&
                       !p->no_fused_cv;
  if (fuse_cv) {
    // the [B,2C,D,H,W] volume is never written: dres0.0's TMA producer assembles each plane from the two feature maps
    mark(-1);
    if (p->x2) {
      RUN(launch_ncdhw_to_blocked_x2(left, (__nv_bfloat16 *)b.feaL, B, C, (int64_t)Hf * Wf, s, p->range_flag)); ++launches;
      RUN(launch_ncdhw_to_blocked_x2(right, (__nv_bfloat16 *)b.feaR, B, C, (int64_t)Hf * Wf, s, p->range_flag)); ++launches;
    } else {
      RUN(launch_ncdhw_to_blocked_h(left, (__nv_bfloat16 *)b.feaL, B, C, (int64_t)Hf * Wf, p->f16, s)); ++launches;
      RUN(launch_ncdhw_to_blocked_h(right, (__nv_bfloat16 *)b.feaR, B, C, (int64_t)Hf * Wf, p->f16, s)); ++launches;
    }
  } else {
    if (p->f16) { set_error("plan_forward: fp16 mode needs the fused cost volume (C in {16,32}, D <= 64)"); return IDISP_ERR_UNSUPPORTED; }
    mark(-1);
    RUN(launch_cost_volume_blocked<T>(left, right, B, C, Hf, Wf, p->mindisp, D, (T *)b.cv, s)); ++launches;
  }
  // ---- the conv section: reads/writes only the workspace -> one CUDA graph per (B, Hf, Wf, workspace) ----
  auto conv_section = [&]() -> int {
  if (fuse_cv) {
    TcCostVolume cvd;
    cvd.left = (const __nv_bfloat16 *)b.feaL; cvd.right = (const __nv_bfloat16 *)b.feaR;
    cvd.shift0 = p->mindisp >= 0 ? p->mindisp / 4 : -((-p->mindisp + 3) / 4);
    mark(0);
    ++launches;
    RUN(tc_layer(p, 0, nullptr, 0, &cvd, B, D, Hf, Wf, nullptr, 1, (__nv_bfloat16 *)b.a, nullptr, 0, nullptr, nullptr, b.split, b.part, s, launches));
  } else {
    RUN(conv(0, b.cv, D, Hf, Wf, nullptr, 1, b.a));
  }
  // dres0 (second conv), dres1 (:130-131)
  RUN(conv(1, b.a, D, Hf, Wf, nullptr, 1, b.t0));
  RUN(conv(2, b.t0, D, Hf, Wf, nullptr, 1, b.a));
  // fused hand-off: cost0 is stored ONLY in the parity layout (in b.cost0): its readers are dres2.conv1 (stride 2) and the
  // residual adds of the three conv6 epilogues, which all address that layout directly
  RUN(conv(3, b.a, D, Hf, Wf, b.t0, 0, b.cost0, fuse_split, false, fuse_split ? 4 : 0, b.cost0));
  // three hourglasses (:133-140); `x` is cost0 / out1 / out2, all outputs land in b.out
  const int D2 = D / 2, H2 = Hf / 2, W2 = Wf / 2, D4 = D / 4, H4 = Hf / 4, W4 = Wf / 4;
  for (int k = 0; k < 3; ++k) {
    const int l0 = 4 + 6 * k;
    const void *x = k == 0 ? b.cost0 : b.out;
    void *pre = k == 0 ? b.pre1 : b.prek;
    const void *postsqu = k == 0 ? nullptr : (k == 1 ? b.postA : b.postB);  // post1 / post2
    void *post = k == 1 ? b.postB : b.postA;                                  // post1,post3 -> A; post2 -> B
    const void *presqu = k == 0 ? b.pre1 /* own pre */ : b.pre1;              // pre1 for dres3 AND dres4 (:136,:139)
    if (fuse_split && k == 0)  // dres2.conv1 reads cost0's parity layout in place
    { mark(l0); ++launches;
      RUN(tc_layer(p, l0, (const __nv_bfloat16 *)b.cost0, 1, nullptr, B, D, Hf, Wf, nullptr, 1, (__nv_bfloat16 *)b.h1, nullptr, 0, nullptr, nullptr,
                   b.split, b.part, s, launches)); }
    else
      RUN(conv(l0 + 0, x, D, Hf, Wf, nullptr, 1, b.h1, false, fuse_split));        // reads b.split (out_k)
    RUN(conv(l0 + 1, b.h1, D2, H2, W2, postsqu, 1, pre, fuse_split));              // writes b.split (pre_k)
    RUN(conv(l0 + 2, pre, D2, H2, W2, nullptr, 1, b.q1, false, fuse_split));       // reads b.split
    RUN(conv(l0 + 3, b.q1, D4, H4, W4, nullptr, 1, b.q2));
    RUN(conv(l0 + 4, b.q2, D4, H4, W4, presqu, 1, post));
    // out_k has two readers: classif{k+1}.0 (natural layout) and the next hourglass's stride-2 conv1 (parity layout).  conv6 is memory-
    // bound, so in the split-precision mode it writes only the natural copy and classif{k+1}.0 -- MMA-bound, DRAM at 30 % -- writes the
    // parity copy of its input as the tiles pass through its shared memory (side_copy); otherwise conv6's epilogue writes both.
    const bool side_copy = fuse_split && k < 2 && p->x2 && !p->no_side_copy;
    RUN(conv(l0 + 5, post, D2, H2, W2, b.cost0, 0, b.out, fuse_split && k < 2 && !side_copy, false, fuse_split ? 2 : 0));  // (+ b.split = out_k)
    // classifier head k on out_k (:142-144), running sum fused
    if (side_copy) p->x_copy_next = (__nv_bfloat16 *)b.split;
    RUN(conv(22 + k, b.out, D, Hf, Wf, nullptr, 1, b.c));
    float *dst = (k == 1) ? b.costY : b.costX;
    const float *prev = k == 0 ? nullptr : (k == 1 ? b.costX : b.costY);
    mark(25 + k);
    if (p->x2 && p->x2_simt_heads)  // A/B switch
      // split precision, 1-channel head on the CUDA cores straight from the hi|lo words (f32 weights and FMAs).  Measured at
      // B=32: 3.8 ms against 1.36 ms for the tensor-core form (w_lo in output column 1) -- kept only as a cross-check.
      RUN(launch_conv3d_to1_x2((const __nv_bfloat16 *)b.c, B, 32, D, Hf, Wf, p->dev[25 + k].w_tap, prev, dst, s));
    else if (std::is_same<T, __nv_bfloat16>::value && tc_head_supported(p->dev[25 + k].head, D, Hf, Wf))
      RUN(tc_head_conv(p->dev[25 + k].head, (const __nv_bfloat16 *)b.c, B, D, Hf, Wf, prev, dst, s));
    else if (std::is_same<T, __nv_bfloat16>::value && tc_supported(IDISP_CONV_S1, 32, 1, D, Hf, Wf))
      RUN(tc_layer(p, 25 + k, (const __nv_bfloat16 *)b.c, 0, nullptr, B, D, Hf, Wf, nullptr, 0, nullptr, nullptr, 0, prev, dst, b.split, b.part, s,
                   launches));
    else if (p->f16) { set_error("plan_forward: fp16 mode needs the tensor-core classifier head"); return IDISP_ERR_UNSUPPORTED; }
    else
      RUN(launch_conv3d_to1<T>((const T *)b.c, B, 32, D, Hf, Wf, p->dev[25 + k].w_tap, prev, dst, s));
    ++launches;
  }
  return IDISP_OK;
  };  // conv_section
  {
    cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
    const bool graph_ok = !p->no_graph && !p->timing && std::is_same<T, __nv_bfloat16>::value &&
                          cudaStreamIsCapturing(s, &cst) == cudaSuccess && cst == cudaStreamCaptureStatusNone;
    if (!graph_ok) {
      RUN(conv_section());
    } else {
      idisp_plan::GraphEntry *hit = nullptr;
      for (auto &g : p->graphs)
        if (g.B == B && g.Hf == Hf && g.Wf == Wf && g.ws == workspace) { hit = &g; break; }
      if (!hit) {
        // Capture on the plan's own stream (nothing executes there; the caller's stream may be the legacy default stream, on
        // which capture is not permitted); the instantiated graph is then launched into the caller's stream.
        const int before = launches;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool ok = p->cap_stream || cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          cudaStream_t user = s;
          s = p->cap_stream;
          const int crc = conv_section();
          s = user;
          const cudaError_t ce = cudaStreamEndCapture(p->cap_stream, &graph);
          if (crc != IDISP_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return crc; }
          ok = ce == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
          if (graph) cudaGraphDestroy(graph);
        }
        if (!ok) {  // no graphs on this driver / in this context: launch eagerly from now on
          cudaGetLastError();
          p->no_graph = true;
          launches = before;
          RUN(conv_section());
        } else {
          if (p->graphs.size() >= 48) {  // least recently used out
            size_t lru = 0;
            for (size_t i = 1; i < p->graphs.size(); ++i) if (p->graphs[i].stamp < p->graphs[lru].stamp) lru = i;
            cudaGraphExecDestroy(p->graphs[lru].exec);
            p->graphs.erase(p->graphs.begin() + lru);
          }
          p->graphs.push_back({B, Hf, Wf, workspace, exec, launches - before, 0ull});
          hit = &p->graphs.back();
          ++p->graph_captures;
          launches = before;
        }
      } else {
        ++p->graph_hits;
      }
      if (hit) {
        hit->stamp = ++p->graph_clock;
        IDISP_CUDA(cudaGraphLaunch(hit->exec, s));
        launches += hit->launches;
      }
    }
  }
  // upsample + softmax + regression (:169-174)
  mark(-2);
  // (b.part -- the fp32 partial of the multi-launch layers, >= B*32*V floats -- is free here: scratch for the per-cell maxima)
  RUN(launch_softargmin(b.costX, B, D, Hf, Wf, p->mindisp, p->maxdisp, H, W, out, s, (float *)b.part)); ++launches;
  if (p->maxdisp - p->mindisp == 4 * D && (D == 24 || D == 48) && !getenv("IDISP_SOFTARGMIN_GENERIC") && !getenv("IDISP_SOFTARGMIN_NO_CELLMAX")) ++launches;  // + cell_max_kernel
  mark(-99);  // closing event
#undef RUN
  p->last_logits = b.costX; p->last_B = B; p->last_Hf = Hf; p->last_Wf = Wf;
  p->launches = launches;
  return IDISP_OK;
}

extern "C" int idisp_plan_forward(idisp_plan_t *p, const float *left, const float *right, int B, int Hf, int Wf,
                                  int H, int W, void *workspace, size_t workspace_bytes, float *out, void *stream)
{
  IDISP_REQUIRE(p != nullptr, "plan_forward: NULL plan");
  if (!p->finalized) { set_error("plan_forward: plan not finalised (call idisp_plan_finalize after loading weights)"); return IDISP_ERR_STATE; }
  IDISP_REQUIRE(B >= 0 && Hf > 0 && Wf > 0 && H >= Hf && W >= Wf, "plan_forward: bad shape B=%d Hf=%d Wf=%d H=%d W=%d", B, Hf, Wf, H, W);
  IDISP_REQUIRE(Hf % 4 == 0 && Wf % 4 == 0, "plan_forward: Hf=%d, Wf=%d must be multiples of 4 (stackhourglass.py:34-49)", Hf, Wf);
  if (B == 0) return IDISP_OK;
  IDISP_REQUIRE(left && right && out && workspace, "plan_forward: NULL pointer");
  const size_t needb = idisp_plan_workspace_bytes(p, B, Hf, Wf);
  IDISP_REQUIRE(workspace_bytes >= needb, "plan_forward: workspace %zu B < required %zu B", workspace_bytes, needb);
  if (p->precision == IDISP_PREC_FP32)
    return forward_impl<float>(p, left, right, B, Hf, Wf, H, W, workspace, out, (cudaStream_t)stream);
  return forward_impl<__nv_bfloat16>(p, left, right, B, Hf, Wf, H, W, workspace, out, (cudaStream_t)stream);
}

extern "C" int idisp_plan_range_exceeded(idisp_plan_t *p, int *exceeded, void *stream)
{
  IDISP_REQUIRE(p != nullptr && exceeded != nullptr, "plan_range_exceeded: NULL argument");
  *exceeded = 0;
  if (!p->range_flag) return IDISP_OK;  // fp32 / bf16 plans, or no forward yet
  cudaStream_t s = (cudaStream_t)stream;
  IDISP_CUDA(cudaMemcpyAsync(exceeded, p->range_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  IDISP_CUDA(cudaStreamSynchronize(s));
  return IDISP_OK;
}

extern "C" int idisp_plan_forward_host(idisp_plan_t *p, const float *left_host, const float *right_host, int B,
                                       int Hf, int Wf, int H, int W, float *out_host, void *stream)
{
  IDISP_REQUIRE(p != nullptr, "plan_forward_host: NULL plan");
  if (B == 0) return IDISP_OK;
  IDISP_REQUIRE(left_host && right_host && out_host && B > 0 && Hf > 0 && Wf > 0 && H > 0 && W > 0, "plan_forward_host: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t fea = (size_t)B * p->C * Hf * Wf * sizeof(float), outb = (size_t)B * H * W * sizeof(float);
  const size_t ws = idisp_plan_workspace_bytes(p, B, Hf, Wf);
  auto up = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  const size_t total = up(fea) * 2 + up(outb) + ws;
  if (total > p->stage_bytes) {
    IDISP_CUDA(cudaStreamSynchronize(s));
    if (p->stage) cudaFree(p->stage);
    p->stage = nullptr; p->stage_bytes = 0;
    IDISP_CUDA(cudaMalloc(&p->stage, total));
    p->stage_bytes = total;
  }
  char *base = (char *)p->stage;
  float *dl = (float *)base, *dr = (float *)(base + up(fea)), *dout = (float *)(base + 2 * up(fea));
  void *wsp = base + 2 * up(fea) + up(outb);
  IDISP_CUDA(cudaMemcpyAsync(dl, left_host, fea, cudaMemcpyHostToDevice, s));
  IDISP_CUDA(cudaMemcpyAsync(dr, right_host, fea, cudaMemcpyHostToDevice, s));
  int rc = idisp_plan_forward(p, dl, dr, B, Hf, Wf, H, W, wsp, ws, dout, stream);
  if (rc) return rc;
  IDISP_CUDA(cudaMemcpyAsync(out_host, dout, outb, cudaMemcpyDeviceToHost, s));
  return IDISP_OK;
}

// Pipelined host-buffer call for a STREAM of batches (the reference's data loader hands DispRCNN3D one batch after another,
// engine/inference.py:24-50).  Call i uses staging slot i & 1:
//   copy-in stream : wait "kernels of call i-2 done"  -> H2D left/right -> record in[slot]
//   caller's stream: wait in[slot], wait "result copy of call i-2 done" -> forward -> record done[slot]
//   copy-out stream: wait done[slot] -> D2H result -> record out[slot]
// so the H2D of call i overlaps the kernels of call i-1 and the D2H of call i overlaps the kernels of call i+1; the kernels
// themselves stay serialised on the caller's stream (one workspace).  Nothing blocks the host here.
extern "C" int idisp_plan_forward_host_async(idisp_plan_t *p, const float *left_host, const float *right_host, int B, int Hf, int Wf,
                                             int H, int W, float *out_host, void *stream, unsigned long long *ticket)
{
  IDISP_REQUIRE(p != nullptr && ticket != nullptr, "plan_forward_host_async: NULL plan / ticket");
  IDISP_REQUIRE(left_host && right_host && out_host && B > 0 && Hf > 0 && Wf > 0 && H > 0 && W > 0, "plan_forward_host_async: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (!p->s_in) {
    IDISP_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    IDISP_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      IDISP_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
      IDISP_CUDA(cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming));
      IDISP_CUDA(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
    }
  }
  const size_t fea = (size_t)B * p->C * Hf * Wf * sizeof(float), outb = (size_t)B * H * W * sizeof(float);
  const size_t ws = idisp_plan_workspace_bytes(p, B, Hf, Wf);
  auto up = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  const size_t slot_bytes = 2 * up(fea) + up(outb), total = 2 * slot_bytes + ws;
  if (total > p->pstage_bytes) {  // (re)allocation: drain everything that may still use the old staging
    IDISP_CUDA(cudaStreamSynchronize(p->s_in));
    IDISP_CUDA(cudaStreamSynchronize(s));
    IDISP_CUDA(cudaStreamSynchronize(p->s_out));
    if (p->pstage) cudaFree(p->pstage);
    p->pstage = nullptr; p->pstage_bytes = 0;
    IDISP_CUDA(cudaMalloc(&p->pstage, total));
    p->pstage_bytes = total;
    p->async_base = p->n_async;   // no earlier call owns a slot of the new buffer
  }
  const unsigned long long i = p->n_async;
  const int slot = (int)(i & 1);
  char *base = (char *)p->pstage + (size_t)slot * slot_bytes;
  float *dl = (float *)base, *dr = (float *)(base + up(fea)), *dout = (float *)(base + 2 * up(fea));
  void *wsp = (char *)p->pstage + 2 * slot_bytes;
  const bool reuse = i - p->async_base >= 2;   // the slot's previous owner (call i-2) may still be in flight
  if (reuse) IDISP_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_done[slot], 0));
  IDISP_CUDA(cudaMemcpyAsync(dl, left_host, fea, cudaMemcpyHostToDevice, p->s_in));
  IDISP_CUDA(cudaMemcpyAsync(dr, right_host, fea, cudaMemcpyHostToDevice, p->s_in));
  IDISP_CUDA(cudaEventRecord(p->ev_in[slot], p->s_in));
  IDISP_CUDA(cudaStreamWaitEvent(s, p->ev_in[slot], 0));
  if (reuse) IDISP_CUDA(cudaStreamWaitEvent(s, p->ev_out[slot], 0));
  const int rc = idisp_plan_forward(p, dl, dr, B, Hf, Wf, H, W, wsp, ws, dout, stream);
  if (rc) return rc;
  IDISP_CUDA(cudaEventRecord(p->ev_done[slot], s));
  IDISP_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_done[slot], 0));
  IDISP_CUDA(cudaMemcpyAsync(out_host, dout, outb, cudaMemcpyDeviceToHost, p->s_out));
  IDISP_CUDA(cudaEventRecord(p->ev_out[slot], p->s_out));
  *ticket = i;
  p->n_async = i + 1;
  return IDISP_OK;
}

// Blocks the host until the result of the call that returned `ticket` has landed in its out_host buffer.
extern "C" int idisp_plan_host_wait(idisp_plan_t *p, unsigned long long ticket)
{
  IDISP_REQUIRE(p != nullptr, "plan_host_wait: NULL plan");
  IDISP_REQUIRE(ticket < p->n_async, "plan_host_wait: ticket %llu was never issued by this plan", ticket);
  // (the slot's event may already have been re-recorded by call ticket+2: the copy-out stream is in order, so that later
  //  record completing implies this call's copy has completed)
  IDISP_CUDA(cudaEventSynchronize(p->ev_out[ticket & 1]));
  return IDISP_OK;
}

extern "C" int idisp_plan_get_logits(idisp_plan_t *p, float *logits, void *stream)
{
  IDISP_REQUIRE(p && logits, "plan_get_logits: NULL argument");
  if (!p->last_logits) { set_error("plan_get_logits: no forward has run on this plan"); return IDISP_ERR_STATE; }
  IDISP_CUDA(cudaMemcpyAsync(logits, p->last_logits, (size_t)p->last_B * p->D * p->last_Hf * p->last_Wf * sizeof(float),
                             cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return IDISP_OK;
}

extern "C" int idisp_plan_launches_per_forward(const idisp_plan_t *p) { return p ? p->launches : 0; }

extern "C" int idisp_plan_graph_stats(const idisp_plan_t *p, int *captures, int *replays)
{
  IDISP_REQUIRE(p != nullptr, "plan_graph_stats: NULL plan");
  if (captures) *captures = p->graph_captures;
  if (replays) *replays = p->graph_hits;
  return IDISP_OK;
}

extern "C" int idisp_plan_enable_timing(idisp_plan_t *p, int on)
{
  IDISP_REQUIRE(p != nullptr, "plan_enable_timing: NULL plan");
  p->timing = on != 0;
  return IDISP_OK;
}

extern "C" int idisp_plan_get_timing(idisp_plan_t *p, float *ms, int *layer, int capacity)
{
  IDISP_REQUIRE(p && ms && layer, "plan_get_timing: NULL argument");
  const int n = (int)p->ev_layer.size();
  if (!p->timing || n == 0 || (int)p->ev.size() < n + 1) { set_error("plan_get_timing: no timed forward on this plan"); return IDISP_ERR_STATE; }
  IDISP_REQUIRE(capacity >= n, "plan_get_timing: capacity %d < %d launches", capacity, n);
  IDISP_CUDA(cudaEventSynchronize(p->ev[n]));
  for (int i = 0; i < n; ++i) {
    IDISP_CUDA(cudaEventElapsedTime(&ms[i], p->ev[i], p->ev[i + 1]));
    layer[i] = p->ev_layer[i];
  }
  return IDISP_OK;
}

// ---------------------------------------------------------------------------------------
// per-layer test hook (NCDHW f32 in/out)
// ---------------------------------------------------------------------------------------
template <typename T>
static int conv3d_hook(const float *x, int B, int Cin, int D, int H, int W, const std::vector<float> &w_tap, int Cout,
                       int kind, const std::vector<float> &bias, const float *residual, int relu, int precision,
                       float *y, cudaStream_t s)
{
  int Do = D, Ho = H, Wo = W;
  if (kind == IDISP_CONV_S2) { Do = (D + 1) / 2; Ho = (H + 1) / 2; Wo = (W + 1) / 2; }
  if (kind == IDISP_DECONV_S2) { Do = 2 * D; Ho = 2 * H; Wo = 2 * W; }
  const int64_t Vi = (int64_t)D * H * W, Vo = (int64_t)Do * Ho * Wo;
  T *xb = nullptr, *yb = nullptr, *rb = nullptr;
  float *wd = nullptr, *bd = nullptr, *y1 = nullptr;
  void *scratch = nullptr;
  int rc = IDISP_OK;
  TcWeights tcw;
  auto cleanup = [&]() { cudaFree(xb); cudaFree(yb); cudaFree(rb); cudaFree(wd); cudaFree(bd); cudaFree(y1); cudaFree(scratch); tc_weights_free(tcw); };
#define HK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr, __FILE__, __LINE__); } } while (0)
#define HR(expr) do { if ((rc = (expr)) != IDISP_OK) { cudaStreamSynchronize(s); cleanup(); return rc; } } while (0)
  HK(cudaMalloc(&xb, (size_t)B * Cin * Vi * sizeof(T)));
  HK(cudaMalloc(&wd, w_tap.size() * sizeof(float)));
  HK(cudaMemcpyAsync(wd, w_tap.data(), w_tap.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  HK(cudaMalloc(&bd, bias.size() * sizeof(float)));
  HK(cudaMemcpyAsync(bd, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  const int f16 = precision == IDISP_PREC_FP16;
  const bool tcp = precision != IDISP_PREC_FP32 && tc_supported(kind, Cin, Cout, D, H, W);
  if (f16 && !tcp) { cleanup(); set_error("conv3d: fp16 precision needs a tensor-core-supported layer shape"); return IDISP_ERR_UNSUPPORTED; }
  if (sizeof(T) == 2) HR(launch_ncdhw_to_blocked_h(x, (__nv_bfloat16 *)xb, B, Cin, Vi, f16, s));
  else HR(launch_ncdhw_to_blocked<T>(x, xb, B, Cin, Vi, s));
  if (Cout == 1) {
    TcHeadWeights hw;
    if (tcp && Cin == 32) {
      rc = tc_head_weights_prepare(w_tap.data(), Cin, f16, 0, hw, s);
      if (rc == IDISP_OK && tc_head_supported(hw, D, H, W)) {
        rc = tc_head_conv(hw, (const __nv_bfloat16 *)xb, B, D, H, W, residual, y, s);
        cudaStreamSynchronize(s);
        tc_head_weights_free(hw);
        if (rc) { cleanup(); return rc; }
        cleanup();
        return IDISP_OK;
      }
      tc_head_weights_free(hw);
      if (rc) { cleanup(); return rc; }
    }
    if (tcp) {
      HR(tc_weights_prepare(w_tap.data(), kind, Cin, Cout, f16, tcw, s));
      HR(tc_conv3d(tcw, (const __nv_bfloat16 *)xb, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, residual, y, nullptr, 0, nullptr, nullptr, s));
    } else {
      HR(launch_conv3d_to1<T>(xb, B, Cin, D, H, W, wd, residual, y, s));
    }
    // bias/relu for the 1-channel hook are not part of any reference layer
  } else {
    HK(cudaMalloc(&yb, (size_t)B * Cout * Vo * sizeof(T)));
    if (residual) {
      HK(cudaMalloc(&rb, (size_t)B * Cout * Vo * sizeof(T)));
      if (sizeof(T) == 2) HR(launch_ncdhw_to_blocked_h(residual, (__nv_bfloat16 *)rb, B, Cout, Vo, f16, s));
      else HR(launch_ncdhw_to_blocked<T>(residual, rb, B, Cout, Vo, s));
    }
    if (tcp) {
      HR(tc_weights_prepare(w_tap.data(), kind, Cin, Cout, f16, tcw, s));
      const size_t sb = tc_scratch_bytes(kind, B, Cin, D, H, W);
      if (sb) HK(cudaMalloc(&scratch, sb));
      HR(tc_conv3d(tcw, (const __nv_bfloat16 *)xb, B, Cin, D, H, W, Cout, kind, bd, (const __nv_bfloat16 *)rb, relu,
                   (__nv_bfloat16 *)yb, nullptr, nullptr, scratch, 0, nullptr, nullptr, s));
    } else {
      HR(launch_conv3d_simt<T>(xb, B, Cin, D, H, W, wd, Cout, kind, bd, rb, relu, yb, s));
    }
    if (sizeof(T) == 2) HR(launch_blocked_to_ncdhw_h((const __nv_bfloat16 *)yb, y, B, Cout, Vo, f16, s));
    else HR(launch_blocked_to_ncdhw<T>(yb, y, B, Cout, Vo, s));
  }
  HK(cudaStreamSynchronize(s));
  cleanup();
#undef HK
#undef HR
  return IDISP_OK;
}

// split-precision variant of the hook: hi|lo IEEE-half words, three tensor-core passes (see tc_layer)
static int conv3d_hook_x2(const float *x, int B, int Cin, int D, int H, int W, const std::vector<float> &w_tap, int Cout, int kind,
                          const std::vector<float> &bias, const float *residual, int relu, float *y, cudaStream_t s)
{
  int Do = D, Ho = H, Wo = W;
  if (kind == IDISP_CONV_S2) { Do = (D + 1) / 2; Ho = (H + 1) / 2; Wo = (W + 1) / 2; }
  if (kind == IDISP_DECONV_S2) { Do = 2 * D; Ho = 2 * H; Wo = 2 * W; }
  const int64_t Vi = (int64_t)D * H * W, Vo = (int64_t)Do * Ho * Wo;
  if (!tc_supported(kind, Cin, Cout, D, H, W)) { set_error("conv3d: fp16x2 precision needs a tensor-core-supported layer shape"); return IDISP_ERR_UNSUPPORTED; }
  __nv_bfloat16 *xb = nullptr, *yb = nullptr, *rb = nullptr;
  float *bd = nullptr, *part = nullptr;
  void *scratch = nullptr;
  int rc = IDISP_OK;
  TcSplitWeights sw;
  auto cleanup = [&]() { cudaFree(xb); cudaFree(yb); cudaFree(rb); cudaFree(bd); cudaFree(part); cudaFree(scratch); tc_split_weights_free(sw); };
#define HK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr, __FILE__, __LINE__); } } while (0)
#define HR(expr) do { if ((rc = (expr)) != IDISP_OK) { cudaStreamSynchronize(s); cleanup(); return rc; } } while (0)
  HK(cudaMalloc(&xb, (size_t)B * Cin * Vi * 4));
  HK(cudaMalloc(&bd, bias.size() * sizeof(float)));
  HK(cudaMemcpyAsync(bd, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  HR(launch_ncdhw_to_blocked_x2(x, xb, B, Cin, Vi, s));
  HR(tc_split_weights_prepare(w_tap.data(), kind, Cin, Cout, sw, s));
  const size_t sb = 2 * tc_scratch_bytes(kind, B, Cin, D, H, W);
  if (sb) HK(cudaMalloc(&scratch, sb));
  if (Cout == 1) {
    TcHeadWeights hw;
    rc = Cin == 32 ? tc_head_weights_prepare(w_tap.data(), Cin, 1, 1, hw, s) : IDISP_OK;
    if (rc == IDISP_OK && tc_head_supported(hw, D, H, W)) {
      rc = tc_head_conv(hw, xb, B, D, H, W, residual, y, s);
      cudaStreamSynchronize(s);
      tc_head_weights_free(hw);
      if (rc) { cleanup(); return rc; }
    } else {
      tc_head_weights_free(hw);
      if (rc) { cleanup(); return rc; }
      HR(tc_conv3d_split(sw, xb, B, Cin, D, H, W, Cout, kind, nullptr, nullptr, 0, nullptr, residual, y, scratch, 0, nullptr, nullptr, nullptr, s));
    }
  } else {
    HK(cudaMalloc(&yb, (size_t)B * Cout * Vo * 4));
    HK(cudaMalloc(&part, (size_t)B * Cout * Vo * 4));
    if (residual) {
      HK(cudaMalloc(&rb, (size_t)B * Cout * Vo * 4));
      HR(launch_ncdhw_to_blocked_x2(residual, rb, B, Cout, Vo, s));
    }
    HR(tc_conv3d_split(sw, xb, B, Cin, D, H, W, Cout, kind, bd, rb, relu, yb, nullptr, nullptr, scratch, 0, nullptr, nullptr, part, s));
    HR(launch_blocked_x2_to_ncdhw(yb, y, B, Cout, Vo, s));
  }
  HK(cudaStreamSynchronize(s));
  cleanup();
#undef HK
#undef HR
  return IDISP_OK;
}

extern "C" int idisp_conv3d(const float *x, int B, int Cin, int D, int H, int W, const float *weight, int Cout, int kind,
                            const float *scale, const float *bias, const float *residual, int relu, int precision,
                            float *y, void *stream)
{
  IDISP_REQUIRE(B >= 0 && Cin > 0 && Cin % 8 == 0 && D > 0 && H > 0 && W > 0, "conv3d: bad input shape B=%d Cin=%d D=%d H=%d W=%d", B, Cin, D, H, W);
  IDISP_REQUIRE(Cout == 1 || Cout % 8 == 0, "conv3d: Cout=%d must be 1 or a multiple of 8", Cout);
  IDISP_REQUIRE(kind == IDISP_CONV_S1 || kind == IDISP_CONV_S2 || kind == IDISP_DECONV_S2, "conv3d: unknown kind %d", kind);
  IDISP_REQUIRE(precision == IDISP_PREC_FP32 || precision == IDISP_PREC_BF16 || precision == IDISP_PREC_FP16 || precision == IDISP_PREC_FP16X2, "conv3d: unknown precision %d", precision);
  IDISP_REQUIRE(Cout != 1 || (kind == IDISP_CONV_S1 && !scale && !bias && !relu), "conv3d: the 1-channel conv is stride-1, no affine, no ReLU");
  if (B == 0) return IDISP_OK;
  IDISP_REQUIRE(x && weight && y, "conv3d: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t wn = (size_t)27 * Cin * Cout;
  std::vector<float> hw(wn), hs(Cout, 1.f), hb(Cout, 0.f);
  IDISP_CUDA(cudaStreamSynchronize(s));
  IDISP_CUDA(cudaMemcpy(hw.data(), weight, wn * sizeof(float), cudaMemcpyDeviceToHost));
  if (scale) IDISP_CUDA(cudaMemcpy(hs.data(), scale, Cout * sizeof(float), cudaMemcpyDeviceToHost));
  if (bias) IDISP_CUDA(cudaMemcpy(hb.data(), bias, Cout * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<double> sc(hs.begin(), hs.end());
  std::vector<float> w_tap;
  relayout_taps(hw.data(), kind, Cin, Cout, sc.data(), w_tap);
  if (precision == IDISP_PREC_FP16X2) return conv3d_hook_x2(x, B, Cin, D, H, W, w_tap, Cout, kind, hb, residual, relu, y, s);
  if (precision == IDISP_PREC_FP32)
    return conv3d_hook<float>(x, B, Cin, D, H, W, w_tap, Cout, kind, hb, residual, relu, precision, y, s);
  return conv3d_hook<__nv_bfloat16>(x, B, Cin, D, H, W, w_tap, Cout, kind, hb, residual, relu, precision, y, s);
}

// ---------------------------------------------------------------------------------------
// test hook: the cost volume AS THE TENSOR-CORE PATH SEES IT.  Runs dres0.0's fused TMA loader with centre-tap identity
// kernels (no BN, no ReLU) so that the layer output equals its input, i.e. the bf16 cost volume, and returns it NCDHW f32.
// ---------------------------------------------------------------------------------------
extern "C" int idisp_debug_fused_cost_volume(const float *left, const float *right, int B, int C, int Hf, int Wf, int mindisp,
                                             int maxdisp, float *cost, void *stream)
{
  IDISP_REQUIRE(B > 0 && (C == 16 || C == 32) && Hf > 0 && Wf > 0 && maxdisp > mindisp && mindisp % 4 == 0 && maxdisp % 4 == 0,
                "debug_fused_cost_volume: unsupported shape (C must be 16 or 32)");
  IDISP_REQUIRE(left && right && cost, "debug_fused_cost_volume: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int D = (maxdisp - mindisp) / 4, cin = 2 * C;
  IDISP_REQUIRE(D <= 64 && tc_supported(IDISP_CONV_S1, cin, 32, D, Hf, Wf), "debug_fused_cost_volume: layer not on the tensor-core path");
  const int64_t V = (int64_t)D * Hf * Wf;
  __nv_bfloat16 *fl = nullptr, *fr = nullptr, *yb = nullptr;
  TcWeights tcw;
  int rc = IDISP_OK;
  auto cleanup = [&]() { cudaFree(fl); cudaFree(yb); tc_weights_free(tcw); };
#define DK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr, __FILE__, __LINE__); } } while (0)
#define DR(expr) do { if ((rc = (expr)) != IDISP_OK) { cudaStreamSynchronize(s); cleanup(); return rc; } } while (0)
  const size_t fbytes = ((size_t)B * C * Hf * Wf * 2 + 1023) / 1024 * 1024;
  DK(cudaMalloc(&fl, 2 * fbytes));  // the fused loader wants the right view's features after the left ones
  fr = fl + fbytes / 2;
  DK(cudaMalloc(&yb, (size_t)B * 32 * V * 2));
  DR(launch_ncdhw_to_blocked<__nv_bfloat16>(left, fl, B, C, (int64_t)Hf * Wf, s));
  DR(launch_ncdhw_to_blocked<__nv_bfloat16>(right, fr, B, C, (int64_t)Hf * Wf, s));
  TcCostVolume cvd;
  cvd.left = fl; cvd.right = fr;
  cvd.shift0 = mindisp >= 0 ? mindisp / 4 : -((-mindisp + 3) / 4);
  for (int c0 = 0; c0 < cin; c0 += 32) {  // 32 output channels per pass: input channels [c0, c0+32) copied through
    std::vector<float> w((size_t)27 * cin * 32, 0.f);
    const char *te = getenv("IDISP_DEBUG_TAP");  // which of the 27 taps carries the identity (default: centre)
    const int tap = te ? atoi(te) : 13;
    for (int co = 0; co < 32; ++co) w[((size_t)tap * cin + c0 + co) * 32 + co] = 1.f;
    DR(tc_weights_prepare(w.data(), IDISP_CONV_S1, cin, 32, 0, tcw, s));
    DR(tc_conv3d(tcw, nullptr, B, cin, D, Hf, Wf, 32, IDISP_CONV_S1, nullptr, nullptr, 0, yb, nullptr, nullptr, nullptr, 0, nullptr, &cvd, s));
    // blocked [B][4][V][8] -> channels [c0, c0+32) of NCDHW [B][2C][V]
    for (int b = 0; b < B; ++b)
      DR(launch_blocked_to_ncdhw<__nv_bfloat16>(yb + (size_t)b * 32 * V, cost + ((size_t)b * cin + c0) * V, 1, 32, V, s));
  }
  DK(cudaStreamSynchronize(s));
  cleanup();
#undef DK
#undef DR
  return IDISP_OK;
}
