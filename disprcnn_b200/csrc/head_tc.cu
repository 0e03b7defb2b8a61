// head_tc.cu -- the 32 -> 1 classifier convolutions (classif{1,2,3}.2: Conv3d(32, 1, k3, p1, bias=False),
// disprcnn/modeling/psmnet/stackhourglass.py:80,84,88, applied :142-144 with the running sums cost2 += cost1, cost3 += cost2)
// as ONE small GEMM per input plane followed by a shifted sum, for the tensor-core modes.
//
// Why not the generic 3x3x3 kernel (conv3d_tc.cu): with a single output channel its MMAs are N = 16..48 wide, so the
// 128 B/clk shared-memory port spends 4 KB of A traffic on every one of the 36 MMAs per plane tile (measured 1.15-1.2 ms per
// head at BASELINE configs[1], 1.9 % of the tensor peak, tensor pipe 38 %).  Here the 27 taps move from K into N:
//     P[v, t] = sum_c x[v, c] * w[t, c]          (M = voxels of the HALOED tile, K = 32 channels, N = 27 taps -> 32 columns)
//     out[z, y, x] = sum_{kd,kh,kw} P[(z + kd - 1, y + kh - 1, x + kw - 1), (kd, kh, kw)]
// i.e. every input voxel is multiplied with all 27 taps at once (4 MMAs per k-step pair instead of 36) and the epilogue adds
// each product to the output voxel it belongs to: in-plane through shared memory, across planes through three rolling
// register sums.  Split precision (fp16x2): x_hi feeds [w_hi | w_lo] in one N = 64 MMA, x_lo feeds w_hi (N = 32, added into the
// x_hi*w_lo columns: both are 2^-11-sized correction terms, four adds per column); main and correction columns are summed in
// fp32 round-to-nearest.  The epilogue is bounded by the TMEM read port (64 B/clk: /opt/skills/guides/B300_MICROARCH.md), so it
// reads as few columns as it can: 64 per M half, and the warps whose rows lie beyond the 180 haloed voxels read nothing.
//
// Tiling: a CTA walks one (n, 16-row, 8-column) output tile through all D planes.  Its haloed input tile is 18 x 10 = 180
// voxels = GEMM rows 0..179, i.e. two M = 128 MMAs per operand pair (rows >= 180 read whatever follows in shared memory and
// are never used).  One TMA box per plane lands [blocks][18][10][8 ch] = consecutive 16-byte rows: exactly the no-swizzle
// K-major A layout (SBO = 128 B, LBO = one channel-block plane), zero fill outside the volume = the conv padding.
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue group A (TMEM -> tap-major products in
// shared memory), warps 8-11 epilogue group B (shifted sums -> logits); A and B hand two product buffers back and forth through
// named barriers, so a plane's extraction overlaps the previous plane's gather.  Accumulators are
// double-buffered (2 x 192 TMEM columns); no zero-on-drain: the first k-step of a plane overwrites (accumulate = 0).
// Roofline: HBM (reads the 32-channel activation once: 2.47 GB per 32 ROIs at configs[1] = 0.38 ms), not tensor.
#include "conv3d_tc.cuh"
#include "sm100_ptx.cuh"

#include <vector>

namespace idisp {
namespace headtc {

constexpr int TW = 8, TH = 16, SW = TW + 2, SH = TH + 2;   // output tile, haloed tile
constexpr int NVOX = SW * SH;                               // 180 haloed voxels = GEMM rows
constexpr int PLANE_BYTES = NVOX * 16;                      // one channel block of the haloed tile
constexpr int SROW = 184;                                   // padded row length of the tap-major product buffer

// OCC = CTAs per SM.  The kernel is bound by memory latency (DRAM 46-50 %, tensor pipe 13 % in the ncu capture): with two
// resident CTAs (two input stages each, half of TMEM each) a second chain of TMA loads / drains fills the first one's bubbles.
template <bool X2, int OCC = 1> struct Cfg {
  static constexpr int AW = X2 ? 2 : 1;                     // activation words
  static constexpr int NB = X2 ? 64 : 32;                   // B rows (taps, hi | lo)
  static constexpr int GCOLS = X2 ? 64 : 32;                // TMEM columns of one M half: [main 32 | corrections (x_hi*w_lo + x_lo*w_hi) 32]
  static constexpr int PCOLS = 2 * GCOLS;                   // one plane buffer (two M halves)
  static constexpr int WBYTES = 2 * 2 * NB * 16;            // [2 k-steps][2 kcores][NB rows][8] 16-bit
  static constexpr int STAGE_BYTES = AW * 4 * PLANE_BYTES;  // 4 channel blocks per word
  static constexpr int STAGES = OCC == 2 ? 2 : 4;
  static constexpr int TCOLS = 512 / OCC;                   // TMEM columns of this CTA (two plane buffers = 2 * PCOLS <= 256)
  static constexpr int RING_PAD = 2048;                     // rows 128..255 of the last block read past the stage
  static constexpr int S_BYTES = 2 * 27 * SROW * 4;         // two product buffers [27 taps][SROW] f32 (group A fills one while B reads the other)
  static constexpr int S_OFF = WBYTES + STAGES * STAGE_BYTES + RING_PAD;
  static constexpr int BAR_OFF = S_OFF + S_BYTES;
  static constexpr int SMEM = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
  static constexpr int NTHREADS = 384;
  static_assert(WBYTES % 128 == 0 && STAGE_BYTES % 128 == 0 && S_OFF % 16 == 0 && BAR_OFF % 8 == 0, "alignment");
};

struct Params {
  const __nv_bfloat16 *w;   // packed B operand (see tc_head_weights_prepare)
  const float *res1;        // running sum of the earlier heads [B][D][H][W] f32 or nullptr
  float *y1;                // logits out [B][D][H][W] f32
  int B, D, H, W, tiles_h, tiles_w;
};

template <bool X2, bool F16, int OCC>
__global__ void __launch_bounds__(384, OCC) head_tc_kernel(const __grid_constant__ CUtensorMap xmap, const Params p)
{
  using C = Cfg<X2, OCC>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t w_addr = smem_base, stage0 = smem_base + C::WBYTES, bar0 = smem_base + C::BAR_OFF;
  float *S = reinterpret_cast<float *>(smem + C::S_OFF);
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (C::STAGES + s); };
  auto accf_bar = [&](uint32_t t) { return bar0 + 8u * (2 * C::STAGES + t); };
  auto acce_bar = [&](uint32_t t) { return bar0 + 8u * (2 * C::STAGES + 2 + t); };
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + C::BAR_OFF + (2 * C::STAGES + 4) * 8);

  ptx::pdl_launch_dependents();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int ncols = p.B * p.tiles_h * p.tiles_w, D = p.D;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&xmap);
    for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int t = 0; t < 2; ++t) { ptx::mbar_init(accf_bar(t), 1); ptx::mbar_init(acce_bar(t), 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<C::TCOLS>(ptx::smem_u32(tmem_ptr_smem));
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.w);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = threadIdx.x; i < C::WBYTES / 16; i += C::NTHREADS) dst[i] = __ldg(src + i);
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  ptx::pdl_wait();   // (programmatic dependent launch: the prologue above overlapped the previous kernel's tail)

  if (warp == 0) {
    // ================= TMA producer =================
    const bool lead = ptx::elect_one();
    uint32_t q = 0;
    for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
      const int tw = col % p.tiles_w, th = (col / p.tiles_w) % p.tiles_h, n = col / (p.tiles_w * p.tiles_h);
      for (int z = 0; z < D; ++z, ++q) {
        const uint32_t s = q % C::STAGES;
        ptx::mbar_wait(empty_bar(s), ((q / C::STAGES) & 1) ^ 1);
        if (lead) {
          ptx::mbar_arrive_expect_tx(full_bar(s), C::STAGE_BYTES);
          ptx::tma_load_4d(stage0 + s * C::STAGE_BYTES, &xmap, full_bar(s), (tw * TW - 1) * 8, th * TH - 1, z, n * (C::AW * 4));
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp converged, one elected lane issues) =================
    const bool lead = ptx::elect_one();
    const uint64_t a_desc0 = ptx::make_smem_desc(stage0, PLANE_BYTES, 128);
    const uint64_t b_desc0 = ptx::make_smem_desc(w_addr, C::NB * 16, 128);
    const uint32_t idm = ptx::make_idesc_h<F16>(128, C::NB), ids = ptx::make_idesc_h<F16>(128, 32);
    uint32_t q = 0;
    for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
      for (int z = 0; z < D; ++z, ++q) {
        const uint32_t s = q % C::STAGES, t = q & 1;
        ptx::mbar_wait(acce_bar(t), ((q >> 1) & 1) ^ 1);
        ptx::mbar_wait(full_bar(s), (q / C::STAGES) & 1);
        ptx::tc_fence_after();
        const uint64_t a0 = a_desc0 + (uint64_t)((s * C::STAGE_BYTES) >> 4);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t d = tmem_base + t * C::PCOLS + h * C::GCOLS;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t b = b_desc0 + (uint64_t)((ks * 2 * C::NB * 16) >> 4);
            // k-step ks = channel blocks 2ks, 2ks+1 of a word; rows 128.. = the second M half
            const uint64_t a_hi = a0 + (uint64_t)(((2 * ks) * PLANE_BYTES + h * 2048) >> 4);
            if (lead) ptx::umma_bf16_ss(d, a_hi, b, idm, ks);                      // x_hi * [w_hi | w_lo]
            if (X2) {
              const uint64_t a_lo = a0 + (uint64_t)(((4 + 2 * ks) * PLANE_BYTES + h * 2048) >> 4);
              if (lead) ptx::umma_bf16_ss(d + 32, a_lo, b, ids, 1u);               // x_lo * w_hi, into the correction columns
            }
          }
        }
        if (lead) { ptx::umma_commit(empty_bar(s)); ptx::umma_commit(accf_bar(t)); }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= epilogue group A: products TMEM -> shared memory (tap-major) =================
    // Named barriers: 1 + b = "S[b] full" (A arrives, B waits), 3 + b = "S[b] free" (B arrives, A waits); 256 = both groups.
    const int et = threadIdx.x - 128;                       // 0..127 = TMEM lane = GEMM row inside an M half
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t q = 0;
    for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
      for (int z = 0; z < D; ++z, ++q) {
        const uint32_t t = q & 1;
        float *Sb = S + (q & 1) * (27 * SROW);
        ptx::mbar_wait(accf_bar(t), (q >> 1) & 1);
        ptx::tc_fence_after();
        if (q & 1) asm volatile("bar.sync 4, 256;" ::: "memory"); else asm volatile("bar.sync 3, 256;" ::: "memory");   // S[b] free
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int vox = h * 128 + et;
          if (h * 128 + (warp & 3) * 32 >= NVOX) continue;   // (warp-uniform) this warp's 32 rows lie beyond the haloed tile: nothing to read
          const uint32_t ta = tmem_base + lane_addr + t * C::PCOLS + h * C::GCOLS;
          uint32_t v[32];
          ptx::tmem_ld_32x32(ta, v);
          if (X2) {
            uint32_t u[32];
            ptx::tmem_ld_32x32(ta + 32, u);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 27; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          } else {
            ptx::tmem_ld_wait();
          }
          if (vox < NVOX) {
#pragma unroll
            for (int j = 0; j < 27; ++j) Sb[j * SROW + vox] = __uint_as_float(v[j]);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acce_bar(t));       // TMEM buffer t may be overwritten
        if (q & 1) asm volatile("bar.arrive 2, 256;" ::: "memory"); else asm volatile("bar.arrive 1, 256;" ::: "memory");   // S[b] full
      }
    }
  } else if (warp >= 8) {
    // ================= epilogue group B: shifted sums shared memory -> logits =================
    const int et = threadIdx.x - 256;
    const int wl = et & 7, hl = et >> 3;                    // output position inside the 8 x 16 tile
    const int64_t HW = (int64_t)p.H * p.W;
    asm volatile("bar.arrive 3, 256;" ::: "memory");       // both product buffers start free
    asm volatile("bar.arrive 4, 256;" ::: "memory");
    uint32_t q = 0;
    for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
      const int tw = col % p.tiles_w, th = (col / p.tiles_w) % p.tiles_h, n = col / (p.tiles_w * p.tiles_h);
      const int hr = th * TH + hl, wr = tw * TW + wl;
      const bool valid = hr < p.H && wr < p.W;
      const int64_t obase = (int64_t)n * D * HW + (int64_t)hr * p.W + wr;
      float r_prev = 0.f, r_cur = 0.f;                      // running sums of output planes z-1 and z
      for (int z = 0; z < D; ++z, ++q) {
        const float *Sb = S + (q & 1) * (27 * SROW);
        // the residual (earlier heads' logits) of the plane this step completes: request it before waiting
        float res = 0.f;
        if (valid && p.res1 && z >= 1) res = __ldg(p.res1 + obase + (int64_t)(z - 1) * HW);
        if (q & 1) asm volatile("bar.sync 2, 256;" ::: "memory"); else asm volatile("bar.sync 1, 256;" ::: "memory");   // S[b] full
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;                 // kd = 0, 1, 2 contributions of input plane z
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const int off = (hl + k / 3) * SW + wl + k % 3;
          a0 += Sb[k * SROW + off];
          a1 += Sb[(9 + k) * SROW + off];
          a2 += Sb[(18 + k) * SROW + off];
        }
        if (q & 1) asm volatile("bar.arrive 4, 256;" ::: "memory"); else asm volatile("bar.arrive 3, 256;" ::: "memory");   // S[b] free
        // out[q] = sum_kd w[kd] . in[q + kd - 1]: input plane z feeds q = z+1 (kd 0), z (kd 1), z-1 (kd 2)
        if (z >= 1 && valid) p.y1[obase + (int64_t)(z - 1) * HW] = r_prev + a2 + res;
        r_prev = r_cur + a1;
        r_cur = a0;
        if (z == D - 1) {
          if (valid) p.y1[obase + (int64_t)z * HW] = r_prev + (p.res1 ? __ldg(p.res1 + obase + (int64_t)z * HW) : 0.f);
          r_prev = 0.f; r_cur = 0.f;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<C::TCOLS>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

}  // namespace headtc

void tc_head_weights_free(TcHeadWeights &w)
{
  if (w.dev) cudaFree(w.dev);
  w.dev = nullptr;
}

// w_tap: HOST [27][32][1] f32 (tap = (kd*3+kh)*3+kw).  B operand: [2 k-steps][2 kcores][NB rows][8]; row t < 27 = tap t's
// 16-bit weights, rows 27..31 zero; split precision: rows 32 + t = half(w - half(w)).
int tc_head_weights_prepare(const float *w_tap, int cin, int f16, int x2, TcHeadWeights &out, cudaStream_t s)
{
  tc_head_weights_free(out);
  out.f16 = f16; out.x2 = x2;
  if (cin != 32 || (x2 && !f16)) return IDISP_OK;  // not covered: the generic kernel runs the layer
  const int NB = x2 ? 64 : 32;
  std::vector<__nv_bfloat16> h((size_t)2 * 2 * NB * 8);
  auto cvt = [f16](float v) -> __nv_bfloat16 {
    if (!f16) return __float2bfloat16_rn(v);
    const __half hh = __float2half_rn(v);
    __nv_bfloat16 o;
    memcpy(&o, &hh, 2);
    return o;
  };
  for (int ks = 0; ks < 2; ++ks)
    for (int kc = 0; kc < 2; ++kc)
      for (int r = 0; r < NB; ++r)
        for (int e = 0; e < 8; ++e) {
          const int t = r % 32, c = ks * 16 + kc * 8 + e;
          float v = 0.f;
          if (t < 27) {
            v = w_tap[(size_t)t * cin + c];
            if (r >= 32) v = v - __half2float(__float2half_rn(v));
          }
          h[(((size_t)ks * 2 + kc) * NB + r) * 8 + e] = cvt(v);
        }
  IDISP_CUDA(cudaMalloc(&out.dev, h.size() * 2));
  IDISP_CUDA(cudaMemcpyAsync(out.dev, h.data(), h.size() * 2, cudaMemcpyHostToDevice, s));
  IDISP_CUDA(cudaStreamSynchronize(s));
  return IDISP_OK;
}

bool tc_head_supported(const TcHeadWeights &w, int D, int H, int W)
{
  static int off = -1;
  if (off < 0) { const char *e = getenv("IDISP_OLD_HEAD"); off = (e && e[0] == '1') ? 1 : 0; }
  return !off && w.dev != nullptr && D >= 1 && H >= 1 && W >= 1;
}

// x: blocked 16-bit activations [B][(x2 ? 2 : 1) * 4][D][H][W][8] (hi blocks, then lo blocks); y1 / res1: [B][D][H][W] f32
int tc_head_conv(const TcHeadWeights &w, const __nv_bfloat16 *x, int B, int D, int H, int W, const float *res1, float *y1, cudaStream_t s)
{
  if (B == 0) return IDISP_OK;
  if (!w.dev || !x || !y1) { set_error("tc_head_conv: weights not prepared / NULL pointer"); return IDISP_ERR_INVALID; }
  headtc::EncodeTiledFn enc = headtc::get_encode();
  if (!enc) { set_error("tc_head_conv: cuTensorMapEncodeTiled not available from the driver"); return IDISP_ERR_CUDA; }
  const int aw = w.x2 ? 2 : 1;
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * aw * 4};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  const cuuint32_t box[4] = {8 * headtc::SW, headtc::SH, 1, (cuuint32_t)(aw * 4)};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16 *>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tc_head_conv: cuTensorMapEncodeTiled failed (%d) for W=%d H=%d D=%d", (int)r, W, H, D); return IDISP_ERR_CUDA; }
  headtc::Params p;
  p.w = (const __nv_bfloat16 *)w.dev; p.res1 = res1; p.y1 = y1;
  p.B = B; p.D = D; p.H = H; p.W = W;
  p.tiles_h = ceil_div(H, headtc::TH); p.tiles_w = ceil_div(W, headtc::TW);
  const int ncols = B * p.tiles_h * p.tiles_w;
  int dev = 0;
  cudaGetDevice(&dev);
  static int sm_count[64];
  if (dev < 0 || dev >= 64) { set_error("tc_head_conv: device ordinal %d out of range", dev); return IDISP_ERR_INVALID; }
  if (!sm_count[dev]) cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
  static int occ2 = -1;   // two CTAs per SM (default); IDISP_HEAD_OCC1=1: the one-CTA form (four input stages)
  if (occ2 < 0) { const char *e = getenv("IDISP_HEAD_OCC1"); occ2 = (e && e[0] == '1') ? 0 : 1; }
  const int slots = sm_count[dev] * (occ2 ? 2 : 1);
  const int grid = ncols < slots ? ncols : slots;
  auto go = [&](auto kern, int smem_bytes, bool *opted) -> int {
    if (!opted[dev]) {
      IDISP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      opted[dev] = true;
    }
    IDISP_CUDA(launch_ex(kern, grid, 384, (size_t)smem_bytes, s, false, ncols <= 4 * sm_count[dev], map, p));
    return IDISP_OK;
  };
  static bool o0[64], o1[64], o2[64], o3[64], o4[64], o5[64];
  if (occ2) {
    if (w.x2) return go(headtc::head_tc_kernel<true, true, 2>, headtc::Cfg<true, 2>::SMEM, o3);
    if (w.f16) return go(headtc::head_tc_kernel<false, true, 2>, headtc::Cfg<false, 2>::SMEM, o4);
    return go(headtc::head_tc_kernel<false, false, 2>, headtc::Cfg<false, 2>::SMEM, o5);
  }
  if (w.x2) return go(headtc::head_tc_kernel<true, true, 1>, headtc::Cfg<true, 1>::SMEM, o0);
  if (w.f16) return go(headtc::head_tc_kernel<false, true, 1>, headtc::Cfg<false, 1>::SMEM, o1);
  return go(headtc::head_tc_kernel<false, false, 1>, headtc::Cfg<false, 1>::SMEM, o2);
}

}  // namespace idisp
