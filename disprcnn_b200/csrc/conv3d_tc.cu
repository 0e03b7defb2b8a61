// conv3d_tc.cu -- tcgen05 (5th-gen tensor core) implicit-GEMM 3x3x3 convolution for sm_100a.
//
// Replaces the cuDNN Conv3d+BatchNorm3d+ReLU(+add) chains of
// disprcnn/modeling/psmnet/stackhourglass.py:63-88 (applied :130-144) for the stride-1 layers
// (83 % of the 388 GFLOP/ROI at BASELINE config 2).  No im2col, no materialised patches.
//
// GEMM view.  Activations are bf16, channel-blocked-8: [N][C/8][D][H][W][8] (common.cuh); one
// voxel of one channel block is exactly the 16-byte row of a NO-SWIZZLE K-major UMMA core matrix.
//   M = 128 output voxels = an 8 (w) x 16 (h) tile of ONE depth plane,
//   K = Cin per filter tap (two K=16 MMAs per 32 channels),
//   N = 32 output channels per filter-depth tap kd.
// A operand.  For an input plane z, TMA (cp.async.bulk.tensor.4d over (8ch*W, H, D, N*Cin/8), box 80 x 18 x 1 x Cin/8,
//   zero OOB fill == conv padding) lands the haloed tile in shared memory as [Cin/8][18][10][8]:
//   rows (w) 16 B apart, 8-row groups (h) 160 B apart (SBO), K core matrices 2880 B apart (LBO).
//   The nine in-plane taps (kh,kw) are the SAME bytes read through descriptors whose start address
//   is shifted by (kh*10+kw)*16 B -- the halo is loaded once and reused 9 (in-plane) x 3 (depth) times.
// Depth streaming + kd stacking.  A CTA walks one (n, h-tile, w-tile) column through all D planes.
//   Input plane z contributes to output planes z-1, z, z+1 with kd = 2, 1, 0, so ONE MMA of
//   N = 96 = [W(kd=2) | W(kd=1) | W(kd=0)] updates three neighbouring accumulators at once.  This
//   matters on Blackwell: an SS-mode MMA with N = 32 needs 5 KB of shared-memory operands per 16
//   tensor cycles (320 B/clk vs the 128 B/clk port); N = 96 needs 7 KB per 48 cycles (146 B/clk).
//   Accumulators live in a ring of 16 TMEM slots (16 x 32 fp32 columns = all 512), slot = plane mod 16;
//   a stacked MMA is split only where the ring wraps or an accumulator is touched for the first
//   time (its accumulate flag must be 0).
// Warp roles (256 threads, 1 CTA/SM, persistent over columns): warp 0 = TMA producer, warp 1 = MMA
//   issuer (one thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld 32x32b.x32 ->
//   +bias (+residual) (ReLU) -> bf16 -> 16-byte stores, 128 B contiguous per 8-voxel row).
// Pipelines: full/empty mbarriers over the input-plane ring (TMA <-> MMA), acc_full/acc_empty over
//   the TMEM ring (MMA <-> epilogue); weights (55/110 KB) stay resident in shared memory.
// Roofline: tensor (dense bf16): 2*128*32*27*Cin FLOP per plane-tile vs 18-36 MMAs of 48 cycles.
#include "conv3d_tc.cuh"
#include "sm100_ptx.cuh"

#include <vector>

namespace idisp {

namespace tc {
constexpr int TW = 8, TH = 16;              // output tile (w x h) of one plane = 128 GEMM rows
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int PLANE_BYTES = HALO_W * HALO_H * 16;  // one channel block of one haloed plane: 2880 B
constexpr int NT = 32;                      // output channels per CTA and per kd
constexpr int NSLOT = 16;                   // TMEM accumulator ring (16 x 32 columns)
constexpr int WCHUNK = 2 * 3 * NT * 16;     // B operand of one (kh,kw,kstep): [2 kcores][96 n][8 ch] bf16 = 3072 B
constexpr int NTHREADS = 256;

template <int CIN> struct Cfg {
  static constexpr int KS = CIN / 16;           // K=16 steps per tap
  static constexpr int CBLK = CIN / 8;          // channel blocks
  static constexpr int STAGE_BYTES = CBLK * PLANE_BYTES;
  static constexpr int STAGES = CIN == 32 ? 8 : 4;
  static constexpr int WBYTES = 9 * KS * WCHUNK;
  static constexpr int BAR_OFF = WBYTES + STAGES * STAGE_BYTES;
  static constexpr int SMEM = BAR_OFF + (2 * STAGES + 2 * NSLOT) * 8 + 16 + NT * 4;
};

struct Params {
  const __nv_bfloat16 *w;         // [NH][9][KS][2][96][8] bf16
  const float *bias;              // [Cout] or nullptr
  const __nv_bfloat16 *residual;  // blocked, output shape, or nullptr
  __nv_bfloat16 *y;               // blocked [B][Cout/8][D][H][W][8]
  int B, D, H, W, Cout, relu;
  int tiles_h, tiles_w, nh;       // spatial tiling, number of 32-wide output-channel halves
  int stack;                      // 1: kd-stacked N=96 MMAs, 0: one N=32 MMA per kd (debug / A-B)
};

template <int CIN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap xmap, const Params p)
{
  using C = Cfg<CIN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t w_addr = smem_base;
  const uint32_t stage_addr0 = smem_base + C::WBYTES;
  const uint32_t bar0 = smem_base + C::BAR_OFF;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  auto accf_bar = [&](int r) { return bar0 + 8u * (2 * C::STAGES + r); };
  auto acce_bar = [&](int r) { return bar0 + 8u * (2 * C::STAGES + NSLOT + r); };
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + C::BAR_OFF + (2 * C::STAGES + 2 * NSLOT) * 8);
  float *bias_s = reinterpret_cast<float *>(smem + C::BAR_OFF + (2 * C::STAGES + 2 * NSLOT) * 8 + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's fixed output-channel half and its strided share of the (n, h-tile, w-tile) columns
  const int nh = blockIdx.x % p.nh;
  const int cta = blockIdx.x / p.nh, ncta = gridDim.x / p.nh;
  const int ncols = p.B * p.tiles_h * p.tiles_w;

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&xmap);
    for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int r = 0; r < NSLOT; ++r) { ptx::mbar_init(accf_bar(r), 1); ptx::mbar_init(acce_bar(r), 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(ptx::smem_u32(tmem_ptr_smem));
  {  // weights of this half -> shared memory (generic proxy), then make them visible to the async proxy
    const uint4 *src = reinterpret_cast<const uint4 *>(p.w) + (size_t)nh * (C::WBYTES / 16);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = threadIdx.x; i < C::WBYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
    if (threadIdx.x < NT) bias_s[threadIdx.x] = p.bias ? p.bias[nh * NT + threadIdx.x] : 0.f;
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int D = p.D;
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t q = 0;
      for (int col = cta; col < ncols; col += ncta) {
        const int tw = col % p.tiles_w, th = (col / p.tiles_w) % p.tiles_h, n = col / (p.tiles_w * p.tiles_h);
        for (int z = 0; z < D; ++z, ++q) {
          const int s = q % C::STAGES;
          ptx::mbar_wait(empty_bar(s), ((q / C::STAGES) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(full_bar(s), C::STAGE_BYTES);
          ptx::tma_load_4d(stage_addr0 + s * C::STAGE_BYTES, &xmap, full_bar(s), (tw * TW - 1) * 8, th * TH - 1, z, n * C::CBLK);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (single thread) =================
    if (lane == 0) {
      uint32_t q = 0, g0 = 0;
      for (int col = cta; col < ncols; col += ncta, g0 += D) {
        for (int z = 0; z < D; ++z, ++q) {
          // accumulators touched for the first time by this input plane must have been drained
          if (z == 0) {
            ptx::mbar_wait(acce_bar(g0 % NSLOT), ((g0 / NSLOT) & 1) ^ 1);
            if (D > 1) ptx::mbar_wait(acce_bar((g0 + 1) % NSLOT), (((g0 + 1) / NSLOT) & 1) ^ 1);
          } else if (z + 1 < D) {
            const uint32_t g = g0 + z + 1;
            ptx::mbar_wait(acce_bar(g % NSLOT), ((g / NSLOT) & 1) ^ 1);
          }
          const int s = q % C::STAGES;
          ptx::mbar_wait(full_bar(s), (q / C::STAGES) & 1);
          ptx::tc_fence_after();
          const uint32_t a_stage = stage_addr0 + s * C::STAGE_BYTES;
          const int jlo = z == 0 ? 1 : 0, jhi = z == D - 1 ? 1 : 2;  // j <-> output plane z-1+j, kd = 2-j
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int kh = tap / 3, kw = tap - 3 * kh;
#pragma unroll
            for (int ks = 0; ks < C::KS; ++ks) {
              const uint64_t a_desc = ptx::make_smem_desc(a_stage + (kh * HALO_W + kw) * 16 + ks * 2 * PLANE_BYTES, PLANE_BYTES, HALO_W * 16);
              const uint32_t b_chunk = w_addr + (tap * C::KS + ks) * WCHUNK;
              const bool first = (tap == 0 && ks == 0);
              int j = jlo;
              while (j <= jhi) {
                const uint32_t slot = (g0 + z - 1 + j) % NSLOT;
                const bool fresh = first && (z == 0 || j == 2);
                int len = 1;
                if (p.stack) {
                  while (j + len <= jhi && slot + len < NSLOT && (first && (z == 0 || j + len == 2)) == fresh) ++len;
                }
                const uint64_t b_desc = ptx::make_smem_desc(b_chunk + j * NT * 16, 3 * NT * 16, 128);
                ptx::umma_bf16_ss(tmem_base + slot * NT, a_desc, b_desc, ptx::make_idesc_bf16(128, NT * len), fresh ? 0u : 1u);
                j += len;
              }
            }
          }
          ptx::umma_commit(empty_bar(s));  // smem stage reusable once these MMAs have read it
          if (z >= 1) ptx::umma_commit(accf_bar((g0 + z - 1) % NSLOT));  // plane z-1 complete
          if (z == D - 1) ptx::umma_commit(accf_bar((g0 + z) % NSLOT));   // last plane complete
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (4 warps = 128 TMEM lanes) =================
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;       // GEMM row = TMEM lane
    const int wl = m & 7, hl = m >> 3;       // voxel inside the 8 x 16 tile
    const int64_t V = (int64_t)D * p.H * p.W;
    const int cblk_out = p.Cout / 8;
    uint32_t g0 = 0;
    for (int col = cta; col < ncols; col += ncta, g0 += D) {
      const int tw = col % p.tiles_w, th = (col / p.tiles_w) % p.tiles_h, n = col / (p.tiles_w * p.tiles_h);
      const int h = th * TH + hl, w = tw * TW + wl;
      const bool valid = h < p.H && w < p.W;
      for (int z = 0; z < D; ++z) {
        const uint32_t g = g0 + z, r = g % NSLOT;
        ptx::mbar_wait(accf_bar(r), (g / NSLOT) & 1);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + r * NT, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acce_bar(r));  // accumulator slot may be overwritten
        if (valid) {
          const int64_t pos = ((int64_t)z * p.H + h) * p.W + w;
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            const int64_t o = (((int64_t)n * cblk_out + nh * 4 + cb) * V + pos) * 8;
            F8 r8;
#pragma unroll
            for (int c = 0; c < 8; ++c) r8.v[c] = __uint_as_float(v[cb * 8 + c]) + bias_s[cb * 8 + c];
            if (p.residual) {
              const F8 q8 = load8<__nv_bfloat16>(p.residual + o);
#pragma unroll
              for (int c = 0; c < 8; ++c) r8.v[c] += q8.v[c];
            }
            if (p.relu) {
#pragma unroll
              for (int c = 0; c < 8; ++c) r8.v[c] = fmaxf(r8.v[c], 0.f);
            }
            store8<__nv_bfloat16>(p.y + o, r8);
          }
        }
      }
    }
  }

  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

static int debug_nostack()
{
  static int v = -1;
  if (v < 0) { const char *e = getenv("IDISP_TC_NOSTACK"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

}  // namespace tc

int tc_weights_prepare(const float *w_tap, int kind, int cin, int cout, TcWeights &out, cudaStream_t s)
{
  tc_weights_free(out);
  out.kind = kind; out.cin = cin; out.cout = cout;
  if (!(kind == IDISP_CONV_S1 && (cin == 32 || cin == 64) && (cout == 32 || cout == 64))) return IDISP_OK;  // SIMT layer
  const int KS = cin / 16, NH = cout / 32;
  std::vector<__nv_bfloat16> h((size_t)NH * 9 * KS * 2 * 96 * 8);
  for (int nh = 0; nh < NH; ++nh)
    for (int t2 = 0; t2 < 9; ++t2)
      for (int ks = 0; ks < KS; ++ks)
        for (int kc = 0; kc < 2; ++kc)
          for (int n = 0; n < 96; ++n)
            for (int e = 0; e < 8; ++e) {
              const int j = n / 32, co = n % 32, kd = 2 - j, kh = t2 / 3, kw = t2 % 3;
              const int ci = ks * 16 + kc * 8 + e;
              const float v = w_tap[((size_t)((kd * 3 + kh) * 3 + kw) * cin + ci) * cout + nh * 32 + co];
              h[(((((size_t)nh * 9 + t2) * KS + ks) * 2 + kc) * 96 + n) * 8 + e] = __float2bfloat16_rn(v);
            }
  out.bytes = h.size() * sizeof(__nv_bfloat16);
  IDISP_CUDA(cudaMalloc(&out.dev, out.bytes));
  IDISP_CUDA(cudaMemcpyAsync(out.dev, h.data(), out.bytes, cudaMemcpyHostToDevice, s));
  IDISP_CUDA(cudaStreamSynchronize(s));
  return IDISP_OK;
}

void tc_weights_free(TcWeights &w)
{
  if (w.dev) cudaFree(w.dev);
  w.dev = nullptr; w.bytes = 0;
}

bool tc_supported(int kind, int cin, int cout, int D, int H, int W)
{
  static int disabled = -1;
  if (disabled < 0) { const char *e = getenv("IDISP_TC_DISABLE"); disabled = (e && e[0] == '1') ? 1 : 0; }
  if (disabled) return false;
  return kind == IDISP_CONV_S1 && (cin == 32 || cin == 64) && (cout == 32 || cout == 64) && D >= 1 && H >= 1 && W >= 1;
}

template <int CIN>
static int tc_launch(const TcWeights &w, const __nv_bfloat16 *x, int B, int D, int H, int W, int Cout, const float *bias,
                     const __nv_bfloat16 *residual, int relu, __nv_bfloat16 *y, cudaStream_t s)
{
  using C = tc::Cfg<CIN>;
  tc::EncodeTiledFn enc = tc::get_encode();
  if (!enc) { set_error("tc_conv3d: cuTensorMapEncodeTiled not available from the driver"); return IDISP_ERR_CUDA; }
  CUtensorMap map;
  // (8 ch, W) are contiguous in the blocked layout -> ONE tensor dimension of 8*W elements, so a box row is
  // 10 voxels x 16 B = 160 contiguous bytes (a 16-byte inner box made TMA issue one request per voxel and
  // capped the kernel at ~10 cycles per 16 B; measured in profiles/r01_notes.md)
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * C::CBLK};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  const cuuint32_t box[4] = {8 * tc::HALO_W, tc::HALO_H, 1, (cuuint32_t)C::CBLK};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16 *>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tc_conv3d: cuTensorMapEncodeTiled failed (%d) for dims W=%d H=%d D=%d", (int)r, W, H, D); return IDISP_ERR_CUDA; }
  tc::Params p;
  p.w = (const __nv_bfloat16 *)w.dev; p.bias = bias; p.residual = residual; p.y = y;
  p.B = B; p.D = D; p.H = H; p.W = W; p.Cout = Cout; p.relu = relu;
  p.tiles_h = ceil_div(H, tc::TH); p.tiles_w = ceil_div(W, tc::TW); p.nh = Cout / 32;
  p.stack = tc::debug_nostack() ? 0 : 1;
  const int ncols = B * p.tiles_h * p.tiles_w;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_half = sms / p.nh;
  if (per_half > ncols) per_half = ncols;
  const int grid = per_half * p.nh;
  auto kern = tc::conv3d_tc_kernel<CIN>;
  IDISP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  kern<<<grid, tc::NTHREADS, C::SMEM, s>>>(map, p);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

int tc_conv3d(const TcWeights &w, const __nv_bfloat16 *x, int B, int Cin, int D, int H, int W, int Cout, int kind,
              const float *bias, const __nv_bfloat16 *residual, int relu, __nv_bfloat16 *y, cudaStream_t s)
{
  if (!tc_supported(kind, Cin, Cout, D, H, W) || !w.dev || w.cin != Cin || w.cout != Cout) {
    set_error("tc_conv3d: layer (kind=%d, %d->%d) not prepared for the tensor-core path", kind, Cin, Cout);
    return IDISP_ERR_INVALID;
  }
  if (B == 0) return IDISP_OK;
  return Cin == 32 ? tc_launch<32>(w, x, B, D, H, W, Cout, bias, residual, relu, y, s)
                   : tc_launch<64>(w, x, B, D, H, W, Cout, bias, residual, relu, y, s);
}

}  // namespace idisp
