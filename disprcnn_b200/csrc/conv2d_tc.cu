// conv2d_tc.cu -- tcgen05 implicit-GEMM 3x3 stride-1 2-D convolution (dilation 1 or 2) in split precision, for the feature
// extractor of iDispNet (disprcnn/modeling/psmnet/submodule.py:60-139): 53 of its 62 convolutions and 99 % of its 22.2 GFLOP per
// 224 x 224 crop (firstconv.2/.4, every BasicBlock conv of layer1-4 except the stride-2 one, lastconv.0) -- SURVEY.md 8(f) row 1.
//
// Same operand scheme as the 3-D stack (conv3d_tc.cu): activations are two IEEE-half words per value in the channel-blocked-8
// layout [N][hi blocks C/8 | lo blocks C/8][H][W][8]; a product is x_hi*w_hi + x_hi*w_lo + x_lo*w_hi accumulated in fp32.
// GEMM view: M = 128 = an 8 (w) x 16 (h) output tile, K = 32 input channels per chunk and tap (two K = 16 MMAs), N = 32 output
// channels per CTA.  One TMA box pair per (tile, 32-channel chunk) lands the haloed tile [word][4 blocks][16+2d][8+2d][8] in
// shared memory; the nine taps are the same bytes read through descriptors shifted by (kh*d*row + kw*d) * 16 B.
// Weight-stationary: a CTA keeps ALL taps of its 32 output channels for up to 128 input channels resident in shared memory
// (36 KB per 32-channel chunk: [9 taps][2 k-steps][2 kcores][lo 32 rows | hi 32 rows][8]) and streams tiles past them, so B
// is read from L2 once per CTA instead of once per tile.  MMAs: x_hi * [w_hi | w_lo] is ONE N = 64 MMA (main | corrections), x_lo *
// w_hi an N = 32 MMA into the corrections.  tcgen05.mma adds into the fp32 accumulator by TRUNCATION: a bias towards
// zero that grows with the number of adds a column collects, and through 53 chained layers it showed (18 adds per main column: features
// 5e-6 relative too small, disparity 1.1e-3 px).  So every 32-channel chunk gets a bank of THREE column blocks [M0 | C | M1]: the main
// products of k-step 0 go to M0 and those of k-step 1 to M1 (9 adds each), both k-steps share the correction block between them
// (B rows are packed [hi | lo] for k-step 0 and [lo | hi] for k-step 1, so each merged MMA's D stays contiguous); the epilogue sums
// blocks and banks in fp32 round-to-nearest.  Cin = 320 (lastconv.0) runs as three launches chained through an fp32 partial.
// Epilogue: + bias (+ residual hi + lo) (ReLU) -> hi = half(v), lo = half(v - hi) -> 16-byte stores, channel-block offsets on
// input / output / residual so that `raw`, `skip` and the SPP branches live inside the 320-channel concat tensor.
// Roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue; accumulators are
// double-buffered where two sets of banks fit the 512 TMEM columns (up to 64 input channels per launch).  Roofline: tensor (3 half-precision MMAs per algorithmic product) / shared-memory port.
#include "conv2d_tc.cuh"
#include "sm100_ptx.cuh"

#include <cstring>
#include <vector>

namespace idisp {
namespace c2d {

constexpr int TW = 8, TH = 16;

// DIL = 0 selects the 1x1 convolution (one tap, no halo); DIL = 1 / 2 the 3x3 convolution with that dilation (padding = dilation)
template <int DIL> struct Geo {
  static constexpr int TAPS = DIL == 0 ? 1 : 9;
  static constexpr int CHUNK_W = TAPS * 2 * 2 * 64 * 16;   // weights of one 32-input-channel chunk for 32 output channels
  static constexpr int SUB_W = TW + 2 * DIL, SUB_H = TH + 2 * DIL;
  static constexpr int PLANE = SUB_W * SUB_H * 16;      // one channel block of the haloed tile
  static constexpr int STAGE = 8 * PLANE;               // [hi: 4 blocks | lo: 4 blocks]
  static_assert((4 * PLANE) % 128 == 0, "TMA destinations (the hi and the lo half of a stage) must stay 128 B aligned");
};

struct Params {
  const __nv_bfloat16 *w;
  const float *bias;
  const __nv_bfloat16 *res;
  __nv_bfloat16 *y;
  const float *part_in;
  float *part_out;
  int *range_flag;
  int B, H, W, relu, nchunks, stages, nbuf;
  int in_blocks, in_blk0, in_lo;
  int out_blocks, out_blk0, out_lo;
  int res_blocks, res_blk0, res_lo;
  int part_cblk;     // channel blocks per sample of the fp32 partial (= Cout / 8)
  int w_slice_bytes; // distance between the weight sets of consecutive output-channel slices (all chunks of the layer)
  int tiles_h, tiles_w, nslices;
};

template <int DIL>
__global__ void __launch_bounds__(256, 1) conv2d_tc_kernel(const __grid_constant__ CUtensorMap xmap, const Params p)
{
  using G = Geo<DIL>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t wbytes = (uint32_t)p.nchunks * G::CHUNK_W;
  const uint32_t w_addr = smem_base, stage0 = smem_base + wbytes, bar0 = stage0 + (uint32_t)p.stages * G::STAGE;
  const uint32_t S = (uint32_t)p.stages;
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (S + s); };
  auto accf_bar = [&](uint32_t t) { return bar0 + 8u * (2 * S + t); };
  auto acce_bar = [&](uint32_t t) { return bar0 + 8u * (2 * S + 2 + t); };
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + wbytes + (size_t)p.stages * G::STAGE + (2 * S + 4) * 8);

  ptx::pdl_launch_dependents();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int nh = blockIdx.x % p.nslices;                 // this CTA's 32-wide output-channel slice (its weights stay resident)
  const int cta = blockIdx.x / p.nslices, ncta = gridDim.x / p.nslices;
  const int ntiles = p.B * p.tiles_h * p.tiles_w;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&xmap);
    for (uint32_t s = 0; s < S; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int t = 0; t < 2; ++t) { ptx::mbar_init(accf_bar(t), 1); ptx::mbar_init(acce_bar(t), 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(ptx::smem_u32(tmem_ptr_smem));
  // weights (36 KB per 32-channel chunk) -> shared memory as bulk TMA copies on their own mbarrier; only the MMA warp waits for it
  // (the copy loop by all threads cost ~10 us of dependent L2 round trips per launch -- most of a launch at the live shape)
  const uint32_t wbar = bar0 + 8u * (2 * S + 4) + 8u;   // (second half of the 16-byte slot that holds the TMEM pointer)
  if (warp == 0 && lane == 0) {
    ptx::mbar_init(wbar, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(wbar, wbytes);
    const char *src = reinterpret_cast<const char *>(p.w) + (size_t)nh * p.w_slice_bytes;
    for (uint32_t off = 0; off < wbytes; off += 16384u) ptx::bulk_g2s(w_addr + off, src + off, wbytes - off < 16384u ? wbytes - off : 16384u, wbar);
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
    for (int tile = cta; tile < ntiles; tile += ncta) {
      const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
      for (int c = 0; c < p.nchunks; ++c, ++q) {
        const uint32_t s = q % S;
        ptx::mbar_wait(empty_bar(s), ((q / S) & 1) ^ 1);
        if (lead) {
          ptx::mbar_arrive_expect_tx(full_bar(s), G::STAGE);
          const int blk = n * p.in_blocks + p.in_blk0 + c * 4;
          ptx::tma_load_3d(stage0 + s * G::STAGE, &xmap, full_bar(s), (tw * TW - DIL) * 8, th * TH - DIL, blk);
          ptx::tma_load_3d(stage0 + s * G::STAGE + 4 * G::PLANE, &xmap, full_bar(s), (tw * TW - DIL) * 8, th * TH - DIL, blk + p.in_lo);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp converged, one elected lane issues) =================
    const bool lead = ptx::elect_one();
    const uint64_t a_desc0 = ptx::make_smem_desc(stage0, G::PLANE, G::SUB_W * 16);
    const uint64_t b_desc0 = ptx::make_smem_desc(w_addr, 64 * 16, 128);
    const uint32_t id64 = ptx::make_idesc_h<true>(128, 64), id32 = ptx::make_idesc_h<true>(128, 32);
    const uint32_t nbuf = (uint32_t)p.nbuf, bufcols = (uint32_t)p.nchunks * 96;
    uint32_t q = 0, it = 0;
    ptx::mbar_wait(wbar, 0);   // the weights have landed
    for (int tile = cta; tile < ntiles; tile += ncta, ++it) {
      const uint32_t t = it % nbuf;
      ptx::mbar_wait(acce_bar(t), ((it / nbuf) & 1) ^ 1);
      for (int c = 0; c < p.nchunks; ++c, ++q) {
        const uint32_t s = q % S;
        ptx::mbar_wait(full_bar(s), (q / S) & 1);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + t * bufcols + c * 96;               // this chunk's bank [M0 | C | M1]
        const uint64_t a0 = a_desc0 + (uint64_t)((s * G::STAGE) >> 4);
        const uint64_t b0 = b_desc0 + (uint64_t)(((uint32_t)c * G::CHUNK_W) >> 4);
#pragma unroll
        for (int tap = 0; tap < G::TAPS; ++tap) {
          const uint32_t shift = ((tap / 3) * DIL * G::SUB_W + (tap % 3) * DIL) * 16;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t a_hi = a0 + (uint64_t)(((2 * ks) * G::PLANE + shift) >> 4);
            const uint64_t a_lo = a0 + (uint64_t)(((4 + 2 * ks) * G::PLANE + shift) >> 4);
            const uint64_t b = b0 + (uint64_t)(((tap * 2 + ks) * 2048) >> 4);
            if (lead) {
              if (ks == 0) {   // B rows [hi | lo] -> D = [M0 | C]; the very first MMA of the tile overwrites both blocks
                ptx::umma_bf16_ss(d, a_hi, b, id64, tap ? 1u : 0u);
                ptx::umma_bf16_ss(d + 32, a_lo, b, id32, 1u);                                     // x_lo * w_hi (rows 0..31) -> C
              } else {         // B rows [lo | hi] -> D = [C | M1]; M1's first MMA must overwrite it, C must be kept: split once
                if (tap == 0) {
                  ptx::umma_bf16_ss(d + 32, a_hi, b, id32, 1u);                                   // x_hi * w_lo -> C
                  ptx::umma_bf16_ss(d + 64, a_hi, b + (uint64_t)((32 * 16) >> 4), id32, 0u);      // x_hi * w_hi -> M1 (overwrite)
                } else {
                  ptx::umma_bf16_ss(d + 32, a_hi, b, id64, 1u);
                }
                ptx::umma_bf16_ss(d + 32, a_lo, b + (uint64_t)((32 * 16) >> 4), id32, 1u);        // x_lo * w_hi (rows 32..63) -> C
              }
            }
          }
        }
        if (lead) ptx::umma_commit(empty_bar(s));
      }
      if (lead) ptx::umma_commit(accf_bar(t));
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int m = (warp & 3) * 32 + lane;                  // GEMM row = TMEM lane
    const int wl = m & 7, hl = m >> 3;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const int64_t HW = (int64_t)p.H * p.W;
    bool bad = false;
    const uint32_t nbuf = (uint32_t)p.nbuf, bufcols = (uint32_t)p.nchunks * 96;
    uint32_t it = 0;
    for (int tile = cta; tile < ntiles; tile += ncta, ++it) {
      const uint32_t t = it % nbuf;
      const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
      const int hr = th * TH + hl, wr = tw * TW + wl;
      const bool valid = hr < p.H && wr < p.W;
      const int64_t pos = (int64_t)hr * p.W + wr;
      // operands that do not depend on the accumulator: requested before the wait
      uint4 rh[4], rl[4];
      if (valid && p.res) {
        const int64_t rb = ((int64_t)n * p.res_blocks + p.res_blk0 + nh * 4) * HW + pos;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          rh[cb] = __ldg(reinterpret_cast<const uint4 *>(p.res + (rb + (int64_t)cb * HW) * 8));
          rl[cb] = __ldg(reinterpret_cast<const uint4 *>(p.res + (rb + (int64_t)(cb + p.res_lo) * HW) * 8));
        }
      }
      ptx::mbar_wait(accf_bar(t), (it / nbuf) & 1);
      ptx::tc_fence_after();
      float acc[32];
      for (int c = 0; c < p.nchunks; ++c) {                 // (warp-uniform trip count: tcgen05.ld is warp-collective)
        const uint32_t ta = tmem_base + lane_addr + t * bufcols + c * 96;   // [M0 | C | M1]
        uint32_t m0[32], cc[32];
        ptx::tmem_ld_32x32(ta, m0);
        ptx::tmem_ld_32x32(ta + 32, cc);
        ptx::tmem_ld_wait();
        float bank[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) bank[i] = __uint_as_float(m0[i]) + __uint_as_float(cc[i]);
        ptx::tmem_ld_32x32(ta + 64, m0);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) bank[i] += __uint_as_float(m0[i]);
        if (c == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = bank[i];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] += bank[i];
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acce_bar(t));
      if (!valid) continue;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float a[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = acc[cb * 8 + c];
        const int64_t po = (((int64_t)n * p.part_cblk + nh * 4 + cb) * HW + pos) * 8;
        if (p.part_in) {
          const float4 p0 = __ldg(reinterpret_cast<const float4 *>(p.part_in + po)), p1 = __ldg(reinterpret_cast<const float4 *>(p.part_in + po) + 1);
          a[0] += p0.x; a[1] += p0.y; a[2] += p0.z; a[3] += p0.w; a[4] += p1.x; a[5] += p1.y; a[6] += p1.z; a[7] += p1.w;
        }
        if (p.part_out) {
          float4 *o = reinterpret_cast<float4 *>(p.part_out + po);
          o[0] = make_float4(a[0], a[1], a[2], a[3]);
          o[1] = make_float4(a[4], a[5], a[6], a[7]);
          continue;
        }
        if (p.bias) {
#pragma unroll
          for (int c = 0; c < 8; ++c) a[c] += __ldg(p.bias + nh * 32 + cb * 8 + c);
        }
        if (p.res) {
          const F8 r0 = unpack8h<true>(rh[cb]), r1 = unpack8h<true>(rl[cb]);
#pragma unroll
          for (int c = 0; c < 8; ++c) a[c] += r0.v[c];
#pragma unroll
          for (int c = 0; c < 8; ++c) a[c] += r1.v[c];
        }
        if (p.relu) {
#pragma unroll
          for (int c = 0; c < 8; ++c) a[c] = fmaxf(a[c], 0.f);
        }
        F8 f;
#pragma unroll
        for (int c = 0; c < 8; ++c) f.v[c] = a[c];
        const uint4 hi = pack8h<true>(f);
        {  // a half whose exponent field is all ones: the value left the IEEE-half range (or was NaN)
          const uint32_t mm = ((hi.x & 0x7fff7fffu) + 0x04000400u) | ((hi.y & 0x7fff7fffu) + 0x04000400u) |
                              ((hi.z & 0x7fff7fffu) + 0x04000400u) | ((hi.w & 0x7fff7fffu) + 0x04000400u);
          bad |= (mm & 0x80008000u) != 0;
        }
        const F8 h = unpack8h<true>(hi);
#pragma unroll
        for (int c = 0; c < 8; ++c) f.v[c] -= h.v[c];
        const uint4 lo = pack8h<true>(f);
        const int64_t ob = ((int64_t)n * p.out_blocks + p.out_blk0 + nh * 4 + cb) * HW + pos;
        *reinterpret_cast<uint4 *>(p.y + ob * 8) = hi;
        *reinterpret_cast<uint4 *>(p.y + (ob + (int64_t)p.out_lo * HW) * 8) = lo;
      }
    }
    if (bad && p.range_flag) *p.range_flag = 1;
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

// NCHW f32 [B][C][HW] (batch stride src_bs) -> channel blocks [blk0, blk0 + C/8) (hi) and + lo_off (lo) of a blocked tensor
__global__ void nchw_to_x2_kernel(const float *__restrict__ src, long long src_bs, uint4 *__restrict__ dst, int blocks, int blk0, int lo_off, int C,
                                  long long HW, int *range_flag)
{
  const int n = blockIdx.y, cblks = C / 8;
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)cblks * HW; i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i % HW;
    const int cb = (int)(i / HW);
    F8 f;
#pragma unroll
    for (int c = 0; c < 8; ++c) f.v[c] = __ldg(src + (long long)n * src_bs + (long long)(cb * 8 + c) * HW + pos);
    const uint4 hi = pack8h<true>(f);
    const uint32_t mm = ((hi.x & 0x7fff7fffu) + 0x04000400u) | ((hi.y & 0x7fff7fffu) + 0x04000400u) | ((hi.z & 0x7fff7fffu) + 0x04000400u) |
                        ((hi.w & 0x7fff7fffu) + 0x04000400u);
    bad |= (mm & 0x80008000u) != 0;
    const F8 h = unpack8h<true>(hi);
#pragma unroll
    for (int c = 0; c < 8; ++c) f.v[c] -= h.v[c];
    const long long o = ((long long)n * blocks + blk0 + cb) * HW + pos;
    dst[o] = hi;
    dst[o + (long long)lo_off * HW] = pack8h<true>(f);
  }
  if (bad && range_flag) *range_flag = 1;
}

__global__ void x2_to_nchw_kernel(const uint4 *__restrict__ src, int blocks, int blk0, int lo_off, float *__restrict__ dst, long long dst_bs, int C,
                                  long long HW)
{
  const int n = blockIdx.y, cblks = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)cblks * HW; i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i % HW;
    const int cb = (int)(i / HW);
    const long long o = ((long long)n * blocks + blk0 + cb) * HW + pos;
    const F8 h = unpack8h<true>(__ldg(src + o)), l = unpack8h<true>(__ldg(src + o + (long long)lo_off * HW));
#pragma unroll
    for (int c = 0; c < 8; ++c) dst[(long long)n * dst_bs + (long long)(cb * 8 + c) * HW + pos] = h.v[c] + l.v[c];
  }
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

}  // namespace c2d

void c2d_weights_free(C2dWeights &w)
{
  if (w.dev) cudaFree(w.dev);
  w.dev = nullptr;
}

// w: HOST [Cin][9][Cout] f32 (tap = kh*3+kw, BN scale folded in) -> [Cout/32 slices][Cin/32 chunks][9 taps][2 k-steps][2 kcores][64 rows][8]:
// k-step 0: rows 0..31 = half(w) (hi), rows 32..63 = half(w - half(w)) (lo) of output channel slice*32 + row % 32; k-step 1: [lo | hi]
int c2d_weights_prepare(const float *w, int cin, int cout, int taps, C2dWeights &out, cudaStream_t s)
{
  c2d_weights_free(out);
  out.cin = cin; out.cout = cout; out.taps = taps;
  if (taps != 1 && taps != 9) { set_error("c2d_weights_prepare: 1 or 9 taps expected, got %d", taps); return IDISP_ERR_INVALID; }
  if (cin % 32 || cout % 32) { set_error("c2d_weights_prepare: Cin=%d / Cout=%d must be multiples of 32", cin, cout); return IDISP_ERR_INVALID; }
  const int nsl = cout / 32, nch = cin / 32;
  std::vector<__half> h((size_t)nsl * nch * taps * 2 * 2 * 64 * 8);
  for (int sl = 0; sl < nsl; ++sl)
    for (int ch = 0; ch < nch; ++ch)
      for (int tap = 0; tap < taps; ++tap)
        for (int ks = 0; ks < 2; ++ks)
          for (int kc = 0; kc < 2; ++kc)
            for (int row = 0; row < 64; ++row)
              for (int e = 0; e < 8; ++e) {
                const int ci = ch * 32 + ks * 16 + kc * 8 + e, co = sl * 32 + row % 32;
                const float v = w[((size_t)ci * taps + tap) * cout + co];
                const __half hi = __float2half_rn(v);
                const bool is_hi = ks == 0 ? row < 32 : row >= 32;
                const __half val = is_hi ? hi : __float2half_rn(v - __half2float(hi));
                h[((((((size_t)sl * nch + ch) * taps + tap) * 2 + ks) * 2 + kc) * 64 + row) * 8 + e] = val;
              }
  IDISP_CUDA(cudaMalloc(&out.dev, h.size() * 2));
  IDISP_CUDA(cudaMemcpyAsync(out.dev, h.data(), h.size() * 2, cudaMemcpyHostToDevice, s));
  IDISP_CUDA(cudaStreamSynchronize(s));
  return IDISP_OK;
}

int c2d_nchw_to_x2(const float *src, long long src_bs, __nv_bfloat16 *dst, int blocks, int blk0, int lo_off, int B, int C, long long HW, int *range_flag,
                   cudaStream_t s)
{
  if (B == 0) return IDISP_OK;
  if (C % 8) { set_error("c2d_nchw_to_x2: C=%d must be a multiple of 8", C); return IDISP_ERR_INVALID; }
  const long long total = (long long)(C / 8) * HW;
  const int gx = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  c2d::nchw_to_x2_kernel<<<dim3(gx, B), 256, 0, s>>>(src, src_bs, (uint4 *)dst, blocks, blk0, lo_off, C, HW, range_flag);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

int c2d_x2_to_nchw(const __nv_bfloat16 *src, int blocks, int blk0, int lo_off, float *dst, long long dst_bs, int B, int C, long long HW, cudaStream_t s)
{
  if (B == 0) return IDISP_OK;
  if (C % 8) { set_error("c2d_x2_to_nchw: C=%d must be a multiple of 8", C); return IDISP_ERR_INVALID; }
  const long long total = (long long)(C / 8) * HW;
  const int gx = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  c2d::x2_to_nchw_kernel<<<dim3(gx, B), 256, 0, s>>>((const uint4 *)src, blocks, blk0, lo_off, dst, dst_bs, C, HW);
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

// One 3x3 stride-1 conv (padding = dilation) on blocked split-precision tensors.  `chunk0` / `nchunks`: the 32-channel input
// chunks this launch covers (a layer with Cin > 128 is several launches chained through part_in / part_out).
int c2d_conv(const C2dWeights &w, int dil, const C2dTensor &x, int chunk0, int nchunks, int B, int H, int W, const float *bias, const C2dTensor *res,
             int relu, const C2dTensor &y, const float *part_in, float *part_out, int *range_flag, cudaStream_t s)
{
  if (B == 0) return IDISP_OK;
  if (!w.dev || dil < 0 || dil > 2 || w.taps != (dil == 0 ? 1 : 9) || nchunks < 1 || nchunks > 4 || (chunk0 + nchunks) * 32 > w.cin) {
    set_error("c2d_conv: bad arguments (dil=%d chunks [%d,%d) of Cin=%d)", dil, chunk0, chunk0 + nchunks, w.cin);
    return IDISP_ERR_INVALID;
  }
  c2d::EncodeTiledFn enc = c2d::get_encode();
  if (!enc) { set_error("c2d_conv: cuTensorMapEncodeTiled not available from the driver"); return IDISP_ERR_CUDA; }
  const int sub_w = c2d::TW + 2 * dil, sub_h = c2d::TH + 2 * dil;
  const int stage = 8 * sub_w * sub_h * 16;
  const int chunk_w = w.taps * 2 * 2 * 64 * 16;
  const int wbytes = nchunks * chunk_w;
  int stages = (232448 - 2048 - wbytes) / stage;
  if (stages > 6) stages = 6;
  if (stages < 2) { set_error("c2d_conv: shared memory does not hold two input stages next to %d weight chunks", nchunks); return IDISP_ERR_INVALID; }
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)B * x.blocks};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  const cuuint32_t box[3] = {(cuuint32_t)(8 * sub_w), (cuuint32_t)sub_h, 4};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16 *>(x.p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("c2d_conv: cuTensorMapEncodeTiled failed (%d) for W=%d H=%d blocks=%d", (int)r, W, H, B * x.blocks); return IDISP_ERR_CUDA; }
  c2d::Params p;
  const int nch_total = w.cin / 32;
  p.w = (const __nv_bfloat16 *)w.dev + (size_t)chunk0 * (chunk_w / 2);   // [slice][chunk][...]: this launch's chunks of slice 0
  p.w_slice_bytes = nch_total * chunk_w;
  p.bias = bias; p.res = res ? res->p : nullptr; p.y = y.p; p.part_in = part_in; p.part_out = part_out; p.range_flag = range_flag;
  p.B = B; p.H = H; p.W = W; p.relu = relu; p.nchunks = nchunks; p.stages = stages;
  p.nbuf = 2 * nchunks * 96 <= 512 ? 2 : 1;   // two sets of accumulator banks where they fit the 512 TMEM columns
  p.in_blocks = x.blocks; p.in_blk0 = x.blk0 + chunk0 * 4; p.in_lo = x.lo;
  p.out_blocks = y.blocks; p.out_blk0 = y.blk0; p.out_lo = y.lo;
  p.res_blocks = res ? res->blocks : 0; p.res_blk0 = res ? res->blk0 : 0; p.res_lo = res ? res->lo : 0;
  p.part_cblk = w.cout / 8;
  p.tiles_h = ceil_div(H, c2d::TH); p.tiles_w = ceil_div(W, c2d::TW); p.nslices = w.cout / 32;
  const int ntiles = B * p.tiles_h * p.tiles_w;
  int dev = 0;
  cudaGetDevice(&dev);
  static int sm_count[64];
  if (dev < 0 || dev >= 64) { set_error("c2d_conv: device ordinal %d out of range", dev); return IDISP_ERR_INVALID; }
  if (!sm_count[dev]) cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
  int per_slice = sm_count[dev] / p.nslices;
  if (per_slice < 1) per_slice = 1;
  if (per_slice > ntiles) per_slice = ntiles;
  const int grid = per_slice * p.nslices;
  const int smem = wbytes + stages * stage + (2 * stages + 4) * 8 + 16;
  // short launches only (see launch_ex): rounds x chunks <= 64 is below ~100 us -- every extractor launch of the live call (R <= 15)
  const bool pdl = (long long)ntiles * p.nslices * p.nchunks <= 64ll * sm_count[dev];
  static bool o0[64], o1[64], o2[64];
  if (dil == 0) {
    if (!o0[dev]) { IDISP_CUDA(cudaFuncSetAttribute(c2d::conv2d_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)); o0[dev] = true; }
    IDISP_CUDA(launch_ex(c2d::conv2d_tc_kernel<0>, grid, 256, (size_t)smem, s, false, pdl, map, p));
  } else if (dil == 1) {
    if (!o1[dev]) { IDISP_CUDA(cudaFuncSetAttribute(c2d::conv2d_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)); o1[dev] = true; }
    IDISP_CUDA(launch_ex(c2d::conv2d_tc_kernel<1>, grid, 256, (size_t)smem, s, false, pdl, map, p));
  } else {
    if (!o2[dev]) { IDISP_CUDA(cudaFuncSetAttribute(c2d::conv2d_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)); o2[dev] = true; }
    IDISP_CUDA(launch_ex(c2d::conv2d_tc_kernel<2>, grid, 256, (size_t)smem, s, false, pdl, map, p));
  }
  IDISP_LAUNCH_CHECK();
  return IDISP_OK;
}

}  // namespace idisp
