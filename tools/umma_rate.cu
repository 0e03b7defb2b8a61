// umma_rate.cu -- microbenchmark: tcgen05.mma (bf16, M=128, K=16, cta_group::1) throughput as a function
// of N and of the A-operand source/layout, one CTA per SM.  Issue pattern = the conv kernel's: a CONVERGED warp,
// one elect.sync lane, descriptors advanced by compile-time constants on the uniform datapath, 16 MMAs unrolled.
// (Two earlier versions of this tool measured the issuing thread instead of the tensor pipe: 78 cycles/MMA with
//  constant-divisor address math, 301 with runtime divisions, identical for every layout -- see profiles/r01_notes.md.)
// Data are zeros; only timing matters.  Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../disprcnn_b200/csrc/sm100_ptx.cuh"
using namespace idisp;

struct Variant { const char *name; int layout; int lbo, sbo; int tmem_a; int step_bytes; int nacc; };

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

template <int TMEM_A, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int outer, Variant v, long long *cycles)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t base = ptx::smem_u32(smem);
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) ptx::tmem_alloc<512>(ptx::smem_u32(&tmem_ptr));
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1) {
    const bool lead = ptx::elect_one();
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const uint64_t a0 = ptx::make_smem_desc(base + 65536, v.lbo, v.sbo) | ((uint64_t)v.layout << 61);
    const uint64_t b0 = ptx::make_smem_desc(base, N * 16, 128);
    const uint32_t step = (uint32_t)v.step_bytes >> 4;
    const long long t0 = clock64();
    for (int o = 0; o < outer; ++o) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t d = tmem + (NACC >= 100 ? (NACC - 100) : (i % NACC) * 128);  // NACC>=100: one accumulator at column offset NACC-100
        if (TMEM_A) { if (lead) umma_bf16_ts(d, tmem + 448 + (i & 7) * 8, b0 + (uint64_t)((i & 3) * 512), idesc, 1); }
        else if (lead) ptx::umma_bf16_ss(d, a0 + (uint64_t)((i % 9) * step), b0 + (uint64_t)((i & 3) * 512), idesc, 1);
      }
    }
    if (lead) ptx::umma_commit(ptx::smem_u32(&bar));
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (cycles && lead) cycles[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}

int main()
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long *d;
  cudaMalloc(&d, sizeof(long long) * sms);
  cudaFuncSetAttribute(rate_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(rate_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(rate_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(rate_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(rate_kernel<0, 132>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(rate_kernel<0, 164>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(rate_kernel<0, 196>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(rate_kernel<0, 116>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int outer = 256, iters = outer * 16;
  const Variant vars[] = {
      {"ns_conv_halo(lbo2880,sbo160,shift16B) 1acc", 0, 2880, 160, 0, 16, 1},   // the conv kernel's haloed layout, tap-shifted starts
      {"ns_conv_halo(lbo2880,sbo160,shift16B) 2acc", 0, 2880, 160, 0, 16, 2},
      {"ns_conv_halo(lbo2880,sbo160,shift16B) 4acc", 0, 2880, 160, 0, 16, 4},
      {"ns_conv_halo D@col32", 0, 2880, 160, 0, 16, 132},
      {"ns_conv_halo D@col64", 0, 2880, 160, 0, 16, 164},
      {"ns_conv_halo D@col96", 0, 2880, 160, 0, 16, 196},
      {"ns_conv_halo D@col16", 0, 2880, 160, 0, 16, 116},
      {"sw128(sbo1024) 1acc", 2, 16, 1024, 0, 32, 1},
      {"sw128(sbo1024) 4acc", 2, 16, 1024, 0, 32, 4},
      {"sw64(sbo512) 4acc", 4, 16, 512, 0, 32, 4},
      {"A_in_TMEM 2acc", 0, 0, 0, 1, 0, 2},
  };
  printf("{\"sms\": %d, \"iters\": %d, \"results\": [\n", sms, iters);
  bool firstrow = true;
  for (const Variant &v : vars) {
    for (int N : {32, 64, 96, 128, 192, 256}) {
      if (N > 128 && v.nacc > 2) continue;  // 4 x 128 columns / offset sweeps
      auto launch = [&](int o, long long *out) {
        if (v.tmem_a) rate_kernel<1, 2><<<sms, 128, 160 * 1024>>>(N, o, v, out);
        else if (v.nacc == 1) rate_kernel<0, 1><<<sms, 128, 160 * 1024>>>(N, o, v, out);
        else if (v.nacc == 2) rate_kernel<0, 2><<<sms, 128, 160 * 1024>>>(N, o, v, out);
        else if (v.nacc == 132) rate_kernel<0, 132><<<sms, 128, 160 * 1024>>>(N, o, v, out);
        else if (v.nacc == 164) rate_kernel<0, 164><<<sms, 128, 160 * 1024>>>(N, o, v, out);
        else if (v.nacc == 196) rate_kernel<0, 196><<<sms, 128, 160 * 1024>>>(N, o, v, out);
        else if (v.nacc == 116) rate_kernel<0, 116><<<sms, 128, 160 * 1024>>>(N, o, v, out);
        else rate_kernel<0, 4><<<sms, 128, 160 * 1024>>>(N, o, v, out);
      };
      launch(4, nullptr);
      launch(outer, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("\n], \"error\": \"%s at %s N=%d\"}\n", cudaGetErrorString(e), v.name, N); return 1; }
      std::vector<long long> h(sms);
      cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      const double cyc = (double)mx / iters;
      printf("%s  {\"A\": \"%s\", \"N\": %d, \"cycles_per_mma\": %.2f, \"tensor_floor\": %.1f, \"frac_of_floor\": %.3f, \"smem_B_per_clk\": %.1f}",
             firstrow ? "" : ",\n", v.name, N, cyc, N / 2.0, (N / 2.0) / cyc, ((v.tmem_a ? 0 : 4096.0) + N * 32) / cyc);
      firstrow = false;
    }
  }
  printf("\n]}\n");
  return 0;
}
