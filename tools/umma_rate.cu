// umma_rate.cu -- microbenchmark: tcgen05.mma (bf16, M=128, K=16, cta_group::1) issue rate as a function
// of N and of the A-operand source/layout, one CTA per SM.  Questions it answers for the conv kernel:
//   * is a small-N SS-mode MMA bound by the shared-memory A read rather than by the tensor pipe?
//   * does the no-swizzle "interleave" layout (16-byte rows) read slower than the 32/64/128-byte swizzles?
//   * what does A-from-TMEM (TS mode) cost?
// Data are zeros; only timing matters.  Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../disprcnn_b200/csrc/sm100_ptx.cuh"
using namespace idisp;

struct Variant { const char *name; int layout; int lbo, sbo; int tmem_a; int kstep_bytes; int ksteps; };

__device__ __forceinline__ uint64_t desc_lt(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
  return ptx::make_smem_desc(addr, lbo, sbo) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, Variant v, long long *cycles)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t base = ptx::smem_u32(smem);
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(ptx::smem_u32(&tmem_ptr));
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const uint32_t a_region = base + 65536;  // 1024-aligned
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t b = base + (i % 4) * 8192;
      const uint64_t bd = ptx::make_smem_desc(b, N * 16, 128);
      const uint32_t d = tmem + (N <= 128 ? (i % 2) * 128 : 0);
      if (v.tmem_a) {
        umma_bf16_ts(d, tmem + 256 + (i % 8) * 8, bd, idesc, 1);
      } else {
        const uint32_t a = a_region + (i % v.ksteps) * v.kstep_bytes + ((i / v.ksteps) % 2) * 32768;
        ptx::umma_bf16_ss(d, desc_lt(a, v.lbo, v.sbo, v.layout), bd, idesc, 1);
      }
    }
    ptx::umma_commit(ptx::smem_u32(&bar));
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

int main()
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long *d;
  cudaMalloc(&d, sizeof(long long) * sms);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 4096;
  const Variant vars[] = {
      {"ns_conv_halo(lbo2880,sbo160)", 0, 2880, 160, 0, 16, 9},   // the conv kernel's haloed layout, tap-shifted starts
      {"ns_dense(lbo128,sbo256)", 0, 128, 256, 0, 4096, 4},
      {"ns_dense(lbo2048,sbo128)", 0, 2048, 128, 0, 4096, 4},
      {"sw128(sbo1024)", 2, 16, 1024, 0, 32, 4},
      {"sw64(sbo512)", 4, 16, 512, 0, 32, 2},
      {"sw32(sbo256)", 6, 16, 256, 0, 4096, 4},
      {"A_in_TMEM", 0, 0, 0, 1, 0, 1},
  };
  printf("{\"sms\": %d, \"iters\": %d, \"results\": [\n", sms, iters);
  bool firstrow = true;
  for (const Variant &v : vars) {
    for (int N : {32, 64, 96, 128, 256}) {
      if (v.tmem_a && N > 128) continue;
      rate_kernel<<<sms, 128, 160 * 1024>>>(N, 64, v, nullptr);  // warm-up
      rate_kernel<<<sms, 128, 160 * 1024>>>(N, iters, v, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("\n], \"error\": \"%s at %s N=%d\"}\n", cudaGetErrorString(e), v.name, N); return 1; }
      std::vector<long long> h(sms);
      cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      const double cyc = (double)mx / iters;
      printf("%s  {\"A\": \"%s\", \"N\": %d, \"cycles_per_mma\": %.2f, \"tensor_floor\": %.1f, \"frac_of_floor\": %.3f}",
             firstrow ? "" : ",\n", v.name, N, cyc, N / 2.0, (N / 2.0) / cyc);
      firstrow = false;
    }
  }
  printf("\n]}\n");
  return 0;
}
