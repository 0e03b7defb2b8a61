// conv3d_tc.cu -- tcgen05 implicit-GEMM 3x3x3 convolution (placeholder until the kernel lands).
#include "conv3d_tc.cuh"

namespace idisp {
int tc_weights_prepare(const float *, int kind, int cin, int cout, TcWeights &out, cudaStream_t)
{
  out.kind = kind; out.cin = cin; out.cout = cout;
  return IDISP_OK;
}
void tc_weights_free(TcWeights &w)
{
  if (w.dev) cudaFree(w.dev);
  w.dev = nullptr; w.bytes = 0;
}
bool tc_supported(int, int, int, int, int, int) { return false; }
int tc_conv3d(const TcWeights &, const __nv_bfloat16 *, int, int, int, int, int, int, int, const float *,
              const __nv_bfloat16 *, int, __nv_bfloat16 *, cudaStream_t)
{
  set_error("tc_conv3d: not built");
  return IDISP_ERR_UNSUPPORTED;
}
}  // namespace idisp
