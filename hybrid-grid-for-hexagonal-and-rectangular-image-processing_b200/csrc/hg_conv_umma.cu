// hg_conv_umma.cu -- tcgen05 / TMEM implicit-GEMM hex convolution (placeholder until the kernel lands:
// reports "not eligible", so hg_conv.cu routes everything to the direct stencil).
#include "hg_conv.cuh"

namespace hg {
bool conv_umma_eligible(const hg_conv_desc*, int) { return false; }
int conv_fwd_umma(const hg_conv_desc*, const ConvGeom&, const ConvTaps&, const void*, const float*, const float*, void*, cudaStream_t) {
  set_error("tcgen05 path not built");
  return HG_E_UNSUPPORTED;
}
int conv_dgrad_umma(const hg_conv_desc*, const ConvGeom&, const ConvTaps&, const void*, const float*, void*, cudaStream_t) {
  set_error("tcgen05 path not built");
  return HG_E_UNSUPPORTED;
}
int conv_wgrad_umma(const hg_conv_desc*, const ConvGeom&, const ConvTaps&, const void*, const void*, float*, float*, cudaStream_t) {
  set_error("tcgen05 path not built");
  return HG_E_UNSUPPORTED;
}
}  // namespace hg
