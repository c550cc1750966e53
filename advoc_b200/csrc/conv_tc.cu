// tcgen05 / TMA implicit-GEMM convolution (placeholder until the kernel lands).
#include "epilogue.cuh"
namespace advoc {
bool conv_fwd_tc_eligible(const advoc_conv_desc*, int) { return false; }
bool conv_transposed_tc_eligible(const advoc_conv_desc*, int) { return false; }
int conv_fwd_tc(const advoc_conv_desc*, const float*, int, const float*, const advoc_epilogue*, void*) {
  return fail(ADVOC_UNSUPPORTED, "tcgen05 path not built");
}
int conv_transposed_tc(const advoc_conv_desc*, const float*, int, const float*, const advoc_epilogue*, void*) {
  return fail(ADVOC_UNSUPPORTED, "tcgen05 path not built");
}
}
