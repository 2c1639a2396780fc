// placeholder until the tcgen05 kernel lands
#include "common.cuh"
#include "../../include/bcp_b200.h"
extern "C" {
int bcp_conv_tc_supported(int, int, const int*, const int*) { return 0; }
int bcp_conv_tc_fwd(const void*, const void*, const float*, void*, int, int, int, const int*, const int*, cudaStream_t) {
  bcp::set_last_error("conv_tc_fwd: not built");
  return BCP_ERR_UNSUPPORTED;
}
}
