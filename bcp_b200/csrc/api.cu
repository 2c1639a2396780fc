// Error plumbing and library-wide queries for the C ABI (include/bcp_b200.h).
#include "common.cuh"
#include "../../include/bcp_b200.h"
#include <stdarg.h>
#include <string.h>

namespace bcp {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_last_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return BCP_ERR_CUDA;
  }
  return BCP_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace bcp

extern "C" {

const char* bcp_last_error(void) { return bcp::g_last_error; }

int bcp_abi_version(void) { return BCP_B200_ABI_VERSION; }

int bcp_device_sm_count(void) { return bcp::sm_count(); }

// launches a counter kernel-free query: number of kernels is tracked by the host wrapper, not here.

}  // extern "C"
