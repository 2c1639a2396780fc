// Sliding-window validation on the device (reference utils/test_3d_patch.py:82-141,
// test_single_case): the class-`cls` softmax probability of one window's logits is added into a score map and the visit
// count is bumped, one launch per window in the reference's x,y,z window order (windows overlap, so accumulating them one
// after the other keeps the sums in the reference's order and needs no atomics); a second kernel divides by the count
// and thresholds.  Parity: tests/test_gpu_networks.py::test_sliding_window_validation (fixture minted from the reference function).
#include "common.cuh"
#include "../../include/bcp_b200.h"

namespace bcp {

template <int C>
__global__ void window_accumulate_kernel(const float* __restrict__ logits, float* __restrict__ score, float* __restrict__ count,
                                         int px, int py, int pz, int H, int D, int ox, int oy, int oz, int cls) {
  const long long PV = (long long)px * py * pz;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < PV; i += stride) {
    const int z = (int)(i % pz);
    const long long r = i / pz;
    const int y = (int)(r % py), x = (int)(r / py);
    float v[C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { v[c] = logits[(long long)c * PV + i]; m = fmaxf(m, v[c]); }
    float s = 0.f, e = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float t = expf(v[c] - m);
      s += t;
      if (c == cls) e = t;
    }
    const long long g = ((long long)(ox + x) * H + (oy + y)) * D + (oz + z);
    score[g] += __fdiv_rn(e, s);
    count[g] += 1.f;
  }
}

__global__ void window_finalize_kernel(float* __restrict__ score, const float* __restrict__ count, unsigned char* __restrict__ label,
                                       long long n, float thr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float p = __fdiv_rn(score[i], count[i]);        // every voxel of the (padded) volume is covered by >= 1 window
    score[i] = p;
    label[i] = p > thr ? 1 : 0;
  }
}

static inline int grid_1d(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace bcp

using namespace bcp;

extern "C" {

int bcp_window_accumulate(const float* logits, float* score, float* count, int c, int cls, const int* patch3, const int* vol3,
                          const int* origin3, cudaStream_t stream) {
  BCP_REQUIRE(logits && score && count && patch3 && vol3 && origin3, "window_accumulate: null pointer");
  BCP_REQUIRE((c == 2 || c == 4) && cls >= 0 && cls < c, "window_accumulate: classes %d / %d", c, cls);
  for (int i = 0; i < 3; ++i)
    BCP_REQUIRE(patch3[i] > 0 && origin3[i] >= 0 && origin3[i] + patch3[i] <= vol3[i], "window_accumulate: window leaves the volume");
  const long long pv = (long long)patch3[0] * patch3[1] * patch3[2];
  if (c == 2)
    window_accumulate_kernel<2><<<grid_1d(pv), 256, 0, stream>>>(logits, score, count, patch3[0], patch3[1], patch3[2], vol3[1], vol3[2],
                                                                  origin3[0], origin3[1], origin3[2], cls);
  else
    window_accumulate_kernel<4><<<grid_1d(pv), 256, 0, stream>>>(logits, score, count, patch3[0], patch3[1], patch3[2], vol3[1], vol3[2],
                                                                  origin3[0], origin3[1], origin3[2], cls);
  return check_launch("window_accumulate");
}

int bcp_window_finalize(float* score, const float* count, unsigned char* label, long long n, float threshold, cudaStream_t stream) {
  BCP_REQUIRE(score && count && label && n > 0, "window_finalize: bad args");
  window_finalize_kernel<<<grid_1d(n), 256, 0, stream>>>(score, count, label, n, threshold);
  return check_launch("window_finalize");
}

}  // extern "C"
