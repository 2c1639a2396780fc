// Shared device/host helpers for the bcp_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace bcp {

// ---- error plumbing (C ABI never throws/aborts; see include/bcp_b200.h) ----
void set_last_error(const char* fmt, ...);
int  check_launch(const char* what);

#define BCP_OK 0
#define BCP_ERR_ARG (-1)
#define BCP_ERR_UNSUPPORTED (-2)
#define BCP_ERR_CUDA (-3)

#define BCP_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      ::bcp::set_last_error(__VA_ARGS__);                       \
      return BCP_ERR_ARG;                                       \
    }                                                           \
  } while (0)

// ---- bf16 x8 (one 16-byte channel block of the CB8 layout) ----
struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction of K floats per thread; result valid in thread 0 (all K).
template <int K, int THREADS>
__device__ __forceinline__ void block_sum(float (&v)[K], float* smem /* K * THREADS/32 floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) smem[warp * K + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float s = 0.f;
      for (int w = 0; w < THREADS / 32; ++w) s += smem[w * K + k];
      v[k] = s;
    }
  }
  __syncthreads();
}

// true in every thread of exactly one block: the last block of the grid to arrive (threadfence-reduction pattern).
// `counter` must be zero on entry and is reset to zero by the last block => reusable by the next launch on the stream.
__device__ __forceinline__ bool last_block_arrives(int* counter) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = gridDim.x * gridDim.y * gridDim.z;
    const int prev = atomicAdd(counter, 1);
    is_last = (prev == total - 1);
    if (is_last) *counter = 0;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last != 0;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();

}  // namespace bcp
