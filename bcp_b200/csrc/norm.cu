// Train-mode BatchNorm / InstanceNorm statistics, fused normalise+activation(+dropout)(+residual)
// forward, and the two-pass backward, on channel-blocked bf16 activations (CB8: [N][C/8][S][8]).
// Statistics are per (group, channel) where a group is `spg` consecutive samples: BatchNorm over a
// reference forward call of batch b uses spg=b (several calls may be batched as several groups);
// InstanceNorm uses spg=1.  All reductions are fixed-order two-stage (deterministic, no atomics).
// Reference modules: nn.BatchNorm3d/2d (networks/VNet.py:19, networks/unet.py:21,25),
// nn.InstanceNorm3d (pancreas/Vnet.py:25,49,76), ReLU / LeakyReLU(0.01), Dropout3d / Dropout.
#include "common.cuh"
#include "../../include/bcp_b200.h"

namespace bcp {

constexpr int NT = 256;

// Per-channel finalize, executed by ONE block (the last one to finish its partial): fixed-order double-precision
// reduce of the partials => deterministic no matter which block happens to be last.
// stat[g][C][2] = {mean, invstd}; coef[g][C][2] = {scale, shift}.
__device__ void bn_finalize_block(const float* __restrict__ partial, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ running_mean,
                                  float* __restrict__ running_var, long long* __restrict__ nbt,
                                  float* __restrict__ stat, float* __restrict__ coef,
                                  int N, int C, long long S, int chunks, int spg, float eps, float momentum) {
  const int Cb = (C + 7) / 8;
  const int G = N / spg;
  if (threadIdx.x == 0 && nbt != nullptr) nbt[0] += G;
  const double M = (double)spg * (double)S;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int cb = c >> 3, k = c & 7;
    float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
    for (int g = 0; g < G; ++g) {
      double s = 0.0, q = 0.0;
      for (int n = g * spg; n < (g + 1) * spg; ++n) {
        const float* p = partial + ((long long)n * Cb + cb) * chunks * 16;
        for (int ch = 0; ch < chunks; ++ch) { s += (double)p[ch * 16 + k]; q += (double)p[ch * 16 + 8 + k]; }
      }
      const double mean = s / M;
      double var = q / M - mean * mean;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)eps));
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float scale = ga * invstd;
      stat[((long long)g * C + c) * 2 + 0] = (float)mean;
      stat[((long long)g * C + c) * 2 + 1] = invstd;
      coef[((long long)g * C + c) * 2 + 0] = scale;
      coef[((long long)g * C + c) * 2 + 1] = be - (float)mean * scale;
      if (running_mean) {   // sequential per-call update, like calling the module once per group
        const double unbiased = (M > 1.0) ? var * M / (M - 1.0) : var;
        rm = (1.f - momentum) * rm + momentum * (float)mean;
        rv = (1.f - momentum) * rv + momentum * (float)unbiased;
      }
    }
    if (running_mean) { running_mean[c] = rm; running_var[c] = rv; }
  }
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

// partial[((n*Cb + cb)*chunks + chunk)*16 + {0..7: sum, 8..15: sumsq}]; the last block finalises
__global__ void __launch_bounds__(NT) bn_stats_kernel(const uint4* __restrict__ y, float* __restrict__ partial,
                                                       long long S, int chunks, int* __restrict__ counter,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                                       long long* __restrict__ nbt, float* __restrict__ stat,
                                                       float* __restrict__ coef, int N, int C, int spg, float eps, float momentum) {
  const int chunk = blockIdx.x, cb = blockIdx.y, n = blockIdx.z, Cb = gridDim.y;
  const uint4* base = y + ((long long)n * Cb + cb) * S;
  const long long per = (S + chunks - 1) / chunks;
  const long long s0 = (long long)chunk * per;
  const long long s1 = min(S, s0 + per);
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (long long s = s0 + threadIdx.x; s < s1; s += NT) {
    float f[8];
    unpack8(ldg_nc_u4(base + s), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] += f[k] * f[k]; }
  }
  __shared__ float red[16 * (NT / 32)];
  block_sum<16, NT>(acc, red);
  if (threadIdx.x == 0) {
    float* dst = partial + (((long long)n * Cb + cb) * chunks + chunk) * 16;
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k] = acc[k];
  }
  if (last_block_arrives(counter))
    bn_finalize_block(partial, gamma, beta, running_mean, running_var, nbt, stat, coef, N, C, S, chunks, spg, eps, momentum);
}

// eval-mode coefficients from running statistics
__global__ void bn_eval_coef_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv,
                                    float* __restrict__ stat, float* __restrict__ coef, int C, int G, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(rv[c] + eps);
  const float scale = (gamma ? gamma[c] : 1.f) * invstd;
  for (int g = 0; g < G; ++g) {
    stat[((long long)g * C + c) * 2 + 0] = rm[c];
    stat[((long long)g * C + c) * 2 + 1] = invstd;
    coef[((long long)g * C + c) * 2 + 0] = scale;
    coef[((long long)g * C + c) * 2 + 1] = (beta ? beta[c] : 0.f) - rm[c] * scale;
  }
}

// out = act(y*scale + shift) [* chan_scale[n][c]] [* keep*elem_scale] [+ residual]
__global__ void __launch_bounds__(NT) bn_apply_kernel(const uint4* __restrict__ y, uint4* __restrict__ out,
                                                       const float* __restrict__ coef, const float* __restrict__ chan_scale,
                                                       const unsigned char* __restrict__ elem_keep, float elem_scale,
                                                       const uint4* __restrict__ residual, int C, long long S, int spg,
                                                       float slope) {
  const int cb = blockIdx.y, n = blockIdx.z, Cb = gridDim.y;
  const int g = n / spg;
  __shared__ float sc[8], sh[8], cs[8];
  if (threadIdx.x < 8) {
    const int c = cb * 8 + threadIdx.x;
    const bool ok = c < C;
    sc[threadIdx.x] = ok ? coef[((long long)g * C + c) * 2] : 0.f;
    sh[threadIdx.x] = ok ? coef[((long long)g * C + c) * 2 + 1] : 0.f;
    cs[threadIdx.x] = (ok && chan_scale) ? chan_scale[(long long)n * C + c] : 1.f;
  }
  __syncthreads();
  const long long plane = ((long long)n * Cb + cb) * S;
  const long long stride = (long long)gridDim.x * NT;
  for (long long s = (long long)blockIdx.x * NT + threadIdx.x; s < S; s += stride) {
    float f[8];
    unpack8(ldg_nc_u4(y + plane + s), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = f[k] * sc[k] + sh[k];
      v = v > 0.f ? v : v * slope;
      f[k] = v * cs[k];
    }
    if (elem_keep) {
      const uint2 kp = *reinterpret_cast<const uint2*>(elem_keep + (plane + s) * 8);
      const unsigned char* kb = reinterpret_cast<const unsigned char*>(&kp);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = kb[k] ? f[k] * elem_scale : 0.f;
    }
    if (residual) {
      float r[8];
      unpack8(ldg_nc_u4(residual + plane + s), r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
    out[plane + s] = pack8(f);
  }
}

__device__ void bn_bwd_finalize_block(const float* __restrict__ partial, float* __restrict__ sums, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, int N, int C, long long S, int chunks, int spg, int accumulate);

// backward pass 1: per (n, cb, chunk) partials of  s1 = sum g,  s2 = sum g*xhat   with
// g = da * dropout * act'(pre)
__global__ void __launch_bounds__(NT) bn_bwd_reduce_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y,
                                                            const float* __restrict__ stat, const float* __restrict__ coef,
                                                            const float* __restrict__ chan_scale,
                                                            const unsigned char* __restrict__ elem_keep, float elem_scale,
                                                            float* __restrict__ partial, int C, long long S, int chunks,
                                                            int spg, float slope, int* __restrict__ counter,
                                                            float* __restrict__ sums, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int N, int accumulate) {
  const int chunk = blockIdx.x, cb = blockIdx.y, n = blockIdx.z, Cb = gridDim.y;
  const int g = n / spg;
  __shared__ float sc[8], sh[8], cs[8], mu[8], is[8];
  if (threadIdx.x < 8) {
    const int c = cb * 8 + threadIdx.x;
    const bool ok = c < C;
    sc[threadIdx.x] = ok ? coef[((long long)g * C + c) * 2] : 0.f;
    sh[threadIdx.x] = ok ? coef[((long long)g * C + c) * 2 + 1] : 0.f;
    mu[threadIdx.x] = ok ? stat[((long long)g * C + c) * 2] : 0.f;
    is[threadIdx.x] = ok ? stat[((long long)g * C + c) * 2 + 1] : 0.f;
    cs[threadIdx.x] = (ok && chan_scale) ? chan_scale[(long long)n * C + c] : 1.f;
  }
  __syncthreads();
  const long long plane = ((long long)n * Cb + cb) * S;
  const long long per = (S + chunks - 1) / chunks;
  const long long s0 = (long long)chunk * per, s1 = min(S, s0 + per);
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (long long s = s0 + threadIdx.x; s < s1; s += NT) {
    float fy[8], fd[8];
    unpack8(ldg_nc_u4(y + plane + s), fy);
    unpack8(ldg_nc_u4(da + plane + s), fd);
    unsigned char kb[8];
    if (elem_keep) *reinterpret_cast<uint2*>(kb) = *reinterpret_cast<const uint2*>(elem_keep + (plane + s) * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float pre = fy[k] * sc[k] + sh[k];
      float gk = fd[k] * cs[k];
      if (elem_keep) gk = kb[k] ? gk * elem_scale : 0.f;
      gk = pre > 0.f ? gk : gk * slope;
      acc[k] += gk;
      acc[8 + k] += gk * ((fy[k] - mu[k]) * is[k]);
    }
  }
  __shared__ float red[16 * (NT / 32)];
  block_sum<16, NT>(acc, red);
  if (threadIdx.x == 0) {
    float* dst = partial + (((long long)n * Cb + cb) * chunks + chunk) * 16;
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k] = acc[k];
  }
  if (last_block_arrives(counter)) bn_bwd_finalize_block(partial, sums, dgamma, dbeta, N, C, S, chunks, spg, accumulate);
}

// sums[g][C][2] = {s1/M, s2/M};  dgamma[c] = sum_g s2, dbeta[c] = sum_g s1   (run by the last block of the reduce pass)
__device__ void bn_bwd_finalize_block(const float* __restrict__ partial, float* __restrict__ sums,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                      int N, int C, long long S, int chunks, int spg, int accumulate) {
  const int Cb = (C + 7) / 8, G = N / spg;
  const double M = (double)spg * (double)S;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int cb = c >> 3, k = c & 7;
    double tg = 0.0, tb = 0.0;
    for (int g = 0; g < G; ++g) {
      double a = 0.0, b = 0.0;
      for (int n = g * spg; n < (g + 1) * spg; ++n) {
        const float* p = partial + ((long long)n * Cb + cb) * chunks * 16;
        for (int ch = 0; ch < chunks; ++ch) { a += (double)p[ch * 16 + k]; b += (double)p[ch * 16 + 8 + k]; }
      }
      sums[((long long)g * C + c) * 2 + 0] = (float)(a / M);
      sums[((long long)g * C + c) * 2 + 1] = (float)(b / M);
      tb += a; tg += b;
    }
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)tg : (float)tg;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)tb : (float)tb;
  }
}

// backward pass 2:  dy = scale * (g - mean(g) - xhat*mean(g*xhat))      (stats_grad=1)
//                   dy = scale * g                                       (stats_grad=0: eval / no norm)
__global__ void __launch_bounds__(NT) bn_bwd_apply_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y,
                                                           uint4* __restrict__ dy, const float* __restrict__ stat,
                                                           const float* __restrict__ coef, const float* __restrict__ sums,
                                                           const float* __restrict__ chan_scale,
                                                           const unsigned char* __restrict__ elem_keep, float elem_scale,
                                                           int C, long long S, int spg, float slope, int stats_grad) {
  const int cb = blockIdx.y, n = blockIdx.z, Cb = gridDim.y;
  const int g = n / spg;
  __shared__ float sc[8], sh[8], cs[8], mu[8], is[8], m1[8], m2[8];
  if (threadIdx.x < 8) {
    const int c = cb * 8 + threadIdx.x;
    const bool ok = c < C;
    sc[threadIdx.x] = ok ? coef[((long long)g * C + c) * 2] : 0.f;
    sh[threadIdx.x] = ok ? coef[((long long)g * C + c) * 2 + 1] : 0.f;
    mu[threadIdx.x] = ok ? stat[((long long)g * C + c) * 2] : 0.f;
    is[threadIdx.x] = ok ? stat[((long long)g * C + c) * 2 + 1] : 0.f;
    m1[threadIdx.x] = (ok && stats_grad) ? sums[((long long)g * C + c) * 2] : 0.f;
    m2[threadIdx.x] = (ok && stats_grad) ? sums[((long long)g * C + c) * 2 + 1] : 0.f;
    cs[threadIdx.x] = (ok && chan_scale) ? chan_scale[(long long)n * C + c] : 1.f;
  }
  __syncthreads();
  const long long plane = ((long long)n * Cb + cb) * S;
  const long long stride = (long long)gridDim.x * NT;
  for (long long s = (long long)blockIdx.x * NT + threadIdx.x; s < S; s += stride) {
    float fy[8], fd[8];
    unpack8(ldg_nc_u4(y + plane + s), fy);
    unpack8(ldg_nc_u4(da + plane + s), fd);
    unsigned char kb[8];
    if (elem_keep) *reinterpret_cast<uint2*>(kb) = *reinterpret_cast<const uint2*>(elem_keep + (plane + s) * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float pre = fy[k] * sc[k] + sh[k];
      float gk = fd[k] * cs[k];
      if (elem_keep) gk = kb[k] ? gk * elem_scale : 0.f;
      gk = pre > 0.f ? gk : gk * slope;
      const float xhat = (fy[k] - mu[k]) * is[k];
      fd[k] = sc[k] * (gk - m1[k] - xhat * m2[k]);
    }
    dy[plane + s] = pack8(fd);
  }
}

// Reduction grid = chunks x (C/8) x N blocks.  Aim at ~4 blocks per SM so mid-size layers are not latency-bound on a
// handful of blocks, but keep at least 1024 voxels (16 KB) per block and at most 256 partials per (n, c/8) plane.
static inline int pick_chunks(long long S, int planes) {
  const long long target = 4LL * sm_count();
  long long c = (target + planes - 1) / planes;
  const long long cmax = (S + 1023) / 1024;
  if (c > cmax) c = cmax;
  if (c > 256) c = 256;
  if (c < 1) c = 1;
  return (int)c;
}

}  // namespace bcp

using namespace bcp;

extern "C" {

int bcp_norm_chunks(int n, int c, long long s) { return pick_chunks(s, n * ((c + 7) / 8)); }

long long bcp_norm_workspace_floats(int n, int c, long long s) {
  return (long long)n * ((c + 7) / 8) * pick_chunks(s, n * ((c + 7) / 8)) * 16;
}

int bcp_norm_stats(const void* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                   long long* num_batches_tracked, float* stat, float* coef, float* workspace, int* counter,
                   int n, int c, long long s, int spg, float eps, float momentum, cudaStream_t stream) {
  BCP_REQUIRE(y && stat && coef && workspace && counter, "norm_stats: null pointer");
  BCP_REQUIRE(n > 0 && c > 0 && s > 0 && spg > 0 && n % spg == 0, "norm_stats: bad shape n=%d spg=%d", n, spg);
  const int Cb = (c + 7) / 8, chunks = pick_chunks(s, n * Cb);
  dim3 grid(chunks, Cb, n);
  bn_stats_kernel<<<grid, NT, 0, stream>>>((const uint4*)y, workspace, s, chunks, counter, gamma, beta, running_mean, running_var,
                                           num_batches_tracked, stat, coef, n, c, spg, eps, momentum);
  return check_launch("norm_stats");
}

int bcp_norm_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float* stat, float* coef, int c, int groups, float eps, cudaStream_t stream) {
  BCP_REQUIRE(running_mean && running_var && stat && coef && c > 0 && groups > 0, "norm_eval_coef: bad args");
  bn_eval_coef_kernel<<<(c + 127) / 128, 128, 0, stream>>>(gamma, beta, running_mean, running_var, stat, coef, c, groups, eps);
  return check_launch("norm_eval_coef");
}

int bcp_norm_apply(const void* y, void* out, const float* coef, const float* chan_scale, const unsigned char* elem_keep,
                   float elem_scale, const void* residual, int n, int c, long long s, int spg, float slope,
                   cudaStream_t stream) {
  BCP_REQUIRE(y && out && coef, "norm_apply: null pointer");
  BCP_REQUIRE(n > 0 && c > 0 && s > 0 && spg > 0 && n % spg == 0, "norm_apply: bad shape");
  const int Cb = (c + 7) / 8;
  int gx = (int)((s + NT * 4 - 1) / (NT * 4));
  if (gx < 1) gx = 1;
  dim3 grid(gx, Cb, n);
  bn_apply_kernel<<<grid, NT, 0, stream>>>((const uint4*)y, (uint4*)out, coef, chan_scale, elem_keep, elem_scale,
                                           (const uint4*)residual, c, s, spg, slope);
  return check_launch("norm_apply");
}

int bcp_norm_bwd(const void* dact, const void* y, void* dy, const float* stat, const float* coef, const float* chan_scale,
                 const unsigned char* elem_keep, float elem_scale, float* dgamma, float* dbeta, float* sums,
                 float* workspace, int* counter, int n, int c, long long s, int spg, float slope, int stats_grad,
                 int accumulate, cudaStream_t stream) {
  BCP_REQUIRE(dact && y && dy && stat && coef && sums && workspace && counter, "norm_bwd: null pointer");
  BCP_REQUIRE(n > 0 && c > 0 && s > 0 && spg > 0 && n % spg == 0, "norm_bwd: bad shape");
  const int Cb = (c + 7) / 8, chunks = pick_chunks(s, n * Cb);
  if (stats_grad || dgamma || dbeta) {
    dim3 grid(chunks, Cb, n);
    bn_bwd_reduce_kernel<<<grid, NT, 0, stream>>>((const uint4*)dact, (const uint4*)y, stat, coef, chan_scale, elem_keep,
                                                  elem_scale, workspace, c, s, chunks, spg, slope, counter, sums, dgamma, dbeta, n, accumulate);
  }
  int gx = (int)((s + NT * 4 - 1) / (NT * 4));
  if (gx < 1) gx = 1;
  dim3 grid2(gx, Cb, n);
  bn_bwd_apply_kernel<<<grid2, NT, 0, stream>>>((const uint4*)dact, (const uint4*)y, (uint4*)dy, stat, coef, sums, chan_scale,
                                                elem_keep, elem_scale, c, s, spg, slope, stats_grad);
  return check_launch("norm_bwd");
}

}  // extern "C"
