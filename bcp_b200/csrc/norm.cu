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

// ---- finalize stage shared by the forward statistics and the backward reduction -------------------------------------
// The partial buffer holds one 16-float record per (sample, channel octet, chunk): {8 x sum_a, 8 x sum_b}.  The LAST block to
// arrive reduces them to per-(group, channel) double-precision totals in shared memory: one warp per (group, octet) task,
// one record per lane per iteration (four independent 16-byte loads in flight per lane, so the L2 round trips of a task
// overlap instead of forming the dependent chain the first version had: 64 sequential loads ~ 20 us on a 64-channel layer),
// then a fixed butterfly => bit-identical run to run no matter which block happens to be last.
constexpr int FIN_PAIRS = 1024;          // (group, channel) pairs per shared-memory tile

__device__ __forceinline__ void fin_reduce_tile(const float* __restrict__ partial, double* __restrict__ sm_a,
                                                double* __restrict__ sm_b, int G, int Cb, int cb0, int ncb, int chunks, int spg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per_group = spg * chunks;
  for (int task = warp; task < G * ncb; task += nwarps) {
    const int g = task / ncb, cbl = task - g * ncb, cb = cb0 + cbl;
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (int e = lane; e < per_group; e += 32) {
      const int ns = e / chunks, ch = e - ns * chunks;
      const float4* p = reinterpret_cast<const float4*>(partial + ((((long long)(g * spg + ns)) * Cb + cb) * chunks + ch) * 16);
      const float4 v0 = __ldcg(p), v1 = __ldcg(p + 1), v2 = __ldcg(p + 2), v3 = __ldcg(p + 3);
      acc[0] += (double)v0.x; acc[1] += (double)v0.y; acc[2] += (double)v0.z; acc[3] += (double)v0.w;
      acc[4] += (double)v1.x; acc[5] += (double)v1.y; acc[6] += (double)v1.z; acc[7] += (double)v1.w;
      acc[8] += (double)v2.x; acc[9] += (double)v2.y; acc[10] += (double)v2.z; acc[11] += (double)v2.w;
      acc[12] += (double)v3.x; acc[13] += (double)v3.y; acc[14] += (double)v3.z; acc[15] += (double)v3.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { sm_a[(g * ncb + cbl) * 8 + k] = acc[k]; sm_b[(g * ncb + cbl) * 8 + k] = acc[8 + k]; }
    }
  }
}

// Per-channel finalize, executed by ONE block (the last one to finish its partial).
// stat[g][C][2] = {mean, invstd}; coef[g][C][2] = {scale, shift}.
__device__ __noinline__ void bn_finalize_block(const float* __restrict__ partial, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ running_mean,
                                  float* __restrict__ running_var, long long* __restrict__ nbt,
                                  float* __restrict__ stat, float* __restrict__ coef,
                                  int N, int C, long long S, int chunks, int spg, float eps, float momentum) {
  const int Cb = (C + 7) / 8;
  const int G = N / spg;
  if (threadIdx.x == 0 && nbt != nullptr) nbt[0] += G;
  const double M = (double)spg * (double)S;
  __shared__ double sm_s[FIN_PAIRS], sm_q[FIN_PAIRS];
  int tile_cb = FIN_PAIRS / (8 * G);
  if (tile_cb < 1) tile_cb = 1;                       // G > 128 groups: the host wrapper rejects such shapes
  for (int cb0 = 0; cb0 < Cb; cb0 += tile_cb) {
    const int ncb = min(tile_cb, Cb - cb0);
    __syncthreads();
    fin_reduce_tile(partial, sm_s, sm_q, G, Cb, cb0, ncb, chunks, spg);
    __syncthreads();
    // the double-precision mean / variance / running-statistics chain runs one channel per THREAD
    for (int t = threadIdx.x; t < ncb * 8; t += blockDim.x) {
      const int c = cb0 * 8 + t;
      if (c >= C) continue;
      float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
      for (int g = 0; g < G; ++g) {
        const double s = sm_s[g * ncb * 8 + t], q = sm_q[g * ncb * 8 + t];
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
}

// partial[((n*Cb + cb)*chunks + chunk)*16 + {0..7: sum, 8..15: sumsq}]; the last block finalises
// (minimum 6 resident blocks per SM: the streaming loop sets the register budget, the rarely-run finalize may spill)
__global__ void __launch_bounds__(NT, 6) bn_stats_kernel(const uint4* __restrict__ y, float* __restrict__ partial,
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
  long long s = s0 + threadIdx.x;
  for (; s + 3 * NT < s1; s += 4 * NT) {          // four independent 16-byte loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldg_nc_u4(base + s + u * NT);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] += f[k] * f[k]; }
    }
  }
  for (; s < s1; s += NT) {
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

// Tiny layers (N*S <= SMALL_LIMIT voxels): ONE block per channel octet walks all samples and finishes the statistics
// itself -- no partial buffer, no grid-wide arrival counter (their fixed cost dominated the deep layers).
constexpr long long SMALL_LIMIT = 2048;     // beyond this the C/8 blocks of the single-launch path are latency-bound
__global__ void __launch_bounds__(NT) bn_stats_small_kernel(const uint4* __restrict__ y, long long S,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* __restrict__ running_mean, float* __restrict__ running_var,
                                                             long long* __restrict__ nbt, float* __restrict__ stat,
                                                             float* __restrict__ coef, int N, int C, int spg, float eps, float momentum) {
  const int cb = blockIdx.x, Cb = gridDim.x, G = N / spg;
  __shared__ float red[16 * (NT / 32)];
  __shared__ float tot[16];
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) nbt[0] += G;
  const double M = (double)spg * (double)S;
  const int c = cb * 8 + (int)threadIdx.x;
  const bool owner = threadIdx.x < 8 && c < C;
  float rm = (owner && running_mean) ? running_mean[c] : 0.f, rv = (owner && running_var) ? running_var[c] : 0.f;
  for (int g = 0; g < G; ++g) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    const long long total = (long long)spg * S;
    for (long long i = threadIdx.x; i < total; i += NT) {
      const int n = g * spg + (int)(i / S);
      const long long sp = i - (i / S) * S;
      float f[8];
      unpack8(ldg_nc_u4(y + ((long long)n * Cb + cb) * S + sp), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] += f[k] * f[k]; }
    }
    block_sum<16, NT>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 16; ++k) tot[k] = acc[k];
    }
    __syncthreads();
    if (owner) {
      const double mean = (double)tot[threadIdx.x] / M;
      double var = (double)tot[8 + threadIdx.x] / M - mean * mean;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)eps));
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float scale = ga * invstd;
      stat[((long long)g * C + c) * 2 + 0] = (float)mean;
      stat[((long long)g * C + c) * 2 + 1] = invstd;
      coef[((long long)g * C + c) * 2 + 0] = scale;
      coef[((long long)g * C + c) * 2 + 1] = be - (float)mean * scale;
      if (running_mean) {
        const double unbiased = (M > 1.0) ? var * M / (M - 1.0) : var;
        rm = (1.f - momentum) * rm + momentum * (float)mean;
        rv = (1.f - momentum) * rv + momentum * (float)unbiased;
      }
    }
    __syncthreads();
  }
  if (owner && running_mean) { running_mean[c] = rm; running_var[c] = rv; }
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

__device__ __noinline__ void bn_bwd_finalize_block(const float* __restrict__ partial, float* __restrict__ sums, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, int N, int C, long long S, int chunks, int spg, int accumulate);

// backward pass 1: per (n, cb, chunk) partials of  s1 = sum g,  s2 = sum g*xhat   with
// g = da * dropout * act'(pre)
__global__ void __launch_bounds__(NT, 4) bn_bwd_reduce_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y,
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
  auto accumulate_one = [&](const uint4& vy, const uint4& vd, const uint2& kp) {
    float fy[8], fd[8];
    unpack8(vy, fy);
    unpack8(vd, fd);
    const unsigned char* kb = reinterpret_cast<const unsigned char*>(&kp);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float pre = fy[k] * sc[k] + sh[k];
      float gk = fd[k] * cs[k];
      if (elem_keep) gk = kb[k] ? gk * elem_scale : 0.f;
      gk = pre > 0.f ? gk : gk * slope;
      acc[k] += gk;
      acc[8 + k] += gk * ((fy[k] - mu[k]) * is[k]);
    }
  };
  long long s = s0 + threadIdx.x;
  for (; s + NT < s1; s += 2 * NT) {               // two voxels = four independent 16-byte loads in flight per thread
    const uint4 y0 = ldg_nc_u4(y + plane + s), y1 = ldg_nc_u4(y + plane + s + NT);
    const uint4 d0 = ldg_nc_u4(da + plane + s), d1 = ldg_nc_u4(da + plane + s + NT);
    uint2 k0 = make_uint2(0, 0), k1 = make_uint2(0, 0);
    if (elem_keep) {
      k0 = *reinterpret_cast<const uint2*>(elem_keep + (plane + s) * 8);
      k1 = *reinterpret_cast<const uint2*>(elem_keep + (plane + s + NT) * 8);
    }
    accumulate_one(y0, d0, k0);
    accumulate_one(y1, d1, k1);
  }
  for (; s < s1; s += NT) {
    uint2 k0 = make_uint2(0, 0);
    if (elem_keep) k0 = *reinterpret_cast<const uint2*>(elem_keep + (plane + s) * 8);
    accumulate_one(ldg_nc_u4(y + plane + s), ldg_nc_u4(da + plane + s), k0);
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
__device__ __noinline__ void bn_bwd_finalize_block(const float* __restrict__ partial, float* __restrict__ sums,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                      int N, int C, long long S, int chunks, int spg, int accumulate) {
  const int Cb = (C + 7) / 8, G = N / spg;
  const double M = (double)spg * (double)S;
  __shared__ double sm_a[FIN_PAIRS], sm_b[FIN_PAIRS];
  int tile_cb = FIN_PAIRS / (8 * G);
  if (tile_cb < 1) tile_cb = 1;
  for (int cb0 = 0; cb0 < Cb; cb0 += tile_cb) {
    const int ncb = min(tile_cb, Cb - cb0);
    __syncthreads();
    fin_reduce_tile(partial, sm_a, sm_b, G, Cb, cb0, ncb, chunks, spg);
    __syncthreads();
    for (int t = threadIdx.x; t < ncb * 8; t += blockDim.x) {
      const int c = cb0 * 8 + t;
      if (c >= C) continue;
      double tg = 0.0, tb = 0.0;
      for (int g = 0; g < G; ++g) {
        const double a = sm_a[g * ncb * 8 + t], b = sm_b[g * ncb * 8 + t];
        sums[((long long)g * C + c) * 2 + 0] = (float)(a / M);
        sums[((long long)g * C + c) * 2 + 1] = (float)(b / M);
        tb += a; tg += b;
      }
      if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)tg : (float)tg;
      if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)tb : (float)tb;
    }
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

// Small layers: reduce + apply of the backward in one launch, one block per channel octet (see bn_stats_small_kernel).
__global__ void __launch_bounds__(NT) bn_bwd_small_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y,
                                                           uint4* __restrict__ dy, const float* __restrict__ stat,
                                                           const float* __restrict__ coef, const float* __restrict__ chan_scale,
                                                           const unsigned char* __restrict__ elem_keep, float elem_scale,
                                                           float* __restrict__ sums, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int N, int C, long long S, int spg,
                                                           float slope, int stats_grad, int reduce, int accumulate) {
  const int cb = blockIdx.x, Cb = gridDim.x, G = N / spg;
  __shared__ float red[16 * (NT / 32)];
  __shared__ float sc[8], sh[8], mu[8], is[8], m1[8], m2[8];
  const double M = (double)spg * (double)S;
  const int c = cb * 8 + (int)threadIdx.x;
  const bool owner = threadIdx.x < 8 && c < C;
  double tg = 0.0, tb = 0.0;
  for (int g = 0; g < G; ++g) {
    __syncthreads();
    if (threadIdx.x < 8) {
      sc[threadIdx.x] = owner ? coef[((long long)g * C + c) * 2] : 0.f;
      sh[threadIdx.x] = owner ? coef[((long long)g * C + c) * 2 + 1] : 0.f;
      mu[threadIdx.x] = owner ? stat[((long long)g * C + c) * 2] : 0.f;
      is[threadIdx.x] = owner ? stat[((long long)g * C + c) * 2 + 1] : 0.f;
      m1[threadIdx.x] = 0.f; m2[threadIdx.x] = 0.f;
    }
    __syncthreads();
    const long long total = (long long)spg * S;
    auto grad_of = [&](long long i, float* gk, float* xh) {
      const int n = g * spg + (int)(i / S);
      const long long off = ((long long)n * Cb + cb) * S + (i - (i / S) * S);
      float fy[8], fd[8];
      unpack8(ldg_nc_u4(y + off), fy);
      unpack8(ldg_nc_u4(da + off), fd);
      unsigned char kb[8];
      if (elem_keep) *reinterpret_cast<uint2*>(kb) = *reinterpret_cast<const uint2*>(elem_keep + off * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float pre = fy[k] * sc[k] + sh[k];
        float v = fd[k] * ((chan_scale && cb * 8 + k < C) ? chan_scale[(long long)n * C + cb * 8 + k] : 1.f);
        if (elem_keep) v = kb[k] ? v * elem_scale : 0.f;
        gk[k] = pre > 0.f ? v : v * slope;
        xh[k] = (fy[k] - mu[k]) * is[k];
      }
      return off;
    };
    if (reduce) {
      float acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = 0.f;
      for (long long i = threadIdx.x; i < total; i += NT) {
        float gk[8], xh[8];
        grad_of(i, gk, xh);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] += gk[k]; acc[8 + k] += gk[k] * xh[k]; }
      }
      block_sum<16, NT>(acc, red);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { m1[k] = (float)((double)acc[k] / M); m2[k] = (float)((double)acc[8 + k] / M); }
#pragma unroll
        for (int k = 0; k < 16; ++k) red[k] = acc[k];
      }
      __syncthreads();
      if (owner) {
        sums[((long long)g * C + c) * 2 + 0] = m1[threadIdx.x];
        sums[((long long)g * C + c) * 2 + 1] = m2[threadIdx.x];
        tb += (double)red[threadIdx.x];
        tg += (double)red[8 + threadIdx.x];
      }
    }
    for (long long i = threadIdx.x; i < total; i += NT) {
      float gk[8], xh[8];
      const long long off = grad_of(i, gk, xh);
#pragma unroll
      for (int k = 0; k < 8; ++k) gk[k] = stats_grad ? sc[k] * (gk[k] - m1[k] - xh[k] * m2[k]) : sc[k] * gk[k];
      dy[off] = pack8(gk);
    }
  }
  if (owner && reduce) {
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)tg : (float)tg;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)tb : (float)tb;
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
  BCP_REQUIRE(n / spg <= 128, "norm_stats: more than 128 normalisation groups in one call (%d)", n / spg);
  const int Cb = (c + 7) / 8, chunks = pick_chunks(s, n * Cb);
  if ((long long)n * s <= SMALL_LIMIT) {
    bn_stats_small_kernel<<<Cb, NT, 0, stream>>>((const uint4*)y, s, gamma, beta, running_mean, running_var, num_batches_tracked,
                                                  stat, coef, n, c, spg, eps, momentum);
    return check_launch("norm_stats");
  }
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
  BCP_REQUIRE(n / spg <= 128, "norm_bwd: more than 128 normalisation groups in one call (%d)", n / spg);
  const int Cb = (c + 7) / 8, chunks = pick_chunks(s, n * Cb);
  if ((long long)n * s <= SMALL_LIMIT) {
    const int reduce = (stats_grad || dgamma || dbeta) ? 1 : 0;
    bn_bwd_small_kernel<<<Cb, NT, 0, stream>>>((const uint4*)dact, (const uint4*)y, (uint4*)dy, stat, coef, chan_scale, elem_keep,
                                                elem_scale, sums, dgamma, dbeta, n, c, s, spg, slope, stats_grad, reduce, accumulate);
    return check_launch("norm_bwd");
  }
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
