// Single-launch train-mode normalisation for the mid-size and deep layers (<= 32 Ki voxels per statistics group):
// batch statistics + finalize + normalise/activation/dropout/skip-add in ONE kernel, and likewise the whole backward
// (both reductions + the gradient) in one.  norm.cu's two-launch path (partials -> last-block finalize -> apply) costs
// ~35 us on an 8 MB tensor whose HBM/L2 traffic is worth ~3 us: two launches, a serial last-block tail and grids that are
// either too small to pull bandwidth or pay a cross-grid arrival counter.
//
// Here every (statistics group, channel octet) is owned by ONE THREAD-BLOCK CLUSTER of CS <= 8 CTAs:
//   pass 1  each CTA reduces its contiguous slice of the group's voxels (the slice stays in REGISTERS when it is at most
//           KEEP 16-byte voxels per thread, else it is re-read from L2 in pass 2);
//   exchange the CTAs' 16 partial sums meet through distributed shared memory: barrier.cluster, then every CTA reads all CS
//           partials in rank order and forms the double-precision totals itself (fixed order => deterministic, and
//           identical in every CTA of the cluster, so no broadcast is needed);
//   pass 2  normalise + activation (+dropout, +residual) / input gradient, written once.
// Cross-group work that must happen in call order (running-statistic updates, num_batches_tracked, d(gamma)/d(beta) summed
// over groups) is done by the last cluster to finish, from a small fp32 table in global memory, one channel per thread.
// The slicing depends only on (spg, S), never on how many groups ride in the launch, so a batched call of G groups is
// bit-identical to G separate calls (tests/test_gpu_networks.py::test_vnet_grouped_equals_two_calls).
//
// Reference modules: nn.BatchNorm3d/2d (networks/VNet.py:19, networks/unet.py:21,25), nn.InstanceNorm3d
// (pancreas/Vnet.py:25,49,76), ReLU / LeakyReLU(0.01), Dropout3d / Dropout.
// HBM-bound: algorithmic bytes fwd = 2*|y| (read once, write once), bwd = 3*|y|.
#include "common.cuh"
#include "../../include/bcp_b200.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace bcp {

constexpr int NF = 256;            // threads per CTA (one CTA per SM-slot: up to 255 registers hold the slice)
constexpr int KEEP_F = 16;         // forward: voxels (uint4) a thread may hold between the passes
constexpr int KEEP_B = 16;         // backward: voxels of y AND da per thread
constexpr long long FUSED_MAX_GROUP_VOX = 8LL * KEEP_F * NF;   // spg * S above this (32 Ki voxels) -> norm.cu's streaming path

struct FusedGeom {
  int CS;                          // CTAs per cluster
  long long per;                   // voxels per CTA slice
};

static inline FusedGeom fused_geom(long long total, int keep) {
  FusedGeom g;
  g.CS = 1;
  while (g.CS < 8 && (total + g.CS - 1) / g.CS > (long long)keep * NF) g.CS *= 2;
  g.per = (total + g.CS - 1) / g.CS;
  return g;
}

__device__ __forceinline__ float cluster_partial(cg::cluster_group& cl, float* local, int rank, int k) {
  return cl.map_shared_rank(local, rank)[k];
}

// true in every thread of the CTA for exactly one caller: the last of `total` arrivals (one per cluster).  Self-resetting.
__device__ __forceinline__ bool last_cluster_arrives(int* counter, int total) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(counter, 1);
    is_last = (prev == total - 1);
    if (is_last) *counter = 0;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last != 0;
}

struct FwdArgs {
  const uint4* y; uint4* out;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* nbt;
  float* stat; float* coef; float* gstat;      // gstat[g][C][2] = {mean, unbiased var} for the in-order running update
  int* counter;
  const float* chan_scale; const unsigned char* elem_keep; float elem_scale;
  const uint4* residual;
  int N, C, spg, G;
  long long S, per;
  float eps, momentum, slope;
};

template <bool CACHED>
__global__ void __launch_bounds__(NF, 1) bn_fused_fwd_kernel(const FwdArgs a) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank(), CS = (int)cl.num_blocks();
  const int g = blockIdx.y, cb = blockIdx.z, Cb = gridDim.z;
  const long long total = (long long)a.spg * a.S;
  const long long lo = (long long)rank * a.per, hi = min(total, lo + a.per);
  __shared__ float red[16 * (NF / 32)];
  __shared__ float part[16];
  __shared__ float sc[8], sh[8];

  uint4 keep[CACHED ? KEEP_F : 1];
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  const int S32 = (int)a.S;                                // spg * S <= 32 Ki: group-linear indices fit 32 bits
  const long long base_g = ((long long)g * a.spg * Cb + cb) * a.S, sample_stride = (long long)Cb * a.S;
  auto plane_off = [&](long long i) {                      // group-linear voxel index -> element offset in the CB8 tensor
    const int ii = (int)i, ns = ii / S32;
    return base_g + ns * sample_stride + (ii - ns * S32);
  };
  if (CACHED) {
#pragma unroll
    for (int u = 0; u < KEEP_F; ++u) {
      const long long i = lo + threadIdx.x + (long long)u * NF;
      keep[u] = (i < hi) ? ldg_nc_u4(a.y + plane_off(i)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < KEEP_F; ++u) {
      float f[8];
      unpack8(keep[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] += f[k] * f[k]; }
    }
  } else {
    long long i = lo + threadIdx.x;
    for (; i + 3 * NF < hi; i += 4 * NF) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcg(a.y + plane_off(i + u * NF));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] += f[k] * f[k]; }
      }
    }
    for (; i < hi; i += NF) {
      float f[8];
      unpack8(__ldcg(a.y + plane_off(i)), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] += f[k] * f[k]; }
    }
  }
  block_sum<16, NF>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) part[k] = acc[k];
  }
  cl.sync();                                               // every CTA's partial is visible cluster-wide
  if (threadIdx.x < 8) {
    const int k = threadIdx.x, c = cb * 8 + k;
    double s = 0.0, q = 0.0;
    for (int r = 0; r < CS; ++r) { s += (double)cluster_partial(cl, part, r, k); q += (double)cluster_partial(cl, part, r, 8 + k); }
    const double M = (double)total;
    const double mean = s / M;
    double var = q / M - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const bool ok = c < a.C;
    const float ga = (ok && a.gamma) ? a.gamma[c] : 1.f, be = (ok && a.beta) ? a.beta[c] : 0.f;
    const float scale = ga * invstd, shift = be - (float)mean * scale;
    sc[k] = ok ? scale : 0.f;
    sh[k] = ok ? shift : 0.f;
    if (rank == 0 && ok) {
      const long long o = ((long long)g * a.C + c) * 2;
      a.stat[o] = (float)mean; a.stat[o + 1] = invstd;
      a.coef[o] = scale; a.coef[o + 1] = shift;
      if (a.running_mean) {
        a.gstat[o] = (float)mean;
        a.gstat[o + 1] = (float)((M > 1.0) ? var * M / (M - 1.0) : var);
      }
    }
  }
  cl.sync();                                               // remote reads of `part` are done; sc/sh visible in this CTA

  // ---- pass 2: out = act(y*scale + shift) [* chan_scale] [* keep*elem_scale] [+ residual]
  auto apply_one = [&](long long i, const uint4& v) {
    const int ii = (int)i, ns = ii / S32;
    const int n = g * a.spg + ns;
    const long long off = base_g + ns * sample_stride + (ii - ns * S32);
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = f[k] * sc[k] + sh[k];
      t = t > 0.f ? t : t * a.slope;
      if (a.chan_scale) t *= (cb * 8 + k < a.C) ? __ldg(a.chan_scale + (long long)n * a.C + cb * 8 + k) : 1.f;
      f[k] = t;
    }
    if (a.elem_keep) {
      const uint2 kp = *reinterpret_cast<const uint2*>(a.elem_keep + off * 8);
      const unsigned char* kb = reinterpret_cast<const unsigned char*>(&kp);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = kb[k] ? f[k] * a.elem_scale : 0.f;
    }
    if (a.residual) {
      float r[8];
      unpack8(ldg_nc_u4(a.residual + off), r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
    a.out[off] = pack8(f);
  };
  if (CACHED) {
#pragma unroll
    for (int u = 0; u < KEEP_F; ++u) {
      const long long i = lo + threadIdx.x + (long long)u * NF;
      if (i < hi) apply_one(i, keep[u]);
    }
  } else {
    long long i = lo + threadIdx.x;
    for (; i + 3 * NF < hi; i += 4 * NF) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcg(a.y + plane_off(i + u * NF));
#pragma unroll
      for (int u = 0; u < 4; ++u) apply_one(i + u * NF, v[u]);
    }
    for (; i < hi; i += NF) apply_one(i, __ldcg(a.y + plane_off(i)));
  }

  // ---- running statistics: one in-order update over the groups, by the last cluster of the launch to get here
  if (a.running_mean != nullptr || a.nbt != nullptr) {
    if (rank == 0) {
      if (last_cluster_arrives(a.counter, a.G * Cb)) {
        if (threadIdx.x == 0 && a.nbt != nullptr) a.nbt[0] += a.G;
        if (a.running_mean != nullptr) {
          for (int c = threadIdx.x; c < a.C; c += NF) {
            float rm = a.running_mean[c], rv = a.running_var[c];
            for (int gg = 0; gg < a.G; ++gg) {
              const float m = __ldcg(a.gstat + ((long long)gg * a.C + c) * 2), u = __ldcg(a.gstat + ((long long)gg * a.C + c) * 2 + 1);
              rm = (1.f - a.momentum) * rm + a.momentum * m;
              rv = (1.f - a.momentum) * rv + a.momentum * u;
            }
            a.running_mean[c] = rm;
            a.running_var[c] = rv;
          }
        }
      }
    }
  }
}

struct BwdArgs {
  const uint4* da; const uint4* y; uint4* dy;
  const float* stat; const float* coef;
  const float* chan_scale; const unsigned char* elem_keep; float elem_scale;
  float* sums; float* dgamma; float* dbeta;
  int* counter;
  int N, C, spg, G, stats_grad, reduce, accumulate;
  long long S, per;
  float slope;
};

template <bool CACHED>
__global__ void __launch_bounds__(NF, 1) bn_fused_bwd_kernel(const BwdArgs a) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank(), CS = (int)cl.num_blocks();
  const int g = blockIdx.y, cb = blockIdx.z, Cb = gridDim.z;
  const long long total = (long long)a.spg * a.S;
  const long long lo = (long long)rank * a.per, hi = min(total, lo + a.per);
  __shared__ float red[16 * (NF / 32)];
  __shared__ float part[16];
  __shared__ float sc[8], sh[8], mu[8], is[8], m1[8], m2[8];
  if (threadIdx.x < 8) {
    const int c = cb * 8 + threadIdx.x;
    const bool ok = c < a.C;
    sc[threadIdx.x] = ok ? a.coef[((long long)g * a.C + c) * 2] : 0.f;
    sh[threadIdx.x] = ok ? a.coef[((long long)g * a.C + c) * 2 + 1] : 0.f;
    mu[threadIdx.x] = ok ? a.stat[((long long)g * a.C + c) * 2] : 0.f;
    is[threadIdx.x] = ok ? a.stat[((long long)g * a.C + c) * 2 + 1] : 0.f;
    m1[threadIdx.x] = 0.f; m2[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const int S32 = (int)a.S;
  const long long base_g = ((long long)g * a.spg * Cb + cb) * a.S, sample_stride = (long long)Cb * a.S;
  auto plane_off = [&](long long i, int& n) {
    const int ii = (int)i, ns = ii / S32;
    n = g * a.spg + ns;
    return base_g + ns * sample_stride + (ii - ns * S32);
  };
  // g_k = da * chan_scale * dropout * act'(pre),  xhat_k
  auto grad_of = [&](long long off, int n, const uint4& vy, const uint4& vd, float* gk, float* xh) {
    float fy[8], fd[8];
    unpack8(vy, fy);
    unpack8(vd, fd);
    unsigned char kb[8];
    if (a.elem_keep) *reinterpret_cast<uint2*>(kb) = *reinterpret_cast<const uint2*>(a.elem_keep + off * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float pre = fy[k] * sc[k] + sh[k];
      float v = fd[k];
      if (a.chan_scale) v *= (cb * 8 + k < a.C) ? __ldg(a.chan_scale + (long long)n * a.C + cb * 8 + k) : 1.f;
      if (a.elem_keep) v = kb[k] ? v * a.elem_scale : 0.f;
      gk[k] = pre > 0.f ? v : v * a.slope;
      xh[k] = (fy[k] - mu[k]) * is[k];
    }
  };
  uint4 ky[CACHED ? KEEP_B : 1], kd[CACHED ? KEEP_B : 1];
  if (CACHED) {
#pragma unroll
    for (int u = 0; u < KEEP_B; ++u) {
      const long long i = lo + threadIdx.x + (long long)u * NF;
      int n;
      if (i < hi) { const long long off = plane_off(i, n); ky[u] = ldg_nc_u4(a.y + off); kd[u] = ldg_nc_u4(a.da + off); }
      else { ky[u] = make_uint4(0, 0, 0, 0); kd[u] = make_uint4(0, 0, 0, 0); }
    }
  }
  if (a.reduce) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    if (CACHED) {
#pragma unroll
      for (int u = 0; u < KEEP_B; ++u) {
        const long long i = lo + threadIdx.x + (long long)u * NF;
        if (i < hi) {
          int n;
          const long long off = plane_off(i, n);
          float gk[8], xh[8];
          grad_of(off, n, ky[u], kd[u], gk, xh);
#pragma unroll
          for (int k = 0; k < 8; ++k) { acc[k] += gk[k]; acc[8 + k] += gk[k] * xh[k]; }
        }
      }
    } else {
      long long i = lo + threadIdx.x;
      for (; i + NF < hi; i += 2 * NF) {
        int n0, n1;
        const long long o0 = plane_off(i, n0), o1 = plane_off(i + NF, n1);
        const uint4 y0 = __ldcg(a.y + o0), y1 = __ldcg(a.y + o1), d0 = __ldcg(a.da + o0), d1 = __ldcg(a.da + o1);
        float gk[8], xh[8];
        grad_of(o0, n0, y0, d0, gk, xh);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] += gk[k]; acc[8 + k] += gk[k] * xh[k]; }
        grad_of(o1, n1, y1, d1, gk, xh);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] += gk[k]; acc[8 + k] += gk[k] * xh[k]; }
      }
      for (; i < hi; i += NF) {
        int n;
        const long long off = plane_off(i, n);
        float gk[8], xh[8];
        grad_of(off, n, __ldcg(a.y + off), __ldcg(a.da + off), gk, xh);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] += gk[k]; acc[8 + k] += gk[k] * xh[k]; }
      }
    }
    block_sum<16, NF>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 16; ++k) part[k] = acc[k];
    }
    cl.sync();
    if (threadIdx.x < 8) {
      const int k = threadIdx.x, c = cb * 8 + k;
      double s1 = 0.0, s2 = 0.0;
      for (int r = 0; r < CS; ++r) { s1 += (double)cluster_partial(cl, part, r, k); s2 += (double)cluster_partial(cl, part, r, 8 + k); }
      const double M = (double)total;
      const float a1 = (float)(s1 / M), a2 = (float)(s2 / M);
      m1[k] = a1; m2[k] = a2;
      if (rank == 0 && c < a.C) {
        a.sums[((long long)g * a.C + c) * 2 + 0] = a1;
        a.sums[((long long)g * a.C + c) * 2 + 1] = a2;
      }
    }
    cl.sync();
  }
  auto write_one = [&](long long off, int n, const uint4& vy, const uint4& vd) {
    float gk[8], xh[8];
    grad_of(off, n, vy, vd, gk, xh);
#pragma unroll
    for (int k = 0; k < 8; ++k) gk[k] = a.stats_grad ? sc[k] * (gk[k] - m1[k] - xh[k] * m2[k]) : sc[k] * gk[k];
    a.dy[off] = pack8(gk);
  };
  if (CACHED) {
#pragma unroll
    for (int u = 0; u < KEEP_B; ++u) {
      const long long i = lo + threadIdx.x + (long long)u * NF;
      if (i < hi) { int n; const long long off = plane_off(i, n); write_one(off, n, ky[u], kd[u]); }
    }
  } else {
    long long i = lo + threadIdx.x;
    for (; i + NF < hi; i += 2 * NF) {
      int n0, n1;
      const long long o0 = plane_off(i, n0), o1 = plane_off(i + NF, n1);
      const uint4 y0 = __ldcg(a.y + o0), y1 = __ldcg(a.y + o1), d0 = __ldcg(a.da + o0), d1 = __ldcg(a.da + o1);
      write_one(o0, n0, y0, d0);
      write_one(o1, n1, y1, d1);
    }
    for (; i < hi; i += NF) { int n; const long long off = plane_off(i, n); write_one(off, n, __ldcg(a.y + off), __ldcg(a.da + off)); }
  }
  // ---- d(gamma), d(beta): sums over the groups, by the last cluster to finish (from the per-group means in `sums`)
  if (a.reduce && (a.dgamma != nullptr || a.dbeta != nullptr)) {
    if (rank == 0) {
      if (last_cluster_arrives(a.counter, a.G * Cb)) {
        const double M = (double)total;
        for (int c = threadIdx.x; c < a.C; c += NF) {
          double tb = 0.0, tg = 0.0;
          for (int gg = 0; gg < a.G; ++gg) {
            tb += (double)__ldcg(a.sums + ((long long)gg * a.C + c) * 2) * M;
            tg += (double)__ldcg(a.sums + ((long long)gg * a.C + c) * 2 + 1) * M;
          }
          if (a.dgamma) a.dgamma[c] = a.accumulate ? a.dgamma[c] + (float)tg : (float)tg;
          if (a.dbeta) a.dbeta[c] = a.accumulate ? a.dbeta[c] + (float)tb : (float)tb;
        }
      }
    }
  }
}

template <typename Kern, typename Args>
static int launch_cluster(Kern kern, const Args& args, int CS, int G, int Cb, cudaStream_t stream, const char* what) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)CS, (unsigned)G, (unsigned)Cb);
  cfg.blockDim = dim3(NF, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args);
  if (e != cudaSuccess) { set_last_error("%s: cluster launch failed: %s", what, cudaGetErrorString(e)); cudaGetLastError(); return BCP_ERR_CUDA; }
  return check_launch(what);
}

}  // namespace bcp

using namespace bcp;

extern "C" {

int bcp_norm_fused_supported(int n, int c, long long s, int spg) {
  if (n <= 0 || c <= 0 || s <= 0 || spg <= 0 || n % spg) return 0;
  if (n / spg > 65535 || (c + 7) / 8 > 65535) return 0;
  return (long long)spg * s <= FUSED_MAX_GROUP_VOX ? 1 : 0;
}

int bcp_norm_fused_fwd(const void* y, void* out, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float* stat, float* coef, float* workspace, int* counter,
                       const float* chan_scale, const unsigned char* elem_keep, float elem_scale, const void* residual,
                       int n, int c, long long s, int spg, float eps, float momentum, float slope, cudaStream_t stream) {
  BCP_REQUIRE(y && out && stat && coef && workspace && counter, "norm_fused_fwd: null pointer");
  BCP_REQUIRE(bcp_norm_fused_supported(n, c, s, spg), "norm_fused_fwd: shape not eligible (n=%d c=%d s=%lld spg=%d)", n, c, s, spg);
  BCP_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "norm_fused_fwd: running_mean / running_var must come together");
  const int Cb = (c + 7) / 8, G = n / spg;
  const FusedGeom geo = fused_geom((long long)spg * s, KEEP_F);
  FwdArgs a{(const uint4*)y, (uint4*)out, gamma, beta, running_mean, running_var, num_batches_tracked, stat, coef, workspace, counter,
            chan_scale, elem_keep, elem_scale, (const uint4*)residual, n, c, spg, G, s, geo.per, eps, momentum, slope};
  if (geo.per <= (long long)KEEP_F * NF) return launch_cluster(bn_fused_fwd_kernel<true>, a, geo.CS, G, Cb, stream, "norm_fused_fwd");
  return launch_cluster(bn_fused_fwd_kernel<false>, a, geo.CS, G, Cb, stream, "norm_fused_fwd");
}

int bcp_norm_fused_bwd(const void* dact, const void* y, void* dy, const float* stat, const float* coef, const float* chan_scale,
                       const unsigned char* elem_keep, float elem_scale, float* dgamma, float* dbeta, float* sums, int* counter,
                       int n, int c, long long s, int spg, float slope, int stats_grad, int accumulate, cudaStream_t stream) {
  BCP_REQUIRE(dact && y && dy && stat && coef && sums && counter, "norm_fused_bwd: null pointer");
  BCP_REQUIRE(bcp_norm_fused_supported(n, c, s, spg), "norm_fused_bwd: shape not eligible (n=%d c=%d s=%lld spg=%d)", n, c, s, spg);
  const int Cb = (c + 7) / 8, G = n / spg;
  const FusedGeom geo = fused_geom((long long)spg * s, KEEP_B);
  const int reduce = (stats_grad || dgamma || dbeta) ? 1 : 0;
  BwdArgs a{(const uint4*)dact, (const uint4*)y, (uint4*)dy, stat, coef, chan_scale, elem_keep, elem_scale, sums, dgamma, dbeta, counter,
            n, c, spg, G, stats_grad, reduce, accumulate, s, geo.per, slope};
  if (geo.per <= (long long)KEEP_B * NF) return launch_cluster(bn_fused_bwd_kernel<true>, a, geo.CS, G, Cb, stream, "norm_fused_bwd");
  return launch_cluster(bn_fused_bwd_kernel<false>, a, geo.CS, G, Cb, stream, "norm_fused_bwd");
}

}  // extern "C"
