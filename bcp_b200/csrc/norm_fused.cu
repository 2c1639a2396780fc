// Single-launch train-mode normalisation for the mid-size and deep layers (<= 64 Ki voxels per statistics group):
// batch statistics + finalize + normalise/activation/dropout/skip-add in ONE kernel, and likewise the whole backward
// (both reductions + the gradient) in one.  norm.cu's two-launch path (partials -> last-block finalize -> apply) costs
// ~35 us on an 8 MB tensor whose HBM/L2 traffic is worth ~3 us: two launches, a serial last-block tail and grids that are
// either too small to pull bandwidth or pay a cross-grid arrival counter.
//
// Here every (statistics group, channel octet) is owned by ONE THREAD-BLOCK CLUSTER of CS <= 8 CTAs:
//   pass 1  each CTA reduces its contiguous slice of the group's voxels; a slice of <= KEEP 16-byte voxels per thread (about
//           one at the usual sizes) stays in REGISTERS for pass 2, a larger one is re-read from L2;
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

constexpr int NF = 1024;           // threads per CTA: the kernels are instruction-issue-bound, not byte-bound (an 8 MB layer is
                                   // ~3 us of HBM time), so they run 32 warps per SM and keep the per-voxel path lean
constexpr int KEEP = 4;            // voxels (16 B each) a thread holds in registers between the two passes
constexpr long long FUSED_MAX_GROUP_VOX = 65536;    // spg * S above this -> norm.cu's multi-launch path (measured: no gain beyond)

#ifdef BCP_NORM_PROBE
__device__ unsigned long long g_norm_probe[16];
__device__ __forceinline__ void probe(int slot) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_norm_probe[slot] = t;
  }
}
#define PROBE(i) probe(i)
#else
#define PROBE(i)
#endif

struct FusedGeom {
  int CS;                          // CTAs per cluster
  int per;                         // voxels per CTA slice
};

// About one voxel per thread, at most 8 CTAs per cluster -- and no more clusters than the GPU keeps co-resident: a CTA of
// 1024 threads owns an SM, and cudaOccupancyMaxActiveClusters (tools/micro/norm_probe.cu) gives 15 clusters of 8, 33 of 4,
// 74 of 2 on the B200; a second wave would double the kernel's latency.  The choice must not depend on how many groups ride in
// the launch (a batched call has to slice -- and therefore round -- exactly like separate calls), so the cluster count is
// taken as that of the usual two-group launch, 2 * Cb.
static inline FusedGeom fused_geom(long long total, int Cb) {
  FusedGeom g;
  const int nclusters = 2 * Cb;
  const int cap = nclusters <= 15 ? 8 : nclusters <= 33 ? 4 : nclusters <= 74 ? 2 : 1;
  g.CS = 1;
  while (g.CS < cap && (total + g.CS - 1) / g.CS > NF) g.CS *= 2;
  g.per = (int)((total + g.CS - 1) / g.CS);
  return g;
}

// One step of a transpose-reduce: NV live values per lane -> NV/2, lanes exchanging with lane ^ OFF.  After the steps
// (16,16),(8,8),(4,4),(2,2) lane l holds, in v[0], the sum of value (l & 15) over the 16 lanes that share l >> 4 ... see below.
template <int NV, int OFF>
__device__ __forceinline__ void treduce_step(float (&v)[16], int lane) {
  const bool up = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) {
    const float send = up ? v[i] : v[i + NV / 2];
    const float keep = up ? v[i + NV / 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

// Deterministic CTA-wide sum of 16 floats per thread (1024 threads).  Warp level: a transpose-reduce (8+4+2+1 shuffles
// leave value index k(lane) summed over 16 lanes, one more shuffle folds the two lane halves: 16 shuffles instead of the
// 80 of sixteen butterflies); then warp 0 combines the 32 warp rows.  Result: part[0..15].
__device__ __forceinline__ void cta_sum16(float (&acc)[16], float* red /* 32*16 */, float* part /* 16 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  treduce_step<16, 16>(acc, lane);      // lanes with bit 16 set keep values 8..15
  treduce_step<8, 8>(acc, lane);
  treduce_step<4, 4>(acc, lane);
  treduce_step<2, 2>(acc, lane);
  // lane now holds value k = 8*b16 + 4*b8 + 2*b4 + b2 (bits of the lane index) summed over lane pairs differing in bit 1
  const float tot = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 1);
  if ((lane & 1) == 0) {
    const int k = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    red[warp * 16 + k] = tot;
  }
  __syncthreads();
  if (warp == 0) {
    const int k = lane & 15, half = lane >> 4;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) s += red[(half * 16 + w) * 16 + k];
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (lane < 16) part[k] = s;
  }
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
  int N, C, spg, G, S, per;
  float eps, momentum, slope;
};

// EXTRA: the layer has a channel-dropout scale, an element-dropout mask or a skip tensor (kept out of the common path)
template <bool EXTRA, int KP>
__global__ void __launch_bounds__(NF, 1) bn_fused_fwd_kernel(const FwdArgs a) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank(), CS = (int)cl.num_blocks();
  const int g = blockIdx.y, cb = blockIdx.z, Cb = gridDim.z;
  const int total = a.spg * a.S;
  const int lo = rank * a.per, hi = min(total, lo + a.per);
  __shared__ float red[16 * (NF / 32)];
  __shared__ float part[16];
  __shared__ float sc[8], sh[8];
  PROBE(0);
  float ga = 1.f, be = 0.f;                                // affine parameters of the channel thread k < 8 finalises (prefetched)
  if (threadIdx.x < 8 && cb * 8 + (int)threadIdx.x < a.C) {
    if (a.gamma) ga = __ldg(a.gamma + cb * 8 + threadIdx.x);
    if (a.beta) be = __ldg(a.beta + cb * 8 + threadIdx.x);
  }
  // this thread's voxels: group-linear index i = lo + tid + u*NF  ->  element offset relative to the group's first plane
  const long long base = ((long long)g * a.spg * Cb + cb) * a.S;
  const unsigned sample_stride = (unsigned)Cb * (unsigned)a.S;
  auto rel_of = [&](int i) { const int ns = i / a.S; return (unsigned)ns * sample_stride + (unsigned)(i - ns * a.S); };
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  auto stat_acc = [&](const uint4& v) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] = fmaf(f[k], f[k], acc[8 + k]); }
  };
  unsigned rel[KP > 0 ? KP : 1];
  uint4 keep[KP > 0 ? KP : 1];
  if constexpr (KP > 0) {
#pragma unroll
    for (int u = 0; u < KP; ++u) {
      const int i = lo + (int)threadIdx.x + u * NF;
      rel[u] = rel_of(i);
      keep[u] = (i < hi) ? ldg_nc_u4(a.y + base + rel[u]) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < KP; ++u) stat_acc(keep[u]);
  } else {                                                 // streaming: the slice is re-read (from L2) in pass 2
#pragma unroll 4
    for (int i = lo + (int)threadIdx.x; i < hi; i += NF) stat_acc(__ldcg(a.y + base + rel_of(i)));
  }
  PROBE(1);
  cta_sum16(acc, red, part);
  PROBE(2);
  cl.sync();                                               // every CTA's partial is visible cluster-wide
  PROBE(3);
  if (threadIdx.x < 8) {
    const int k = threadIdx.x, c = cb * 8 + k;
    double s = 0.0, q = 0.0;
    for (int r = 0; r < CS; ++r) {
      const float* rp = cl.map_shared_rank(part, r);
      s += (double)rp[k];
      q += (double)rp[8 + k];
    }
    const double M = (double)total;
    const double mean = s / M;
    double var = q / M - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const bool ok = c < a.C;
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
  PROBE(4);
  cl.sync();                                               // remote reads of `part` are done; sc/sh visible in this CTA
  PROBE(5);
  // ---- pass 2: out = act(y*scale + shift) [* chan_scale] [* keep*elem_scale] [+ residual]
  float scr[8], shr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { scr[k] = sc[k]; shr[k] = sh[k]; }
  auto apply_store = [&](int i, unsigned r, const uint4& v) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = fmaf(f[k], scr[k], shr[k]);
      f[k] = t > 0.f ? t : t * a.slope;
    }
    const long long off = base + r;
    if (EXTRA) {
      if (a.chan_scale) {
        const int n = g * a.spg + i / a.S;
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] *= (cb * 8 + k < a.C) ? __ldg(a.chan_scale + (long long)n * a.C + cb * 8 + k) : 1.f;
      }
      if (a.elem_keep) {
        const uint2 kp = *reinterpret_cast<const uint2*>(a.elem_keep + off * 8);
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&kp);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = kb[k] ? f[k] * a.elem_scale : 0.f;
      }
      if (a.residual) {
        float rr[8];
        unpack8(ldg_nc_u4(a.residual + off), rr);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += rr[k];
      }
    }
    a.out[off] = pack8(f);
  };
  if constexpr (KP > 0) {
#pragma unroll
    for (int u = 0; u < KP; ++u) {
      const int i = lo + (int)threadIdx.x + u * NF;
      if (i < hi) apply_store(i, rel[u], keep[u]);
    }
  } else {
#pragma unroll 4
    for (int i = lo + (int)threadIdx.x; i < hi; i += NF) {
      const unsigned r = rel_of(i);
      apply_store(i, r, __ldcg(a.y + base + r));
    }
  }
  PROBE(6);
  // ---- running statistics: one in-order update over the groups, by the last cluster of the launch to get here
  if ((a.running_mean != nullptr || a.nbt != nullptr) && rank == 0) {
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
  PROBE(7);
}

struct BwdArgs {
  const uint4* da; const uint4* y; uint4* dy;
  const float* stat; const float* coef;
  const float* chan_scale; const unsigned char* elem_keep; float elem_scale;
  float* sums; float* dgamma; float* dbeta;
  int* counter;
  int N, C, spg, G, stats_grad, reduce, accumulate, S, per;
  float slope;
};

template <bool EXTRA, int KP>
__global__ void __launch_bounds__(NF, 1) bn_fused_bwd_kernel(const BwdArgs a) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank(), CS = (int)cl.num_blocks();
  const int g = blockIdx.y, cb = blockIdx.z, Cb = gridDim.z;
  const int total = a.spg * a.S;
  const int lo = rank * a.per, hi = min(total, lo + a.per);
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
  const long long base = ((long long)g * a.spg * Cb + cb) * a.S;
  const unsigned sample_stride = (unsigned)Cb * (unsigned)a.S;
  auto rel_of = [&](int i) { const int ns = i / a.S; return (unsigned)ns * sample_stride + (unsigned)(i - ns * a.S); };
  unsigned rel[KP > 0 ? KP : 1];
  uint4 ky[KP > 0 ? KP : 1], kd[KP > 0 ? KP : 1];
  if constexpr (KP > 0) {
#pragma unroll
    for (int u = 0; u < KP; ++u) {
      const int i = lo + (int)threadIdx.x + u * NF;
      rel[u] = rel_of(i);
      if (i < hi) { ky[u] = ldg_nc_u4(a.y + base + rel[u]); kd[u] = ldg_nc_u4(a.da + base + rel[u]); }
      else { ky[u] = make_uint4(0, 0, 0, 0); kd[u] = make_uint4(0, 0, 0, 0); }
    }
  }
  __syncthreads();
  // g = da * chan_scale * dropout * act'(pre); formed twice (pass 1 for the sums, pass 2 for dy) from the same operands
  auto grad8 = [&](int i, unsigned r, const uint4& vy, const uint4& vd, float* gk, float* fy) {
    float fd[8];
    unpack8(vy, fy);
    unpack8(vd, fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float pre = fmaf(fy[k], sc[k], sh[k]);
      gk[k] = pre > 0.f ? fd[k] : fd[k] * a.slope;
    }
    if (EXTRA) {
      if (a.chan_scale) {
        const int n = g * a.spg + i / a.S;
#pragma unroll
        for (int k = 0; k < 8; ++k) gk[k] *= (cb * 8 + k < a.C) ? __ldg(a.chan_scale + (long long)n * a.C + cb * 8 + k) : 1.f;
      }
      if (a.elem_keep) {
        const uint2 kp = *reinterpret_cast<const uint2*>(a.elem_keep + (base + r) * 8);
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&kp);
#pragma unroll
        for (int k = 0; k < 8; ++k) gk[k] = kb[k] ? gk[k] * a.elem_scale : 0.f;
      }
    }
  };
  if (a.reduce) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    auto red_acc = [&](int i, unsigned r, const uint4& vy, const uint4& vd) {
      float gk[8], fy[8];
      grad8(i, r, vy, vd, gk, fy);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[k] += gk[k]; acc[8 + k] = fmaf(gk[k], (fy[k] - mu[k]) * is[k], acc[8 + k]); }
    };
    if constexpr (KP > 0) {
#pragma unroll
      for (int u = 0; u < KP; ++u) {
        const int i = lo + (int)threadIdx.x + u * NF;
        if (i < hi) red_acc(i, rel[u], ky[u], kd[u]);
      }
    } else {
#pragma unroll 2
      for (int i = lo + (int)threadIdx.x; i < hi; i += NF) {
        const unsigned r = rel_of(i);
        red_acc(i, r, __ldcg(a.y + base + r), __ldcg(a.da + base + r));
      }
    }
    cta_sum16(acc, red, part);
    cl.sync();
    if (threadIdx.x < 8) {
      const int k = threadIdx.x, c = cb * 8 + k;
      double s1 = 0.0, s2 = 0.0;
      for (int r = 0; r < CS; ++r) {
        const float* rp = cl.map_shared_rank(part, r);
        s1 += (double)rp[k];
        s2 += (double)rp[8 + k];
      }
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
  auto write_dy = [&](int i, unsigned r, const uint4& vy, const uint4& vd) {
    float gk[8], fy[8];
    grad8(i, r, vy, vd, gk, fy);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      gk[k] = a.stats_grad ? sc[k] * (gk[k] - m1[k] - ((fy[k] - mu[k]) * is[k]) * m2[k]) : sc[k] * gk[k];
    a.dy[base + r] = pack8(gk);
  };
  if constexpr (KP > 0) {
#pragma unroll
    for (int u = 0; u < KP; ++u) {
      const int i = lo + (int)threadIdx.x + u * NF;
      if (i < hi) write_dy(i, rel[u], ky[u], kd[u]);
    }
  } else {
#pragma unroll 2
    for (int i = lo + (int)threadIdx.x; i < hi; i += NF) {
      const unsigned r = rel_of(i);
      write_dy(i, r, __ldcg(a.y + base + r), __ldcg(a.da + base + r));
    }
  }
  // ---- d(gamma), d(beta): sums over the groups, by the last cluster to finish (from the per-group means in `sums`)
  if (a.reduce && (a.dgamma != nullptr || a.dbeta != nullptr) && rank == 0) {
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
  if ((long long)spg * s > FUSED_MAX_GROUP_VOX) return 0;
  return (long long)spg * ((c + 7) / 8) * s < (1LL << 31) ? 1 : 0;      // 32-bit offsets inside a group
}

int bcp_norm_fused_fwd(const void* y, void* out, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float* stat, float* coef, float* workspace, int* counter,
                       const float* chan_scale, const unsigned char* elem_keep, float elem_scale, const void* residual,
                       int n, int c, long long s, int spg, float eps, float momentum, float slope, cudaStream_t stream) {
  BCP_REQUIRE(y && out && stat && coef && workspace && counter, "norm_fused_fwd: null pointer");
  BCP_REQUIRE(bcp_norm_fused_supported(n, c, s, spg), "norm_fused_fwd: shape not eligible (n=%d c=%d s=%lld spg=%d)", n, c, s, spg);
  BCP_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "norm_fused_fwd: running_mean / running_var must come together");
  const int Cb = (c + 7) / 8, G = n / spg;
  const FusedGeom geo = fused_geom((long long)spg * s, Cb);
  FwdArgs a{(const uint4*)y, (uint4*)out, gamma, beta, running_mean, running_var, num_batches_tracked, stat, coef, workspace, counter,
            chan_scale, elem_keep, elem_scale, (const uint4*)residual, n, c, spg, G, (int)s, geo.per, eps, momentum, slope};
  // instance: 1 = one voxel per thread (lean), KEEP = the slice stays in registers, 0 = streaming (pass 2 re-reads from L2)
  const bool extra = chan_scale || elem_keep || residual;
  const int kp = geo.per <= NF ? 1 : geo.per <= KEEP * NF ? KEEP : 0;
#define BCP_FWD(E, K) launch_cluster(bn_fused_fwd_kernel<E, K>, a, geo.CS, G, Cb, stream, "norm_fused_fwd")
  if (extra) return kp == 1 ? BCP_FWD(true, 1) : kp == KEEP ? BCP_FWD(true, KEEP) : BCP_FWD(true, 0);
  return kp == 1 ? BCP_FWD(false, 1) : kp == KEEP ? BCP_FWD(false, KEEP) : BCP_FWD(false, 0);
#undef BCP_FWD
}

int bcp_norm_fused_bwd(const void* dact, const void* y, void* dy, const float* stat, const float* coef, const float* chan_scale,
                       const unsigned char* elem_keep, float elem_scale, float* dgamma, float* dbeta, float* sums, int* counter,
                       int n, int c, long long s, int spg, float slope, int stats_grad, int accumulate, cudaStream_t stream) {
  BCP_REQUIRE(dact && y && dy && stat && coef && sums && counter, "norm_fused_bwd: null pointer");
  BCP_REQUIRE(bcp_norm_fused_supported(n, c, s, spg), "norm_fused_bwd: shape not eligible (n=%d c=%d s=%lld spg=%d)", n, c, s, spg);
  const int Cb = (c + 7) / 8, G = n / spg;
  const FusedGeom geo = fused_geom((long long)spg * s, Cb);
  const int reduce = (stats_grad || dgamma || dbeta) ? 1 : 0;
  BwdArgs a{(const uint4*)dact, (const uint4*)y, (uint4*)dy, stat, coef, chan_scale, elem_keep, elem_scale, sums, dgamma, dbeta, counter,
            n, c, spg, G, stats_grad, reduce, accumulate, (int)s, geo.per, slope};
  // two operand tensors: only the one-voxel-per-thread instance keeps them in registers (64 registers per thread)
  const bool extra = chan_scale || elem_keep, one = geo.per <= NF;
#define BCP_BWD(E, K) launch_cluster(bn_fused_bwd_kernel<E, K>, a, geo.CS, G, Cb, stream, "norm_fused_bwd")
  if (extra) return one ? BCP_BWD(true, 1) : BCP_BWD(true, 0);
  return one ? BCP_BWD(false, 1) : BCP_BWD(false, 0);
#undef BCP_BWD
}

#ifdef BCP_NORM_PROBE
int bcp_norm_probe_read(unsigned long long* host16) {   // tools/micro/norm_probe.cu only
  return cudaMemcpyFromSymbol(host16, g_norm_probe, 16 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -1;
}
int bcp_norm_probe_max_clusters(int cs) {               // co-resident clusters of `cs` CTAs of the forward kernel
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cs, 2, 8);
  cfg.blockDim = dim3(NF, 1, 1);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = -1;
  cudaOccupancyMaxActiveClusters(&n, bn_fused_fwd_kernel<false, KEEP>, &cfg);
  return n;
}
#endif

}  // extern "C"
