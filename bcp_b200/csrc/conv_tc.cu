// tcgen05 + TMA implicit-GEMM convolution for sm_100a: 3x3x3 / 1x3x3, stride 1, zero 'same' padding, on
// channel-blocked bf16 activations (CB8).  Serves nn.Conv3d(k=3,p=1) / nn.Conv2d(k=3,p=1) forward
// (networks/VNet.py:17, networks/unet.py:20,24) and, with the flipped operand pack, their data gradient.
//
// Design ("halo brick + shifted views"):
//   * The output volume is cut into bricks of BX x BY x BZ voxels.  For a 16-channel chunk of the input, ONE
//     TMA box load brings the brick plus its one-voxel halo into shared memory (out-of-bounds = zero fill =
//     the convolution's zero padding) as two 8-channel planes of [HX*HY*HZ rows][8 ch] = 16 bytes per row.  The
//     tensor map describes the CB8 tensor as 8-byte elements with (z, 8 channels) merged into the innermost
//     dimension (encode_cb8), so a box row is a whole z-run, not a 16-byte voxel.
//   * That is exactly the UMMA "K-major, no swizzle" canonical layout with 16-byte rows (SBO = 128 B between
//     8-row groups, LBO = plane stride between the two K chunks).  A filter tap (dx,dy,dz) is therefore nothing
//     but a start-address offset of ((dx*HY+dy)*HZ+dz)*16 bytes in the A descriptor: all 27 taps are MMAs over
//     the SAME shared-memory bytes -- the halo is read from L2/HBM once, not 27 times.
//   * Output rows are the linearised halo'd brick; rows that fall on halo positions are computed and dropped
//     (efficiency BY*BZ/((BY+2)*(BZ+2)) per x-slab, chosen per layer by a small cost model).
//   * M = 128 rows per MMA, N = Cout (16..256), K = 16; accumulators live in TMEM (MT tiles of N columns,
//     double-buffered when 2*MT*N <= 512) and are drained by four epilogue warps with tcgen05.ld.
//   * Warp roles: warp 0 = TMA producer (A bricks through a ring of 16-channel slots; the weights of a chunk --
//     all taps, or one dx-group of nine -- as one TMA box per stage of a second ring), warp 1 = MMA issuer (ONE
//     elected thread runs the whole role: the tensor pipe is paced by that thread's instruction stream, see the
//     comment at the issuer and tools/micro/mma_bench.cu), warps 2..5 = epilogue (TMEM -> registers -> +bias ->
//     bf16 -> 16-byte coalesced stores; optionally the following normalisation's batch statistics).  All hand-offs
//     are mbarriers.
//   * Work item = (brick, slice of Cout/NS output channels); persistent grid, one CTA per SM, items assigned
//     round-robin (static => deterministic).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/bcp_b200.h"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace bcp {

struct TcParams {
  int N, X, Y, Z, Cin, Cout, kx;
  int BX, BY, BZ, HX, HY, HZ;
  int nbx, nby, nbz, nbricks;
  int rows_h;          // HX*HY*HZ rows delivered by TMA per 8-channel plane
  int MT;              // 128-row MMA tiles per brick
  int nchunks;         // Cin / 16
  int T;               // taps
  int SA, SB;          // ring depths
  int TG;              // taps per B stage (divides T)
  int NS, Ns;          // output-channel slices per brick (work item = brick x slice) and channels per slice
  int AS;              // accumulator stages in TMEM (1 or 2)
  int tmem_cols;       // power of two >= AS*MT*Cout
  unsigned slotA_bytes, stageB_bytes;
  unsigned offA, offB, offBar;   // shared-memory carve-up (bytes from the 128B-aligned base)
  int mergedA, mergedB;          // tensor maps use the 8-byte-element "merged inner dimension" form (tma_load_cb8)
  unsigned reserve, offStats;    // fused-BN-statistics table: bytes kept out of the operand rings / its offset
};

// Fused train-mode normalisation statistics (conv_tc_kernel<.., true>): the epilogue accumulates per-(group, channel)
// sum and sum of squares of the bf16-ROUNDED outputs it stores, so the separate pass over y (norm.cu bn_stats_kernel)
// disappears.  A thread sums its (at most MT) rows of one item in fp32; every combination across lanes, items, warps and
// CTAs is double precision in a fixed order (deterministic run to run; a different brick tiling regroups only those
// short fp32 sums, i.e. moves the statistics by ~1e-8 relative).
struct StatsArgs {
  double* partial;               // [gridDim.x][G][Cout][2]
  int* counter;                  // zero on entry, reset by the last CTA
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* nbt;
  float* stat; float* coef;      // [G][Cout][2]: {mean, invstd}, {scale, shift}
  int spg, G;
  float eps, momentum;
};

constexpr int TC_THREADS = 192;

// one step of the transpose-reduce: NV live values per lane -> NV/2, lanes exchanging with lane ^ OFF.  After the five
// steps (32,16),(16,8),(8,4),(4,2),(2,1) lane l holds the warp-wide total of value l in acc[0] (fixed order, 31 shuffles).
template <int NV, int OFF>
__device__ __forceinline__ void xreduce_step(double (&acc)[32], int lane) {
  const bool up = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) {
    const double send = up ? acc[i] : acc[i + NV / 2];
    const double keep = up ? acc[i + NV / 2] : acc[i];
    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

// run by the last CTA of a fused-statistics launch: fixed-order reduce of the per-CTA partials, then the same
// double-precision mean / variance / running-statistics chain as norm.cu bn_finalize_block
__device__ void conv_stats_finalize(const StatsArgs& sa, int C, double M, unsigned char* smem) {
  const int G = sa.G, GC2 = G * C * 2, nct = gridDim.x;
  double* sc = reinterpret_cast<double*>(smem);          // scratch [G][C][2] in the (now idle) operand slots
  if (threadIdx.x == 0 && sa.nbt != nullptr) sa.nbt[0] += G;
  // column-parallel reduce of partial[nct][GC2]: thread = (segment of the CTA range, output column); consecutive threads
  // read consecutive doubles, each keeps several loads in flight; segments are combined in fixed order afterwards
  const int W = GC2 < TC_THREADS ? GC2 : TC_THREADS;
  const int SEG = TC_THREADS / W;
  const int seg = threadIdx.x / W, col = threadIdx.x - seg * W;
  const int per = (nct + SEG - 1) / SEG;
  for (int i0 = 0; i0 < GC2; i0 += W) {
    const int i = i0 + col;
    if (seg < SEG && i < GC2) {
      const int b0 = seg * per, b1 = (b0 + per < nct) ? b0 + per : nct;
      double v = 0.0;
#pragma unroll 8
      for (int b = b0; b < b1; ++b) v += sa.partial[(size_t)b * GC2 + i];
      sc[(size_t)seg * GC2 + i] = v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < GC2; i += TC_THREADS) {
    double v = sc[i];
    for (int k = 1; k < SEG; ++k) v += sc[(size_t)k * GC2 + i];
    sc[i] = v;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += TC_THREADS) {
    float rm = sa.running_mean ? sa.running_mean[c] : 0.f, rv = sa.running_var ? sa.running_var[c] : 0.f;
    for (int g = 0; g < G; ++g) {
      const double mean = sc[(g * C + c) * 2] / M;
      double var = sc[(g * C + c) * 2 + 1] / M - mean * mean;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)sa.eps));
      const float ga = sa.gamma ? sa.gamma[c] : 1.f, be = sa.beta ? sa.beta[c] : 0.f;
      const float scale = ga * invstd;
      sa.stat[((long long)g * C + c) * 2 + 0] = (float)mean;
      sa.stat[((long long)g * C + c) * 2 + 1] = invstd;
      sa.coef[((long long)g * C + c) * 2 + 0] = scale;
      sa.coef[((long long)g * C + c) * 2 + 1] = be - (float)mean * scale;
      if (sa.running_mean) {
        const double unbiased = (M > 1.0) ? var * M / (M - 1.0) : var;
        rm = (1.f - sa.momentum) * rm + sa.momentum * (float)mean;
        rv = (1.f - sa.momentum) * rv + sa.momentum * (float)unbiased;
      }
    }
    if (sa.running_mean) { sa.running_mean[c] = rm; sa.running_var[c] = rv; }
  }
}

template <bool PROF, bool STATS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_w,
               const float* __restrict__ bias, uint4* __restrict__ out, const TcParams p, unsigned long long* __restrict__ prof,
               const StatsArgs sa) {
  // PROF: per-CTA cycle counters for tools/debug_conv_tc.py --prof (never instantiated on the product path)
  long long t_start = 0, w0 = 0, w1 = 0, w2 = 0;
  if (PROF) t_start = clock64();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a_base = sbase + p.offA, b_base = sbase + p.offB, bar_base = sbase + p.offBar;
  // barrier slots (8 bytes each): full_a[SA], empty_a[SA], full_b[SB], empty_b[SB], tmem_full[AS], tmem_empty[AS]
  const uint32_t full_a = bar_base, empty_a = full_a + 8 * p.SA, full_b = empty_a + 8 * p.SA, empty_b = full_b + 8 * p.SB;
  const uint32_t tmem_full = empty_b + 8 * p.SB, tmem_empty = tmem_full + 8 * p.AS;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(full_a + 8 * i, 1); mbar_init(empty_a + 8 * i, 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 1); }
    for (int i = 0; i < p.AS; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (STATS) {            // per-epilogue-warp tables [4][G][Cout][2] of doubles
    double* tbl = reinterpret_cast<double*>(smem + p.offStats);
    for (int i = threadIdx.x; i < 4 * sa.G * p.Cout * 2; i += TC_THREADS) tbl[i] = 0.0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int Cib = p.Cin >> 3, Cob = p.Cout >> 3;
  const int bricks_per_n = p.nbx * p.nby * p.nbz;
  const int nitems = p.nbricks * p.NS;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform; one elected lane issues)
    {
      Ring ra, rb;                                   // (slot, phase) of the A / B rings
      const uint32_t a_bytes = (uint32_t)p.rows_h * 32u;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int brick = item / p.NS, n0 = (item - brick * p.NS) * p.Ns;
        const int n = brick / bricks_per_n;
        int r = brick - n * bricks_per_n;
        const int bz = r % p.nbz; r /= p.nbz;
        const int by = r % p.nby;
        const int bx = r / p.nby;
        const int x0 = bx * p.BX - (p.kx >> 1), y0 = by * p.BY - 1, z0 = bz * p.BZ - 1;
        for (int c = 0; c < p.nchunks; ++c) {
          const uint32_t sa = ra.s, pa = ra.ph;
          { long long t0 = PROF ? clock64() : 0; mbar_wait(empty_a + 8 * sa, pa ^ 1); if (PROF) w0 += clock64() - t0; }
          if (elect_one()) {
            mbar_expect_tx(full_a + 8 * sa, a_bytes);
            tma_load_cb8(a_base + sa * p.slotA_bytes, &tmap, full_a + 8 * sa, p.mergedA, z0, y0, x0, n * Cib + 2 * c);
          }
          __syncwarp();
          ra.advance(p.SA);
          for (int t = 0; t < p.T; t += p.TG) {
            const uint32_t sb = rb.s, pb = rb.ph;
            { long long t0 = PROF ? clock64() : 0; mbar_wait(empty_b + 8 * sb, pb ^ 1); if (PROF) w1 += clock64() - t0; }
            if (elect_one()) {
              // ONE 4-D TMA box {8, Ns channels, 2 planes, TG taps} of the operand pack [T][Cin/8][Cout][8] per stage
              // (per-tap bulk copies of 1-2 KB each made the TMA unit's per-operation cost the bottleneck on deep layers)
              mbar_expect_tx(full_b + 8 * sb, p.stageB_bytes);
              if (p.mergedB) tma_load_3d(b_base + sb * p.stageB_bytes, &tmap_w, full_b + 8 * sb, 2 * n0, 2 * c, t);
              else tma_load_4d(b_base + sb * p.stageB_bytes, &tmap_w, full_b + 8 * sb, 0, n0, 2 * c, t);
            }
            __syncwarp();
            rb.advance(p.SB);
          }
        }
      }
    }
    if (PROF && lane == 0) { prof[blockIdx.x * 16 + 1] = w0; prof[blockIdx.x * 16 + 2] = w1; }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: ONE elected thread runs the whole role
    // (waits, issue, commits).  Measured (tools/micro/mma_bench.cu): the tensor pipe retires a M=128,K=16 MMA every
    // max(N/2, 32+N/4) cycles regardless of accumulator reuse or operand alignment -- but only if the issuing thread
    // spends a handful of uniform-datapath instructions per MMA.  So: descriptor high words are constants, the nine
    // (dy,dz) taps of a dx-group are unrolled with immediate row offsets, and only 32-bit low words are updated.
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Ns >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t adesc0 = make_desc(0, (uint32_t)p.rows_h * 16u, 128u);
      const uint64_t bdesc0 = make_desc(0, (uint32_t)p.Ns * 16u, 128u);
      const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
      const uint32_t bstep = (uint32_t)p.Ns * 2u;                          // 16-byte units per tap inside a B stage
      const uint32_t HZ = (uint32_t)p.HZ, plane = (uint32_t)(p.HY * p.HZ);
      const int groups = p.TG / 9;                                          // dx-groups per B stage (TG is 9 or 27)
      Ring ra, rb, rt;
      uint32_t it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it, rt.advance(p.AS)) {
        const uint32_t as = rt.s, ap = rt.ph;
        { long long t0 = PROF ? clock64() : 0; mbar_wait(tmem_empty + 8 * as, ap ^ 1); if (PROF) w2 += clock64() - t0; }
        tc_fence_after();
        const uint32_t d0 = tmem_base + as * (uint32_t)(p.MT * p.Ns);
        for (int c = 0; c < p.nchunks; ++c) {
          const uint32_t sa = ra.s, pa = ra.ph;
          { long long t0 = PROF ? clock64() : 0; mbar_wait(full_a + 8 * sa, pa); if (PROF) w0 += clock64() - t0; }
          tc_fence_after();
          uint32_t a_dx = a_lo0 + ((a_base + sa * p.slotA_bytes) >> 4);   // advances one x-plane per dx-group
          for (int t = 0; t < p.T; t += p.TG) {
            const uint32_t sb = rb.s, pb = rb.ph;
            { long long t0 = PROF ? clock64() : 0; mbar_wait(full_b + 8 * sb, pb); if (PROF) w1 += clock64() - t0; }
            tc_fence_after();
            uint32_t b_g = b_lo0 + ((b_base + sb * p.stageB_bytes) >> 4);
            for (int g = 0; g < groups; ++g) {
              const uint32_t first = (uint32_t)(c | t | g);                 // 0 only for the very first tap of the item
              uint32_t d = d0, a_m = a_dx;
              for (int mt = 0; mt < p.MT; ++mt) {
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                  const uint32_t a_lo = a_m + (uint32_t)(k / 3) * HZ + (uint32_t)(k % 3);
                  const uint32_t b_lo = b_g + (uint32_t)k * bstep;
                  umma_bf16(d, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, k ? 1u : first);
                }
                d += (uint32_t)p.Ns;
                a_m += 128u;
              }
              a_dx += plane;
              b_g += 9u * bstep;
            }
            umma_commit(empty_b + 8 * sb);
            rb.advance(p.SB);
          }
          umma_commit(empty_a + 8 * sa);
          ra.advance(p.SA);
        }
        umma_commit(tmem_full + 8 * as);
      }
      if (PROF) { prof[blockIdx.x * 16 + 3] = w0; prof[blockIdx.x * 16 + 4] = w1; prof[blockIdx.x * 16 + 5] = w2; prof[blockIdx.x * 16 + 8] = clock64() - t_start; prof[blockIdx.x * 16 + 9] = it; }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    Ring rt;
    const long long S = (long long)p.X * p.Y * p.Z;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, rt.advance(p.AS)) {
      const int brick = item / p.NS, n0 = (item - brick * p.NS) * p.Ns;
      const int n = brick / bricks_per_n;
      int r = brick - n * bricks_per_n;
      const int bz = r % p.nbz; r /= p.nbz;
      const int by = r % p.nby;
      const int bx = r / p.nby;
      const uint32_t as = rt.s, ap = rt.ph;
      long long t_e0 = 0;
      { long long t0 = PROF ? clock64() : 0; mbar_wait(tmem_full + 8 * as, ap); if (PROF) { t_e0 = clock64(); w0 += t_e0 - t0; } }
      tc_fence_after();
      const uint32_t d0 = tmem_base + as * (uint32_t)(p.MT * p.Ns) + ((uint32_t)(q * 32) << 16);
      if (STATS) {
        // channel-group outer, tile inner: 32 double accumulators live per thread; one transpose-reduce per (item, group)
        double* tbl = reinterpret_cast<double*>(smem + p.offStats) + (size_t)q * sa.G * p.Cout * 2;
        const int g = n / sa.spg;
        for (int c16 = 0; c16 < p.Ns; c16 += 16) {
          float facc[32];              // fp32 over this thread's <= MT rows of the item (FP64 here cost more than the
#pragma unroll                         // separate statistics pass saved); everything across threads / items is double
          for (int k = 0; k < 32; ++k) facc[k] = 0.f;
          for (int mt = 0; mt < p.MT; ++mt) {
            const int L = mt * 128 + q * 32 + lane;
            const int iz = L % p.HZ;
            const int ry = L / p.HZ;
            const int iy = ry % p.HY, ix = ry / p.HY;
            const int x = bx * p.BX + ix, y = by * p.BY + iy, z = bz * p.BZ + iz;
            const bool valid = (ix < p.BX) && (iy < p.BY) && (iz < p.BZ) && (x < p.X) && (y < p.Y) && (z < p.Z);
            const long long sp = ((long long)x * p.Y + y) * p.Z + z;
            uint32_t v[16];
            tmem_ld16(d0 + (uint32_t)(mt * p.Ns + c16), v);
            tmem_ld_wait();
            if (valid) {
              float f[16];
#pragma unroll
              for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v[k]) + (bias ? __ldg(bias + n0 + c16 + k) : 0.f);
              const uint4 p0 = pack8(f), p1 = pack8(f + 8);
              uint4* dst = out + ((long long)n * Cob + ((n0 + c16) >> 3)) * S + sp;
              dst[0] = p0;
              dst[S] = p1;
              unpack8(p0, f);                    // the statistics are those of the stored (bf16-rounded) tensor
              unpack8(p1, f + 8);
#pragma unroll
              for (int k = 0; k < 16; ++k) { facc[k] += f[k]; facc[16 + k] = fmaf(f[k], f[k], facc[16 + k]); }
            }
          }
          double acc[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) acc[k] = (double)facc[k];
          xreduce_step<32, 16>(acc, lane);
          xreduce_step<16, 8>(acc, lane);
          xreduce_step<8, 4>(acc, lane);
          xreduce_step<4, 2>(acc, lane);
          xreduce_step<2, 1>(acc, lane);
          tbl[((size_t)g * p.Cout + n0 + c16 + (lane & 15)) * 2 + (lane >> 4)] += acc[0];
        }
      } else {
      for (int mt = 0; mt < p.MT; ++mt) {
        const int L = mt * 128 + q * 32 + lane;
        const int iz = L % p.HZ;
        const int ry = L / p.HZ;
        const int iy = ry % p.HY, ix = ry / p.HY;
        const int x = bx * p.BX + ix, y = by * p.BY + iy, z = bz * p.BZ + iz;
        const bool valid = (ix < p.BX) && (iy < p.BY) && (iz < p.BZ) && (x < p.X) && (y < p.Y) && (z < p.Z);
        const long long sp = ((long long)x * p.Y + y) * p.Z + z;
        for (int c16 = 0; c16 < p.Ns; c16 += 16) {
          uint32_t v[16];
          tmem_ld16(d0 + (uint32_t)(mt * p.Ns + c16), v);
          tmem_ld_wait();
          if (valid) {
            float f[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v[k]) + (bias ? __ldg(bias + n0 + c16 + k) : 0.f);
            uint4* dst = out + ((long long)n * Cob + ((n0 + c16) >> 3)) * S + sp;
            dst[0] = pack8(f);
            dst[S] = pack8(f + 8);
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + 8 * as);
      if (PROF) w1 += clock64() - t_e0;
    }
    if (PROF && warp == 2 && lane == 0) { prof[blockIdx.x * 16 + 6] = w0; prof[blockIdx.x * 16 + 7] = w1; }
  }

  tc_fence_before();
  __syncthreads();
  if (PROF && threadIdx.x == 0) prof[blockIdx.x * 16 + 0] = clock64() - t_start;
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
  if (STATS) {
    // this CTA's totals (the four epilogue warps' tables, fixed order) -> global; the last CTA to arrive finalises
    const int GC2 = sa.G * p.Cout * 2;
    const double* t0 = reinterpret_cast<const double*>(smem + p.offStats);
    for (int i = threadIdx.x; i < GC2; i += TC_THREADS)
      sa.partial[(size_t)blockIdx.x * GC2 + i] = ((t0[i] + t0[GC2 + i]) + t0[2 * GC2 + i]) + t0[3 * GC2 + i];
    if (last_block_arrives(sa.counter))
      conv_stats_finalize(sa, p.Cout, (double)sa.spg * (double)p.X * (double)p.Y * (double)p.Z, smem);
  }
}

// =========================================================================================================
// dz-folded forward / data-gradient kernel, the default for 16- and 32-channel layers (ops._TC_FOLD; parity:
// tests/test_gpu_primitives.py::test_conv_tc_fold_vs_torch, tests/test_gpu_conv_shapes.py): the same convolution with the
// three dz taps of a (dx,dy) pair folded into the MMA N dimension.
//   B tile of a group g=(dx,dy): [W(g,dz=0) | W(g,dz=1) | W(g,dz=2)]  -> N = 3*Ns, D'[row][dz*Ns + co]
//   A tile: 8-row core-matrix groups start every SIX rows (descriptor SBO = 96 B): tile row i is frame row
//           96*mt + 6*(i/8) + (i%8), so an 8-lane group of the epilogue holds rows r..r+7 and produces outputs r..r+5 as
//           out[r] = D'[r][0] + D'[r+1][1] + D'[r+2][2] with two intra-group shuffles (tools/model_dzfold.py checks the
//           index algebra).  9 MMAs of N = 3*Ns per 96 output rows instead of 27 of N = Ns per 128.
// Same roles / barriers as conv_tc_kernel; one weight stage = all taps of a 16-channel chunk laid out [g][plane][dz][Ns].
// =========================================================================================================
struct FoldParams {
  int N, X, Y, Z, Cin, Cout, kx;
  int BX, BY, BZ, HX, HY, HZ;
  int nbx, nby, nbz, nbricks;
  int rows_h, MT, nchunks, ngrp;   // MT = tiles of 96 output rows per brick, ngrp = (dx,dy) groups (9 or 3)
  int SA, SB, NS, Ns, AS, tmem_cols;
  unsigned slotA_bytes, stageB_bytes, offA, offB, offBar;
  int mergedA;
};

// EW = epilogue warps (multiple of 4: TMEM lane quarter = warp % 4), EW / 4 tile subsets.  The epilogue (three TMEM loads,
// 32 shuffles, bf16 pack, two 16-byte stores per 16 channels of a 96-row tile) is latency-bound per warp, so it takes
// several warps per scheduler to keep up with the MMA stream (9 MMAs of N = 3*Ns per tile).
constexpr int FOLD_EPI_WARPS_DEFAULT = 16;

template <bool PROF, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_tc_fold_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_w,
                    const float* __restrict__ bias, uint4* __restrict__ out, const FoldParams p,
                    unsigned long long* __restrict__ prof) {
  // PROF: per-CTA cycle counters (same slots as conv_tc_kernel<true, .>), tools/debug_conv_tc.py --prof-fold only
  long long t_start = 0, w0 = 0, w1 = 0, w2 = 0;
  if (PROF) t_start = clock64();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a_base = sbase + p.offA, b_base = sbase + p.offB, bar_base = sbase + p.offBar;
  const uint32_t full_a = bar_base, empty_a = full_a + 8 * p.SA, full_b = empty_a + 8 * p.SA, empty_b = full_b + 8 * p.SB;
  const uint32_t tmem_full = empty_b + 8 * p.SB, tmem_empty = tmem_full + 8 * p.AS;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(full_a + 8 * i, 1); mbar_init(empty_a + 8 * i, 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 1); }
    for (int i = 0; i < p.AS; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __shared__ float bias_s[32];                       // Ns <= 32 (NS = 1): the layer's bias, read as shared-memory broadcasts
  if (threadIdx.x < 32) bias_s[threadIdx.x] = (bias != nullptr && (int)threadIdx.x < p.Ns) ? __ldg(bias + threadIdx.x) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int Cib = p.Cin >> 3, Cob = p.Cout >> 3;
  const int bricks_per_n = p.nbx * p.nby * p.nbz;
  const int nitems = p.nbricks * p.NS;
  const int N3 = 3 * p.Ns;

  if (warp == 0) {
    Ring ra, rb;
    const uint32_t a_bytes = (uint32_t)p.rows_h * 32u;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int brick = item / p.NS, n0 = (item - brick * p.NS) * p.Ns;
      const int n = brick / bricks_per_n;
      int r = brick - n * bricks_per_n;
      const int bz = r % p.nbz; r /= p.nbz;
      const int by = r % p.nby;
      const int bx = r / p.nby;
      const int x0 = bx * p.BX - (p.kx >> 1), y0 = by * p.BY - 1, z0 = bz * p.BZ - 1;
      for (int c = 0; c < p.nchunks; ++c) {
        { long long t0 = PROF ? clock64() : 0; mbar_wait(empty_a + 8 * ra.s, ra.ph ^ 1); if (PROF) w0 += clock64() - t0; }
        { long long t0 = PROF ? clock64() : 0; mbar_wait(empty_b + 8 * rb.s, rb.ph ^ 1); if (PROF) w1 += clock64() - t0; }
        if (elect_one()) {
          mbar_expect_tx(full_a + 8 * ra.s, a_bytes);
          tma_load_cb8(a_base + ra.s * p.slotA_bytes, &tmap, full_a + 8 * ra.s, p.mergedA, z0, y0, x0, n * Cib + 2 * c);
          // weight box {Ns*2 (8-byte units), dz 3, plane 2, group ngrp} -> shared memory [g][plane][dz][Ns rows]
          mbar_expect_tx(full_b + 8 * rb.s, p.stageB_bytes);
          tma_load_4d(b_base + rb.s * p.stageB_bytes, &tmap_w, full_b + 8 * rb.s, 2 * n0, 0, 2 * c, 0);
        }
        __syncwarp();
        ra.advance(p.SA);
        rb.advance(p.SB);
      }
    }
    if (PROF && lane == 0) { prof[blockIdx.x * 16 + 1] = w0; prof[blockIdx.x * 16 + 2] = w1; }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N3 >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t adesc0 = make_desc(0, (uint32_t)p.rows_h * 16u, 96u);        // SBO = 96 B: row groups overlap by two rows
      const uint64_t bdesc0 = make_desc(0, (uint32_t)N3 * 16u, 128u);
      const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
      const uint32_t HZ = (uint32_t)p.HZ, plane = (uint32_t)(p.HY * p.HZ);
      const uint32_t bgroup = 2u * (uint32_t)N3;                                    // 16-byte units per (dx,dy) group: 2 planes x 3*Ns rows
      Ring ra, rb, rt;
      uint32_t it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, rt.advance(p.AS), ++it) {
        { long long t0 = PROF ? clock64() : 0; mbar_wait(tmem_empty + 8 * rt.s, rt.ph ^ 1); if (PROF) w2 += clock64() - t0; }
        tc_fence_after();
        const uint32_t d0 = tmem_base + rt.s * (uint32_t)(p.MT * N3);
        for (int c = 0; c < p.nchunks; ++c) {
          { long long t0 = PROF ? clock64() : 0; mbar_wait(full_a + 8 * ra.s, ra.ph); if (PROF) w0 += clock64() - t0; }
          { long long t0 = PROF ? clock64() : 0; mbar_wait(full_b + 8 * rb.s, rb.ph); if (PROF) w1 += clock64() - t0; }
          tc_fence_after();
          const uint32_t a_slot = a_lo0 + ((a_base + ra.s * p.slotA_bytes) >> 4);
          const uint32_t b_slot = b_lo0 + ((b_base + rb.s * p.stageB_bytes) >> 4);
          uint32_t d = d0, a_m = a_slot;
          for (int mt = 0; mt < p.MT; ++mt) {
#pragma unroll
            for (int g = 0; g < 9; ++g) {
              if (g < p.ngrp) {
                const uint32_t goff = (p.kx == 3) ? (uint32_t)(g / 3) * plane + (uint32_t)(g % 3) * HZ : (uint32_t)g * HZ;
                umma_bf16(d, ((uint64_t)a_hi << 32) | (a_m + goff), ((uint64_t)b_hi << 32) | (b_slot + (uint32_t)g * bgroup), idesc,
                          g ? 1u : (uint32_t)c);
              }
            }
            d += (uint32_t)N3;
            a_m += 96u;                                                               // next tile: 96 frame rows further
          }
          umma_commit(empty_a + 8 * ra.s);
          umma_commit(empty_b + 8 * rb.s);
          ra.advance(p.SA);
          rb.advance(p.SB);
        }
        umma_commit(tmem_full + 8 * rt.s);
      }
      if (PROF) { prof[blockIdx.x * 16 + 3] = w0; prof[blockIdx.x * 16 + 4] = w1; prof[blockIdx.x * 16 + 5] = w2; prof[blockIdx.x * 16 + 8] = clock64() - t_start; prof[blockIdx.x * 16 + 9] = it; }
    }
    __syncwarp();
  } else {
    // EW epilogue warps: warp w reads TMEM lane quarter w % 4 and takes the tiles mt = sub, sub + nsub, ...
    // The frame row of a lane advances by 96*nsub per step; (ix,iy,iz) follow incrementally (no per-row divisions).
    const int q = warp & 3, sub = (warp - 2) >> 2, nsub = EW / 4;

    Ring rt;
    const long long S = (long long)p.X * p.Y * p.Z;
    const int grp = q * 4 + (lane >> 3), k8 = lane & 7;                               // tile row i = 32q + lane = 8*grp + k8
    const int step = 96 * nsub;
    const int sz = step % p.HZ, sy = (step / p.HZ) % p.HY, sx = (step / p.HZ) / p.HY;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, rt.advance(p.AS)) {
      const int brick = item / p.NS, n0 = (item - brick * p.NS) * p.Ns;
      const int n = brick / bricks_per_n;
      int r = brick - n * bricks_per_n;
      const int bz = r % p.nbz; r /= p.nbz;
      const int by = r % p.nby;
      const int bx = r / p.nby;
      long long t_e0 = 0;
      { long long t0 = PROF ? clock64() : 0; mbar_wait(tmem_full + 8 * rt.s, rt.ph); if (PROF) { t_e0 = clock64(); w0 += t_e0 - t0; } }
      tc_fence_after();
      const uint32_t d0 = tmem_base + rt.s * (uint32_t)(p.MT * N3) + ((uint32_t)(q * 32) << 16);
      // Lean drain.  ncu (profiles/fold_c16_ncu_r02.txt) showed the epilogue warps latency-bound on short dependencies
      // (5.6 cycles per issued instruction per warp, two warps per scheduler) and spending a third of their instructions
      // on per-tile index arithmetic; software-pipelining the TMEM loads did not help.  So: many warps (EW = 16, four per
      // scheduler), few registers, 32-bit offsets, the row -> voxel map advanced incrementally, no division in the loop.
      const int L0 = sub * 96 + grp * 6 + k8;                                          // frame row of this lane in its first tile
      int iz = L0 % p.HZ, iy = (L0 / p.HZ) % p.HY, ix = (L0 / p.HZ) / p.HY;
      const int x0 = bx * p.BX, y0 = by * p.BY, z0 = bz * p.BZ;
      const int xlim = min(p.BX, p.X - x0), ylim = min(p.BY, p.Y - y0), zlim = min(p.BZ, p.Z - z0);
      uint4* const out_item = out + ((long long)n * Cob + (n0 >> 3)) * S + ((long long)x0 * p.Y + y0) * p.Z + z0;
      const uint32_t YZ = (uint32_t)(p.Y * p.Z), Zs = (uint32_t)p.Z, S32 = (uint32_t)S;
      uint32_t taddr = d0 + (uint32_t)(sub * N3);
      const uint32_t tstep = (uint32_t)(nsub * N3);
      for (int mt = sub; mt < p.MT; mt += nsub, taddr += tstep) {
        const bool valid = (k8 < 6) & (ix < xlim) & (iy < ylim) & (iz < zlim);
        const uint32_t sp = (uint32_t)ix * YZ + (uint32_t)iy * Zs + (uint32_t)iz;
        uint32_t ca = taddr;
        uint4* dst = out_item + sp;
        for (int c16 = 0; c16 < p.Ns; c16 += 16, ca += 16, dst += 2 * S32) {
          uint32_t v0[16], v1[16], v2[16];
          tmem_ld16(ca, v0);
          tmem_ld16(ca + (uint32_t)p.Ns, v1);
          tmem_ld16(ca + (uint32_t)(2 * p.Ns), v2);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v1[k]), 1, 8);   // D'[r+1][dz=1]
            const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[k]), 2, 8);   // D'[r+2][dz=2]
            f[k] = (__uint_as_float(v0[k]) + a1) + (a2 + bias_s[c16 + k]);
          }
          if (valid) {
            dst[0] = pack8(f);
            dst[S32] = pack8(f + 8);
          }
        }
        iz += sz;
        if (iz >= p.HZ) { iz -= p.HZ; ++iy; }
        iy += sy;
        if (iy >= p.HY) { iy -= p.HY; ++ix; }
        ix += sx;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + 8 * rt.s);
      if (PROF) w1 += clock64() - t_e0;
    }
    if (PROF && warp == 2 && lane == 0) { prof[blockIdx.x * 16 + 6] = w0; prof[blockIdx.x * 16 + 7] = w1; }
  }
  tc_fence_before();
  __syncthreads();
  if (PROF && threadIdx.x == 0) prof[blockIdx.x * 16 + 0] = clock64() - t_start;
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// =========================================================================================================
// stride-2 family on tensor cores ("tap GEMM"): 2x2x2 stride-2 conv (gather) and transposed conv (scatter).
//   rows = voxels of the HALF-resolution grid, compact per brick (no halo, no dropped rows except tile padding)
//   gather : out_half[o][co]        = sum_t sum_ci in_full[2o+t][ci] * W[t][ci][co]
//            A item = (16-ch chunk, tap): ONE TMA box load with elementStrides (1,2,2,2,1) starting at tap t picks
//            exactly the 2o+t voxels -> every input byte is read once; B item = W[t] chunk (kind-0 pack)
//   scatter: out_full[2i+t][co]     = sum_ci in_half[i][ci] * W[ci][t][co]
//            one GEMM with N = taps*Cout, cut into pieces of TPc taps; A item = 16-ch chunk; B item = 2 planes of the
//            kind-3 pack [Cin/8][T][Cout][8]; the epilogue scatters column group tl to voxel 2i + tap(t0+tl)
// Same warp roles / barriers as conv_tc_kernel.  Used for nn.Conv3d(k=2,s=2) fwd + nn.ConvTranspose3d dgrad (gather)
// and nn.ConvTranspose3d fwd + nn.Conv3d(k=2,s=2) dgrad (scatter)  (networks/VNet.py:74,101).
// =========================================================================================================
struct S2Params {
  int N, Cin, Cout, mode;          // mode 1 = gather, 2 = scatter
  int Xh, Yh, Zh;                  // half-resolution grid
  int BX, BY, BZ, nbx, nby, nbz, nbricks;
  int rows, MT, nchunks;
  int TPc, npieces, Npiece;        // taps per piece, pieces, MMA N
  int SA, SB, AS, tmem_cols;
  unsigned slotA_bytes, stageB_bytes, offA, offB, offBar;
  int mergedA;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_s2_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* __restrict__ wpack, const float* __restrict__ bias,
                  uint4* __restrict__ out, const S2Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a_base = sbase + p.offA, b_base = sbase + p.offB, bar_base = sbase + p.offBar;
  const uint32_t full_a = bar_base, empty_a = full_a + 8 * p.SA, full_b = empty_a + 8 * p.SA, empty_b = full_b + 8 * p.SB;
  const uint32_t tmem_full = empty_b + 8 * p.SB, tmem_empty = tmem_full + 8 * p.AS;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(full_a + 8 * i, 1); mbar_init(empty_a + 8 * i, 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(full_b + 8 * i, 1); mbar_init(empty_b + 8 * i, 1); }
    for (int i = 0; i < p.AS; ++i) { mbar_init(tmem_full + 8 * i, 1); mbar_init(tmem_empty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int Cib = p.Cin >> 3, Cob = p.Cout >> 3;
  const int bricks_per_n = p.nbx * p.nby * p.nbz;
  const int kitems = (p.mode == 1) ? p.nchunks * 8 : p.nchunks;     // A/B items per (brick, piece)
  const int ntiles = p.nbricks * p.npieces;

  if (warp == 0) {
    {
      Ring ra, rb;
      const uint32_t a_bytes = (uint32_t)p.rows * 32u;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int brick = tile / p.npieces, piece = tile - brick * p.npieces;
        const int n = brick / bricks_per_n;
        int r = brick - n * bricks_per_n;
        const int bz = r % p.nbz; r /= p.nbz;
        const int by = r % p.nby;
        const int bx = r / p.nby;
        for (int k = 0; k < kitems; ++k) {
          const int c = (p.mode == 1) ? (k >> 3) : k;
          const int t = (p.mode == 1) ? (k & 7) : 0;
          const uint32_t sa = ra.s, pa = ra.ph;
          mbar_wait(empty_a + 8 * sa, pa ^ 1);
          const uint32_t sb = rb.s, pb = rb.ph;
          mbar_wait(empty_b + 8 * sb, pb ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full_a + 8 * sa, a_bytes);
            if (p.mode == 1)
              tma_load_5d(a_base + sa * p.slotA_bytes, &tmap, full_a + 8 * sa, 0, 2 * bz * p.BZ + (t & 1), 2 * by * p.BY + ((t >> 1) & 1),
                          2 * bx * p.BX + (t >> 2), n * Cib + 2 * c);
            else
              tma_load_cb8(a_base + sa * p.slotA_bytes, &tmap, full_a + 8 * sa, p.mergedA, bz * p.BZ, by * p.BY, bx * p.BX, n * Cib + 2 * c);
            mbar_expect_tx(full_b + 8 * sb, p.stageB_bytes);
            if (p.mode == 1) {
              bulk_load(b_base + sb * p.stageB_bytes, wpack + ((size_t)t * Cib + 2 * c) * (size_t)p.Cout * 8, p.stageB_bytes, full_b + 8 * sb);
            } else {
              const uint32_t half = p.stageB_bytes / 2;
              const size_t t0 = (size_t)piece * p.TPc;
              bulk_load(b_base + sb * p.stageB_bytes, wpack + (((size_t)(2 * c) * 8 + t0) * p.Cout) * 8, half, full_b + 8 * sb);
              bulk_load(b_base + sb * p.stageB_bytes + half, wpack + (((size_t)(2 * c + 1) * 8 + t0) * p.Cout) * 8, half, full_b + 8 * sb);
            }
          }
          __syncwarp();
          ra.advance(p.SA);
          rb.advance(p.SB);
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Npiece >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t adesc0 = make_desc(0, (uint32_t)p.rows * 16u, 128u), bdesc0 = make_desc(0, (uint32_t)p.Npiece * 16u, 128u);
      const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
      if (elect_one()) {                  // one elected thread runs the whole role (see conv_tc_kernel)
        Ring ra, rb, rt;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, rt.advance(p.AS)) {
          const uint32_t as = rt.s, ap = rt.ph;
          mbar_wait(tmem_empty + 8 * as, ap ^ 1);
          tc_fence_after();
          const uint32_t d0 = tmem_base + as * (uint32_t)(p.MT * p.Npiece);
          for (int k = 0; k < kitems; ++k) {
            const uint32_t sa = ra.s, pa = ra.ph;
            const uint32_t sb = rb.s, pb = rb.ph;
            mbar_wait(full_a + 8 * sa, pa);
            mbar_wait(full_b + 8 * sb, pb);
            tc_fence_after();
            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (b_lo0 + ((b_base + sb * p.stageB_bytes) >> 4));
            uint32_t a_lo = a_lo0 + ((a_base + sa * p.slotA_bytes) >> 4), d = d0;
            for (int mt = 0; mt < p.MT; ++mt) {
              umma_bf16(d, ((uint64_t)a_hi << 32) | a_lo, bdesc, idesc, k ? 1u : 0u);
              d += (uint32_t)p.Npiece;
              a_lo += 128u;
            }
            umma_commit(empty_a + 8 * sa);
            umma_commit(empty_b + 8 * sb);
            ra.advance(p.SA); rb.advance(p.SB);
          }
          umma_commit(tmem_full + 8 * as);
        }
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    Ring rt;
    const int Xo = (p.mode == 1) ? p.Xh : 2 * p.Xh, Yo = (p.mode == 1) ? p.Yh : 2 * p.Yh, Zo = (p.mode == 1) ? p.Zh : 2 * p.Zh;
    const long long So = (long long)Xo * Yo * Zo;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, rt.advance(p.AS)) {
      const int brick = tile / p.npieces, piece = tile - brick * p.npieces;
      const int n = brick / bricks_per_n;
      int r = brick - n * bricks_per_n;
      const int bz = r % p.nbz; r /= p.nbz;
      const int by = r % p.nby;
      const int bx = r / p.nby;
      const uint32_t as = rt.s, ap = rt.ph;
      mbar_wait(tmem_full + 8 * as, ap);
      tc_fence_after();
      const uint32_t d0 = tmem_base + as * (uint32_t)(p.MT * p.Npiece) + ((uint32_t)(q * 32) << 16);
      for (int mt = 0; mt < p.MT; ++mt) {
        const int L = mt * 128 + q * 32 + lane;
        const int iz = L % p.BZ;
        const int ry = L / p.BZ;
        const int iy = ry % p.BY, ix = ry / p.BY;
        const int x = bx * p.BX + ix, y = by * p.BY + iy, z = bz * p.BZ + iz;
        const bool valid = (L < p.rows) && (x < p.Xh) && (y < p.Yh) && (z < p.Zh);
        if (p.mode == 2 && (p.TPc & 1) == 0) {
          // scatter, tap pairs: taps t (even) and t + 1 differ only in dz, so their outputs are the z-neighbours 2z and
          // 2z + 1 -- ONE 32-byte store per channel octet.  Consecutive lanes are consecutive half-resolution z, so a warp
          // writes 1 KB contiguous per instruction; per-tap 16-byte stores at a 32-byte stride half-filled every sector
          // twice (82 us for the 128 MB full-resolution output of the 32 -> 16 up-convolution).
          for (int tp = 0; tp < p.TPc; tp += 2) {
            const int t = piece * p.TPc + tp;
            const int ox = 2 * x + (t >> 2), oy = 2 * y + ((t >> 1) & 1);
            uint4* dst_t = out + (long long)n * Cob * So + ((long long)ox * Yo + oy) * Zo + 2 * z;
            for (int cg = 0; cg < p.Cout; cg += 16) {
              uint32_t v0[16], v1[16];
              tmem_ld16(d0 + (uint32_t)(mt * p.Npiece + tp * p.Cout + cg), v0);
              tmem_ld16(d0 + (uint32_t)(mt * p.Npiece + (tp + 1) * p.Cout + cg), v1);
              tmem_ld_wait();
              if (valid) {
                float f0[16], f1[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const float b = bias ? __ldg(bias + cg + k) : 0.f;
                  f0[k] = __uint_as_float(v0[k]) + b;
                  f1[k] = __uint_as_float(v1[k]) + b;
                }
                uint4* dst = dst_t + (long long)(cg >> 3) * So;
                stg256(dst, pack8(f0), pack8(f1));
                stg256(dst + So, pack8(f0 + 8), pack8(f1 + 8));
              }
            }
          }
          continue;
        }
        for (int c16 = 0; c16 < p.Npiece; c16 += 16) {
          uint32_t v[16];
          tmem_ld16(d0 + (uint32_t)(mt * p.Npiece + c16), v);
          tmem_ld_wait();
          if (valid) {
            int co = c16, ox = x, oy = y, oz = z;
            if (p.mode == 2) {
              const int tl = c16 / p.Cout;
              co = c16 - tl * p.Cout;
              const int t = piece * p.TPc + tl;
              ox = 2 * x + (t >> 2); oy = 2 * y + ((t >> 1) & 1); oz = 2 * z + (t & 1);
            }
            float f[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v[k]) + (bias ? __ldg(bias + co + k) : 0.f);
            uint4* dst = out + ((long long)n * Cob + (co >> 3)) * So + ((long long)ox * Yo + oy) * Zo + oz;
            dst[0] = pack8(f);
            dst[So] = pack8(f + 8);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + 8 * as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// =========================================================================================================
// weight gradient on tensor cores:  dW[t][co][ci] = sum_rows dy[row][co] * a[row + tap(t)][ci]
//   D[M = co (128 lanes)][N = ci] += A[M][K] * B[N][K],  K = 16 voxels along z per MMA, both operands "MN-major,
//   no swizzle" views of the same CB8 bricks the forward kernel uses (channels contiguous, voxels strided 16 B):
//   A = dy brick (compact [BX][BY][ZP] rows, ZP = Z rounded up to 16, tail rows zero-filled by TMA OOB),
//   B = halo'd a brick shifted by the tap offset -- again only a start-address change.
//   Each CTA owns a group of taps (TP*Cin <= 512 TMEM columns) and one 128-channel half of Cout, streams its
//   share of the bricks through a TMA ring and accumulates in TMEM across ALL of them; partial results go to a
//   workspace [split][T][Cout][Cin] and are summed in fixed order by a finalize kernel (deterministic).
// =========================================================================================================
struct WgParams {
  int N, X, Y, Z, Cin, Cout, kx;
  int BX, BY, HX, HY, HZ, ZP;
  int nbx, nby, nbricks;
  int rows_a, rows_dy;
  int T, TP, npass_t, MH, PL;     // taps, taps per pass, tap passes, M halves, dy planes loaded per brick
  int S, splits, tmem_cols;
  int s2;                         // 1: stride-2 family (B bricks are per-tap strided gathers of the full-res tensor)
  int zoff;                       // z-tiled launch (z-lines longer than a TMA box): first column of this launch's window
  int MM, m64map;                 // MMA M (128 or 64) and the TMEM row->lane map assumed for M = 64
  unsigned a_tx_bytes, dy_tx_bytes, a_alloc_bytes, dy_alloc_bytes, slot_bytes, offBar, tap_bytes;
  int mergedA, mergedD;
  // dz-folded mode (3*Cout <= 128): the three dz taps of a (dx,dy) pair ride in the M dimension.  The dy brick is loaded
  // three times, copy dz shifted by dz rows along z (a TMA coordinate), so ONE MMA per (dx,dy) produces
  // D[(dz,co)][ci]; a 16-channel layer then issues 9 MMAs per K-step with 48 useful rows instead of 27 with 16.
  int fz;                         // 1 or 3
  int KG, ngrp;                   // folded: (dx,dy) groups per pass / in total
  int KS, accumulate;             // in-kernel finalize: lanes sharing one output quad (power of two <= 32); dw += or =
  // flush mode (layers whose per-CTA accumulation chain is thousands of MMAs long): the tensor core adds into its fp32
  // accumulator with truncation, a bias that grows linearly with the chain (measured 3e-4 relative on the 112x112x80
  // 16-channel layer against 1.3e-5 for fp32 cuDNN).  Every `flush` bricks the accumulators of one of two TMEM sets are
  // drained into round-to-nearest fp32 REGISTER accumulators of the epilogue warps while the MMAs continue on the other set.
  int flush;                      // 0 = off, else bricks per flush (requires TP*Cin <= 144 columns, i.e. <= 9 accumulators x 16)
  // dx-folded mode (dz-folded layers with 3*Cin <= 256 and bricks one x-plane thick, BX = 1): the halo'd a brick holds its
  // three x-planes [plane][x = dx][y][z]; with HX = 3 the channel-plane stride is exactly 3 x-plane strides, so an MN-major
  // B descriptor whose N-group stride is ONE x-plane walks (plane, dx) uniformly: N = 3*Cin columns = the three dx taps of
  // every input channel in ONE MMA.  A K-step then issues 3 MMAs (dy = address offset) of N = 3*Cin instead of 9 of N = Cin:
  // 132 instead of 351 tensor-pipe cycles for 16 channels, 168 instead of 360 for 32 (cost max(N/2, 32 + N/4) per MMA).
  // Accumulator dy: column (plane*3 + dx)*8 + c8.
  int dxf;
  // per-tap compact mode (s2 == 2; small-volume layers, Cout > 32): instead of ONE halo'd a brick whose taps are address
  // offsets -- which forces z-lines padded to 16 rows (a 7x7x5 layer runs at 31 % K efficiency), tiny bricks with 6x halo
  // inflation and all Cin planes in every CTA -- every tap gets its OWN shifted copy of the brick, loaded by TMA at
  // coordinates (x+dx-1, y+dy-1, z+dz-1) with out-of-bounds zero fill, in the SAME compact [BX][BY][Z] row order as the dy
  // brick.  K then runs over the linear row index in steps of 16 (BX*BY*Z % 16 == 0, no padding inside the volume), and the
  // input channels are split NH ways across CTAs (pass = (tap group, M half, N slice)): more passes, fewer or no split-K
  // partials.  Deep layers are tiny, so the 27 shifted reads of a all hit L2.
  int NH;                         // N slices (input-channel groups) across passes; 1 outside the per-tap mode
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_dy,
                     float* __restrict__ partial, float* __restrict__ dw, int* __restrict__ counter, const WgParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + p.offBar;
  const uint32_t full = bar_base, empty = full + 8 * p.S, tmem_full = empty + 8 * p.S;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + p.offBar + 8 * (2 * p.S + 1));
  // flush mode: tmem_full / tmem_empty per accumulator set live behind the tap-offset table
  const uint32_t fl_full = bar_base + 8 * (2 * p.S + 1) + 128, fl_empty = fl_full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, pass = blockIdx.y;
  const int nh = pass % p.NH, mh = (pass / p.NH) % p.MH, tg = pass / (p.NH * p.MH);
  const int t0 = (p.fz == 3) ? tg * p.KG : tg * p.TP;                           // first tap (folded: first (dx,dy) group)
  const int ntap = (p.fz == 3) ? min(p.KG, p.ngrp - t0) : min(p.TP, p.T - t0);  // accumulators this CTA owns
  const int acc_cols = p.dxf ? 3 * p.Cin : p.Cin / p.NH;                         // columns per accumulator
  const int Cib = p.Cin >> 3, Cob = p.Cout >> 3;

  // zero the regions TMA never writes (unused dy planes, slack rows after the a planes): they feed MMAs
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < p.S; ++s) {
      uint4* slot = reinterpret_cast<uint4*>(smem + (size_t)s * p.slot_bytes);
      const unsigned a0 = p.a_tx_bytes / 16, a1 = p.a_alloc_bytes / 16;
      for (unsigned i = a0 + threadIdx.x; i < a1; i += TC_THREADS) slot[i] = z;
      const unsigned d0 = (p.a_alloc_bytes + p.dy_tx_bytes) / 16, d1 = (p.a_alloc_bytes + p.dy_alloc_bytes) / 16;
      for (unsigned i = d0 + threadIdx.x; i < d1; i += TC_THREADS) slot[i] = z;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.S; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
    mbar_init(tmem_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(fl_full + 8 * i, 1); mbar_init(fl_empty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int bricks_per_n = p.nbx * p.nby;
  const bool has_work = split < p.nbricks;

  if (warp == 0) {
    {
      Ring rs;
      for (int brick = split; brick < p.nbricks; brick += p.splits, rs.advance(p.S)) {
        const int n = brick / bricks_per_n;
        const int r = brick - n * bricks_per_n;
        const int by = r % p.nby, bx = r / p.nby;
        const uint32_t s = rs.s, ph = rs.ph;
        mbar_wait(empty + 8 * s, ph ^ 1);
        const uint32_t slot = sbase + s * p.slot_bytes;
        if (elect_one()) {
        if (p.s2 == 2) {               // per-tap compact mode: shifted copies of the (Cin/NH)-channel brick
          mbar_expect_tx(full + 8 * s, (uint32_t)ntap * p.tap_bytes + p.dy_tx_bytes);
          const int pl0 = n * Cib + nh * (Cib / p.NH);
          for (int tl = 0; tl < ntap; ++tl) {
            const int t = t0 + tl;
            const int tz = t % 3, ty = (t / 3) % 3, tx = (p.kx == 3) ? t / 9 : 1;
            tma_load_cb8(slot + (uint32_t)tl * p.tap_bytes, &map_a, full + 8 * s, p.mergedA, tz - 1, by * p.BY + ty - 1, bx * p.BX + tx - 1, pl0);
          }
        } else if (p.s2) {
          mbar_expect_tx(full + 8 * s, (uint32_t)ntap * p.tap_bytes + p.dy_tx_bytes);
          for (int tl = 0; tl < ntap; ++tl) {
            const int t = t0 + tl;
            tma_load_5d(slot + (uint32_t)tl * p.tap_bytes, &map_a, full + 8 * s, 0, (t & 1), 2 * by * p.BY + ((t >> 1) & 1),
                        2 * bx * p.BX + (t >> 2), n * Cib);
          }
        } else {
          mbar_expect_tx(full + 8 * s, p.a_tx_bytes + p.dy_tx_bytes);
          tma_load_cb8(slot, &map_a, full + 8 * s, p.mergedA, p.zoff - 1, by * p.BY - 1, bx * p.BX - (p.kx >> 1), n * Cib);
        }
        if (p.fz == 3) {
          const uint32_t copy_bytes = (uint32_t)Cob * (uint32_t)p.rows_dy * 16u;
          for (int dz = 0; dz < 3; ++dz)
            tma_load_cb8(slot + p.a_alloc_bytes + (uint32_t)dz * copy_bytes, &map_dy, full + 8 * s, p.mergedD, -dz, by * p.BY, bx * p.BX, n * Cob);
        } else {
          tma_load_cb8(slot + p.a_alloc_bytes, &map_dy, full + 8 * s, p.mergedD, 0, by * p.BY, bx * p.BX, n * Cob + mh * 16);
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (has_work) {
      // M = 128, N = Cin, both operands MN-major (bits 15, 16)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(acc_cols >> 3) << 17) | ((uint32_t)(p.MM >> 4) << 24);
      // constant descriptor parts; per MMA only the 16-byte-unit start address changes
      const uint64_t adesc0 = make_desc(0, 128u, (uint32_t)p.rows_dy * 16u);
      // B N-group stride: the next 8-channel plane -- or, dx-folded, the next x-plane of the halo'd brick
      const uint64_t bdesc0 = make_desc(0, 128u, p.dxf ? (uint32_t)(p.HY * p.HZ) * 16u : (uint32_t)p.rows_a * 16u);
      // per-accumulator B row offset (same-conv: halo shift of the tap; folded: (dx,dy) shift only; stride-2: separate
      // gathered brick per tap)
      uint32_t* tapoff = reinterpret_cast<uint32_t*>(smem + p.offBar + 8 * (2 * p.S + 1) + 16);
      if (lane < 27) {
        const int t = t0 + lane;
        uint32_t off;
        if (p.s2) off = (uint32_t)lane * (p.tap_bytes >> 4);
        else if (p.dxf) off = (uint32_t)(t * p.HZ);                                 // accumulator = dy
        else if (p.fz == 3) off = (p.kx == 3) ? (uint32_t)(((t / 3) * p.HY + (t % 3)) * p.HZ) : (uint32_t)(t * p.HZ);
        else { const int tz = t % 3, ty = (t / 3) % 3, tx = t / 9; off = (uint32_t)((tx * p.HY + ty) * p.HZ + tz); }
        tapoff[lane] = off;
      }
      __syncwarp();
      // one elected thread runs the role; descriptor high words are constants, only 32-bit start fields move
      if (elect_one()) {
        const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
        const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
        const uint32_t cin = (uint32_t)acc_cols;
        Ring rs;
        uint32_t acc = 0;
        const uint32_t row_step_y = (uint32_t)(p.s2 ? p.ZP : p.HZ);                 // B-row distance between consecutive lines
        const uint32_t row_step_x = (uint32_t)(p.s2 ? p.BY * p.ZP : p.HY * p.HZ);
        const uint32_t set_cols = (uint32_t)(p.TP * acc_cols);                       // flush mode: columns per accumulator set
        uint32_t fcount = 0, fidx = 0, tm = tmem_base;                               // bricks in the current flush group, group index
        for (int brick = split; brick < p.nbricks; brick += p.splits, rs.advance(p.S)) {
          const uint32_t s = rs.s, ph = rs.ph;
          if (p.flush && fcount == 0) {                                               // first brick of a flush group: claim a set
            const uint32_t set = fidx & 1u;
            mbar_wait(fl_empty + 8 * set, ((fidx >> 1) & 1u) ^ 1u);
            tc_fence_after();
            tm = tmem_base + set * set_cols;
            acc = 0;
          }
          mbar_wait(full + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_slot = sbase + s * p.slot_bytes, dy_slot = a_slot + p.a_alloc_bytes;
          uint32_t arow_x = a_lo0 + (dy_slot >> 4), brow_x = b_lo0 + (a_slot >> 4);
          if (p.s2 == 2) {                                                            // compact bricks: K over the linear row index
            for (uint32_t zc = 0; zc < (uint32_t)p.rows_dy; zc += 16) {
              const uint64_t adesc = ((uint64_t)a_hi << 32) | (arow_x + zc);
              const uint32_t bz = brow_x + zc;
#pragma unroll 4
              for (int tl = 0; tl < ntap; ++tl)
                umma_bf16(tm + (uint32_t)tl * cin, adesc, ((uint64_t)b_hi << 32) | (bz + tapoff[tl]), idesc, acc);
              acc = 1;
            }
          } else
          for (int ix = 0; ix < p.BX; ++ix) {
            uint32_t arow = arow_x, brow = brow_x;
            for (int iy = 0; iy < p.BY; ++iy) {
              for (uint32_t zc = 0; zc < (uint32_t)p.ZP; zc += 16) {
                const uint64_t adesc = ((uint64_t)a_hi << 32) | (arow + zc);
                const uint32_t bz = brow + zc;
                if (ntap == 9) {
#pragma unroll
                  for (int tl = 0; tl < 9; ++tl)
                    umma_bf16(tm + (uint32_t)tl * cin, adesc, ((uint64_t)b_hi << 32) | (bz + tapoff[tl]), idesc, acc);
                } else if (ntap == 3) {
#pragma unroll
                  for (int tl = 0; tl < 3; ++tl)
                    umma_bf16(tm + (uint32_t)tl * cin, adesc, ((uint64_t)b_hi << 32) | (bz + tapoff[tl]), idesc, acc);
                } else {
#pragma unroll 4
                  for (int tl = 0; tl < ntap; ++tl)
                    umma_bf16(tm + (uint32_t)tl * cin, adesc, ((uint64_t)b_hi << 32) | (bz + tapoff[tl]), idesc, acc);
                }
                acc = 1;
              }
              arow += (uint32_t)p.ZP;
              brow += row_step_y;
            }
            arow_x += (uint32_t)(p.BY * p.ZP);
            brow_x += row_step_x;
          }
          umma_commit(empty + 8 * s);
          if (p.flush && (++fcount == (uint32_t)p.flush || brick + p.splits >= p.nbricks)) {   // group complete: hand the set over
            umma_commit(fl_full + 8 * (fidx & 1u));
            ++fidx;
            fcount = 0;
          }
        }
        if (!p.flush) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    int co = mh * 128 + q * 32 + lane;
    if (p.MM == 64) {
      // M = 64 accumulators occupy 64 of the 128 TMEM lanes; m64map selects the row -> lane convention
      //   0: row r at lane (r & 15) + 32 * (r >> 4)      1: row r at lane r      2: row r at lane (r & 31) + 64 * (r >> 5)
      const int l = q * 32 + lane;
      if (p.m64map == 0) co = ((l & 31) < 16) ? ((l >> 5) * 16 + (l & 15)) : (1 << 30);
      else if (p.m64map == 1) co = (l < 64) ? l : (1 << 30);
      else co = ((l & 63) < 32) ? ((l >> 6) * 32 + (l & 31)) : (1 << 30);
    }
    int dzl = 0;
    if (p.fz == 3) {                      // accumulator row r = dz * Cout + co
      const int r = co;
      if (r < 3 * p.Cout) { dzl = r / p.Cout; co = r - dzl * p.Cout; } else co = 1 << 30;
    }
    if (p.flush) {
      // register accumulators: <= 9 accumulators x 16 input channels per TMEM lane (row)
      float accr[9][16];
#pragma unroll
      for (int tl = 0; tl < 9; ++tl)
#pragma unroll
        for (int k = 0; k < 16; ++k) accr[tl][k] = 0.f;
      const int nb_cta = has_work ? (p.nbricks - split + p.splits - 1) / p.splits : 0;
      const int ngroups = (nb_cta + p.flush - 1) / p.flush;
      const uint32_t set_cols = (uint32_t)(p.TP * acc_cols);
      const int nblk = p.dxf ? 9 : ntap;                 // 16-column blocks per accumulator set (dx-folded: 3 accumulators x 48)
      for (int f = 0; f < ngroups; ++f) {
        const uint32_t set = (uint32_t)f & 1u;
        mbar_wait(fl_full + 8 * set, ((uint32_t)f >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int tl = 0; tl < 9; ++tl) {
          if (tl < nblk) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + set * set_cols + (uint32_t)(tl * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) accr[tl][k] += __uint_as_float(v[k]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(fl_empty + 8 * set);
      }
#pragma unroll
      for (int tl = 0; tl < 9; ++tl) {
        if (tl < nblk && co < p.Cout) {
          if (p.dxf) {                     // block tl = (dy, 16-column chunk): two (plane, dx) octets (Cin = 16)
            const int dy = tl / 3, ch = tl - dy * 3;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int gidx = ch * 2 + h, plane = gidx / 3, dx = gidx - plane * 3;
              const int t = (dx * 3 + dy) * 3 + dzl;
              float4* d4 = reinterpret_cast<float4*>(partial + (((size_t)split * p.T + t) * p.Cout + co) * p.Cin + plane * 8);
              d4[0] = make_float4(accr[tl][8 * h + 0], accr[tl][8 * h + 1], accr[tl][8 * h + 2], accr[tl][8 * h + 3]);
              d4[1] = make_float4(accr[tl][8 * h + 4], accr[tl][8 * h + 5], accr[tl][8 * h + 6], accr[tl][8 * h + 7]);
            }
          } else {
          const int t = (p.fz == 3) ? (t0 + tl) * 3 + dzl : t0 + tl;
          float4* d4 = reinterpret_cast<float4*>(partial + (((size_t)split * p.T + t) * p.Cout + co) * p.Cin);
          d4[0] = make_float4(accr[tl][0], accr[tl][1], accr[tl][2], accr[tl][3]);
          d4[1] = make_float4(accr[tl][4], accr[tl][5], accr[tl][6], accr[tl][7]);
          d4[2] = make_float4(accr[tl][8], accr[tl][9], accr[tl][10], accr[tl][11]);
          d4[3] = make_float4(accr[tl][12], accr[tl][13], accr[tl][14], accr[tl][15]);
          }
        }
      }
    } else {
    if (has_work) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    if (p.dxf) {
      // accumulator tl = dy; its 3*Cin columns are (plane, dx) octets
      for (int tl = 0; tl < ntap; ++tl) {
        for (int c16 = 0; c16 < acc_cols; c16 += 16) {
          uint32_t v[16];
          if (has_work) {
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * acc_cols + c16), v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = 0u;
          }
          if (co < p.Cout) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int gidx = (c16 >> 3) + h, plane = gidx / 3, dx = gidx - plane * 3;
              const int t = (dx * 3 + tl) * 3 + dzl;
              float4* d4 = reinterpret_cast<float4*>(partial + (((size_t)split * p.T + t) * p.Cout + co) * p.Cin + plane * 8);
              d4[0] = make_float4(__uint_as_float(v[8 * h + 0]), __uint_as_float(v[8 * h + 1]), __uint_as_float(v[8 * h + 2]), __uint_as_float(v[8 * h + 3]));
              d4[1] = make_float4(__uint_as_float(v[8 * h + 4]), __uint_as_float(v[8 * h + 5]), __uint_as_float(v[8 * h + 6]), __uint_as_float(v[8 * h + 7]));
            }
          }
        }
      }
    } else
    for (int tl = 0; tl < ntap; ++tl) {
      const int t = (p.fz == 3) ? (t0 + tl) * 3 + dzl : t0 + tl;
      float* dst = partial + (((size_t)split * p.T + t) * p.Cout + co) * p.Cin + nh * acc_cols;
      for (int c16 = 0; c16 < acc_cols; c16 += 16) {
        uint32_t v[16];
        if (has_work) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * acc_cols + c16), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = 0u;
        }
        if (co < p.Cout) {
          float4* d4 = reinterpret_cast<float4*>(dst + c16);
          d4[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          d4[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          d4[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
          d4[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
        }
      }
    }
    }
  }
  tc_fence_before();
  __threadfence();                       // this CTA's partial is visible device-wide before it arrives at the barrier
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
  // ---- in-kernel finalize (replaces a separate finalize launch and its ~25 us of launch gap + cold reads): the grid is
  // at most one CTA per SM, all co-resident, every CTA arrives at a device-wide barrier, then ALL
  // CTAs sum the partials -- dw[co][ci][t] = sum_split partial[split][t][co][ci] -- each output quad by a fixed set of KS
  // lanes in a fixed order (deterministic, no float atomics).  counter[0] = arrivals, counter[1] = departures; the last
  // CTA to leave resets both, so the pair is reusable by the next launch on the stream.
  const unsigned nct = gridDim.x * gridDim.y;
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen < nct) __nanosleep(64);
    } while (seen < nct);
  }
  __syncthreads();
  {
    const long long per = (long long)p.T * p.Cout * p.Cin;
    const long long items = per >> 2;                                 // float4 outputs (Cin % 16 == 0)
    const float4* part4 = reinterpret_cast<const float4*>(partial);
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    const unsigned KS = (unsigned)p.KS, opw = 32u / KS;               // output quads per warp per sub-pass
    const unsigned ks = lane & (KS - 1), ol = lane / KS;
    constexpr int NO = 4;                                             // independent output quads per thread per pass (ILP:
                                                                      // the grid has only ~28k threads for up to 9M loads)
    const long long wstride = (long long)nct * (TC_THREADS / 32) * opw * NO;
    for (long long base = ((long long)cta * (TC_THREADS / 32) + warp) * opw * NO; base < items; base += wstride) {
      float4 acc[NO];
      long long o[NO];
#pragma unroll
      for (int j = 0; j < NO; ++j) { acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); o[j] = base + (long long)j * opw + ol; }
#pragma unroll 2
      for (int k = (int)ks; k < p.splits; k += (int)KS) {
        const float4* src = part4 + (long long)k * items;
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          if (o[j] < items) {
            const float4 v = __ldcg(src + o[j]);
            acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
          }
        }
      }
      for (unsigned m = 1; m < KS; m <<= 1) {
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, m);
          acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, m);
          acc[j].z += __shfl_xor_sync(0xffffffffu, acc[j].z, m);
          acc[j].w += __shfl_xor_sync(0xffffffffu, acc[j].w, m);
        }
      }
      if (ks == 0) {
        float* dst[NO];
        float old[NO][4];
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          const long long i = o[j] << 2;
          const int ci = (int)(i % p.Cin);
          const int co = (int)((i / p.Cin) % p.Cout);
          const int t = (int)(i / ((long long)p.Cin * p.Cout));
          dst[j] = dw + ((long long)co * p.Cin + ci) * p.T + t;
#pragma unroll
          for (int q = 0; q < 4; ++q) old[j][q] = (p.accumulate && o[j] < items) ? dst[j][q * p.T] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          if (o[j] < items) {
            dst[j][0] = old[j][0] + acc[j].x; dst[j][p.T] = old[j][1] + acc[j].y;
            dst[j][2 * p.T] = old[j][2] + acc[j].z; dst[j][3 * p.T] = old[j][3] + acc[j].w;
          }
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = (unsigned)atomicAdd(counter + 1, 1);
    if (prev == nct - 1) { counter[0] = 0; counter[1] = 0; __threadfence(); }
  }
}

// stand-alone finalize kept for reference / debugging: dw[co][ci][t] = sum_split partial[split][t][co][ci]   (fixed order)
__global__ void conv_tc_wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int T,
                                              int Cout, int Cin, int accumulate) {
  const long long per = (long long)T * Cout * Cin;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per) return;
  const int ci = (int)(i % Cin);
  const int co = (int)((i / Cin) % Cout);
  const int t = (int)(i / ((long long)Cin * Cout));
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partial[(long long)k * per + i];
  float* dst = dw + ((long long)co * Cin + ci) * T + t;
  *dst = accumulate ? *dst + s : s;
}

// ---------------------------------------------------------------------------------------------------------
// host side: brick-shape selection, tensor map, launch
// ---------------------------------------------------------------------------------------------------------
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

EncodeTiledFn tma_encoder() { return get_encode(); }

// Tensor map over a CB8 tensor [planes][X][Y][Z][8] bf16 for a box of (bp planes, bx, by, bz voxels).  When the z-run
// fits a 256-element box the map is built over 8-byte elements with (z, channel-octet) folded into the innermost
// dimension: the box's innermost extent becomes bz*16 bytes instead of 16, which is what the TMA unit's throughput
// depends on (a 16-byte inner box caps a SM's TMA at roughly 8-10 B/clk -- measured on the c64..c256 layers).
// Out-of-bounds z (the conv padding) still zero-fills because z*2 stays the coordinate of its own dimension.
CUresult encode_cb8(EncodeTiledFn enc, CUtensorMap* map, const void* base, long long Z, long long Y, long long X,
                    long long planes, int bz, int by, int bx, int bp, int* merged, long long Zpitch) {
  // Zpitch != 0: `base` points at a z-window of a wider tensor -- Z is the window's extent (out-of-bounds beyond it),
  // Zpitch the real row length the strides follow (z-tiled weight gradient)
  const int want = (bz <= 128) ? 1 : 0;
  *merged = want;
  const long long ZS = Zpitch ? Zpitch : Z;
  if (want) {
    const cuuint64_t gdim[4] = {(cuuint64_t)(2 * Z), (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)planes};
    const cuuint64_t gstr[3] = {(cuuint64_t)ZS * 16, (cuuint64_t)ZS * Y * 16, (cuuint64_t)ZS * Y * X * 16};
    const cuuint32_t box[4] = {(cuuint32_t)(2 * bz), (cuuint32_t)by, (cuuint32_t)bx, (cuuint32_t)bp};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  const cuuint64_t gdim[5] = {8, (cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)planes};
  const cuuint64_t gstr[4] = {16, (cuuint64_t)ZS * 16, (cuuint64_t)ZS * Y * 16, (cuuint64_t)ZS * Y * X * 16};
  const cuuint32_t box[5] = {8, (cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, (cuuint32_t)bp};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

constexpr unsigned SMEM_BUDGET = 220 * 1024;   // of the 227 KB a CTA may opt into

static bool plan(TcParams& p, int nsm) {
  const int T = p.kx * 9;
  p.T = T;
  p.nchunks = p.Cin / 16;
  double best = 1e300;
  TcParams bestp = p;
  bool found = false;
  int bz_opts[4];
  int nz = 0;
  for (int d = 1; d <= 8 && nz < 4; d *= 2) {
    const int bz = (p.Z + d - 1) / d;
    if (bz + 2 <= 256 && (nz == 0 || bz_opts[nz - 1] != bz)) bz_opts[nz++] = bz;
  }
  for (int NS = 1; NS <= 8; NS *= 2) {
    if (p.Cout % (16 * NS)) break;
    const int Ns = p.Cout / NS;
    if (Ns > 128) continue;            // an N = 128 MMA already runs at the tensor pipe's rate (N/2 cycles); wider buys nothing
    // a weight stage holds whole dx-groups of nine (dy,dz) taps (the issuer unrolls them): all T taps of a chunk when two
    // such stages fit next to the A slots (one barrier round trip per chunk), else one dx-group
    const double per_mma = (Ns / 2.0 > 32.0 + Ns / 4.0) ? Ns / 2.0 : 32.0 + Ns / 4.0;
    for (int zi = 0; zi < nz; ++zi) {
      const int BZ = bz_opts[zi], HZ = BZ + 2;
      for (int BY = 1; BY <= p.Y && BY + 2 <= 256; ++BY) {
        const int HY = BY + 2;
        const int bxmax = (p.kx == 1) ? 1 : p.X;
        for (int BX = 1; BX <= bxmax; ++BX) {
          const int HX = BX + p.kx - 1;
          if (HX > 256) break;
          const long long rows_h = (long long)HX * HY * HZ;
          if (rows_h >= 16384) break;
          const long long lmax = ((long long)(BX - 1) * HY + (BY - 1)) * HZ + BZ;
          const int MT = (int)((lmax + 127) / 128);
          if (MT * Ns > 512) break;
          const long long rows_alloc = (long long)MT * 128 + ((long long)(p.kx - 1) * HY + 2) * HZ + 2;
          const long long slotA = ((rows_h + (rows_alloc > rows_h ? rows_alloc : rows_h)) * 16 + 127) / 128 * 128;
          // A slots: 2..3 (one (brick, 16-channel) chunk each); the rest of shared memory goes to weight stages, whose
          // depth hides the L2 latency of the weight stream
          int TG = T;
          unsigned stageB = (unsigned)TG * (unsigned)Ns * 32u;
          if ((long long)SMEM_BUDGET - (long long)p.reserve - 1024 - 2 * slotA < 2ll * stageB) { TG = 9; stageB = 9u * (unsigned)Ns * 32u; }
          const int sb_min = (TG == T) ? 2 : 3;
          long long avail = (long long)SMEM_BUDGET - (long long)p.reserve - 1024 - (long long)sb_min * stageB;
          int SA = (int)(avail / slotA);
          if (SA < 2) break;
          if (SA > 3) SA = 3;
          int SB = (int)(((long long)SMEM_BUDGET - (long long)p.reserve - 1024 - (long long)SA * slotA) / stageB);
          if (SB > 8) SB = 8;
          if (SB < sb_min) SB = sb_min;
          const int nbx = (p.X + BX - 1) / BX, nby = (p.Y + BY - 1) / BY, nbz = (p.Z + BZ - 1) / BZ;
          const long long nb = (long long)p.N * nbx * nby * nbz;
          const long long items = nb * NS;
          const long long waves = (items + nsm - 1) / nsm;
          const double mma_cyc = (double)MT * T * p.nchunks * per_mma;
          const double item_bytes = (double)rows_h * p.Cin * 2.0 + (double)T * p.nchunks * Ns * 32.0;
          const double load_cyc = item_bytes / 40.0;                      // one SM's TMA intake
          const double epi_cyc = (double)MT * Ns * 12.0;
          const int AS = (2 * MT * Ns <= 512) ? 2 : 1;
          const double per_item = (mma_cyc > load_cyc ? mma_cyc : load_cyc) + (AS == 2 ? 0.0 : epi_cyc) + 1500.0;
          const double l2_bound = (double)items * item_bytes / 2500.0;   // chip-wide L2 -> SM bandwidth (bytes/clk)
          const double t_sm = (double)waves * per_item;
          const double cost = t_sm > l2_bound ? t_sm : l2_bound;
          if (cost < best) {
            best = cost;
            found = true;
            bestp = p;
            bestp.BX = BX; bestp.BY = BY; bestp.BZ = BZ; bestp.HX = HX; bestp.HY = HY; bestp.HZ = HZ;
            bestp.nbx = nbx; bestp.nby = nby; bestp.nbz = nbz; bestp.nbricks = (int)nb;
            bestp.rows_h = (int)rows_h; bestp.MT = MT; bestp.SA = SA; bestp.SB = SB; bestp.AS = AS;
            bestp.slotA_bytes = (unsigned)slotA; bestp.stageB_bytes = stageB; bestp.TG = TG; bestp.NS = NS; bestp.Ns = Ns;
          }
        }
      }
    }
  }
  if (!found) return false;
  p = bestp;
  int cols = 32;
  while (cols < p.AS * p.MT * p.Ns) cols *= 2;
  p.tmem_cols = cols;
  p.offA = 0;
  p.offB = p.SA * p.slotA_bytes;
  p.offBar = p.offB + p.SB * p.stageB_bytes;
  return true;
}

struct PlanKey {
  int v[8];
  bool operator==(const PlanKey& o) const { for (int i = 0; i < 8; ++i) if (v[i] != o.v[i]) return false; return true; }
};
// immutable plans memoised per shape (pure function of the key; guarded by a mutex)
static bool plan_cached(TcParams& p, int nsm) {
  static std::mutex mu;
  static PlanKey keys[256];
  static TcParams vals[256];
  static int count = 0;
  const PlanKey k = {{p.N, p.X, p.Y, p.Z, p.Cin, p.Cout, p.kx, (int)p.reserve}};
  {
    std::lock_guard<std::mutex> g(mu);
    for (int i = 0; i < count; ++i) if (keys[i] == k) { p = vals[i]; return true; }
  }
  if (!plan(p, nsm)) return false;
  std::lock_guard<std::mutex> g(mu);
  if (count < 256) { keys[count] = k; vals[count] = p; ++count; }
  return true;
}

static bool shape_ok(int cin, int cout, const int* dims, const int* kernel) {
  if (cin % 16 || cout % 16 || cin < 16 || cout < 16 || cout > 256) return false;
  if (!((kernel[0] == 3 || kernel[0] == 1) && kernel[1] == 3 && kernel[2] == 3)) return false;
  if (kernel[0] == 1 && dims[0] != 1) return false;
  if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1) return false;
  return true;
}



static bool s2_shape_ok(int cin, int cout, const int* half_dims) {
  if (cin % 16 || cout % 16 || cin < 16 || cout < 16 || cin > 256 || cout > 256) return false;
  if (half_dims[0] < 1 || half_dims[1] < 1 || half_dims[2] < 1) return false;
  return true;
}

static bool s2_plan(S2Params& p) {
  p.nchunks = p.Cin / 16;
  if (p.mode == 1) { p.TPc = 8; p.npieces = 1; p.Npiece = p.Cout; }
  else {
    int tp = 8;
    while (tp > 1 && tp * p.Cout > 256) tp >>= 1;
    if (tp * p.Cout > 256) return false;
    p.TPc = tp; p.npieces = 8 / tp; p.Npiece = tp * p.Cout;
  }
  int mtmax = 512 / (2 * p.Npiece);
  if (mtmax < 1) mtmax = 1;
  if (mtmax > 4) mtmax = 4;
  const int target = mtmax * 128;
  p.BZ = p.Zh < 128 ? p.Zh : 128;
  int by = target / p.BZ; if (by < 1) by = 1; if (by > p.Yh) by = p.Yh; if (by > 128) by = 128;
  p.BY = by;
  int bx = target / (p.BZ * p.BY); if (bx < 1) bx = 1; if (bx > p.Xh) bx = p.Xh; if (bx > 128) bx = 128;
  p.BX = bx;
  p.rows = p.BX * p.BY * p.BZ;
  p.MT = (p.rows + 127) / 128;
  if (p.MT * p.Npiece > 512) return false;
  p.AS = (2 * p.MT * p.Npiece <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p.AS * p.MT * p.Npiece) cols *= 2;
  p.tmem_cols = cols;
  p.nbx = (p.Xh + p.BX - 1) / p.BX; p.nby = (p.Yh + p.BY - 1) / p.BY; p.nbz = (p.Zh + p.BZ - 1) / p.BZ;
  p.nbricks = p.N * p.nbx * p.nby * p.nbz;
  p.slotA_bytes = (unsigned)(((long long)(p.rows + p.MT * 128) * 16 + 127) / 128 * 128);
  p.stageB_bytes = (unsigned)p.Npiece * 32u;
  p.SB = 4;
  long long avail = (long long)SMEM_BUDGET - 1024 - (long long)p.SB * p.stageB_bytes;
  int sa = (int)(avail / p.slotA_bytes);
  if (sa < 2) return false;
  p.SA = sa > 8 ? 8 : sa;
  p.offA = 0; p.offB = p.SA * p.slotA_bytes; p.offBar = p.offB + p.SB * p.stageB_bytes;
  return true;
}

// M = 64 accumulate mode for the weight-gradient kernels (output channels <= 64): halves the A-operand shared-memory
// traffic.  TMEM row->lane map 0 (row r at lane (r & 15) + 32 * (r >> 4)) was validated against torch on the B200
// (tests/test_gpu_primitives.py::test_conv_production_shapes); the kernel keeps the other two conventions for reference.
static constexpr int wg_m64_mode() { return 0; }

// A z-line (+halo) must fit a TMA box of 256 rows.  Longer lines (the 256-wide ACDC slices) are cut into windows of at most
// 128 columns, one launch per window: the dy map is based at the window (nothing outside it contributes), the `a` map keeps
// the whole line and the kernel shifts its z coordinate by WgParams::zoff, so the halo columns are the real neighbours;
// launches after the first accumulate into dw.
static int wg_ztiles(int Z) { return (Z + 2 <= 256) ? 1 : (Z + 127) / 128; }
static int wg_tile_width(int Z) {
  const int nt = wg_ztiles(Z);
  return nt == 1 ? Z : ((Z + nt - 1) / nt + 15) / 16 * 16;
}
static bool wg_shape_ok(int cin, int cout, const int* dims, const int* kernel) {
  if (!shape_ok(cin, cout, dims, kernel)) return false;
  if (wg_tile_width(dims[2]) + 2 > 256) return false;
  return true;
}

// Plan of the per-tap compact mode (WgParams::NH).  Returns the modelled cost in cycles, or a negative value.
static double wg_plan_pertap(WgParams& p, int nsm) {
  p.T = p.kx * 9;
  p.MH = (p.Cout > 128) ? 2 : 1;
  p.PL = (p.Cout / 8 < 16) ? p.Cout / 8 : 16;
  p.MM = (p.Cout <= 64 && wg_m64_mode() >= 0) ? 64 : 128;
  p.m64map = wg_m64_mode() < 0 ? 0 : wg_m64_mode();
  p.ZP = p.Z; p.HX = p.HY = p.HZ = 0; p.s2 = 2;
  p.fz = 1; p.KG = 0; p.ngrp = 0; p.dxf = 0; p.flush = 0;
  if (p.Z > 128) return -1.0;
  const double out_bytes = (double)p.T * p.Cout * p.Cin * 4.0;
  double best = 1e300; bool found = false; WgParams bp = p;
  for (int NH = 1; NH <= 4; NH *= 2) {
    const int accc = p.Cin / NH;
    if (accc < 32 || accc % 16) break;
    const int tpmax = (512 / accc < p.T) ? 512 / accc : p.T;
    const double per_mma = (accc / 2.0 > 32.0 + accc / 4.0) ? accc / 2.0 : 32.0 + accc / 4.0;
    for (int TP = 1; TP <= tpmax; ++TP) {
      const int npass_t = (p.T + TP - 1) / TP, npass = npass_t * p.MH * NH;
      if (npass > nsm) continue;
      const int bxmax = (p.kx == 1) ? 1 : p.X;
      for (int BY = 1; BY <= p.Y && BY <= 128; ++BY) {
        for (int BX = 1; BX <= bxmax; ++BX) {
          const long long rows = (long long)BX * BY * p.Z;
          if (rows >= 16384) break;
          if (rows % 16) continue;
          const long long tap_bytes = rows * 16 * (accc / 8), a_alloc = ((long long)TP * tap_bytes + 127) / 128 * 128;
          const long long dy_tx = rows * 16 * p.PL, dy_alloc = (rows * 16 * (p.MM / 8) + 127) / 128 * 128;
          const long long slot = a_alloc + dy_alloc;
          int S = (int)((SMEM_BUDGET - 1024) / slot);
          if (S < 2) break;
          if (S > 4) S = 4;
          const int nbx = (p.X + BX - 1) / BX, nby = (p.Y + BY - 1) / BY;
          const long long nb = (long long)p.N * nbx * nby;
          long long splits = nsm / npass; if (splits < 1) splits = 1; if (splits > nb) splits = nb;
          const long long per_cta = (nb + splits - 1) / splits;
          const double mma_cyc = (double)(rows / 16) * TP * per_mma;
          const double load_cyc = (double)(TP * tap_bytes + dy_tx) / 40.0;
          // split-K partials: written by every CTA, read back by the in-kernel finalize (through L2, ~32 B/clk per SM)
          const double fin_cyc = 2.0 * (double)splits * out_bytes / ((double)nsm * 32.0);
          const double cost = (double)per_cta * ((mma_cyc > load_cyc ? mma_cyc : load_cyc) + 1200.0) + fin_cyc;
          if (cost < best) {
            best = cost; found = true; bp = p;
            bp.NH = NH; bp.TP = TP; bp.npass_t = npass_t;
            bp.BX = BX; bp.BY = BY; bp.nbx = nbx; bp.nby = nby; bp.nbricks = (int)nb;
            bp.rows_a = (int)rows; bp.rows_dy = (int)rows; bp.S = S; bp.splits = (int)splits;
            bp.tap_bytes = (unsigned)tap_bytes; bp.a_tx_bytes = (unsigned)a_alloc; bp.a_alloc_bytes = (unsigned)a_alloc;
            bp.dy_tx_bytes = (unsigned)dy_tx; bp.dy_alloc_bytes = (unsigned)dy_alloc; bp.slot_bytes = (unsigned)slot;
          }
        }
      }
    }
  }
  if (!found) return -1.0;
  p = bp;
  int cols = 32;
  while (cols < p.TP * (p.Cin / p.NH)) cols *= 2;
  p.tmem_cols = cols;
  p.offBar = p.S * p.slot_bytes;
  return best;
}

static bool wg_plan(WgParams& p, int nsm, int max_splits = 0) {
  // small volumes with wide channels: the per-tap compact mode (no z padding, channel-sliced passes)
  if (max_splits >= 0 && 3 * p.Cout > 128 && (long long)p.X * p.Y * p.Z <= 16384 && p.Cin >= 32) {
    WgParams q = p;
    if (wg_plan_pertap(q, nsm) > 0.0) { p = q; return true; }
  }
  if (max_splits < 0) max_splits = 0;
  p.T = p.kx * 9;
  p.TP = 512 / p.Cin;
  if (p.TP > p.T) p.TP = p.T;
  if (p.TP < 1) return false;
  p.npass_t = (p.T + p.TP - 1) / p.TP;
  p.MH = (p.Cout > 128) ? 2 : 1;
  p.PL = (p.Cout / 8 < 16) ? p.Cout / 8 : 16;
  p.MM = (p.Cout <= 64 && wg_m64_mode() >= 0) ? 64 : 128;
  p.m64map = wg_m64_mode() < 0 ? 0 : wg_m64_mode();
  p.HZ = p.Z + 2;
  p.ZP = (p.Z + 15) / 16 * 16;
  p.fz = 1; p.KG = 0; p.ngrp = 0; p.dxf = 0; p.NH = 1;
  if (3 * p.Cout <= 128 && (p.Z + 2 + 15) / 16 * 16 <= 256 && p.Cin <= 512 / 3 && (wg_m64_mode() >= 0 || 3 * p.Cout > 64)) {
    // dz-folded: K runs over the halo'd z-line (Z+2 rows, rounded up to 16); both bricks use that z extent
    p.fz = 3;
    p.ZP = (p.Z + 2 + 15) / 16 * 16;
    p.HZ = p.ZP;
    p.ngrp = p.T / 3;
    p.KG = 512 / p.Cin;
    if (p.KG > p.ngrp) p.KG = p.ngrp;
    p.npass_t = (p.ngrp + p.KG - 1) / p.KG;
    p.TP = p.KG;                                  // accumulators per CTA
    p.MH = 1;
    p.PL = p.Cout / 8;
    p.MM = (3 * p.Cout <= 64) ? 64 : 128;
    if (p.kx == 3 && 3 * p.Cin <= 256 && 9 * p.Cin <= 512) {      // dx-folded (WgParams::dxf): three dy accumulators of 3*Cin columns
      p.dxf = 1;
      p.ngrp = 3; p.KG = 3; p.TP = 3; p.npass_t = 1;
    }
  }
  const int npass = p.npass_t * p.MH;
  const int acc_cols = p.dxf ? 3 * p.Cin : p.Cin;
  int cols = 32;
  while (cols < p.TP * acc_cols) cols *= 2;
  p.tmem_cols = cols;
  double best = 1e300;
  bool found = false;
  WgParams bp = p;
  const int Cib = p.Cin / 8;
  const int bxmax = (p.kx == 1 || p.dxf) ? 1 : p.X;            // dx-folded: HX must be 3 (plane stride = 3 x-plane strides)
  for (int BY = 1; BY <= p.Y && BY + 2 <= 256; ++BY) {
    for (int BX = 1; BX <= bxmax; ++BX) {
      const int HX = BX + p.kx - 1, HY = BY + 2;
      if (HX > 256) break;
      const long long rows_a = (long long)HX * HY * p.HZ, rows_dy = (long long)BX * BY * p.ZP;
      if (rows_a >= 16384 || rows_dy >= 16384) break;
      const long long a_tx = rows_a * 16 * Cib, dy_tx = rows_dy * 16 * p.PL * p.fz;
      const long long a_alloc = (a_tx + 256 + 127) / 128 * 128, dy_alloc = (rows_dy * 16 * (p.MM / 8) + 127) / 128 * 128;
      const long long slot = a_alloc + dy_alloc;
      int S = (int)((SMEM_BUDGET - 1024) / slot);
      if (S < 2) break;
      if (S > 4) S = 4;
      const int nbx = (p.X + BX - 1) / BX, nby = (p.Y + BY - 1) / BY;
      const long long nb = (long long)p.N * nbx * nby;
      long long splits = nsm / npass;
      if (max_splits > 0 && splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
      if (splits > nb) splits = nb;
      const long long per_cta = (nb + splits - 1) / splits;
      const double per_mma = (acc_cols / 2.0 > 32.0 + acc_cols / 4.0) ? acc_cols / 2.0 : 32.0 + acc_cols / 4.0;
      const double mma_cyc = (double)BX * BY * (p.ZP / 16) * p.TP * per_mma;
      const double load_cyc = (double)(a_tx + dy_tx) / 40.0;
      const double cost = (double)per_cta * ((mma_cyc > load_cyc ? mma_cyc : load_cyc) + 1200.0);
      if (cost < best) {
        best = cost; found = true; bp = p;
        bp.BX = BX; bp.BY = BY; bp.HX = HX; bp.HY = HY; bp.nbx = nbx; bp.nby = nby; bp.nbricks = (int)nb;
        bp.rows_a = (int)rows_a; bp.rows_dy = (int)rows_dy; bp.S = S; bp.splits = (int)splits;
        bp.a_tx_bytes = (unsigned)a_tx; bp.dy_tx_bytes = (unsigned)dy_tx; bp.a_alloc_bytes = (unsigned)a_alloc;
        bp.dy_alloc_bytes = (unsigned)dy_alloc; bp.slot_bytes = (unsigned)slot;
      }
    }
  }
  if (!found) return false;
  p = bp;
  p.offBar = p.S * p.slot_bytes;
  // flush mode for long accumulation chains (see WgParams::flush): 16 input channels, <= 9 accumulators, two sets in TMEM
  p.flush = 0;
  if (p.Cin == 16 && p.TP * acc_cols <= 144 && 2 * p.TP * acc_cols <= 512) {
    const long long vox_cta = (long long)p.N * p.X * p.Y * p.Z / p.splits;
    if (vox_cta > 4096) {
      const int ks = p.BX * p.BY * (p.ZP / 16);          // MMA K-steps per brick
      p.flush = 96 / ks > 1 ? 96 / ks : 1;
      int cols2 = 32;
      while (cols2 < 2 * p.TP * acc_cols) cols2 *= 2;
      p.tmem_cols = cols2;
    }
  }
  return true;
}

}  // namespace bcp

using namespace bcp;

extern "C" {

int bcp_conv_tc_supported(int cin, int cout, const int* dims, const int* kernel) {
  if (!dims || !kernel) return 0;
  if (!shape_ok(cin, cout, dims, kernel)) return 0;
  return get_encode() != nullptr ? 1 : 0;
}

// writes the chosen plan for inspection/tests: {BX,BY,BZ,MT,SA,SB,AS,nbricks,tmem_cols,smem_bytes}
int bcp_conv_tc_plan(int n, int cin, int cout, const int* dims, const int* kernel, int* plan10) {
  BCP_REQUIRE(dims && kernel && plan10, "conv_tc_plan: null pointer");
  if (!shape_ok(cin, cout, dims, kernel)) { set_last_error("conv_tc_plan: unsupported shape"); return BCP_ERR_UNSUPPORTED; }
  TcParams p{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = dims[2]; p.Cin = cin; p.Cout = cout; p.kx = kernel[0];
  if (!plan(p, sm_count())) { set_last_error("conv_tc_plan: no brick shape fits"); return BCP_ERR_UNSUPPORTED; }
  const int v[10] = {p.BX, p.BY, p.BZ, p.MT, p.SA, p.NS * 100 + p.TG, p.AS, p.nbricks, p.tmem_cols, (int)(p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS) + 16 + 128)};
  for (int i = 0; i < 10; ++i) plan10[i] = v[i];
  return BCP_OK;
}


// fused-statistics eligibility: big layers only (on the deep ones the standalone statistics launch is a few microseconds
// and the per-CTA partial reduce would cost as much), table of 4 x G x Cout x 2 doubles <= 16 KB
static unsigned stats_table_bytes(int n, int cout, const int* dims, int spg) {
  if (spg <= 0 || n % spg) return 0;
  const long long S = (long long)dims[0] * dims[1] * dims[2];
  const int G = n / spg;
  if (S < 8192 || G * cout > 128) return 0;
  return (unsigned)(4 * G * cout * 2 * sizeof(double));
}

static int conv_tc_launch(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                          const int* dims, const int* kernel, const StatsArgs* sa, unsigned long long* prof, cudaStream_t stream) {
  BCP_REQUIRE(in && wpack && out && dims && kernel, "conv_tc_fwd: null pointer");
  if (!shape_ok(cin, cout, dims, kernel)) { set_last_error("conv_tc_fwd: unsupported shape cin=%d cout=%d", cin, cout); return BCP_ERR_UNSUPPORTED; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("conv_tc_fwd: cuTensorMapEncodeTiled unavailable"); return BCP_ERR_CUDA; }
  TcParams p{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = dims[2]; p.Cin = cin; p.Cout = cout; p.kx = kernel[0];
  p.reserve = sa ? stats_table_bytes(n, cout, dims, sa->spg) + 128 : 0;
  const int nsm = sm_count();
  if (!plan_cached(p, nsm)) { set_last_error("conv_tc_fwd: no brick shape fits shared memory / TMEM"); return BCP_ERR_UNSUPPORTED; }

  CUtensorMap tmap;
  const CUresult cr = encode_cb8(enc, &tmap, in, p.Z, p.Y, p.X, (long long)n * (cin / 8), p.HZ, p.HY, p.HX, 2, &p.mergedA);
  if (cr != CUDA_SUCCESS) { set_last_error("conv_tc_fwd: cuTensorMapEncodeTiled failed (%d)", (int)cr); return BCP_ERR_CUDA; }

  CUtensorMap tmap_w;
  {
    // pack [T][Cin/8][Cout][8]: the Ns-channel run of a (tap, plane) is contiguous, so fold (cout, 8) into 8-byte elements
    p.mergedB = (p.Ns <= 128) ? 1 : 0;
    CUresult cw;
    if (p.mergedB) {
      const cuuint64_t wdim[3] = {(cuuint64_t)cout * 2, (cuuint64_t)(cin / 8), (cuuint64_t)p.T};
      const cuuint64_t wstr[2] = {(cuuint64_t)cout * 16, (cuuint64_t)cout * 16 * (cin / 8)};
      const cuuint32_t wbox[3] = {(cuuint32_t)p.Ns * 2, 2, (cuuint32_t)p.TG};
      const cuuint32_t wes[3] = {1, 1, 1};
      cw = enc(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(wpack), wdim, wstr, wbox, wes, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t wdim[4] = {8, (cuuint64_t)cout, (cuuint64_t)(cin / 8), (cuuint64_t)p.T};
      const cuuint64_t wstr[3] = {16, (cuuint64_t)cout * 16, (cuuint64_t)cout * 16 * (cin / 8)};
      const cuuint32_t wbox[4] = {8, (cuuint32_t)p.Ns, 2, (cuuint32_t)p.TG};
      const cuuint32_t wes[4] = {1, 1, 1, 1};
      cw = enc(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(wpack), wdim, wstr, wbox, wes, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (cw != CUDA_SUCCESS) { set_last_error("conv_tc_fwd: weight tensor map failed (%d)", (int)cw); return BCP_ERR_CUDA; }
  }
  size_t smem = (size_t)p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS) + 16 + 128;
  if (sa) {
    p.offStats = (unsigned)((p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS) + 16 + 127) / 128 * 128);
    smem = (size_t)p.offStats + (p.reserve - 128) + 128;
  }
  // dynamic + static shared memory share the 227 KB opt-in limit; the statistics variant has a few static bytes, so
  // every variant opts into 226 KB of dynamic memory (plans stay below SMEM_BUDGET + barriers ~ 221 KB)
  constexpr int kMaxDyn = 226 * 1024;
  BCP_REQUIRE(smem <= (size_t)kMaxDyn, "conv_tc_fwd: shared memory plan overflow (%zu bytes)", smem);
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
    attr_err = e;
    if (e != cudaSuccess) cudaGetLastError();
  });
  if (attr_err != cudaSuccess) { set_last_error("conv_tc_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err)); return BCP_ERR_CUDA; }
  const int nitems = p.nbricks * p.NS;
  const int grid = nitems < nsm ? nitems : nsm;
  if (sa) conv_tc_kernel<false, true><<<grid, TC_THREADS, smem, stream>>>(tmap, tmap_w, bias, (uint4*)out, p, nullptr, *sa);
  else if (prof) conv_tc_kernel<true, false><<<grid, TC_THREADS, smem, stream>>>(tmap, tmap_w, bias, (uint4*)out, p, prof, StatsArgs{});
  else conv_tc_kernel<false, false><<<grid, TC_THREADS, smem, stream>>>(tmap, tmap_w, bias, (uint4*)out, p, nullptr, StatsArgs{});
  return check_launch("conv_tc_fwd");
}

int bcp_conv_tc_fwd(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                    const int* dims, const int* kernel, cudaStream_t stream) {
  return conv_tc_launch(in, wpack, bias, out, n, cin, cout, dims, kernel, nullptr, nullptr, stream);
}

long long bcp_conv_tc_stats_workspace_bytes(int n, int cin, int cout, const int* dims, const int* kernel, int spg) {
  if (!dims || !kernel || !shape_ok(cin, cout, dims, kernel) || get_encode() == nullptr) return 0;
  const unsigned tb = stats_table_bytes(n, cout, dims, spg);
  if (!tb) return 0;
  TcParams p{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = dims[2]; p.Cin = cin; p.Cout = cout; p.kx = kernel[0];
  p.reserve = tb + 128;
  if (!plan_cached(p, sm_count())) return 0;
  return (long long)sm_count() * (n / spg) * cout * 2 * (long long)sizeof(double);
}

int bcp_conv_tc_fwd_stats(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                          const int* dims, const int* kernel, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, long long* num_batches_tracked, float* stat, float* coef, void* workspace,
                          int* counter, int spg, float eps, float momentum, cudaStream_t stream) {
  BCP_REQUIRE(stat && coef && workspace && counter && dims, "conv_tc_fwd_stats: null pointer");
  if (!stats_table_bytes(n, cout, dims, spg)) { set_last_error("conv_tc_fwd_stats: layer not eligible for fused statistics"); return BCP_ERR_UNSUPPORTED; }
  StatsArgs sa{};
  sa.partial = (double*)workspace; sa.counter = counter; sa.gamma = gamma; sa.beta = beta;
  sa.running_mean = running_mean; sa.running_var = running_var; sa.nbt = num_batches_tracked;
  sa.stat = stat; sa.coef = coef; sa.spg = spg; sa.G = n / spg; sa.eps = eps; sa.momentum = momentum;
  return conv_tc_launch(in, wpack, bias, out, n, cin, cout, dims, kernel, &sa, nullptr, stream);
}

// ---- dz-folded forward (conv_tc_fold_kernel): brick plan + launch
static bool fold_shape_ok(int cin, int cout, const int* dims, const int* kernel) {
  if (!shape_ok(cin, cout, dims, kernel)) return false;
  return cout == 16 || cout == 32;                 // N = 3*Cout <= 96; wider layers are already near the tensor rate
}

static bool fold_plan(FoldParams& p, int nsm) {
  const int T = p.kx * 9;
  p.ngrp = T / 3;
  p.nchunks = p.Cin / 16;
  p.NS = 1; p.Ns = p.Cout;
  const int N3 = 3 * p.Ns;
  const unsigned stageB = (unsigned)p.ngrp * 2u * (unsigned)N3 * 16u;
  const double per_mma = (N3 / 2.0 > 32.0 + N3 / 4.0) ? N3 / 2.0 : 32.0 + N3 / 4.0;
  double best = 1e300;
  bool found = false;
  FoldParams bp = p;
  int bz_opts[4];
  int nz = 0;
  for (int d = 1; d <= 8 && nz < 4; d *= 2) {
    const int bz = (p.Z + d - 1) / d;
    if (bz + 2 <= 256 && (nz == 0 || bz_opts[nz - 1] != bz)) bz_opts[nz++] = bz;
  }
  for (int zi = 0; zi < nz; ++zi) {
    const int BZ = bz_opts[zi], HZ = BZ + 2;
    for (int BY = 1; BY <= p.Y && BY + 2 <= 256; ++BY) {
      const int HY = BY + 2;
      const int bxmax = (p.kx == 1) ? 1 : p.X;
      for (int BX = 1; BX <= bxmax; ++BX) {
        const int HX = BX + p.kx - 1;
        const long long rows_h = (long long)HX * HY * HZ;
        if (rows_h >= 16384) break;
        const long long lmax = ((long long)(BX - 1) * HY + (BY - 1)) * HZ + BZ;
        const int MT = (int)((lmax + 95) / 96);
        int AS = 2;
        if (2 * MT * N3 > 512) AS = 1;
        if (MT * N3 > 512) break;
        // rows the MMAs may touch: last tile start + 98 rows + the largest (dx,dy) shift
        const long long rows_alloc = 96ll * (MT - 1) + 98 + ((long long)(p.kx - 1) * HY + 2) * HZ;
        const long long slotA = ((rows_h + (rows_alloc > rows_h ? rows_alloc : rows_h)) * 16 + 127) / 128 * 128;
        int SB = 2;
        int SA = (int)(((long long)SMEM_BUDGET - 1024 - (long long)SB * stageB) / slotA);
        if (SA < 2) break;
        if (SA > 3) SA = 3;
        SB = (int)(((long long)SMEM_BUDGET - 1024 - (long long)SA * slotA) / stageB);
        if (SB > 4) SB = 4;
        const int nbx = (p.X + BX - 1) / BX, nby = (p.Y + BY - 1) / BY, nbz = (p.Z + BZ - 1) / BZ;
        const long long nb = (long long)p.N * nbx * nby * nbz;
        const long long waves = (nb + nsm - 1) / nsm;
        const double mma_cyc = (double)MT * p.ngrp * p.nchunks * per_mma;
        const double epi_cyc = (double)MT * (p.Ns / 16) * 900.0;            // three TMEM loads + 32 shuffles per 16 channels
        const double item_bytes = (double)rows_h * p.Cin * 2.0 + (double)stageB * p.nchunks;
        const double load_cyc = item_bytes / 40.0;
        double per_item = mma_cyc > load_cyc ? mma_cyc : load_cyc;
        per_item = (AS == 2) ? (per_item > epi_cyc ? per_item : epi_cyc) : per_item + epi_cyc;
        const double l2_bound = (double)nb * item_bytes / 2500.0;
        const double t_sm = (double)waves * (per_item + 1500.0);
        const double cost = t_sm > l2_bound ? t_sm : l2_bound;
        if (cost < best) {
          best = cost; found = true; bp = p;
          bp.BX = BX; bp.BY = BY; bp.BZ = BZ; bp.HX = HX; bp.HY = HY; bp.HZ = HZ;
          bp.nbx = nbx; bp.nby = nby; bp.nbz = nbz; bp.nbricks = (int)nb;
          bp.rows_h = (int)rows_h; bp.MT = MT; bp.SA = SA; bp.SB = SB; bp.AS = AS;
          bp.slotA_bytes = (unsigned)slotA; bp.stageB_bytes = stageB;
        }
      }
    }
  }
  if (!found) return false;
  p = bp;
  int cols = 32;
  while (cols < p.AS * p.MT * N3) cols *= 2;
  p.tmem_cols = cols;
  p.offA = 0;
  p.offB = p.SA * p.slotA_bytes;
  p.offBar = p.offB + p.SB * p.stageB_bytes;
  return true;
}

int bcp_conv_tc_fold_supported(int cin, int cout, const int* dims, const int* kernel) {
  if (!dims || !kernel || !fold_shape_ok(cin, cout, dims, kernel)) return 0;
  return get_encode() != nullptr ? 1 : 0;
}

// chosen tiling for inspection: {BX,BY,BZ,MT,SA,SB,AS,nbricks,tmem_cols,smem_bytes}
int bcp_conv_tc_fold_plan(int n, int cin, int cout, const int* dims, const int* kernel, int* plan10) {
  BCP_REQUIRE(dims && kernel && plan10, "conv_tc_fold_plan: null pointer");
  if (!fold_shape_ok(cin, cout, dims, kernel)) { set_last_error("conv_tc_fold_plan: unsupported shape"); return BCP_ERR_UNSUPPORTED; }
  FoldParams p{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = dims[2]; p.Cin = cin; p.Cout = cout; p.kx = kernel[0];
  if (!fold_plan(p, sm_count())) { set_last_error("conv_tc_fold_plan: no brick shape fits"); return BCP_ERR_UNSUPPORTED; }
  const int v[10] = {p.BX, p.BY, p.BZ, p.MT, p.SA, p.SB, p.AS, p.nbricks, p.tmem_cols, (int)(p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS) + 16 + 128)};
  for (int i = 0; i < 10; ++i) plan10[i] = v[i];
  return BCP_OK;
}

static int conv_tc_fold_launch(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                               const int* dims, const int* kernel, unsigned long long* prof, int ew, cudaStream_t stream) {
  BCP_REQUIRE(in && wpack && out && dims && kernel, "conv_tc_fold_fwd: null pointer");
  if (!fold_shape_ok(cin, cout, dims, kernel)) { set_last_error("conv_tc_fold_fwd: unsupported shape"); return BCP_ERR_UNSUPPORTED; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("conv_tc_fold_fwd: cuTensorMapEncodeTiled unavailable"); return BCP_ERR_CUDA; }
  FoldParams p{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = dims[2]; p.Cin = cin; p.Cout = cout; p.kx = kernel[0];
  const int nsm = sm_count();
  if (!fold_plan(p, nsm)) { set_last_error("conv_tc_fold_fwd: no brick shape fits"); return BCP_ERR_UNSUPPORTED; }
  CUtensorMap tmap, tmap_w;
  CUresult cr = encode_cb8(enc, &tmap, in, p.Z, p.Y, p.X, (long long)n * (cin / 8), p.HZ, p.HY, p.HX, 2, &p.mergedA);
  if (cr != CUDA_SUCCESS) { set_last_error("conv_tc_fold_fwd: tensor map failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  {
    // pack [T][Cin/8][Cout][8] bf16 with t = g*3 + dz, seen as 8-byte elements: dims (inner -> outer) cout*2, dz, plane, g
    // (the dz stride is larger than the plane stride; if the driver rejects non-monotonic strides, add a repack kind
    // that stores [g][Cin/8][dz][Cout][8] instead)
    const cuuint64_t tap_bytes = (cuuint64_t)(cin / 8) * cout * 16;
    const cuuint64_t wdim[4] = {(cuuint64_t)cout * 2, 3, (cuuint64_t)(cin / 8), (cuuint64_t)p.ngrp};
    const cuuint64_t wstr[3] = {tap_bytes, (cuuint64_t)cout * 16, 3 * tap_bytes};
    const cuuint32_t wbox[4] = {(cuuint32_t)p.Ns * 2, 3, 2, (cuuint32_t)p.ngrp};
    const cuuint32_t wes[4] = {1, 1, 1, 1};
    cr = enc(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(wpack), wdim, wstr, wbox, wes, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_last_error("conv_tc_fold_fwd: weight tensor map failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS) + 16 + 128;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_fold_kernel<false, FOLD_EPI_WARPS_DEFAULT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(conv_tc_fold_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(conv_tc_fold_kernel<true, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(conv_tc_fold_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (attr_err != cudaSuccess) cudaGetLastError();
  });
  if (attr_err != cudaSuccess) { set_last_error("conv_tc_fold_fwd: cudaFuncSetAttribute failed"); return BCP_ERR_CUDA; }
  BCP_REQUIRE(smem <= 226 * 1024, "conv_tc_fold_fwd: shared memory plan overflow");
  const int grid = p.nbricks < nsm ? p.nbricks : nsm;
  if (prof && ew == 8) conv_tc_fold_kernel<true, 8><<<grid, 64 + 32 * 8, smem, stream>>>(tmap, tmap_w, bias, (uint4*)out, p, prof);
  else if (prof && ew == 12) conv_tc_fold_kernel<true, 12><<<grid, 64 + 32 * 12, smem, stream>>>(tmap, tmap_w, bias, (uint4*)out, p, prof);
  else if (prof) conv_tc_fold_kernel<true, 16><<<grid, 64 + 32 * 16, smem, stream>>>(tmap, tmap_w, bias, (uint4*)out, p, prof);
  else conv_tc_fold_kernel<false, FOLD_EPI_WARPS_DEFAULT><<<grid, 64 + 32 * FOLD_EPI_WARPS_DEFAULT, smem, stream>>>(tmap, tmap_w, bias, (uint4*)out, p, nullptr);
  return check_launch("conv_tc_fold_fwd");
}

int bcp_conv_tc_fold_fwd(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                         const int* dims, const int* kernel, cudaStream_t stream) {
  return conv_tc_fold_launch(in, wpack, bias, out, n, cin, cout, dims, kernel, nullptr, FOLD_EPI_WARPS_DEFAULT, stream);
}

// Instrumented variants for tools/debug_conv_tc.py: `prof` = device buffer of 16 x uint64 per CTA (>= #SMs CTAs) that
// receives per-role wait cycles.  fold = 0: unfolded kernel; fold = 8 / 12 / 16: the dz-folded kernel with that many epilogue warps.  The caller passes the buffer explicitly: the
// library keeps no profiling state.
int bcp_conv_tc_fwd_profiled(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                             const int* dims, const int* kernel, int fold, void* prof, cudaStream_t stream) {
  BCP_REQUIRE(prof, "conv_tc_fwd_profiled: null profile buffer");
  if (fold) return conv_tc_fold_launch(in, wpack, bias, out, n, cin, cout, dims, kernel, (unsigned long long*)prof, fold, stream);
  return conv_tc_launch(in, wpack, bias, out, n, cin, cout, dims, kernel, nullptr, (unsigned long long*)prof, stream);
}


// Launch of the weight-gradient kernel (its in-kernel finalize needs every CTA of the grid co-resident: the grid is
// splits x passes <= #SMs with one CTA per SM).
static int wg_launch(const CUtensorMap& map_a, const CUtensorMap& map_dy, float* workspace, float* dw, int* counter, WgParams& p,
                     cudaStream_t stream, const char* what) {
  const size_t smem = (size_t)p.offBar + 8 * (2 * p.S + 1) + 16 + 128 + 128 + 64;   // barriers, TMEM slot, tap table, flush barriers, alignment slack
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err != cudaSuccess) cudaGetLastError();
  });
  if (attr_err != cudaSuccess) { set_last_error("%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(attr_err)); return BCP_ERR_CUDA; }
  if (p.NH < 1) p.NH = 1;
  const int nct = p.splits * p.npass_t * p.MH * p.NH;
  if (nct > sm_count()) { set_last_error("%s: grid of %d CTAs cannot be co-resident", what, nct); return BCP_ERR_UNSUPPORTED; }
  // lanes per output quad: spread the `splits` partials of an output over KS lanes when there are more threads than quads
  const long long items = (long long)p.T * p.Cout * p.Cin / 4, threads = (long long)nct * TC_THREADS;
  int ks = 1;
  while (ks < 32 && ks * 2 <= p.splits && items * ks * 2 <= threads) ks *= 2;
  p.KS = ks;
  // Plain launch: the device-wide barrier inside the kernel needs every CTA co-resident, which holds because the grid is at most
  // one CTA per SM (220 KB of shared memory each) and a stream-ordered launch starts on an idle device.  (A cooperative
  // launch would make the driver guarantee it, but measured ~10 us of extra launch latency per call inside a CUDA graph:
  // 28 calls per step.)  Concurrent kernels of other streams only delay the barrier, they cannot deadlock it, as long as
  // they terminate on their own.
  dim3 grid(p.splits, p.npass_t * p.MH * p.NH);
  conv_tc_wgrad_kernel<<<grid, TC_THREADS, smem, stream>>>(map_a, map_dy, workspace, dw, counter, p);
  return check_launch(what);
}

int bcp_conv_tc_wgrad_supported(int cin, int cout, const int* dims, const int* kernel) {
  if (!dims || !kernel) return 0;
  if (!wg_shape_ok(cin, cout, dims, kernel)) return 0;
  return get_encode() != nullptr ? 1 : 0;
}

static int wg_setup(WgParams& p, int n, int cin, int cout, const int* dims, const int* kernel, int max_splits = 0) {
  p = WgParams{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = wg_tile_width(dims[2]); p.Cin = cin; p.Cout = cout; p.kx = kernel[0];
  if (!wg_plan(p, sm_count(), max_splits)) return -1;
  if (wg_ztiles(dims[2]) > 1 && p.s2 != 0) return -1;          // windows exist for the halo-brick path only
  return 0;
}

long long bcp_conv_tc_wgrad_workspace_floats(int n, int cin, int cout, const int* dims, const int* kernel) {
  if (!dims || !kernel || !wg_shape_ok(cin, cout, dims, kernel)) return 0;
  WgParams p{};
  if (wg_setup(p, n, cin, cout, dims, kernel) != 0) return 0;
  return (long long)p.splits * p.T * cout * cin;
}

// dw[cout][cin][T] fp32; `a` = layer input (cin channels), `dy` = output gradient (cout channels), both CB8 at `dims`
static int conv_tc_wgrad_impl(const void* a, const void* dy, float* dw, float* workspace, int* counter, int n, int cin, int cout,
                              const int* dims, const int* kernel, int accumulate, int max_splits, cudaStream_t stream);

int bcp_conv_tc_wgrad(const void* a, const void* dy, float* dw, float* workspace, int* counter, int n, int cin, int cout,
                      const int* dims, const int* kernel, int accumulate, cudaStream_t stream) {
  return conv_tc_wgrad_impl(a, dy, dw, workspace, counter, n, cin, cout, dims, kernel, accumulate, 0, stream);
}

// tools/debug_conv_tc.py --wgrad-splits: the same launch with the split-K factor capped (plan exploration; workspace as above)
int bcp_conv_tc_wgrad_capped(const void* a, const void* dy, float* dw, float* workspace, int* counter, int n, int cin, int cout,
                             const int* dims, const int* kernel, int accumulate, int max_splits, cudaStream_t stream) {
  return conv_tc_wgrad_impl(a, dy, dw, workspace, counter, n, cin, cout, dims, kernel, accumulate, max_splits, stream);
}

static int conv_tc_wgrad_impl(const void* a, const void* dy, float* dw, float* workspace, int* counter, int n, int cin, int cout,
                              const int* dims, const int* kernel, int accumulate, int max_splits, cudaStream_t stream) {
  BCP_REQUIRE(a && dy && dw && workspace && counter && dims && kernel, "conv_tc_wgrad: null pointer");
  if (!wg_shape_ok(cin, cout, dims, kernel)) { set_last_error("conv_tc_wgrad: unsupported shape"); return BCP_ERR_UNSUPPORTED; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("conv_tc_wgrad: cuTensorMapEncodeTiled unavailable"); return BCP_ERR_CUDA; }
  WgParams p{};
  if (wg_setup(p, n, cin, cout, dims, kernel, max_splits) != 0) { set_last_error("conv_tc_wgrad: no brick shape fits"); return BCP_ERR_UNSUPPORTED; }
  CUtensorMap map_a, map_dy;
  const int Zfull = dims[2], nt = wg_ztiles(Zfull), tw = p.Z;
  {
    const CUresult cr = (p.s2 == 2)
        ? encode_cb8(enc, &map_a, a, Zfull, p.Y, p.X, (long long)n * (cin / 8), p.Z, p.BY, p.BX, (cin / 8) / p.NH, &p.mergedA)
        : encode_cb8(enc, &map_a, a, Zfull, p.Y, p.X, (long long)n * (cin / 8), p.HZ, p.HY, p.HX, cin / 8, &p.mergedA);
    if (cr != CUDA_SUCCESS) { set_last_error("conv_tc_wgrad: tensor map (a) failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  }
  for (int t = 0; t < nt; ++t) {
    const int zoff = t * tw;
    const int ext = (Zfull - zoff < tw) ? Zfull - zoff : tw;
    const CUresult cr = encode_cb8(enc, &map_dy, (const char*)dy + (size_t)zoff * 16, ext, p.Y, p.X, (long long)n * (cout / 8), p.ZP, p.BY,
                                   p.BX, p.PL, &p.mergedD, Zfull);
    if (cr != CUDA_SUCCESS) { set_last_error("conv_tc_wgrad: tensor map (dy) failed (%d)", (int)cr); return BCP_ERR_CUDA; }
    p.zoff = zoff;
    p.accumulate = (t == 0) ? accumulate : 1;
    const int rc = wg_launch(map_a, map_dy, workspace, dw, counter, p, stream, "conv_tc_wgrad");
    if (rc != 0) return rc;
  }
  return BCP_OK;
}

// ---- stride-2 family.  half_dims = dims of the half-resolution grid (full = 2x).  mode 1 (gather): `in` is full-res with
// cin channels, wpack = kind-0 pack, out half-res cout channels.  mode 2 (scatter): `in` is half-res with cin channels,
// wpack = kind-3 pack [cin/8][8][cout][8], out full-res cout channels.
int bcp_conv_tc_s2_supported(int cin, int cout, const int* half_dims, int mode) {
  if (!half_dims || (mode != 1 && mode != 2)) return 0;
  if (!s2_shape_ok(cin, cout, half_dims)) return 0;
  S2Params p{};
  p.N = 1; p.Cin = cin; p.Cout = cout; p.mode = mode; p.Xh = half_dims[0]; p.Yh = half_dims[1]; p.Zh = half_dims[2];
  if (!s2_plan(p)) return 0;
  return get_encode() != nullptr ? 1 : 0;
}

int bcp_conv_tc_s2_fwd(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                       const int* half_dims, int mode, cudaStream_t stream) {
  BCP_REQUIRE(in && wpack && out && half_dims, "conv_tc_s2_fwd: null pointer");
  BCP_REQUIRE(mode == 1 || mode == 2, "conv_tc_s2_fwd: mode");
  if (!s2_shape_ok(cin, cout, half_dims)) { set_last_error("conv_tc_s2_fwd: unsupported shape"); return BCP_ERR_UNSUPPORTED; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("conv_tc_s2_fwd: cuTensorMapEncodeTiled unavailable"); return BCP_ERR_CUDA; }
  S2Params p{};
  p.N = n; p.Cin = cin; p.Cout = cout; p.mode = mode; p.Xh = half_dims[0]; p.Yh = half_dims[1]; p.Zh = half_dims[2];
  if (!s2_plan(p)) { set_last_error("conv_tc_s2_fwd: no tiling fits"); return BCP_ERR_UNSUPPORTED; }
  const int f = (mode == 1) ? 2 : 1;      // input resolution factor
  const cuuint64_t X = (cuuint64_t)p.Xh * f, Y = (cuuint64_t)p.Yh * f, Z = (cuuint64_t)p.Zh * f;
  CUtensorMap tmap;
  CUresult cr;
  if (mode == 2) {
    cr = encode_cb8(enc, &tmap, in, (long long)Z, (long long)Y, (long long)X, (long long)n * (cin / 8), p.BZ, p.BY, p.BX, 2, &p.mergedA);
  } else {
    p.mergedA = 0;      // stride-2 gather: elementStrides select every other 16-byte voxel, not expressible on a merged dimension
    const cuuint64_t gdim[5] = {8, Z, Y, X, (cuuint64_t)n * (cin / 8)};
    const cuuint64_t gstr[4] = {16, Z * 16, Z * Y * 16, Z * Y * X * 16};
    const cuuint32_t box[5] = {8, (cuuint32_t)(p.BZ * f), (cuuint32_t)(p.BY * f), (cuuint32_t)(p.BX * f), 2};
    const cuuint32_t estr[5] = {1, (cuuint32_t)f, (cuuint32_t)f, (cuuint32_t)f, 1};
    cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(in), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (cr != CUDA_SUCCESS) { set_last_error("conv_tc_s2_fwd: cuTensorMapEncodeTiled failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  const size_t smem = (size_t)p.offBar + 8 * (2 * p.SA + 2 * p.SB + 2 * p.AS) + 16 + 128;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] { cudaFuncSetAttribute(conv_tc_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  const int nsm = sm_count();
  const int ntiles = p.nbricks * p.npieces;
  conv_tc_s2_kernel<<<ntiles < nsm ? ntiles : nsm, TC_THREADS, smem, stream>>>(tmap, (const __nv_bfloat16*)wpack, bias, (uint4*)out, p);
  return check_launch("conv_tc_s2_fwd");
}

// dw[c_half][c_full][8] = sum_i half[i][c_half] * full[2i+t][c_full]   (Conv3d k2s2: half = dy, full = input;
// ConvTranspose3d k2s2: half = input, full = dy)
static bool wg_s2_plan(WgParams& p, int nsm) {
  p.T = 8;
  p.TP = 512 / p.Cin; if (p.TP > 8) p.TP = 8; if (p.TP < 1) return false;
  p.npass_t = (8 + p.TP - 1) / p.TP;
  p.MH = (p.Cout > 128) ? 2 : 1;
  p.PL = (p.Cout / 8 < 16) ? p.Cout / 8 : 16;
  p.MM = (p.Cout <= 64 && wg_m64_mode() >= 0) ? 64 : 128;
  p.m64map = wg_m64_mode() < 0 ? 0 : wg_m64_mode();
  p.ZP = (p.Z + 15) / 16 * 16;
  if (p.ZP > 128) return false;
  p.HX = p.HY = p.HZ = 0; p.s2 = 1; p.kx = 2;
  int cols = 32;
  while (cols < p.TP * p.Cin) cols *= 2;
  p.tmem_cols = cols;
  const int npass = p.npass_t * p.MH, Cib = p.Cin / 8;
  double best = 1e300; bool found = false; WgParams bp = p;
  for (int BY = 1; BY <= p.Y && BY <= 128; ++BY) {
    for (int BX = 1; BX <= p.X && BX <= 128; ++BX) {
      const long long rows = (long long)BX * BY * p.ZP;
      if (rows >= 16384) break;
      const long long tap_bytes = rows * 16 * Cib, a_alloc = ((long long)p.TP * tap_bytes + 127) / 128 * 128;
      const long long dy_tx = rows * 16 * p.PL, dy_alloc = (rows * 16 * (p.MM / 8) + 127) / 128 * 128;
      const long long slot = a_alloc + dy_alloc;
      int S = (int)((SMEM_BUDGET - 1024) / slot);
      if (S < 2) break;
      if (S > 4) S = 4;
      const int nbx = (p.X + BX - 1) / BX, nby = (p.Y + BY - 1) / BY;
      const long long nb = (long long)p.N * nbx * nby;
      long long splits = nsm / npass; if (splits < 1) splits = 1; if (splits > nb) splits = nb;
      const long long per_cta = (nb + splits - 1) / splits;
      const double per_mma = (p.Cin / 2.0 > 32.0 + p.Cin / 4.0) ? p.Cin / 2.0 : 32.0 + p.Cin / 4.0;
      const double mma_cyc = (double)BX * BY * (p.ZP / 16) * p.TP * per_mma;
      const double load_cyc = (double)(p.TP * tap_bytes + dy_tx) / 40.0;
      const double cost = (double)per_cta * ((mma_cyc > load_cyc ? mma_cyc : load_cyc) + 1200.0);
      if (cost < best) {
        best = cost; found = true; bp = p;
        bp.BX = BX; bp.BY = BY; bp.nbx = nbx; bp.nby = nby; bp.nbricks = (int)nb;
        bp.rows_a = (int)rows; bp.rows_dy = (int)rows; bp.S = S; bp.splits = (int)splits;
        bp.tap_bytes = (unsigned)tap_bytes; bp.a_tx_bytes = (unsigned)a_alloc; bp.a_alloc_bytes = (unsigned)a_alloc;
        bp.dy_tx_bytes = (unsigned)dy_tx; bp.dy_alloc_bytes = (unsigned)dy_alloc; bp.slot_bytes = (unsigned)slot;
      }
    }
  }
  if (!found) return false;
  p = bp;
  p.offBar = p.S * p.slot_bytes;
  return true;
}

static int wg_s2_setup(WgParams& p, int n, int c_half, int c_full, const int* half_dims) {
  p = WgParams{};
  p.N = n; p.X = half_dims[0]; p.Y = half_dims[1]; p.Z = half_dims[2]; p.Cin = c_full; p.Cout = c_half;
  return wg_s2_plan(p, sm_count()) ? 0 : -1;
}

int bcp_conv_tc_s2_wgrad_supported(int c_half, int c_full, const int* half_dims) {
  if (!half_dims || !s2_shape_ok(c_full, c_half, half_dims)) return 0;
  WgParams p{};
  if (wg_s2_setup(p, 1, c_half, c_full, half_dims) != 0) return 0;
  return get_encode() != nullptr ? 1 : 0;
}

long long bcp_conv_tc_s2_wgrad_workspace_floats(int n, int c_half, int c_full, const int* half_dims) {
  if (!half_dims || !s2_shape_ok(c_full, c_half, half_dims)) return 0;
  WgParams p{};
  if (wg_s2_setup(p, n, c_half, c_full, half_dims) != 0) return 0;
  return (long long)p.splits * 8 * c_half * c_full;
}

int bcp_conv_tc_s2_wgrad(const void* full, const void* half, float* dw, float* workspace, int* counter, int n, int c_half,
                         int c_full, const int* half_dims, int accumulate, cudaStream_t stream) {
  BCP_REQUIRE(full && half && dw && workspace && counter && half_dims, "conv_tc_s2_wgrad: null pointer");
  if (!s2_shape_ok(c_full, c_half, half_dims)) { set_last_error("conv_tc_s2_wgrad: unsupported shape"); return BCP_ERR_UNSUPPORTED; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("conv_tc_s2_wgrad: cuTensorMapEncodeTiled unavailable"); return BCP_ERR_CUDA; }
  WgParams p{};
  if (wg_s2_setup(p, n, c_half, c_full, half_dims) != 0) { set_last_error("conv_tc_s2_wgrad: no tiling fits"); return BCP_ERR_UNSUPPORTED; }
  CUtensorMap map_a, map_dy;
  {
    const cuuint64_t X = 2ull * p.X, Y = 2ull * p.Y, Z = 2ull * p.Z;
    const cuuint64_t gdim[5] = {8, Z, Y, X, (cuuint64_t)n * (c_full / 8)};
    const cuuint64_t gstr[4] = {16, Z * 16, Z * Y * 16, Z * Y * X * 16};
    const cuuint32_t box[5] = {8, (cuuint32_t)(2 * p.ZP), (cuuint32_t)(2 * p.BY), (cuuint32_t)(2 * p.BX), (cuuint32_t)(c_full / 8)};
    const cuuint32_t estr[5] = {1, 2, 2, 2, 1};
    const CUresult cr = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(full), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_last_error("conv_tc_s2_wgrad: tensor map (full) failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  }
  {
    const CUresult cr = encode_cb8(enc, &map_dy, half, p.Z, p.Y, p.X, (long long)n * (c_half / 8), p.ZP, p.BY, p.BX, p.PL, &p.mergedD);
    if (cr != CUDA_SUCCESS) { set_last_error("conv_tc_s2_wgrad: tensor map (half) failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  }
  p.accumulate = accumulate;
  return wg_launch(map_a, map_dy, workspace, dw, counter, p, stream, "conv_tc_s2_wgrad");
}

}  // extern "C"
