// First-layer (Cin = 1) kernels fed by TMA: the operands of a slab are staged in shared memory by bulk tensor copies
// (out-of-bounds zero fill = the convolution's padding in x, y AND z), the CUDA cores only run the multiply-adds.
//
// Weight gradient (networks/VNet.py:151 block_one, networks/unet.py:72 in_conv -- autograd's dW of that Conv3d/Conv2d):
//   dW[co][tx][ty][tz] = sum_{n,x,y,z} dy[n][co][x][y][z] * in[n][x+tx-px][y+ty-1][z+tz-1]
// This launch is the LAST weight gradient of a backward pass (nothing is left to overlap it with), so it has to use the
// whole GPU on its own.  The earlier kernel (conv_direct.cu conv_first_wgrad_partial_kernel) fetched every operand with
// per-thread global loads one step ahead and sat at ~17 % of the fp32 rate, latency-bound.  Here:
//   * slab = (n, x, YR rows, ZT columns); one producer warp issues two TMA box loads per slab -- the dy slab for all
//     output-channel octets (CB8, 16 B per voxel) and the input planes x-px..x+px with a one-voxel halo -- into a ring of
//     stages guarded by full/empty mbarriers;
//   * NW consumer warps = (channel octet, tx) combinations x S splits; a warp keeps its 9 x 8 accumulators in registers
//     for the whole kernel and walks the slab 32 flat voxels per step: one 16-byte dy load, nine 4-byte window loads
//     (conflict-free, consecutive lanes = consecutive z), 72 FMAs;
//   * one partial [octet][tx][9][8] per CTA, summed across CTAs in fixed order by conv_first_wgrad_finalize_kernel
//     (deterministic: static slab -> CTA assignment, fixed shuffle / split order).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/bcp_b200.h"
#include <mutex>

namespace bcp {

// The input box starts FIRST_ZLEAD floats before the slab's first z: TMA wants the innermost start coordinate on a 16-byte
// boundary (a start at z0 - 1 = -4 bytes is an illegal instruction -- tools/micro/tma_probe.cu), so the z = -1 halo column
// sits at row offset FIRST_ZLEAD - 1 = 3 and a window row is ZT + 8 floats long.
constexpr int FIRST_ZLEAD = 4;
constexpr int FW_NW = 12;                         // consumer warps
constexpr int FW_THREADS = 32 * (FW_NW + 1);      // + the producer warp
constexpr unsigned FW_SMEM_BUDGET = 200 * 1024;


// (n, x, y tile, z tile) of a slab index, advanced by the grid stride with carries instead of three divisions per slab
struct SlabPos {
  int zt, yt, x, n;
  __device__ __forceinline__ void set(int slab, int nzt, int nyt, int X) {
    zt = slab % nzt; slab /= nzt;
    yt = slab % nyt; slab /= nyt;
    x = slab % X;
    n = slab / X;
  }
  __device__ __forceinline__ void advance(const SlabPos& d, int nzt, int nyt, int X) {
    zt += d.zt; if (zt >= nzt) { zt -= nzt; ++yt; }
    yt += d.yt; if (yt >= nyt) { yt -= nyt; ++x; }
    x += d.x;   if (x >= X) { x -= X; ++n; }
    n += d.n;
  }
};

struct FirstWgParams {
  int N, X, Y, Z, kx, Cob;
  int YR, ZT, ZP;            // slab rows / columns, padded window row length (floats)
  int nyt, nzt, nslabs;
  int nvox, nsteps;          // voxels per slab, 32-voxel steps per slab
  int S;                     // splits per (octet, tx) combination: FW_NW / (Cob * kx)
  int NSTG;
  unsigned x_box_bytes, x_bytes, dy_bytes, stage_bytes;   // exact TMA box bytes / 128-byte-padded offset of the dy slab
  unsigned rcp_zt;           // ceil(2^32 / ZT): f / ZT == umulhi(f, rcp_zt) for f < 2^16
  int merged;
};

__global__ void __launch_bounds__(FW_THREADS, 1)
conv_first_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                            float* __restrict__ partial, const FirstWgParams p) {
  extern __shared__ unsigned char smem_raw[];
  // aligned by pointer arithmetic on the __shared__ symbol (an integer round trip would turn every access into a generic load)
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) unsigned long long bars[16];
  __shared__ float red[FW_NW * 72];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t full = smem_u32(bars), empty = full + 8 * p.NSTG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.NSTG; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, FW_NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
  }
  __syncthreads();

  if (warp == FW_NW) {
    // ---------------------------------------------------------------- producer: two box loads per slab
    Ring r;
    SlabPos sp, step;
    sp.set(blockIdx.x, p.nzt, p.nyt, p.X);
    step.set(gridDim.x, p.nzt, p.nyt, p.X);
    for (int slab = blockIdx.x; slab < p.nslabs; slab += gridDim.x, sp.advance(step, p.nzt, p.nyt, p.X)) {
      mbar_wait(empty + 8 * r.s, r.ph ^ 1);
      if (elect_one()) {
        const uint32_t dst = sbase + r.s * p.stage_bytes;
        mbar_expect_tx(full + 8 * r.s, p.x_box_bytes + p.dy_bytes);
        tma_load_4d(dst, &tmap_x, full + 8 * r.s, sp.zt * p.ZT - FIRST_ZLEAD, sp.yt * p.YR - 1, sp.x - (p.kx >> 1), sp.n);
        tma_load_cb8(dst + p.x_bytes, &tmap_dy, full + 8 * r.s, p.merged, sp.zt * p.ZT, sp.yt * p.YR, sp.x, sp.n * p.Cob);
      }
      __syncwarp();
      r.advance(p.NSTG);
    }
  } else {
    // ---------------------------------------------------------------- consumers
    const int combo = warp / p.S, split = warp - combo * p.S;
    const int cob = combo / p.kx, tx = combo - cob * p.kx;
    // fp32 multiply-adds as packed pairs (fma.rn.f32x2): the three-register scalar FFMA issues every other cycle per
    // scheduler on sm_100, the packed form carries two per issue -- same IEEE results, twice the rate
    float2 acc[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    const int xoff = tx * (p.YR + 2) * p.ZP;                  // this warp's input plane inside the window (floats)
    const int doff = cob * p.YR * p.ZT;                       // this warp's octet inside the dy slab (16-byte units)
    const int ZP = p.ZP;
    Ring r;
    for (int slab = blockIdx.x; slab < p.nslabs; slab += gridDim.x) {
      mbar_wait(full + 8 * r.s, r.ph);
      const unsigned char* st = smem + (size_t)r.s * p.stage_bytes;
      const float* xs = reinterpret_cast<const float*>(st) + xoff;
      const uint4* ds = reinterpret_cast<const uint4*>(st + p.x_bytes) + doff;
      for (int k = split; k < p.nsteps; k += p.S) {
        const int f = k * 32 + lane;
        if (f < p.nvox) {
          const int yy = (int)__umulhi((unsigned)f, p.rcp_zt);
          const int zz = f - yy * p.ZT;
          const uint4 dv = ds[f];
          const float* xp = xs + yy * ZP + zz + (FIRST_ZLEAD - 1);
          float w[9];
#pragma unroll
          for (int ty = 0; ty < 3; ++ty)
#pragma unroll
            for (int tz = 0; tz < 3; ++tz) w[ty * 3 + tz] = xp[ty * ZP + tz];
          float d[8];
          unpack8(dv, d);
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const float2 ww = make_float2(w[i], w[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(ww, make_float2(d[2 * j], d[2 * j + 1]), acc[i][j]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * r.s);
      r.advance(p.NSTG);
    }
    // lanes -> warp total (fixed butterfly), one row of `red` per warp
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = warp_sum((j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x);
        if (lane == 0) red[warp * 72 + i * 8 + j] = v;
      }
  }
  __syncthreads();
  // splits of a combination in fixed order -> partial[cta][cob][tx][72]
  const int ncombo = p.Cob * p.kx;
  for (int i = threadIdx.x; i < ncombo * 72; i += FW_THREADS) {
    const int combo = i / 72, e = i - combo * 72;
    float s = 0.f;
    for (int k = 0; k < p.S; ++k) s += red[(combo * p.S + k) * 72 + e];
    partial[(size_t)blockIdx.x * ncombo * 72 + i] = s;
  }
}

// plan + tensor maps + launch; returns the number of partial chunks written (= CTAs), 0 when the shape is not eligible
// (the caller then uses the register-window kernel), < 0 on error
static bool first_wg_plan(FirstWgParams& p) {
  const int ncombo = p.Cob * p.kx;
  if (ncombo < 1 || FW_NW % ncombo) return false;
  if (p.Z % 4) return false;                                   // 16-byte global strides of the fp32 input map
  p.S = FW_NW / ncombo;
  p.nzt = (p.Z + FIRST_ZLEAD + 1 <= 256) ? 1 : (p.Z + 127) / 128;
  p.ZT = ((p.Z + p.nzt - 1) / p.nzt + 3) & ~3;
  p.ZP = (p.ZT + FIRST_ZLEAD + 1 + 3) & ~3;
  if (p.ZP > 256) return false;
  const int cap = 1152 * 2 / p.Cob;                            // dy slab <= 36 KB
  int yr = cap / p.ZT;
  if (yr < 1) return false;
  if (yr > p.Y) yr = p.Y;
  if (yr > 254) yr = 254;
  for (int c = yr; c >= (yr + 1) / 2 && c >= 1; --c)
    if (p.Y % c == 0) { yr = c; break; }
  p.YR = yr;
  p.nyt = (p.Y + yr - 1) / yr;
  p.nvox = p.YR * p.ZT;
  if (p.nvox >= 65536) return false;
  p.nsteps = (p.nvox + 31) / 32;
  p.rcp_zt = (unsigned)((0x100000000ull + (unsigned)p.ZT - 1) / (unsigned)p.ZT);
  p.x_box_bytes = (unsigned)(p.kx * (p.YR + 2) * p.ZP * 4);
  p.x_bytes = (p.x_box_bytes + 127u) & ~127u;                      // keeps the dy slab 128-byte aligned (the tail is never read)
  p.dy_bytes = (unsigned)(p.Cob * p.YR * p.ZT * 16);
  p.stage_bytes = (p.x_bytes + p.dy_bytes + 127u) & ~127u;
  p.NSTG = (int)(FW_SMEM_BUDGET / p.stage_bytes);
  if (p.NSTG > 6) p.NSTG = 6;
  if (p.NSTG < 2) return false;
  p.nslabs = p.N * p.X * p.nyt * p.nzt;
  return true;
}

int first_wgrad_tma_chunks(int n, int cout, const int* dims, const int* kernel) {
  if (tma_encoder() == nullptr || cout % 8) return 0;
  FirstWgParams p{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = dims[2]; p.kx = kernel[0]; p.Cob = cout / 8;
  if (!first_wg_plan(p)) return 0;
  const int nsm = sm_count();
  return p.nslabs < nsm ? p.nslabs : nsm;
}

int first_wgrad_tma_launch(const float* in, const void* outgrad, float* partial, int n, int cout, const int* dims,
                           const int* kernel, cudaStream_t stream) {
  EncodeTiledFn enc = tma_encoder();
  if (enc == nullptr || cout % 8 || ((uintptr_t)in & 15) || ((uintptr_t)outgrad & 15)) return 0;
  FirstWgParams p{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = dims[2]; p.kx = kernel[0]; p.Cob = cout / 8;
  if (!first_wg_plan(p)) return 0;
  CUtensorMap tmap_x, tmap_dy;
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)p.Z, (cuuint64_t)p.Y, (cuuint64_t)p.X, (cuuint64_t)p.N};
    const cuuint64_t gstr[3] = {(cuuint64_t)p.Z * 4, (cuuint64_t)p.Z * p.Y * 4, (cuuint64_t)p.Z * p.Y * p.X * 4};
    const cuuint32_t box[4] = {(cuuint32_t)p.ZP, (cuuint32_t)(p.YR + 2), (cuuint32_t)p.kx, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = enc(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_last_error("conv_first_wgrad: input tensor map failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  }
  {
    const CUresult cr = encode_cb8(enc, &tmap_dy, outgrad, p.Z, p.Y, p.X, (long long)p.N * p.Cob, p.ZT, p.YR, 1, p.Cob, &p.merged);
    if (cr != CUDA_SUCCESS) { set_last_error("conv_first_wgrad: gradient tensor map failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  }
  const int nsm = sm_count();
  const int grid = p.nslabs < nsm ? p.nslabs : nsm;
  const size_t smem = (size_t)p.NSTG * p.stage_bytes + 128;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_first_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FW_SMEM_BUDGET + 1024);
  });
  if (attr_err != cudaSuccess) { set_last_error("conv_first_wgrad: cannot opt into %u bytes of shared memory", FW_SMEM_BUDGET); return BCP_ERR_CUDA; }
  conv_first_wgrad_tma_kernel<<<grid, FW_THREADS, smem, stream>>>(tmap_x, tmap_dy, partial, p);
  return grid;
}

// ---------------------------------------------------------------------------------------------------------
// Forward: out[n][co][x][y][z] = b[co] + sum_t w[co][t] * in[n][x+tx-px][y+ty-1][z+tz-1], CB8 bf16 output.
// Same staging: per slab (n, x, YR rows, ZT columns) ONE TMA box load brings the kx input planes with their halo (padding =
// out-of-bounds zero fill).  A consumer thread owns FF_ZR consecutive z voxels of one row for ALL 16 output channels
// (64 fp32 accumulators): per (tx,ty) window row three shared-memory loads, per tap four broadcast 16-byte weight
// loads and 64 FMAs -- 93 % of the issued instructions are FMAs (the register-window kernel: 54 global loads per item
// ahead of the arithmetic, 4 warps per scheduler to hide them, 38 % of the fp32 rate).
// ---------------------------------------------------------------------------------------------------------
constexpr int FF_NW = 11;                         // consumer warps (+ the producer = 12 warps: 168 registers per thread, no spills)
constexpr int FF_THREADS = 32 * (FF_NW + 1);
constexpr int FF_ZR = 4;

struct FirstFwdParams {
  int N, X, Y, Z, kx;
  int YR, ZT, ZP, zg;        // slab rows / columns, window row length (floats), z groups per row
  int nyt, nzt, nslabs, nitems;
  int NSTG;
  unsigned box_bytes, stage_bytes;   // exact TMA box bytes / 128-byte-padded stage stride
  unsigned rcp_zg;
};

__global__ void __launch_bounds__(FF_THREADS, 1)
conv_first_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ w, const float* __restrict__ bias,
                          uint4* __restrict__ out, const FirstFwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  // aligned by pointer arithmetic on the __shared__ symbol (an integer round trip would turn every access into a generic load)
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) unsigned long long bars[16];
  __shared__ __align__(16) float wsm[27 * 16 + 16];          // [T][16 co], then the bias
  const uint32_t sbase = smem_u32(smem);
  const uint32_t full = smem_u32(bars), empty = full + 8 * p.NSTG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.kx * 9;
  for (int i = threadIdx.x; i < T * 16; i += FF_THREADS) {
    const int t = i >> 4, c = i & 15;
    wsm[i] = w[c * T + t];                                   // PyTorch layout [co][1][T]
  }
  if (threadIdx.x < 16) wsm[27 * 16 + threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.NSTG; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, FF_NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  __syncthreads();

  if (warp == FF_NW) {
    Ring r;
    SlabPos sp, step;
    sp.set(blockIdx.x, p.nzt, p.nyt, p.X);
    step.set(gridDim.x, p.nzt, p.nyt, p.X);
    for (int slab = blockIdx.x; slab < p.nslabs; slab += gridDim.x, sp.advance(step, p.nzt, p.nyt, p.X)) {
      mbar_wait(empty + 8 * r.s, r.ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full + 8 * r.s, p.box_bytes);
        tma_load_4d(sbase + r.s * p.stage_bytes, &tmap_x, full + 8 * r.s, sp.zt * p.ZT - FIRST_ZLEAD, sp.yt * p.YR - 1, sp.x - (p.kx >> 1), sp.n);
      }
      __syncwarp();
      r.advance(p.NSTG);
    }
  } else {
    const int item = threadIdx.x;                              // (row, z group) of the slab, fixed for the whole kernel
    const int yy = (int)__umulhi((unsigned)item, p.rcp_zg);
    const int zz = (item - yy * p.zg) * FF_ZR;
    const bool active = item < p.nitems;
    const long long S = (long long)p.X * p.Y * p.Z;
    const int ZP = p.ZP, rows = p.YR + 2;
    Ring r;
    SlabPos sp, step;
    sp.set(blockIdx.x, p.nzt, p.nyt, p.X);
    step.set(gridDim.x, p.nzt, p.nyt, p.X);
    for (int slab = blockIdx.x; slab < p.nslabs; slab += gridDim.x, sp.advance(step, p.nzt, p.nyt, p.X)) {
      const int n = sp.n, x = sp.x;
      mbar_wait(full + 8 * r.s, r.ph);
      const int y = sp.yt * p.YR + yy, z = sp.zt * p.ZT + zz;
      if (active && y < p.Y && z < p.Z) {
        const float* xs = reinterpret_cast<const float*>(smem + (size_t)r.s * p.stage_bytes) + yy * ZP + zz;
        float2 acc[FF_ZR][8];                                    // packed pairs: see the weight-gradient kernel
#pragma unroll
        for (int k = 0; k < FF_ZR; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k][j] = make_float2(wsm[27 * 16 + 2 * j], wsm[27 * 16 + 2 * j + 1]);
        for (int tx = 0; tx < p.kx; ++tx) {
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {
            // window elements z-1 .. z+4 sit at row offsets 3 .. 8 (the box starts FIRST_ZLEAD = 4 floats early)
            const float a0 = xs[(tx * rows + ty) * ZP + 3];
            const float4 a1 = *reinterpret_cast<const float4*>(xs + (tx * rows + ty) * ZP + 4);
            const float a2 = xs[(tx * rows + ty) * ZP + 8];
            const float2 v[6] = {make_float2(a0, a0), make_float2(a1.x, a1.x), make_float2(a1.y, a1.y),
                                 make_float2(a1.z, a1.z), make_float2(a1.w, a1.w), make_float2(a2, a2)};
            const float* wrow = wsm + (tx * 9 + ty * 3) * 16;
#pragma unroll
            for (int tz = 0; tz < 3; ++tz) {
              float2 wv[8];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 t4 = *reinterpret_cast<const float4*>(wrow + tz * 16 + j4 * 4);
                wv[j4 * 2] = make_float2(t4.x, t4.y);
                wv[j4 * 2 + 1] = make_float2(t4.z, t4.w);
              }
#pragma unroll
              for (int k = 0; k < FF_ZR; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[k][j] = __ffma2_rn(v[k + tz], wv[j], acc[k][j]);
            }
          }
        }
        uint4* dst = out + (long long)n * 2 * S + ((long long)x * p.Y + y) * p.Z + z;
#pragma unroll
        for (int k = 0; k < FF_ZR; ++k)
          if (z + k < p.Z) {
            const float* a = reinterpret_cast<const float*>(acc[k]);
            dst[k] = pack8(a);
            dst[S + k] = pack8(a + 8);
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * r.s);
      r.advance(p.NSTG);
    }
  }
}

static bool first_fwd_plan(FirstFwdParams& p) {
  if (p.Z % 4) return false;
  p.nzt = (p.Z + FIRST_ZLEAD + 1 <= 256) ? 1 : (p.Z + 127) / 128;
  p.ZT = ((p.Z + p.nzt - 1) / p.nzt + 3) & ~3;
  p.ZP = (p.ZT + FIRST_ZLEAD + 1 + 3) & ~3;
  if (p.ZP > 256) return false;
  p.zg = p.ZT / FF_ZR;
  int yr = (FF_NW * 32) / p.zg;                                 // one item per consumer thread
  if (yr < 1) return false;
  if (yr > p.Y) yr = p.Y;
  if (yr > 254) yr = 254;
  p.YR = yr;
  p.nyt = (p.Y + yr - 1) / yr;
  p.nitems = p.YR * p.zg;
  p.rcp_zg = (unsigned)((0x100000000ull + (unsigned)p.zg - 1) / (unsigned)p.zg);
  p.box_bytes = (unsigned)(p.kx * (p.YR + 2) * p.ZP * 4);
  p.stage_bytes = (p.box_bytes + 127u) & ~127u;
  p.NSTG = (int)(FW_SMEM_BUDGET / p.stage_bytes);
  if (p.NSTG > 6) p.NSTG = 6;
  if (p.NSTG < 2) return false;
  p.nslabs = p.N * p.X * p.nyt * p.nzt;
  return true;
}

// 1 = launched, 0 = shape not eligible (caller falls back to the register-window kernel), < 0 = error
int first_fwd_tma_launch(const float* in, const float* w, const float* bias, void* out, int n, int cout, const int* dims,
                         const int* kernel, cudaStream_t stream) {
  EncodeTiledFn enc = tma_encoder();
  if (enc == nullptr || cout != 16 || ((uintptr_t)in & 15)) return 0;
  FirstFwdParams p{};
  p.N = n; p.X = dims[0]; p.Y = dims[1]; p.Z = dims[2]; p.kx = kernel[0];
  if (!first_fwd_plan(p)) return 0;
  CUtensorMap tmap_x;
  const cuuint64_t gdim[4] = {(cuuint64_t)p.Z, (cuuint64_t)p.Y, (cuuint64_t)p.X, (cuuint64_t)p.N};
  const cuuint64_t gstr[3] = {(cuuint64_t)p.Z * 4, (cuuint64_t)p.Z * p.Y * 4, (cuuint64_t)p.Z * p.Y * p.X * 4};
  const cuuint32_t box[4] = {(cuuint32_t)p.ZP, (cuuint32_t)(p.YR + 2), (cuuint32_t)p.kx, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_last_error("conv_first_fwd: input tensor map failed (%d)", (int)cr); return BCP_ERR_CUDA; }
  const int nsm = sm_count();
  const int grid = p.nslabs < nsm ? p.nslabs : nsm;
  const size_t smem = (size_t)p.NSTG * p.stage_bytes + 128;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_first_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FW_SMEM_BUDGET + 1024);
  });
  if (attr_err != cudaSuccess) { set_last_error("conv_first_fwd: cannot opt into %u bytes of shared memory", FW_SMEM_BUDGET); return BCP_ERR_CUDA; }
  conv_first_fwd_tma_kernel<<<grid, FF_THREADS, smem, stream>>>(tmap_x, w, bias, (uint4*)out, p);
  return 1;
}

}  // namespace bcp
