// CUDA-core convolution kernels on channel-blocked bf16 activations (CB8: [N][C/8][X][Y][Z][8]).
// They cover the HBM-bound / odd-shaped members of the family and are the generic path for shapes the
// tcgen05 implicit-GEMM kernel (conv_tc.cu) does not take:
//   * first layer  (Cin = 1, fp32 planar input)            networks/VNet.py:151, networks/unet.py:72
//   * generic kx*ky*kz conv, stride 1|2, zero padding      nn.Conv3d / nn.Conv2d call sites
//   * 2x2x2 stride-2 transposed conv ("scatter")           nn.ConvTranspose3d, networks/VNet.py:101
//   * weight gradients for all of the above (fixed-order two-stage reduction: deterministic)
//   * classifier head to planar fp32 logits                networks/VNet.py:210, networks/unet.py:102
// fp32 accumulation everywhere; bf16 only as the storage format of activations/packed weights.
#include "common.cuh"
#include "../../include/bcp_b200.h"

namespace bcp {

struct Geom {
  int N;
  int Xi, Yi, Zi;   // input spatial dims
  int Xo, Yo, Zo;   // output spatial dims
  int kx, ky, kz;   // kernel
  int sx, sy, sz;   // stride
  int px, py, pz;   // padding (low side)
  int transposed;   // 1: out[o] = sum_ci in[o/s] * W[tap = o % s]   (k == s, no padding)
};

__device__ __forceinline__ void decompose(long long o, const Geom& g, int& n, int& x, int& y, int& z) {
  z = (int)(o % g.Zo); o /= g.Zo;
  y = (int)(o % g.Yo); o /= g.Yo;
  x = (int)(o % g.Xo);
  n = (int)(o / g.Xo);
}

// -------------------------------------------------------------------------------------------------
// generic forward (also dgrad with the flipped pack, and the stride-2 gather/scatter members)
// block = 128 threads: lane -> output voxel (32 consecutive linear outputs), warp -> 8 output channels
// weights for the current tap are staged in shared memory as fp32 [Cin/8][32 co][8 ci]
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_direct_fwd_kernel(const uint4* __restrict__ in, const uint4* __restrict__ wpack,
                                                               const float* __restrict__ bias, uint4* __restrict__ out,
                                                               Geom g, int Cin, int Cout) {
  extern __shared__ float wsm[];   // [Cib][32][8]
  const int Cib = (Cin + 7) / 8, Cob = (Cout + 7) / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co0 = blockIdx.y * 32;
  const int cob = blockIdx.y * 4 + warp;
  const long long So = (long long)g.Xo * g.Yo * g.Zo, Si = (long long)g.Xi * g.Yi * g.Zi;
  const long long total = (long long)g.N * So;
  const long long o = (long long)blockIdx.x * 32 + lane;
  const bool vox_ok = o < total;
  const bool co_ok = cob < Cob;
  int n = 0, x = 0, y = 0, z = 0;
  if (vox_ok) decompose(o, g, n, x, y, z);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int T = g.kx * g.ky * g.kz;
  int my_tap = 0;
  long long in_sp = 0;
  if (g.transposed) {
    my_tap = ((x % g.sx) * g.ky + (y % g.sy)) * g.kz + (z % g.sz);
    in_sp = ((long long)(x / g.sx) * g.Yi + (y / g.sy)) * g.Zi + (z / g.sz);
  }
  for (int t = 0; t < T; ++t) {
    __syncthreads();
    // stage W[t][cib][co0..co0+31][8] as fp32
    const int rows = Cib * 32;
    for (int r = threadIdx.x; r < rows; r += 128) {
      const int cib = r / 32, j = r % 32;
      float f[8];
      if (co0 + j < Cout) {
        unpack8(wpack[((long long)t * Cib + cib) * Cout + co0 + j], f);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(wsm + (long long)r * 8);
      dst[0] = make_float4(f[0], f[1], f[2], f[3]);
      dst[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    bool ok = vox_ok && co_ok;
    long long sp = 0;
    if (g.transposed) {
      ok = ok && (t == my_tap);
      sp = in_sp;
    } else {
      const int tz = t % g.kz, ty = (t / g.kz) % g.ky, tx = t / (g.kz * g.ky);
      const int ix = x * g.sx + tx - g.px, iy = y * g.sy + ty - g.py, iz = z * g.sz + tz - g.pz;
      ok = ok && ix >= 0 && ix < g.Xi && iy >= 0 && iy < g.Yi && iz >= 0 && iz < g.Zi;
      sp = ((long long)ix * g.Yi + iy) * g.Zi + iz;
    }
    if (ok) {
      const uint4* ip = in + (long long)n * Cib * Si + sp;
      for (int cib = 0; cib < Cib; ++cib) {
        float a[8];
        unpack8(__ldg(ip + (long long)cib * Si), a);
        const float4* w = reinterpret_cast<const float4*>(wsm + ((long long)cib * 32 + warp * 8) * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w0 = w[2 * j], w1 = w[2 * j + 1];
          acc[j] += a[0] * w0.x + a[1] * w0.y + a[2] * w0.z + a[3] * w0.w + a[4] * w1.x + a[5] * w1.y + a[6] * w1.z + a[7] * w1.w;
        }
      }
    }
  }
  if (vox_ok && co_ok) {
    if (bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (cob * 8 + j < Cout) ? bias[cob * 8 + j] : 0.f;
    }
    const long long so = ((long long)x * g.Yo + y) * g.Zo + z;
    out[((long long)n * Cob + cob) * So + so] = pack8(acc);
  }
}

// -------------------------------------------------------------------------------------------------
// generic weight gradient:  dW[a][b][t] = sum_{n,o} outgrad[o][a] * in[o*s + t - p][b]
// grid (chunks, Cib*ceil(Cob/4), T); block 128: lane -> voxel, warp -> 8 'a' channels; acc[8 b][8 a]
// partial[chunk][t][cib][cob][8 b][8 a]   then fixed-order sum over chunks
// -------------------------------------------------------------------------------------------------
constexpr int WG_VOX_PER_BLOCK = 2048;

__global__ void __launch_bounds__(128) conv_wgrad_partial_kernel(const uint4* __restrict__ in, const uint4* __restrict__ og,
                                                                  float* __restrict__ partial, Geom g, int Cin, int Cout) {
  const int Cib = (Cin + 7) / 8, Cob = (Cout + 7) / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cobt = (Cob + 3) / 4;
  const int cib = blockIdx.y / cobt;
  const int cob = (blockIdx.y % cobt) * 4 + warp;
  const int t = blockIdx.z, T = gridDim.z;
  const long long So = (long long)g.Xo * g.Yo * g.Zo, Si = (long long)g.Xi * g.Yi * g.Zi;
  const long long total = (long long)g.N * So;
  const long long o0 = (long long)blockIdx.x * WG_VOX_PER_BLOCK;
  const long long o1 = min(total, o0 + WG_VOX_PER_BLOCK);
  const int tz = t % g.kz, ty = (t / g.kz) % g.ky, tx = t / (g.kz * g.ky);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (cob < Cob) {
    for (long long o = o0 + lane; o < o1; o += 32) {
      int n, x, y, z;
      decompose(o, g, n, x, y, z);
      const int ix = x * g.sx + tx - g.px, iy = y * g.sy + ty - g.py, iz = z * g.sz + tz - g.pz;
      if (ix < 0 || ix >= g.Xi || iy < 0 || iy >= g.Yi || iz < 0 || iz >= g.Zi) continue;
      const long long so = ((long long)x * g.Yo + y) * g.Zo + z;
      float a[8], d[8];
      unpack8(__ldg(in + ((long long)n * Cib + cib) * Si + ((long long)ix * g.Yi + iy) * g.Zi + iz), a);
      unpack8(__ldg(og + ((long long)n * Cob + cob) * So + so), d);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += a[i] * d[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = warp_sum(acc[i][j]);
  if (lane == 0 && cob < Cob) {
    float* dst = partial + ((((long long)blockIdx.x * T + t) * Cib + cib) * Cob + cob) * 64;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[i * 8 + j] = acc[i][j];
  }
}

// dst layout selectable: [A][B][T] (PyTorch conv weight layout; A = channels of `og`, B = channels of `in`)
__global__ void conv_wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int chunks, int T,
                                           int Cin, int Cout, int accumulate) {
  const int Cib = (Cin + 7) / 8, Cob = (Cout + 7) / 8;
  const long long per = (long long)T * Cib * Cob * 64;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per) return;
  long long r = i;
  const int a8 = (int)(r % 8); r /= 8;
  const int b8 = (int)(r % 8); r /= 8;
  const int cob = (int)(r % Cob); r /= Cob;
  const int cib = (int)(r % Cib); r /= Cib;
  const int t = (int)r;
  const int a = cob * 8 + a8, b = cib * 8 + b8;
  if (a >= Cout || b >= Cin) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partial[(long long)c * per + i];
  float* dst = dw + ((long long)a * Cin + b) * T + t;
  *dst = accumulate ? *dst + s : s;
}

// per-channel sum of a CB8 tensor (conv bias gradient): partial then fixed-order sum
__global__ void __launch_bounds__(256) chan_sum_partial_kernel(const uint4* __restrict__ x, float* __restrict__ partial,
                                                               long long S, int chunks) {
  const int chunk = blockIdx.x, cb = blockIdx.y, n = blockIdx.z, Cb = gridDim.y;
  const uint4* base = x + ((long long)n * Cb + cb) * S;
  const long long per = (S + chunks - 1) / chunks;
  const long long s0 = (long long)chunk * per, s1 = min(S, s0 + per);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (long long s = s0 + threadIdx.x; s < s1; s += 256) {
    float f[8];
    unpack8(ldg_nc_u4(base + s), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  __shared__ float red[8 * 8];
  block_sum<8, 256>(acc, red);
  if (threadIdx.x == 0) {
    float* dst = partial + (((long long)n * Cb + cb) * chunks + chunk) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = acc[k];
  }
}

__global__ void chan_sum_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int N, int C, int chunks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int Cb = (C + 7) / 8, cb = c >> 3, k = c & 7;
  double s = 0.0;
  for (int n = 0; n < N; ++n)
    for (int ch = 0; ch < chunks; ++ch) s += (double)partial[(((long long)n * Cb + cb) * chunks + ch) * 8 + k];
  out[c] = (float)s;
}

// -------------------------------------------------------------------------------------------------
// first layer: Cin = 1, fp32 planar input [N][X][Y][Z]; weights fp32 [Cout][1][T] (PyTorch layout)
// -------------------------------------------------------------------------------------------------
constexpr int FIRST_ZR = 4;       // consecutive z outputs per thread in the first-layer forward
template <int KX>
__global__ void __launch_bounds__(128) conv_first_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                              const float* __restrict__ bias, uint4* __restrict__ out,
                                                              Geom g, int Cout) {
  // kernel KX x 3 x 3, 'same' padding.  A thread produces FIRST_ZR consecutive z voxels: the KX*3 input rows it needs are
  // read once as (FIRST_ZR + 2)-wide windows and every weight fetched from shared memory feeds FIRST_ZR voxels.
  constexpr int T = KX * 9, ZR = FIRST_ZR;
  extern __shared__ float wsm[];   // [T][Cob*8]
  const int Cob = (Cout + 7) / 8;
  for (int i = threadIdx.x; i < T * Cob * 8; i += 128) {
    const int t = i / (Cob * 8), c = i % (Cob * 8);
    wsm[i] = (c < Cout) ? w[(long long)c * T + t] : 0.f;
  }
  __syncthreads();
  const int zg = (g.Zo + ZR - 1) / ZR;
  const long long So = (long long)g.Xo * g.Yo * g.Zo;
  const long long total = (long long)g.N * g.Xo * g.Yo * zg;
  const long long stride = (long long)gridDim.x * 128;
  for (long long o = (long long)blockIdx.x * 128 + threadIdx.x; o < total; o += stride) {
    const int gz = (int)(o % zg);
    long long r = o / zg;
    const int y = (int)(r % g.Yo); r /= g.Yo;
    const int x = (int)(r % g.Xo);
    const int n = (int)(r / g.Xo);
    const int z0 = gz * ZR;
    float v[KX * 3][ZR + 2];
    const float* base = in + (long long)n * g.Xi * g.Yi * g.Zi;
#pragma unroll
    for (int tx = 0; tx < KX; ++tx)
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        const int ix = x + tx - (KX >> 1), iy = y + ty - 1;
        const bool rok = ix >= 0 && ix < g.Xi && iy >= 0 && iy < g.Yi;
        const float* row = base + ((long long)ix * g.Yi + iy) * g.Zi;
#pragma unroll
        for (int k = 0; k < ZR + 2; ++k) {
          const int iz = z0 + k - 1;
          v[tx * 3 + ty][k] = (rok && iz >= 0 && iz < g.Zi) ? __ldg(row + iz) : 0.f;
        }
      }
    const long long so = ((long long)x * g.Yo + y) * g.Zo + z0;
    for (int cob = 0; cob < Cob; ++cob) {
      float acc[ZR][8];
#pragma unroll
      for (int k = 0; k < ZR; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = (bias && cob * 8 + j < Cout) ? bias[cob * 8 + j] : 0.f;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + t * Cob * 8 + cob * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + t * Cob * 8 + cob * 8 + 4);
#pragma unroll
        for (int k = 0; k < ZR; ++k) {
          const float a = v[t / 3][k + (t % 3)];
          acc[k][0] += a * w0.x; acc[k][1] += a * w0.y; acc[k][2] += a * w0.z; acc[k][3] += a * w0.w;
          acc[k][4] += a * w1.x; acc[k][5] += a * w1.y; acc[k][6] += a * w1.z; acc[k][7] += a * w1.w;
        }
      }
      uint4* dst = out + ((long long)n * Cob + cob) * So + so;
#pragma unroll
      for (int k = 0; k < ZR; ++k)
        if (z0 + k < g.Zo) dst[k] = pack8(acc[k]);
    }
  }
}

constexpr int FIRST_WG_XC = 8;    // x planes per block in the first-layer weight gradient
// dW[co][t] for Cin = 1.  grid (ztiles * xchunks, N, Cob * kx), 4 warps.  A warp owns 32 consecutive z and a quarter of
// the y range and walks along y: per step one coalesced 16-byte dy load, three input loads (the new y row of a 3x3
// register window) and 72 FMAs -- no per-voxel index arithmetic.  partial[chunk][cob][tx][9][8], chunk = (block.x, n).
__global__ void __launch_bounds__(128) conv_first_wgrad_partial_kernel(const float* __restrict__ in, const uint4* __restrict__ og,
                                                                        float* __restrict__ partial, Geom g, int Cout) {
  const int Cob = (Cout + 7) / 8;
  const int cob = blockIdx.z / g.kx, tx = blockIdx.z - cob * g.kx;
  const int n = blockIdx.y;
  const int nzt = (g.Zo + 31) / 32;
  const int xchunk = blockIdx.x / nzt, ztile = blockIdx.x - xchunk * nzt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = ztile * 32 + lane;
  const bool zok = z < g.Zo;
  const int seg = (g.Yo + 3) / 4;
  const int y0 = warp * seg, y1 = min(g.Yo, y0 + seg);
  const long long So = (long long)g.Xo * g.Yo * g.Zo;
  float acc[9][8];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int x_end = min(g.Xo, (xchunk + 1) * FIRST_WG_XC);
  for (int x = xchunk * FIRST_WG_XC; x < x_end; ++x) {
    const int ix = x + tx - g.px;
    if (ix < 0 || ix >= g.Xi || y0 >= y1) continue;
    const float* plane = in + ((long long)n * g.Xi + ix) * g.Yi * g.Zi;
    auto load_row = [&](int iy, float* r3) {
      const bool rok = iy >= 0 && iy < g.Yi;
      const float* row = plane + (long long)iy * g.Zi;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int iz = z + k - 1;
        r3[k] = (rok && iz >= 0 && iz < g.Zi) ? __ldg(row + iz) : 0.f;
      }
    };
    float win[3][3];
    load_row(y0 - 1, win[0]);
    load_row(y0, win[1]);
    load_row(y0 + 1, win[2]);
    const uint4* dyp = og + ((long long)n * Cob + cob) * So + ((long long)x * g.Yo + y0) * g.Zo + z;
    uint4 dcur = zok ? __ldg(dyp) : make_uint4(0, 0, 0, 0);
    for (int y = y0; y < y1; ++y, dyp += g.Zo) {
      // software prefetch: the NEXT step's dy vector and input row are in flight during this step's 72 FMAs (the register
      // budget allows only ~4 warps per scheduler, too few to hide an L2 round trip per step otherwise)
      float nxt[3];
      load_row(y + 2, nxt);
      const uint4 dnext = (zok && y + 1 < y1) ? __ldg(dyp + g.Zo) : make_uint4(0, 0, 0, 0);
      float d[8];
      unpack8(dcur, d);
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float v = win[i / 3][i % 3];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += v * d[j];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) { win[0][k] = win[1][k]; win[1][k] = win[2][k]; win[2][k] = nxt[k]; }
      dcur = dnext;
    }
  }
  __shared__ float red[72 * 4];
  float flat[72];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) flat[i * 8 + j] = acc[i][j];
  block_sum<72, 128>(flat, red);
  if (threadIdx.x == 0) {
    const long long chunk = (long long)blockIdx.y * gridDim.x + blockIdx.x;
    float* dst = partial + ((chunk * Cob + cob) * g.kx + tx) * 72;
#pragma unroll
    for (int k = 0; k < 72; ++k) dst[k] = flat[k];
  }
}

// one WARP per output element: lanes stride the chunk list (independent loads in flight), fixed butterfly => deterministic
__global__ void conv_first_wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int chunks,
                                                 int Cout, int kx, int ky, int kz, int accumulate) {
  const int Cob = (Cout + 7) / 8, T = kx * ky * kz;
  const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= Cout * T) return;
  const int co = i / T, t = i % T;
  const int tz = t % kz, ty = (t / kz) % ky, tx = t / (kz * ky);
  const int cob = co >> 3, j = co & 7;
  float s = 0.f;
  for (int c = lane; c < chunks; c += 32) s += partial[(((long long)c * Cob + cob) * kx + tx) * 72 + (ty * 3 + tz) * 8 + j];
  s = warp_sum(s);
  if (lane == 0) dw[i] = accumulate ? dw[i] + s : s;
}

// -------------------------------------------------------------------------------------------------
// classifier head: CB8 bf16 -> planar fp32 logits (ncls <= 8), kernel kx*ky*kz (1 or 3 per dim), stride 1
// -------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(128) head_fwd_kernel(const uint4* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out, Geom g, int Cin) {
  extern __shared__ float wsm[];   // [T][Cib*8][NC]
  const int Cib = (Cin + 7) / 8, T = g.kx * g.ky * g.kz;
  for (int i = threadIdx.x; i < T * Cib * 8 * NC; i += 128) {
    const int k = i % NC, ci = (i / NC) % (Cib * 8), t = i / (NC * Cib * 8);
    wsm[i] = (ci < Cin) ? w[((long long)k * Cin + ci) * T + t] : 0.f;
  }
  __syncthreads();
  const long long S = (long long)g.Xo * g.Yo * g.Zo;
  const long long total = (long long)g.N * S;
  const long long stride = (long long)gridDim.x * 128;
  for (long long o = (long long)blockIdx.x * 128 + threadIdx.x; o < total; o += stride) {
    int n, x, y, z;
    decompose(o, g, n, x, y, z);
    float acc[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[k] = bias ? bias[k] : 0.f;
    for (int t = 0; t < T; ++t) {
      const int tz = t % g.kz, ty = (t / g.kz) % g.ky, tx = t / (g.kz * g.ky);
      const int ix = x + tx - g.px, iy = y + ty - g.py, iz = z + tz - g.pz;
      if (ix < 0 || ix >= g.Xi || iy < 0 || iy >= g.Yi || iz < 0 || iz >= g.Zi) continue;
      const long long sp = ((long long)ix * g.Yi + iy) * g.Zi + iz;
      for (int cib = 0; cib < Cib; ++cib) {
        float a[8];
        unpack8(__ldg(in + ((long long)n * Cib + cib) * S + sp), a);
        const float* wr = wsm + ((long long)t * Cib * 8 + cib * 8) * NC;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int k = 0; k < NC; ++k) acc[k] += a[i] * wr[i * NC + k];
      }
    }
    const long long sp = o - (long long)n * S;
#pragma unroll
    for (int k = 0; k < NC; ++k) out[((long long)n * NC + k) * S + sp] = acc[k];
  }
}

// ---- 1x1x1 head (the out_conv of every network here: networks/VNet.py:210, networks/unet.py:102, pancreas/Vnet.py:128): no taps, no
// borders, input index == output index.  The generic kernels above spend most of their instructions on 64-bit index
// decomposition; these stream: one thread per voxel, all channel octets in registers, coalesced 16-byte / 4-byte accesses.
// HBM-bound: fwd / dgrad / wgrad each move |activation| + |logits| bytes.
template <int NC, int CIB>
__global__ void __launch_bounds__(256) head1_fwd_kernel(const uint4* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, long long S, int Cin) {
  __shared__ float wsm[CIB * 8 * NC];
  for (int i = threadIdx.x; i < CIB * 8 * NC; i += 256) {
    const int k = i % NC, ci = i / NC;
    wsm[i] = (ci < Cin) ? w[(long long)k * Cin + ci] : 0.f;
  }
  __syncthreads();
  const int n = blockIdx.y;
  const uint4* src = in + (long long)n * CIB * S;
  float* dst = out + (long long)n * NC * S;
  const long long stride = (long long)gridDim.x * 256;
#pragma unroll 2
  for (long long sp = (long long)blockIdx.x * 256 + threadIdx.x; sp < S; sp += stride) {
    uint4 v[CIB];
#pragma unroll
    for (int cib = 0; cib < CIB; ++cib) v[cib] = ldg_nc_u4(src + (long long)cib * S + sp);
    float acc[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[k] = bias ? __ldg(bias + k) : 0.f;
#pragma unroll
    for (int cib = 0; cib < CIB; ++cib) {
      float a[8];
      unpack8(v[cib], a);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < NC; ++k) acc[k] = fmaf(a[i], wsm[(cib * 8 + i) * NC + k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) dst[(long long)k * S + sp] = acc[k];
  }
}

template <int NC, int CIB>
__global__ void __launch_bounds__(256) head1_dgrad_kernel(const float* __restrict__ dlog, const float* __restrict__ w,
                                                           uint4* __restrict__ din, long long S, int Cin) {
  __shared__ float wsm[CIB * 8 * NC];
  for (int i = threadIdx.x; i < CIB * 8 * NC; i += 256) {
    const int k = i % NC, ci = i / NC;
    wsm[i] = (ci < Cin) ? w[(long long)k * Cin + ci] : 0.f;
  }
  __syncthreads();
  const int n = blockIdx.y;
  const float* src = dlog + (long long)n * NC * S;
  uint4* dst = din + (long long)n * CIB * S;
  const long long stride = (long long)gridDim.x * 256;
#pragma unroll 2
  for (long long sp = (long long)blockIdx.x * 256 + threadIdx.x; sp < S; sp += stride) {
    float d[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) d[k] = __ldg(src + (long long)k * S + sp);
#pragma unroll
    for (int cib = 0; cib < CIB; ++cib) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) a += d[k] * wsm[(cib * 8 + i) * NC + k];     // same order as the generic kernel
        acc[i] = a;
      }
      dst[(long long)cib * S + sp] = pack8(acc);
    }
  }
}

// partial[chunk][cib][NC*8 + NC] (the generic layout with T = 1): one pass over the voxels for ALL channel octets
template <int NC, int CIB>
__global__ void __launch_bounds__(128) head1_wgrad_partial_kernel(const uint4* __restrict__ in, const float* __restrict__ dlog,
                                                                   float* __restrict__ partial, long long S, int N) {
  constexpr int K = NC * 8 + NC;
  const long long total = (long long)N * S;
  const long long o0 = (long long)blockIdx.x * 4096, o1 = min(total, o0 + 4096);
  float acc[CIB][NC * 8], accb[NC];
#pragma unroll
  for (int c = 0; c < CIB; ++c)
#pragma unroll
    for (int k = 0; k < NC * 8; ++k) acc[c][k] = 0.f;
#pragma unroll
  for (int k = 0; k < NC; ++k) accb[k] = 0.f;
  for (long long o = o0 + threadIdx.x; o < o1; o += 128) {
    const long long n = o / S, sp = o - n * S;
    float d[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { d[k] = __ldg(dlog + (n * NC + k) * S + sp); accb[k] += d[k]; }
#pragma unroll
    for (int c = 0; c < CIB; ++c) {
      float a[8];
      unpack8(ldg_nc_u4(in + (n * CIB + c) * S + sp), a);
#pragma unroll
      for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][k * 8 + i] += d[k] * a[i];
    }
  }
  __shared__ float red[K * 4];
#pragma unroll
  for (int c = 0; c < CIB; ++c) {
    float v[K];
#pragma unroll
    for (int k = 0; k < NC * 8; ++k) v[k] = acc[c][k];
#pragma unroll
    for (int k = 0; k < NC; ++k) v[NC * 8 + k] = accb[k];
    block_sum<K, 128>(v, red);
    if (threadIdx.x == 0) {
      float* dst = partial + ((long long)blockIdx.x * CIB + c) * K;
#pragma unroll
      for (int k = 0; k < K; ++k) dst[k] = v[k];
    }
  }
}

// d_in[v][ci] = sum_t sum_k dlogits[v - t + p][k] * W[k][ci][t]
template <int NC>
__global__ void __launch_bounds__(128) head_dgrad_kernel(const float* __restrict__ dlog, const float* __restrict__ w,
                                                          uint4* __restrict__ din, Geom g, int Cin) {
  extern __shared__ float wsm[];   // [T][Cib*8][NC]
  const int Cib = (Cin + 7) / 8, T = g.kx * g.ky * g.kz;
  for (int i = threadIdx.x; i < T * Cib * 8 * NC; i += 128) {
    const int k = i % NC, ci = (i / NC) % (Cib * 8), t = i / (NC * Cib * 8);
    wsm[i] = (ci < Cin) ? w[((long long)k * Cin + ci) * T + t] : 0.f;
  }
  __syncthreads();
  const long long S = (long long)g.Xo * g.Yo * g.Zo;
  const long long total = (long long)g.N * S;
  const long long stride = (long long)gridDim.x * 128;
  for (long long o = (long long)blockIdx.x * 128 + threadIdx.x; o < total; o += stride) {
    int n, x, y, z;
    decompose(o, g, n, x, y, z);
    for (int cib = 0; cib < Cib; ++cib) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int t = 0; t < T; ++t) {
        const int tz = t % g.kz, ty = (t / g.kz) % g.ky, tx = t / (g.kz * g.ky);
        const int ox = x - tx + g.px, oy = y - ty + g.py, oz = z - tz + g.pz;   // output voxel that read us through tap t
        if (ox < 0 || ox >= g.Xo || oy < 0 || oy >= g.Yo || oz < 0 || oz >= g.Zo) continue;
        const long long sp = ((long long)ox * g.Yo + oy) * g.Zo + oz;
        const float* wr = wsm + ((long long)t * Cib * 8 + cib * 8) * NC;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const float d = __ldg(dlog + ((long long)n * NC + k) * S + sp);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += d * wr[i * NC + k];
        }
      }
      din[((long long)n * Cib + cib) * S + (o - (long long)n * S)] = pack8(acc);
    }
  }
}

// dW[k][ci][t], db[k]: grid (chunks, Cib, T); partial[chunk][t][cib][NC*8 + NC]
template <int NC>
__global__ void __launch_bounds__(128) head_wgrad_partial_kernel(const uint4* __restrict__ in, const float* __restrict__ dlog,
                                                                  float* __restrict__ partial, Geom g, int Cin) {
  const int Cib = (Cin + 7) / 8;
  const int cib = blockIdx.y, t = blockIdx.z, T = gridDim.z;
  const int tz = t % g.kz, ty = (t / g.kz) % g.ky, tx = t / (g.kz * g.ky);
  const long long S = (long long)g.Xo * g.Yo * g.Zo;
  const long long total = (long long)g.N * S;
  const long long o0 = (long long)blockIdx.x * 4096, o1 = min(total, o0 + 4096);
  constexpr int K = NC * 8 + NC;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  for (long long o = o0 + threadIdx.x; o < o1; o += 128) {
    int n, x, y, z;
    decompose(o, g, n, x, y, z);
    float d[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { d[k] = __ldg(dlog + ((long long)n * NC + k) * S + (o - (long long)n * S)); acc[NC * 8 + k] += d[k]; }
    const int ix = x + tx - g.px, iy = y + ty - g.py, iz = z + tz - g.pz;
    if (ix < 0 || ix >= g.Xi || iy < 0 || iy >= g.Yi || iz < 0 || iz >= g.Zi) continue;
    float a[8];
    unpack8(__ldg(in + ((long long)n * Cib + cib) * S + ((long long)ix * g.Yi + iy) * g.Zi + iz), a);
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[k * 8 + i] += d[k] * a[i];
  }
  __shared__ float red[K * 4];
  block_sum<K, 128>(acc, red);
  if (threadIdx.x == 0) {
    float* dst = partial + (((long long)blockIdx.x * T + t) * Cib + cib) * K;
#pragma unroll
    for (int k = 0; k < K; ++k) dst[k] = acc[k];
  }
}

// one WARP per output element (see conv_first_wgrad_finalize_kernel)
template <int NC>
__global__ void head_wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                           int chunks, int T, int Cin, int accumulate) {
  const int Cib = (Cin + 7) / 8;
  constexpr int K = NC * 8 + NC;
  const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i < NC * Cin * T) {
    const int t = i % T, ci = (i / T) % Cin, k = i / (T * Cin);
    float s = 0.f;
    for (int c = lane; c < chunks; c += 32) s += partial[(((long long)c * T + t) * Cib + (ci >> 3)) * K + k * 8 + (ci & 7)];
    s = warp_sum(s);
    if (lane == 0) dw[i] = accumulate ? dw[i] + s : s;
  } else if (i < NC * Cin * T + NC && db) {
    const int k = i - NC * Cin * T;
    float s = 0.f;
    for (int c = lane; c < chunks; c += 32) s += partial[(((long long)c * T + 0) * Cib + 0) * K + NC * 8 + k];
    s = warp_sum(s);
    if (lane == 0) db[k] = accumulate ? db[k] + s : s;
  }
}

static int make_geom(Geom& g, int n, const int* in_dims, const int* kernel, const int* stride, const int* pad, int transposed) {
  g.N = n;
  g.Xi = in_dims[0]; g.Yi = in_dims[1]; g.Zi = in_dims[2];
  g.kx = kernel[0]; g.ky = kernel[1]; g.kz = kernel[2];
  g.sx = stride[0]; g.sy = stride[1]; g.sz = stride[2];
  g.px = pad[0]; g.py = pad[1]; g.pz = pad[2];
  g.transposed = transposed;
  if (transposed) {
    if (g.kx != g.sx || g.ky != g.sy || g.kz != g.sz || g.px || g.py || g.pz) return -1;
    g.Xo = g.Xi * g.sx; g.Yo = g.Yi * g.sy; g.Zo = g.Zi * g.sz;
  } else {
    g.Xo = (g.Xi + 2 * g.px - g.kx) / g.sx + 1;
    g.Yo = (g.Yi + 2 * g.py - g.ky) / g.sy + 1;
    g.Zo = (g.Zi + 2 * g.pz - g.kz) / g.sz + 1;
  }
  if (g.Xo <= 0 || g.Yo <= 0 || g.Zo <= 0) return -1;
  return 0;
}

// conv_first_tma.cu: TMA-staged first-layer kernels.  launch: > 0 = launched (weight gradient: number of per-CTA partials),
// 0 = shape not eligible -> the register-window kernels above, < 0 = error
int first_wgrad_tma_chunks(int n, int cout, const int* dims, const int* kernel);
int first_wgrad_tma_launch(const float* in, const void* outgrad, float* partial, int n, int cout, const int* dims,
                           const int* kernel, cudaStream_t stream);
int first_fwd_tma_launch(const float* in, const float* w, const float* bias, void* out, int n, int cout, const int* dims,
                         const int* kernel, cudaStream_t stream);

}  // namespace bcp

using namespace bcp;

extern "C" {

int bcp_conv_direct_fwd(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                        const int* in_dims, const int* kernel, const int* stride, const int* pad, int transposed,
                        cudaStream_t stream) {
  BCP_REQUIRE(in && wpack && out && in_dims && kernel && stride && pad, "conv_direct_fwd: null pointer");
  BCP_REQUIRE(n > 0 && cin > 0 && cout > 0, "conv_direct_fwd: bad channels");
  Geom g;
  BCP_REQUIRE(make_geom(g, n, in_dims, kernel, stride, pad, transposed) == 0, "conv_direct_fwd: bad geometry");
  const int Cib = (cin + 7) / 8, Cob = (cout + 7) / 8;
  const long long total = (long long)n * g.Xo * g.Yo * g.Zo;
  const size_t smem = (size_t)Cib * 32 * 8 * sizeof(float);
  BCP_REQUIRE(smem <= 200 * 1024, "conv_direct_fwd: Cin %d too large", cin);
  if (smem > 48 * 1024) cudaFuncSetAttribute(conv_direct_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)((total + 31) / 32), (Cob + 3) / 4);
  conv_direct_fwd_kernel<<<grid, 128, smem, stream>>>((const uint4*)in, (const uint4*)wpack, bias, (uint4*)out, g, cin, cout);
  return check_launch("conv_direct_fwd");
}

long long bcp_conv_wgrad_workspace_floats(int n, int cin, int cout, const int* out_dims, const int* kernel) {
  const long long total = (long long)n * out_dims[0] * out_dims[1] * out_dims[2];
  const long long chunks = (total + WG_VOX_PER_BLOCK - 1) / WG_VOX_PER_BLOCK;
  const long long T = (long long)kernel[0] * kernel[1] * kernel[2];
  return chunks * T * ((cin + 7) / 8) * ((cout + 7) / 8) * 64;
}

// dw[cout][cin][T] where `outgrad` has `cout` channels at output resolution and `in` has `cin` channels
int bcp_conv_direct_wgrad(const void* in, const void* outgrad, float* dw, float* workspace, int n, int cin, int cout,
                          const int* in_dims, const int* kernel, const int* stride, const int* pad, int accumulate,
                          cudaStream_t stream) {
  BCP_REQUIRE(in && outgrad && dw && workspace, "conv_direct_wgrad: null pointer");
  Geom g;
  BCP_REQUIRE(make_geom(g, n, in_dims, kernel, stride, pad, 0) == 0, "conv_direct_wgrad: bad geometry");
  const int Cib = (cin + 7) / 8, Cob = (cout + 7) / 8, T = g.kx * g.ky * g.kz;
  const long long total = (long long)n * g.Xo * g.Yo * g.Zo;
  const int chunks = (int)((total + WG_VOX_PER_BLOCK - 1) / WG_VOX_PER_BLOCK);
  dim3 grid(chunks, Cib * ((Cob + 3) / 4), T);
  conv_wgrad_partial_kernel<<<grid, 128, 0, stream>>>((const uint4*)in, (const uint4*)outgrad, workspace, g, cin, cout);
  const long long per = (long long)T * Cib * Cob * 64;
  conv_wgrad_finalize_kernel<<<(unsigned)((per + 255) / 256), 256, 0, stream>>>(workspace, dw, chunks, T, cin, cout, accumulate);
  return check_launch("conv_direct_wgrad");
}

long long bcp_chan_sum_workspace_floats(int n, int c, long long s) {
  long long chunks = (s + 16383) / 16384;
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  return (long long)n * ((c + 7) / 8) * chunks * 8;
}

int bcp_chan_sum(const void* x, float* out, float* workspace, int n, int c, long long s, cudaStream_t stream) {
  BCP_REQUIRE(x && out && workspace && n > 0 && c > 0 && s > 0, "chan_sum: bad args");
  long long chunks = (s + 16383) / 16384;
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (c + 7) / 8, n);
  chan_sum_partial_kernel<<<grid, 256, 0, stream>>>((const uint4*)x, workspace, s, (int)chunks);
  chan_sum_finalize_kernel<<<(c + 127) / 128, 128, 0, stream>>>(workspace, out, n, c, (int)chunks);
  return check_launch("chan_sum");
}

int bcp_conv_first_fwd(const float* in, const float* w, const float* bias, void* out, int n, int cout,
                       const int* dims, const int* kernel, cudaStream_t stream) {
  BCP_REQUIRE(in && w && out && dims && kernel, "conv_first_fwd: null pointer");
  const int stride[3] = {1, 1, 1};
  const int pad[3] = {kernel[0] / 2, kernel[1] / 2, kernel[2] / 2};
  Geom g;
  BCP_REQUIRE(make_geom(g, n, dims, kernel, stride, pad, 0) == 0, "conv_first_fwd: bad geometry");
  BCP_REQUIRE((g.kx == 3 || g.kx == 1) && g.ky == 3 && g.kz == 3, "conv_first_fwd: kernel must be 3x3x3 or 1x3x3");
  {
    const int rc = first_fwd_tma_launch(in, w, bias, out, n, cout, dims, kernel, stream);
    if (rc < 0) return rc;
    if (rc == 1) return check_launch("conv_first_fwd");
  }
  const int Cob = (cout + 7) / 8;
  const long long total = (long long)n * g.Xo * g.Yo * ((g.Zo + FIRST_ZR - 1) / FIRST_ZR);
  long long blocks = (total + 127) / 128;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)g.kx * g.ky * g.kz * Cob * 8 * sizeof(float);
  if (g.kx == 3)
    conv_first_fwd_kernel<3><<<(unsigned)blocks, 128, smem, stream>>>(in, w, bias, (uint4*)out, g, cout);
  else
    conv_first_fwd_kernel<1><<<(unsigned)blocks, 128, smem, stream>>>(in, w, bias, (uint4*)out, g, cout);
  return check_launch("conv_first_fwd");
}

long long bcp_conv_first_wgrad_workspace_floats(int n, int cout, const int* dims, const int* kernel) {
  long long chunks = (long long)n * ((dims[2] + 31) / 32) * ((dims[0] + FIRST_WG_XC - 1) / FIRST_WG_XC);
  const long long tma = first_wgrad_tma_chunks(n, cout, dims, kernel);
  if (tma > chunks) chunks = tma;
  return chunks * ((cout + 7) / 8) * kernel[0] * 72;
}

int bcp_conv_first_wgrad_tma_chunks(int n, int cout, const int* dims, const int* kernel) {
  if (!dims || !kernel || kernel[1] != 3 || kernel[2] != 3 || (kernel[0] != 1 && kernel[0] != 3)) return 0;
  return first_wgrad_tma_chunks(n, cout, dims, kernel);
}

int bcp_conv_first_wgrad(const float* in, const void* outgrad, float* dw, float* workspace, int n, int cout,
                         const int* dims, const int* kernel, int accumulate, cudaStream_t stream) {
  BCP_REQUIRE(in && outgrad && dw && workspace, "conv_first_wgrad: null pointer");
  const int stride[3] = {1, 1, 1};
  const int pad[3] = {kernel[0] / 2, kernel[1] / 2, kernel[2] / 2};
  Geom g;
  BCP_REQUIRE(make_geom(g, n, dims, kernel, stride, pad, 0) == 0, "conv_first_wgrad: bad geometry");
  BCP_REQUIRE(g.ky == 3 && g.kz == 3 && (g.kx == 1 || g.kx == 3), "conv_first_wgrad: kernel must be 3x3x3 or 1x3x3");
  const int Cob = (cout + 7) / 8;
  int chunks = first_wgrad_tma_launch(in, outgrad, workspace, n, cout, dims, kernel, stream);
  if (chunks < 0) return chunks;
  if (chunks == 0) {
    const int per_n = ((g.Zo + 31) / 32) * ((g.Xo + FIRST_WG_XC - 1) / FIRST_WG_XC);
    chunks = per_n * n;
    dim3 grid(per_n, n, Cob * g.kx);
    conv_first_wgrad_partial_kernel<<<grid, 128, 0, stream>>>(in, (const uint4*)outgrad, workspace, g, cout);
  }
  const int T = g.kx * g.ky * g.kz;
  conv_first_wgrad_finalize_kernel<<<(cout * T * 32 + 127) / 128, 128, 0, stream>>>(workspace, dw, chunks, cout, g.kx, g.ky, g.kz, accumulate);
  return check_launch("conv_first_wgrad");
}

#define HEAD_DISPATCH(NCV, CALL) \
  switch (NCV) {                 \
    case 2: { constexpr int NC = 2; CALL; } break; \
    case 4: { constexpr int NC = 4; CALL; } break; \
    default: set_last_error("head: %d classes unsupported (2 or 4)", NCV); return BCP_ERR_UNSUPPORTED; }

static int head_geom(Geom& g, int n, const int* dims, const int* kernel) {
  const int stride[3] = {1, 1, 1};
  const int pad[3] = {kernel[0] / 2, kernel[1] / 2, kernel[2] / 2};
  return make_geom(g, n, dims, kernel, stride, pad, 0);
}

int bcp_head_fwd(const void* in, const float* w, const float* bias, float* logits, int n, int cin, int ncls,
                 const int* dims, const int* kernel, cudaStream_t stream) {
  BCP_REQUIRE(in && w && logits && dims && kernel, "head_fwd: null pointer");
  Geom g;
  BCP_REQUIRE(head_geom(g, n, dims, kernel) == 0, "head_fwd: bad geometry");
  const long long total = (long long)n * g.Xo * g.Yo * g.Zo;
  long long blocks = (total + 127) / 128;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (g.kx * g.ky * g.kz == 1 && (cin == 16 || cin == 8) && (ncls == 2 || ncls == 4)) {
    const long long S = (long long)g.Xo * g.Yo * g.Zo;
    long long bx = (S + 511) / 512;
    const long long capx = (long long)sm_count() * 8 / n + 1;
    if (bx > capx) bx = capx;
    dim3 grid((unsigned)bx, (unsigned)n);
#define H1F(NC, CIB) head1_fwd_kernel<NC, CIB><<<grid, 256, 0, stream>>>((const uint4*)in, w, bias, logits, S, cin)
    if (ncls == 2 && cin == 16) H1F(2, 2); else if (ncls == 4 && cin == 16) H1F(4, 2); else if (ncls == 2) H1F(2, 1); else H1F(4, 1);
#undef H1F
    return check_launch("head_fwd");
  }
  const size_t smem = (size_t)g.kx * g.ky * g.kz * ((cin + 7) / 8) * 8 * ncls * sizeof(float);
  BCP_REQUIRE(smem <= 48 * 1024, "head_fwd: weights do not fit shared memory");
  HEAD_DISPATCH(ncls, (head_fwd_kernel<NC><<<(unsigned)blocks, 128, smem, stream>>>((const uint4*)in, w, bias, logits, g, cin)));
  return check_launch("head_fwd");
}

int bcp_head_dgrad(const float* dlogits, const float* w, void* din, int n, int cin, int ncls, const int* dims,
                   const int* kernel, cudaStream_t stream) {
  BCP_REQUIRE(dlogits && w && din && dims && kernel, "head_dgrad: null pointer");
  Geom g;
  BCP_REQUIRE(head_geom(g, n, dims, kernel) == 0, "head_dgrad: bad geometry");
  const long long total = (long long)n * g.Xo * g.Yo * g.Zo;
  long long blocks = (total + 127) / 128;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (g.kx * g.ky * g.kz == 1 && (cin == 16 || cin == 8) && (ncls == 2 || ncls == 4)) {
    const long long S = (long long)g.Xo * g.Yo * g.Zo;
    long long bx = (S + 511) / 512;
    const long long capx = (long long)sm_count() * 8 / n + 1;
    if (bx > capx) bx = capx;
    dim3 grid((unsigned)bx, (unsigned)n);
#define H1D(NC, CIB) head1_dgrad_kernel<NC, CIB><<<grid, 256, 0, stream>>>(dlogits, w, (uint4*)din, S, cin)
    if (ncls == 2 && cin == 16) H1D(2, 2); else if (ncls == 4 && cin == 16) H1D(4, 2); else if (ncls == 2) H1D(2, 1); else H1D(4, 1);
#undef H1D
    return check_launch("head_dgrad");
  }
  const size_t smem = (size_t)g.kx * g.ky * g.kz * ((cin + 7) / 8) * 8 * ncls * sizeof(float);
  BCP_REQUIRE(smem <= 48 * 1024, "head_dgrad: weights do not fit shared memory");
  HEAD_DISPATCH(ncls, (head_dgrad_kernel<NC><<<(unsigned)blocks, 128, smem, stream>>>(dlogits, w, (uint4*)din, g, cin)));
  return check_launch("head_dgrad");
}

long long bcp_head_wgrad_workspace_floats(int n, int cin, int ncls, const int* dims, const int* kernel) {
  const long long total = (long long)n * dims[0] * dims[1] * dims[2];
  return ((total + 4095) / 4096) * kernel[0] * kernel[1] * kernel[2] * ((cin + 7) / 8) * (ncls * 8 + ncls);
}

int bcp_head_wgrad(const void* in, const float* dlogits, float* dw, float* db, float* workspace, int n, int cin, int ncls,
                   const int* dims, const int* kernel, int accumulate, cudaStream_t stream) {
  BCP_REQUIRE(in && dlogits && dw && workspace, "head_wgrad: null pointer");
  Geom g;
  BCP_REQUIRE(head_geom(g, n, dims, kernel) == 0, "head_wgrad: bad geometry");
  const long long total = (long long)n * g.Xo * g.Yo * g.Zo;
  const int chunks = (int)((total + 4095) / 4096), T = g.kx * g.ky * g.kz, Cib = (cin + 7) / 8;
  if (T == 1 && (cin == 16 || cin == 8) && (ncls == 2 || ncls == 4)) {
    const long long S = (long long)g.Xo * g.Yo * g.Zo;
#define H1W(NC, CIB) head1_wgrad_partial_kernel<NC, CIB><<<chunks, 128, 0, stream>>>((const uint4*)in, dlogits, workspace, S, n)
    if (ncls == 2 && cin == 16) H1W(2, 2); else if (ncls == 4 && cin == 16) H1W(4, 2); else if (ncls == 2) H1W(2, 1); else H1W(4, 1);
#undef H1W
  } else {
  dim3 grid(chunks, Cib, T);
  HEAD_DISPATCH(ncls, (head_wgrad_partial_kernel<NC><<<grid, 128, 0, stream>>>((const uint4*)in, dlogits, workspace, g, cin)));
  }
  const int nout = ncls * cin * T + ncls;
  HEAD_DISPATCH(ncls, (head_wgrad_finalize_kernel<NC><<<(nout * 32 + 127) / 128, 128, 0, stream>>>(workspace, dw, db, chunks, T, cin, accumulate)));
  return check_launch("head_wgrad");
}

}  // extern "C"
