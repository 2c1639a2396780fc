// Resampling ops of the 2-D U-Net (networks/unet.py:37,50) and the V-Net's feature pool
// (networks/VNet.py:249,289) on CB8 bf16 activations.  All gather-style (deterministic, no atomics).
#include "common.cuh"
#include "../../include/bcp_b200.h"

namespace bcp {

// MaxPool (2,2) over the last two spatial dims; dims: in [N][Cb][X][Y][Z], out [N][Cb][X][Y/2][Z/2]
__global__ void maxpool2_fwd_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long planes, int Y, int Z) {
  const int Yo = Y / 2, Zo = Z / 2;
  const long long total = planes * Yo * Zo;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int zo = (int)(i % Zo);
    const long long r = i / Zo;
    const int yo = (int)(r % Yo);
    const long long p = r / Yo;
    const uint4* b = in + (p * Y + 2 * yo) * Z + 2 * zo;
    float m[8], f[8];
    unpack8(b[0], m);
    unpack8(b[1], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = f[k] > m[k] ? f[k] : m[k];
    unpack8(b[Z], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = f[k] > m[k] ? f[k] : m[k];
    unpack8(b[Z + 1], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = f[k] > m[k] ? f[k] : m[k];
    out[i] = pack8(m);
  }
}

// gradient goes to the first maximum in window scan order (like ATen's max_pool2d backward)
__global__ void maxpool2_bwd_kernel(const uint4* __restrict__ in, const uint4* __restrict__ dout, uint4* __restrict__ din,
                                    long long planes, int Y, int Z) {
  const int Yo = Y / 2, Zo = Z / 2;
  const long long total = planes * Yo * Zo;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int zo = (int)(i % Zo);
    const long long r = i / Zo;
    const int yo = (int)(r % Yo);
    const long long p = r / Yo;
    const long long base = (p * Y + 2 * yo) * Z + 2 * zo;
    float v[4][8], d[8], o[4][8];
    unpack8(in[base], v[0]);
    unpack8(in[base + 1], v[1]);
    unpack8(in[base + Z], v[2]);
    unpack8(in[base + Z + 1], v[3]);
    unpack8(dout[i], d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int best = 0;
      float m = v[0][k];
#pragma unroll
      for (int q = 1; q < 4; ++q) if (v[q][k] > m) { m = v[q][k]; best = q; }
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][k] = (q == best) ? d[k] : 0.f;
    }
    din[base] = pack8(o[0]);
    din[base + 1] = pack8(o[1]);
    din[base + Z] = pack8(o[2]);
    din[base + Z + 1] = pack8(o[3]);
  }
  // odd trailing rows/cols (not reached by any window) get zero gradient
  if ((Y & 1) || (Z & 1)) {
    const long long tot_in = planes * Y * Z;
    float zf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot_in; i += stride) {
      const int z = (int)(i % Z), y = (int)((i / Z) % Y);
      if (y >= 2 * Yo || z >= 2 * Zo) din[i] = pack8(zf);
    }
  }
}

__device__ __forceinline__ void src_index(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float src = scale * (float)o;          // align_corners=True: scale = (in-1)/(out-1)
  i0 = (int)src;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = src - (float)i0;
}

// bilinear x2, align_corners=True over the last two spatial dims
__global__ void upsample2_fwd_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long planes, int Y, int Z) {
  const int Yo = 2 * Y, Zo = 2 * Z;
  const float sy = (Yo > 1) ? (float)(Y - 1) / (float)(Yo - 1) : 0.f, sz = (Zo > 1) ? (float)(Z - 1) / (float)(Zo - 1) : 0.f;
  const long long total = planes * Yo * Zo;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int zo = (int)(i % Zo);
    const long long r = i / Zo;
    const int yo = (int)(r % Yo);
    const long long p = r / Yo;
    int y0, y1, z0, z1;
    float ly, lz;
    src_index(yo, sy, Y, y0, y1, ly);
    src_index(zo, sz, Z, z0, z1, lz);
    const uint4* b = in + p * Y * Z;
    float a[8], c[8], d[8], e[8], o[8];
    unpack8(b[(long long)y0 * Z + z0], a);
    unpack8(b[(long long)y0 * Z + z1], c);
    unpack8(b[(long long)y1 * Z + z0], d);
    unpack8(b[(long long)y1 * Z + z1], e);
    const float hy = 1.f - ly, hz = 1.f - lz;
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = hy * (hz * a[k] + lz * c[k]) + ly * (hz * d[k] + lz * e[k]);
    out[i] = pack8(o);
  }
}

__global__ void upsample2_bwd_kernel(const uint4* __restrict__ dout, uint4* __restrict__ din, long long planes, int Y, int Z) {
  const int Yo = 2 * Y, Zo = 2 * Z;
  const float sy = (Yo > 1) ? (float)(Y - 1) / (float)(Yo - 1) : 0.f, sz = (Zo > 1) ? (float)(Z - 1) / (float)(Zo - 1) : 0.f;
  const long long total = planes * Y * Z;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int z = (int)(i % Z);
    const long long r = i / Z;
    const int y = (int)(r % Y);
    const long long p = r / Y;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const uint4* b = dout + p * Yo * Zo;
    for (int yo = max(0, 2 * y - 2); yo <= min(Yo - 1, 2 * y + 4); ++yo) {
      int y0, y1; float ly;
      src_index(yo, sy, Y, y0, y1, ly);
      const float wy = ((y0 == y) ? (1.f - ly) : 0.f) + ((y1 == y) ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int zo = max(0, 2 * z - 2); zo <= min(Zo - 1, 2 * z + 4); ++zo) {
        int z0, z1; float lz;
        src_index(zo, sz, Z, z0, z1, lz);
        const float wz = ((z0 == z) ? (1.f - lz) : 0.f) + ((z1 == z) ? lz : 0.f);
        if (wz == 0.f) continue;
        float d[8];
        unpack8(b[(long long)yo * Zo + zo], d);
        const float w = wy * wz;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += w * d[k];
      }
    }
    din[i] = pack8(acc);
  }
}

// MaxPool3d(kernel 3, stride 2, no padding): CB8 in -> planar fp32 out [N][C][Xo][Yo][Zo]
__global__ void maxpool3d_k3s2_kernel(const uint4* __restrict__ in, float* __restrict__ out, int N, int C, int X, int Y, int Z,
                                      int Xo, int Yo, int Zo) {
  const int Cb = (C + 7) / 8;
  const long long total = (long long)N * Cb * Xo * Yo * Zo;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long r = i;
    const int zo = (int)(r % Zo); r /= Zo;
    const int yo = (int)(r % Yo); r /= Yo;
    const int xo = (int)(r % Xo); r /= Xo;
    const int cb = (int)(r % Cb);
    const int n = (int)(r / Cb);
    const uint4* b = in + ((long long)n * Cb + cb) * X * Y * Z;
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
    for (int dx = 0; dx < 3; ++dx)
      for (int dy = 0; dy < 3; ++dy)
        for (int dz = 0; dz < 3; ++dz) {
          float f[8];
          unpack8(b[((long long)(2 * xo + dx) * Y + (2 * yo + dy)) * Z + 2 * zo + dz], f);
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
        }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cb * 8 + k;
      if (c < C) out[((((long long)n * C + c) * Xo + xo) * Yo + yo) * Zo + zo] = m[k];
    }
  }
}

static inline int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace bcp

using namespace bcp;

extern "C" {

int bcp_maxpool2_fwd(const void* in, void* out, long long planes, int y, int z, cudaStream_t stream) {
  BCP_REQUIRE(in && out && planes > 0 && y >= 2 && z >= 2, "maxpool2_fwd: bad args");
  maxpool2_fwd_kernel<<<grid_for(planes * (y / 2) * (z / 2), 256), 256, 0, stream>>>((const uint4*)in, (uint4*)out, planes, y, z);
  return check_launch("maxpool2_fwd");
}

int bcp_maxpool2_bwd(const void* in, const void* dout, void* din, long long planes, int y, int z, cudaStream_t stream) {
  BCP_REQUIRE(in && dout && din && planes > 0 && y >= 2 && z >= 2, "maxpool2_bwd: bad args");
  maxpool2_bwd_kernel<<<grid_for(planes * (y / 2) * (z / 2), 256), 256, 0, stream>>>((const uint4*)in, (const uint4*)dout, (uint4*)din, planes, y, z);
  return check_launch("maxpool2_bwd");
}

int bcp_upsample2_fwd(const void* in, void* out, long long planes, int y, int z, cudaStream_t stream) {
  BCP_REQUIRE(in && out && planes > 0 && y > 0 && z > 0, "upsample2_fwd: bad args");
  upsample2_fwd_kernel<<<grid_for(planes * 4 * y * z, 256), 256, 0, stream>>>((const uint4*)in, (uint4*)out, planes, y, z);
  return check_launch("upsample2_fwd");
}

int bcp_upsample2_bwd(const void* dout, void* din, long long planes, int y, int z, cudaStream_t stream) {
  BCP_REQUIRE(dout && din && planes > 0 && y > 0 && z > 0, "upsample2_bwd: bad args");
  upsample2_bwd_kernel<<<grid_for(planes * y * z, 256), 256, 0, stream>>>((const uint4*)dout, (uint4*)din, planes, y, z);
  return check_launch("upsample2_bwd");
}

int bcp_maxpool3d_k3s2(const void* in, float* out, int n, int c, int x, int y, int z, cudaStream_t stream) {
  BCP_REQUIRE(in && out && n > 0 && c > 0 && x >= 3 && y >= 3 && z >= 3, "maxpool3d_k3s2: bad args");
  const int xo = (x - 3) / 2 + 1, yo = (y - 3) / 2 + 1, zo = (z - 3) / 2 + 1;
  const long long total = (long long)n * ((c + 7) / 8) * xo * yo * zo;
  maxpool3d_k3s2_kernel<<<grid_for(total, 128), 128, 0, stream>>>((const uint4*)in, out, n, c, x, y, z, xo, yo, zo);
  return check_launch("maxpool3d_k3s2");
}

}  // extern "C"
