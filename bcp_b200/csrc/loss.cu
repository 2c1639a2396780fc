// Fused mask-weighted Dice + cross-entropy ("mix_loss") forward and backward.
//   form 0 (LA / Pancreas): utils/BCP_utils.py:58-69 + utils/losses.py:47-77 (== pancreas/losses.py:82-141)
//           per-(n,c) soft Dice on softmax probabilities, smooth 1e-5, mean over (n,c)
//   form 1 (ACDC): ACDC_BCP_train.py:167-179 + utils/losses.py:102-134
//           batch-global per-class Dice (2*sum(p*t*m)+1e-10)/(sum(p^2*m)+sum(t*m)+1e-10), mean over classes
// The box mask M is implicit (0 inside the box, 1 outside): voxels outside use target `lab_img` with
// weight w_img, voxels inside use `lab_patch` with weight w_patch -- exactly the two masked terms of
// the reference, evaluated in ONE pass over the logits (the reference runs ~60 kernels and 8 host
// syncs per call).  Reductions: per-thread -> warp shuffle -> block -> fixed-order final reduce in
// double (deterministic).  Backward recomputes the softmax and writes dlogits in one pass.
#include "common.cuh"
#include "../../include/bcp_b200.h"
#include <math.h>

namespace bcp {

constexpr int LT = 256;

struct BoxArgs {
  int X, Y, Z, x0, y0, z0, x1, y1, z1;
};

// box {x0,y0,z0,px,py,pz} in device memory (replayable CUDA graphs); clipped like Python slicing
__device__ __forceinline__ BoxArgs load_box(const int* __restrict__ box, int X, int Y, int Z) {
  BoxArgs b;
  b.X = X; b.Y = Y; b.Z = Z;
  b.x0 = __ldg(box); b.y0 = __ldg(box + 1); b.z0 = __ldg(box + 2);
  b.x1 = min(b.x0 + __ldg(box + 3), X); b.y1 = min(b.y0 + __ldg(box + 4), Y); b.z1 = min(b.z0 + __ldg(box + 5), Z);
  return b;
}

__device__ __forceinline__ int in_box(long long v, const BoxArgs& b) {
  const int z = (int)(v % b.Z);
  const long long r = v / b.Z;
  const int y = (int)(r % b.Y);
  const int x = (int)(r / b.Y);
  return (x >= b.x0) & (x < b.x1) & (y >= b.y0) & (y < b.y1) & (z >= b.z0) & (z < b.z1);
}

// partial[(n*blocks + blk)*K + ...], K = 2*C*3 + 4:   [s][c][{I,A,B}] then ce[s], count[s]
template <int C, int FORM>
__global__ void __launch_bounds__(LT) mix_loss_fwd_kernel(const float* __restrict__ logits,
                                                           const unsigned char* __restrict__ lab_img,
                                                           const unsigned char* __restrict__ lab_patch,
                                                           const unsigned char* __restrict__ mask,
                                                           float* __restrict__ partial, long long V, int X, int Y, int Z,
                                                           const int* __restrict__ box_dev) {
  constexpr int K = 2 * C * 3 + 4;
  const BoxArgs box = load_box(box_dev, X, Y, Z);
  const int n = blockIdx.y, blocks = gridDim.x;
  const bool vec4 = (Z % 4 == 0) && ((((uintptr_t)logits | (uintptr_t)lab_img | (uintptr_t)lab_patch | (uintptr_t)mask) & 15) == 0);
  long long per = (V + blocks - 1) / blocks;
  if (vec4) per = (per + 3) & ~3ll;
  const long long v0 = min(V, (long long)blockIdx.x * per), v1 = min(V, v0 + per);
  const float* lg = logits + (long long)n * C * V;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  auto voxel = [&](const float (&x)[C], int s, int t) {
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) m = fmaxf(m, x[c]);
    float p[C];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { p[c] = expf(x[c] - m); sum += p[c]; }
    const float inv = 1.f / sum, lse = m + logf(sum);
    float ce = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float pc = p[c] * inv;
      const float oh = (t == c) ? 1.f : 0.f;
      if (t == c) ce = lse - x[c];
      const float vi = pc * oh;
      const float va = (FORM == 0) ? (pc + oh) : oh;
      const float vb = (FORM == 0) ? 0.f : pc * pc;
      // branch-free select of the set (s = 0/1)
      const float w1 = (float)s, w0 = 1.f - w1;
      acc[(0 * C + c) * 3 + 0] += w0 * vi; acc[(0 * C + c) * 3 + 1] += w0 * va; acc[(0 * C + c) * 3 + 2] += w0 * vb;
      acc[(1 * C + c) * 3 + 0] += w1 * vi; acc[(1 * C + c) * 3 + 1] += w1 * va; acc[(1 * C + c) * 3 + 2] += w1 * vb;
    }
    acc[2 * C * 3 + 0] += (1.f - (float)s) * ce;
    acc[2 * C * 3 + 1] += (float)s * ce;
    acc[2 * C * 3 + 2] += 1.f - (float)s;      // voxel counts per set (exact in fp32 up to 2^24 per thread)
    acc[2 * C * 3 + 3] += (float)s;
  };
  if (vec4) {
    // four consecutive z per thread (Z % 4 == 0: a run never leaves its row): 16-byte logit loads, 4-byte label loads,
    // one (x, y) decomposition per run -- four times the bytes in flight per thread of the scalar loop
    for (long long v = v0 + 4 * threadIdx.x; v < v1; v += 4 * LT) {
      float4 xv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(lg + (long long)c * V + v);
      const uchar4 li = *reinterpret_cast<const uchar4*>(lab_img + (long long)n * V + v);
      const uchar4 lp = *reinterpret_cast<const uchar4*>(lab_patch + (long long)n * V + v);
      int s4[4];
      if (mask) {
        const uchar4 mk = *reinterpret_cast<const uchar4*>(mask + (long long)n * V + v);
        s4[0] = mk.x ? 0 : 1; s4[1] = mk.y ? 0 : 1; s4[2] = mk.z ? 0 : 1; s4[3] = mk.w ? 0 : 1;
      } else {
        const int z = (int)(v % box.Z);
        const long long r = v / box.Z;
        const int y = (int)(r % box.Y), xx = (int)(r / box.Y);
        const int xy = (xx >= box.x0) & (xx < box.x1) & (y >= box.y0) & (y < box.y1);
#pragma unroll
        for (int k = 0; k < 4; ++k) s4[k] = xy & (z + k >= box.z0) & (z + k < box.z1);
      }
      const int ti[4] = {li.x, li.y, li.z, li.w}, tp[4] = {lp.x, lp.y, lp.z, lp.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float x[C];
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = reinterpret_cast<const float*>(&xv[c])[k];
        voxel(x, s4[k], s4[k] ? tp[k] : ti[k]);
      }
    }
  } else {
#pragma unroll 2
    for (long long v = v0 + threadIdx.x; v < v1; v += LT) {
      float x[C];
#pragma unroll
      for (int c = 0; c < C; ++c) x[c] = lg[(long long)c * V + v];
      const int s = mask ? (mask[(long long)n * V + v] ? 0 : 1) : in_box(v, box);
      const int t = s ? lab_patch[(long long)n * V + v] : lab_img[(long long)n * V + v];
      voxel(x, s, t);
    }
  }
  __shared__ float red[K * (LT / 32)];
  block_sum<K, LT>(acc, red);
  if (threadIdx.x == 0) {
    float* dst = partial + ((long long)n * blocks + blockIdx.x) * K;
#pragma unroll
    for (int k = 0; k < K; ++k) dst[k] = acc[k];
  }
}

// ctx layout (floats): [0]=loss=(dice+ce)/2, [1]=dice, [2]=ce, [3]=unused, [4..5]=ce coef per set,
// then [N][2][C][3] = {cA, cB, cC}:   dDice/dp_c(v) = cA*[T==c] + cB + cC*p_c   (v in set s, sample n)
// deterministic warp-parallel sum of partial[(n*blocks + b)*K + k] over b (lane-strided, then a fixed shuffle tree)
template <int K>
__device__ __forceinline__ void sum_partials(const float* __restrict__ partial, int n, int blocks, double (&a)[K]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < K; ++k) a[k] = 0.0;
  for (int b = lane; b < blocks; b += 32) {
    const float* src = partial + ((long long)n * blocks + b) * K;
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] += (double)src[k];
  }
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
}

// ctx layout (floats): [0]=loss=(dice+ce)/2, [1]=dice, [2]=ce, [3]=unused, [4..5]=ce coef per set,
// then [N][2][C][3] = {cA, cB, cC}:   dDice/dp_c(v) = cA*[T==c] + cB + cC*p_c   (v in set s, sample n)
template <int C, int FORM>
__global__ void mix_loss_finalize_kernel(const float* __restrict__ partial, float* __restrict__ ctx, int N, int blocks,
                                         float w_img, float w_patch) {
  constexpr int K = 2 * C * 3 + 4;
  const bool writer = threadIdx.x == 0;
  const double w[2] = {(double)w_img, (double)w_patch};
  double cnt[2] = {0.0, 0.0};     // mask.sum() / (1-mask).sum() over the whole [N, ...] loss mask
  double dice = 0.0, ce = 0.0;
  double cesum[2] = {0.0, 0.0};
  float* tab = ctx + 6;
  if (FORM == 0) {
    const double eps = 1e-5;
    double dsum[2] = {0.0, 0.0};
    for (int n = 0; n < N; ++n) {
      double a[K];
      sum_partials<K>(partial, n, blocks, a);
      for (int s = 0; s < 2; ++s) {
        for (int c = 0; c < C; ++c) {
          const double I = a[(s * C + c) * 3], U = a[(s * C + c) * 3 + 1];
          const double D = (2.0 * I + eps) / (U + eps);
          dsum[s] += D;
          const double k0 = -w[s] / ((double)N * C);
          if (writer) {
            float* t = tab + (((long long)n * 2 + s) * C + c) * 3;
            t[0] = (float)(k0 * 2.0 / (U + eps));
            t[1] = (float)(-k0 * D / (U + eps));
            t[2] = 0.f;
          }
        }
        cesum[s] += a[2 * C * 3 + s];
        cnt[s] += a[2 * C * 3 + 2 + s];
      }
    }
    for (int s = 0; s < 2; ++s) dice += w[s] * (1.0 - dsum[s] / ((double)N * C));
  } else {
    const double eps = 1e-10;
    double a[K];
    for (int k = 0; k < K; ++k) a[k] = 0.0;
    for (int n = 0; n < N; ++n) {
      double an[K];
      sum_partials<K>(partial, n, blocks, an);
      for (int k = 0; k < K; ++k) a[k] += an[k];
    }
    for (int s = 0; s < 2; ++s) {
      double dl = 0.0;
      for (int c = 0; c < C; ++c) {
        const double I = a[(s * C + c) * 3], Y = a[(s * C + c) * 3 + 1], Zs = a[(s * C + c) * 3 + 2];
        const double den = Zs + Y + eps;
        dl += 1.0 - (2.0 * I + eps) / den;
        const double cA = -w[s] / C * 2.0 / den;
        const double cC = w[s] / C * 2.0 * (2.0 * I + eps) / (den * den);
        if (writer)
          for (int n = 0; n < N; ++n) {
            float* t = tab + (((long long)n * 2 + s) * C + c) * 3;
            t[0] = (float)cA; t[1] = 0.f; t[2] = (float)cC;
          }
      }
      dice += w[s] * dl / C;
      cesum[s] = a[2 * C * 3 + s];
      cnt[s] = a[2 * C * 3 + 2 + s];
    }
  }
  if (writer) {
    for (int s = 0; s < 2; ++s) {
      ce += w[s] * cesum[s] / (cnt[s] + 1e-16);
      ctx[4 + s] = (float)(w[s] / (cnt[s] + 1e-16));
    }
    ctx[0] = (float)((dice + ce) * 0.5);
    ctx[1] = (float)dice;
    ctx[2] = (float)ce;
    ctx[3] = 0.f;
  }
}

template <int C>
__global__ void __launch_bounds__(LT) mix_loss_bwd_kernel(const float* __restrict__ logits,
                                                           const unsigned char* __restrict__ lab_img,
                                                           const unsigned char* __restrict__ lab_patch,
                                                           const unsigned char* __restrict__ mask,
                                                           const float* __restrict__ ctx, const float* __restrict__ grad3,
                                                           float* __restrict__ dlogits, int N, long long V, int X, int Y, int Z,
                                                           const int* __restrict__ box_dev) {
  const BoxArgs box = load_box(box_dev, X, Y, Z);
  // outputs were {loss=(dice+ce)/2, dice, ce}: fold the three upstream gradients
  const float gd = grad3[1] + 0.5f * grad3[0], gc = grad3[2] + 0.5f * grad3[0];
  const long long total = (long long)N * V;
  const long long stride = (long long)gridDim.x * LT;
  const float* tab = ctx + 6;
  auto voxel = [&](const float (&x)[C], long long n, int s, int t, float (&d)[C]) {
    float p[C], g[C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) m = fmaxf(m, x[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { p[c] = expf(x[c] - m); sum += p[c]; }
    const float inv = 1.f / sum;
    const float cec = gc * ctx[4 + s];
    const float* tb = tab + ((n * 2 + s) * C) * 3;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      p[c] *= inv;
      const float oh = (t == c) ? 1.f : 0.f;
      g[c] = gd * (tb[c * 3] * oh + tb[c * 3 + 1] + tb[c * 3 + 2] * p[c]);
      dot += p[c] * g[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float oh = (t == c) ? 1.f : 0.f;
      d[c] = p[c] * (g[c] - dot) + cec * (p[c] - oh);
    }
  };
  const bool vec4 = (Z % 4 == 0) &&
                    ((((uintptr_t)logits | (uintptr_t)dlogits | (uintptr_t)lab_img | (uintptr_t)lab_patch | (uintptr_t)mask) & 15) == 0);
  if (vec4) {
    for (long long i = 4 * ((long long)blockIdx.x * LT + threadIdx.x); i < total; i += 4 * stride) {
      const long long n = i / V, v = i - n * V;
      const float* lg = logits + n * C * V + v;
      float4 xv[C], dv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(lg + (long long)c * V);
      const uchar4 li = *reinterpret_cast<const uchar4*>(lab_img + n * V + v);
      const uchar4 lp = *reinterpret_cast<const uchar4*>(lab_patch + n * V + v);
      int s4[4];
      if (mask) {
        const uchar4 mk = *reinterpret_cast<const uchar4*>(mask + n * V + v);
        s4[0] = mk.x ? 0 : 1; s4[1] = mk.y ? 0 : 1; s4[2] = mk.z ? 0 : 1; s4[3] = mk.w ? 0 : 1;
      } else {
        const int z = (int)(v % box.Z);
        const long long r = v / box.Z;
        const int y = (int)(r % box.Y), xx = (int)(r / box.Y);
        const int xy = (xx >= box.x0) & (xx < box.x1) & (y >= box.y0) & (y < box.y1);
#pragma unroll
        for (int k = 0; k < 4; ++k) s4[k] = xy & (z + k >= box.z0) & (z + k < box.z1);
      }
      const int ti[4] = {li.x, li.y, li.z, li.w}, tp[4] = {lp.x, lp.y, lp.z, lp.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float x[C], d[C];
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = reinterpret_cast<const float*>(&xv[c])[k];
        voxel(x, n, s4[k], s4[k] ? tp[k] : ti[k], d);
#pragma unroll
        for (int c = 0; c < C; ++c) reinterpret_cast<float*>(&dv[c])[k] = d[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) *reinterpret_cast<float4*>(dlogits + n * C * V + (long long)c * V + v) = dv[c];
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * LT + threadIdx.x; i < total; i += stride) {
    const long long n = i / V, v = i - n * V;
    const float* lg = logits + n * C * V + v;
    float x[C], d[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = lg[(long long)c * V];
    const int s = mask ? (mask[n * V + v] ? 0 : 1) : in_box(v, box);
    const int t = s ? lab_patch[n * V + v] : lab_img[n * V + v];
    voxel(x, n, s, t, d);
#pragma unroll
    for (int c = 0; c < C; ++c) dlogits[n * C * V + (long long)c * V + v] = d[c];
  }
}


// ------------------------------------------------------------------------------------------------------------
// DiceLoss on PROBABILITIES (utils/losses.py:113-134 as called from ACDC_BCP_train.py:170,175 with softmax=False):
// batch-global per-class  1 - (2*sum(s*t*m) + 1e-10) / (sum(s*s*m) + sum(t*t*m) + 1e-10), mean over classes.
// The drop-in utils.losses.DiceLoss uses this pair so the reference's own mix_loss body (which calls F.softmax itself)
// runs unchanged.  partial[(blk*C + c)*3 + {I,Y,Z}]; ctx = {loss, -, cA[C], cC[C]}: dL/ds_c(v) = m*(cA_c*[t==c] + cC_c*s).
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(LT) dice_prob_fwd_kernel(const float* __restrict__ probs, const unsigned char* __restrict__ target,
                                                            const unsigned char* __restrict__ mask, float* __restrict__ partial,
                                                            int N, long long V) {
  float acc[3 * C];
#pragma unroll
  for (int k = 0; k < 3 * C; ++k) acc[k] = 0.f;
  const long long total = (long long)N * V, stride = (long long)gridDim.x * LT;
  for (long long i = (long long)blockIdx.x * LT + threadIdx.x; i < total; i += stride) {
    const long long n = i / V, v = i - n * V;
    const float m = mask ? (mask[i] ? 1.f : 0.f) : 1.f;
    const int t = target[i];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float sc = probs[(n * C + c) * V + v];
      const float oh = (t == c) ? 1.f : 0.f;
      acc[c * 3 + 0] += sc * oh * m;
      acc[c * 3 + 1] += oh * m;
      acc[c * 3 + 2] += sc * sc * m;
    }
  }
  __shared__ float red[3 * C * (LT / 32)];
  block_sum<3 * C, LT>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 3 * C; ++k) partial[(long long)blockIdx.x * 3 * C + k] = acc[k];
  }
}

template <int C>
__global__ void dice_prob_finalize_kernel(const float* __restrict__ partial, float* __restrict__ ctx, int blocks) {
  double a[3 * C];
  sum_partials<3 * C>(partial, 0, blocks, a);
  if (threadIdx.x != 0) return;
  const double eps = 1e-10;
  double loss = 0.0;
  for (int c = 0; c < C; ++c) {
    const double I = a[c * 3], Y = a[c * 3 + 1], Z = a[c * 3 + 2];
    const double den = Z + Y + eps;
    loss += 1.0 - (2.0 * I + eps) / den;
    ctx[2 + c] = (float)(-2.0 / den / C);
    ctx[2 + C + c] = (float)(2.0 * (2.0 * I + eps) / (den * den) / C);
  }
  ctx[0] = (float)(loss / C);
  ctx[1] = 0.f;
}

template <int C>
__global__ void __launch_bounds__(LT) dice_prob_bwd_kernel(const float* __restrict__ probs, const unsigned char* __restrict__ target,
                                                            const unsigned char* __restrict__ mask, const float* __restrict__ ctx,
                                                            const float* __restrict__ gout, float* __restrict__ dprobs, int N, long long V) {
  const float g = gout[0];
  const long long total = (long long)N * V, stride = (long long)gridDim.x * LT;
  for (long long i = (long long)blockIdx.x * LT + threadIdx.x; i < total; i += stride) {
    const long long n = i / V, v = i - n * V;
    const float m = mask ? (mask[i] ? 1.f : 0.f) : 1.f;
    const int t = target[i];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const long long j = (n * C + c) * V + v;
      const float oh = (t == c) ? 1.f : 0.f;
      dprobs[j] = g * m * (ctx[2 + c] * oh + ctx[2 + C + c] * probs[j]);
    }
  }
}

static inline int loss_blocks(long long V) {
  long long b = (V + 4095) / 4096;       // with the per-sample grid dimension: ~4 resident blocks per SM on the LA volumes
  if (b < 1) b = 1;
  if (b > 296) b = 296;
  return (int)b;
}

}  // namespace bcp

using namespace bcp;

extern "C" {

long long bcp_mix_loss_ctx_floats(int n, int c) { return 6 + (long long)n * 2 * c * 3; }
long long bcp_mix_loss_workspace_floats(int n, int c, long long v) { return (long long)n * loss_blocks(v) * (2 * c * 3 + 4); }

int bcp_mix_loss_fwd(const float* logits, const unsigned char* lab_img, const unsigned char* lab_patch,
                     const unsigned char* mask, float* ctx,
                     float* workspace, int n, int c, int X, int Y, int Z, const int* box6, int form, float w_img,
                     float w_patch, cudaStream_t stream) {
  BCP_REQUIRE(logits && lab_img && lab_patch && ctx && workspace && box6, "mix_loss_fwd: null pointer");
  BCP_REQUIRE(c == 2 || c == 4, "mix_loss_fwd: %d classes unsupported (2 or 4)", c);
  BCP_REQUIRE(form == 0 || form == 1, "mix_loss_fwd: form");
  BCP_REQUIRE(n > 0 && X > 0 && Y > 0 && Z > 0, "mix_loss_fwd: bad shape");
  const long long V = (long long)X * Y * Z;
  const int blocks = loss_blocks(V);
  dim3 grid(blocks, n);
#define LAUNCH(CC, FF)                                                                                         \
  mix_loss_fwd_kernel<CC, FF><<<grid, LT, 0, stream>>>(logits, lab_img, lab_patch, mask, workspace, V, X, Y, Z, box6);     \
  mix_loss_finalize_kernel<CC, FF><<<1, 32, 0, stream>>>(workspace, ctx, n, blocks, w_img, w_patch);
  if (c == 2 && form == 0) { LAUNCH(2, 0) }
  else if (c == 4 && form == 0) { LAUNCH(4, 0) }
  else if (c == 2 && form == 1) { LAUNCH(2, 1) }
  else { LAUNCH(4, 1) }
#undef LAUNCH
  return check_launch("mix_loss_fwd");
}

int bcp_mix_loss_bwd(const float* logits, const unsigned char* lab_img, const unsigned char* lab_patch,
                     const unsigned char* mask, const float* ctx, const float* grad3, float* dlogits,
                     int n, int c, int X, int Y, int Z, const int* box6, cudaStream_t stream) {
  BCP_REQUIRE(logits && lab_img && lab_patch && ctx && grad3 && dlogits && box6, "mix_loss_bwd: null pointer");
  BCP_REQUIRE(c == 2 || c == 4, "mix_loss_bwd: %d classes unsupported", c);
  const long long V = (long long)X * Y * Z;
  long long blocks = ((long long)n * V + LT - 1) / LT;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (c == 2)
    mix_loss_bwd_kernel<2><<<(int)blocks, LT, 0, stream>>>(logits, lab_img, lab_patch, mask, ctx, grad3, dlogits, n, V, X, Y, Z, box6);
  else
    mix_loss_bwd_kernel<4><<<(int)blocks, LT, 0, stream>>>(logits, lab_img, lab_patch, mask, ctx, grad3, dlogits, n, V, X, Y, Z, box6);
  return check_launch("mix_loss_bwd");
}

long long bcp_dice_prob_ctx_floats(int c) { return 2 + 2 * (long long)c; }
long long bcp_dice_prob_workspace_floats(int n, int c, long long v) { return (long long)loss_blocks((long long)n * v) * 3 * c; }

int bcp_dice_prob_fwd(const float* probs, const unsigned char* target, const unsigned char* mask, float* ctx, float* workspace,
                      int n, int c, long long v, cudaStream_t stream) {
  BCP_REQUIRE(probs && target && ctx && workspace, "dice_prob_fwd: null pointer");
  BCP_REQUIRE(c == 2 || c == 4, "dice_prob_fwd: %d classes unsupported (2 or 4)", c);
  BCP_REQUIRE(n > 0 && v > 0, "dice_prob_fwd: bad shape");
  const int blocks = loss_blocks((long long)n * v);
  if (c == 2) { dice_prob_fwd_kernel<2><<<blocks, LT, 0, stream>>>(probs, target, mask, workspace, n, v); dice_prob_finalize_kernel<2><<<1, 32, 0, stream>>>(workspace, ctx, blocks); }
  else { dice_prob_fwd_kernel<4><<<blocks, LT, 0, stream>>>(probs, target, mask, workspace, n, v); dice_prob_finalize_kernel<4><<<1, 32, 0, stream>>>(workspace, ctx, blocks); }
  return check_launch("dice_prob_fwd");
}

int bcp_dice_prob_bwd(const float* probs, const unsigned char* target, const unsigned char* mask, const float* ctx,
                      const float* grad_out, float* dprobs, int n, int c, long long v, cudaStream_t stream) {
  BCP_REQUIRE(probs && target && ctx && grad_out && dprobs, "dice_prob_bwd: null pointer");
  BCP_REQUIRE(c == 2 || c == 4, "dice_prob_bwd: %d classes unsupported", c);
  long long blocks = ((long long)n * v + LT - 1) / LT;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (c == 2) dice_prob_bwd_kernel<2><<<(int)blocks, LT, 0, stream>>>(probs, target, mask, ctx, grad_out, dprobs, n, v);
  else dice_prob_bwd_kernel<4><<<(int)blocks, LT, 0, stream>>>(probs, target, mask, ctx, grad_out, dprobs, n, v);
  return check_launch("dice_prob_bwd");
}

}  // extern "C"
