// Bandwidth-bound step primitives: box mask-mix, pseudo-labels, fused SGD/Adam+EMA, weight repack,
// layout converts.  All kernels: 128-bit vectorised where alignment allows, grid-stride, no atomics.
#include "common.cuh"
#include "../../include/bcp_b200.h"
#include <math.h>

namespace bcp {

// ------------------------------------------------------------------------------------------
// mask-mix:  out = a*M + b*(1-M), M = 0 inside the box, 1 outside (reference: LA_BCP_train.py:248-249,
// ACDC_BCP_train.py:372-373, train_pancreas.py:155-156).  Computed literally with two multiplies and
// one add (no select, no FMA) so -0.0 / inf / NaN behave like the reference's tensor expression.
// ------------------------------------------------------------------------------------------
// The box {x0,y0,z0,px,py,pz} lives in device memory so a captured CUDA graph can be replayed with a new box each step;
// it is clipped to the volume like Python slicing (mask[w:w+px, ...]).
__global__ void mask_mix_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                long long total, int X, int Y, int Z, const int* __restrict__ box) {
  const int bx0 = __ldg(box), by0 = __ldg(box + 1), bz0 = __ldg(box + 2);
  const int bx1 = min(bx0 + __ldg(box + 3), X), by1 = min(by0 + __ldg(box + 4), Y), bz1 = min(bz0 + __ldg(box + 5), Z);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long r = i;
    const int z = (int)(r % Z); r /= Z;
    const int y = (int)(r % Y); r /= Y;
    const int x = (int)(r % X);
    const bool inside = (x >= bx0) & (x < bx1) & (y >= by0) & (y < by1) & (z >= bz0) & (z < bz1);
    const float m = inside ? 0.f : 1.f;
    out[i] = __fadd_rn(__fmul_rn(a[i], m), __fmul_rn(b[i], 1.f - m));
  }
}

__global__ void label_mix_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b,
                                 unsigned char* __restrict__ out, long long total, int X, int Y, int Z,
                                 const int* __restrict__ box) {
  const int bx0 = __ldg(box), by0 = __ldg(box + 1), bz0 = __ldg(box + 2);
  const int bx1 = min(bx0 + __ldg(box + 3), X), by1 = min(by0 + __ldg(box + 4), Y), bz1 = min(bz0 + __ldg(box + 5), Z);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long r = i;
    const int z = (int)(r % Z); r /= Z;
    const int y = (int)(r % Y); r /= Y;
    const int x = (int)(r % X);
    const bool inside = (x >= bx0) & (x < bx1) & (y >= by0) & (y < by1) & (z >= bz0) & (z < bz1);
    out[i] = inside ? b[i] : a[i];
  }
}

// 128-bit versions (Z % 4 == 0 / Z % 16 == 0, 16-byte aligned bases, < 2^31 elements): one thread moves a run of 4 floats
// (16 labels) that lies inside ONE z-line, so the (x, y) part of the box test and the 32-bit index decomposition happen once
// per run; only the z compare is per element.  Same literal arithmetic per element as the scalar kernel (bit-exact).
__global__ void __launch_bounds__(256) mask_mix_vec4_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                            float4* __restrict__ out, unsigned total4, unsigned X, unsigned Y,
                                                            unsigned Z4, int Z, const int* __restrict__ box) {
  const int bx0 = __ldg(box), by0 = __ldg(box + 1), bz0 = __ldg(box + 2);
  const int bx1 = min(bx0 + __ldg(box + 3), (int)X), by1 = min(by0 + __ldg(box + 4), (int)Y), bz1 = min(bz0 + __ldg(box + 5), Z);
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    unsigned r = i;
    const int z = (int)(r % Z4) * 4; r /= Z4;
    const int y = (int)(r % Y); r /= Y;
    const int x = (int)(r % X);
    const bool row = (x >= bx0) & (x < bx1) & (y >= by0) & (y < by1);
    const float4 va = __ldg(a + i), vb = __ldg(b + i);
    const float m0 = (row & (z >= bz0) & (z < bz1)) ? 0.f : 1.f;
    const float m1 = (row & (z + 1 >= bz0) & (z + 1 < bz1)) ? 0.f : 1.f;
    const float m2 = (row & (z + 2 >= bz0) & (z + 2 < bz1)) ? 0.f : 1.f;
    const float m3 = (row & (z + 3 >= bz0) & (z + 3 < bz1)) ? 0.f : 1.f;
    float4 o;
    o.x = __fadd_rn(__fmul_rn(va.x, m0), __fmul_rn(vb.x, 1.f - m0));
    o.y = __fadd_rn(__fmul_rn(va.y, m1), __fmul_rn(vb.y, 1.f - m1));
    o.z = __fadd_rn(__fmul_rn(va.z, m2), __fmul_rn(vb.z, 1.f - m2));
    o.w = __fadd_rn(__fmul_rn(va.w, m3), __fmul_rn(vb.w, 1.f - m3));
    out[i] = o;
  }
}

__global__ void __launch_bounds__(256) label_mix_vec16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                              uint4* __restrict__ out, unsigned total16, unsigned X, unsigned Y,
                                                              unsigned Z16, int Z, const int* __restrict__ box) {
  const int bx0 = __ldg(box), by0 = __ldg(box + 1), bz0 = __ldg(box + 2);
  const int bx1 = min(bx0 + __ldg(box + 3), (int)X), by1 = min(by0 + __ldg(box + 4), (int)Y), bz1 = min(bz0 + __ldg(box + 5), Z);
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += stride) {
    unsigned r = i;
    const int z = (int)(r % Z16) * 16; r /= Z16;
    const int y = (int)(r % Y); r /= Y;
    const int x = (int)(r % X);
    const bool row = (x >= bx0) & (x < bx1) & (y >= by0) & (y < by1);
    const uint4 va = __ldg(a + i), vb = __ldg(b + i);
    // byte-select mask: 0xFF where the voxel lies inside the box (take b)
    unsigned sel[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      unsigned mword = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int zz = z + w * 4 + k;
        if (row & (zz >= bz0) & (zz < bz1)) mword |= 0xFFu << (8 * k);
      }
      sel[w] = mword;
    }
    uint4 o;
    o.x = (va.x & ~sel[0]) | (vb.x & sel[0]);
    o.y = (va.y & ~sel[1]) | (vb.y & sel[1]);
    o.z = (va.z & ~sel[2]) | (vb.z & sel[2]);
    o.w = (va.w & ~sel[3]) | (vb.w & sel[3]);
    out[i] = o;
  }
}

// ------------------------------------------------------------------------------------------
// pseudo labels.  thresh: (softmax(x,1) >= thr)[:,1]   (LA_BCP_train.py:57-60, pancreas_utils.py:275-278)
//                 argmax: torch.max(softmax(x,1),1)[1] (ACDC_BCP_train.py:112-114)
// fp32 softmax is replicated step by step (max, expf(x-max), sum, IEEE divide) because
// "p1 >= 0.5" is not "x1 >= x0" once exp rounds to 1.0f.
// ------------------------------------------------------------------------------------------
// exp(d) for d <= 0 rounded like a <=1-ulp libm (the reference's CPU path) where it matters: for |d| < 2^-10 the
// result is 1 + d + d^2/2 to well below half an ulp, so evaluate that directly; elsewhere p is far from any tie.
__device__ __forceinline__ float exp_near_one(float d) {
  if (fabsf(d) < 9.765625e-4f) return __fadd_rn(1.0f, __fadd_rn(d, __fmul_rn(__fmul_rn(d, d), 0.5f)));
  return expf(d);
}

template <int C>
__global__ void pseudo_label_kernel(const float* __restrict__ logits, unsigned char* __restrict__ out,
                                    int N, long long V, int mode, float thr) {
  const long long total = (long long)N * V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long n = i / V, v = i - n * V;
    const float* p = logits + n * C * V + v;
    float x[C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { x[c] = p[(long long)c * V]; m = fmaxf(m, x[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { x[c] = exp_near_one(x[c] - m); s += x[c]; }
    if (mode == 0) {
      out[i] = (__fdiv_rn(x[1], s) >= thr) ? 1 : 0;
    } else {
      float best = __fdiv_rn(x[0], s);
      int bi = 0;
#pragma unroll
      for (int c = 1; c < C; ++c) {
        const float pc = __fdiv_rn(x[c], s);
        if (pc > best) { best = pc; bi = c; }
      }
      out[i] = (unsigned char)bi;
    }
  }
}

// 128-bit version (V % 4 == 0): four voxels per thread, one float4 per class plane, one 32-bit store of four labels.
template <int C>
__global__ void __launch_bounds__(256) pseudo_label_vec4_kernel(const float4* __restrict__ logits, uchar4* __restrict__ out,
                                                                unsigned N, unsigned V4, int mode, float thr) {
  const unsigned total4 = N * V4, stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const unsigned n = i / V4, v = i - n * V4;
    const float4* p = logits + (size_t)n * C * V4 + v;
    float x[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 t = __ldg(p + (size_t)c * V4);
      x[c][0] = t.x; x[c][1] = t.y; x[c][2] = t.z; x[c][3] = t.w;
    }
    unsigned char lab[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < C; ++c) m = fmaxf(m, x[c][k]);
      float e[C];
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) { e[c] = exp_near_one(x[c][k] - m); s += e[c]; }
      if (mode == 0) {
        lab[k] = (__fdiv_rn(e[1], s) >= thr) ? 1 : 0;
      } else {
        float best = __fdiv_rn(e[0], s);
        int bi = 0;
#pragma unroll
        for (int c = 1; c < C; ++c) {
          const float pc = __fdiv_rn(e[c], s);
          if (pc > best) { best = pc; bi = c; }
        }
        lab[k] = (unsigned char)bi;
      }
    }
    out[i] = make_uchar4(lab[0], lab[1], lab[2], lab[3]);
  }
}

// ------------------------------------------------------------------------------------------
// fused SGD(momentum, weight decay) + EMA over a flat fp32 arena (optim.SGD semantics,
// LA_BCP_train.py:218 + utils/BCP_utils.py:78-81).  hyper = {lr, momentum, wd, ema_alpha, grad_scale, 1-ema_alpha}
// lives on the device so a captured CUDA graph survives LR decay.
//   g = grad*grad_scale + wd*p ; buf = mom*buf + g ; p -= lr*buf ; e = e*alpha + (1-alpha)*p
// elements [n_train, n_total) are EMA-only (never-trained MLP heads / BN buffers for ACDC).
// ------------------------------------------------------------------------------------------
// One thread owns four consecutive floats: 128-bit loads/stores of p, g, buf, e (the arenas are allocator-aligned); the
// n_train boundary (EMA-only tail) is a per-element predicate inside the vector.
__global__ void __launch_bounds__(256) sgd_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                                      float* __restrict__ e, const float* __restrict__ hyper, long long n_train,
                                                      long long n_total) {
  const float lr = hyper[0], mom = hyper[1], wd = hyper[2], alpha = hyper[3], gs = hyper[4];
  const float one_m_alpha = hyper[5];   // float(1 - alpha) evaluated in double on the host, like the Python expression
  const long long n4 = (n_total + 3) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    const long long i0 = q << 2;
    if (i0 + 4 <= n_total) {
      float4 pv = *reinterpret_cast<const float4*>(p + i0);
      float* pf = reinterpret_cast<float*>(&pv);
      if (i0 < n_train) {
        if (i0 + 4 <= n_train) {
          const float4 gv = *reinterpret_cast<const float4*>(g + i0);
          float4 bv = *reinterpret_cast<const float4*>(buf + i0);
          const float* gf = reinterpret_cast<const float*>(&gv);
          float* bf = reinterpret_cast<float*>(&bv);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float gk = gf[k] * gs + wd * pf[k];
            bf[k] = mom * bf[k] + gk;
            pf[k] = pf[k] - lr * bf[k];
          }
          *reinterpret_cast<float4*>(buf + i0) = bv;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (i0 + k < n_train) {
              const float gk = g[i0 + k] * gs + wd * pf[k];
              const float bk = mom * buf[i0 + k] + gk;
              buf[i0 + k] = bk;
              pf[k] = pf[k] - lr * bk;
            }
        }
        *reinterpret_cast<float4*>(p + i0) = pv;
      }
      if (e != nullptr) {
        float4 ev = *reinterpret_cast<const float4*>(e + i0);
        float* ef = reinterpret_cast<float*>(&ev);
#pragma unroll
        for (int k = 0; k < 4; ++k) ef[k] = __fadd_rn(__fmul_rn(ef[k], alpha), __fmul_rn(one_m_alpha, pf[k]));
        *reinterpret_cast<float4*>(e + i0) = ev;
      }
    } else {
      for (long long i = i0; i < n_total; ++i) {
        float pv = p[i];
        if (i < n_train) {
          const float gv = g[i] * gs + wd * pv;
          const float bv = mom * buf[i] + gv;
          buf[i] = bv;
          pv = pv - lr * bv;
          p[i] = pv;
        }
        if (e != nullptr) e[i] = __fadd_rn(__fmul_rn(e[i], alpha), __fmul_rn(one_m_alpha, pv));
      }
    }
  }
}

// Adam (optim.Adam defaults, pancreas/dataloaders.py:182) + EMA.  hyper = {lr, beta1, beta2, eps, ema_alpha, grad_scale,
// bias_corr1, bias_corr2_sqrt, 1-ema_alpha, step_size = lr / bias_corr1, 1-beta1, 1-beta2} (the two complements are formed
// in double on the host like torch's Python expressions: 1.f - 0.999f is 1.3e-5 away from float(1 - 0.999)).  Slots 6, 7, 9 are produced ON THE DEVICE by
// adam_tick_kernel from a device-resident step counter (so a captured CUDA graph advances the bias corrections on every
// replay), in double precision like the Python expressions of torch.optim.Adam.
__global__ void adam_tick_kernel(float* __restrict__ hyper, long long* __restrict__ step) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const long long t = step[0] + 1;
    step[0] = t;
    const double b1 = (double)hyper[1], b2 = (double)hyper[2];
    const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
    hyper[6] = (float)bc1;
    hyper[7] = (float)sqrt(bc2);
    hyper[9] = (float)((double)hyper[0] / bc1);
  }
}

__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ e, const float* __restrict__ hyper,
                                                       long long n_train, long long n_total) {
  const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3], alpha = hyper[4], gs = hyper[5];
  const float bc2s = hyper[7], step_size = hyper[9];
  const float one_m_alpha = hyper[8];
  const float w1 = hyper[10], w2 = hyper[11];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += stride) {
    float pv = p[i];
    if (i < n_train) {
      const float gv = g[i] * gs;
      const float mv = fmaf(w1, gv - m[i], m[i]);              // exp_avg.lerp_(grad, 1 - beta1)
      const float vv = fmaf(w2 * gv, gv, v[i] * b2);            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
      m[i] = mv; v[i] = vv;
      const float denom = sqrtf(vv) / bc2s + eps;
      pv = pv - step_size * (mv / denom);
      p[i] = pv;
    }
    if (e != nullptr) e[i] = __fadd_rn(__fmul_rn(e[i], alpha), __fmul_rn(one_m_alpha, pv));
  }
}

// ACDC's state_dict EMA also blends int64 num_batches_tracked through float and truncates
// (ACDC_BCP_train.py:123-129).
__global__ void ema_i64_kernel(long long* __restrict__ e, const long long* __restrict__ p, int n, float alpha,
                               float one_m_alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float r = __fadd_rn(__fmul_rn(alpha, (float)e[i]), __fmul_rn(one_m_alpha, (float)p[i]));
    e[i] = (long long)r;  // truncation toward zero like tensor.copy_(float -> int64)
  }
}

// ------------------------------------------------------------------------------------------
// weight repack: fp32 master weights (PyTorch layouts) -> bf16 operand layouts used by the conv kernels.
// One launch handles every layer of a network from a device-resident job table.
//   source is always [A][B][T] fp32 (Conv3d/2d: A=Cout,B=Cin; Conv k2s2: A=Cout(half-res),B=Cin(full-res);
//   ConvTranspose k2s2: A=Cin(half-res),B=Cout(full-res)).
//   kind 0: [T][B/8][A][8]  inner = B index            -- conv fwd operand; stride-2 "gather" (full->half) operand
//   kind 1: [T][A/8][B][8]  inner = A index, taps flipped (t -> T-1-t)   -- conv dgrad operand
//   kind 2: [T][A/8][B][8]  inner = A index            -- stride-2 "scatter" (half->full) operand, CUDA-core kernel
//   kind 3: [A/8][T][B][8]  inner = A index            -- stride-2 "scatter" operand of the tcgen05 kernel
// Channel counts that are not multiples of 8 are zero-padded in the blocked dimension.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) repack_kernel(const float* __restrict__ arena, __nv_bfloat16* __restrict__ packed,
                                                     const bcp_repack_job* __restrict__ jobs, int njobs) {
  // One 8(a) x 8(b) x T tile of the PyTorch-layout weight [A][B][T] per iteration: read as eight contiguous runs of
  // 8*T floats, staged in shared memory, written as T runs of 64 bf16 (128 bytes) -- both sides coalesced.
  //   kind 0: dst [T][ceil(B/8)][A][8]  (inner = b)          kind 3: dst [ceil(A/8)][T][B][8]  (inner = a)
  //   kind 1/2: dst [T][ceil(A/8)][B][8] (inner = a), kind 1 flips the taps
  __shared__ float tile[64 * 27];
  for (int j = blockIdx.y; j < njobs; j += gridDim.y) {
    const bcp_repack_job job = jobs[j];
    const float* src = arena + job.src_off;
    __nv_bfloat16* dst = packed + job.dst_off;
    const int A = job.dim_a, B = job.dim_b, T = job.taps;
    const int Ab = (A + 7) / 8, Bb = (B + 7) / 8;
    const int n64 = 64 * T;
    for (int tl = blockIdx.x; tl < Ab * Bb; tl += gridDim.x) {
      const int ab = tl / Bb, bb = tl - ab * Bb;
      __syncthreads();
      for (int i = threadIdx.x; i < n64; i += 128) {
        const int pair = i / T, t = i - pair * T;
        const int a = ab * 8 + (pair >> 3), b = bb * 8 + (pair & 7);
        tile[i] = (a < A && b < B) ? src[((long long)a * B + b) * T + t] : 0.f;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < n64; i += 128) {
        const int t = i >> 6, r = i & 63;
        long long o;
        float v;
        if (job.kind == 0) {                       // r = a_local * 8 + b8
          v = tile[r * T + t];
          o = (((long long)t * Bb + bb) * A + ab * 8 + (r >> 3)) * 8 + (r & 7);
          if (ab * 8 + (r >> 3) >= A) continue;
        } else {                                   // r = b_local * 8 + a8
          const int bl = r >> 3, a8 = r & 7;
          if (bb * 8 + bl >= B) continue;
          const int ts = (job.kind == 1) ? (T - 1 - t) : t;
          v = tile[(a8 * 8 + bl) * T + ts];
          if (job.kind == 3) o = (((long long)ab * T + t) * B + bb * 8 + bl) * 8 + a8;
          else o = (((long long)t * Ab + ab) * B + bb * 8 + bl) * 8 + a8;
        }
        dst[o] = __float2bfloat16_rn(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// layout converts between planar fp32 NC(D)HW and channel-blocked bf16 CB8 [N][C/8][S][8]
// ------------------------------------------------------------------------------------------
__global__ void planar_to_cb8_kernel(const float* __restrict__ in, uint4* __restrict__ out, int N, int C, long long S) {
  const int Cb = (C + 7) / 8;
  const long long total = (long long)N * Cb * S;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long s = i % S;
    const long long r = i / S;
    const int cb = (int)(r % Cb);
    const long long n = r / Cb;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cb * 8 + k;
      f[k] = (c < C) ? in[(n * C + c) * S + s] : 0.f;
    }
    out[i] = pack8(f);
  }
}

__global__ void cb8_to_planar_kernel(const uint4* __restrict__ in, float* __restrict__ out, int N, int C, long long S) {
  const int Cb = (C + 7) / 8;
  const long long total = (long long)N * Cb * S;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long s = i % S;
    const long long r = i / S;
    const int cb = (int)(r % Cb);
    const long long n = r / Cb;
    float f[8];
    unpack8(in[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cb * 8 + k;
      if (c < C) out[(n * C + c) * S + s] = f[k];
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

int bcp_mask_mix(const float* a, const float* b, float* out, int n, int c, int X, int Y, int Z,
                 const int* box6_dev, cudaStream_t stream) {
  BCP_REQUIRE(a && b && out && box6_dev, "mask_mix: null pointer");
  BCP_REQUIRE(n > 0 && c > 0 && X > 0 && Y > 0 && Z > 0, "mask_mix: bad shape");
  const long long total = (long long)n * c * X * Y * Z;
  const bool aligned = ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0;
  if (Z % 4 == 0 && aligned && total < (1ll << 31))
    mask_mix_vec4_kernel<<<grid_for(total / 4, 256), 256, 0, stream>>>((const float4*)a, (const float4*)b, (float4*)out,
                                                                         (unsigned)(total / 4), X, Y, Z / 4, Z, box6_dev);
  else
    mask_mix_kernel<<<grid_for(total, 256), 256, 0, stream>>>(a, b, out, total, X, Y, Z, box6_dev);
  return check_launch("mask_mix");
}

int bcp_label_mix(const unsigned char* a, const unsigned char* b, unsigned char* out, int n, int X, int Y, int Z,
                  const int* box6_dev, cudaStream_t stream) {
  BCP_REQUIRE(a && b && out && box6_dev && n > 0 && X > 0 && Y > 0 && Z > 0, "label_mix: bad args");
  const long long total = (long long)n * X * Y * Z;
  const bool aligned = ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0;
  if (Z % 16 == 0 && aligned && total < (1ll << 31))
    label_mix_vec16_kernel<<<grid_for(total / 16, 256), 256, 0, stream>>>((const uint4*)a, (const uint4*)b, (uint4*)out,
                                                                            (unsigned)(total / 16), X, Y, Z / 16, Z, box6_dev);
  else
    label_mix_kernel<<<grid_for(total, 256), 256, 0, stream>>>(a, b, out, total, X, Y, Z, box6_dev);
  return check_launch("label_mix");
}

int bcp_pseudo_label(const float* logits, unsigned char* out, int n, int c, long long v, int mode, float thr,
                     cudaStream_t stream) {
  BCP_REQUIRE(logits && out, "pseudo_label: null pointer");
  BCP_REQUIRE(c == 2 || c == 4, "pseudo_label: class count %d unsupported (2 or 4)", c);
  BCP_REQUIRE(mode == 0 || mode == 1, "pseudo_label: mode");
  BCP_REQUIRE(!(mode == 0 && c != 2), "pseudo_label: threshold mode takes channel 1 of 2");
  const long long total = (long long)n * v;
  const bool aligned = ((((uintptr_t)logits) & 15) | (((uintptr_t)out) & 3)) == 0;
  if (v % 4 == 0 && aligned && total * c < (1ll << 31)) {
    const int grid = grid_for(total / 4, 256);
    if (c == 2) pseudo_label_vec4_kernel<2><<<grid, 256, 0, stream>>>((const float4*)logits, (uchar4*)out, n, (unsigned)(v / 4), mode, thr);
    else pseudo_label_vec4_kernel<4><<<grid, 256, 0, stream>>>((const float4*)logits, (uchar4*)out, n, (unsigned)(v / 4), mode, thr);
  } else if (c == 2)
    pseudo_label_kernel<2><<<grid_for(total, 256), 256, 0, stream>>>(logits, out, n, v, mode, thr);
  else
    pseudo_label_kernel<4><<<grid_for(total, 256), 256, 0, stream>>>(logits, out, n, v, mode, thr);
  return check_launch("pseudo_label");
}

int bcp_sgd_ema_step(float* params, const float* grads, float* momentum, float* ema, const float* hyper,
                     long long n_train, long long n_total, cudaStream_t stream) {
  BCP_REQUIRE(params && grads && momentum && hyper, "sgd_ema_step: null pointer");
  BCP_REQUIRE(n_train >= 0 && n_total >= n_train, "sgd_ema_step: bad sizes");
  if (n_total == 0) return BCP_OK;
  BCP_REQUIRE(((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)momentum) | ((uintptr_t)ema)) & 15) == 0,
              "sgd_ema_step: arenas must be 16-byte aligned");
  sgd_ema_kernel<<<grid_for((n_total + 3) / 4, 256), 256, 0, stream>>>(params, grads, momentum, ema, hyper, n_train, n_total);
  return check_launch("sgd_ema_step");
}

int bcp_adam_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema,
                      const float* hyper, long long n_train, long long n_total, cudaStream_t stream) {
  BCP_REQUIRE(params && grads && exp_avg && exp_avg_sq && hyper, "adam_ema_step: null pointer");
  BCP_REQUIRE(n_train >= 0 && n_total >= n_train, "adam_ema_step: bad sizes");
  if (n_total == 0) return BCP_OK;
  adam_ema_kernel<<<grid_for(n_total, 256), 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, ema, hyper, n_train, n_total);
  return check_launch("adam_ema_step");
}

int bcp_adam_tick(float* hyper, long long* step, cudaStream_t stream) {
  BCP_REQUIRE(hyper && step, "adam_tick: null pointer");
  adam_tick_kernel<<<1, 32, 0, stream>>>(hyper, step);
  return check_launch("adam_tick");
}

int bcp_ema_i64(long long* ema, const long long* model, int n, float alpha, float one_minus_alpha, cudaStream_t stream) {
  BCP_REQUIRE(ema && model && n >= 0, "ema_i64: bad args");
  if (n == 0) return BCP_OK;
  ema_i64_kernel<<<(n + 127) / 128, 128, 0, stream>>>(ema, model, n, alpha, one_minus_alpha);
  return check_launch("ema_i64");
}

int bcp_weights_repack(const float* arena, void* packed, const bcp_repack_job* jobs_dev, int njobs,
                       cudaStream_t stream) {
  BCP_REQUIRE(arena && packed && jobs_dev && njobs > 0, "weights_repack: bad args");
  // x covers the 8x8 tiles of the largest layers (256x256 -> 1024 tiles) in a few iterations; blocks beyond a small
  // layer's tile count exit at once
  dim3 grid(256, njobs < 512 ? njobs : 512);
  repack_kernel<<<grid, 128, 0, stream>>>(arena, (__nv_bfloat16*)packed, jobs_dev, njobs);
  return check_launch("weights_repack");
}

int bcp_planar_to_cb8(const float* in, void* out, int n, int c, long long s, cudaStream_t stream) {
  BCP_REQUIRE(in && out && n > 0 && c > 0 && s > 0, "planar_to_cb8: bad args");
  const long long total = (long long)n * ((c + 7) / 8) * s;
  planar_to_cb8_kernel<<<grid_for(total, 256), 256, 0, stream>>>(in, (uint4*)out, n, c, s);
  return check_launch("planar_to_cb8");
}

int bcp_cb8_to_planar(const void* in, float* out, int n, int c, long long s, cudaStream_t stream) {
  BCP_REQUIRE(in && out && n > 0 && c > 0 && s > 0, "cb8_to_planar: bad args");
  const long long total = (long long)n * ((c + 7) / 8) * s;
  cb8_to_planar_kernel<<<grid_for(total, 256), 256, 0, stream>>>((const uint4*)in, out, n, c, s);
  return check_launch("cb8_to_planar");
}

}  // extern "C"
