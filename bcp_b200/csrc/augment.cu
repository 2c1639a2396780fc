// Device-side input pipeline for the LA / Pancreas volumes: the training volumes stay resident
// in HBM and every step's patches are cut out of them by ONE gather kernel per sample that composes the reference's numpy
// transforms -- RandomRotFlip (np.rot90 by k quarter turns in the (x, y) plane, then np.flip along axis 0 or 1), the
// zero padding RandomCrop applies to volumes smaller than the patch, and the random crop itself
// (dataloaders/dataset.py:52-60,173-225) -- and ToTensor (:267-277: image -> fp32 [1,X,Y,Z]; the label stays uint8 here).
// Pure data movement: bit-exact against the numpy pipeline.  HBM-bound: reads <= one patch of fp32 + uint8, writes the same.
#include "common.cuh"
#include "../../include/bcp_b200.h"

namespace bcp {

struct AugArgs {
  int W, H, D;          // stored volume
  int OX, OY, OZ;       // patch
  int k, axis;          // quarter turns (0..3), flip axis (0 or 1) applied after the rotation
  int pw, ph, pd;       // zero padding per side (after rotation/flip)
  int w1, h1, d1;       // crop origin in the padded frame
};

__global__ void __launch_bounds__(256) aug_crop_rotflip_kernel(const float* __restrict__ img, const unsigned char* __restrict__ lab,
                                                               float* __restrict__ out_img, unsigned char* __restrict__ out_lab,
                                                               AugArgs a) {
  // dims of the rotated (and flipped) volume
  const int RW = (a.k & 1) ? a.H : a.W, RH = (a.k & 1) ? a.W : a.H;
  const long long total = (long long)a.OX * a.OY * a.OZ;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
    const int z = (int)(o % a.OZ);
    const long long r = o / a.OZ;
    const int y = (int)(r % a.OY), x = (int)(r / a.OY);
    int i = x + a.w1 - a.pw, j = y + a.h1 - a.ph;
    const int d = z + a.d1 - a.pd;
    float v = 0.f;
    unsigned char l = 0;
    if (i >= 0 && i < RW && j >= 0 && j < RH && d >= 0 && d < a.D) {
      if (a.axis == 0) i = RW - 1 - i; else j = RH - 1 - j;           // undo np.flip
      int si, sj;                                                    // undo np.rot90(m, k): r[i, j] = m[si, sj]
      switch (a.k) {
        case 0: si = i; sj = j; break;
        case 1: si = j; sj = a.H - 1 - i; break;
        case 2: si = a.W - 1 - i; sj = a.H - 1 - j; break;
        default: si = a.W - 1 - j; sj = i; break;
      }
      const long long s = ((long long)si * a.H + sj) * a.D + d;
      v = __ldg(img + s);
      l = __ldg(lab + s);
    }
    out_img[o] = v;
    out_lab[o] = l;
  }
}

}  // namespace bcp

using namespace bcp;

extern "C" {

int bcp_aug_crop_rotflip(const float* img, const unsigned char* lab, float* out_img, unsigned char* out_lab, const int* src_dims,
                         const int* out_dims, int k, int flip_axis, const int* pad, const int* origin, cudaStream_t stream) {
  BCP_REQUIRE(img && lab && out_img && out_lab && src_dims && out_dims && pad && origin, "aug_crop_rotflip: null pointer");
  BCP_REQUIRE(k >= 0 && k < 4 && (flip_axis == 0 || flip_axis == 1), "aug_crop_rotflip: k in 0..3, flip axis 0 or 1");
  AugArgs a{src_dims[0], src_dims[1], src_dims[2], out_dims[0], out_dims[1], out_dims[2], k, flip_axis,
            pad[0], pad[1], pad[2], origin[0], origin[1], origin[2]};
  BCP_REQUIRE(a.W > 0 && a.H > 0 && a.D > 0 && a.OX > 0 && a.OY > 0 && a.OZ > 0, "aug_crop_rotflip: bad dims");
  const long long total = (long long)a.OX * a.OY * a.OZ;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  aug_crop_rotflip_kernel<<<(unsigned)blocks, 256, 0, stream>>>(img, lab, out_img, out_lab, a);
  return check_launch("aug_crop_rotflip");
}

}  // extern "C"
