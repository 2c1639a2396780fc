// One TMA box load of an fp32 tensor with a given (dims, box, start): does the descriptor + instruction run, and does the
// tile (with out-of-bounds zero fill) arrive as expected?   usage: tma_probe rank gz gy gx gn  bz by bx bn  cz cy cx cn [u32]
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/micro/tma_probe_bin tools/micro/tma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int nfloats, int rank, int c0, int c1, int c2, int c3, unsigned bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    if (rank == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(d), "l"(&map), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(d), "l"(&map), "r"(b), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}

int main(int argc, char** argv) {
  if (argc < 14) { printf("usage\n"); return 2; }
  const int rank = atoi(argv[1]);
  long long g[4]; int bx[4], c[4];
  for (int i = 0; i < 4; ++i) { g[i] = atoll(argv[2 + i]); bx[i] = atoi(argv[6 + i]); c[i] = atoi(argv[10 + i]); }
  const bool u32 = argc > 14;
  const long long total = g[0] * g[1] * g[2] * g[3];
  std::vector<float> h(total);
  for (long long i = 0; i < total; ++i) h[i] = (float)(i + 1);
  float *dx, *dout;
  cudaMalloc(&dx, total * 4);
  cudaMemcpy(dx, h.data(), total * 4, cudaMemcpyHostToDevice);
  const int nf = bx[0] * bx[1] * bx[2] * bx[3];
  cudaMalloc(&dout, nf * 4);
  cudaMemset(dout, 0xff, nf * 4);
  CUtensorMap map;
  cuuint64_t gdim[4] = {(cuuint64_t)g[0], (cuuint64_t)g[1], (cuuint64_t)g[2], (cuuint64_t)g[3]};
  cuuint64_t gstr[3] = {(cuuint64_t)g[0] * 4, (cuuint64_t)g[0] * g[1] * 4, (cuuint64_t)g[0] * g[1] * g[2] * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx[0], (cuuint32_t)bx[1], (cuuint32_t)bx[2], (cuuint32_t)bx[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  cuInit(0);
  CUresult cr = cuTensorMapEncodeTiled(&map, u32 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, dx, gdim, gstr, box, es,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("rank %d dims %lld %lld %lld %lld box %d %d %d %d at %d %d %d %d %s: encode rc %d", rank, g[0], g[1], g[2], g[3], bx[0], bx[1], bx[2], bx[3],
         c[0], c[1], c[2], c[3], u32 ? "u32" : "f32", (int)cr);
  if (cr != CUDA_SUCCESS) { printf("\n"); return 1; }
  int nb = nf; if (rank == 3) nb = bx[0] * bx[1] * bx[2];
  probe<<<1, 128, nb * 4 + 256>>>(map, dout, nb, rank, c[0], c[1], c[2], c[3], (unsigned)nb * 4);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  run: %s", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> o(nb);
    cudaMemcpy(o.data(), dout, nb * 4, cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (int i = 0; i < nb; ++i) {
      int r = i;
      const int iz = r % bx[0]; r /= bx[0];
      const int iy = r % bx[1]; r /= bx[1];
      const int ix = r % bx[2]; r /= bx[2];
      const int in = r;
      const long long z = c[0] + iz, y = c[1] + iy, x = c[2] + ix, n = (rank == 4 ? c[3] : 0) + in;
      float want = 0.f;
      if (z >= 0 && z < g[0] && y >= 0 && y < g[1] && x >= 0 && x < g[2] && n >= 0 && n < g[3]) want = h[((n * g[2] + x) * g[1] + y) * g[0] + z];
      if (o[i] != want) ++bad;
    }
    printf("  mismatches %lld / %d", bad, nb);
  }
  printf("\n");
  return 0;
}
