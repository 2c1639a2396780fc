// Phase timing of the cluster-fused normalisation kernel against the two-launch path (csrc/norm.cu) on one layer shape.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DBCP_NORM_PROBE -o /tmp/norm_probe \
//        tools/micro/norm_probe.cu bcp_b200/csrc/norm_fused.cu bcp_b200/csrc/norm.cu bcp_b200/csrc/api.cu && /tmp/norm_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../include/bcp_b200.h"

extern "C" int bcp_norm_probe_read(unsigned long long* host16);
extern "C" int bcp_norm_probe_max_clusters(int cs);

static void run(int n, int c, long long s, int spg) {
  const long long elems = (long long)n * c * s;
  std::vector<__nv_bfloat16> h(elems);
  for (long long i = 0; i < elems; ++i) h[i] = __float2bfloat16((float)((i * 2654435761u) % 1000) / 500.f - 0.7f);
  void *y, *out, *da, *dy;
  cudaMalloc(&y, elems * 2); cudaMalloc(&out, elems * 2); cudaMalloc(&da, elems * 2); cudaMalloc(&dy, elems * 2);
  cudaMemcpy(y, h.data(), elems * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(da, h.data(), elems * 2, cudaMemcpyHostToDevice);
  const int G = n / spg;
  float *gamma, *beta, *rm, *rv, *stat, *coef, *ws, *sums, *dg, *db;
  long long* nbt; int* counter;
  cudaMalloc(&gamma, c * 4); cudaMalloc(&beta, c * 4); cudaMalloc(&rm, c * 4); cudaMalloc(&rv, c * 4); cudaMalloc(&dg, c * 4); cudaMalloc(&db, c * 4);
  cudaMemset(gamma, 0, c * 4); cudaMemset(beta, 0, c * 4); cudaMemset(rm, 0, c * 4); cudaMemset(rv, 0, c * 4);
  cudaMalloc(&stat, G * c * 8); cudaMalloc(&coef, G * c * 8); cudaMalloc(&sums, G * c * 8);
  const long long wsf = bcp_norm_workspace_floats(n, c, s) + G * c * 2 + 64;
  cudaMalloc(&ws, wsf * 4); cudaMalloc(&nbt, 8); cudaMalloc(&counter, 64); cudaMemset(counter, 0, 64); cudaMemset(nbt, 0, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int mode = 0; mode < 4; ++mode) {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(e0);
      int rc = 0;
      if (mode == 0) rc = bcp_norm_fused_fwd(y, out, gamma, beta, rm, rv, nbt, stat, coef, ws, counter, nullptr, nullptr, 1.f, nullptr, n, c, s, spg, 1e-5f, 0.1f, 0.f, 0);
      if (mode == 1) { rc = bcp_norm_stats(y, gamma, beta, rm, rv, nbt, stat, coef, ws, counter + 4, n, c, s, spg, 1e-5f, 0.1f, 0);
                       rc |= bcp_norm_apply(y, out, coef, nullptr, nullptr, 1.f, nullptr, n, c, s, spg, 0.f, 0); }
      if (mode == 2) rc = bcp_norm_fused_bwd(da, y, dy, stat, coef, nullptr, nullptr, 1.f, dg, db, sums, counter + 8, n, c, s, spg, 0.f, 1, 0, 0);
      if (mode == 3) rc = bcp_norm_bwd(da, y, dy, stat, coef, nullptr, nullptr, 1.f, dg, db, sums, ws, counter + 4, n, c, s, spg, 0.f, 1, 0, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      if (rc) { printf("rc=%d %s\n", rc, bcp_last_error()); return; }
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const char* names[4] = {"fused fwd", "stats+apply", "fused bwd", "reduce+apply"};
    printf("n=%d c=%d s=%lld spg=%d  %-13s %7.1f us", n, c, s, spg, names[mode], best * 1e3f);
    if (mode == 0) {
      unsigned long long p[16];
      bcp_norm_probe_read(p);
      printf("   phases(ns): load+acc %llu  block_sum %llu  sync1 %llu  stats %llu  sync2 %llu  apply %llu  tail %llu", p[1] - p[0], p[2] - p[1],
             p[3] - p[2], p[4] - p[3], p[5] - p[4], p[6] - p[5], p[7] - p[6]);
    }
    printf("\n");
  }
  cudaFree(y); cudaFree(out); cudaFree(da); cudaFree(dy);
}

int main() {
  for (int cs = 1; cs <= 8; cs *= 2) printf("max co-resident clusters of %d CTAs x 1024 threads: %d\n", cs, bcp_norm_probe_max_clusters(cs));
  run(4, 64, 28 * 28 * 20, 2);
  run(4, 128, 14 * 14 * 10, 2);
  run(4, 256, 7 * 7 * 5, 2);
  run(4, 32, 56 * 56 * 40, 2);
  run(12, 32, 128 * 128, 6);
  run(2, 64, 24 * 24 * 24, 1);
  return 0;
}
