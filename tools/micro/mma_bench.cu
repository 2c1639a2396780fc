// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, no-swizzle K-major operands in shared memory) as a
// function of N, the number of independent accumulators interleaved (MT), and the alignment of the A start row.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mma_bench tools/micro/mma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct Cfg { int N, MT, step_rows, reps, lbo_rows, mshift, same_b, tile_rows; };

__global__ void __launch_bounds__(64, 1) bench(Cfg c, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned long long bar;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 128 * 1024;
    const uint64_t adesc0 = make_desc(a_base, (uint32_t)c.lbo_rows * 16u, 128u);
    const uint64_t bdesc0 = make_desc(b_base, (uint32_t)c.N * 16u, 128u);
    const bool leader = elect_one();
    long long t0 = clock64();
    if (leader) {
      int row = c.mshift;
      for (int i = 0; i < c.reps; ++i) {
        uint32_t d = tmem;
        for (int mt = 0; mt < c.MT; ++mt) {
          umma(d, adesc0 + (uint64_t)(row + mt * c.tile_rows), bdesc0 + (c.same_b ? 0 : (uint64_t)((i % 9) * c.N * 2)), idesc, 1u);
          d += (uint32_t)c.N;
        }
        row += c.step_rows;
        if (row > 1000) row = c.mshift;
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    } while (!done);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// Lean issue structure: one elected thread, hi descriptor words constant, 9 taps unrolled with immediate row offsets,
// runtime loops over (group, mt).  Emulates the production loop nest: chunk -> 3 tap groups -> MT tiles -> 9 taps.
__global__ void __launch_bounds__(64, 1) bench_lean(Cfg c, long long* out, int HZ) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned long long bar;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 128 * 1024;
    const uint64_t adesc0 = make_desc(a_base, (uint32_t)c.lbo_rows * 16u, 128u);
    const uint64_t bdesc0 = make_desc(b_base, (uint32_t)c.N * 16u, 128u);
    const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
    const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
    const uint32_t bstep = (uint32_t)c.N * 2;     // 16-byte units per tap
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < c.reps; ++i) {
        for (int g = 0; g < 3; ++g) {
          const uint32_t a_g = a_lo0 + (uint32_t)(g * 9 * HZ);
          uint32_t d = tmem;
          for (int mt = 0; mt < c.MT; ++mt) {
            const uint32_t a_m = a_g + (uint32_t)(mt * 128);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint32_t a_lo = a_m + (uint32_t)((t / 3) * HZ + (t % 3));
              const uint32_t b_lo = b_lo0 + (uint32_t)t * bstep;
              umma(d, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, (i | g | t) ? 1u : 0u);
            }
            d += (uint32_t)c.N;
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    } while (!done);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  cudaFuncSetAttribute(bench_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("LEAN grid N MT HZ cycles_per_mma\n");
  for (int grid : {1, 148})
    for (int N : Ns)
      for (int MT : {1, 2, 4, 8}) {
        if (MT * N > 512) continue;
        for (int HZ : {8, 7}) {
          Cfg c{N, MT, 0, 200, 2048, 0, 0, 128};
          bench_lean<<<grid, 64, 220 * 1024>>>(c, d_out, HZ);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, d_out, grid * 8, cudaMemcpyDeviceToHost);
          double avg = 0;
          for (int i = 0; i < grid; ++i) avg += (double)h[i];
          avg /= grid;
          printf("LEAN %d %d %d %d %.1f\n", grid, N, MT, HZ, avg / (c.reps * 27.0 * MT));
        }
      }
  return 0;
  printf("grid N MT step_rows lbo_rows mshift same_b cycles_per_mma\n");
  for (int grid : {1, 148})
    for (int N : Ns)
      for (int MT : {1, 2, 4, 8}) {
        if (MT * N > 512) continue;
        for (int variant = 0; variant < 4; ++variant) {
          // 0: aligned rows, moving; 1: unaligned (+1 row per step); 2: fixed address; 3: aligned, MT tiles overlap (tile stride 8 rows)
          Cfg c{N, MT, variant == 1 ? 1 : (variant == 2 ? 0 : 8), 2000 / MT, 2048, variant == 1 ? 1 : 0, 0, variant == 3 ? 8 : 128};
          if (grid == 148 && variant >= 2) continue;
          bench<<<grid, 64, 220 * 1024>>>(c, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, d_out, grid * 8, cudaMemcpyDeviceToHost);
          double avg = 0;
          for (int i = 0; i < grid; ++i) avg += (double)h[i];
          avg /= grid;
          printf("%d %d %d %d %d %d v%d %.1f\n", grid, N, MT, c.step_rows, c.lbo_rows, c.mshift, variant, avg / (c.reps * MT));
        }
      }
  return 0;
}
