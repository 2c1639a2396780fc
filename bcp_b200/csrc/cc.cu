// Largest-connected-component filter on the GPU (replaces the reference's per-sample
// GPU->CPU->GPU round trip through skimage.measure.label: LA_BCP_train.py:65-77,
// pancreas/pancreas_utils.py:284-296, ACDC_BCP_train.py:89-109).
//   * two-phase union-find: tile-local labelling in shared memory, then unions across tile borders only
//     (CAS hooks of the larger root under the smaller: roots = smallest raster index of a component)
//   * integer histogram of component sizes (atomicAdd on ints: order-independent result)
//   * per (sample, class) arg-max with ties -> smallest root == first component in raster order,
//     which is what np.argmax(np.bincount(labels.flat)[1:]) + 1 picks from skimage's raster labelling
//   * classes 1..3 are handled in one pass (unions only between equal non-zero labels)
// connectivity = max number of axes along which neighbours may differ (skimage semantics):
// 3-D: 1 -> 6, 2 -> 18, 3 -> 26 neighbours; 2-D (X == 1): 1 -> 4, 2 -> 8.
#include "common.cuh"
#include "../../include/bcp_b200.h"

namespace bcp {

// Union-find in the style of ECL-CC (Jaiganesh & Burtscher): find with intermediate pointer jumping, hooking by
// atomicCAS on ROOTS only (larger root under the smaller), so parent links always decrease and are never lost.
__device__ __forceinline__ int uf_find(volatile int* L, int v) {
  int curr = L[v];
  if (curr != v) {
    int prev = v, next;
    while (curr > (next = L[curr])) {
      L[prev] = next;          // pointer jumping; racy by design, values only move towards the root
      prev = curr;
      curr = next;
    }
  }
  return curr;
}

// read-only root lookup for the flatten pass: with no pointer-jumping writes in flight, the only stores are final
// roots, so every voxel ends up labelled with its root (a jumping write could otherwise overwrite a final label
// with a non-root ancestor).
__device__ __forceinline__ int uf_root(const volatile int* L, int v) {
  int curr = L[v], next;
  while (curr > (next = L[curr])) curr = next;
  return curr;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  int ra = uf_find(L, a), rb = uf_find(L, b);
  while (ra != rb) {
    if (ra < rb) { const int t = ra; ra = rb; rb = t; }       // ra > rb: hook ra under rb if ra is still a root
    const int seen = atomicCAS(&L[ra], ra, rb);
    if (seen == ra) return;
    ra = seen;                                                // somebody hooked ra meanwhile: continue from its parent
  }
}

// ---- phase 1: tile-local labelling in shared memory (tile = 8 x 8 x 32 voxels, CC_THREADS threads, 4 voxels each).  Local
// indices are monotone in the global raster order, so a local root is the first voxel of its local component; it is
// written out as a GLOBAL index, which makes phase 2 a plain continuation of the same union-find forest.
// The tile shape sets the share of voxels that still need global-memory unions in phase 2 (any voxel on a low-x,
// y or z face): 4x4x32 left 65 % of them, 8x8x32 leaves 38 %.
constexpr int TX = 8, TY = 8, TZ = 32, TV = TX * TY * TZ, CC_THREADS = 512;

__device__ __forceinline__ int sfind(volatile int* L, int v) {
  int curr = L[v];
  if (curr != v) {
    int prev = v, next;
    while (curr > (next = L[curr])) { L[prev] = next; prev = curr; curr = next; }
  }
  return curr;
}
__device__ __forceinline__ void sunion(int* L, int a, int b) {
  int ra = sfind(L, a), rb = sfind(L, b);
  while (ra != rb) {
    if (ra < rb) { const int t = ra; ra = rb; rb = t; }
    const int seen = atomicCAS(&L[ra], ra, rb);
    if (seen == ra) return;
    ra = seen;
  }
}

__global__ void __launch_bounds__(CC_THREADS) cc_local_kernel(const unsigned char* __restrict__ seg, int* __restrict__ L,
                                                              int* __restrict__ cnt, unsigned long long* __restrict__ best,
                                                              int N, int X, int Y, int Z, int conn, int nbest) {
  __shared__ int Ls[TV];
  __shared__ unsigned char Ss[TV];
  const int tilesz = (Z + TZ - 1) / TZ, tilesy = (Y + TY - 1) / TY, tilesx = (X + TX - 1) / TX;
  int b = blockIdx.x;
  const int tz = b % tilesz; b /= tilesz;
  const int ty = b % tilesy; b /= tilesy;
  const int tx = b % tilesx;
  const int n = b / tilesx;
  const long long V = (long long)X * Y * Z;
  if (blockIdx.x == 0 && threadIdx.x < nbest) best[threadIdx.x] = 0ull;
  for (int l = threadIdx.x; l < TV; l += CC_THREADS) {
    const int lz = l % TZ, ly = (l / TZ) % TY, lx = l / (TZ * TY);
    const int x = tx * TX + lx, y = ty * TY + ly, z = tz * TZ + lz;
    const bool inside = x < X && y < Y && z < Z;
    Ss[l] = inside ? seg[(long long)n * V + ((long long)x * Y + y) * Z + z] : 0;
    Ls[l] = l;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < TV; l += CC_THREADS) {
    const unsigned char c = Ss[l];
    if (!c) continue;
    const int lz = l % TZ, ly = (l / TZ) % TY, lx = l / (TZ * TY);
    for (int dx = -1; dx <= 0; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) {
          const int off = (dx * TY + dy) * TZ + dz;
          if (off >= 0) continue;
          if (abs(dx) + abs(dy) + abs(dz) > conn) continue;
          const int xx = lx + dx, yy = ly + dy, zz = lz + dz;
          if (xx < 0 || yy < 0 || yy >= TY || zz < 0 || zz >= TZ) continue;   // other tile: phase 2
          if (Ss[l + off] == c) sunion(Ls, l, l + off);
        }
  }
  __syncthreads();
  for (int l = threadIdx.x; l < TV; l += CC_THREADS) {
    const int lz = l % TZ, ly = (l / TZ) % TY, lx = l / (TZ * TY);
    const int x = tx * TX + lx, y = ty * TY + ly, z = tz * TZ + lz;
    if (!(x < X && y < Y && z < Z)) continue;
    const long long gi = (long long)n * V + ((long long)x * Y + y) * Z + z;
    int out = -1;
    if (Ss[l]) {
      int r = Ls[l], nx;
      while (r > (nx = Ls[r])) r = nx;                                          // read-only root lookup
      const int rz = r % TZ, ry = (r / TZ) % TY, rx = r / (TZ * TY);
      out = ((tx * TX + rx) * Y + (ty * TY + ry)) * Z + (tz * TZ + rz);
    }
    L[gi] = out;
    cnt[gi] = 0;
  }
}

// ---- phase 2: unions across tile borders only (global forest)
__global__ void cc_border_kernel(const unsigned char* __restrict__ seg, int* __restrict__ L, int N, int X, int Y, int Z, int conn) {
  const int V = X * Y * Z;
  const long long total = (long long)N * V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += stride) {
    const unsigned char c = seg[gi];
    if (!c) continue;
    const int n = (int)(gi / V), i = (int)(gi - (long long)n * V);
    const int z = i % Z, y = (i / Z) % Y, x = i / (Z * Y);
    const int lx = x % TX, ly = y % TY, lz = z % TZ;
    if (lx != 0 && ly != 0 && ly != TY - 1 && lz != 0 && lz != TZ - 1) continue;  // interior: no backward neighbour leaves the tile
    int* Ls = L + (long long)n * V;
    const unsigned char* ss = seg + (long long)n * V;
    for (int dx = -1; dx <= 0; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) {
          const int off = (dx * Y + dy) * Z + dz;
          if (off >= 0) continue;
          if (abs(dx) + abs(dy) + abs(dz) > conn) continue;
          const int xx = x + dx, yy = y + dy, zz = z + dz;
          if (xx < 0 || yy < 0 || yy >= Y || zz < 0 || zz >= Z) continue;
          const int tlx = lx + dx, tly = ly + dy, tlz = lz + dz;
          if (tlx >= 0 && tly >= 0 && tly < TY && tlz >= 0 && tlz < TZ) continue;   // same tile: done in phase 1
          const int j = i + off;
          if (ss[j] == c) uf_union(Ls, i, j);
        }
  }
}

__global__ void cc_count_kernel(const unsigned char* __restrict__ seg, int* __restrict__ L, int* __restrict__ cnt, int N, int V) {
  const long long total = (long long)N * V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // the loop bound is rounded up to a whole warp so every lane takes part in the match; lanes of one warp that share a
  // root (the common case inside a blob) issue ONE atomicAdd of their population count
  const long long total_w = (total + 31) / 32 * 32;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total_w; gi += stride) {
    long long key = -1;
    if (gi < total && seg[gi]) {
      const int n = (int)(gi / V), i = (int)(gi - (long long)n * V);
      int* Ls = L + (long long)n * V;
      const int r = uf_root(Ls, i);
      Ls[i] = r;
      key = (long long)n * V + r;
    }
    const unsigned m = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&cnt[key], __popc(m));
  }
}

__global__ void cc_best_kernel(const unsigned char* __restrict__ seg, const int* __restrict__ L, const int* __restrict__ cnt,
                               unsigned long long* __restrict__ best, int N, int V) {
  const long long total = (long long)N * V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += stride) {
    const unsigned char c = seg[gi];
    if (!c) continue;
    const int n = (int)(gi / V), i = (int)(gi - (long long)n * V);
    if (L[gi] != i) continue;   // roots only
    const unsigned long long key = ((unsigned long long)(unsigned)cnt[gi] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    atomicMax(&best[n * 4 + (c & 3)], key);
  }
}

__global__ void cc_select_kernel(const unsigned char* __restrict__ seg, const int* __restrict__ L,
                                 const unsigned long long* __restrict__ best, unsigned char* __restrict__ out_u8,
                                 float* __restrict__ out_f32, int N, int V) {
  const long long total = (long long)N * V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += stride) {
    const unsigned char c = seg[gi];
    unsigned char o = 0;
    if (c) {
      const int n = (int)(gi / V);
      const unsigned long long key = best[n * 4 + (c & 3)];
      const int root = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
      o = (L[gi] == root) ? c : 0;
    }
    if (out_u8) out_u8[gi] = o;
    if (out_f32) out_f32[gi] = (float)o;
  }
}

}  // namespace bcp

using namespace bcp;

extern "C" {

long long bcp_largest_cc_workspace_bytes(int n, long long v) { return (long long)n * v * 8 + (long long)n * 4 * 8 + 64; }

int bcp_largest_cc(const unsigned char* seg, unsigned char* out_u8, float* out_f32, void* workspace, int n, int X, int Y, int Z,
                   int connectivity, cudaStream_t stream) {
  BCP_REQUIRE(seg && workspace && (out_u8 || out_f32), "largest_cc: null pointer");
  BCP_REQUIRE(n > 0 && X > 0 && Y > 0 && Z > 0, "largest_cc: bad shape");
  BCP_REQUIRE((long long)X * Y * Z < (1ll << 31), "largest_cc: volume too large");
  BCP_REQUIRE(connectivity >= 1 && connectivity <= 3, "largest_cc: connectivity %d", connectivity);
  const int V = X * Y * Z;
  const long long total = (long long)n * V;
  unsigned long long* best = (unsigned long long*)workspace;          // [n][4], 8-byte aligned at the front
  int* L = (int*)((char*)workspace + (((long long)n * 4 * 8 + 63) / 64) * 64);
  int* cnt = L + total;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  const int g = (int)blocks;
  const int tiles = n * ((X + TX - 1) / TX) * ((Y + TY - 1) / TY) * ((Z + TZ - 1) / TZ);
  BCP_REQUIRE(n * 4 <= CC_THREADS, "largest_cc: batch too large");
  cc_local_kernel<<<tiles, CC_THREADS, 0, stream>>>(seg, L, cnt, best, n, X, Y, Z, connectivity, n * 4);
  cc_border_kernel<<<g, 256, 0, stream>>>(seg, L, n, X, Y, Z, connectivity);
  cc_count_kernel<<<g, 256, 0, stream>>>(seg, L, cnt, n, V);
  cc_best_kernel<<<g, 256, 0, stream>>>(seg, L, cnt, best, n, V);
  cc_select_kernel<<<g, 256, 0, stream>>>(seg, L, best, out_u8, out_f32, n, V);
  return check_launch("largest_cc");
}

}  // extern "C"
