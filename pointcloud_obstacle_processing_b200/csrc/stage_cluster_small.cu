// Fused Euclidean clustering for small clouds: ONE thread block per frame, everything in shared memory.
//
// After ground removal the cloud that reaches pcl::EuclideanClusterExtraction (od.cpp:796) is a few thousand
// points (BASELINE configs[1]: ~5 k of 120 k).  At that size the generic path (stage_cluster.cu: ~30 launches,
// three global radix sorts, union-find in L2) is pure launch and latency overhead.  For frames with at most
// ES_MAX points this kernel runs the whole stage -- search grid, connected components, size filter, canonical
// ordering, CSR indices and the per-cluster centroid + bounding radius -- in one launch:
//
//   1. min/max of the finite points -> grid (same rules as the generic path, ece_common.cuh);
//   2. (cell key, point index) packed in 64 bits, sorted by a normalised bitonic network in shared memory
//      (index in the low word => stable, a cell's first point is its smallest original index);
//   3. occupied-cell list + an open-addressing hash table cell key -> cell (replaces the kd-tree / binary search);
//   4. union-find in shared memory, atomic-min hooking (roots = smallest original index, so labels are
//      deterministic).  Clique mode: one warp per cell, the 62 forward neighbour cells are looked up by the
//      lanes in parallel, point pairs are tested with the exact float predicate only until the first hit and
//      only while the two cells are still in different sets.  Point mode (extent too large for clique cells):
//      one thread per point scans its forward neighbour cells;
//   5. sizes, kept roots sorted by (size desc, smallest index asc), CSR offsets, members sorted by
//      (cluster rank, index) -> cluster_indices;
//   6. one warp per cluster: centroid (double sums) and max distance (od.cpp:457-464 arithmetic).
//
// The result is identical to the generic path (and the oracle): clusters are the connected components of the
// graph "float ((dx*dx)+(dy*dy))+(dz*dz) < r2", which does not depend on the acceleration structure.
#include "ece_common.cuh"
#include "internal.cuh"
#include "primitives.cuh"

namespace pcop {
#ifdef PCOP_ECE_DEBUG_CLK
__device__ long long g_ece_small_clk[16];
#endif

namespace {

constexpr int ES_THREADS = 1024;
constexpr int ES_WARPS = ES_THREADS / 32;
constexpr int ES_HASH_BITS = 14;
constexpr int ES_HASH = 1 << ES_HASH_BITS;
constexpr int ES_MAX = ECE_SMALL_MAX;
typedef unsigned long long u64;

// Shared-memory layout (bytes).  Regions are reused between phases:
//   R1  u64 sortbuf[ES_MAX]        sorts            | float x[ES_MAX], y[ES_MAX] during the union phase
//   R2  float z[ES_MAX]            union phase      | int csize[ES_MAX] afterwards
//   R3  int parent[ES_MAX]         union-find over SORTED positions, then labels
//   R4  u16 idx16[ES_MAX]          original index of each sorted position
//   R5  u16 cell_start[ES_MAX+8]   first sorted position of each occupied cell   \  afterwards: int minidx[ES_MAX]
//   R6  u16 hash[ES_HASH]          cell hash table (value = cell + 1, 0 = empty) /  then u16 rank_of[ES_MAX]
constexpr int ES_R1 = 0;
constexpr int ES_R2 = ES_R1 + 8 * ES_MAX;
constexpr int ES_R3 = ES_R2 + 4 * ES_MAX;
constexpr int ES_R4 = ES_R3 + 4 * ES_MAX;
constexpr int ES_R5 = ES_R4 + 2 * ES_MAX;
constexpr int ES_R6 = ES_R5 + 2 * (ES_MAX + 8);
constexpr int ES_MISC = ES_R6 + 2 * ES_HASH;
struct EceSmallMisc {
  float red[ES_WARPS][6];
  int wscan[ES_WARPS + 1];
  EceFrame ef;
  int n_cells, n_clusters, n_members;
};
constexpr int ES_SMEM_BYTES = ES_MISC + (int)sizeof(EceSmallMisc);
static_assert(ES_SMEM_BYTES <= 227 * 1024, "fused clustering kernel: shared memory budget");
static_assert(4 * ES_MAX <= 2 * (ES_MAX + 8) + 2 * ES_HASH, "minidx overlays cell_start + hash");
static_assert(ES_MAX <= 16383, "positions / indices are packed in 14 bits");
static_assert(ES_R5 % 4 == 0 && ES_R2 % 16 == 0 && ES_MISC % 8 == 0, "alignment");

__device__ __forceinline__ uint32_t hi32(u64 v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ uint32_t lo32(u64 v) { return (uint32_t)v; }

// Normalised bitonic network (every comparator puts the smaller element at the lower index), so elements past n
// behave as +infinity without being stored: a comparator whose upper end is >= n is a no-op.
__device__ void bitonic_sort_smem(u64* a, int n) {
  if (n <= 1) {
    __syncthreads();
    return;
  }
  int npow = 1;
  while (npow < n) npow <<= 1;
  const int half_n = npow >> 1;
  for (int k = 2; k <= npow; k <<= 1) {
    const int hk = k >> 1;
    for (int t = threadIdx.x; t < half_n; t += ES_THREADS) {
      const int off = t & (hk - 1);
      const int base = (t - off) << 1;
      const int i = base + off, p = base + k - 1 - off;
      if (p < n) {
        const u64 x = a[i], y = a[p];
        if (x > y) {
          a[i] = y;
          a[p] = x;
        }
      }
    }
    __syncthreads();
    for (int j = hk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < half_n; t += ES_THREADS) {
        const int off = t & (j - 1);
        const int i = ((t - off) << 1) + off, p = i + j;
        if (p < n) {
          const u64 x = a[i], y = a[p];
          if (x > y) {
            a[i] = y;
            a[p] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

// exclusive prefix of v over the block in thread order; total = block sum.  All threads must call.
__device__ __forceinline__ int block_excl_scan(int v, int* wscan, int& total) {
  const int lane = lane_id(), warp = warp_id();
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += up;
  }
  __syncthreads();  // wscan may still be read from a previous call
  if (lane == 31) wscan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = wscan[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(FULL, wi, o);
      if (lane >= o) wi += up;
    }
    wscan[lane] = wi - w;
    if (lane == 31) wscan[ES_WARPS] = wi;
  }
  __syncthreads();
  total = wscan[ES_WARPS];
  return wscan[warp] + incl - v;
}

__device__ __forceinline__ uint32_t hash_slot(uint32_t key) { return (key * 2654435761u) >> (32 - ES_HASH_BITS); }

struct EceSmallView {
  const float *x, *y, *z;
  const unsigned short *idx16, *cell_start, *hash;
  EceFrame e;
  __device__ __forceinline__ uint32_t cell_key(int c) const {  // recomputed from the cell's first point
    const int j = cell_start[c];
    return ece_point_key(make_float4(x[j], y[j], z[j], 0.f), e, (int)idx16[j]);
  }
  // cell index of `key` or -1
  __device__ __forceinline__ int lookup(uint32_t key) const {
    uint32_t s = hash_slot(key);
    while (true) {
      const unsigned v = hash[s];
      if (v == 0u) return -1;
      if (cell_key((int)v - 1) == key) return (int)v - 1;
      s = (s + 1) & (ES_HASH - 1);
    }
  }
};

// phase timestamps of block 0 (clock64), for tools/ece_phases.py; written only when PCOP_ECE_DEBUG_CLK is defined
#ifdef PCOP_ECE_DEBUG_CLK
#define ES_CLK(k)                                                      \
  do {                                                                 \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_ece_small_clk[k] = clock64(); \
  } while (0)
#else
#define ES_CLK(k) \
  do {            \
  } while (0)
#endif

__global__ void __launch_bounds__(ES_THREADS, 1)
    k_ece_small(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in, float tol, float r2,
                int min_size, int max_size, int small_max, int* __restrict__ offsets, int* __restrict__ indices,
                int* __restrict__ n_clusters, int* __restrict__ n_cluster_pts, float4* __restrict__ obstacles, int cap) {
  const int f = blockIdx.x;
  const int n = n_in[f];
  if (n > small_max) return;  // this frame takes the generic path
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* sortbuf = reinterpret_cast<u64*>(smem_raw + ES_R1);
  float* px = reinterpret_cast<float*>(smem_raw + ES_R1);
  float* py = px + ES_MAX;
  float* pz = reinterpret_cast<float*>(smem_raw + ES_R2);
  int* csize = reinterpret_cast<int*>(smem_raw + ES_R2);
  int* parent = reinterpret_cast<int*>(smem_raw + ES_R3);
  unsigned short* idx16 = reinterpret_cast<unsigned short*>(smem_raw + ES_R4);
  unsigned short* cell_start = reinterpret_cast<unsigned short*>(smem_raw + ES_R5);
  unsigned short* hash = reinterpret_cast<unsigned short*>(smem_raw + ES_R6);
  int* minidx = reinterpret_cast<int*>(smem_raw + ES_R5);
  unsigned short* rank_of = reinterpret_cast<unsigned short*>(smem_raw + ES_R5);
  EceSmallMisc& sm = *reinterpret_cast<EceSmallMisc*>(smem_raw + ES_MISC);

  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const float4* src = in + (size_t)f * in_stride;
  int* offs = offsets + (size_t)f * (cap + 1);
  int* idx_out = indices + (size_t)f * cap;
  float4* obs = obstacles + (size_t)f * cap;
  if (n <= 0) {
    if (tid == 0) {
      n_clusters[f] = 0;
      n_cluster_pts[f] = 0;
      offs[0] = 0;
    }
    return;
  }

  ES_CLK(0);
  // ---- 1. min/max over the finite points, grid ---------------------------------------------------------
  {
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (int i = tid; i < n; i += ES_THREADS) {
      const float4 p = __ldg(src + i);
      const float v[3] = {p.x, p.y, p.z};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (fabsf(v[a]) <= 3.0e38f) {  // finite only (NaN fails the compare)
          mn[a] = fminf(mn[a], v[a]);
          mx[a] = fmaxf(mx[a], v[a]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        sm.red[warp][a] = mn[a];
        sm.red[warp][3 + a] = mx[a];
      }
    }
    __syncthreads();
    if (tid == 0) {
      float gmn[3], gmx[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        gmn[a] = sm.red[0][a];
        gmx[a] = sm.red[0][3 + a];
        for (int w = 1; w < ES_WARPS; ++w) {
          gmn[a] = fminf(gmn[a], sm.red[w][a]);
          gmx[a] = fmaxf(gmx[a], sm.red[w][3 + a]);
        }
      }
      sm.ef = ece_make_frame(gmn, gmx, n, tol, /*clique=*/1);
    }
    __syncthreads();
  }
  const EceFrame e = sm.ef;

  ES_CLK(1);
  // ---- 2. sort (cell key, index); index in the low word => a cell's points ascend by original index ----------
  for (int i = tid; i < n; i += ES_THREADS) {
    const float4 p = __ldg(src + i);
    sortbuf[i] = ((u64)ece_point_key(p, e, i) << 32) | (u64)(uint32_t)i;
  }
  __syncthreads();
  bitonic_sort_smem(sortbuf, n);

  ES_CLK(2);
  // ---- 3. unpack: run heads, original indices, then the coordinates overwrite the sort buffer ----------------
  const int ipt = cdiv(n, ES_THREADS);  // <= ES_MAX / ES_THREADS
  {
    constexpr int IPT_MAX = ES_MAX / ES_THREADS;
    const int j0 = min(tid * ipt, n), j1 = min(j0 + ipt, n);
    unsigned short my_idx[IPT_MAX];
    unsigned headmask = 0u;
    uint32_t prev = (j0 > 0 && j0 < n) ? hi32(sortbuf[j0 - 1]) : 0xffffffffu;
#pragma unroll
    for (int k = 0; k < IPT_MAX; ++k) {
      const int j = j0 + k;
      if (j < j1) {
        const u64 v = sortbuf[j];
        my_idx[k] = (unsigned short)lo32(v);
        if (j == 0 || hi32(v) != prev) headmask |= 1u << k;
        prev = hi32(v);
      }
    }
    int total;
    int pos = block_excl_scan(__popc(headmask), sm.wscan, total);  // (syncs: every thread has read its entries)
#pragma unroll
    for (int k = 0; k < IPT_MAX; ++k) {
      const int j = j0 + k;
      if (j < j1) {
        const float4 p = __ldg(src + my_idx[k]);
        px[j] = p.x;
        py[j] = p.y;
        pz[j] = p.z;
        idx16[j] = my_idx[k];
        parent[j] = j;
        if ((headmask >> k) & 1u) cell_start[pos++] = (unsigned short)j;
      }
    }
    if (tid == 0) {
      cell_start[total] = (unsigned short)n;
      sm.n_cells = total;
    }
    for (int i = tid; i < ES_HASH / 2; i += ES_THREADS) reinterpret_cast<unsigned*>(hash)[i] = 0u;
    __syncthreads();
  }
  ES_CLK(3);
  const int nc = sm.n_cells;
  EceSmallView view{px, py, pz, idx16, cell_start, hash, e};
  for (int c = tid; c < nc; c += ES_THREADS) {
    const uint32_t key = view.cell_key(c);
    if (key >= 0x40000000u) continue;  // private cell of a non-finite point: never looked up
    uint32_t s = hash_slot(key);
    while (atomicCAS(&hash[s], (unsigned short)0, (unsigned short)(c + 1)) != 0) s = (s + 1) & (ES_HASH - 1);
  }
  __syncthreads();

  ES_CLK(4);
  // ---- 4. union-find over sorted positions ------------------------------------------------------------------
  const int dimx = e.dim[0], dimy = e.dim[1], dimz = e.dim[2];
  if (e.mode == 0) {
    for (int c = warp; c < nc; c += ES_WARPS) {
      const int j0 = cell_start[c], j1 = cell_start[c + 1];
      for (int j = j0 + 1 + lane; j < j1; j += 32) parent[j] = j0;  // a clique cell is one super-node
      const uint32_t keyA = view.cell_key(c);
      if (keyA >= 0x40000000u) continue;
      const int cx = (int)(keyA % (uint32_t)dimx);
      const int cy = (int)((keyA / (uint32_t)dimx) % (uint32_t)dimy);
      const int cz = (int)(keyA / ((uint32_t)dimx * (uint32_t)dimy));
      // the 62 forward cells within 2 per axis: q < 2: (dx,0,0) = (q+1,0,0); 2 <= q < 12: dz = 0, dy = 1,2;
      // q >= 12: dz = 1,2, dy = -2..2; dx = -2..2
      for (int q = lane; q < 62; q += 32) {
        int dx, dy, dz;
        if (q < 2) {
          dx = q + 1;
          dy = 0;
          dz = 0;
        } else if (q < 12) {
          const int t = q - 2;
          dz = 0;
          dy = 1 + t / 5;
          dx = t % 5 - 2;
        } else {
          const int t = q - 12;
          dz = 1 + t / 25;
          dy = (t % 25) / 5 - 2;
          dx = t % 5 - 2;
        }
        const int xx = cx + dx, yy = cy + dy, zz = cz + dz;
        if (xx < 0 || xx >= dimx || yy < 0 || yy >= dimy || zz >= dimz) continue;
        const uint32_t keyB = (uint32_t)xx + (uint32_t)dimx * ((uint32_t)yy + (uint32_t)dimy * (uint32_t)zz);
        const int cb = view.lookup(keyB);
        if (cb < 0) continue;
        const int jb0 = cell_start[cb], jb1 = cell_start[cb + 1];
        if (uf_find(parent, j0) == uf_find(parent, jb0)) continue;
        bool hit = false;
        for (int a = j0; a < j1 && !hit; ++a) {
          const float ax = px[a], ay = py[a], az = pz[a];
          for (int b = jb0; b < jb1; ++b) {
            if (dist2(ax, ay, az, px[b], py[b], pz[b]) < r2) {
              hit = true;
              break;
            }
          }
        }
        if (hit) uf_union(parent, j0, jb0);
      }
    }
  } else {
    for (int j = tid; j < n; j += ES_THREADS) {
      const float x = px[j], y = py[j], z = pz[j];
      const uint32_t key = ece_point_key(make_float4(x, y, z, 0.f), e, (int)idx16[j]);
      const int cx = (int)(key % (uint32_t)dimx);
      const int cy = (int)((key / (uint32_t)dimx) % (uint32_t)dimy);
      const int cz = (int)(key / ((uint32_t)dimx * (uint32_t)dimy));
      const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, dimx - 1);
      // the rest of the own cell directly follows j in sorted order
      {
        const int c_own = view.lookup(key);
        const int j_end = cell_start[c_own + 1];
        for (int q = j + 1; q < j_end; ++q)
          if (dist2(x, y, z, px[q], py[q], pz[q]) < r2) uf_union(parent, j, q);
      }
      // forward cells: (+1,0,0), then the rows (dy,dz) = (+1,0), (-1,+1), (0,+1), (+1,+1) with dx = -1..1
      for (int r = -1; r < 4; ++r) {
        const int dy = (r <= 0) ? (r + 1) : (r - 2);
        const int dz = (r <= 0) ? 0 : 1;
        const int yy = cy + dy, zz = cz + dz;
        if (yy < 0 || yy >= dimy || zz >= dimz) continue;
        const int xa = (r < 0) ? cx + 1 : x_lo;
        for (int xx = xa; xx <= x_hi; ++xx) {
          const uint32_t keyB = (uint32_t)xx + (uint32_t)dimx * ((uint32_t)yy + (uint32_t)dimy * (uint32_t)zz);
          const int cb = view.lookup(keyB);
          if (cb < 0) continue;
          const int jb0 = cell_start[cb], jb1 = cell_start[cb + 1];
          for (int q = jb0; q < jb1; ++q)
            if (dist2(x, y, z, px[q], py[q], pz[q]) < r2) uf_union(parent, j, q);
        }
      }
    }
  }
  __syncthreads();

  ES_CLK(5);
  // ---- 5. labels, sizes, smallest original index per component, kept roots in canonical order, CSR ----------
  for (int i = tid; i < n; i += ES_THREADS) {
    csize[i] = 0;                // (z is dead)
    minidx[i] = 0x7fffffff;      // (cell_start and the hash table are dead)
  }
  __syncthreads();
  for (int j = tid; j < n; j += ES_THREADS) {
    int root = parent[j];
    while (true) {
      const int up = parent[root];
      if (up == root) break;
      root = up;
    }
    parent[j] = root;  // racing readers see either an ancestor or the root
    atomicAdd(&csize[root], 1);
    atomicMin(&minidx[root], (int)idx16[j]);
  }
  __syncthreads();
  ES_CLK(6);
  {
    const int i0 = min(tid * ipt, n), i1 = min(i0 + ipt, n);
    int cnt = 0;
    for (int j = i0; j < i1; ++j) {
      const int sz = csize[j];
      cnt += (parent[j] == j && sz >= min_size && sz <= max_size) ? 1 : 0;
    }
    int total;
    int pos = block_excl_scan(cnt, sm.wscan, total);
    for (int j = i0; j < i1; ++j) {
      const int sz = csize[j];
      if (parent[j] == j && sz >= min_size && sz <= max_size)  // size desc, smallest original index asc
        sortbuf[pos++] = ((u64)(uint32_t)(n - sz) << 32) | ((u64)(uint32_t)minidx[j] << 16) | (u64)(uint32_t)j;
    }
    if (tid == 0) sm.n_clusters = total;
    __syncthreads();
  }
  ES_CLK(7);
  const int C = sm.n_clusters;
  bitonic_sort_smem(sortbuf, C);
  ES_CLK(8);
  {
    const int cpt = cdiv(max(C, 1), ES_THREADS);
    const int r0 = min(tid * cpt, C), r1 = min(r0 + cpt, C);
    int sum = 0;
    for (int r = r0; r < r1; ++r) sum += n - (int)hi32(sortbuf[r]);
    int total;
    int run = block_excl_scan(sum, sm.wscan, total);  // (syncs: minidx is dead from here on)
    for (int r = r0; r < r1; ++r) {
      const u64 v = sortbuf[r];
      offs[r] = run;
      rank_of[lo32(v) & 0xffffu] = (unsigned short)r;
      run += n - (int)hi32(v);
    }
    if (tid == 0) {
      offs[C] = total;
      sm.n_members = total;
      n_clusters[f] = C;
      n_cluster_pts[f] = total;
    }
    __syncthreads();
  }
  ES_CLK(9);
  const int L = sm.n_members;
  for (int j = tid; j < n; j += ES_THREADS) {
    const int root = parent[j];
    const int sz = csize[root];
    const bool kept = sz >= min_size && sz <= max_size;
    sortbuf[j] = kept ? (((u64)rank_of[root] << 32) | (u64)idx16[j]) : ~0ull;
  }
  __syncthreads();
  ES_CLK(10);
  bitonic_sort_smem(sortbuf, n);
  for (int j = tid; j < L; j += ES_THREADS) idx_out[j] = (int)lo32(sortbuf[j]);

  ES_CLK(11);
  // ---- 6. centroid + bounding radius, one warp per cluster --------------------------------------------------
  for (int c = warp; c < C; c += ES_WARPS) {
    const int b = offs[c], en = offs[c + 1];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int j = b + lane; j < en; j += 32) {
      const float4 p = __ldg(src + lo32(sortbuf[j]));
      sx += (double)p.x;
      sy += (double)p.y;
      sz += (double)p.z;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      sx += __shfl_xor_sync(FULL, sx, o);
      sy += __shfl_xor_sync(FULL, sy, o);
      sz += __shfl_xor_sync(FULL, sz, o);
    }
    const double cnt = (double)(en - b);
    const float cx = (float)(sx / cnt), cy = (float)(sy / cnt), cz = (float)(sz / cnt);
    float r = 0.0f;
    for (int j = b + lane; j < en; j += 32) {
      const float4 p = __ldg(src + lo32(sortbuf[j]));
      r = fmaxf(r, sqrtf(dist2(p.x, p.y, p.z, cx, cy, cz)));  // od.cpp:457-464 arithmetic
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) r = fmaxf(r, __shfl_xor_sync(FULL, r, o));
    if (lane == 0) obs[c] = make_float4(cx, cy, cz, r);
  }
  __syncthreads();
  ES_CLK(12);
}

}  // namespace

void run_cluster_small(const Ctx& c, const ClusterArgs& a, int small_max) {
  cudaFuncSetAttribute(k_ece_small, cudaFuncAttributeMaxDynamicSharedMemorySize, ES_SMEM_BYTES);  // per device, idempotent
  const float r2 = (float)((double)a.tol * (double)a.tol);  // KdTreeFLANN::radiusSearch: (float)(radius*radius)
  KL(c, "k_ece_small", k_ece_small<<<c.B, ES_THREADS, ES_SMEM_BYTES, c.stream>>>(
      a.in, a.in_stride, a.n_in, a.tol, r2, a.min_size, a.max_size, small_max, a.offsets, a.indices, a.n_clusters,
      a.n_cluster_pts, a.obstacles, c.cap));
  count_launch(c);
}

}  // namespace pcop

#ifdef PCOP_ECE_DEBUG_CLK
extern "C" int pcop_debug_ece_small_cycles(long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, pcop::g_ece_small_clk, sizeof(long long) * 16);
}
#endif
