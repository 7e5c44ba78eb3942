// Fused Euclidean clustering for small clouds: ONE thread block per frame, everything in shared memory.
//
// After ground removal the cloud that reaches pcl::EuclideanClusterExtraction (od.cpp:796) is a few thousand
// points (BASELINE configs[1]: ~5 k of 120 k).  At that size the generic path (stage_cluster.cu: ~30 launches,
// three global radix sorts, union-find in L2) is pure launch and latency overhead.  For frames with at most
// ES_MAX points this kernel runs the whole stage -- search grid, connected components, size filter, canonical
// ordering, CSR indices and the per-cluster centroid + bounding radius -- in one launch:
//
//   1. min/max of the finite points -> grid (same rules as the generic path, ece_common.cuh);
//   2. points grouped by cell WITHOUT a sort: cell keys (10 bits per axis) go into an open-addressing hash table
//      (atomicCAS), occupied slots get dense cell ids by a block scan, points take a slot inside their cell with
//      one shared-memory atomicAdd, coordinates are scattered into cell order.  The order of the points inside a
//      cell depends on the atomics, but nothing observable does: components do not depend on the test order and
//      every output is ordered by original index afterwards;
//   3. union-find in shared memory, atomic-min hooking.
//      Clique mode (cell edge < tol/sqrt(3)): the nodes are the CELLS.  One thread per cell looks up its 62
//      forward neighbour cells (hash probes only), then walks the cells it found: point pairs are tested with the
//      exact float predicate only while the two cells are still in different sets and only until the first hit.
//      Point mode (extent too large for clique cells): the nodes are the points; one thread per point scans the
//      rest of its cell and the 13 forward neighbour cells;
//   4. per component: size and smallest original index (shared-memory atomics per node), kept roots ranked by
//      (size desc, smallest index asc), CSR offsets;
//   5. cluster_indices: labels by original index, then a block-wide stable multi-way split by cluster rank
//      (<= ES_SPLIT_MAXC clusters; a bitonic sort of (rank, index) otherwise);
//   6. member coordinates staged in shared memory; one warp per cluster: centroid (double sums) and bounding radius
//      (od.cpp:457-464 arithmetic).
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
constexpr int ES_MAX = ECE_SMALL_MAX;
constexpr int ES_IPT = (ES_MAX + ES_THREADS - 1) / ES_THREADS;  // points (cells) per thread
constexpr int ES_HASH_BITS = 14;
constexpr int ES_HASH = 1 << ES_HASH_BITS;
constexpr int ES_SCAN_MAXC = 128;   // cluster offsets of up to this many clusters are mirrored in shared memory
constexpr int ES_SPLIT_MAXC = 1024;  // up to this many clusters the members are listed by a block-wide stable split
constexpr uint32_t ES_EMPTY = 0xffffffffu;
typedef unsigned long long u64;

// Shared-memory regions (bytes) and what lives in them per phase:
//   A  12*ES_MAX  build: u32 tk[ES_HASH] (keys of the hash slots), int cnt[ES_MAX] (points per cell) behind it
//                 union: float x[], y[], z[] in cell order
//                 after the union: int csize[], int minidx[], u16 lab16[] (by original index), u16 rankc[] (by node)
//   D  2*ES_HASH  u16 hash[] (cell id + 1 per slot)            \ after the union: u32 rootkey[], u16 rootnode[] (kept
//   E  4*ES_MAX   u32 cell_key[] (by cell id)                  /  roots), then u32 member sort keys (many clusters)
//   B  2*ES_MAX   u16 parent[] (by node; nodes < 2^16, atomic-min by 32-bit CAS on the containing word)
//   C  2*ES_MAX   u16 idx16[] (original index of each cell-ordered position), then the member list
//   F  2*ES_MAX+16  u16 node_start[] (first position of each cell; identity in point mode after the union)
constexpr int ES_A = 0;
constexpr int ES_D = ES_A + 12 * ES_MAX;
constexpr int ES_E = ES_D + 2 * ES_HASH;
constexpr int ES_B = ES_E + 4 * ES_MAX;
constexpr int ES_C = ES_B + 2 * ES_MAX;
constexpr int ES_F = ES_C + 2 * ES_MAX;
constexpr int ES_MISC = ES_F + 2 * ES_MAX + 16;
struct EceSmallMisc {
  float red[ES_WARPS][6];
  int wscan[ES_WARPS + 1];
  EceFrame ef;
  int n_cells, n_clusters, n_members;
  int soff[ES_SCAN_MAXC + 1];
  int fwd[64];  // packed-key offset of forward neighbour q
};
constexpr int ES_SMEM_BYTES = ES_MISC + (int)sizeof(EceSmallMisc);
static_assert(ES_SMEM_BYTES <= 227 * 1024, "fused clustering kernel: shared memory budget");
static_assert(4 * ES_HASH + 4 * ES_MAX <= 12 * ES_MAX, "build-phase key table + cell counts overlay the coordinates");
static_assert(2 * ES_HASH + 4 * ES_MAX >= 6 * ES_MAX, "kept-root keys + nodes overlay hash + cell keys");
static_assert(ES_MAX < 16384, "positions / indices / sizes are packed in 14 bits");
static_assert(ES_MAX * 16 <= ES_HASH * 9, "hash load factor <= 0.5625");
static_assert(2 * ES_WARPS * ES_SPLIT_MAXC <= 8 * ES_MAX, "split offsets overlay sizes + smallest indices");
static_assert(ES_MISC % 8 == 0 && ES_D % 16 == 0 && ES_B % 16 == 0 && ES_F % 4 == 0 && ES_C % 4 == 0, "alignment");

__device__ __forceinline__ uint32_t hi32(u64 v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ uint32_t lo32(u64 v) { return (uint32_t)v; }

// Normalised bitonic network (every comparator puts the smaller element at the lower index), so elements past n
// behave as +infinity without being stored: a comparator whose upper end is >= n is a no-op.
template <class T>
__device__ void bitonic_sort_smem(T* a, int n) {
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
        const T x = a[i], y = a[p];
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
          const T x = a[i], y = a[p];
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

// cell key with 10 bits per axis (the grid has at most 1024 cells per axis); clique mode: a point with a
// non-finite coordinate is within tol of nothing and gets a private cell (bit 30 set)
__device__ __forceinline__ uint32_t es_point_key(const float4 p, const EceFrame& e, int i) {
  const int cx = cell_coord(p.x, e.mn[0], e.inv[0], e.dim[0]);
  const int cy = cell_coord(p.y, e.mn[1], e.inv[1], e.dim[1]);
  const int cz = cell_coord(p.z, e.mn[2], e.inv[2], e.dim[2]);
  uint32_t key = (uint32_t)cx | ((uint32_t)cy << 10) | ((uint32_t)cz << 20);
  if (e.mode == 0 && !(fabsf(p.x) <= 3.0e38f && fabsf(p.y) <= 3.0e38f && fabsf(p.z) <= 3.0e38f))
    key = 0x40000000u | (uint32_t)i;
  return key;
}

// cell id of `key` or -1
__device__ __forceinline__ int es_lookup(const unsigned short* hash, const uint32_t* cell_key, uint32_t key) {
  uint32_t s = hash_slot(key);
  while (true) {
    const unsigned v = hash[s];
    if (v == 0u) return -1;
    if (cell_key[v - 1] == key) return (int)v - 1;
    s = (s + 1) & (ES_HASH - 1);
  }
}

// Union-find on shared memory.  Parents only ever point to smaller node ids and hooks only ever touch roots, so a
// plain store of an ancestor into a non-root entry (path compression) is safe next to concurrent atomicMin hooks.
__device__ __forceinline__ int es_find(unsigned short* parent, int v) {
  volatile unsigned short* par = parent;
  while (true) {  // path halving
    const int p = par[v];
    if (p == v) return v;
    const int gp = par[p];
    if (gp == p) return p;
    par[v] = (unsigned short)gp;
    v = gp;
  }
}
// atomicMin on a 16-bit shared-memory entry: 32-bit CAS on the containing word (a concurrent plain store to the
// other half makes the CAS fail and retry, so it is never lost); returns the old value
__device__ __forceinline__ int es_atomic_min16(unsigned short* base, int idx, int val) {
  unsigned* w = reinterpret_cast<unsigned*>(base + (idx & ~1));
  const int shift = (idx & 1) * 16;
  unsigned cur = *reinterpret_cast<volatile unsigned*>(w);
  while (true) {
    const int old16 = (int)((cur >> shift) & 0xffffu);
    if (old16 <= val) return old16;
    const unsigned nw = (cur & ~(0xffffu << shift)) | ((unsigned)val << shift);
    const unsigned prev = atomicCAS(w, cur, nw);
    if (prev == cur) return old16;
    cur = prev;
  }
}
__device__ __forceinline__ void es_union(unsigned short* parent, int a, int b) {
  while (true) {
    a = es_find(parent, a);
    b = es_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = es_atomic_min16(parent, a, b);  // hook the larger root under the smaller
    if (old == a) return;
    a = old;  // a was no longer a root: merge its (former) parent with b instead
  }
}

// phase timestamps of block 0 (clock64), for tools/ece_phases.py; written only when PCOP_ECE_DEBUG_CLK is defined
#ifdef PCOP_ECE_DEBUG_CLK
#define ES_CLK(k)                                                      \
  do {                                                                 \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_ece_small_clk[k] = clock64(); \
  } while (0)
#define ES_CLKW(k)                                                                      \
  do {                                                                                  \
    if (blockIdx.x == 0 && threadIdx.x == 0 && c0 == 0) g_ece_small_clk[k] = clock64(); \
  } while (0)
#else
#define ES_CLKW(k) \
  do {             \
  } while (0)
#define ES_CLK(k) \
  do {            \
  } while (0)
#endif

// The 62 forward neighbours of a clique cell = the cells within 2 per axis that follow it in (z, y, x) order, listed by
// increasing distance: near cells are the likely links, so once they are merged most of the far cells are already in
// the same set when their turn comes and their pair tests are skipped.  X(q, dx, dy, dz)
#define ES_FWD_LIST(X) \
  X(0, 1, 0, 0) X(1, 0, 1, 0) X(2, 0, 0, 1) X(3, -1, 1, 0) X(4, 1, 1, 0) X(5, 0, -1, 1) \
  X(6, -1, 0, 1) X(7, 1, 0, 1) X(8, 0, 1, 1) X(9, -1, -1, 1) X(10, 1, -1, 1) X(11, -1, 1, 1) \
  X(12, 1, 1, 1) X(13, 2, 0, 0) X(14, 0, 2, 0) X(15, 0, 0, 2) X(16, -2, 1, 0) X(17, 2, 1, 0) \
  X(18, -1, 2, 0) X(19, 1, 2, 0) X(20, 0, -2, 1) X(21, -2, 0, 1) X(22, 2, 0, 1) X(23, 0, 2, 1) \
  X(24, 0, -1, 2) X(25, -1, 0, 2) X(26, 1, 0, 2) X(27, 0, 1, 2) X(28, -1, -2, 1) X(29, 1, -2, 1) \
  X(30, -2, -1, 1) X(31, 2, -1, 1) X(32, -2, 1, 1) X(33, 2, 1, 1) X(34, -1, 2, 1) X(35, 1, 2, 1) \
  X(36, -1, -1, 2) X(37, 1, -1, 2) X(38, -1, 1, 2) X(39, 1, 1, 2) X(40, -2, 2, 0) X(41, 2, 2, 0) \
  X(42, 0, -2, 2) X(43, -2, 0, 2) X(44, 2, 0, 2) X(45, 0, 2, 2) X(46, -2, -2, 1) X(47, 2, -2, 1) \
  X(48, -2, 2, 1) X(49, 2, 2, 1) X(50, -1, -2, 2) X(51, 1, -2, 2) X(52, -2, -1, 2) X(53, 2, -1, 2) \
  X(54, -2, 1, 2) X(55, 2, 1, 2) X(56, -1, 2, 2) X(57, 1, 2, 2) X(58, -2, -2, 2) X(59, 2, -2, 2) \
  X(60, -2, 2, 2) X(61, 2, 2, 2)
#define ES_FWD_OFF_ENTRY(q, dx, dy, dz) (dx) + (dy)*1024 + (dz)*1048576,
__device__ const int ES_FWD_OFF[62] = {ES_FWD_LIST(ES_FWD_OFF_ENTRY)};  // packed-key offset of forward neighbour q

__global__ void __launch_bounds__(ES_THREADS, 1)
    k_ece_small(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in, float tol, float r2,
                int min_size, int max_size, int small_max, int zero_skipped, int* __restrict__ offsets,
                int* __restrict__ indices, int* __restrict__ n_clusters, int* __restrict__ n_cluster_pts,
                float4* __restrict__ obstacles, int cap) {
  const int f = blockIdx.x;
  const int n = n_in[f];
  if (n > small_max) {  // this frame takes the generic path (zero_skipped: it has not run yet -- empty result for now)
    if (zero_skipped && threadIdx.x == 0) {
      n_clusters[f] = 0;
      n_cluster_pts[f] = 0;
      offsets[(size_t)f * (cap + 1)] = 0;
    }
    return;
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* tk = reinterpret_cast<uint32_t*>(smem_raw + ES_A);
  float* px = reinterpret_cast<float*>(smem_raw + ES_A);
  float* py = px + ES_MAX;
  float* pz = py + ES_MAX;
  int* csize = reinterpret_cast<int*>(smem_raw + ES_A);
  int* minidx = csize + ES_MAX;
  unsigned short* lab16 = reinterpret_cast<unsigned short*>(smem_raw + ES_A + 8 * ES_MAX);
  unsigned short* rankc = lab16 + ES_MAX;
  unsigned short* hash = reinterpret_cast<unsigned short*>(smem_raw + ES_D);
  uint32_t* cell_key = reinterpret_cast<uint32_t*>(smem_raw + ES_E);
  uint32_t* rootkey = reinterpret_cast<uint32_t*>(smem_raw + ES_D);  // ((n - size) << 16) | smallest index
  unsigned short* rootnode = reinterpret_cast<unsigned short*>(smem_raw + ES_D + 4 * ES_MAX);
  uint32_t* memkey = reinterpret_cast<uint32_t*>(smem_raw + ES_D);   // (rank << 14) | index
  int* cnt = reinterpret_cast<int*>(smem_raw + ES_A + 4 * ES_HASH);
  unsigned short* parent = reinterpret_cast<unsigned short*>(smem_raw + ES_B);
  unsigned short* idx16 = reinterpret_cast<unsigned short*>(smem_raw + ES_C);
  unsigned short* node_start = reinterpret_cast<unsigned short*>(smem_raw + ES_F);
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

#ifdef PCOP_ECE_DEBUG_CLK
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int k = 10; k < 16; ++k) g_ece_small_clk[k] = 0;
#endif
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
    if (tid < 62) sm.fwd[tid] = ES_FWD_OFF[tid];
    for (int s = tid; s < ES_HASH; s += ES_THREADS) tk[s] = ES_EMPTY;
    for (int i = tid; i < n; i += ES_THREADS) cnt[i] = 0;
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
  // ---- 2. group the points by cell ---------------------------------------------------------------------------
  // point k of this thread is original index tid + k*ES_THREADS; mine[k] = hash slot, then cell id | slot in cell << 16
  uint32_t mine[ES_IPT];
#pragma unroll
  for (int k = 0; k < ES_IPT; ++k) {
    const int i = tid + k * ES_THREADS;
    mine[k] = 0u;
    if (i < n) {
      const uint32_t key = es_point_key(__ldg(src + i), e, i);
      uint32_t s = hash_slot(key);
      while (true) {
        const uint32_t old = atomicCAS(&tk[s], ES_EMPTY, key);
        if (old == ES_EMPTY || old == key) break;
        s = (s + 1) & (ES_HASH - 1);
      }
      mine[k] = s;
    }
  }
  __syncthreads();
  {  // dense cell ids for the occupied slots (any fixed bijection will do: slot t + k*ES_THREADS, thread-major)
    unsigned occ = 0u;
#pragma unroll
    for (int k = 0; k < ES_HASH / ES_THREADS; ++k) occ |= (tk[tid + k * ES_THREADS] != ES_EMPTY) ? (1u << k) : 0u;
    int total;
    int pos = block_excl_scan(__popc(occ), sm.wscan, total);
#pragma unroll
    for (int k = 0; k < ES_HASH / ES_THREADS; ++k) {
      const int s = tid + k * ES_THREADS;
      if ((occ >> k) & 1u) {
        cell_key[pos] = tk[s];
        hash[s] = (unsigned short)(pos + 1);
        ++pos;
      } else {
        hash[s] = 0;
      }
    }
    if (tid == 0) sm.n_cells = total;
    __syncthreads();
  }
  const int nc = sm.n_cells;
  ES_CLK(2);
#pragma unroll
  for (int k = 0; k < ES_IPT; ++k) {
    const int i = tid + k * ES_THREADS;
    if (i < n) {
      const uint32_t id = (uint32_t)hash[mine[k]] - 1u;
      const uint32_t r = (uint32_t)atomicAdd(&cnt[id], 1);
      mine[k] = id | (r << 16);
    }
  }
  __syncthreads();
  {  // exclusive scan of the cell counts -> first position of each cell
    int c0 = min(tid * ES_IPT, nc), c1 = min(c0 + ES_IPT, nc);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += cnt[c];
    int total;
    int run = block_excl_scan(sum, sm.wscan, total);
    for (int c = c0; c < c1; ++c) {
      node_start[c] = (unsigned short)run;
      run += cnt[c];
    }
    if (tid == 0) node_start[nc] = (unsigned short)n;
    __syncthreads();  // (cnt and tk are dead from here on)
  }
#pragma unroll
  for (int k = 0; k < ES_IPT; ++k) {
    const int i = tid + k * ES_THREADS;
    if (i < n) {
      const float4 p = __ldg(src + i);
      const int j = (int)node_start[mine[k] & 0xffffu] + (int)(mine[k] >> 16);
      px[j] = p.x;
      py[j] = p.y;
      pz[j] = p.z;
      idx16[j] = (unsigned short)i;
    }
  }
  const int nn = (e.mode == 0) ? nc : n;  // union-find nodes: cells (clique mode) or points
  for (int v = tid; v < nn; v += ES_THREADS) parent[v] = (unsigned short)v;
  __syncthreads();

  ES_CLK(3);
  // ---- 3. union-find ----------------------------------------------------------------------------------------
  const int dimx = e.dim[0], dimy = e.dim[1], dimz = e.dim[2];
  if (e.mode == 0) {
    // warp-uniform loops with explicit reconvergence: without it the lanes of a warp drift apart in the
    // data-dependent loops below and execute them one lane at a time
    for (int c0 = warp * 32; c0 < nc; c0 += ES_THREADS) {
      const int c = c0 + lane;
      const uint32_t keyA = (c < nc) ? cell_key[c] : 0x40000000u;
      const bool live = !(keyA & 0x40000000u);  // (private cells of non-finite points have no neighbours)
      const int cx = (int)(keyA & 1023u), cy = (int)((keyA >> 10) & 1023u), cz = (int)((keyA >> 20) & 1023u);
      // per-axis validity of the offsets -2..2 (bit d+2); inside the grid the packed key of a neighbour is keyA + a
      // constant (no carries between the 10-bit fields)
      unsigned xm = 0u, ym = 0u, zm = 0u;
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        xm |= (live && cx + d >= 0 && cx + d < dimx) ? (1u << (d + 2)) : 0u;
        ym |= (live && cy + d >= 0 && cy + d < dimy) ? (1u << (d + 2)) : 0u;
        zm |= (live && cz + d >= 0 && cz + d < dimz) ? (1u << (d + 2)) : 0u;
      }
      // Which of the 62 forward cells may exist: the first probe of every hash lookup, unconditional and independent
      // (62 loads in flight, no divergence).  The table is < 10 % full, so an empty first slot settles most of them;
      // a non-empty one is resolved (full lookup) when its turn comes in the walk below.
      u64 found = 0ull;
#define ES_PROBE(q, dx, dy, dz)                                                                           \
  {                                                                                                       \
    const bool ok = ((xm >> ((dx) + 2)) & (ym >> ((dy) + 2)) & (zm >> ((dz) + 2)) & 1u) != 0u;            \
    const unsigned v = hash[hash_slot(keyA + (uint32_t)((dx) + (dy)*1024 + (dz)*1048576))];               \
    if (ok && v != 0u) found |= 1ull << (q);                                                              \
  }
      ES_FWD_LIST(ES_PROBE)
#undef ES_PROBE
      __syncwarp();
      ES_CLKW(10);
      const int j0 = live ? node_start[c] : 0, j1 = live ? node_start[c + 1] : 0;
      int rc = live ? c : 0;
      // bounding box of the cell's own points: a point b of another cell whose box distance exceeds r2 cannot pass the
      // exact predicate against any of them, so two cells that are NOT linked cost |B| box tests instead of |A| x |B|
      // exact tests.  The pre-test is exact, not approximate: lo <= a <= hi per axis and rounding is monotonic, so
      // fl(lo - b) or fl(b - hi) never exceeds |fl(a - b)|, and the squares are summed in dist2's order without FMA,
      // hence box_d2 <= dist2(a, b) in float for every a of the cell (the 1e-4 margin is slack on top of that)
      float lox = 3.402823466e+38f, loy = 3.402823466e+38f, loz = 3.402823466e+38f;
      float hix = -3.402823466e+38f, hiy = -3.402823466e+38f, hiz = -3.402823466e+38f;
      for (int a = j0; a < j1; ++a) {
        lox = fminf(lox, px[a]);
        hix = fmaxf(hix, px[a]);
        loy = fminf(loy, py[a]);
        hiy = fmaxf(hiy, py[a]);
        loz = fminf(loz, pz[a]);
        hiz = fmaxf(hiz, pz[a]);
      }
      const float r2m = r2 * 1.0001f;
      while (__any_sync(FULL, found != 0ull)) {
        if (found) {
          const int qq = __ffsll((long long)found) - 1;
          found &= found - 1ull;
          const int cb = es_lookup(hash, cell_key, keyA + (uint32_t)sm.fwd[qq]);  // -1: the slot held another cell
          if (cb >= 0) rc = es_find(parent, rc);
          if (cb >= 0 && rc != es_find(parent, cb)) {
            const int jb0 = node_start[cb], jb1 = node_start[cb + 1];
            bool hit = false;
            for (int b = jb0; b < jb1 && !hit; ++b) {
              const float bx = px[b], by = py[b], bz = pz[b];
              const float ex = fmaxf(fmaxf(lox - bx, bx - hix), 0.0f), ey = fmaxf(fmaxf(loy - by, by - hiy), 0.0f),
                          ez = fmaxf(fmaxf(loz - bz, bz - hiz), 0.0f);
              if (ex * ex + ey * ey + ez * ez > r2m) continue;
              for (int a = j0; a < j1; ++a) {
                if (dist2(px[a], py[a], pz[a], bx, by, bz) < r2) {
                  hit = true;
                  break;
                }
              }
            }
            if (hit) es_union(parent, rc, cb);
          }
        }
        __syncwarp();
      }
      ES_CLKW(11);
    }
  } else {
    for (int c = tid; c < nc; c += ES_THREADS) {
      const uint32_t keyA = cell_key[c];
      const int cx = (int)(keyA & 1023u), cy = (int)((keyA >> 10) & 1023u), cz = (int)(keyA >> 20);
      const int j0 = node_start[c], j1 = node_start[c + 1];
      // inside the cell
      for (int a = j0; a < j1; ++a) {
        const float ax = px[a], ay = py[a], az = pz[a];
        for (int b = a + 1; b < j1; ++b)
          if (dist2(ax, ay, az, px[b], py[b], pz[b]) < r2) es_union(parent, a, b);
      }
      // the 13 forward cells: (+1,0,0), then the rows (dy,dz) = (+1,0), (-1,+1), (0,+1), (+1,+1) with dx = -1..1
      for (int r = -1; r < 4; ++r) {
        const int dy = (r <= 0) ? (r + 1) : (r - 2);
        const int dz = (r <= 0) ? 0 : 1;
        const int yy = cy + dy, zz = cz + dz;
        if (yy < 0 || yy >= dimy || zz >= dimz) continue;
        for (int xx = (r < 0) ? cx + 1 : cx - 1; xx <= cx + 1; ++xx) {
          if (xx < 0 || xx >= dimx) continue;
          const uint32_t keyB = (uint32_t)xx | ((uint32_t)yy << 10) | ((uint32_t)zz << 20);
          const int cb = es_lookup(hash, cell_key, keyB);
          if (cb < 0) continue;
          const int jb0 = node_start[cb], jb1 = node_start[cb + 1];
          for (int a = j0; a < j1; ++a) {
            const float ax = px[a], ay = py[a], az = pz[a];
            for (int b = jb0; b < jb1; ++b)
              if (dist2(ax, ay, az, px[b], py[b], pz[b]) < r2) es_union(parent, a, b);
          }
        }
      }
    }
  }
  __syncthreads();

  ES_CLK(4);
  // ---- 4. sizes, smallest original index per component, kept roots in canonical order, CSR offsets -----------
  if (e.mode != 0) {  // point mode: from here on every point is its own node
    for (int j = tid; j <= n; j += ES_THREADS) node_start[j] = (unsigned short)j;
  }
  for (int v = tid; v < nn; v += ES_THREADS) {  // (the coordinates are dead)
    csize[v] = 0;
    minidx[v] = 0x7fffffff;
  }
  __syncthreads();
  for (int v = tid; v < nn; v += ES_THREADS) {
    int root = v;
    while (true) {
      const int up = parent[root];
      if (up == root) break;
      root = up;
    }
    const int j0 = node_start[v], j1 = node_start[v + 1];
    int mi = 0x7fffffff;
    for (int j = j0; j < j1; ++j) mi = min(mi, (int)idx16[j]);
    atomicAdd(&csize[root], j1 - j0);
    atomicMin(&minidx[root], mi);
    parent[v] = (unsigned short)root;  // racing readers see either an ancestor or the root
  }
  __syncthreads();
  ES_CLK(5);
  const int npt = cdiv(nn, ES_THREADS);
  {
    const int v0 = min(tid * npt, nn), v1 = min(v0 + npt, nn);
    int k = 0;
    for (int v = v0; v < v1; ++v) {
      const int sz = csize[v];
      k += ((int)parent[v] == v && sz >= min_size && sz <= max_size) ? 1 : 0;
    }
    int total;
    int pos = block_excl_scan(k, sm.wscan, total);
    for (int v = v0; v < v1; ++v) {
      const int sz = csize[v];
      if ((int)parent[v] == v && sz >= min_size && sz <= max_size) {  // size desc, smallest original index asc
        rootkey[pos] = ((uint32_t)(n - sz) << 16) | (uint32_t)minidx[v];
        rootnode[pos] = (unsigned short)v;
        ++pos;
      }
    }
    if (tid == 0) sm.n_clusters = total;
    __syncthreads();
  }
  const int C = sm.n_clusters;
  {  // rank by counting, in place (the keys are distinct: a smallest index belongs to one component)
    uint32_t mykey[ES_IPT];
    unsigned short mynode[ES_IPT], myrank[ES_IPT];
#pragma unroll
    for (int k = 0; k < ES_IPT; ++k) {
      const int r = tid + k * ES_THREADS;
      mykey[k] = (r < C) ? rootkey[r] : 0xffffffffu;
      mynode[k] = (r < C) ? rootnode[r] : (unsigned short)0;
      myrank[k] = 0;
    }
    for (int o = 0; o < C; ++o) {
      const uint32_t other = rootkey[o];  // (broadcast)
#pragma unroll
      for (int k = 0; k < ES_IPT; ++k) myrank[k] += (other < mykey[k]) ? 1 : 0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ES_IPT; ++k) {
      if (tid + k * ES_THREADS < C) {
        rootkey[myrank[k]] = mykey[k];
        rootnode[myrank[k]] = mynode[k];
      }
    }
    __syncthreads();
  }
  ES_CLK(6);
  {
    const int cpt = cdiv(max(C, 1), ES_THREADS);
    const int r0 = min(tid * cpt, C), r1 = min(r0 + cpt, C);
    int sum = 0;
    for (int r = r0; r < r1; ++r) sum += n - (int)(rootkey[r] >> 16);
    int total;
    int run = block_excl_scan(sum, sm.wscan, total);
    for (int r = r0; r < r1; ++r) {
      offs[r] = run;
      if (r < ES_SCAN_MAXC) sm.soff[r] = run;
      rankc[rootnode[r]] = (unsigned short)r;
      run += n - (int)(rootkey[r] >> 16);
    }
    if (tid == 0) {
      offs[C] = total;
      if (C <= ES_SCAN_MAXC) sm.soff[C] = total;
      sm.n_members = total;
      n_clusters[f] = C;
      n_cluster_pts[f] = total;
    }
    __syncthreads();
  }
  const int L = sm.n_members;
  // ---- 5. labels by original index, cluster_indices, centroids ---------------------------------------------------
  for (int v = tid; v < nn; v += ES_THREADS) {
    const int root = parent[v];
    const int sz = csize[root];
    const unsigned short lab = (sz >= min_size && sz <= max_size) ? rankc[root] : (unsigned short)0xffffu;
    const int j0 = node_start[v], j1 = node_start[v + 1];
    for (int j = j0; j < j1; ++j) lab16[idx16[j]] = lab;
  }
  __syncthreads();
  ES_CLK(7);
  unsigned short* mem16 = idx16;  // member list (the cell-ordered index list is dead)
  if (C <= ES_SPLIT_MAXC) {
    // stable multi-way split of the points by cluster rank: warp w owns the index range [w*chunk, (w+1)*chunk);
    // per-(warp, cluster) counts -> running offsets -> positions (match_any ranks the lanes of one cluster)
    unsigned short* woff = reinterpret_cast<unsigned short*>(smem_raw + ES_A);  // [ES_WARPS][C]  (sizes are dead)
    const int chunk = cdiv(cdiv(n, ES_WARPS), 32) * 32;
    const int i0 = warp * chunk, i1 = min(i0 + chunk, n);
    for (int k = tid; k < ES_WARPS * C; k += ES_THREADS) woff[k] = 0;
    __syncthreads();
    for (int base = i0; base < i1; base += 32) {
      const int i = base + lane;
      const unsigned lab = (i < i1) ? (unsigned)lab16[i] : 0xffffu;
      const unsigned peers = __match_any_sync(FULL, lab);
      if (lab != 0xffffu && (peers & lanemask_lt()) == 0u) woff[warp * C + lab] += (unsigned short)__popc(peers);
      __syncwarp();
    }
    __syncthreads();
    for (int c = tid; c < C; c += ES_THREADS) {
      int run = (C <= ES_SCAN_MAXC) ? sm.soff[c] : offs[c];
      for (int w = 0; w < ES_WARPS; ++w) {
        const int t = woff[w * C + c];
        woff[w * C + c] = (unsigned short)run;
        run += t;
      }
    }
    __syncthreads();
    for (int base = i0; base < i1; base += 32) {
      const int i = base + lane;
      const unsigned lab = (i < i1) ? (unsigned)lab16[i] : 0xffffu;
      const unsigned peers = __match_any_sync(FULL, lab);
      if (lab != 0xffffu) {
        const int first = (int)woff[warp * C + lab];
        const int pos = first + __popc(peers & lanemask_lt());
        mem16[pos] = (unsigned short)i;
        idx_out[pos] = i;
        __syncwarp(peers);
        if ((peers & lanemask_lt()) == 0u) woff[warp * C + lab] = (unsigned short)(first + __popc(peers));
      }
      __syncwarp();
    }
  } else {
    for (int i = tid; i < n; i += ES_THREADS) {  // (the kept-root arrays are dead)
      const unsigned short lab = lab16[i];
      memkey[i] = (lab != 0xffffu) ? (((uint32_t)lab << 14) | (uint32_t)i) : 0xffffffffu;
    }
    __syncthreads();
    bitonic_sort_smem(memkey, n);
    for (int j = tid; j < L; j += ES_THREADS) {
      const uint32_t i = memkey[j] & 0x3fffu;
      mem16[j] = (unsigned short)i;
      idx_out[j] = (int)i;
    }
  }
  __syncthreads();
  ES_CLK(8);
  // ---- 6. member coordinates staged in shared memory (labels, sizes are dead), centroid + radius per cluster -------
  for (int j = tid; j < L; j += ES_THREADS) {
    const float4 p = __ldg(src + mem16[j]);
    px[j] = p.x;
    py[j] = p.y;
    pz[j] = p.z;
  }
  __syncthreads();
  for (int c = warp; c < C; c += ES_WARPS) {
    const int b = (C <= ES_SCAN_MAXC) ? sm.soff[c] : offs[c], en = (C <= ES_SCAN_MAXC) ? sm.soff[c + 1] : offs[c + 1];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int j = b + lane; j < en; j += 32) {
      sx += (double)px[j];
      sy += (double)py[j];
      sz += (double)pz[j];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      sx += __shfl_xor_sync(FULL, sx, o);
      sy += __shfl_xor_sync(FULL, sy, o);
      sz += __shfl_xor_sync(FULL, sz, o);
    }
    const double dn = (double)(en - b);
    const float cx = (float)(sx / dn), cy = (float)(sy / dn), cz = (float)(sz / dn);
    float r = 0.0f;
    for (int j = b + lane; j < en; j += 32)
      r = fmaxf(r, sqrtf(dist2(px[j], py[j], pz[j], cx, cy, cz)));  // od.cpp:457-464 arithmetic
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) r = fmaxf(r, __shfl_xor_sync(FULL, r, o));
    if (lane == 0) obs[c] = make_float4(cx, cy, cz, r);
  }
  ES_CLK(9);
}

}  // namespace

void run_cluster_small(const Ctx& c, const ClusterArgs& a, int small_max, bool zero_skipped) {
  cudaFuncSetAttribute(k_ece_small, cudaFuncAttributeMaxDynamicSharedMemorySize, ES_SMEM_BYTES);  // per device, idempotent
  const float r2 = (float)((double)a.tol * (double)a.tol);  // KdTreeFLANN::radiusSearch: (float)(radius*radius)
  KL(c, "k_ece_small", k_ece_small<<<c.B, ES_THREADS, ES_SMEM_BYTES, c.stream>>>(
      a.in, a.in_stride, a.n_in, a.tol, r2, a.min_size, a.max_size, small_max, zero_skipped ? 1 : 0, a.offsets, a.indices,
      a.n_clusters, a.n_cluster_pts, a.obstacles, c.cap));
  count_launch(c);
}

}  // namespace pcop

#ifdef PCOP_ECE_DEBUG_CLK
extern "C" int pcop_debug_ece_small_cycles(long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, pcop::g_ece_small_clk, sizeof(long long) * 16);
}
#endif
