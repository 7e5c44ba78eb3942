// Euclidean cluster extraction (reference: extract_euclidian_clusters od.cpp:430-455 with the
// kd-tree of od.cpp:791-792 -> pcl::EuclideanClusterExtraction) and per-cluster
// centroid + bounding radius (msg/PointWithRad.msg:1-4; distance arithmetic od.cpp:457-464).
//
// The kd-tree radius search is replaced by a uniform grid with cell >= tolerance*(1+2^-8):
// points are radix-sorted by cell key, every point tests the 13 "forward" neighbour cells plus
// the rest of its own cell with the exact float predicate d2 < r2, and matching pairs are merged
// in an atomic-min union-find whose roots are the smallest original index of each component, so
// labels are deterministic.  Clusters = connected components, filtered by size, ordered by
// size descending then smallest index ascending, indices ascending inside a cluster.
#include <algorithm>
#include <cstdlib>

#include "ece_common.cuh"
#include "internal.cuh"
#include "primitives.cuh"

namespace pcop {

namespace {

__global__ void k_ece_setup(const MinMax* __restrict__ minmax, const int* __restrict__ n_in, float tol, int clique,
                            EceFrame* __restrict__ ef, int B) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  float mn[3], mx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = ord2f(minmax[f].mn[a]);
    mx[a] = ord2f(minmax[f].mx[a]);
  }
  ef[f] = ece_make_frame(mn, mx, n_in[f], tol, clique);
}

__global__ void __launch_bounds__(CT_THREADS)
    k_ece_keys(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in,
               const EceFrame* __restrict__ ef, uint32_t* __restrict__ keys, uint32_t* __restrict__ maxkey,
               int* __restrict__ parent, int* __restrict__ csize, int cap) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * CT_TILE >= n) return;
  const EceFrame e = ef[f];
  const float4* src = in + (size_t)f * in_stride;
  uint32_t mk = 0;
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int i = ct_index(tile, k);
    if (i < n) {
      const float4 p = __ldg(src + i);
      const uint32_t key = ece_point_key(p, e, i);
      keys[(size_t)f * cap + i] = key;
      if (parent) {
        parent[(size_t)f * cap + i] = i;
        csize[(size_t)f * cap + i] = 0;
      }
      mk = max(mk, key);
    }
  }
  mk = __reduce_max_sync(FULL, mk);
  if (lane_id() == 0 && mk) atomicMax(&maxkey[f], mk);
}

// sorted copy of the cloud: xyz + original index in w
__global__ void __launch_bounds__(CT_THREADS)
    k_ece_gather(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in,
                 const uint32_t* __restrict__ val0, const uint32_t* __restrict__ val1, const int* __restrict__ npass,
                 float4* __restrict__ sorted_pts, int cap) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * CT_TILE >= n) return;
  const uint32_t* vs = ((npass[f] & 1) ? val1 : val0) + (size_t)f * cap;
  const float4* src = in + (size_t)f * in_stride;
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int j = ct_index(tile, k);
    if (j < n) {
      const uint32_t i = vs[j];
      float4 p = __ldg(src + i);
      p.w = __uint_as_float(i);
      sorted_pts[(size_t)f * cap + j] = p;
    }
  }
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t* a, int n, uint32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
    k_ece_union(const float4* __restrict__ sorted_pts, const uint32_t* __restrict__ key0,
                const uint32_t* __restrict__ key1, const int* __restrict__ npass, const int* __restrict__ n_in,
                const EceFrame* __restrict__ ef, int* __restrict__ parent, float r2, int cap) {
  const int f = blockIdx.y;
  const int n = n_in[f];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (ef[f].mode != 1) return;
  const uint32_t* ks = ((npass[f] & 1) ? key1 : key0) + (size_t)f * cap;
  const float4* sp = sorted_pts + (size_t)f * cap;
  int* par = parent + (size_t)f * cap;
  const EceFrame e = ef[f];
  const float4 p = sp[j];
  const int pi = (int)__float_as_uint(p.w);
  const uint32_t key = ks[j];
  const int cx = (int)(key % (uint32_t)e.dim[0]);
  const int cy = (int)((key / (uint32_t)e.dim[0]) % (uint32_t)e.dim[1]);
  const int cz = (int)(key / ((uint32_t)e.dim[0] * (uint32_t)e.dim[1]));
  const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, e.dim[0] - 1);

  // own row: the rest of the own cell and the +x cell directly follow j in sorted order
  {
    const uint32_t hi_key = key - (uint32_t)cx + (uint32_t)x_hi;
    for (int q = j + 1; q < n && ks[q] <= hi_key; ++q) {
      const float4 o = sp[q];
      if (dist2(p.x, p.y, p.z, o.x, o.y, o.z) < r2) uf_union(par, pi, (int)__float_as_uint(o.w));
    }
  }
  // the four forward rows: (dy,dz) = (+1,0), (-1,+1), (0,+1), (+1,+1)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int dy = (r == 0) ? 1 : (r - 2);
    const int dz = (r == 0) ? 0 : 1;
    const int yy = cy + dy, zz = cz + dz;
    if (yy < 0 || yy >= e.dim[1] || zz >= e.dim[2]) continue;
    const uint32_t row = (uint32_t)e.dim[0] * ((uint32_t)yy + (uint32_t)e.dim[1] * (uint32_t)zz);
    const uint32_t lo_key = row + (uint32_t)x_lo, hi_key = row + (uint32_t)x_hi;
    for (int q = lower_bound_u32(ks, n, lo_key); q < n && ks[q] <= hi_key; ++q) {
      const float4 o = sp[q];
      if (dist2(p.x, p.y, p.z, o.x, o.y, o.z) < r2) uf_union(par, pi, (int)__float_as_uint(o.w));
    }
  }
}

// clique mode: first sorted position + key of every occupied cell (stable compaction of the run heads)
__global__ void __launch_bounds__(CT_THREADS)
    k_ece_cell_heads(const uint32_t* __restrict__ key0, const uint32_t* __restrict__ key1, const int* __restrict__ npass,
                     const int* __restrict__ n_in, const EceFrame* __restrict__ ef,
                     const float4* __restrict__ sorted_pts, int* __restrict__ cell_start, uint32_t* __restrict__ cell_key,
                     int* __restrict__ cell_rep, int* __restrict__ n_cells, unsigned* __restrict__ desc, int cap,
                     int tiles) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * CT_TILE >= n || ef[f].mode != 0) {
    if (tile == 0 && threadIdx.x == 0) n_cells[f] = 0;
    return;
  }
  __shared__ CompactSmem sm;
  const uint32_t* ks = ((npass[f] & 1) ? key1 : key0) + (size_t)f * cap;
  bool keep[CT_ITEMS];
  unsigned pos[CT_ITEMS];
  uint32_t kk[CT_ITEMS];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int j = ct_index(tile, k);
    keep[k] = false;
    kk[k] = 0;
    if (j < n) {
      kk[k] = ks[j];
      keep[k] = j == 0 || kk[k] != ks[j - 1];
    }
  }
  const unsigned incl_total = tile_compact_positions(keep, pos, desc + (size_t)f * tiles, tile, sm);
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k)
    if (keep[k]) {
      cell_start[(size_t)f * cap + pos[k]] = ct_index(tile, k);
      cell_key[(size_t)f * cap + pos[k]] = kk[k];
      // representative = first sorted point of the cell = its smallest original index (stable sort)
      cell_rep[(size_t)f * cap + pos[k]] = (int)__float_as_uint(sorted_pts[(size_t)f * cap + ct_index(tile, k)].w);
    }
  if ((tile + 1) * CT_TILE >= n && threadIdx.x == 0) n_cells[f] = (int)incl_total;
}

// clique mode: one warp per occupied cell A.  The cell's representative is its first sorted point, which (stable
// sort) is its smallest original index; the other points are hung under it.  A is merged with each of the 62
// "forward" cells B within 2 cells per axis as soon as ONE pair (a in A, b in B) passes the exact predicate; pairs
// are only tested while A and B are still in different sets.  Lanes 0..12 each locate the candidate cells of one
// forward row (binary search in the sorted cell-key list) in parallel; the warp then walks the found cells.
__global__ void __launch_bounds__(256)
    k_ece_cell_union(const float4* __restrict__ sorted_pts, const int* __restrict__ cell_start,
                     const uint32_t* __restrict__ cell_key, const int* __restrict__ cell_rep,
                     const int* __restrict__ n_cells, const int* __restrict__ n_in, const EceFrame* __restrict__ ef,
                     int* __restrict__ parent, float r2, int cap) {
  const int f = blockIdx.y;
  const int nc = n_cells[f];
  const int warps_per_frame = gridDim.x * (blockDim.x >> 5);
  const int lane = lane_id();
  int c = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (c >= nc) return;
  const int n = n_in[f];
  const EceFrame e = ef[f];
  const float4* sp = sorted_pts + (size_t)f * cap;
  const int* cs = cell_start + (size_t)f * cap;
  const uint32_t* ck = cell_key + (size_t)f * cap;
  const int* crep = cell_rep + (size_t)f * cap;
  int* par = parent + (size_t)f * cap;
  const uint32_t dimx = (uint32_t)e.dim[0], dimy = (uint32_t)e.dim[1];
  for (; c < nc; c += warps_per_frame) {
    const int j0 = cs[c];
    const int j1 = (c + 1 < nc) ? cs[c + 1] : n;
    const uint32_t keyA = ck[c];
    const int repA = crep[c];
    for (int j = j0 + 1 + lane; j < j1; j += 32) par[(int)__float_as_uint(sp[j].w)] = repA;
    if (keyA >= 0x40000000u) continue;  // private cell of a non-finite point
    const int cx = (int)(keyA % dimx);
    const int cy = (int)((keyA / dimx) % dimy);
    const int cz = (int)(keyA / (dimx * dimy));
    const int x_lo = max(cx - 2, 0), x_hi = min(cx + 2, e.dim[0] - 1);
    // lane r < 13: first candidate cell and number of candidate cells (<= 5) of forward row r.
    // row 0 = own row (cells x+1, x+2 follow A directly); r = 1,2: dz = 0, dy = 1,2;
    // r = 3..12: dz = 1 + (r-3)/5, dy = (r-3)%5 - 2
    int my_start = nc, my_cnt = 0;
    if (lane < 13) {
      uint32_t hi_key = 0;
      bool ok = true;
      if (lane == 0) {
        my_start = c + 1;
        hi_key = keyA - (uint32_t)cx + (uint32_t)x_hi;
      } else {
        const int dz = (lane < 3) ? 0 : 1 + (lane - 3) / 5;
        const int dy = (lane < 3) ? lane : (lane - 3) % 5 - 2;
        const int yy = cy + dy, zz = cz + dz;
        ok = !(yy < 0 || yy >= e.dim[1] || zz >= e.dim[2]);
        if (ok) {
          const uint32_t row = dimx * ((uint32_t)yy + dimy * (uint32_t)zz);
          const uint32_t lo_key = row + (uint32_t)x_lo;
          hi_key = row + (uint32_t)x_hi;
          int lo = c + 1, hi = nc;  // forward rows have larger keys than A
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ck[mid] < lo_key) lo = mid + 1;
            else hi = mid;
          }
          my_start = lo;
        }
      }
      if (ok) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int cb = my_start + k;
          const uint32_t kb = (cb < nc) ? ck[cb] : 0xffffffffu;
          if (kb <= hi_key) my_cnt = k + 1;  // keys ascend, so the matches are a prefix
        }
      }
    }
    // spread the candidate cells (<= 62) over the lanes: lane q handles candidate q, q + 32
    int excl = my_cnt;  // inclusive scan over lanes (rows), then made exclusive
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int up = __shfl_up_sync(FULL, excl, o);
      if (lane >= o) excl += up;
    }
    const int total_cand = __shfl_sync(FULL, excl, 12);
    excl -= my_cnt;
    for (int q0 = 0; q0 < total_cand; q0 += 32) {
      const int q = q0 + lane;
      int cb = -1;
#pragma unroll
      for (int r = 0; r < 13; ++r) {
        const int e_r = __shfl_sync(FULL, excl, r);
        const int c_r = __shfl_sync(FULL, my_cnt, r);
        const int s_r = __shfl_sync(FULL, my_start, r);
        if (q >= e_r && q < e_r + c_r) cb = s_r + (q - e_r);
      }
      if (cb >= 0) {
        const int repB = crep[cb];
        if (uf_find(par, repA) != uf_find(par, repB)) {
          const int jb0 = cs[cb];
          const int jb1 = (cb + 1 < nc) ? cs[cb + 1] : n;
          bool hit = false;
          for (int a = j0; a < j1 && !hit; ++a) {
            const float4 pa = sp[a];
            for (int b = jb0; b < jb1; ++b) {
              const float4 pb = sp[b];
              if (dist2(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z) < r2) {
                hit = true;
                break;
              }
            }
          }
          if (hit) uf_union(par, repA, repB);
        }
      }
      __syncwarp();
    }
  }
}

// label = root; component sizes (warp-aggregated atomics)
__global__ void __launch_bounds__(256) k_ece_flatten(int* __restrict__ parent, int* __restrict__ csize,
                                                       const int* __restrict__ n_in, int cap) {
  const int f = blockIdx.y;
  const int n = n_in[f];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  int* par = parent + (size_t)f * cap;
  int root = -1 - (int)threadIdx.x;  // unique dummy for idle lanes
  if (valid) {
    root = par[i];
    while (true) {
      const int up = par[root];
      if (up == root) break;
      root = up;
    }
    par[i] = root;
  }
  const unsigned m = __match_any_sync(FULL, root);
  if (valid && (m & lanemask_lt()) == 0u) atomicAdd(&csize[(size_t)f * cap + root], __popc(m));
}

// kept roots in ascending index order + their sort key (n - size: ascending key = descending size)
__global__ void __launch_bounds__(CT_THREADS)
    k_ece_roots(const int* __restrict__ parent, const int* __restrict__ csize, const int* __restrict__ n_in,
                int min_size, int max_size, int* __restrict__ roots, uint32_t* __restrict__ keys,
                uint32_t* __restrict__ maxkey, int* __restrict__ n_clusters, unsigned* __restrict__ desc, int cap,
                int tiles) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * CT_TILE >= n) {
    if (tile == 0 && threadIdx.x == 0) n_clusters[f] = 0;
    return;
  }
  __shared__ CompactSmem sm;
  bool keep[CT_ITEMS];
  unsigned pos[CT_ITEMS];
  int sz[CT_ITEMS];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int i = ct_index(tile, k);
    keep[k] = false;
    sz[k] = 0;
    if (i < n && parent[(size_t)f * cap + i] == i) {
      sz[k] = csize[(size_t)f * cap + i];
      keep[k] = sz[k] >= min_size && sz[k] <= max_size;
    }
  }
  const unsigned incl_total = tile_compact_positions(keep, pos, desc + (size_t)f * tiles, tile, sm);
  uint32_t mk = 0;
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    if (keep[k]) {
      roots[(size_t)f * cap + pos[k]] = ct_index(tile, k);
      const uint32_t key = (uint32_t)(n - sz[k]);
      keys[(size_t)f * cap + pos[k]] = key;
      mk = max(mk, key);
    }
  }
  mk = __reduce_max_sync(FULL, mk);
  if (lane_id() == 0 && mk) atomicMax(&maxkey[f], mk);
  if ((tile + 1) * CT_TILE >= n && threadIdx.x == 0) n_clusters[f] = (int)incl_total;
}

// one block per frame: walk the size-sorted roots, assign ranks and CSR offsets
__global__ void __launch_bounds__(256)
    k_ece_rank(const int* __restrict__ roots, const uint32_t* __restrict__ val0, const uint32_t* __restrict__ val1,
               const int* __restrict__ npass, const int* __restrict__ csize, const int* __restrict__ n_clusters,
               int* __restrict__ rank_of, int* __restrict__ offsets, int* __restrict__ n_cluster_pts, int cap) {
  const int f = blockIdx.x;
  const int C = n_clusters[f];
  const uint32_t* vs = ((npass[f] & 1) ? val1 : val0) + (size_t)f * cap;
  int* offs = offsets + (size_t)f * (cap + 1);
  __shared__ int wsum[8];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < C; base += 256) {
    const int r = base + threadIdx.x;
    int sz = 0, root = -1;
    if (r < C) {
      root = roots[(size_t)f * cap + vs[r]];
      sz = csize[(size_t)f * cap + root];
      rank_of[(size_t)f * cap + root] = r;
    }
    int incl = sz;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(FULL, incl, o);
      if (lane_id() >= o) incl += up;
    }
    if (lane_id() == 31) wsum[warp_id()] = incl;
    __syncthreads();
    int wb = carry;
    for (int w = 0; w < warp_id(); ++w) wb += wsum[w];
    if (r < C) offs[r] = wb + incl - sz;
    __syncthreads();
    if (threadIdx.x == 255) carry = wb + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    offs[C] = carry;
    n_cluster_pts[f] = carry;
  }
}

// per point: sort key = rank of its cluster, or C for points outside every kept cluster
__global__ void __launch_bounds__(CT_THREADS)
    k_ece_member_keys(const int* __restrict__ parent, const int* __restrict__ csize, const int* __restrict__ rank_of,
                      const int* __restrict__ n_in, const int* __restrict__ n_clusters, int min_size, int max_size,
                      uint32_t* __restrict__ keys, uint32_t* __restrict__ maxkey, int cap) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * CT_TILE >= n) return;
  const uint32_t C = (uint32_t)n_clusters[f];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int i = ct_index(tile, k);
    if (i < n) {
      const int root = parent[(size_t)f * cap + i];
      const int sz = csize[(size_t)f * cap + root];
      const bool kept = sz >= min_size && sz <= max_size;
      keys[(size_t)f * cap + i] = kept ? (uint32_t)rank_of[(size_t)f * cap + root] : C;
    }
  }
  if (tile == 0 && threadIdx.x == 0 && C) atomicMax(&maxkey[f], C);
}

// the first L sorted values are the CSR indices
__global__ void __launch_bounds__(256)
    k_ece_indices(const uint32_t* __restrict__ val0, const uint32_t* __restrict__ val1, const int* __restrict__ npass,
                  const int* __restrict__ n_cluster_pts, int* __restrict__ indices, int cap) {
  const int f = blockIdx.y;
  const int L = n_cluster_pts[f];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L) return;
  const uint32_t* vs = ((npass[f] & 1) ? val1 : val0) + (size_t)f * cap;
  indices[(size_t)f * cap + j] = (int)vs[j];
}

// Centroid + bounding radius of every kept cluster (msg/PointWithRad.msg; distance arithmetic of od.cpp:457-464).
// Work items: a cluster of up to CR_CHUNK members is one item (one block sums it, takes the mean and the radius); a
// larger cluster is cut into CR_CHUNK-member pieces whose partial sums are combined in piece order by a second kernel
// (deterministic, whatever the block scheduling), and a third pass takes the radius (a maximum: order-free).  Every
// block enumerates the items by walking the cluster offsets (the host does not know the cluster sizes).
constexpr int CR_CHUNK = 8192;
constexpr int CR_GRID = 128;

struct CrItem {
  int c, piece, pieces, pslot;  // cluster, piece of it, number of pieces, first partial slot of the cluster
};
// item `want` of the frame (items in cluster order, pieces in order), or c = -1 past the end
__device__ __forceinline__ CrItem cr_find_item(const int* __restrict__ offs, int C, int want, bool big_only) {
  int item = 0, pslot = 0;
  for (int c = 0; c < C; ++c) {
    const int pieces = max(1, cdiv(offs[c + 1] - offs[c], CR_CHUNK));
    const int cnt = (big_only && pieces == 1) ? 0 : pieces;
    if (want < item + cnt) return CrItem{c, want - item, pieces, pslot};
    item += cnt;
    if (pieces > 1) pslot += pieces;
  }
  return CrItem{-1, 0, 0, 0};
}

__device__ __forceinline__ void cr_block_sum3(double& sx, double& sy, double& sz, double (*sh)[3]) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    sx += __shfl_xor_sync(FULL, sx, o);
    sy += __shfl_xor_sync(FULL, sy, o);
    sz += __shfl_xor_sync(FULL, sz, o);
  }
  __syncthreads();  // (sh may still be read from the previous item)
  if (lane_id() == 0) {
    sh[warp_id()][0] = sx;
    sh[warp_id()][1] = sy;
    sh[warp_id()][2] = sz;
  }
  __syncthreads();
}
__device__ __forceinline__ float cr_block_max(float r, float* shr) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) r = fmaxf(r, __shfl_xor_sync(FULL, r, o));
  __syncthreads();
  if (lane_id() == 0) shr[warp_id()] = r;
  __syncthreads();
  float rr = shr[0];
  for (int w = 1; w < 8; ++w) rr = fmaxf(rr, shr[w]);
  return rr;
}

// pass 1: every item's sum; single-piece clusters are finished here
__global__ void __launch_bounds__(256)
    k_centroid_radius(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ offsets,
                      const int* __restrict__ indices, const int* __restrict__ n_clusters, float4* __restrict__ obstacles,
                      double* __restrict__ partial, int partial_stride, int cap) {
  const int f = blockIdx.y;
  const int C = n_clusters[f];
  const float4* src = in + (size_t)f * in_stride;
  const int* offs = offsets + (size_t)f * (cap + 1);
  const int* idx = indices + (size_t)f * cap;
  __shared__ double sh[8][3];
  __shared__ float shc[3];
  __shared__ float shr[8];
  for (int item = blockIdx.x;; item += gridDim.x) {
    const CrItem it = cr_find_item(offs, C, item, false);
    if (it.c < 0) break;
    const int b = offs[it.c] + it.piece * CR_CHUNK, e = min(offs[it.c + 1], b + CR_CHUNK);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int j = b + threadIdx.x; j < e; j += blockDim.x) {
      const float4 p = __ldg(src + idx[j]);
      sx += (double)p.x;
      sy += (double)p.y;
      sz += (double)p.z;
    }
    cr_block_sum3(sx, sy, sz, sh);
    if (it.pieces > 1) {  // a piece of a large cluster: partial sum for k_centroid_combine
      if (threadIdx.x < 3) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
        if (it.pslot + it.piece < partial_stride / 3) partial[(size_t)f * partial_stride + (size_t)(it.pslot + it.piece) * 3 + threadIdx.x] = s;
      }
      continue;
    }
    if (threadIdx.x < 3) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
      shc[threadIdx.x] = (float)(s / (double)(e - b));
    }
    __syncthreads();
    const float cx = shc[0], cy = shc[1], cz = shc[2];
    float r = 0.0f;
    for (int j = b + threadIdx.x; j < e; j += blockDim.x) {
      const float4 p = __ldg(src + idx[j]);
      r = fmaxf(r, sqrtf(dist2(p.x, p.y, p.z, cx, cy, cz)));  // od.cpp:457-464 arithmetic
    }
    const float rr = cr_block_max(r, shr);
    if (threadIdx.x == 0) obstacles[(size_t)f * cap + it.c] = make_float4(cx, cy, cz, rr);
  }
}

// pass 2 (one warp per frame): centroids of the large clusters from their pieces' sums, in piece order; radius 0 for now
__global__ void __launch_bounds__(32)
    k_centroid_combine(const int* __restrict__ offsets, const int* __restrict__ n_clusters, float4* __restrict__ obstacles,
                       const double* __restrict__ partial, int partial_stride, int cap) {
  const int f = blockIdx.x;
  const int C = n_clusters[f];
  const int* offs = offsets + (size_t)f * (cap + 1);
  int pslot = 0;
  for (int c = 0; c < C; ++c) {
    const int n = offs[c + 1] - offs[c];
    const int pieces = max(1, cdiv(n, CR_CHUNK));
    if (pieces == 1) continue;
    if (threadIdx.x < 3) {
      double s = 0.0;
      for (int k = 0; k < pieces; ++k)
        if (pslot + k < partial_stride / 3) s += partial[(size_t)f * partial_stride + (size_t)(pslot + k) * 3 + threadIdx.x];
      (&obstacles[(size_t)f * cap + c].x)[threadIdx.x] = (float)(s / (double)n);
    }
    if (threadIdx.x == 3) obstacles[(size_t)f * cap + c].w = 0.0f;
    pslot += pieces;
  }
}

// pass 3: radius of the large clusters, piece by piece (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256)
    k_radius_big(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ offsets,
                 const int* __restrict__ indices, const int* __restrict__ n_clusters, float4* __restrict__ obstacles, int cap) {
  const int f = blockIdx.y;
  const int C = n_clusters[f];
  const float4* src = in + (size_t)f * in_stride;
  const int* offs = offsets + (size_t)f * (cap + 1);
  const int* idx = indices + (size_t)f * cap;
  __shared__ float shr[8];
  for (int item = blockIdx.x;; item += gridDim.x) {
    const CrItem it = cr_find_item(offs, C, item, true);
    if (it.c < 0) break;
    const int b = offs[it.c] + it.piece * CR_CHUNK, e = min(offs[it.c + 1], b + CR_CHUNK);
    const float4 cen = obstacles[(size_t)f * cap + it.c];
    float r = 0.0f;
    for (int j = b + threadIdx.x; j < e; j += blockDim.x) {
      const float4 p = __ldg(src + idx[j]);
      r = fmaxf(r, sqrtf(dist2(p.x, p.y, p.z, cen.x, cen.y, cen.z)));
    }
    const float rr = cr_block_max(r, shr);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned*>(&obstacles[(size_t)f * cap + it.c].w), __float_as_uint(rr));
  }
}

}  // namespace

void run_grid_sort(const Ctx& c, const float4* in, size_t in_stride, const int* n_in, float cell, MinMax* minmax,
                   EceFrame* ef, const SortBufs& sort, float4* sorted_pts, int* parent, int* csize, int clique) {
  const int tiles = cdiv(c.cap, CT_TILE);        // descriptor stride
  const int gtiles = cdiv(c.grid_cap, CT_TILE);  // blocks actually launched per frame
  run_minmax(c, in, in_stride, n_in, minmax, /*finite_only=*/true);
  KL(c, "k_ece_setup", k_ece_setup<<<cdiv(c.B, 128), 128, 0, c.stream>>>(minmax, n_in, cell, clique, ef, c.B));
  sort_reset_maxkey(c, sort);
  KL(c, "k_ece_keys", k_ece_keys<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(in, in_stride, n_in, ef, sort.key[0], sort.maxkey, parent,
                                                            csize, c.cap));
  count_launch(c, 2);
  radix_sort_batched(c, sort, n_in, true);
  KL(c, "k_ece_gather", k_ece_gather<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(in, in_stride, n_in, sort.val[0], sort.val[1], sort.npass,
                                                              sorted_pts, c.cap));
  count_launch(c);
}

// frames the fused shared-memory kernel handles look empty to the generic kernels
__global__ void k_ece_route(const int* __restrict__ n_in, int* __restrict__ n_generic, int small_max, int B) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < B) n_generic[f] = (n_in[f] > small_max) ? n_in[f] : 0;
}

static void run_cluster_generic(const Ctx& c, const ClusterArgs& a) {
  const int tiles = cdiv(c.cap, CT_TILE);        // descriptor stride
  const int gtiles = cdiv(c.grid_cap, CT_TILE);  // blocks actually launched per frame
  const float r2 = (float)((double)a.tol * (double)a.tol);  // KdTreeFLANN::radiusSearch: (float)(radius*radius)
  run_grid_sort(c, a.in, a.in_stride, a.n_in, a.tol, a.minmax, a.ef, a.sort, a.sorted_pts, a.parent, a.csize, /*clique=*/1);
  // frames in clique mode: union-find over cells
  cudaMemsetAsync(a.desc, 0, (size_t)c.B * tiles * sizeof(unsigned), c.stream);
  KL(c, "k_ece_cell_heads", k_ece_cell_heads<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(
      a.sort.key[0], a.sort.key[1], a.sort.npass, a.n_in, a.ef, a.sorted_pts, a.cell_start, a.cell_key,
      /*cell_rep (scratch, reused for the root list afterwards)=*/a.roots, a.n_cells, a.desc, c.cap, tiles));
  // (a warp per cell, grid-stride: the kernel waits on its binary searches, so a call of one or two large frames gets
  // enough blocks to fill every SM with eight of them)
  const int union_blocks = max(1, min(cdiv(c.grid_cap, 16), max(512, 1184 / c.B)));
  KL(c, "k_ece_cell_union", k_ece_cell_union<<<dim3(union_blocks, c.B), 256, 0, c.stream>>>(
      a.sorted_pts, a.cell_start, a.cell_key, a.roots, a.n_cells, a.n_in, a.ef, a.parent, r2, c.cap));
  count_launch(c, 2);
  // frames whose extent does not fit 1024 clique cells per axis: per-point neighbour scan
  KL(c, "k_ece_union", k_ece_union<<<dim3(cdiv(c.grid_cap, 256), c.B), 256, 0, c.stream>>>(a.sorted_pts, a.sort.key[0], a.sort.key[1], a.sort.npass,
                                                                 a.n_in, a.ef, a.parent, r2, c.cap));
  KL(c, "k_ece_flatten", k_ece_flatten<<<dim3(cdiv(c.grid_cap, 256), c.B), 256, 0, c.stream>>>(a.parent, a.csize, a.n_in, c.cap));
  count_launch(c, 2);
  // kept roots, ordered by size descending (stable => smallest index first among equals)
  cudaMemsetAsync(a.desc, 0, (size_t)c.B * tiles * sizeof(unsigned), c.stream);
  sort_reset_maxkey(c, a.sort);
  KL(c, "k_ece_roots", k_ece_roots<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(a.parent, a.csize, a.n_in, a.min_size, a.max_size, a.roots,
                                                             a.sort.key[0], a.sort.maxkey, a.n_clusters, a.desc, c.cap,
                                                             tiles));
  count_launch(c);
  radix_sort_batched(c, a.sort, a.n_clusters, true);
  KL(c, "k_ece_rank", k_ece_rank<<<c.B, 256, 0, c.stream>>>(a.roots, a.sort.val[0], a.sort.val[1], a.sort.npass, a.csize, a.n_clusters,
                                        a.rank_of, a.offsets, a.n_cluster_pts, c.cap));
  // CSR indices: stable sort of the points by cluster rank
  sort_reset_maxkey(c, a.sort);
  KL(c, "k_ece_member_keys", k_ece_member_keys<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(a.parent, a.csize, a.rank_of, a.n_in, a.n_clusters,
                                                                   a.min_size, a.max_size, a.sort.key[0], a.sort.maxkey,
                                                                   c.cap));
  count_launch(c, 2);
  radix_sort_batched(c, a.sort, a.n_in, true);
  KL(c, "k_ece_indices", k_ece_indices<<<dim3(cdiv(c.grid_cap, 256), c.B), 256, 0, c.stream>>>(a.sort.val[0], a.sort.val[1], a.sort.npass,
                                                                   a.n_cluster_pts, a.indices, c.cap));
  count_launch(c);
}

int ece_small_limit() {
  const char* s = getenv("PCOP_ECE_SMALL_MAX");
  if (!s || !*s) return ECE_SMALL_MAX;
  const long v = strtol(s, nullptr, 10);
  return (int)std::min<long>(std::max<long>(v, 0), ECE_SMALL_MAX);
}

// Frames with at most a.small_max points are clustered by the fused shared-memory kernel (which also writes
// their obstacles); the others by the generic path.  The host does not know the remaining-cloud sizes when it
// enqueues the stage (nothing synchronises between the input upload and the wave's counts), so:
//   with_generic = false  only the fused kernel is launched; frames above small_max get empty results and the caller
//                         repeats the stage with with_generic = true once it has seen the counts (pcop_api.cu);
//   with_generic = true   the generic kernels run as well, on the frames above small_max only.
// Returns true when the generic centroid/radius kernel still has to run.
bool run_cluster(const Ctx& c, const ClusterArgs& a, bool with_generic) {
  const int small_max = a.small_max;
  const bool maybe_big = c.grid_cap > small_max;  // grid_cap bounds every frame's point count
  const bool generic = maybe_big && (with_generic || small_max <= 0);
  if (generic) {
    ClusterArgs g = a;
    if (small_max > 0) {
      KL(c, "k_ece_route", k_ece_route<<<cdiv(c.B, 256), 256, 0, c.stream>>>(a.n_in, a.n_route, small_max, c.B));
      count_launch(c);
      g.n_in = a.n_route;
    }
    run_cluster_generic(c, g);
  }
  if (small_max > 0) run_cluster_small(c, a, small_max, /*zero_skipped=*/!generic);
  return generic;
}

void run_centroid_radius(const Ctx& c, const ClusterArgs& a) {
  KL(c, "k_centroid_radius", k_centroid_radius<<<dim3(CR_GRID, c.B), 256, 0, c.stream>>>(a.in, a.in_stride, a.offsets, a.indices, a.n_clusters,
                                                                                       a.obstacles, a.partial, a.partial_stride, c.cap));
  count_launch(c);
  if (c.grid_cap > CR_CHUNK) {  // some cluster may be larger than one piece
    KL(c, "k_centroid_combine", k_centroid_combine<<<c.B, 32, 0, c.stream>>>(a.offsets, a.n_clusters, a.obstacles, a.partial, a.partial_stride, c.cap));
    KL(c, "k_radius_big", k_radius_big<<<dim3(CR_GRID, c.B), 256, 0, c.stream>>>(a.in, a.in_stride, a.offsets, a.indices, a.n_clusters,
                                                                               a.obstacles, c.cap));
    count_launch(c, 2);
  }
}

}  // namespace pcop
