// StatisticalOutlierRemoval (reference: remove_statistical_outliers od.cpp:316-340 ->
// pcl::StatisticalOutlierRemoval::applyFilterIndices).
//
// Exact k-NN (k = meanK + 1, the query itself included) on a uniform grid: each point grows a
// cube of cells ring by ring, keeps the k smallest float squared distances, and stops once the
// k-th is provably closer than anything outside the cube.  The mean distance uses sqrt in
// double over the sorted d2 (depends only on the multiset of the k smallest, so ties at the
// k-th place are harmless).  The global mean / variance are canonical tree sums in double; the
// keep test is !(distance > mean + mul*stddev); survivors keep their order.
#include "internal.cuh"
#include "primitives.cuh"

namespace pcop {

namespace {

constexpr int SOR_MAX_K = 64;

__global__ void k_sor_setup(const int* __restrict__ n_in, int meanK, double* __restrict__ thr,
                            uint32_t* __restrict__ warnings, int B) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  const int n = n_in[f];
  thr[f] = __longlong_as_double(0x7ff0000000000000ll);  // +inf: keep everything unless a threshold is computed
  if (n > 0 && n <= meanK) atomicOr(&warnings[f], (uint32_t)PCOP_WARN_SOR_TOO_FEW_POINTS);
}

__device__ __forceinline__ int lower_bound_key(const uint32_t* a, int n, uint32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128)
    k_sor_knn(const float4* __restrict__ sorted_pts, const uint32_t* __restrict__ key0, const uint32_t* __restrict__ key1,
              const int* __restrict__ npass, const int* __restrict__ n_in, const EceFrame* __restrict__ ef, int meanK,
              float* __restrict__ dist, int cap) {
  const int f = blockIdx.y;
  const int n = n_in[f];
  if (n <= meanK) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t* ks = ((npass[f] & 1) ? key1 : key0) + (size_t)f * cap;
  const float4* sp = sorted_pts + (size_t)f * cap;
  const EceFrame e = ef[f];
  const float4 p = sp[j];
  const uint32_t key = ks[j];
  const int dimx = e.dim[0], dimy = e.dim[1], dimz = e.dim[2];
  const int cx = (int)(key % (uint32_t)dimx);
  const int cy = (int)((key / (uint32_t)dimx) % (uint32_t)dimy);
  const int cz = (int)(key / ((uint32_t)dimx * (uint32_t)dimy));
  const int K = meanK + 1;
  float best[SOR_MAX_K];  // ascending
  int cnt = 0;
  // smallest real cell edge, shrunk: lower bound on the distance from p to anything outside the cube
  const float cell_lb = fminf(fminf(1.0f / e.inv[0], 1.0f / e.inv[1]), 1.0f / e.inv[2]) * 0.999f;
  const int rmax = max(max(max(cx, dimx - 1 - cx), max(cy, dimy - 1 - cy)), max(cz, dimz - 1 - cz));
  for (int r = 0; r <= rmax; ++r) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, dimz - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, dimy - 1);
    for (int zz = z0; zz <= z1; ++zz) {
      for (int yy = y0; yy <= y1; ++yy) {
        const bool shell_row = (abs(zz - cz) == r) || (abs(yy - cy) == r);
        const uint32_t row = (uint32_t)dimx * ((uint32_t)yy + (uint32_t)dimy * (uint32_t)zz);
        // shell rows scan the whole x range; interior rows only the two end cells x = cx-r, cx+r
        const int nseg = shell_row ? 1 : 2;
        for (int sgm = 0; sgm < nseg; ++sgm) {
          int xa, xb;
          if (shell_row) {
            xa = max(cx - r, 0);
            xb = min(cx + r, dimx - 1);
          } else {
            xa = xb = (sgm == 0) ? (cx - r) : (cx + r);
            if (xa < 0 || xa >= dimx) continue;
          }
          const uint32_t lo_key = row + (uint32_t)xa, hi_key = row + (uint32_t)xb;
          for (int q = lower_bound_key(ks, n, lo_key); q < n && ks[q] <= hi_key; ++q) {
            const float4 o = sp[q];
            const float d2 = dist2(p.x, p.y, p.z, o.x, o.y, o.z);
            if (cnt < K) {
              int t = cnt++;
              while (t > 0 && best[t - 1] > d2) {
                best[t] = best[t - 1];
                --t;
              }
              best[t] = d2;
            } else if (d2 < best[K - 1]) {
              int t = K - 1;
              while (t > 0 && best[t - 1] > d2) {
                best[t] = best[t - 1];
                --t;
              }
              best[t] = d2;
            }
          }
        }
      }
    }
    if (cnt == K) {
      const float reach = (float)r * cell_lb;
      if (best[K - 1] <= reach * reach) break;
    }
  }
  // PCL: dist_sum += sqrt(nn_dists[k]) for k = 1..meanK (element 0 is the query), double accumulator
  double sum = 0.0;
  for (int k = 1; k < K; ++k) sum = dadd(sum, __dsqrt_rn((double)best[k]));
  dist[(size_t)f * cap + (int)__float_as_uint(p.w)] = (float)ddiv(sum, (double)meanK);
}

// canonical tree sums of distances and of (float)(d*d), one 2048 chunk per block
__global__ void __launch_bounds__(256) k_sor_sums(const float* __restrict__ dist, const int* __restrict__ n_in, int meanK,
                                                    double* __restrict__ partial, int chunks, int cap) {
  const int f = blockIdx.y, chunk = blockIdx.x;
  const int n = n_in[f];
  if (n <= meanK || chunk * TS_CHUNK >= n) return;
  double a = 0.0, b = 0.0;
#pragma unroll
  for (int r = 0; r < TS_CHUNK / 256; ++r) {
    const int i = chunk * TS_CHUNK + r * 256 + threadIdx.x;
    double ea = 0.0, eb = 0.0;
    if (i < n) {
      const float d = dist[(size_t)f * cap + i];
      ea = (double)d;
      eb = (double)fmul(d, d);
    }
    a = dadd(a, ea);
    b = dadd(b, eb);
  }
  a = tree_butterfly(a);
  b = tree_butterfly(b);
  __shared__ double sh[8][2];
  if (lane_id() == 0) {
    sh[warp_id()][0] = a;
    sh[warp_id()][1] = b;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = sh[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) s = dadd(s, sh[w][threadIdx.x]);
    partial[((size_t)f * chunks + chunk) * 2 + threadIdx.x] = s;
  }
}

__global__ void k_sor_threshold(const double* __restrict__ partial, const int* __restrict__ n_in, int meanK, double mul,
                                double* __restrict__ thr, int chunks, int B) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  const int n = n_in[f];
  if (n <= meanK) return;
  const int nch = cdiv(n, TS_CHUNK);
  double sum = 0.0, sq = 0.0;
  for (int c = 0; c < nch; ++c) {
    sum = dadd(sum, partial[((size_t)f * chunks + c) * 2 + 0]);
    sq = dadd(sq, partial[((size_t)f * chunks + c) * 2 + 1]);
  }
  const double nn = (double)n;
  const double mean = ddiv(sum, nn);
  const double variance = ddiv(dsub(sq, ddiv(dmul(sum, sum), nn)), dsub(nn, 1.0));
  const double stddev = __dsqrt_rn(variance);
  thr[f] = dadd(mean, dmul(mul, stddev));
}

__global__ void __launch_bounds__(CT_THREADS)
    k_sor_filter(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in,
                 const float* __restrict__ dist, const double* __restrict__ thr, int meanK, float4* __restrict__ out,
                 int* __restrict__ kept_idx, int* __restrict__ n_out, unsigned* __restrict__ desc, int cap, int tiles) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * CT_TILE >= n) {
    if (tile == 0 && threadIdx.x == 0) n_out[f] = 0;
    return;
  }
  __shared__ CompactSmem sm;
  const bool pass_through = n <= meanK;
  const double t = thr[f];
  const float4* src = in + (size_t)f * in_stride;
  float4 p[CT_ITEMS];
  bool keep[CT_ITEMS];
  unsigned pos[CT_ITEMS];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int i = ct_index(tile, k);
    keep[k] = false;
    if (i < n) {
      p[k] = __ldg(src + i);
      keep[k] = pass_through || !((double)dist[(size_t)f * cap + i] > t);
    }
  }
  const unsigned incl_total = tile_compact_positions(keep, pos, desc + (size_t)f * tiles, tile, sm);
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    if (keep[k]) {
      out[(size_t)f * cap + pos[k]] = p[k];
      kept_idx[(size_t)f * cap + pos[k]] = ct_index(tile, k);
    }
  }
  if ((tile + 1) * CT_TILE >= n && threadIdx.x == 0) n_out[f] = (int)incl_total;
}

}  // namespace

void run_sor(const Ctx& c, const SorArgs& a) {
  const int tiles = cdiv(c.cap, CT_TILE);        // descriptor stride
  const int gtiles = cdiv(c.grid_cap, CT_TILE);  // blocks actually launched per frame
  const int chunks = cdiv(c.cap, TS_CHUNK);
  const int gchunks = cdiv(c.grid_cap, TS_CHUNK);
  KL(c, "k_sor_setup", k_sor_setup<<<cdiv(c.B, 128), 128, 0, c.stream>>>(a.n_in, a.meanK, a.thr, a.warnings, c.B));
  count_launch(c);
  run_grid_sort(c, a.in, a.in_stride, a.n_in, a.cell, a.minmax, a.gf, a.sort, a.sorted_pts, nullptr, nullptr);
  KL(c, "k_sor_knn", k_sor_knn<<<dim3(cdiv(c.grid_cap, 128), c.B), 128, 0, c.stream>>>(a.sorted_pts, a.sort.key[0], a.sort.key[1], a.sort.npass,
                                                               a.n_in, a.gf, a.meanK, a.dist, c.cap));
  KL(c, "k_sor_sums", k_sor_sums<<<dim3(gchunks, c.B), 256, 0, c.stream>>>(a.dist, a.n_in, a.meanK, a.partial, chunks, c.cap));
  KL(c, "k_sor_threshold", k_sor_threshold<<<cdiv(c.B, 128), 128, 0, c.stream>>>(a.partial, a.n_in, a.meanK, a.mul, a.thr, chunks, c.B));
  cudaMemsetAsync(a.desc, 0, (size_t)c.B * tiles * sizeof(unsigned), c.stream);
  KL(c, "k_sor_filter", k_sor_filter<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(a.in, a.in_stride, a.n_in, a.dist, a.thr, a.meanK, a.out,
                                                              a.kept_idx, a.n_out, a.desc, c.cap, tiles));
  count_launch(c, 4);
}

}  // namespace pcop
