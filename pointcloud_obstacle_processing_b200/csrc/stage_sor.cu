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

// One thread per point (meanK above 31: the k + 1 best do not fit one per lane).
__global__ void __launch_bounds__(128)
    k_sor_knn_serial(const float4* __restrict__ sorted_pts, const uint32_t* __restrict__ key0, const uint32_t* __restrict__ key1,
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

__device__ __forceinline__ int sor_warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(FULL, v, o);
    if (lane_id() >= o) v += up;
  }
  return v;
}

// One WARP per point.  The cube around the point's cell grows ring by ring; the x-runs of a ring (a whole row of the
// shell, or the two end cells of an interior row) are dealt to the lanes, which locate them in the sorted key list
// with two binary searches each -- 32 searches in flight instead of one -- and the candidates of all runs are then
// scanned 32 at a time.  The k smallest squared distances live one per lane (ascending, lanes >= k hold +inf); a
// candidate below the k-th is inserted with a ballot (its position) and a shuffle (the shift).  Same candidate set,
// same float predicate, same termination rule and same summation order as the one-thread-per-point kernel this
// runs for larger calls (209 -> 50 us on the 8.7 k-point VLP-16 frame: that one walks 35 rows with 13 dependent loads
// each; it wins back once a call holds enough points to fill the GPU with threads, see run_sor).
constexpr int SOR_WARPS = 8;
__global__ void __launch_bounds__(SOR_WARPS * 32)
    k_sor_knn(const float4* __restrict__ sorted_pts, const uint32_t* __restrict__ key0, const uint32_t* __restrict__ key1,
              const int* __restrict__ npass, const int* __restrict__ n_in, const EceFrame* __restrict__ ef, int meanK,
              float* __restrict__ dist, int cap) {
  const int f = blockIdx.y;
  const int n = n_in[f];
  if (n <= meanK) return;
  const int lane = lane_id();
  const int j = blockIdx.x * SOR_WARPS + warp_id();
  if (j >= n) return;  // (whole warp)
  const uint32_t* ks = ((npass[f] & 1) ? key1 : key0) + (size_t)f * cap;
  const float4* sp = sorted_pts + (size_t)f * cap;
  const EceFrame e = ef[f];
  const float4 p = sp[j];
  const uint32_t key = ks[j];
  const int dimx = e.dim[0], dimy = e.dim[1], dimz = e.dim[2];
  const int cx = (int)(key % (uint32_t)dimx);
  const int cy = (int)((key / (uint32_t)dimx) % (uint32_t)dimy);
  const int cz = (int)(key / ((uint32_t)dimx * (uint32_t)dimy));
  const int K = meanK + 1;  // <= 32: one per lane
  const float INF = __int_as_float(0x7f800000);
  float best = INF;  // lane l: the l-th smallest squared distance so far
  int cnt = 0;       // valid entries (warp-uniform)
  float kth = INF;   // best of lane K - 1 (warp-uniform)
  // smallest real cell edge, shrunk: lower bound on the distance from p to anything outside the cube
  const float cell_lb = fminf(fminf(1.0f / e.inv[0], 1.0f / e.inv[1]), 1.0f / e.inv[2]) * 0.999f;
  const int rmax = max(max(max(cx, dimx - 1 - cx), max(cy, dimy - 1 - cy)), max(cz, dimz - 1 - cz));
  for (int r = 0; r <= rmax; ++r) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, dimz - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, dimy - 1);
    const int ny = y1 - y0 + 1;
    const int nrows = (z1 - z0 + 1) * ny;
    // run s of the ring: row s / 2, run s % 2 of that row (a shell row has one run, an interior row two)
    for (int sbase = 0; sbase < 2 * nrows; sbase += 32) {
      const int s = sbase + lane;
      int q0 = 0, q1 = 0;
      if (s < 2 * nrows) {
        const int row_i = s >> 1, sgm = s & 1;
        const int zz = z0 + row_i / ny, yy = y0 + row_i % ny;
        const bool shell_row = (abs(zz - cz) == r) || (abs(yy - cy) == r);
        const uint32_t row = (uint32_t)dimx * ((uint32_t)yy + (uint32_t)dimy * (uint32_t)zz);
        int xa = -1, xb = -2;
        if (shell_row) {
          if (sgm == 0) {
            xa = max(cx - r, 0);
            xb = min(cx + r, dimx - 1);
          }
        } else {  // interior rows only the two end cells x = cx - r, cx + r
          const int x = (sgm == 0) ? (cx - r) : (cx + r);
          if (x >= 0 && x < dimx) xa = xb = x;
        }
        if (xa <= xb) {
          q0 = lower_bound_key(ks, n, row + (uint32_t)xa);
          q1 = q0;  // (runs hold a handful of points: walking to the end costs fewer dependent loads than a second search)
          const uint32_t hi_key = row + (uint32_t)xb;
          while (q1 < n && ks[q1] <= hi_key) ++q1;
        }
      }
      // candidates of the 32 runs, flattened: lane l's run covers [excl, excl + len)
      const int len = q1 - q0;
      const int incl = sor_warp_incl_scan(len);
      const int excl = incl - len;
      const int total = __shfl_sync(FULL, incl, 31);
      for (int t0 = 0; t0 < total; t0 += 32) {
        const int t = t0 + lane;
        // the run that holds candidate t: the last lane whose excl <= t among the runs that are not empty
        int lo = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
          const int probe = lo + step;
          const int ex = __shfl_sync(FULL, excl, probe & 31);
          if (probe < 32 && ex <= t) lo = probe;
        }
        // (empty runs share their successor's excl: `lo` is the last of them, whose incl is > t only if it holds t;
        // walk is not needed because the LAST lane with excl <= t is the one whose run is non-empty or t >= total)
        const int rq0 = __shfl_sync(FULL, q0, lo), rex = __shfl_sync(FULL, excl, lo);
        float d2 = INF;
        if (t < total) {
          const float4 o = sp[rq0 + (t - rex)];
          d2 = dist2(p.x, p.y, p.z, o.x, o.y, o.z);
        }
        // insert every candidate that beats the k-th best (or fills the list), lowest lane first
        unsigned pend = __ballot_sync(FULL, t < total && (cnt < K || d2 < kth));
        while (pend) {
          const int src = __ffs(pend) - 1;
          pend &= pend - 1u;
          const float d = __shfl_sync(FULL, d2, src);
          if (cnt == K && !(d < kth)) continue;  // (the list has tightened since the ballot)
          // the serial kernel's insertion, verbatim: from the back, entries greater than d move up one place (the last
          // one drops out when the list is full); d lands behind the last entry that is not greater (NaNs included)
          const int lim = (cnt < K) ? cnt : K - 1;
          const unsigned stay = __ballot_sync(FULL, lane < lim && !(best > d));
          const int pos = stay ? (32 - __clz(stay)) : 0;
          const float up = __shfl_up_sync(FULL, best, 1);
          if (lane > pos && lane <= lim) best = up;
          if (lane == pos) best = d;
          if (lane >= K) best = INF;
          cnt = min(cnt + 1, K);
          kth = __shfl_sync(FULL, best, K - 1);
        }
      }
    }
    if (cnt == K) {
      const float reach = (float)r * cell_lb;
      if (kth <= reach * reach) break;
    }
  }
  // PCL: dist_sum += sqrt(nn_dists[k]) for k = 1..meanK (element 0 is the query), double accumulator
  double sum = 0.0;
  for (int k = 1; k < K; ++k) {
    const float b = __shfl_sync(FULL, best, k);
    sum = dadd(sum, __dsqrt_rn((double)b));
  }
  if (lane == 0) dist[(size_t)f * cap + (int)__float_as_uint(p.w)] = (float)ddiv(sum, (double)meanK);
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
  // a call of a few small frames cannot fill the GPU with one thread per point (latency: 209 us for 8.7 k points, whatever
  // the count, up to ~300 k points); the warp-per-point kernel does 3-4 x the work in a quarter of the time there
  if (a.meanK + 1 <= 32 && (long long)c.B * c.grid_cap <= 32768)
    KL(c, "k_sor_knn", k_sor_knn<<<dim3(cdiv(c.grid_cap, SOR_WARPS), c.B), SOR_WARPS * 32, 0, c.stream>>>(
                           a.sorted_pts, a.sort.key[0], a.sort.key[1], a.sort.npass, a.n_in, a.gf, a.meanK, a.dist, c.cap));
  else
    KL(c, "k_sor_knn", k_sor_knn_serial<<<dim3(cdiv(c.grid_cap, 128), c.B), 128, 0, c.stream>>>(
                           a.sorted_pts, a.sort.key[0], a.sort.key[1], a.sort.npass, a.n_in, a.gf, a.meanK, a.dist, c.cap));
  KL(c, "k_sor_sums", k_sor_sums<<<dim3(gchunks, c.B), 256, 0, c.stream>>>(a.dist, a.n_in, a.meanK, a.partial, chunks, c.cap));
  KL(c, "k_sor_threshold", k_sor_threshold<<<cdiv(c.B, 128), 128, 0, c.stream>>>(a.partial, a.n_in, a.meanK, a.mul, a.thr, chunks, c.B));
  cudaMemsetAsync(a.desc, 0, (size_t)c.B * tiles * sizeof(unsigned), c.stream);
  KL(c, "k_sor_filter", k_sor_filter<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(a.in, a.in_stride, a.n_in, a.dist, a.thr, a.meanK, a.out,
                                                              a.kept_idx, a.n_out, a.desc, c.cap, tiles));
  count_launch(c, 4);
}

}  // namespace pcop
