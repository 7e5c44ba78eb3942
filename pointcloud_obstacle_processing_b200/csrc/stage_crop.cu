// Crop stage (reference: od.cpp:195-215, the hand-written box test inside
// build_initial_occupancy_grid_dataset).  One float4 load per point, the literal predicate
// (only x is NaN-tested, bounds inclusive), warp-ballot ranks + decoupled look-back for a
// *stable* compaction (kept points keep input order: push_back at od.cpp:214), and a fused
// min/max of the survivors for the VoxelGrid stage that follows.
#include "internal.cuh"
#include "primitives.cuh"

namespace pcop {

namespace {

__global__ void k_minmax_init(MinMax* mm, int B) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < B) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mm[f].mn[a] = ORD_POS_FLT_MAX;
      mm[f].mx[a] = ORD_NEG_FLT_MAX;
    }
  }
}

// compare-based min/max: NaN never updates (oracle voxel_setup)
struct MinMaxAcc {
  float mn[3], mx[3];
  __device__ void init() {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = 3.402823466e+38f;
      mx[a] = -3.402823466e+38f;
    }
  }
  __device__ void add(const float4 p) {
    const float v[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (v[a] < mn[a]) mn[a] = v[a];
      if (v[a] > mx[a]) mx[a] = v[a];
    }
  }
  // block reduce + one atomic pair per axis per block
  __device__ void commit(MinMax* out, float (*sh)[6]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
      }
    }
    if (lane_id() == 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        sh[warp_id()][a] = mn[a];
        sh[warp_id()][3 + a] = mx[a];
      }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      const int a = threadIdx.x;
      float v = sh[0][a];
      for (int w = 1; w < CT_THREADS / 32; ++w) v = (a < 3) ? fminf(v, sh[w][a]) : fmaxf(v, sh[w][a]);
      if (a < 3) {
        atomicMin(&out->mn[a], f2ord(v));
      } else {
        atomicMax(&out->mx[a - 3], f2ord(v));
      }
    }
  }
};

__global__ void __launch_bounds__(CT_THREADS)
    k_crop(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in, float4* __restrict__ out,
           int* __restrict__ kept_idx, int* __restrict__ n_out, MinMax* __restrict__ minmax, unsigned* __restrict__ desc,
           int cap, int tiles, float x_min, float x_max, float y_min, float y_max, float z_min, float z_max) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * BT_TILE >= n) {
    if (tile == 0 && threadIdx.x == 0) n_out[f] = 0;
    return;
  }
  __shared__ CompactSmem sm;
  __shared__ float shmm[CT_THREADS / 32][6];
  const float4* src = in + (size_t)f * in_stride;
  unsigned keepmask = 0u;
  MinMaxAcc acc;
  acc.init();
  // pass 1: 16 independent 16-byte loads per thread, predicate only; nothing is held across the look-back
#pragma unroll
  for (int k = 0; k < BT_ITEMS; ++k) {
    const int i = bt_index<BT_ITEMS>(tile, k);
    if (i < n) {
      const float4 p = __ldg(src + i);
      // od.cpp:197-199, literal: drop iff isnan(x) || x<x_min || x>x_max || z<z_min || z>z_max || y<y_min || y>y_max
      const bool drop = (p.x != p.x) || p.x < x_min || p.x > x_max || p.z < z_min || p.z > z_max || p.y < y_min ||
                        p.y > y_max;
      if (!drop) {
        keepmask |= 1u << k;
        acc.add(p);
      }
    }
  }
  unsigned wbase;
  const unsigned incl_total = big_tile_scan<BT_ITEMS>(keepmask, desc + (size_t)f * tiles, tile, sm, wbase);
  float4* dst = out + (size_t)f * cap;
  int* kdst = kept_idx + (size_t)f * cap;
  // pass 2: survivors are re-read (L1/L2 hits) and written to their stable slots
#pragma unroll
  for (int k = 0; k < BT_ITEMS; ++k) {
    const bool keep = (keepmask >> k) & 1u;
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) {
      const int i = bt_index<BT_ITEMS>(tile, k);
      const unsigned pos = wbase + __popc(m & lanemask_lt());
      dst[pos] = __ldg(src + i);
      kdst[pos] = i;
    }
    wbase += __popc(m);
  }
  if ((tile + 1) * BT_TILE >= n && threadIdx.x == 0) n_out[f] = (int)incl_total;
  acc.commit(minmax + f, shmm);
}

__global__ void __launch_bounds__(CT_THREADS)
    k_minmax(const float4* __restrict__ pts, size_t stride, const int* __restrict__ n_in, MinMax* __restrict__ minmax,
             int finite_only) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * CT_TILE >= n) return;
  __shared__ float shmm[CT_THREADS / 32][6];
  const float4* src = pts + (size_t)f * stride;
  MinMaxAcc acc;
  acc.init();
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int i = ct_index(tile, k);
    if (i < n) {
      float4 p = __ldg(src + i);
      if (finite_only) {  // drop +-inf (NaN is ignored by the compares anyway)
        const float big = 3.0e38f;
        if (fabsf(p.x) > big) p.x = __int_as_float(0x7fc00000);
        if (fabsf(p.y) > big) p.y = __int_as_float(0x7fc00000);
        if (fabsf(p.z) > big) p.z = __int_as_float(0x7fc00000);
      }
      acc.add(p);
    }
  }
  acc.commit(minmax + f, shmm);
}

}  // namespace

void run_crop(const Ctx& c, const CropArgs& a) {
  const int tiles = cdiv(c.cap, BT_TILE);        // descriptor stride
  const int gtiles = cdiv(c.grid_cap, BT_TILE);  // blocks actually launched per frame
  cudaMemsetAsync(a.desc, 0, (size_t)c.B * tiles * sizeof(unsigned), c.stream);
  KL(c, "k_minmax_init", k_minmax_init<<<cdiv(c.B, 256), 256, 0, c.stream>>>(a.minmax, c.B));
  KL(c, "k_crop", k_crop<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(a.in, a.in_stride, a.n_in, a.out, a.kept_idx, a.n_out, a.minmax,
                                                        a.desc, c.cap, tiles, a.lim[0], a.lim[1], a.lim[2], a.lim[3],
                                                        a.lim[4], a.lim[5]));
  count_launch(c, 2);
}

void run_minmax(const Ctx& c, const float4* pts, size_t stride, const int* n, MinMax* minmax, bool finite_only) {
  const int gtiles = cdiv(c.grid_cap, CT_TILE);
  KL(c, "k_minmax_init", k_minmax_init<<<cdiv(c.B, 256), 256, 0, c.stream>>>(minmax, c.B));
  KL(c, "k_minmax", k_minmax<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(pts, stride, n, minmax, finite_only ? 1 : 0));
  count_launch(c, 2);
}

}  // namespace pcop
