// VoxelGrid stage (reference: downsample_cloud, od.cpp:271-296 -> pcl::VoxelGrid::applyFilter,
// SURVEY 8a-2).  Integer voxel keys with PCL's exact float/int sequence, batched onesweep radix
// sort of (key, point index), run-boundary detection by stable compaction of head positions,
// then one thread per voxel sums its run sequentially in ascending original index (the stable
// sort hands that order over for free) and divides by the float count.
#include "internal.cuh"
#include "primitives.cuh"

namespace pcop {

namespace {

__global__ void k_voxel_setup(const MinMax* __restrict__ minmax, const int* __restrict__ n_in, float leaf,
                              VoxelFrame* __restrict__ vf, uint32_t* __restrict__ warnings, int B) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  VoxelFrame v;
  v.inv = fdiv(1.0f, leaf);
  float mn[3], mx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = ord2f(minmax[f].mn[a]);
    mx[a] = ord2f(minmax[f].mx[a]);
  }
  unsigned long long prod = 1ull;
  int max_b[3];
  unsigned div_b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const long long d = (long long)((unsigned long long)cvt_f2l(fmul(fsub(mx[a], mn[a]), v.inv)) + 1ull);
    prod *= (unsigned long long)d;
    v.min_b[a] = cvt_f2i(floorf(fmul(mn[a], v.inv)));
    max_b[a] = cvt_f2i(floorf(fmul(mx[a], v.inv)));
    div_b[a] = (unsigned)max_b[a] - (unsigned)v.min_b[a] + 1u;
  }
  v.overflow = ((long long)prod > 2147483647ll) ? 1 : 0;
  if (n_in[f] <= 0) v.overflow = 0;
  v.mul1 = div_b[0];
  v.mul2 = div_b[0] * div_b[1];
  vf[f] = v;
  if (v.overflow) atomicOr(&warnings[f], (uint32_t)PCOP_WARN_VOXEL_OVERFLOW_FALLBACK);
}

__global__ void __launch_bounds__(CT_THREADS)
    k_voxel_keys(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in,
                 const VoxelFrame* __restrict__ vf, uint32_t* __restrict__ keys, uint32_t* __restrict__ maxkey, int cap) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * CT_TILE >= n) return;
  const VoxelFrame v = vf[f];
  const float4* src = in + (size_t)f * in_stride;
  uint32_t* kd = keys + (size_t)f * cap;
  uint32_t mk = 0;
  const float fb0 = (float)v.min_b[0], fb1 = (float)v.min_b[1], fb2 = (float)v.min_b[2];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int i = ct_index(tile, k);
    if (i < n) {
      const float4 p = __ldg(src + i);
      uint32_t key = 0u;
      if (!v.overflow) {
        const int i0 = cvt_f2i(fsub(floorf(fmul(p.x, v.inv)), fb0));
        const int i1 = cvt_f2i(fsub(floorf(fmul(p.y, v.inv)), fb1));
        const int i2 = cvt_f2i(fsub(floorf(fmul(p.z, v.inv)), fb2));
        key = (uint32_t)i0 + (uint32_t)i1 * v.mul1 + (uint32_t)i2 * v.mul2;
      }
      kd[i] = key;
      mk = max(mk, key);
    }
  }
  mk = __reduce_max_sync(FULL, mk);
  if (lane_id() == 0 && mk) atomicMax(&maxkey[f], mk);
}

// head positions of the sorted key runs -> run_start[], V
__global__ void __launch_bounds__(CT_THREADS)
    k_voxel_heads(const uint32_t* __restrict__ key0, const uint32_t* __restrict__ key1, const int* __restrict__ npass,
                  const int* __restrict__ n_in, const VoxelFrame* __restrict__ vf, int* __restrict__ run_start,
                  int* __restrict__ n_out, unsigned* __restrict__ desc, int cap, int tiles) {
  const int f = blockIdx.y, tile = blockIdx.x;
  const int n = n_in[f];
  if (tile * BT_TILE >= n) {
    if (tile == 0 && threadIdx.x == 0) n_out[f] = 0;
    return;
  }
  __shared__ CompactSmem sm;
  const uint32_t* ks = ((npass[f] & 1) ? key1 : key0) + (size_t)f * cap;
  const bool all_heads = vf[f].overflow != 0;
  unsigned keepmask = 0u;
#pragma unroll
  for (int k = 0; k < BT_ITEMS; ++k) {
    const int j = bt_index<BT_ITEMS>(tile, k);
    if (j < n && (all_heads || j == 0 || ks[j] != ks[j - 1])) keepmask |= 1u << k;
  }
  unsigned wbase;
  const unsigned incl_total = big_tile_scan<BT_ITEMS>(keepmask, desc + (size_t)f * tiles, tile, sm, wbase);
  int* rs = run_start + (size_t)f * cap;
#pragma unroll
  for (int k = 0; k < BT_ITEMS; ++k) {
    const bool keep = (keepmask >> k) & 1u;
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) rs[wbase + __popc(m & lanemask_lt())] = bt_index<BT_ITEMS>(tile, k);
    wbase += __popc(m);
  }
  if ((tile + 1) * BT_TILE >= n && threadIdx.x == 0) n_out[f] = (int)incl_total;
}

__global__ void __launch_bounds__(256)
    k_voxel_centroid(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in,
                     const uint32_t* __restrict__ key0, const uint32_t* __restrict__ key1, const uint32_t* __restrict__ val0,
                     const uint32_t* __restrict__ val1, const int* __restrict__ npass, const VoxelFrame* __restrict__ vf,
                     const int* __restrict__ run_start, const int* __restrict__ n_vox, float4* __restrict__ out,
                     uint32_t* __restrict__ out_keys, int cap) {
  const int f = blockIdx.y;
  const int V = n_vox[f];
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int m = n_in[f];
  const int par = npass[f] & 1;
  const uint32_t* ks = (par ? key1 : key0) + (size_t)f * cap;
  const uint32_t* vs = (par ? val1 : val0) + (size_t)f * cap;
  const int* rs = run_start + (size_t)f * cap;
  const float4* src = in + (size_t)f * in_stride;
  const int j0 = rs[v];
  const int j1 = (v + 1 < V) ? rs[v + 1] : m;
  float4 o;
  if (vf[f].overflow) {  // PCL fallback: output = input, untouched
    o = __ldg(src + vs[j0]);
  } else {
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int j = j0; j < j1; ++j) {
      const float4 p = __ldg(src + vs[j]);
      sx = fadd(sx, p.x);
      sy = fadd(sy, p.y);
      sz = fadd(sz, p.z);
    }
    const float cnt = (float)(j1 - j0);
    o = make_float4(fdiv(sx, cnt), fdiv(sy, cnt), fdiv(sz, cnt), 1.0f);
  }
  out[(size_t)f * cap + v] = o;
  out_keys[(size_t)f * cap + v] = ks[j0];
}

}  // namespace

void run_voxel(const Ctx& c, const VoxelArgs& a) {
  const int tiles = cdiv(c.cap, CT_TILE);        // descriptor stride
  const int gtiles = cdiv(c.grid_cap, CT_TILE);  // blocks actually launched per frame
  KL(c, "k_voxel_setup", k_voxel_setup<<<cdiv(c.B, 128), 128, 0, c.stream>>>(a.minmax, a.n_in, a.leaf, a.vf, a.warnings, c.B));
  sort_reset_maxkey(c, a.sort);
  KL(c, "k_voxel_keys", k_voxel_keys<<<dim3(gtiles, c.B), CT_THREADS, 0, c.stream>>>(a.in, a.in_stride, a.n_in, a.vf, a.sort.key[0],
                                                              a.sort.maxkey, c.cap));
  count_launch(c, 2);
  radix_sort_batched(c, a.sort, a.n_in, /*iota_vals=*/true);
  const int btiles = cdiv(c.cap, BT_TILE), gbtiles = cdiv(c.grid_cap, BT_TILE);
  cudaMemsetAsync(a.desc, 0, (size_t)c.B * btiles * sizeof(unsigned), c.stream);
  KL(c, "k_voxel_heads", k_voxel_heads<<<dim3(gbtiles, c.B), CT_THREADS, 0, c.stream>>>(
      a.sort.key[0], a.sort.key[1], a.sort.npass, a.n_in, a.vf, a.run_start, a.n_out, a.desc, c.cap, btiles));
  KL(c, "k_voxel_centroid", k_voxel_centroid<<<dim3(cdiv(c.grid_cap, 256), c.B), 256, 0, c.stream>>>(
      a.in, a.in_stride, a.n_in, a.sort.key[0], a.sort.key[1], a.sort.val[0], a.sort.val[1], a.sort.npass, a.vf,
      a.run_start, a.n_out, a.out, a.out_keys, c.cap));
  count_launch(c, 2);
}

}  // namespace pcop
