// Fused crop + VoxelGrid fast path (reference: the crop of od.cpp:195-215 followed by downsample_cloud
// od.cpp:271-296 -> pcl::VoxelGrid::applyFilter).
//
// When the crop box is enabled its limits bound the voxel coordinates of every surviving point, so the sort key
// does not have to wait for the min/max of the cropped cloud: with F_a = (int)floorf(p.a * inv) (PCL's float
// multiply + floor) the composite key  (F_z - B_z) * ny * nx + (F_y - B_y) * nx + (F_x - B_x),  B_a =
// (int)floorf(lo_a * inv), orders the points exactly like PCL's  ijk0 + ijk1*div_b0 + ijk2*div_b0*div_b1
// (both are the lexicographic order of (F_z, F_y, F_x); ijk_a = F_a - min_b_a exactly because all values are
// integers below 2^24).  That removes the cropped-cloud round trip through HBM:
//
//   k_vf_crop_key   one read of the input: crop predicate (literal, od.cpp:197-199), composite key, STABLE compaction
//                   of (key, original index) pairs (warp ballots + decoupled look-back), min/max of the survivors
//                   (PCL's key arithmetic and overflow guard need them), all digit histograms of the sort;
//   k_vf_scan       exclusive scan of the digit histograms;
//   k_vf_sort_pass  onesweep LSD radix passes with ceil(bits / npass)-bit digits (<= 8 bits by default; 64-bit
//                   (key, index) elements; warp ranks by MATCH.ANY on the coarse digits, per-bit ballots on the
//                   high-entropy low digits);
//   k_vf_reduce     run heads + voxel count + centroids in one kernel: the tile's points are gathered into shared
//                   memory in sorted order, every run head sums its run sequentially (ascending original index:
//                   the compaction and the sort are stable) and divides by the float count; PCL's own key of the
//                   voxel is computed from the head point.
//
// The path is taken only if the host can prove PCL's int32 overflow guard cannot fire inside the crop box.  A
// surviving point with a NaN y or z (kept by the reference's predicate: only x is NaN-tested) would get a key that
// depends on the cloud's min/max; such a frame raises an internal flag and the whole wave is redone by the generic
// path (pcop_api.cu), so the result is always the reference's.
#include <cmath>
#include <cstdlib>

#include "internal.cuh"
#include "primitives.cuh"

namespace pcop {

namespace {

#ifndef VF_SORT_ITEMS
#define VF_SORT_ITEMS 16  // 4096-element tiles, 4 blocks per SM: -4.5 % against 2048-element tiles at 5 blocks per SM
#endif
constexpr int VF_ITEMS = VF_SORT_ITEMS;         // elements per thread of a sort tile
constexpr int VF_TILE = RS_THREADS * VF_ITEMS;  // elements per sort tile
constexpr int VF_MAX_BITS = 9;
constexpr int VF_MAX_PASSES = 6;  // (digit-width experiments: 6 x 5 bits)
constexpr int VF_MAX_BINS = 1 << VF_MAX_BITS;

struct VfSmemA {
  CompactSmem cs;
  float shmm[CT_THREADS / 32][6];
  uint32_t hist[VF_MAX_PASSES * VF_MAX_BINS];
};

// compare-based min/max: NaN never updates (oracle voxel_setup)
struct VfMinMax {
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
};

__global__ void k_vf_init(MinMax* mm, uint32_t* __restrict__ flags, uint32_t* __restrict__ warnings, int B) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < B) {
    if (warnings) warnings[f] = 0u;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mm[f].mn[a] = ORD_POS_FLT_MAX;
      mm[f].mx[a] = ORD_NEG_FLT_MAX;
    }
    flags[f] = 0u;
  }
}

template <bool NEED_MINMAX, int MINB>
__global__ void __launch_bounds__(CT_THREADS, MINB)
    k_vf_crop_key(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_in, VoxFusedPlan pl,
                  unsigned long long* __restrict__ pairs, int* __restrict__ n_out,
                  MinMax* __restrict__ minmax, uint32_t* __restrict__ hist, uint32_t* __restrict__ flags,
                  unsigned* __restrict__ desc, int cap, int tiles) {
  const int f = blockIdx.x, tile = blockIdx.y;  // frame-major dispatch: see run_voxel_fused
  const int n = n_in[f];
  if (tile * BT_TILE >= n) {
    if (tile == 0 && threadIdx.x == 0) n_out[f] = 0;
    return;
  }
  __shared__ VfSmemA sm;
  const int nbins = 1 << pl.digit_bits;
  for (int i = threadIdx.x; i < pl.npass * nbins; i += CT_THREADS) sm.hist[i] = 0u;
  __syncthreads();
  const float4* src = in + (size_t)f * in_stride;
  unsigned keepmask = 0u;
  uint32_t key[BT_ITEMS];
  VfMinMax acc;
  acc.init();
  bool odd = false;  // a survivor whose y or z is not finite
#pragma unroll
  for (int k = 0; k < BT_ITEMS; ++k) {
    const int i = bt_index<BT_ITEMS>(tile, k);
    key[k] = 0u;
    if (i < n) {
      const float4 p = __ldg(src + i);
      // od.cpp:197-199, literal: drop iff isnan(x) || x<x_min || x>x_max || z<z_min || z>z_max || y<y_min || y>y_max
      const bool drop = (p.x != p.x) || p.x < pl.lim[0] || p.x > pl.lim[1] || p.z < pl.lim[4] || p.z > pl.lim[5] ||
                        p.y < pl.lim[2] || p.y > pl.lim[3];
      if (!drop) {
        keepmask |= 1u << k;
        if (NEED_MINMAX) acc.add(p);
        odd = odd || !(fabsf(p.y) <= 3.0e38f) || !(fabsf(p.z) <= 3.0e38f);
        const int cx = __float2int_rz(floorf(fmul(p.x, pl.inv))) - pl.b0[0];
        const int cy = __float2int_rz(floorf(fmul(p.y, pl.inv))) - pl.b0[1];
        const int cz = __float2int_rz(floorf(fmul(p.z, pl.inv))) - pl.b0[2];
        const uint32_t kk = (uint32_t)cx + pl.nx * ((uint32_t)cy + pl.ny * (uint32_t)cz);
        key[k] = kk;
        for (int q = 0; q < pl.npass; ++q) atomicAdd(&sm.hist[q * nbins + ((kk >> (q * pl.digit_bits)) & (nbins - 1))], 1u);
      }
    }
  }
  if (__any_sync(FULL, odd) && lane_id() == 0) atomicOr(&flags[f], 1u);
  unsigned wbase;
  const unsigned incl_total = big_tile_scan<BT_ITEMS>(keepmask, desc + (size_t)f * tiles, tile, sm.cs, wbase);
  unsigned long long* pd = pairs + (size_t)f * cap;  // (key << 32) | original index
#pragma unroll
  for (int k = 0; k < BT_ITEMS; ++k) {
    const bool keep = (keepmask >> k) & 1u;
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) {
      const unsigned pos = wbase + __popc(m & lanemask_lt());
      pd[pos] = ((unsigned long long)key[k] << 32) | (unsigned long long)(uint32_t)bt_index<BT_ITEMS>(tile, k);
    }
    wbase += __popc(m);
  }
  if ((tile + 1) * BT_TILE >= n && threadIdx.x == 0) n_out[f] = (int)incl_total;
  // min/max (only PCL's own voxel keys need them: PCOP_OUT_VOXEL): block reduce + one atomic per axis per block
  if (NEED_MINMAX) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        acc.mn[a] = fminf(acc.mn[a], __shfl_xor_sync(FULL, acc.mn[a], o));
        acc.mx[a] = fmaxf(acc.mx[a], __shfl_xor_sync(FULL, acc.mx[a], o));
      }
    }
    if (lane_id() == 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        sm.shmm[warp_id()][a] = acc.mn[a];
        sm.shmm[warp_id()][3 + a] = acc.mx[a];
      }
    }
  }
  __syncthreads();  // (also: all histogram atomics of the block are done)
  if (NEED_MINMAX && threadIdx.x < 6) {
    const int a = threadIdx.x;
    float v = sm.shmm[0][a];
    for (int w = 1; w < CT_THREADS / 32; ++w) v = (a < 3) ? fminf(v, sm.shmm[w][a]) : fmaxf(v, sm.shmm[w][a]);
    if (a < 3) atomicMin(&minmax[f].mn[a], f2ord(v));
    else atomicMax(&minmax[f].mx[a - 3], f2ord(v));
  }
  uint32_t* gh = hist + (size_t)f * VF_MAX_PASSES * VF_MAX_BINS;
  for (int i = threadIdx.x; i < pl.npass * nbins; i += CT_THREADS) {
    const uint32_t v = sm.hist[i];
    if (v) atomicAdd(&gh[(i / nbins) * VF_MAX_BINS + (i % nbins)], v);
  }
}

// exclusive scan of each (frame, pass) histogram, in place; 512 threads = VF_MAX_BINS
__device__ void vf_setup_one(const MinMax* __restrict__ minmax, float leaf, VoxelFrame* __restrict__ vf, int f);

// (block (0, f) also derives PCL's voxel frame of frame f from the min/max when the keys are an output)
__global__ void __launch_bounds__(VF_MAX_BINS) k_vf_scan(uint32_t* __restrict__ hist, const MinMax* __restrict__ minmax,
                                                         float leaf, VoxelFrame* __restrict__ vf, int want_keys) {
  const int f = blockIdx.y, p = blockIdx.x;
  if (want_keys && p == 0 && threadIdx.x == 0) vf_setup_one(minmax, leaf, vf, f);
  uint32_t* h = hist + ((size_t)f * VF_MAX_PASSES + p) * VF_MAX_BINS;
  __shared__ uint32_t wsum[VF_MAX_BINS / 32];
  const uint32_t v = h[threadIdx.x];
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(FULL, incl, o);
    if (lane_id() >= o) incl += up;
  }
  if (lane_id() == 31) wsum[warp_id()] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (int w = 0; w < warp_id(); ++w) wbase += wsum[w];
  h[threadIdx.x] = wbase + incl - v;
}

template <int BITS>
struct VfPassSmem {
  static constexpr int BINS = 1 << BITS;
  uint32_t warp_hist[RS_THREADS / 32][BINS];  // per-warp digit counts -> exclusive across warps
  uint32_t tile_off[BINS + 1];                // exclusive scan of the tile histogram
  uint32_t glob_base[BINS];                   // global output slot minus staged position, per digit
  unsigned long long spair[VF_TILE];
  uint32_t wsum[RS_THREADS / 32];
};

// One LSD pass over 2048-key tiles; same scheme as radix_sort.cu's k_sort_pass (warp match_any multi-split ranks,
// per-digit decoupled look-back over the tiles of the frame, tile staged in shared memory in sorted order,
// coalesced run writes), digit width as a template parameter, pass count uniform over the frames.
#ifndef VF_SORT_MINBLOCKS
#define VF_SORT_MINBLOCKS 4
#endif
template <int BITS, bool USE_MATCH, int MINB = VF_SORT_MINBLOCKS>
__global__ void __launch_bounds__(RS_THREADS, MINB)
    k_vf_sort_pass(const unsigned long long* __restrict__ pair_in, unsigned long long* __restrict__ pair_out,
                   const int* __restrict__ count, const uint32_t* __restrict__ bin_base,
                   uint32_t* __restrict__ desc, int pass, int shift, int cap, int tiles,
                   unsigned long long* __restrict__ stats) {
  constexpr int BINS = 1 << BITS;
  constexpr int BPT = (BINS + RS_THREADS - 1) / RS_THREADS;  // bins per thread in the per-digit steps
  const int f = blockIdx.x;
  const int n = count[f];
  const int tile = blockIdx.y;
  const int tbase = tile * VF_TILE;
  if (tbase >= n) return;
  if (tile == 0 && threadIdx.x == 0 && stats) atomicAdd(stats, (unsigned long long)n);  // keys moved by sort passes
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VfPassSmem<BITS>& sm = *reinterpret_cast<VfPassSmem<BITS>*>(smem_raw);
  const int lane = lane_id(), warp = warp_id();
  const unsigned long long* pin = pair_in + (size_t)f * cap;
  unsigned long long* pout = pair_out + (size_t)f * cap;

  for (int i = threadIdx.x; i < (RS_THREADS / 32) * BINS; i += RS_THREADS) (&sm.warp_hist[0][0])[i] = 0u;

  unsigned long long pr[VF_ITEMS];  // (key << 32) | index: one 8-byte load / store per element
  unsigned short rank[VF_ITEMS];
  const int wbase_idx = tbase + warp * (32 * VF_ITEMS) + lane;
#pragma unroll
  for (int k = 0; k < VF_ITEMS; ++k) {
    const int i = wbase_idx + k * 32;
    pr[k] = (i < n) ? pin[i] : ~0ull;
  }
  auto key_of = [&](int k) -> uint32_t { return (uint32_t)(pr[k] >> 32); };
  __syncthreads();
  // rank keys inside the warp, row by row => stable.  All match masks first (independent, so their latencies
  // overlap), then one shared atomic per distinct digit of a row (issued by the lowest lane of the match group);
  // its return value is the group's base, broadcast by shuffle.
  uint32_t* wh = sm.warp_hist[warp];
#pragma unroll
  for (int k0 = 0; k0 < VF_ITEMS; k0 += 8) {  // 8 rows at a time: their match masks first, then the atomics
    unsigned mm[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k = k0 + kk;
      const bool valid = (wbase_idx + k * 32) < n;
      const uint32_t d = (key_of(k) >> shift) & (BINS - 1);
      if (USE_MATCH) {
        mm[kk] = __match_any_sync(FULL, valid ? d : (BINS + lane));  // invalid lanes match nobody
      } else {
        // MATCH.ANY runs on the ADU pipe in time proportional to the number of distinct values in the warp (80 %
        // pipe utilisation on the high-entropy low digits): one ballot per digit bit instead
        unsigned peers = __ballot_sync(FULL, valid);
#pragma unroll
        for (int b = 0; b < BITS; ++b) {
          // s = all ones if bit b of the digit is set, else 0: peers &= (bit ? bal : ~bal) is ONE three-input logic
          // op, peers & ~(bal ^ s)  (the select form cost a complement + a select + an and per bit)
          const int s = (int)(d << (31 - b)) >> 31;
          const unsigned bal = __ballot_sync(FULL, s < 0);
          peers &= ~(bal ^ (unsigned)s);
        }
        mm[kk] = valid ? peers : (1u << lane);
      }
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k = k0 + kk;
      const bool valid = (wbase_idx + k * 32) < n;
      const uint32_t d = (key_of(k) >> shift) & (BINS - 1);
      const unsigned m = mm[kk];
      const int leader = __ffs(m) - 1;
      uint32_t before = 0;
      if (valid && lane == leader) before = atomicAdd(&wh[d], (uint32_t)__popc(m));
      before = __shfl_sync(FULL, before, leader);
      rank[k] = (unsigned short)(before + __popc(m & lanemask_lt()));
    }
  }
  __syncthreads();

  unsigned* dd_frame = desc + (((size_t)pass * gridDim.x + f) * tiles) * BINS;
  // per digit: exclusive scan over the warps, tile count, early publish
  uint32_t tile_count[BPT];
  {
    uint32_t thread_sum = 0;
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
      const int d = threadIdx.x * BPT + b;  // consecutive digits per thread => the tile scan is a plain thread scan
      uint32_t run = 0;
      if (d < BINS) {
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) {
          const uint32_t c = sm.warp_hist[w][d];
          sm.warp_hist[w][d] = run;
          run += c;
        }
        if (tile == 0) st_volatile_u32(dd_frame + d, LB_PREFIX | run);
        else st_volatile_u32(dd_frame + (size_t)tile * BINS + d, LB_AGG | run);
      }
      tile_count[b] = run;
      thread_sum += run;
    }
    uint32_t incl = thread_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) sm.wsum[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (int w = 0; w < warp; ++w) wb += sm.wsum[w];
    uint32_t run = wb + incl - thread_sum;
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
      const int d = threadIdx.x * BPT + b;
      if (d < BINS) sm.tile_off[d] = run;
      run += tile_count[b];
    }
    if (threadIdx.x == RS_THREADS - 1) sm.tile_off[BINS] = run;
  }
  __syncthreads();

  // stage the tile in sorted order
#pragma unroll
  for (int k = 0; k < VF_ITEMS; ++k) {
    const int i = wbase_idx + k * 32;
    if (i < n) {
      const uint32_t d = (key_of(k) >> shift) & (BINS - 1);
      const uint32_t p = sm.tile_off[d] + sm.warp_hist[warp][d] + rank[k];
      sm.spair[p] = pr[k];
    }
  }
  // decoupled look-back per digit over the earlier tiles of this frame
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const int d = threadIdx.x * BPT + b;
    if (d < BINS) {
      uint32_t excl = 0;
      if (tile > 0) {
        unsigned* dd = dd_frame + d;
        for (int t = tile - 1; t >= 0; --t) {
          const unsigned v = lookback_wait(dd + (size_t)t * BINS);
          excl += v & LB_VALUE;
          if ((v >> 30) == 2u) break;
        }
        st_volatile_u32(dd + (size_t)tile * BINS, LB_PREFIX | (excl + tile_count[b]));
      }
      // (global slot of the digit's first key in this tile) - (its position in the staged tile)
      sm.glob_base[d] = bin_base[((size_t)f * VF_MAX_PASSES + pass) * VF_MAX_BINS + d] + excl - sm.tile_off[d];
    }
  }
  __syncthreads();
  const int tile_n = min(VF_TILE, n - tbase);
  for (int i = threadIdx.x; i < tile_n; i += RS_THREADS) {
    const unsigned long long pp = sm.spair[i];
    const uint32_t d = ((uint32_t)(pp >> 32) >> shift) & (BINS - 1);
    pout[sm.glob_base[d] + (uint32_t)i] = pp;
  }
}

// PCL's voxel frame from the min/max of the survivors (same arithmetic as stage_voxel.cu's k_voxel_setup; the host
// has proven that the overflow guard cannot fire)
__device__ void vf_setup_one(const MinMax* __restrict__ minmax, float leaf, VoxelFrame* __restrict__ vf, int f) {
  VoxelFrame v;
  v.inv = fdiv(1.0f, leaf);
  unsigned div_b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float mn = ord2f(minmax[f].mn[a]), mx = ord2f(minmax[f].mx[a]);
    v.min_b[a] = cvt_f2i(floorf(fmul(mn, v.inv)));
    const int max_b = cvt_f2i(floorf(fmul(mx, v.inv)));
    div_b[a] = (unsigned)max_b - (unsigned)v.min_b[a] + 1u;
  }
  v.overflow = 0;
  v.mul1 = div_b[0];
  v.mul2 = div_b[0] * div_b[1];
  vf[f] = v;
}

struct VfSmemR {
  CompactSmem cs;
  uint32_t skey[CT_TILE + 1];  // [0] = key before the tile
  float sx[CT_TILE], sy[CT_TILE], sz[CT_TILE];
  unsigned short head[CT_TILE];  // tile positions of the run heads, dense
};

template <bool WITH_KEYS>
__global__ void __launch_bounds__(CT_THREADS)
    k_vf_reduce(const float4* __restrict__ in, size_t in_stride, const int* __restrict__ n_sorted,
                const unsigned long long* __restrict__ pairs, const VoxelFrame* __restrict__ vf,
                float4* __restrict__ out, uint32_t* __restrict__ out_keys, int* __restrict__ n_out,
                unsigned* __restrict__ desc, int cap, int tiles, int frame_contiguous) {
  // frame_contiguous (experiment, off by default): blockIdx.x = tile, so the tiles of one frame are dispatched
  // together and the frame's input (1.9 MB) stays in L2 while its points are gathered (see run_voxel_fused)
  const int f = frame_contiguous ? blockIdx.y : blockIdx.x, tile = frame_contiguous ? blockIdx.x : blockIdx.y;
  const int m = n_sorted[f];
  const int tbase = tile * CT_TILE;
  if (tbase >= m) {
    if (tile == 0 && threadIdx.x == 0) n_out[f] = 0;
    return;
  }
  __shared__ VfSmemR sm;
  const unsigned long long* ps = pairs + (size_t)f * cap;
  const float4* src = in + (size_t)f * in_stride;
  const int tile_n = min(CT_TILE, m - tbase);
  // sorted (key, index) of the tile; the points are gathered into shared memory in sorted order.  The run heads (and
  // with them the tile's aggregate for the look-back) depend on the keys only, so the gathers are issued first and
  // stay in flight while the heads are ranked and the tile waits for its predecessors.
  unsigned long long pp[CT_ITEMS];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int t = threadIdx.x + k * CT_THREADS;
    pp[k] = (t < tile_n) ? ps[tbase + t] : 0ull;
  }
  float4 p[CT_ITEMS];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int t = threadIdx.x + k * CT_THREADS;
    p[k] = (t < tile_n) ? __ldg(src + (uint32_t)pp[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int t = threadIdx.x + k * CT_THREADS;
    if (t < tile_n) sm.skey[t + 1] = (uint32_t)(pp[k] >> 32);
  }
  if (threadIdx.x == 0) sm.skey[0] = (tbase > 0) ? (uint32_t)(ps[tbase - 1] >> 32) : 0u;
  __syncthreads();
  bool keep[CT_ITEMS];
  unsigned pos[CT_ITEMS];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int t = warp_id() * (32 * CT_ITEMS) + k * 32 + lane_id();
    keep[k] = t < tile_n && (tbase + t == 0 || sm.skey[t + 1] != sm.skey[t]);
  }
  const unsigned incl_total = tile_compact_positions(keep, pos, desc + (size_t)f * tiles, tile, sm.cs);
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const int t = threadIdx.x + k * CT_THREADS;
    if (t < tile_n) {
      sm.sx[t] = p[k].x;
      sm.sy[t] = p[k].y;
      sm.sz[t] = p[k].z;
    }
  }
  // dense list of the tile's heads: the per-head work below then runs with every lane busy (half of the sorted
  // elements are heads; walking them in place left half of each warp idle through four unrolled copies of the loop)
  const unsigned tile_excl = sm.cs.tile_excl, n_heads = sm.cs.tile_total;
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k)
    if (keep[k]) sm.head[pos[k] - tile_excl] = (unsigned short)(warp_id() * (32 * CT_ITEMS) + k * 32 + lane_id());
  __syncthreads();
  const VoxelFrame v = vf[f];
  const float fb0 = (float)v.min_b[0], fb1 = (float)v.min_b[1], fb2 = (float)v.min_b[2];
  for (unsigned h = threadIdx.x; h < n_heads; h += CT_THREADS) {
    const int t0 = sm.head[h];
    const int t1 = (h + 1 < n_heads) ? (int)sm.head[h + 1] : tile_n;  // the run inside the tile is [t0, t1)
    const uint32_t kk = sm.skey[t0 + 1];
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    for (int t = t0; t < t1; ++t) {
      ax = fadd(ax, sm.sx[t]);
      ay = fadd(ay, sm.sy[t]);
      az = fadd(az, sm.sz[t]);
    }
    int cnt = t1 - t0;
    if (t1 == tile_n) {  // ... and its tail in the following tiles
      for (int j = tbase + tile_n; j < m; ++j) {
        const unsigned long long pp = ps[j];
        if ((uint32_t)(pp >> 32) != kk) break;
        const float4 p = __ldg(src + (uint32_t)pp);
        ax = fadd(ax, p.x);
        ay = fadd(ay, p.y);
        az = fadd(az, p.z);
        ++cnt;
      }
    }
    const float c = (float)cnt;
    const size_t o = (size_t)f * cap + tile_excl + h;
    out[o] = make_float4(fdiv(ax, c), fdiv(ay, c), fdiv(az, c), 1.0f);
    if (WITH_KEYS) {  // PCL's key of this voxel, from its first point (voxel_grid.hpp: ijk = floor(p*inv) - min_b)
      const int i0 = cvt_f2i(fsub(floorf(fmul(sm.sx[t0], v.inv)), fb0));
      const int i1 = cvt_f2i(fsub(floorf(fmul(sm.sy[t0], v.inv)), fb1));
      const int i2 = cvt_f2i(fsub(floorf(fmul(sm.sz[t0], v.inv)), fb2));
      out_keys[o] = (uint32_t)i0 + (uint32_t)i1 * v.mul1 + (uint32_t)i2 * v.mul2;
    }
  }
  if (tbase + CT_TILE >= m && threadIdx.x == 0) n_out[f] = (int)incl_total;
}

template <int BITS, bool USE_MATCH, int MINB>
void launch_pass_b(const Ctx& c, const VoxelFusedArgs& a, int pass, int shift, int gtiles) {
  cudaFuncSetAttribute(k_vf_sort_pass<BITS, USE_MATCH, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)sizeof(VfPassSmem<BITS>));
  const int src = pass & 1;
  KL(c, "k_vf_sort_pass", k_vf_sort_pass<BITS, USE_MATCH, MINB><<<dim3(c.B, gtiles), RS_THREADS, sizeof(VfPassSmem<BITS>), c.stream>>>(
      a.pair[src], a.pair[src ^ 1], a.n_crop, a.sort.hist, a.sort.desc, pass, shift, c.cap, gtiles, a.sort.stats));
  count_launch(c);
}
// Blocks per SM the register allocation aims at: measured on B200 (5 x 1024 HDL-64 frames, all four passes): 3 blocks (80
// registers) 10.42 ms, 4 blocks (64 registers, kept) 10.71 ms and the same step time, 5 blocks (48 registers + spills)
// 12.09 ms, 6 blocks 13.14 ms.
template <int BITS, bool USE_MATCH>
void launch_pass_m(const Ctx& c, const VoxelFusedArgs& a, int pass, int shift, int gtiles) {
  launch_pass_b<BITS, USE_MATCH, VF_SORT_MINBLOCKS>(c, a, pass, shift, gtiles);
}

// measured (B200, HDL-64 keys, 4 x 7 bits): MATCH.ANY on the two most significant digits (coarse y / z cells: few
// distinct values per warp) and per-bit ballots on the others: 1.70 ms per 3 x 256 frames; all MATCH 1.94; all ballots 1.82
template <int BITS>
void launch_pass(const Ctx& c, const VoxelFusedArgs& a, int pass, int shift, int gtiles) {
  if (pass >= a.plan.npass - 2) launch_pass_m<BITS, true>(c, a, pass, shift, gtiles);
  else launch_pass_m<BITS, false>(c, a, pass, shift, gtiles);
}

}  // namespace

VoxFusedPlan make_vox_fused_plan(const pcop_params& p) {
  VoxFusedPlan pl{};
  pl.ok = 0;
  if (!p.enable_crop || !p.enable_voxel || !(p.downsample_size > 0.0f)) return pl;
  const float lim[6] = {p.x_min, p.x_max, p.y_min, p.y_max, p.z_min, p.z_max};
  const float inv = 1.0f / p.downsample_size;  // IEEE division, same value as the device's __fdiv_rn
  long long cells[3];
  for (int a = 0; a < 3; ++a) {
    const float lo = lim[2 * a], hi = lim[2 * a + 1];
    if (!(std::fabs(lo) <= 3.0e38f) || !(std::fabs(hi) <= 3.0e38f) || !(hi >= lo)) return pl;
    const float flo = std::floor(lo * inv), fhi = std::floor(hi * inv);
    if (!(std::fabs(flo) < 8388608.0f) || !(std::fabs(fhi) < 8388608.0f)) return pl;  // exact integer floats only
    pl.b0[a] = (int)flo;
    cells[a] = (long long)fhi - (long long)flo + 1;
    pl.lim[2 * a] = lo;
    pl.lim[2 * a + 1] = hi;
  }
  // PCL's guard: dx*dy*dz > INT32_MAX with d_a = int64((max_a - min_a) * inv) + 1 <= cells_a + 1 inside the box
  const long double prod = (long double)(cells[0] + 1) * (long double)(cells[1] + 1) * (long double)(cells[2] + 1);
  if (prod > 2147483647.0L) return pl;
  pl.inv = inv;
  pl.nx = (uint32_t)cells[0];
  pl.ny = (uint32_t)cells[1];
  pl.nz = (uint32_t)cells[2];
  const unsigned long long total = (unsigned long long)cells[0] * cells[1] * cells[2];
  int bits = 1;
  while (bits < 32 && (1ull << bits) < total) ++bits;
  pl.bits = bits;
  // measured on B200: four 7-bit passes over the 25-bit HDL-64 keys beat three 9-bit ones (a pass costs almost the
  // same for 128..512 bins per key, but the per-tile digit work doubles), so digits are capped at 8 bits by default
  const int max_digit = 8;
  pl.npass = (bits + max_digit - 1) / max_digit;
  if (pl.npass > VF_MAX_PASSES) return pl;
  pl.digit_bits = std::max(4, (bits + pl.npass - 1) / pl.npass);
  pl.ok = 1;
  return pl;
}

size_t vox_fused_hist_elems(int B) { return (size_t)B * VF_MAX_PASSES * VF_MAX_BINS; }
size_t vox_fused_desc_bytes(int B, int cap) {
  return (size_t)VF_MAX_PASSES * B * cdiv(cap, VF_TILE) * VF_MAX_BINS * sizeof(uint32_t);
}

// Launch geometry of the look-back kernels: blockIdx.x = frame, blockIdx.y = tile.  Blocks are dispatched x-fastest, so
// tile t of every frame is in flight before any tile t+1: when a tile looks back, its predecessors in the SAME frame
// were dispatched a whole row of frames earlier and have usually published their inclusive prefix already (look-back
// depth ~ resident blocks / frames instead of ~ all tiles of the frame).  Lower tiles still have lower linear block
// ids, which is what the look-back's forward-progress argument needs.
void run_voxel_fused(const Ctx& c, const VoxelFusedArgs& a) {
  const VoxFusedPlan& pl = a.plan;
  const int nbins = 1 << pl.digit_bits;
  const int btiles = cdiv(c.cap, BT_TILE), gbtiles = cdiv(c.grid_cap, BT_TILE);
  const int gtiles = cdiv(c.grid_cap, VF_TILE);
  cudaMemsetAsync(a.desc, 0, (size_t)c.B * btiles * sizeof(unsigned), c.stream);
  cudaMemsetAsync(a.sort.hist, 0, vox_fused_hist_elems(c.B) * sizeof(uint32_t), c.stream);
  cudaMemsetAsync(a.sort.desc, 0, (size_t)pl.npass * c.B * gtiles * nbins * sizeof(uint32_t), c.stream);
  KL(c, "k_vf_init", k_vf_init<<<cdiv(c.B, 256), 256, 0, c.stream>>>(a.minmax, a.flags, a.warnings, c.B));
  // Blocks per SM the register allocation aims at.  Measured on B200 (5 x 1024 HDL-64 frames): 4 blocks (64
  // registers, all 16 loads of a thread in flight) 4.58 ms, 5 blocks (46 registers) 3.97 ms, 6 blocks (32 registers)
  // 3.18 ms, 8 blocks 3.26 ms: the kernel is latency-bound (look-back wait, histogram flush), so resident warps beat
  // loads in flight per thread.
#define VF_CROP_LAUNCH(KEYS, MINB)                                                                              \
  KL(c, "k_vf_crop_key", k_vf_crop_key<KEYS, MINB><<<dim3(c.B, gbtiles), CT_THREADS, 0, c.stream>>>(            \
      a.in, a.in_stride, a.n_in, pl, a.pair[0], a.n_crop, a.minmax, a.sort.hist, a.flags, a.desc, c.cap, btiles))
  if (a.want_keys) VF_CROP_LAUNCH(true, 6);
  else VF_CROP_LAUNCH(false, 6);
#undef VF_CROP_LAUNCH
  KL(c, "k_vf_scan", k_vf_scan<<<dim3(pl.npass, c.B), VF_MAX_BINS, 0, c.stream>>>(a.sort.hist, a.minmax, a.leaf, a.vf, a.want_keys));
  count_launch(c, 3);
  for (int p = 0; p < pl.npass; ++p) {
    const int shift = p * pl.digit_bits;
    switch (pl.digit_bits) {
      case 9: launch_pass<9>(c, a, p, shift, gtiles); break;
      case 8: launch_pass<8>(c, a, p, shift, gtiles); break;
      case 7: launch_pass<7>(c, a, p, shift, gtiles); break;
      case 6: launch_pass<6>(c, a, p, shift, gtiles); break;
      case 5: launch_pass<5>(c, a, p, shift, gtiles); break;
      default: launch_pass<4>(c, a, p, shift, gtiles); break;
    }
  }
  const int fin = pl.npass & 1;
  const int tiles = cdiv(c.cap, CT_TILE), gt = cdiv(c.grid_cap, CT_TILE);
  cudaMemsetAsync(a.desc, 0, (size_t)c.B * tiles * sizeof(unsigned), c.stream);
  // (dispatching a frame's tiles together, so that its input stays in L2 for the gathers, was slower than frame-major
  // dispatch: 1.56 vs 1.13 ms per 4 x 256 frames -- the look-back chain inside a frame costs more than the second DRAM
  // fetch of a sector)
  const int contiguous = 0;
  const dim3 rgrid = contiguous ? dim3(gt, c.B) : dim3(c.B, gt);
  if (a.want_keys)
    KL(c, "k_vf_reduce", k_vf_reduce<true><<<rgrid, CT_THREADS, 0, c.stream>>>(
        a.in, a.in_stride, a.n_crop, a.pair[fin], a.vf, a.out, a.out_keys, a.n_out, a.desc, c.cap, tiles, contiguous));
  else
    KL(c, "k_vf_reduce", k_vf_reduce<false><<<rgrid, CT_THREADS, 0, c.stream>>>(
        a.in, a.in_stride, a.n_crop, a.pair[fin], a.vf, a.out, a.out_keys, a.n_out, a.desc, c.cap, tiles, contiguous));
  count_launch(c);
}

}  // namespace pcop
