// Batched, stable, least-significant-digit radix sort of (key32, val32) pairs, onesweep style:
// one histogram kernel for all digits, then one kernel per 8-bit digit in which every 2048-key
// tile ranks its keys (warp match_any multi-split), learns its per-digit global offset by
// decoupled look-back over the earlier tiles of the same frame, stages the tile in shared
// memory in sorted order and writes it out in coalesced runs.
//
// Frames are independent segments: blockIdx.y = frame.  A frame only runs the passes its
// largest key needs (npass[f], from the atomicMax the key-producing kernel left in maxkey[f]);
// its result sits in buffer npass[f] & 1.
//
// Used by VoxelGrid (voxel keys), the ECE hash grid (cell keys), the canonical
// cluster ordering (size desc) and the CSR build (cluster rank).
#include "internal.cuh"
#include "primitives.cuh"

namespace pcop {

namespace {

__global__ void k_sort_reset(uint32_t* maxkey, int B) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < B) maxkey[f] = 0u;
}

// npass from maxkey; zero this frame's histograms
__global__ void k_sort_setup(const uint32_t* __restrict__ maxkey, int* __restrict__ npass, uint32_t* __restrict__ hist,
                             int B) {
  const int f = blockIdx.x;
  if (threadIdx.x == 0) {
    const uint32_t mk = maxkey[f];
    const int bits = 32 - __clz((int)(mk | 1u));
    npass[f] = max(1, cdiv(bits, RS_RADIX_BITS));
  }
  for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_BINS; i += blockDim.x)
    hist[(size_t)f * RS_MAX_PASSES * RS_BINS + i] = 0u;
}

__global__ void __launch_bounds__(RS_THREADS) k_sort_hist(const uint32_t* __restrict__ keys, const int* __restrict__ count,
                                                           const int* __restrict__ npass, uint32_t* __restrict__ hist,
                                                           int cap) {
  const int f = blockIdx.y;
  const int n = count[f];
  const int base = blockIdx.x * RS_TILE;
  if (base >= n) return;
  const int np = npass[f];
  __shared__ uint32_t sh[RS_MAX_PASSES * RS_BINS];
  for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_BINS; i += RS_THREADS) sh[i] = 0u;
  __syncthreads();
  const uint32_t* k = keys + (size_t)f * cap;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int i = base + it * RS_THREADS + threadIdx.x;
    if (i < n) {
      const uint32_t key = k[i];
      for (int p = 0; p < np; ++p) atomicAdd(&sh[p * RS_BINS + ((key >> (p * RS_RADIX_BITS)) & (RS_BINS - 1))], 1u);
    }
  }
  __syncthreads();
  uint32_t* gh = hist + (size_t)f * RS_MAX_PASSES * RS_BINS;
  for (int i = threadIdx.x; i < np * RS_BINS; i += RS_THREADS) {
    const uint32_t v = sh[i];
    if (v) atomicAdd(&gh[i], v);
  }
}

// exclusive scan of each (frame, pass) histogram, in place; blockDim = 256 = RS_BINS
__global__ void __launch_bounds__(RS_BINS) k_sort_scan(uint32_t* __restrict__ hist, const int* __restrict__ npass) {
  const int f = blockIdx.y, p = blockIdx.x;
  if (p >= npass[f]) return;
  uint32_t* h = hist + ((size_t)f * RS_MAX_PASSES + p) * RS_BINS;
  __shared__ uint32_t wsum[RS_BINS / 32];
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

struct PassSmem {
  uint32_t warp_hist[RS_THREADS / 32][RS_BINS];  // per-warp digit counts -> exclusive across warps
  uint32_t tile_off[RS_BINS];                    // exclusive scan of the tile histogram
  uint32_t glob_base[RS_BINS];                   // global output slot of the tile's first key of each digit
  uint32_t skey[RS_TILE];
  uint32_t sval[RS_TILE];
  uint32_t wsum[RS_BINS / 32];
};

__global__ void __launch_bounds__(RS_THREADS, 6)
    k_sort_pass(uint32_t* key_a, uint32_t* val_a, uint32_t* key_b, uint32_t* val_b, const int* __restrict__ count,
                const int* __restrict__ npass, const uint32_t* __restrict__ bin_base, uint32_t* __restrict__ desc,
                int pass, int cap, int tiles, int iota_vals, unsigned long long* __restrict__ stats) {
  const int f = blockIdx.y;
  const int n = count[f];
  const int tile = blockIdx.x;
  const int tbase = tile * RS_TILE;
  if (tbase >= n) return;
  if (pass >= npass[f]) return;
  if (tile == 0 && threadIdx.x == 0 && stats) atomicAdd(stats, (unsigned long long)n);  // keys moved by sort passes
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PassSmem& sm = *reinterpret_cast<PassSmem*>(smem_raw);
  const int lane = lane_id(), warp = warp_id();
  const int shift = pass * RS_RADIX_BITS;
  // buffers of this pass: even passes read buffer 0
  const uint32_t* kin = ((pass & 1) ? key_b : key_a) + (size_t)f * cap;
  const uint32_t* vin = ((pass & 1) ? val_b : val_a) + (size_t)f * cap;
  uint32_t* kout = ((pass & 1) ? key_a : key_b) + (size_t)f * cap;
  uint32_t* vout = ((pass & 1) ? val_a : val_b) + (size_t)f * cap;

  for (int i = threadIdx.x; i < (RS_THREADS / 32) * RS_BINS; i += RS_THREADS) (&sm.warp_hist[0][0])[i] = 0u;

  // warp-striped: warp w owns tile elements [w*32*ITEMS, (w+1)*32*ITEMS); item k of lane l = that base + k*32 + l
  uint32_t key[RS_ITEMS];
  unsigned short rank[RS_ITEMS];
  const int wbase_idx = tbase + warp * (32 * RS_ITEMS) + lane;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int i = wbase_idx + k * 32;
    key[k] = (i < n) ? kin[i] : 0xffffffffu;
  }
  __syncthreads();
  // rank keys inside the warp, row by row => stable.  One shared atomic per distinct digit of a row (issued by the
  // lowest lane of the match group); its return value is the group's base, broadcast by shuffle.
  uint32_t* wh = sm.warp_hist[warp];
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const bool valid = (wbase_idx + k * 32) < n;
    const uint32_t d = (key[k] >> shift) & (RS_BINS - 1);
    const unsigned m = __match_any_sync(FULL, valid ? d : (RS_BINS + lane));  // invalid lanes match nobody
    const int leader = __ffs(m) - 1;
    uint32_t before = 0;
    if (valid && lane == leader) before = atomicAdd(&wh[d], (uint32_t)__popc(m));
    before = __shfl_sync(FULL, before, leader);
    rank[k] = (unsigned short)(before + __popc(m & lanemask_lt()));
  }
  __syncthreads();

  // thread d: exclusive scan of digit d over the warps, tile count, look-back
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w) {
      const uint32_t c = sm.warp_hist[w][d];
      sm.warp_hist[w][d] = run;
      run += c;
    }
    const uint32_t tile_count = run;
    // publish this tile's digit count as early as possible
    unsigned* dd = desc + (((size_t)pass * gridDim.y + f) * tiles) * RS_BINS + d;
    if (tile == 0) st_volatile_u32(dd, LB_PREFIX | tile_count);
    else st_volatile_u32(dd + (size_t)tile * RS_BINS, LB_AGG | tile_count);
    // tile-local exclusive scan over digits
    uint32_t incl = tile_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) sm.wsum[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (int w = 0; w < warp; ++w) wb += sm.wsum[w];
    sm.tile_off[d] = wb + incl - tile_count;
  }
  __syncthreads();

  // stage the tile in sorted order (values are only loaded now, straight into shared memory)
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int i = wbase_idx + k * 32;
    if (i < n) {
      const uint32_t d = (key[k] >> shift) & (RS_BINS - 1);
      const uint32_t p = sm.tile_off[d] + sm.warp_hist[warp][d] + rank[k];
      sm.skey[p] = key[k];
      sm.sval[p] = (iota_vals && pass == 0) ? (uint32_t)i : vin[i];
    }
  }
  // decoupled look-back for digit d over the earlier tiles of this frame (after the staging loads were issued)
  {
    const int d = threadIdx.x;
    unsigned* dd = desc + (((size_t)pass * gridDim.y + f) * tiles) * RS_BINS + d;
    uint32_t excl = 0;
    if (tile > 0) {
      for (int t = tile - 1; t >= 0; --t) {
        const unsigned v = lookback_wait(dd + (size_t)t * RS_BINS);
        excl += v & LB_VALUE;
        if ((v >> 30) == 2u) break;
      }
      const uint32_t tile_count = ((d + 1 < RS_BINS) ? sm.tile_off[d + 1] : (uint32_t)min(RS_TILE, n - tbase)) - sm.tile_off[d];
      st_volatile_u32(dd + (size_t)tile * RS_BINS, LB_PREFIX | (excl + tile_count));
    }
    sm.glob_base[d] = bin_base[((size_t)f * RS_MAX_PASSES + pass) * RS_BINS + d] + excl;
  }
  __syncthreads();
  const int tile_n = min(RS_TILE, n - tbase);
  for (int i = threadIdx.x; i < tile_n; i += RS_THREADS) {
    const uint32_t kk = sm.skey[i];
    const uint32_t d = (kk >> shift) & (RS_BINS - 1);
    const uint32_t g = sm.glob_base[d] + ((uint32_t)i - sm.tile_off[d]);
    kout[g] = kk;
    vout[g] = sm.sval[i];
  }
}

}  // namespace

size_t sort_desc_bytes(int B, int cap) {
  return (size_t)RS_MAX_PASSES * B * cdiv(cap, RS_TILE) * RS_BINS * sizeof(uint32_t);
}

void sort_reset_maxkey(const Ctx& c, const SortBufs& s) {
  KL(c, "k_sort_reset", k_sort_reset<<<cdiv(c.B, 256), 256, 0, c.stream>>>(s.maxkey, c.B));
  count_launch(c);
}

void radix_sort_batched(const Ctx& c, const SortBufs& s, const int* count, bool iota_vals) {
  // dynamic shared memory opt-in (per device, idempotent and cheap)
  cudaFuncSetAttribute(k_sort_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem));
  const int gtiles = cdiv(c.grid_cap, RS_TILE);
  cudaMemsetAsync(s.desc, 0, sort_desc_bytes(c.B, c.grid_cap), c.stream);  // stride = launched tiles
  KL(c, "k_sort_setup", k_sort_setup<<<c.B, 256, 0, c.stream>>>(s.maxkey, s.npass, s.hist, c.B));
  KL(c, "k_sort_hist", k_sort_hist<<<dim3(gtiles, c.B), RS_THREADS, 0, c.stream>>>(s.key[0], count, s.npass, s.hist, c.cap));
  KL(c, "k_sort_scan", k_sort_scan<<<dim3(RS_MAX_PASSES, c.B), RS_BINS, 0, c.stream>>>(s.hist, s.npass));
  count_launch(c, 3);
  for (int p = 0; p < RS_MAX_PASSES; ++p) {
    KL(c, "k_sort_pass", k_sort_pass<<<dim3(gtiles, c.B), RS_THREADS, sizeof(PassSmem), c.stream>>>(s.key[0], s.val[0], s.key[1], s.val[1], count, s.npass,
                                                               s.hist, s.desc, p, c.cap, gtiles, iota_vals ? 1 : 0, s.stats));
    count_launch(c);
  }
}

}  // namespace pcop
