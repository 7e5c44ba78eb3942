// Block-level primitives shared by the stage kernels:
//   * single-pass *stable* stream compaction over a frame: warp-ballot ranks inside a
//     1024-point tile + decoupled look-back across the tiles of the same frame;
//   * the canonical tree sum ("CT2048", same shape as oracle tree_sum);
//   * small warp helpers.
#pragma once
#include "common.cuh"

namespace pcop {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned* p, unsigned v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Spin until a look-back descriptor is published (status bits non-zero).  Forward progress rests on the dispatch order
// described below; a build with -DPCOP_LOOKBACK_SPIN_LIMIT=<n> turns a wait of more than n polls into a trap (a
// sticky launch error the API reports) instead of a hang -- for runs under tools that serialise or reorder blocks.
__device__ __forceinline__ unsigned lookback_wait(const unsigned* p) {
  unsigned d = ld_volatile_u32(p);
#ifdef PCOP_LOOKBACK_SPIN_LIMIT
  unsigned long long polls = 0ull;
#endif
  while ((d >> 30) == 0u) {
#ifdef PCOP_LOOKBACK_SPIN_LIMIT
    if (++polls > (unsigned long long)(PCOP_LOOKBACK_SPIN_LIMIT)) __trap();
#endif
    d = ld_volatile_u32(p);
  }
  return d;
}

// ---- decoupled look-back ------------------------------------------------------
// One 32-bit descriptor per tile: status in bits 31..30 (0 = not ready, 1 = tile aggregate,
// 2 = inclusive prefix), value in bits 29..0.  Status and value travel in one word, so no
// fence is needed between them.  Tiles of one frame are blockIdx.x = 0,1,2,... of the same
// blockIdx.y; the hardware dispatches lower linear block ids first, so a tile only ever
// waits on tiles that are already resident or finished.
constexpr unsigned LB_AGG = 1u << 30;
constexpr unsigned LB_PREFIX = 2u << 30;
constexpr unsigned LB_VALUE = (1u << 30) - 1u;

// Called by all 32 lanes of one warp.  Returns the exclusive prefix of `aggregate` over the
// preceding tiles of the frame (valid in every lane).
__device__ __forceinline__ unsigned lookback_warp(unsigned* desc_frame, int tile, unsigned aggregate) {
  const int lane = lane_id();
  if (tile == 0) {
    if (lane == 0) st_volatile_u32(desc_frame, LB_PREFIX | aggregate);
    return 0u;
  }
  if (lane == 0) st_volatile_u32(desc_frame + tile, LB_AGG | aggregate);
  unsigned exclusive = 0u;
  int look = tile - 1;
  while (true) {
    const int idx = look - lane;
    unsigned d = LB_PREFIX;  // lanes past the first tile read as "prefix 0"
    if (idx >= 0) {
      d = lookback_wait(desc_frame + idx);
    }
    const unsigned is_prefix = __ballot_sync(FULL, (d >> 30) == 2u);
    const int first = is_prefix ? (__ffs(is_prefix) - 1) : 31;  // nearest tile that already knows its inclusive prefix
    unsigned v = (lane <= first) ? (d & LB_VALUE) : 0u;
    v = __reduce_add_sync(FULL, v);
    exclusive += v;
    if (is_prefix) break;  // always true for the last group because idx<0 lanes read as prefix
    look -= 32;
  }
  if (lane == 0) st_volatile_u32(desc_frame + tile, LB_PREFIX | (exclusive + aggregate));
  return exclusive;
}

struct CompactSmem {
  unsigned warp_total[CT_THREADS / 32];
  unsigned tile_excl;
  unsigned tile_total;
};

// Stable output positions for one tile of CT_TILE points.  Arrangement is warp-striped:
// item k of lane l of warp w is tile element w*128 + k*32 + l, so loads are coalesced and the
// ballot of row k ranks 32 consecutive elements.  On return pos[k] is the frame-wide output
// slot of item k if keep[k]; the function returns the inclusive total up to this tile.
__device__ __forceinline__ unsigned tile_compact_positions(const bool (&keep)[CT_ITEMS], unsigned (&pos)[CT_ITEMS],
                                                           unsigned* desc_frame, int tile, CompactSmem& sm) {
  const int lane = lane_id(), warp = warp_id();
  unsigned row_base = 0;
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) {
    const unsigned m = __ballot_sync(FULL, keep[k]);
    pos[k] = row_base + __popc(m & lanemask_lt());
    row_base += __popc(m);
  }
  if (lane == 0) sm.warp_total[warp] = row_base;
  __syncthreads();
  if (warp == 0) {
    unsigned t = (lane < CT_THREADS / 32) ? sm.warp_total[lane] : 0u;
    // exclusive scan over the 8 warp totals
    unsigned incl = t;
#pragma unroll
    for (int o = 1; o < CT_THREADS / 32; o <<= 1) {
      const unsigned up = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += up;
    }
    const unsigned tile_total = __shfl_sync(FULL, incl, CT_THREADS / 32 - 1);
    const unsigned excl = lookback_warp(desc_frame, tile, tile_total);
    if (lane < CT_THREADS / 32) sm.warp_total[lane] = excl + incl - t;
    if (lane == 0) {
      sm.tile_excl = excl;
      sm.tile_total = tile_total;
    }
  }
  __syncthreads();
  const unsigned wbase = sm.warp_total[warp];
#pragma unroll
  for (int k = 0; k < CT_ITEMS; ++k) pos[k] += wbase;
  return sm.tile_excl + sm.tile_total;
}

// Big-tile variant for the bandwidth-heavy compactions: ITEMS (<= 32) items per thread, keep flags as a
// bit mask, so a thread keeps no payload in registers across the look-back wait (payload is re-read from
// L1/L2 afterwards).  Tile = 256*ITEMS elements, warp-striped: item k of lane l of warp w is tile element
// w*32*ITEMS + k*32 + l.  Returns the inclusive kept total up to this tile; wbase = frame-wide output slot
// of the warp's first kept item.  The caller then walks its rows:
//   m = __ballot_sync(FULL, keepmask >> k & 1); pos = wbase + popc(m & lanemask_lt()); wbase += popc(m);
template <int ITEMS>
__device__ __forceinline__ int bt_index(int tile, int k) {
  return tile * (CT_THREADS * ITEMS) + warp_id() * (32 * ITEMS) + k * 32 + lane_id();
}
template <int ITEMS>
__device__ __forceinline__ unsigned big_tile_scan(unsigned keepmask, unsigned* desc_frame, int tile, CompactSmem& sm,
                                                  unsigned& wbase) {
  const int lane = lane_id(), warp = warp_id();
  unsigned wtotal = __reduce_add_sync(FULL, (unsigned)__popc(keepmask));
  if (lane == 0) sm.warp_total[warp] = wtotal;
  __syncthreads();
  if (warp == 0) {
    unsigned t = (lane < CT_THREADS / 32) ? sm.warp_total[lane] : 0u;
    unsigned incl = t;
#pragma unroll
    for (int o = 1; o < CT_THREADS / 32; o <<= 1) {
      const unsigned up = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += up;
    }
    const unsigned tile_total = __shfl_sync(FULL, incl, CT_THREADS / 32 - 1);
    const unsigned excl = lookback_warp(desc_frame, tile, tile_total);
    if (lane < CT_THREADS / 32) sm.warp_total[lane] = excl + incl - t;
    if (lane == 0) {
      sm.tile_excl = excl;
      sm.tile_total = tile_total;
    }
  }
  __syncthreads();
  wbase = sm.warp_total[warp];
  return sm.tile_excl + sm.tile_total;
}

// element index (inside the frame) of item k of this thread in tile `tile`
__device__ __forceinline__ int ct_index(int tile, int k) {
  return tile * CT_TILE + warp_id() * (32 * CT_ITEMS) + k * 32 + lane_id();
}

// ---- canonical tree sum ---------------------------------------------------------
// Block of 256 threads, one TS_CHUNK of 2048 elements: the caller has already summed its
// 8 strided elements (element 256*r + t for r = 0..7) sequentially into `acc`.
// Returns (in thread 0) the chunk sum in the canonical order: xor-butterfly 16,8,4,2,1 inside
// each warp, then the 8 warp results added sequentially.
__device__ __forceinline__ double tree_butterfly(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = dadd(v, __shfl_xor_sync(FULL, v, off));
  return v;
}

}  // namespace pcop
