// Shadow casting + obstacle marking on the occupancy grid (reference: traceShadow od.cpp:467-538,
// calculate_shadow_cast od.cpp:540-582, handle_shadow_casting od.cpp:584-672, the per-cluster loop and the marking
// loop of cloud_cb od.cpp:817-833).
//
//   k_occ_shadow   one block per cluster: members into the sensor frame, the four extrema by a (value, position)
//                  reduction that keeps the sequential loop's "first occurrence wins" rule, the shadow geometry by
//                  one thread (double arithmetic, deterministic asin / tan), then the fan of lines is dealt to the
//                  threads; every line walks its pixels sequentially (the y intercept is a running float sum).
//                  All lines store the same value, so their order does not matter.
//   k_occ_mark     every remaining point marks its cell 100 (after the shadows, as in the reference).
//
// The arithmetic choices the reference leaves open (double overloads of fabs / sqrt / asin / tan / ceil, step cap of
// the cell search, 64-bit cell indices, the degenerate-fan guard, the bounds check of the marking) are the ones
// written in words in include/pcop.h (pcop_occupancy_shadows).
#include "det_math.cuh"
#include "internal.cuh"

namespace pcop {

namespace {

constexpr int OCC_COUNT_CAP = 1 << 20;
constexpr long long SHADOW_MAX_LINE = 65536;
constexpr int SH_THREADS = 256;

// get_occupancy_grid_x_y's while-loops (od.cpp:139-147) in closed form: a floor estimate, then a short walk over the
// same float edge values fl(a +- fl((k+1)*bs)); the count stops at OCC_COUNT_CAP.
__device__ int occ_up_capped(float x_min, float bs, float x) {  // while (k < CAP && x_min + (k+1)*bs < x) k++
  if (!(x == x)) return 0;
  const float est = floorf(fdiv(fsub(x, x_min), bs));
  if (est >= (float)(OCC_COUNT_CAP + 8)) return OCC_COUNT_CAP;
  int k = (est > 2.0f) ? (int)(est - 2.0f) : 0;
  while (k > 0 && !(fadd(x_min, fmul((float)k, bs)) < x)) --k;
  while (k < OCC_COUNT_CAP && fadd(x_min, fmul((float)(k + 1), bs)) < x) ++k;
  return k;
}
__device__ int occ_down_capped(float y_max, float bs, float y) {  // while (k < CAP && y_max - (k+1)*bs > y) k++
  if (!(y == y)) return 0;
  const float est = floorf(fdiv(fsub(y_max, y), bs));
  if (est >= (float)(OCC_COUNT_CAP + 8)) return OCC_COUNT_CAP;
  int k = (est > 2.0f) ? (int)(est - 2.0f) : 0;
  while (k > 0 && !(fsub(y_max, fmul((float)k, bs)) > y)) --k;
  while (k < OCC_COUNT_CAP && fsub(y_max, fmul((float)(k + 1), bs)) > y) ++k;
  return k;
}

__device__ __forceinline__ float4 xform(const Mat34& t, const float4 p) {
  float4 o = p;
  o.x = fadd(fadd(fadd(fmul(t.m[0], p.x), fmul(t.m[1], p.y)), fmul(t.m[2], p.z)), t.m[3]);
  o.y = fadd(fadd(fadd(fmul(t.m[4], p.x), fmul(t.m[5], p.y)), fmul(t.m[6], p.z)), t.m[7]);
  o.z = fadd(fadd(fadd(fmul(t.m[8], p.x), fmul(t.m[9], p.y)), fmul(t.m[10], p.z)), t.m[11]);
  return o;
}

// static_cast<int>(double) as cvttsd2si
__device__ __forceinline__ int cvt_d2i(double v) {
  if (v != v || v >= 2147483648.0 || v <= -2147483649.0) return (int)0x80000000;
  return __double2int_rz(v);
}

// running extreme of a strict-compare loop: the first occurrence of the smallest (LESS) / largest value; values that
// compare false against everything (NaN) never win.  pos = INT_MAX: nothing seen yet.
struct Ext {
  float v;
  int pos;
};
template <bool LESS>
__device__ __forceinline__ void ext_add(Ext& e, float v, int pos) {
  if (!(v == v)) return;
  if (e.pos == 0x7fffffff || (LESS ? (v < e.v) : (v > e.v))) {
    e.v = v;
    e.pos = pos;
  }
}
template <bool LESS>
__device__ __forceinline__ void ext_merge(Ext& e, float v, int pos) {
  if (pos == 0x7fffffff) return;
  if (e.pos == 0x7fffffff || (LESS ? (v < e.v) : (v > e.v)) || (v == e.v && pos < e.pos)) {
    e.v = v;
    e.pos = pos;
  }
}
template <bool LESS>
__device__ void ext_block_reduce(Ext& e, Ext* sh /* [SH_THREADS / 32] */) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float v = __shfl_xor_sync(FULL, e.v, o);
    const int p = __shfl_xor_sync(FULL, e.pos, o);
    ext_merge<LESS>(e, v, p);
  }
  __syncthreads();  // (sh may still be read from the previous reduction)
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e;
  __syncthreads();
  e = sh[0];
  for (int w = 1; w < SH_THREADS / 32; ++w) ext_merge<LESS>(e, sh[w].v, sh[w].pos);
}

// traceShadow (od.cpp:467-538); false: the line is too long to draw
__device__ bool trace_shadow(float v1x, float v1y, float v2x, float v2y, signed char* __restrict__ grid, int W, long long size,
                             signed char opacity) {
  int x0 = cvt_f2i(v1x), x1 = cvt_f2i(v2x), y0 = cvt_f2i(v1y), y1 = cvt_f2i(v2y);
  const bool steep = llabs((long long)y1 - y0) > llabs((long long)x1 - x0);
  if (steep) {
    int t = x0; x0 = y0; y0 = t;
    t = x1; x1 = y1; y1 = t;
  }
  if (x0 > x1) {
    int t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
  }
  if ((long long)x1 - x0 + 1 > SHADOW_MAX_LINE) return false;
  const float dx = (float)(x1 - x0);
  const float dy = (float)(int)((unsigned)y1 - (unsigned)y0);
  float gradient = fdiv(dy, dx);
  if (dx == 0.0f) gradient = 1.0f;
  float iy = (float)y0;
  for (int x = x0; x <= x1; ++x) {
    const int fl = cvt_f2i(floorf(iy));
    const long long gy = steep ? x : fl, gx = steep ? fl : x;
    long long idx = gy * W + gx;
    if (idx < size && idx > -1) grid[idx] = opacity;
    idx += 1;
    if (idx < size && idx > -1) grid[idx] = opacity;
    iy = fadd(iy, gradient);
  }
  return true;
}

struct ShadowShared {
  Ext red[SH_THREADS / 32];
  int start_x, start_y, end_x, end_y;
  long long n_lines;
  int skipped;
};

__global__ void __launch_bounds__(SH_THREADS) k_occ_shadow(OccShadowArgs a) {
  const int c = blockIdx.x;
  const int o0 = a.offsets[c], o1 = a.offsets[c + 1];
  int* rec = a.records ? a.records + 6 * (size_t)c : nullptr;
  if (o1 - o0 < 2) {  // od.cpp:574
    if (rec && threadIdx.x < 6) rec[threadIdx.x] = 0;
    return;
  }
  __shared__ ShadowShared sm;
  // od.cpp:587-609: extrema of the members in the sensor frame
  Ext vmin{0.f, 0x7fffffff}, vmax{0.f, 0x7fffffff}, hmin{0.f, 0x7fffffff}, hmax{0.f, 0x7fffffff};
  for (int j = o0 + (int)threadIdx.x; j < o1; j += SH_THREADS) {
    const float4 q = xform(a.world_to_sensor, __ldg(a.cloud + a.indices[j]));
    ext_add<true>(vmin, q.x, j);
    ext_add<false>(vmax, q.x, j);
    ext_add<true>(hmin, q.y, j);
    ext_add<false>(hmax, q.y, j);
  }
  ext_block_reduce<true>(vmin, sm.red);
  ext_block_reduce<false>(vmax, sm.red);
  ext_block_reduce<true>(hmin, sm.red);
  ext_block_reduce<false>(hmax, sm.red);
  if (threadIdx.x == 0) {
    // the loop starts from member 0: a NaN there is never replaced (every compare against it is false)
    const float4 q0 = xform(a.world_to_sensor, __ldg(a.cloud + a.indices[o0]));
    const bool x0nan = !(q0.x == q0.x), y0nan = !(q0.y == q0.y);
    const float4 vmin_pt = x0nan ? q0 : xform(a.world_to_sensor, __ldg(a.cloud + a.indices[vmin.pos]));
    const float vertical_max = x0nan ? q0.x : vmax.v;
    const float horizontal_min = y0nan ? q0.y : hmin.v, horizontal_max = y0nan ? q0.y : hmax.v;
    const float width = fabsf(fsub(horizontal_max, horizontal_min));  // od.cpp:616
    // calculate_shadow_cast (od.cpp:540-582)
    const float sa = vmin_pt.z;
    const float sb = fabsf(vmin_pt.x);
    const float sc = (float)__dsqrt_rn((double)fadd(fmul(sa, sa), fmul(sb, sb)));
    const float se = (float)dadd(dsub(fabs((double)vertical_max), fabs((double)vmin_pt.x)), 0.04);
    const float D = (float)det_asin((double)fdiv(sa, sc));
    const float d = (float)dadd(dmul(det_tan((double)D), (double)se), 0.25);
    const float v_len = (float)__dsqrt_rn(
        (double)fadd(fadd(fmul(vmin_pt.x, vmin_pt.x), fmul(vmin_pt.y, vmin_pt.y)), fmul(vmin_pt.z, vmin_pt.z)));
    float4 end = vmin_pt;
    end.x = fadd(fmul(fdiv(vmin_pt.x, v_len), d), vmin_pt.x);
    end.y = fadd(fmul(fdiv(vmin_pt.y, v_len), d), vmin_pt.y);
    end.z = fadd(fmul(fdiv(vmin_pt.z, v_len), d), vmin_pt.z);
    const float4 world_end = xform(a.sensor_to_world, end);
    int end_x = occ_up_capped(a.y_min, a.block_size, world_end.y);     // od.cpp:569: (x, y) := (point.y, point.x)
    const int end_y = occ_down_capped(a.x_max, a.block_size, world_end.x);
    const float4 world_start = xform(a.sensor_to_world, vmin_pt);      // od.cpp:634-637
    int start_x = occ_up_capped(a.y_min, a.block_size, world_start.y);
    const int start_y = occ_down_capped(a.x_max, a.block_size, world_start.x);
    // od.cpp:642-643: first += ceil((width / block_size) / 2)   (int += double)
    const float wb = fdiv(width, a.block_size);
    const double shift = ceil((double)fdiv(wb, 2.0f));
    start_x = cvt_d2i(dadd((double)start_x, shift));
    end_x = cvt_d2i(dadd((double)end_x, shift));
    // od.cpp:645: for (int i = 0; i < ceil(width / block_size) + 3; i++)
    const double lim = dadd(ceil((double)wb), 3.0);
    long long n_lines = 0;
    if (lim == lim && lim > 0.0) n_lines = (lim > 1.0e9) ? 1000000000ll : (long long)ceil(lim);
    int skipped = 0;
    if (n_lines > SHADOW_MAX_LINE) {
      skipped = 1;
      n_lines = 0;
    }
    sm.start_x = start_x;
    sm.start_y = start_y;
    sm.end_x = end_x;
    sm.end_y = end_y;
    sm.n_lines = n_lines;
    sm.skipped = skipped;
  }
  __syncthreads();
  const signed char opacity = (signed char)a.opacity;
  bool bad = false;
  for (long long i = threadIdx.x; i < sm.n_lines; i += SH_THREADS) {  // od.cpp:645-661
    const int sx = (int)((unsigned)sm.start_x - (unsigned)i), ex = (int)((unsigned)sm.end_x - (unsigned)i);
    if (!trace_shadow((float)sx, (float)sm.start_y, (float)ex, (float)sm.end_y, a.grid, a.W, a.size, opacity)) bad = true;
  }
  const int any_bad = __syncthreads_or(bad ? 1 : 0);
  if (threadIdx.x == 0) {
    const int skipped = (sm.skipped || any_bad) ? 1 : 0;
    if (skipped) atomicOr(a.warnings, (uint32_t)PCOP_WARN_SHADOW_DEGENERATE);
    if (rec) {
      rec[0] = sm.start_x;
      rec[1] = sm.start_y;
      rec[2] = sm.end_x;
      rec[3] = sm.end_y;
      rec[4] = (int)sm.n_lines;
      rec[5] = skipped;
    }
  }
}

// od.cpp:823-833: every remaining point with a non-NaN x marks its cell (bounds-checked like od.cpp:205)
__global__ void __launch_bounds__(256)
    k_occ_mark(const float4* __restrict__ cloud, int n, float y_min, float x_max, float bs, int W, long long size,
               signed char* __restrict__ grid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(cloud + i);
  if (p.x != p.x) return;
  const int xc = occ_up_capped(y_min, bs, p.y);
  const int yc = occ_down_capped(x_max, bs, p.x);
  const long long idx = (long long)yc * W + xc;
  if (idx < size) grid[idx] = 100;
}

}  // namespace

void run_occ_shadows(const Ctx& c, const OccShadowArgs& a) {
  if (a.n_clusters > 0) {
    KL(c, "k_occ_shadow", k_occ_shadow<<<a.n_clusters, SH_THREADS, 0, c.stream>>>(a));
    count_launch(c);
  }
  if (a.n > 0) {
    KL(c, "k_occ_mark", k_occ_mark<<<cdiv(a.n, 256), 256, 0, c.stream>>>(a.cloud, a.n, a.y_min, a.x_max, a.block_size, a.W,
                                                                          a.size, a.grid));
    count_launch(c);
  }
}

}  // namespace pcop
