// Device helpers shared by the generic (multi-kernel) and the fused shared-memory Euclidean-clustering paths.
#pragma once
#include "common.cuh"

namespace pcop {

constexpr int ECE_MAX_DIM = 1024;  // cells per axis (30-bit keys)

// Grid guarantees (DESIGN.md "search grid"), with s = fl(fl(x - mn) * inv) and at most 1024 cells per axis:
//  mode 1: cell >= tol*(1+2^-8)  =>  d2 < r2 implies the cell coordinates differ by at most 1 per axis;
//  mode 0: cell  = tol*0.5728 (< tol/sqrt(3) * (1 - 2^-7))  =>  all points of one cell are mutually within tol
//          (a clique), and d2 < r2 implies the cell coordinates differ by at most 2 per axis.
// mn/mx: component-wise min/max over the finite points of the frame.
__device__ __forceinline__ EceFrame ece_make_frame(const float* mn_in, const float* mx_in, int n, float tol, int clique) {
  EceFrame e;
  float mn[3], ext[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = mn_in[a];
    float mx = mx_in[a];
    if (n <= 0 || !(mx >= mn[a])) {
      mn[a] = 0.0f;
      mx = 0.0f;
    }
    ext[a] = mx - mn[a];
  }
  const float fine = tol * 0.5728f;
  e.mode = 1;
  if (clique && fine > 0.0f) {
    const float lim = fine * (float)(ECE_MAX_DIM - 2);
    if (ext[0] < lim && ext[1] < lim && ext[2] < lim) e.mode = 0;
  }
  const float cell_min = (e.mode == 0) ? fine : tol * 1.00390625f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float cell = (e.mode == 0) ? cell_min : fmaxf(cell_min, ext[a] / (float)(ECE_MAX_DIM - 1));
    if (!(cell > 0.0f) || !(cell < 3.0e38f)) cell = 1.0f;
    e.mn[a] = mn[a];
    e.inv[a] = 1.0f / cell;
    int d = (int)floorf(ext[a] * e.inv[a]) + 1;
    e.dim[a] = min(max(d, 1), ECE_MAX_DIM);
  }
  return e;
}

__device__ __forceinline__ int cell_coord(float x, float mn, float inv, int dim) {
  const float s = floorf((x - mn) * inv);
  int c = (s == s) ? (int)fminf(fmaxf(s, 0.0f), (float)(dim - 1)) : 0;
  return c;
}

// cell key of point i; clique mode: a point with a non-finite coordinate is within tol of nothing and gets a
// private cell (key >= 2^30)
__device__ __forceinline__ uint32_t ece_point_key(const float4 p, const EceFrame& e, int i) {
  const int cx = cell_coord(p.x, e.mn[0], e.inv[0], e.dim[0]);
  const int cy = cell_coord(p.y, e.mn[1], e.inv[1], e.dim[1]);
  const int cz = cell_coord(p.z, e.mn[2], e.inv[2], e.dim[2]);
  uint32_t key = (uint32_t)cx + (uint32_t)e.dim[0] * ((uint32_t)cy + (uint32_t)e.dim[1] * (uint32_t)cz);
  if (e.mode == 0 && !(fabsf(p.x) <= 3.0e38f && fabsf(p.y) <= 3.0e38f && fabsf(p.z) <= 3.0e38f))
    key = 0x40000000u | (uint32_t)i;
  return key;
}

// union-find with atomic-min hooking (works on global and shared memory): parents only ever decrease, roots are
// the smallest index of their set
__device__ __forceinline__ int uf_find(int* parent, int v) {
  // path halving with atomicMin: a stale read is still an ancestor
  int p = parent[v];
  while (p != v) {
    const int gp = parent[p];
    if (gp != p) atomicMin(&parent[v], gp);
    v = p;
    p = gp;
  }
  return v;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&parent[a], b);  // hook the larger root under the smaller
    if (old == a) return;
    a = old;  // a was no longer a root: merge its (former) parent with b instead
  }
}

}  // namespace pcop
