// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// PARITY UNPINNED (see pcop_oracle.h): CPU restatement of the PCL algorithms the
// reference node calls; od.cpp = /root/reference/minibot_cr18/src/obstacle_detection.cpp.
//
// Build: g++ -O2 -ffp-contract=off (no -ffast-math, no -march=native): every
// float expression below is evaluated in the written order, one IEEE-754
// rounding per operation, no FMA.
#include "pcop_oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <utility>
#include <vector>

#include "det_math.hpp"

namespace {

using pcop_oracle::det_atan2_ypos;
using pcop_oracle::det_cos;
using pcop_oracle::det_log;
using pcop_oracle::det_sin;

struct P4 {
  float x, y, z, w;
};

// ---------------------------------------------------------------------------
// float -> int conversions with x86 cvttss2si semantics ("integer indefinite"
// for NaN / out of range), which is what a PCL build on the reference's
// platform executes for static_cast<int>(float).
inline int32_t cvt_f2i(float v) {
  if (v != v || v >= 2147483648.0f || v < -2147483648.0f) return INT32_MIN;
  return (int32_t)v;
}
inline int64_t cvt_f2l(float v) {
  if (v != v || v >= 9223372036854775808.0f || v < -9223372036854775808.0f) return INT64_MIN;
  return (int64_t)v;
}

// squared distance, FLANN L2_Simple<float> order
inline float dist2(const P4& a, const P4& b) {
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return ((dx * dx) + (dy * dy)) + (dz * dz);
}

// ---------------------------------------------------------------------------
// Canonical tree sum ("CT2048").  The GPU library implements the same shape
// with 256-thread blocks; written here as plain loops.
//   chunk c = elements [2048c, 2048c+2048); lane t sums elements 2048c + 256r + t,
//   r = 0..7, sequentially from +0.0 (missing elements add +0.0);
//   within each group of 32 lanes a xor-butterfly (16,8,4,2,1): v[t] = v[t] + v[t^off];
//   the 8 group results are added sequentially; chunks are added sequentially.
double tree_sum(const double* e, int64_t n) {
  double total = 0.0;
  for (int64_t c0 = 0; c0 < n; c0 += 2048) {
    double lane[256];
    for (int t = 0; t < 256; ++t) {
      double acc = 0.0;
      for (int r = 0; r < 8; ++r) {
        const int64_t i = c0 + 256 * r + t;
        acc += (i < n) ? e[i] : 0.0;
      }
      lane[t] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
      double nxt[256];
      for (int t = 0; t < 256; ++t) nxt[t] = lane[t] + lane[t ^ off];
      std::memcpy(lane, nxt, sizeof(lane));
    }
    double s = lane[0];
    for (int w = 1; w < 8; ++w) s += lane[32 * w];
    total += s;
  }
  return total;
}

// ---------------------------------------------------------------------------
// a-1 crop: od.cpp:195-215, literal predicate (only x is NaN-tested).
int crop(const pcop_params& pr, const P4* in, int n, P4* out, int32_t* kept) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const P4& p = in[i];
    if (std::isnan(p.x) || p.x < pr.x_min || p.x > pr.x_max || p.z < pr.z_min || p.z > pr.z_max ||
        p.y < pr.y_min || p.y > pr.y_max)
      continue;
    if (out) out[m] = p;
    if (kept) kept[m] = i;
    ++m;
  }
  return m;
}

// ---------------------------------------------------------------------------
// a-2 VoxelGrid: od.cpp:282-285; PCL voxel_grid.hpp::applyFilter.
struct VoxelSetup {
  float inv;
  int32_t min_b[3];
  uint32_t mul1, mul2;
  bool overflow;
};

VoxelSetup voxel_setup(const P4* in, int m, float leaf) {
  VoxelSetup s;
  s.inv = 1.0f / leaf;
  // getMinMax3D on a dense cloud: compare-based min/max (ORACLE CHOICE: NaN never updates)
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  for (int i = 0; i < m; ++i) {
    const float v[3] = {in[i].x, in[i].y, in[i].z};
    for (int a = 0; a < 3; ++a) {
      if (v[a] < mn[a]) mn[a] = v[a];
      if (v[a] > mx[a]) mx[a] = v[a];
    }
  }
  uint64_t prod = 1;
  for (int a = 0; a < 3; ++a) {
    const int64_t d = (int64_t)((uint64_t)cvt_f2l((mx[a] - mn[a]) * s.inv) + 1ull);
    prod *= (uint64_t)d;  // wrapping, like the int64 product on x86
  }
  s.overflow = (int64_t)prod > (int64_t)INT32_MAX;
  int32_t max_b[3];
  uint32_t div_b[3];
  for (int a = 0; a < 3; ++a) {
    s.min_b[a] = cvt_f2i(std::floor(mn[a] * s.inv));
    max_b[a] = cvt_f2i(std::floor(mx[a] * s.inv));
    div_b[a] = (uint32_t)max_b[a] - (uint32_t)s.min_b[a] + 1u;
  }
  s.mul1 = div_b[0];
  s.mul2 = div_b[0] * div_b[1];
  return s;
}

inline uint32_t voxel_key(const VoxelSetup& s, const P4& p) {
  const int32_t i0 = cvt_f2i(std::floor(p.x * s.inv) - (float)s.min_b[0]);
  const int32_t i1 = cvt_f2i(std::floor(p.y * s.inv) - (float)s.min_b[1]);
  const int32_t i2 = cvt_f2i(std::floor(p.z * s.inv) - (float)s.min_b[2]);
  return (uint32_t)i0 + (uint32_t)i1 * s.mul1 + (uint32_t)i2 * s.mul2;
}

int voxel(const P4* in, int m, float leaf, P4* out, uint32_t* out_keys, uint32_t* warnings) {
  if (m <= 0) return 0;
  const VoxelSetup s = voxel_setup(in, m, leaf);
  if (s.overflow) {  // PCL: warn, output = input
    if (warnings) *warnings |= PCOP_WARN_VOXEL_OVERFLOW_FALLBACK;
    for (int i = 0; i < m; ++i) {
      if (out) out[i] = in[i];
      if (out_keys) out_keys[i] = 0;
    }
    return m;
  }
  std::vector<std::pair<uint32_t, int32_t>> kv(m);
  for (int i = 0; i < m; ++i) kv[i] = {voxel_key(s, in[i]), i};
  // ORACLE CHOICE: stable order (ascending original index inside a voxel); PCL's std::sort
  // leaves the within-voxel order unspecified, so this is one legal PCL outcome.
  std::stable_sort(kv.begin(), kv.end(),
                   [](const std::pair<uint32_t, int32_t>& a, const std::pair<uint32_t, int32_t>& b) {
                     return a.first < b.first;
                   });
  int v = 0;
  for (int j = 0; j < m;) {
    int e = j;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    while (e < m && kv[e].first == kv[j].first) {  // sequential float sum, ORACLE CHOICE
      const P4& p = in[kv[e].second];
      sx += p.x;
      sy += p.y;
      sz += p.z;
      ++e;
    }
    const float cnt = (float)(e - j);
    if (out) out[v] = {sx / cnt, sy / cnt, sz / cnt, 1.0f};  // true division, ORACLE CHOICE
    if (out_keys) out_keys[v] = kv[j].first;
    ++v;
    j = e;
  }
  return v;
}

// ---------------------------------------------------------------------------
// kd-tree (median split on the widest axis, leaf size 15 like FLANN's
// KDTreeSingleIndex default used by pcl::search::KdTree).  Only an acceleration
// structure: leaves apply the exact float predicate, pruning is conservative.
struct KdTree {
  struct Node {
    int lo, hi;      // range in idx
    int left, right; // children or -1
    float bmin[3], bmax[3];
  };
  const P4* pts = nullptr;
  std::vector<int> idx;
  std::vector<Node> nodes;

  void build(const P4* p, int n) {
    pts = p;
    idx.resize(n);
    std::iota(idx.begin(), idx.end(), 0);
    nodes.clear();
    nodes.reserve(n / 4 + 16);
    if (n > 0) build_rec(0, n);
  }
  int build_rec(int lo, int hi) {
    Node nd;
    nd.lo = lo;
    nd.hi = hi;
    nd.left = nd.right = -1;
    for (int a = 0; a < 3; ++a) {
      nd.bmin[a] = INFINITY;
      nd.bmax[a] = -INFINITY;
    }
    for (int j = lo; j < hi; ++j) {
      const float v[3] = {pts[idx[j]].x, pts[idx[j]].y, pts[idx[j]].z};
      for (int a = 0; a < 3; ++a) {
        if (v[a] < nd.bmin[a]) nd.bmin[a] = v[a];
        if (v[a] > nd.bmax[a]) nd.bmax[a] = v[a];
      }
    }
    const int me = (int)nodes.size();
    nodes.push_back(nd);
    if (hi - lo > 15) {
      int ax = 0;
      float ext = nd.bmax[0] - nd.bmin[0];
      for (int a = 1; a < 3; ++a)
        if (nd.bmax[a] - nd.bmin[a] > ext) {
          ext = nd.bmax[a] - nd.bmin[a];
          ax = a;
        }
      if (ext > 0.0f) {
        const int mid = (lo + hi) / 2;
        std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
          const float va = ax == 0 ? pts[a].x : (ax == 1 ? pts[a].y : pts[a].z);
          const float vb = ax == 0 ? pts[b].x : (ax == 1 ? pts[b].y : pts[b].z);
          return va < vb;
        });
        const int l = build_rec(lo, mid);
        const int r = build_rec(mid, hi);
        nodes[me].left = l;
        nodes[me].right = r;
      }
    }
    return me;
  }
  // lower bound of the squared distance from q to the node box, in double, shrunk by a
  // relative slack so pruning never removes a point the float predicate would accept.
  double box_d2(const Node& nd, const P4& q) const {
    const double v[3] = {q.x, q.y, q.z};
    double d2 = 0.0;
    for (int a = 0; a < 3; ++a) {
      double d = 0.0;
      if (v[a] < nd.bmin[a]) d = (double)nd.bmin[a] - v[a];
      else if (v[a] > nd.bmax[a]) d = v[a] - (double)nd.bmax[a];
      d2 += d * d;
    }
    return d2 * (1.0 - 1e-5);
  }
  template <class F>
  void radius(int node, const P4& q, float r2, F&& f) const {
    const Node& nd = nodes[node];
    if (box_d2(nd, q) > (double)r2) return;
    if (nd.left < 0) {
      for (int j = nd.lo; j < nd.hi; ++j)
        if (dist2(q, pts[idx[j]]) < r2) f(idx[j]);
      return;
    }
    radius(nd.left, q, r2, f);
    radius(nd.right, q, r2, f);
  }
  // k smallest squared distances (max-heap in `heap`, size <= k)
  void knn(int node, const P4& q, int k, std::vector<float>& heap) const {
    const Node& nd = nodes[node];
    if ((int)heap.size() == k && box_d2(nd, q) > (double)heap.front()) return;
    if (nd.left < 0) {
      for (int j = nd.lo; j < nd.hi; ++j) {
        const float d = dist2(q, pts[idx[j]]);
        if ((int)heap.size() < k) {
          heap.push_back(d);
          std::push_heap(heap.begin(), heap.end());
        } else if (d < heap.front()) {
          std::pop_heap(heap.begin(), heap.end());
          heap.back() = d;
          std::push_heap(heap.begin(), heap.end());
        }
      }
      return;
    }
    const double dl = box_d2(nodes[nd.left], q), dr = box_d2(nodes[nd.right], q);
    if (dl <= dr) {
      knn(nd.left, q, k, heap);
      knn(nd.right, q, k, heap);
    } else {
      knn(nd.right, q, k, heap);
      knn(nd.left, q, k, heap);
    }
  }
};

// ---------------------------------------------------------------------------
// a-3 StatisticalOutlierRemoval: od.cpp:326-330.
inline float sor_mean_dist(std::vector<float>& d2, int meanK) {
  std::sort(d2.begin(), d2.end());  // ascending; element 0 is the query itself
  double sum = 0.0;
  for (int k = 1; k <= meanK; ++k) sum += std::sqrt((double)d2[k]);  // ORACLE CHOICE: sqrt in double
  return (float)(sum / (double)meanK);
}

void sor_distances(const P4* in, int v, int meanK, std::vector<float>& dist) {
  KdTree t;
  t.build(in, v);
  dist.resize(v);
  std::vector<float> heap;
  for (int i = 0; i < v; ++i) {
    heap.clear();
    t.knn(0, in[i], meanK + 1, heap);
    dist[i] = sor_mean_dist(heap, meanK);
  }
}

double sor_threshold(const std::vector<float>& dist, double mul) {
  const int n = (int)dist.size();
  std::vector<double> a(n), b(n);
  for (int i = 0; i < n; ++i) {
    a[i] = (double)dist[i];
    b[i] = (double)(dist[i] * dist[i]);  // float*float rounded to float, then widened (PCL)
  }
  const double sum = tree_sum(a.data(), n);     // ORACLE CHOICE: canonical tree order
  const double sq_sum = tree_sum(b.data(), n);
  const double mean = sum / (double)n;
  const double variance = (sq_sum - sum * sum / (double)n) / ((double)n - 1.0);
  const double stddev = std::sqrt(variance);
  return mean + mul * stddev;
}

int sor(const pcop_params& pr, const P4* in, int v, P4* out, int32_t* kept, uint32_t* warnings, float* distances,
        double* thr_out) {
  const int meanK = pr.statistical_outlier_meanK;
  if (v <= meanK) {  // ORACLE CHOICE: PCL reads past the k-NN result here (undefined); pass through
    if (warnings && v > 0) *warnings |= PCOP_WARN_SOR_TOO_FEW_POINTS;
    for (int i = 0; i < v; ++i) {
      if (out) out[i] = in[i];
      if (kept) kept[i] = i;
      if (distances) distances[i] = 0.0f;
    }
    if (thr_out) *thr_out = 0.0;
    return v;
  }
  std::vector<float> dist;
  sor_distances(in, v, meanK, dist);
  const double thr = sor_threshold(dist, (double)pr.statistical_outlier_stdDevThres);
  if (thr_out) *thr_out = thr;
  int s = 0;
  for (int i = 0; i < v; ++i) {
    if (distances) distances[i] = dist[i];
    if ((double)dist[i] > thr) continue;
    if (out) out[s] = in[i];
    if (kept) kept[s] = i;
    ++s;
  }
  return s;
}

// ---------------------------------------------------------------------------
// a-4 RANSAC plane loop: od.cpp:364-399.
struct Mt19937 {  // boost::mt19937 == std::mt19937 algorithm, restated
  uint32_t mt[624];
  int pos;
  explicit Mt19937(uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    pos = 624;
  }
  uint32_t next() {
    if (pos >= 624) {
      for (int i = 0; i < 624; ++i) {
        const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      pos = 0;
    }
    uint32_t y = mt[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  // boost::uniform_int<>(0, INT_MAX) over a 32-bit engine: bucket size 2, never rejects
  int32_t rnd() { return (int32_t)(next() >> 1); }
};

struct PlaneModelCtx {
  const P4* pts;
  int n;
  float thr;          // (float) distance threshold; compare is float |d| < thr
  double eps_angle;   // radians (od.cpp:371 passes the int through unchanged)
  double cos_eps;     // cos(eps_angle), libm, evaluated once on the host
  double axis[3];
};

const double kHalfPi = 1.57079632679489661923;

// SampleConsensusModelPerpendicularPlane::isModelValid.  ORACLE CHOICE: the angle test
// |angle(n, axis)| folded to [0, pi/2] > eps is evaluated in the cosine domain
// (no acos): reject iff cos_angle < cos(eps); eps >= pi/2 can never reject.
bool model_valid(const PlaneModelCtx& c, const float co[4]) {
  if (!(c.eps_angle > 0.0)) return true;
  if (c.eps_angle >= kHalfPi) return true;
  const double nx = co[0], ny = co[1], nz = co[2];
  const double dot = (nx * c.axis[0] + ny * c.axis[1]) + nz * c.axis[2];
  const double nn = std::sqrt((nx * nx + ny * ny) + nz * nz);
  const double na = std::sqrt((c.axis[0] * c.axis[0] + c.axis[1] * c.axis[1]) + c.axis[2] * c.axis[2]);
  const double cosang = std::fabs(dot) / (nn * na);
  return !(cosang < c.cos_eps);
}

// ORACLE CHOICE: 4-term float dot as ((a*x + b*y) + (c*z + d)).
inline float plane_dist(const float co[4], const P4& p) {
  return std::fabs(((co[0] * p.x) + (co[1] * p.y)) + ((co[2] * p.z) + co[3]));
}

inline bool collinear_ratio_test(const P4& p0, const P4& p1, const P4& p2) {
  // (p1-p0)/(p2-p0) component-wise float; "bad" iff all three ratios equal
  const float ax = (p1.x - p0.x) / (p2.x - p0.x);
  const float ay = (p1.y - p0.y) / (p2.y - p0.y);
  const float az = (p1.z - p0.z) / (p2.z - p0.z);
  return (ax == ay) && (az == ay);
}

bool compute_model(const P4& p0, const P4& p1, const P4& p2, float co[4]) {
  if (collinear_ratio_test(p0, p1, p2)) return false;
  const float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
  const float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
  float nx = ay * bz - az * by;
  float ny = az * bx - ax * bz;
  float nz = ax * by - ay * bx;
  const float norm = std::sqrt(((nx * nx) + (ny * ny)) + (nz * nz));  // ORACLE CHOICE: order, true division
  nx = nx / norm;
  ny = ny / norm;
  nz = nz / norm;
  co[0] = nx;
  co[1] = ny;
  co[2] = nz;
  co[3] = -(((nx * p0.x) + (ny * p0.y)) + (nz * p0.z));
  return true;
}

int count_within(const PlaneModelCtx& c, const float co[4]) {
  if (!model_valid(c, co)) return 0;
  int cnt = 0;
  for (int i = 0; i < c.n; ++i)
    if (plane_dist(co, c.pts[i]) < c.thr) ++cnt;
  return cnt;
}

void select_within(const PlaneModelCtx& c, const float co[4], std::vector<int32_t>& inl) {
  inl.clear();
  if (!model_valid(c, co)) return;
  for (int i = 0; i < c.n; ++i)
    if (plane_dist(co, c.pts[i]) < c.thr) inl.push_back(i);
}

// pcl::eigen33 smallest eigenpair (common/eigen.hpp), restated in double with the
// deterministic elementary functions of det_math.hpp.
void compute_roots2(double b, double c, double r[3]) {
  r[0] = 0.0;
  double d = b * b - 4.0 * c;
  if (d < 0.0) d = 0.0;
  const double sd = std::sqrt(d);
  r[2] = 0.5 * (b + sd);
  r[1] = 0.5 * (b - sd);
}

void compute_roots(const double m[9], double r[3]) {
  const double m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
  const double c0 = m00 * m11 * m22 + 2.0 * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  const double c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  const double c2 = m00 + m11 + m22;
  if (std::fabs(c0) < 2.220446049250313e-16) {
    compute_roots2(c2, c1, r);
    return;
  }
  const double s_inv3 = 1.0 / 3.0;
  const double s_sqrt3 = std::sqrt(3.0);
  const double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0) a_over_3 = 0.0;
  const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0) q = 0.0;
  const double rho = std::sqrt(-a_over_3);
  const double theta = det_atan2_ypos(std::sqrt(-q), half_b) * s_inv3;
  const double cos_theta = det_cos(theta);
  const double sin_theta = det_sin(theta);
  r[0] = c2_over_3 + 2.0 * rho * cos_theta;
  r[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  r[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (r[0] >= r[1]) std::swap(r[0], r[1]);
  if (r[1] >= r[2]) {
    std::swap(r[1], r[2]);
    if (r[0] >= r[1]) std::swap(r[0], r[1]);
  }
  if (r[0] <= 0.0) compute_roots2(c2, c1, r);
}

inline void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

void eigen33_smallest(const double mat[9], double* eval, double evec[3]) {
  double scale = 0.0;
  for (int i = 0; i < 9; ++i) scale = std::fabs(mat[i]) > scale ? std::fabs(mat[i]) : scale;
  if (scale <= 2.2250738585072014e-308) scale = 1.0;
  double sm[9];
  for (int i = 0; i < 9; ++i) sm[i] = mat[i] / scale;
  double r[3];
  compute_roots(sm, r);
  *eval = r[0] * scale;
  sm[0] -= r[0];
  sm[4] -= r[0];
  sm[8] -= r[0];
  double v1[3], v2[3], v3[3];
  cross3(&sm[0], &sm[3], v1);
  cross3(&sm[0], &sm[6], v2);
  cross3(&sm[3], &sm[6], v3);
  const double l1 = (v1[0] * v1[0] + v1[1] * v1[1]) + v1[2] * v1[2];
  const double l2 = (v2[0] * v2[0] + v2[1] * v2[1]) + v2[2] * v2[2];
  const double l3 = (v3[0] * v3[0] + v3[1] * v3[1]) + v3[2] * v3[2];
  const double* v;
  double l;
  if (l1 >= l2 && l1 >= l3) {
    v = v1;
    l = l1;
  } else if (l2 >= l1 && l2 >= l3) {
    v = v2;
    l = l2;
  } else {
    v = v3;
    l = l3;
  }
  const double s = std::sqrt(l);
  evec[0] = v[0] / s;
  evec[1] = v[1] / s;
  evec[2] = v[2] / s;
}

// SampleConsensusModelPlane::optimizeModelCoefficients.  ORACLE CHOICE (documented
// deviation from PCL's float accumulators): the nine moments are accumulated in
// double in the canonical tree order over the pass' point positions (non-inliers
// contribute +0.0), so the result does not depend on a summation schedule.
void refine_model(const PlaneModelCtx& c, const float co[4], float out[4]) {
  std::memcpy(out, co, 4 * sizeof(float));
  std::vector<double> e[9];
  for (auto& v : e) v.assign(c.n, 0.0);
  int64_t cnt = 0;
  const bool valid = model_valid(c, co);
  for (int i = 0; i < c.n; ++i) {
    if (!valid || !(plane_dist(co, c.pts[i]) < c.thr)) continue;
    const double x = c.pts[i].x, y = c.pts[i].y, z = c.pts[i].z;
    e[0][i] = x * x;
    e[1][i] = x * y;
    e[2][i] = x * z;
    e[3][i] = y * y;
    e[4][i] = y * z;
    e[5][i] = z * z;
    e[6][i] = x;
    e[7][i] = y;
    e[8][i] = z;
    ++cnt;
  }
  if (cnt < 4) return;  // PCL: needs at least 4 inliers, else keep the RANSAC model
  double a[9];
  for (int k = 0; k < 9; ++k) a[k] = tree_sum(e[k].data(), c.n) / (double)cnt;
  double cov[9];
  cov[0] = a[0] - a[6] * a[6];
  cov[1] = a[1] - a[6] * a[7];
  cov[2] = a[2] - a[6] * a[8];
  cov[4] = a[3] - a[7] * a[7];
  cov[5] = a[4] - a[7] * a[8];
  cov[8] = a[5] - a[8] * a[8];
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
  double ev, vec[3];
  eigen33_smallest(cov, &ev, vec);
  float o[4];
  o[0] = (float)vec[0];
  o[1] = (float)vec[1];
  o[2] = (float)vec[2];
  const float cx = (float)a[6], cy = (float)a[7], cz = (float)a[8];
  o[3] = -(((o[0] * cx) + (o[1] * cy)) + (o[2] * cz));
  if (!model_valid(c, o)) return;  // revert
  std::memcpy(out, o, sizeof(o));
}

struct SegmentOut {
  bool ok = false;
  float ransac_coeff[4] = {0, 0, 0, 0};
  float coeff[4] = {0, 0, 0, 0};
  int n_ransac_inliers = 0;
  int iterations = 0;
  std::vector<int32_t> inliers;
};

// pcl::SACSegmentation::segment with SAC_RANSAC (ransac.hpp::computeModel), sequential as PCL.
void segment_once(const pcop_params& pr, const P4* pts, int n, SegmentOut& out) {
  out = SegmentOut();
  PlaneModelCtx c;
  c.pts = pts;
  c.n = n;
  c.thr = pr.plane_segment_dist_thres;
  c.eps_angle = (double)pr.plane_segment_angle;
  c.cos_eps = std::cos(c.eps_angle);
  for (int a = 0; a < 3; ++a) c.axis[a] = (double)pr.plane_axis[a];

  if (n < 3) return;  // getSamples: too few points -> empty selection -> computeModel fails
  Mt19937 rng(pr.ransac_seed);  // new model object per segment() call => re-seeded every pass
  std::vector<int32_t> shuf(n);
  std::iota(shuf.begin(), shuf.end(), 0);

  const int max_iter = pr.plane_max_iterations;
  const unsigned max_skip = (unsigned)max_iter * 10u;
  const double log_probability = det_log(1.0 - pr.plane_probability);
  const double one_over_indices = 1.0 / (double)n;
  int iterations = 0;
  int best = -INT_MAX;
  double k = 1.0;
  unsigned skipped = 0;
  bool have_model = false;
  float best_co[4] = {0, 0, 0, 0};

  while ((double)iterations < k && skipped < max_skip) {
    // getSamples: up to 1000 draws until isSampleGood
    int sel[3];
    bool good = false;
    for (int check = 0; check < 1000; ++check) {
      for (int i = 0; i < 3; ++i) {
        const int64_t j = i + (int64_t)((uint64_t)rng.rnd() % (uint64_t)(n - i));
        std::swap(shuf[i], shuf[j]);
      }
      sel[0] = shuf[0];
      sel[1] = shuf[1];
      sel[2] = shuf[2];
      if (!collinear_ratio_test(pts[sel[0]], pts[sel[1]], pts[sel[2]])) {
        good = true;
        break;
      }
    }
    if (!good) break;
    float co[4];
    if (!compute_model(pts[sel[0]], pts[sel[1]], pts[sel[2]], co)) {
      ++skipped;
      continue;
    }
    const int cnt = count_within(c, co);
    if (cnt > best) {
      best = cnt;
      have_model = true;
      std::memcpy(best_co, co, sizeof(co));
      const double w = (double)best * one_over_indices;
      double p_no_outliers = 1.0 - (w * w) * w;  // ORACLE CHOICE: pow(w,3) as (w*w)*w
      p_no_outliers = std::max(2.220446049250313e-16, p_no_outliers);
      p_no_outliers = std::min(1.0 - 2.220446049250313e-16, p_no_outliers);
      k = log_probability / det_log(p_no_outliers);
    }
    ++iterations;
    if (iterations > max_iter) break;
  }
  out.iterations = iterations;
  if (!have_model) return;
  out.ok = true;
  std::memcpy(out.ransac_coeff, best_co, sizeof(best_co));
  select_within(c, best_co, out.inliers);
  out.n_ransac_inliers = (int)out.inliers.size();
  std::memcpy(out.coeff, best_co, sizeof(best_co));
  if (pr.optimize_coefficients) {
    refine_model(c, best_co, out.coeff);
    select_within(c, out.coeff, out.inliers);
  }
}

struct PlaneOut {
  std::vector<P4> remaining;
  std::vector<int32_t> src;
  int n_passes = 0;
  int pass_points[PCOP_MAX_PLANE_PASSES_RECORDED];
  int pass_inliers[PCOP_MAX_PLANE_PASSES_RECORDED];
  float pass_coeff[PCOP_MAX_PLANE_PASSES_RECORDED][4];
  float last_coeff[4] = {0, 0, 0, 0};
  std::vector<int32_t> last_inliers;
  uint32_t warnings = 0;
};

void plane_loop(const pcop_params& pr, const P4* in, int s, PlaneOut& o) {
  o.remaining.assign(in, in + s);
  o.src.resize(s);
  std::iota(o.src.begin(), o.src.end(), 0);
  std::memset(o.pass_points, 0, sizeof(o.pass_points));
  std::memset(o.pass_inliers, 0, sizeof(o.pass_inliers));
  std::memset(o.pass_coeff, 0, sizeof(o.pass_coeff));
  const int nr_points = s;
  while ((double)o.remaining.size() > pr.plane_keep_fraction * (double)nr_points) {  // od.cpp:379
    SegmentOut seg;
    segment_once(pr, o.remaining.data(), (int)o.remaining.size(), seg);
    std::memcpy(o.last_coeff, seg.coeff, sizeof(seg.coeff));
    o.last_inliers = seg.inliers;
    if (seg.inliers.empty()) {  // od.cpp:383-387
      o.warnings |= PCOP_WARN_PLANE_BREAK;
      break;
    }
    if (o.n_passes < PCOP_MAX_PLANE_PASSES_RECORDED) {
      o.pass_points[o.n_passes] = (int)o.remaining.size();
      o.pass_inliers[o.n_passes] = (int)seg.inliers.size();
      std::memcpy(o.pass_coeff[o.n_passes], seg.coeff, sizeof(seg.coeff));
    }
    ++o.n_passes;
    // ExtractIndices negative: remaining points keep their relative order (od.cpp:395-397)
    std::vector<P4> rem;
    std::vector<int32_t> src;
    rem.reserve(o.remaining.size());
    src.reserve(o.remaining.size());
    size_t q = 0;
    for (size_t i = 0; i < o.remaining.size(); ++i) {
      if (q < seg.inliers.size() && (size_t)seg.inliers[q] == i) {
        ++q;
        continue;
      }
      rem.push_back(o.remaining[i]);
      src.push_back(o.src[i]);
    }
    o.remaining.swap(rem);
    o.src.swap(src);
  }
}

// ---------------------------------------------------------------------------
// a-6 Euclidean clustering: od.cpp:446-454 (+ kd-tree od.cpp:791-792), SURVEY 8a-6.
inline float radius2(float tol) { return (float)((double)tol * (double)tol); }

void canonical_clusters(std::vector<std::vector<int32_t>>& comps, int min_sz, int max_sz, int32_t* offsets,
                        int32_t* indices, int32_t* c_out, int32_t* l_out) {
  std::vector<std::vector<int32_t>> kept;
  for (auto& c : comps) {
    if ((int)c.size() < min_sz || (int)c.size() > max_sz) continue;  // oversize dropped whole
    std::sort(c.begin(), c.end());
    kept.push_back(std::move(c));
  }
  // canonical order: size descending, then min index ascending
  std::stable_sort(kept.begin(), kept.end(), [](const std::vector<int32_t>& a, const std::vector<int32_t>& b) {
    if (a.size() != b.size()) return a.size() > b.size();
    return a[0] < b[0];
  });
  int32_t l = 0;
  offsets[0] = 0;
  for (size_t k = 0; k < kept.size(); ++k) {
    for (int32_t i : kept[k]) indices[l++] = i;
    offsets[k + 1] = l;
  }
  *c_out = (int32_t)kept.size();
  *l_out = l;
}

void cluster_kdtree(const pcop_params& pr, const P4* pts, int p, int32_t* offsets, int32_t* indices, int32_t* c_out,
                    int32_t* l_out) {
  const float r2 = radius2(pr.euc_cluster_tolerance);
  KdTree t;
  t.build(pts, p);
  std::vector<char> processed(p, 0);
  std::vector<std::vector<int32_t>> comps;
  std::vector<int32_t> queue;
  for (int i = 0; i < p; ++i) {  // pcl::extractEuclideanClusters: BFS from each unprocessed index
    if (processed[i]) continue;
    queue.clear();
    queue.push_back(i);
    processed[i] = 1;
    for (size_t q = 0; q < queue.size(); ++q) {
      t.radius(0, pts[queue[q]], r2, [&](int j) {
        if (!processed[j]) {
          processed[j] = 1;
          queue.push_back(j);
        }
      });
    }
    comps.emplace_back(queue.begin(), queue.end());
  }
  canonical_clusters(comps, pr.euc_min_cluster_size, pr.euc_max_cluster_size, offsets, indices, c_out, l_out);
}

void cluster_brute(const pcop_params& pr, const P4* pts, int p, int32_t* offsets, int32_t* indices, int32_t* c_out,
                   int32_t* l_out) {
  const float r2 = radius2(pr.euc_cluster_tolerance);
  std::vector<int32_t> parent(p);
  std::iota(parent.begin(), parent.end(), 0);
  auto find = [&](int a) {
    while (parent[a] != a) {
      parent[a] = parent[parent[a]];
      a = parent[a];
    }
    return a;
  };
  for (int i = 0; i < p; ++i)
    for (int j = i + 1; j < p; ++j)
      if (dist2(pts[i], pts[j]) < r2) {
        const int a = find(i), b = find(j);
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
      }
  std::vector<std::vector<int32_t>> comps;
  std::vector<int32_t> slot(p, -1);
  for (int i = 0; i < p; ++i) {
    const int r = find(i);
    if (slot[r] < 0) {
      slot[r] = (int32_t)comps.size();
      comps.emplace_back();
    }
    comps[slot[r]].push_back(i);
  }
  canonical_clusters(comps, pr.euc_min_cluster_size, pr.euc_max_cluster_size, offsets, indices, c_out, l_out);
}

// ---------------------------------------------------------------------------
// a-7 centroid + bounding radius (north_star definition; distance arithmetic od.cpp:457-464).
void centroid_radius(const P4* pts, const int32_t* offsets, const int32_t* indices, int c, P4* obstacles) {
  for (int k = 0; k < c; ++k) {
    double sx = 0.0, sy = 0.0, sz = 0.0;
    const int b = offsets[k], e = offsets[k + 1];
    for (int j = b; j < e; ++j) {
      sx += (double)pts[indices[j]].x;
      sy += (double)pts[indices[j]].y;
      sz += (double)pts[indices[j]].z;
    }
    const double cnt = (double)(e - b);
    const float cx = (float)(sx / cnt), cy = (float)(sy / cnt), cz = (float)(sz / cnt);
    float r = 0.0f;
    for (int j = b; j < e; ++j) {
      const P4& q = pts[indices[j]];
      const float dx = q.x - cx, dy = q.y - cy, dz = q.z - cz;
      const float d = std::sqrt(((dx * dx) + (dy * dy)) + (dz * dz));
      if (d > r) r = d;
    }
    obstacles[k] = {cx, cy, cz, r};
  }
}

template <class T>
T* dup(const std::vector<T>& v) {
  T* p = (T*)std::malloc(std::max<size_t>(1, v.size()) * sizeof(T));
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

}  // namespace

// ===========================================================================
extern "C" {

int pcop_oracle_crop(const pcop_params* pr, const float* xyzw, int32_t n, float* out_xyzw, int32_t* kept_idx,
                     int32_t* m) {
  *m = crop(*pr, (const P4*)xyzw, n, (P4*)out_xyzw, kept_idx);
  return PCOP_OK;
}

int pcop_oracle_voxel(const pcop_params* pr, const float* xyzw, int32_t m, float* out_xyzw, uint32_t* out_keys,
                      int32_t* v, uint32_t* warnings) {
  if (warnings) *warnings = 0;
  *v = voxel((const P4*)xyzw, m, pr->downsample_size, (P4*)out_xyzw, out_keys, warnings);
  return PCOP_OK;
}

int pcop_oracle_voxel_keys(const pcop_params* pr, const float* xyzw, int32_t m, uint32_t* all_keys) {
  if (m <= 0) return PCOP_OK;
  const VoxelSetup s = voxel_setup((const P4*)xyzw, m, pr->downsample_size);
  for (int i = 0; i < m; ++i) all_keys[i] = voxel_key(s, ((const P4*)xyzw)[i]);
  return PCOP_OK;
}

int pcop_oracle_sor(const pcop_params* pr, const float* xyzw, int32_t v, float* out_xyzw, int32_t* kept_idx,
                    int32_t* s, uint32_t* warnings, float* distances, double* thr) {
  if (pr->statistical_outlier_meanK < 1) return PCOP_ERR_BAD_PARAM;
  if (warnings) *warnings = 0;
  *s = sor(*pr, (const P4*)xyzw, v, (P4*)out_xyzw, kept_idx, warnings, distances, thr);
  return PCOP_OK;
}

int pcop_oracle_plane(const pcop_params* pr, const float* xyzw, int32_t s, float* remaining_xyzw,
                      int32_t* remaining_src_idx, int32_t* p, int32_t* n_passes, int32_t* pass_points,
                      int32_t* pass_inliers, float* pass_coeff, float* last_coeff, int32_t* inlier_idx,
                      int32_t* n_inliers, uint32_t* warnings) {
  PlaneOut o;
  plane_loop(*pr, (const P4*)xyzw, s, o);
  *p = (int32_t)o.remaining.size();
  if (remaining_xyzw && !o.remaining.empty())
    std::memcpy(remaining_xyzw, o.remaining.data(), o.remaining.size() * sizeof(P4));
  if (remaining_src_idx && !o.src.empty())
    std::memcpy(remaining_src_idx, o.src.data(), o.src.size() * sizeof(int32_t));
  if (n_passes) *n_passes = o.n_passes;
  if (pass_points) std::memcpy(pass_points, o.pass_points, sizeof(o.pass_points));
  if (pass_inliers) std::memcpy(pass_inliers, o.pass_inliers, sizeof(o.pass_inliers));
  if (pass_coeff) std::memcpy(pass_coeff, o.pass_coeff, sizeof(o.pass_coeff));
  if (last_coeff) std::memcpy(last_coeff, o.last_coeff, sizeof(o.last_coeff));
  if (inlier_idx && !o.last_inliers.empty())
    std::memcpy(inlier_idx, o.last_inliers.data(), o.last_inliers.size() * sizeof(int32_t));
  if (n_inliers) *n_inliers = (int32_t)o.last_inliers.size();
  if (warnings) *warnings = o.warnings;
  return PCOP_OK;
}

int pcop_oracle_cluster(const pcop_params* pr, const float* xyzw, int32_t p, int32_t* cluster_offsets,
                        int32_t* cluster_indices, int32_t* c, int32_t* l) {
  cluster_kdtree(*pr, (const P4*)xyzw, p, cluster_offsets, cluster_indices, c, l);
  return PCOP_OK;
}

int pcop_oracle_cluster_bruteforce(const pcop_params* pr, const float* xyzw, int32_t p, int32_t* cluster_offsets,
                                   int32_t* cluster_indices, int32_t* c, int32_t* l) {
  cluster_brute(*pr, (const P4*)xyzw, p, cluster_offsets, cluster_indices, c, l);
  return PCOP_OK;
}

int pcop_oracle_centroid_radius(const float* xyzw, int32_t p, const int32_t* cluster_offsets,
                                const int32_t* cluster_indices, int32_t c, float* obstacles) {
  (void)p;
  centroid_radius((const P4*)xyzw, cluster_offsets, cluster_indices, c, (P4*)obstacles);
  return PCOP_OK;
}

int pcop_oracle_sor_distances_bruteforce(const float* xyzw, int32_t v, int32_t meanK, float* distances) {
  const P4* pts = (const P4*)xyzw;
  if (v <= meanK) return PCOP_ERR_BAD_PARAM;
  std::vector<float> d2(v);
  for (int i = 0; i < v; ++i) {
    for (int j = 0; j < v; ++j) d2[j] = dist2(pts[i], pts[j]);
    std::partial_sort(d2.begin(), d2.begin() + meanK + 1, d2.end());
    std::vector<float> head(d2.begin(), d2.begin() + meanK + 1);
    distances[i] = sor_mean_dist(head, meanK);
  }
  return PCOP_OK;
}

// pcl_ros::transformPointCloud -> pcl::transformPointCloud(Matrix4f) (od.cpp:696), PCL 1.7/1.8 transforms.hpp:
// explicit coefficient formula, float, evaluated left to right; non-dense clouds skip non-finite points (the output
// starts as a copy of the input).
int pcop_oracle_transform(const float* xyzw, int32_t n, const float* m, int32_t is_dense, float* out) {
  if (!xyzw || !m || !out || n < 0) return PCOP_ERR_BAD_PARAM;
  for (int32_t i = 0; i < n; ++i) {
    const float x = xyzw[4 * i], y = xyzw[4 * i + 1], z = xyzw[4 * i + 2];
    out[4 * i + 3] = xyzw[4 * i + 3];
    if (!is_dense && (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z))) {
      out[4 * i] = x;
      out[4 * i + 1] = y;
      out[4 * i + 2] = z;
      continue;
    }
    for (int r = 0; r < 3; ++r) {
      volatile float a = m[4 * r] * x;  // volatile: one IEEE operation per statement, no contraction, no reassociation
      volatile float b = m[4 * r + 1] * y;
      volatile float c = m[4 * r + 2] * z;
      volatile float s = a + b;
      s = s + c;
      s = s + m[4 * r + 3];
      out[4 * i + r] = s;
    }
  }
  return PCOP_OK;
}

// pcl::fromPCLPointCloud2<pcl::PointXYZ> (od.cpp:689): per point a memcpy of each mapped field into a
// default-constructed PointXYZ {0, 0, 0, 1.0f}
int pcop_oracle_pointcloud2_to_xyz(const unsigned char* data, int32_t n_points, int32_t point_step, int32_t off_x,
                                   int32_t off_y, int32_t off_z, float* out) {
  if (!data || !out || n_points < 0 || point_step < 4 || off_x < 0 || off_y < 0 || off_z < 0 || off_x + 4 > point_step ||
      off_y + 4 > point_step || off_z + 4 > point_step)
    return PCOP_ERR_BAD_PARAM;
  for (int32_t i = 0; i < n_points; ++i) {
    const unsigned char* rec = data + (size_t)i * (size_t)point_step;
    std::memcpy(out + 4 * (size_t)i + 0, rec + off_x, 4);
    std::memcpy(out + 4 * (size_t)i + 1, rec + off_y, 4);
    std::memcpy(out + 4 * (size_t)i + 2, rec + off_z, 4);
    out[4 * (size_t)i + 3] = 1.0f;
  }
  return PCOP_OK;
}

// pcl::toROSMsg(pcl::PointCloud<pcl::PointXYZ>) -> pcl::toPCLPointCloud2 (od.cpp:290-294 and the other debug
// publishers): msg.data is a memcpy of the PointXYZ array, point_step = sizeof(PointXYZ) = 16, fields x, y, z FLOAT32 at
// offsets 0, 4, 8.  Other layouts (not produced by the reference; offered for consumers with their own record format):
// the three fields at their offsets, the rest of the record zero.
int pcop_oracle_xyz_to_pointcloud2(const float* xyzw, int32_t n_points, int32_t point_step, int32_t off_x, int32_t off_y,
                                   int32_t off_z, unsigned char* out_data) {
  if ((n_points > 0 && (!xyzw || !out_data)) || n_points < 0 || point_step < 12 || off_x < 0 || off_y < 0 || off_z < 0 ||
      off_x + 4 > point_step || off_y + 4 > point_step || off_z + 4 > point_step)
    return PCOP_ERR_BAD_PARAM;
  const bool verbatim = point_step == 16 && off_x == 0 && off_y == 4 && off_z == 8;
  for (int32_t i = 0; i < n_points; ++i) {
    unsigned char* rec = out_data + (size_t)i * (size_t)point_step;
    if (verbatim) {
      std::memcpy(rec, xyzw + 4 * (size_t)i, 16);
      continue;
    }
    std::memset(rec, 0, (size_t)point_step);
    std::memcpy(rec + off_x, xyzw + 4 * (size_t)i + 0, 4);
    std::memcpy(rec + off_y, xyzw + 4 * (size_t)i + 1, 4);
    std::memcpy(rec + off_z, xyzw + 4 * (size_t)i + 2, 4);
  }
  return PCOP_OK;
}

// od.cpp:958-960: (int)ceil((fabs(lo) + fabs(hi)) / block_size).  ORACLE CHOICE: the unqualified fabs binds to
// ::fabs(double), so the sum and the division are evaluated in double.
int pcop_oracle_occupancy_dims(const pcop_params* pr, int32_t* width, int32_t* height) {
  if (!pr || !width || !height || !(pr->block_size > 0.0f)) return PCOP_ERR_BAD_PARAM;
  *width = (int32_t)std::ceil((std::fabs((double)pr->y_min) + std::fabs((double)pr->y_max)) / (double)pr->block_size);
  *height = (int32_t)std::ceil((std::fabs((double)pr->x_min) + std::fabs((double)pr->x_max)) / (double)pr->block_size);
  return PCOP_OK;
}

// get_occupancy_grid_x_y (od.cpp:134-150), called as (point.y, point.x, y_min, x_max, block_size) at od.cpp:203
static void occupancy_xy(float x, float y, float x_min, float y_max, float block_size, int* xc, int* yc) {
  int x_count = 0, y_count = 0;
  while (true) {
    volatile float step = (float)(x_count + 1) * block_size;
    volatile float edge = x_min + step;
    if (!(edge < x)) break;
    x_count++;
  }
  while (true) {
    volatile float step = (float)(y_count + 1) * block_size;
    volatile float edge = y_max - step;
    if (!(edge > y)) break;
    y_count++;
  }
  *xc = x_count;
  *yc = y_count;
}

int pcop_oracle_occupancy_grid(const pcop_params* pr, const float* xyzw, int32_t n, int8_t* grid_data, int64_t* counts_out,
                               int64_t* row_avg_out) {
  int32_t W = 0, H = 0;
  if (!xyzw || !grid_data || n < 0 || pcop_oracle_occupancy_dims(pr, &W, &H) != PCOP_OK || W <= 0 || H <= 0) return PCOP_ERR_BAD_PARAM;
  const long long size = (long long)W * H;
  std::vector<long long> counts((size_t)size, 0), row_avg((size_t)H, 0);
  for (int32_t i = 0; i < n; ++i) {  // od.cpp:195-215
    const float x = xyzw[4 * i], y = xyzw[4 * i + 1], z = xyzw[4 * i + 2];
    if (std::isnan(x) || x < pr->x_min || x > pr->x_max || z < pr->z_min || z > pr->z_max || y < pr->y_min || y > pr->y_max)
      continue;
    int xc, yc;
    occupancy_xy(y, x, pr->y_min, pr->x_max, pr->block_size, &xc, &yc);
    const long long index = (long long)yc * W + xc;  // (int arithmetic in the reference; identical in range)
    if (index >= size) continue;                     // od.cpp:205 "OUT OF BOUNDS INDEX ACCESS"
    counts[(size_t)index]++;
  }
  for (int r = 0; r < H; ++r) {  // od.cpp:226-234
    long long row = 0;
    for (int c = 0; c < W; ++c) row += counts[(size_t)r * W + c];
    row_avg[r] = row / W;
  }
  for (long long i = 0; i < size; ++i) {  // od.cpp:241-266
    const long long avg = row_avg[(size_t)(i / W)];
    volatile float one_minus = 1.0f - pr->dev_percent;
    volatile float thr = (float)avg * one_minus;
    grid_data[i] = ((float)counts[(size_t)i] < thr) ? 100 : 0;
  }
  if (counts_out) std::memcpy(counts_out, counts.data(), sizeof(long long) * (size_t)size);
  if (row_avg_out) std::memcpy(row_avg_out, row_avg.data(), sizeof(long long) * (size_t)H);
  return PCOP_OK;
}

// ---------------------------------------------------------------------------
// Shadow casting + obstacle marking on the occupancy grid (od.cpp:467-672, 817-833).
//
// ORACLE CHOICES (the reference leaves these to the platform or runs into undefined behaviour):
//  * unqualified fabs / sqrt / asin / tan / ceil bind to the double overloads of <math.h>; asin and tan are the
//    deterministic det_asin / det_tan (det_math.hpp, < 1e-15 from libm away from |D| = pi/2);
//  * the TF lookups (od.cpp:570, 580, 626) are explicit row-major 4x4 float matrices, applied with
//    pcl::transformPointCloud's coefficient formula (pcop_oracle_transform, dense cloud);
//  * get_occupancy_grid_x_y's while-loops stop at PCOP_OCC_COUNT_CAP = 2^20 steps (the reference would run on, and
//    overflow its int counter, for a point that far outside the arena);
//  * cell indices are formed in 64 bits (reference: int; identical whenever the int does not overflow);
//  * double -> int and float -> int conversions follow cvttsd2si / cvttss2si (NaN / out of range -> INT_MIN);
//  * a shadow line of more than PCOP_SHADOW_MAX_LINE = 65536 pixels, or a fan of more than 65536 lines, is not
//    drawn and raises PCOP_WARN_SHADOW_DEGENERATE (the reference would loop over it for minutes);
//  * the obstacle marking (od.cpp:823-833) is bounds-checked like the initial data set (od.cpp:205); the reference
//    writes unchecked.
namespace {
constexpr int OCC_COUNT_CAP = 1 << 20;
constexpr long long SHADOW_MAX_LINE = 65536;

inline int32_t cvt_d2i(double v) {
  if (v != v || v >= 2147483648.0 || v <= -2147483649.0) return INT32_MIN;
  return (int32_t)v;
}

// get_occupancy_grid_x_y (od.cpp:134-150), literal, with the step cap
static void occupancy_xy_capped(float x, float y, float x_min, float y_max, float block_size, int* xc, int* yc) {
  int x_count = 0, y_count = 0;
  while (x_count < OCC_COUNT_CAP) {
    volatile float step = (float)(x_count + 1) * block_size;
    volatile float edge = x_min + step;
    if (!(edge < x)) break;
    x_count++;
  }
  while (y_count < OCC_COUNT_CAP) {
    volatile float step = (float)(y_count + 1) * block_size;
    volatile float edge = y_max - step;
    if (!(edge > y)) break;
    y_count++;
  }
  *xc = x_count;
  *yc = y_count;
}

// one point through pcl::transformPointCloud's coefficient formula (row-major 4x4)
static P4 transform_one(const float* m, const P4& p) {
  P4 o = p;
  float* oo = &o.x;
  for (int r = 0; r < 3; ++r) {
    volatile float a = m[4 * r] * p.x;
    volatile float b = m[4 * r + 1] * p.y;
    volatile float c = m[4 * r + 2] * p.z;
    volatile float s = a + b;
    s = s + c;
    s = s + m[4 * r + 3];
    oo[r] = s;
  }
  return o;
}

// traceShadow (od.cpp:467-538), literal; returns false when the line is too long to draw
static bool trace_shadow(float v1x, float v1y, float v2x, float v2y, int8_t* grid, int W, long long size, int8_t opacity) {
  int x0 = cvt_f2i(v1x), x1 = cvt_f2i(v2x), y0 = cvt_f2i(v1y), y1 = cvt_f2i(v2y);
  // abs() of the int differences; evaluated in 64 bits (no overflow)
  const bool steep = std::llabs((long long)y1 - y0) > std::llabs((long long)x1 - x0);
  if (steep) {
    std::swap(x0, y0);
    std::swap(x1, y1);
  }
  if (x0 > x1) {
    std::swap(x0, x1);
    std::swap(y0, y1);
  }
  if ((long long)x1 - x0 + 1 > SHADOW_MAX_LINE) return false;
  const float dx = (float)(x1 - x0);
  const float dy = (float)(int32_t)((uint32_t)y1 - (uint32_t)y0);  // int difference, x86 wrap-around
  volatile float gradient = dy / dx;
  if (dx == 0.0f) gradient = 1.0f;
  volatile float intersect_y = (float)y0;
  for (int x = x0; x <= x1; ++x) {
    const int fl = cvt_f2i(std::floor(intersect_y));
    const long long grid_y = steep ? x : fl, grid_x = steep ? fl : x;
    long long idx = grid_y * W + grid_x;
    if (idx < size && idx > -1) grid[idx] = opacity;
    idx += 1;
    if (idx < size && idx > -1) grid[idx] = opacity;
    intersect_y = intersect_y + gradient;
  }
  return true;
}

}  // namespace

int pcop_oracle_occupancy_shadows(const pcop_params* pr, const float* remaining_xyzw, int32_t n_remaining,
                                             const int32_t* cluster_offsets, const int32_t* cluster_indices,
                                             int32_t n_clusters, const float* world_to_sensor16,
                                             const float* sensor_to_world16, int8_t* grid_data, int32_t* shadow_records,
                                             uint32_t* warnings) {
  int32_t W = 0, H = 0;
  if (!pr || !grid_data || n_remaining < 0 || n_clusters < 0 || (n_remaining > 0 && !remaining_xyzw) ||
      (n_clusters > 0 && (!cluster_offsets || !cluster_indices || !world_to_sensor16 || !sensor_to_world16)) ||
      pcop_oracle_occupancy_dims(pr, &W, &H) != PCOP_OK || W <= 0 || H <= 0)
    return PCOP_ERR_BAD_PARAM;
  const long long size = (long long)W * H;
  const P4* cloud = (const P4*)remaining_xyzw;
  const float bs = pr->block_size;
  const int8_t opacity = (int8_t)pr->grid_opacity;  // int stored into a char cell (od.cpp:508)
  uint32_t warn = 0;
  for (int32_t c = 0; c < n_clusters; ++c) {  // handle_shadow_casting (od.cpp:584-672), one call per cluster (od.cpp:817-821)
    int32_t* rec = shadow_records ? shadow_records + 6 * (size_t)c : nullptr;
    if (rec) std::fill(rec, rec + 6, 0);
    const int32_t o0 = cluster_offsets[c], o1 = cluster_offsets[c + 1];
    if (o1 - o0 < 2) continue;  // od.cpp:574
    // od.cpp:587-609: members into the sensor frame, extrema with strict compares (first occurrence wins)
    P4 vmin_pt = transform_one(world_to_sensor16, cloud[cluster_indices[o0]]);
    float vmax = vmin_pt.x, hmin = vmin_pt.y, hmax = vmin_pt.y;
    for (int32_t j = o0 + 1; j < o1; ++j) {
      const P4 q = transform_one(world_to_sensor16, cloud[cluster_indices[j]]);
      if (q.x < vmin_pt.x) vmin_pt = q;
      if (q.x > vmax) vmax = q.x;
      if (q.y < hmin) hmin = q.y;
      if (q.y > hmax) hmax = q.y;
    }
    volatile float hdiff = hmax - hmin;
    const float width = std::fabs(hdiff);  // od.cpp:616
    // calculate_shadow_cast (od.cpp:540-582)
    const float a = vmin_pt.z;
    const float b = std::fabs(vmin_pt.x);
    volatile float aa = a * a, bb = b * b;
    volatile float ab = aa + bb;
    const float cc = (float)std::sqrt((double)ab);
    const float e = (float)((std::fabs((double)vmax) - std::fabs((double)vmin_pt.x)) + 0.04);
    volatile float a_over_c = a / cc;
    const float D = (float)pcop_oracle::det_asin((double)a_over_c);
    const float d = (float)(pcop_oracle::det_tan((double)D) * (double)e + 0.25);
    volatile float xx = vmin_pt.x * vmin_pt.x, yy = vmin_pt.y * vmin_pt.y, zz = vmin_pt.z * vmin_pt.z;
    volatile float s2 = xx + yy;
    s2 = s2 + zz;
    const float v_len = (float)std::sqrt((double)s2);
    P4 end = vmin_pt;
    {
      volatile float nx = vmin_pt.x / v_len, ny = vmin_pt.y / v_len, nz = vmin_pt.z / v_len;
      nx = nx * d;
      ny = ny * d;
      nz = nz * d;
      volatile float ex = nx + vmin_pt.x, ey = ny + vmin_pt.y, ez = nz + vmin_pt.z;
      end.x = ex;
      end.y = ey;
      end.z = ez;
    }
    const P4 world_end = transform_one(sensor_to_world16, end);
    int end_x, end_y, start_x, start_y;
    occupancy_xy_capped(world_end.y, world_end.x, pr->y_min, pr->x_max, bs, &end_x, &end_y);  // od.cpp:569
    const P4 world_start = transform_one(sensor_to_world16, vmin_pt);                        // od.cpp:634-634
    occupancy_xy_capped(world_start.y, world_start.x, pr->y_min, pr->x_max, bs, &start_x, &start_y);
    // od.cpp:642-643: first += ceil((width / block_size) / 2)   (int += double)
    volatile float wb = width / bs;
    volatile float half = wb / 2.0f;
    const double shift = std::ceil((double)half);
    start_x = cvt_d2i((double)start_x + shift);
    end_x = cvt_d2i((double)end_x + shift);
    // od.cpp:645: for (int i = 0; i < ceil(width / block_size) + 3; i++)
    const double lim = std::ceil((double)wb) + 3.0;
    long long n_lines = 0;
    if (lim == lim && lim > 0.0) n_lines = (lim > 1.0e9) ? 1000000000ll : (long long)std::ceil(lim);
    bool skipped = false;
    if (n_lines > SHADOW_MAX_LINE) {
      skipped = true;
      n_lines = 0;
    }
    if (rec) {
      rec[0] = start_x;
      rec[1] = start_y;
      rec[2] = end_x;
      rec[3] = end_y;
    }
    for (long long i = 0; i < n_lines; ++i) {
      // Vertex holds floats (od.cpp:127-131); first -= 1 per line (od.cpp:659-660), int wrap-around as on x86
      const int32_t sx = (int32_t)((uint32_t)start_x - (uint32_t)i), ex = (int32_t)((uint32_t)end_x - (uint32_t)i);
      if (!trace_shadow((float)sx, (float)start_y, (float)ex, (float)end_y, grid_data, W, size, opacity)) skipped = true;
    }
    if (skipped) warn |= PCOP_WARN_SHADOW_DEGENERATE;
    if (rec) {
      rec[4] = (int32_t)n_lines;
      rec[5] = skipped ? 1 : 0;
    }
  }
  for (int32_t i = 0; i < n_remaining; ++i) {  // od.cpp:823-833
    if (std::isnan(cloud[i].x)) continue;
    int xc, yc;
    occupancy_xy_capped(cloud[i].y, cloud[i].x, pr->y_min, pr->x_max, bs, &xc, &yc);
    const long long idx = (long long)yc * W + xc;
    if (idx < size) grid_data[idx] = 100;
  }
  if (warnings) *warnings = warn;
  return PCOP_OK;
}

int pcop_oracle_process(const pcop_params* pr, const float* xyzw, int32_t n, pcop_frame_result* out) {
  std::memset(out, 0, sizeof(*out));
  const P4* in = (const P4*)xyzw;
  out->n_input = n;
  uint32_t warn = 0;

  // crop
  std::vector<P4> a(in, in + n);
  std::vector<int32_t> crop_kept(n);
  if (pr->enable_crop) {
    const int m = crop(*pr, in, n, a.data(), crop_kept.data());
    a.resize(m);
    crop_kept.resize(m);
  } else {
    std::iota(crop_kept.begin(), crop_kept.end(), 0);
  }
  out->n_crop = (int32_t)a.size();

  // voxel
  std::vector<P4> b(a.size());
  std::vector<uint32_t> vkeys(a.size());
  if (pr->enable_voxel) {
    const int v = voxel(a.data(), (int)a.size(), pr->downsample_size, b.data(), vkeys.data(), &warn);
    b.resize(v);
    vkeys.resize(v);
  } else {
    b = a;
    std::fill(vkeys.begin(), vkeys.end(), 0u);
  }
  out->n_voxel = (int32_t)b.size();

  // SOR
  std::vector<P4> c(b.size());
  std::vector<int32_t> sor_kept(b.size());
  if (pr->enable_sor) {
    if (pr->statistical_outlier_meanK < 1) return PCOP_ERR_BAD_PARAM;
    const int s = sor(*pr, b.data(), (int)b.size(), c.data(), sor_kept.data(), &warn, nullptr, nullptr);
    c.resize(s);
    sor_kept.resize(s);
  } else {
    c = b;
    std::iota(sor_kept.begin(), sor_kept.end(), 0);
  }
  out->n_sor = (int32_t)c.size();

  // plane loop
  PlaneOut po;
  if (pr->enable_plane) {
    plane_loop(*pr, c.data(), (int)c.size(), po);
  } else {
    po.remaining = c;
    po.src.resize(c.size());
    std::iota(po.src.begin(), po.src.end(), 0);
    std::memset(po.pass_points, 0, sizeof(po.pass_points));
    std::memset(po.pass_inliers, 0, sizeof(po.pass_inliers));
    std::memset(po.pass_coeff, 0, sizeof(po.pass_coeff));
  }
  warn |= po.warnings;
  out->n_remaining = (int32_t)po.remaining.size();
  out->n_plane_passes = po.n_passes;
  std::memcpy(out->plane_pass_points, po.pass_points, sizeof(po.pass_points));
  std::memcpy(out->plane_pass_inliers, po.pass_inliers, sizeof(po.pass_inliers));
  std::memcpy(out->plane_pass_coeff, po.pass_coeff, sizeof(po.pass_coeff));
  std::memcpy(out->plane_coeff, po.last_coeff, sizeof(po.last_coeff));
  out->n_plane_inliers = (int32_t)po.last_inliers.size();

  // clusters
  const int p = (int)po.remaining.size();
  std::vector<int32_t> offs(p + 2, 0), idx(p + 1, 0);
  int32_t cc = 0, ll = 0;
  if (pr->enable_cluster && p > 0)
    cluster_kdtree(*pr, po.remaining.data(), p, offs.data(), idx.data(), &cc, &ll);
  offs.resize(cc + 1);
  idx.resize(ll);
  out->n_clusters = cc;
  out->n_cluster_points = ll;
  std::vector<P4> obst(cc);
  if (cc > 0) centroid_radius(po.remaining.data(), offs.data(), idx.data(), cc, obst.data());

  out->warnings = warn;
  out->crop_kept_idx = dup(crop_kept);
  out->voxel_keys = dup(vkeys);
  out->voxel_centroids = (const float*)dup(b);
  out->sor_kept_idx = dup(sor_kept);
  out->plane_inlier_idx = dup(po.last_inliers);
  out->remaining_cloud = (const float*)dup(po.remaining);
  out->remaining_src_idx = dup(po.src);
  out->cluster_offsets = dup(offs);
  out->cluster_indices = dup(idx);
  out->obstacles = (const float*)dup(obst);
  return PCOP_OK;
}

void pcop_oracle_free_result(pcop_frame_result* r) {
  std::free((void*)r->crop_kept_idx);
  std::free((void*)r->voxel_keys);
  std::free((void*)r->voxel_centroids);
  std::free((void*)r->sor_kept_idx);
  std::free((void*)r->plane_inlier_idx);
  std::free((void*)r->remaining_cloud);
  std::free((void*)r->remaining_src_idx);
  std::free((void*)r->cluster_offsets);
  std::free((void*)r->cluster_indices);
  std::free((void*)r->obstacles);
  std::memset(r, 0, sizeof(*r));
}

void pcop_oracle_rng_raw(uint32_t seed, int32_t count, uint32_t* raw, int32_t* rnd) {
  Mt19937 g(seed);
  for (int i = 0; i < count; ++i) {
    const uint32_t r = g.next();
    if (raw) raw[i] = r;
    if (rnd) rnd[i] = (int32_t)(r >> 1);
  }
}

void pcop_oracle_draw_samples(uint32_t seed, int32_t n_points, int32_t n_samples, int32_t* samples3) {
  Mt19937 g(seed);
  std::vector<int32_t> shuf(n_points);
  std::iota(shuf.begin(), shuf.end(), 0);
  for (int s = 0; s < n_samples; ++s) {
    for (int i = 0; i < 3; ++i) {
      const int64_t j = i + (int64_t)((uint64_t)g.rnd() % (uint64_t)(n_points - i));
      std::swap(shuf[i], shuf[j]);
    }
    for (int i = 0; i < 3; ++i) samples3[3 * s + i] = shuf[i];
  }
}

float pcop_oracle_radius2(float tolerance) { return radius2(tolerance); }
float pcop_oracle_inverse_leaf(float leaf) { return 1.0f / leaf; }

int pcop_oracle_segment_once(const pcop_params* pr, const float* xyzw, int32_t n, float* ransac_coeff,
                             float* refined_coeff, int32_t* n_ransac_inliers, int32_t* n_refined_inliers,
                             int32_t* iterations) {
  SegmentOut s;
  segment_once(*pr, (const P4*)xyzw, n, s);
  std::memcpy(ransac_coeff, s.ransac_coeff, 4 * sizeof(float));
  std::memcpy(refined_coeff, s.coeff, 4 * sizeof(float));
  *n_ransac_inliers = s.n_ransac_inliers;
  *n_refined_inliers = (int32_t)s.inliers.size();
  *iterations = s.iterations;
  return s.ok ? PCOP_OK : PCOP_ERR_INTERNAL;
}

double pcop_oracle_det_log(double x) { return det_log(x); }
double pcop_oracle_det_atan2_ypos(double y, double x) { return det_atan2_ypos(y, x); }
double pcop_oracle_det_sin(double x) { return det_sin(x); }
double pcop_oracle_det_cos(double x) { return det_cos(x); }
double pcop_oracle_det_asin(double q) { return pcop_oracle::det_asin(q); }
double pcop_oracle_det_tan(double x) { return pcop_oracle::det_tan(x); }
double pcop_oracle_tree_sum(const double* v, int32_t n) { return tree_sum(v, n); }
void pcop_oracle_eigen33_smallest(const double* m9, double* eval, double* evec3) {
  eigen33_smallest(m9, eval, evec3);
}

}  // extern "C"
