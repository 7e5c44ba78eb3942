// Minimal C++ host over the C ABI (include/pcop.h), the shape of the shim INTEGRATION.md describes: parameters by
// their params.yaml names, one pcop_process call per frame in place of od.cpp:727-797, then the occupancy-grid
// product (od.cpp:727 grid part, od.cpp:817-833).  No PCL / ROS types: pcl::PointXYZ is four floats.
//
//   g++ -std=c++11 -I include examples/pcop_host_demo.cpp -L pointcloud_obstacle_processing_b200 -lpcop -o demo
//   LD_LIBRARY_PATH=pointcloud_obstacle_processing_b200 ./demo
//
// Exit codes: 0 ok, 3 no usable CUDA device (the library has no CPU fallback), 1 any other failure.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "pcop.h"

namespace {

struct Lcg {  // tiny deterministic generator for the synthetic frame
  uint64_t s;
  float next() {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((s >> 40) & 0xffffff) / 16777216.0f;
  }
};

std::vector<float> make_frame(int n) {  // floor at z = -0.3 plus two boxes, inside the params.yaml crop box
  std::vector<float> c((size_t)n * 4);
  Lcg g{42};
  for (int i = 0; i < n; ++i) {
    float x = 4.5f * g.next(), y = 3.78f * g.next(), z = -0.3f + 0.004f * (g.next() - 0.5f);
    const int kind = i % 10;
    if (kind == 0) {
      x = 1.0f + 0.25f * g.next(); y = 1.0f + 0.25f * g.next(); z = -0.3f + 0.3f * g.next();
    } else if (kind == 1) {
      x = 3.0f + 0.2f * g.next(); y = 2.5f + 0.3f * g.next(); z = -0.3f + 0.25f * g.next();
    }
    c[4 * (size_t)i + 0] = x;
    c[4 * (size_t)i + 1] = y;
    c[4 * (size_t)i + 2] = z;
    c[4 * (size_t)i + 3] = 1.0f;
  }
  return c;
}

}  // namespace

int main() {
  if (pcop_abi_version() != PCOP_ABI_VERSION) {
    std::fprintf(stderr, "pcop: header / library ABI mismatch\n");
    return 1;
  }
  pcop_params p;
  pcop_params_init_params_yaml(&p);  // minibot_cr18/params.yaml
  const int n = 30000;
  const std::vector<float> cloud = make_frame(n);

  pcop_handle* h = nullptr;
  const int st = pcop_create(&p, /*device=*/0, (size_t)n, /*max_batch=*/1, &h);
  if (st != PCOP_OK) {
    std::fprintf(stderr, "pcop_create failed (status %d): %s\n", st, pcop_global_error());
    return st == PCOP_ERR_CUDA ? 3 : 1;
  }

  int32_t W = 0, H = 0;
  if (pcop_occupancy_dims(h, &W, &H) != PCOP_OK) return 1;
  std::vector<int8_t> grid((size_t)W * H);
  if (pcop_occupancy_grid(h, cloud.data(), n, grid.data(), nullptr, nullptr) != PCOP_OK) {  // od.cpp:727
    std::fprintf(stderr, "pcop_occupancy_grid: %s\n", pcop_last_error(h));
    return 1;
  }
  pcop_frame_result r;
  if (pcop_process(h, cloud.data(), n, &r) != PCOP_OK) {  // od.cpp:727-797
    std::fprintf(stderr, "pcop_process: %s\n", pcop_last_error(h));
    return 1;
  }
  std::printf("N=%d crop=%d voxel=%d sor=%d remaining=%d clusters=%d (device %.0f us)\n", r.n_input, r.n_crop, r.n_voxel,
              r.n_sor, r.n_remaining, r.n_clusters, pcop_last_elapsed_us(h));
  for (int c = 0; c < r.n_clusters && c < 4; ++c)
    std::printf("  obstacle %d: %d points, centre (%.3f, %.3f, %.3f), radius %.3f\n", c,
                r.cluster_offsets[c + 1] - r.cluster_offsets[c], r.obstacles[4 * c], r.obstacles[4 * c + 1],
                r.obstacles[4 * c + 2], r.obstacles[4 * c + 3]);

  // sensor 1.2 m above the far end of the arena, looking along -x (what the two TF lookups of od.cpp:592 / 570 return)
  const float to_world[16] = {-1, 0, 0, 5.3f, 0, -1, 0, 1.89f, 0, 0, 1, 1.2f, 0, 0, 0, 1};
  const float to_sensor[16] = {-1, 0, 0, 5.3f, 0, -1, 0, 1.89f, 0, 0, 1, -1.2f, 0, 0, 0, 1};  // its inverse
  uint32_t warn = 0;
  if (pcop_occupancy_shadows(h, r.remaining_cloud, r.n_remaining, r.cluster_offsets, r.cluster_indices, r.n_clusters,
                             to_sensor, to_world, grid.data(), nullptr, &warn) != PCOP_OK) {  // od.cpp:817-833
    std::fprintf(stderr, "pcop_occupancy_shadows: %s\n", pcop_last_error(h));
    return 1;
  }
  long marked = 0;
  for (size_t i = 0; i < grid.size(); ++i) marked += grid[i] == 100;
  std::printf("occupancy grid %d x %d, %ld cells marked, warnings %u\n", W, H, marked, warn);
  pcop_destroy(h);
  return 0;
}
