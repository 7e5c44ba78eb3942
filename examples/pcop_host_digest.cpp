// C++ host over the C ABI (include/pcop.h) that runs recorded frames through the pipeline and prints a digest of every
// result array, so that a test can compare a pure C++ caller -- the language the node's host side is written in --
// with the CPU oracle's golden fixtures (tests/golden/pipeline_golden.json), bit for bit.
//
//   pcop_host_digest <frames.bin> <points_per_frame> <n_frames> <config>
//
// frames.bin: n_frames x points_per_frame x {x, y, z, 1.0f} float32 (pcl::PointXYZ records).  config 1 =
// minibot_cr18/params.yaml verbatim (pcop_params_init_params_yaml); config 2 = the HDL-64 parameter set of SURVEY 8d.
// Every frame is processed twice: one pcop_process call per frame, and all frames in one pcop_process_batch call;
// one output line per frame and mode:
//   frame <f> <single|batch> counts N M V S P C L passes inliers warnings | crc <9 CRC-32 values> | obstacles ...
//
//   g++ -std=c++11 -I include examples/pcop_host_digest.cpp -L pointcloud_obstacle_processing_b200 -lpcop -o digest
//
// Exit codes: 0 ok, 3 no usable CUDA device (the library has no CPU fallback), 1 any other failure.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pcop.h"

namespace {

uint32_t crc32(const void* data, size_t n) {  // the zlib polynomial, bitwise
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < n; ++i) {
    c ^= p[i];
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
  }
  return c ^ 0xffffffffu;
}

void print_frame(int f, const char* mode, const pcop_frame_result& r) {
  std::printf("frame %d %s counts %d %d %d %d %d %d %d %d %d %u | crc", f, mode, r.n_input, r.n_crop, r.n_voxel, r.n_sor,
              r.n_remaining, r.n_clusters, r.n_cluster_points, r.n_plane_passes, r.n_plane_inliers, r.warnings);
  std::printf(" %u", crc32(r.crop_kept_idx, (size_t)r.n_crop * 4));
  std::printf(" %u", crc32(r.voxel_keys, (size_t)r.n_voxel * 4));
  std::printf(" %u", crc32(r.voxel_centroids, (size_t)r.n_voxel * 16));
  std::printf(" %u", crc32(r.sor_kept_idx, r.sor_kept_idx ? (size_t)r.n_sor * 4 : 0));
  std::printf(" %u", crc32(r.plane_inlier_idx, (size_t)r.n_plane_inliers * 4));
  std::printf(" %u", crc32(r.remaining_src_idx, (size_t)r.n_remaining * 4));
  std::printf(" %u", crc32(r.remaining_cloud, (size_t)r.n_remaining * 16));
  std::printf(" %u", crc32(r.cluster_offsets, (size_t)(r.n_clusters + 1) * 4));
  std::printf(" %u", crc32(r.cluster_indices, (size_t)r.n_cluster_points * 4));
  std::printf(" | obstacles");
  for (int c = 0; c < r.n_clusters; ++c)
    std::printf(" %.9g %.9g %.9g %.9g", r.obstacles[4 * c], r.obstacles[4 * c + 1], r.obstacles[4 * c + 2], r.obstacles[4 * c + 3]);
  std::printf("\n");
}

}  // namespace

int main(int argc, char** argv) {
  if (argc != 5) {
    std::fprintf(stderr, "usage: %s <frames.bin> <points_per_frame> <n_frames> <config 1|2>\n", argv[0]);
    return 1;
  }
  const int n = std::atoi(argv[2]), frames = std::atoi(argv[3]), config = std::atoi(argv[4]);
  if (n <= 0 || frames <= 0 || (config != 1 && config != 2)) return 1;
  std::vector<float> cloud((size_t)frames * n * 4);
  std::FILE* fp = std::fopen(argv[1], "rb");
  if (!fp || std::fread(cloud.data(), sizeof(float), cloud.size(), fp) != cloud.size()) {
    std::fprintf(stderr, "cannot read %s\n", argv[1]);
    return 1;
  }
  std::fclose(fp);

  pcop_params p;
  pcop_params_init_params_yaml(&p);  // minibot_cr18/params.yaml (config 1)
  if (config == 2) {                 // HDL-64 frame, SOR off
    p.x_min = -40.0f; p.x_max = 40.0f; p.y_min = -40.0f; p.y_max = 40.0f; p.z_min = -2.5f; p.z_max = 1.5f;
    p.downsample_size = 0.1f;
    p.enable_sor = 0;
    p.plane_segment_dist_thres = 0.2f;
    p.euc_cluster_tolerance = 0.5f; p.euc_min_cluster_size = 10; p.euc_max_cluster_size = 100000;
  }
  p.publish_point_clouds = 1;  // od.cpp:945: every intermediate comes back
  p.outputs = PCOP_OUT_ALL;

  pcop_handle* h = nullptr;
  const int st = pcop_create(&p, /*device=*/0, (size_t)n, /*max_batch=*/frames, &h);
  if (st != PCOP_OK) {
    std::fprintf(stderr, "pcop_create failed (status %d): %s\n", st, pcop_global_error());
    return st == PCOP_ERR_CUDA ? 3 : 1;
  }
  for (int f = 0; f < frames; ++f) {
    pcop_frame_result r;
    if (pcop_process(h, cloud.data() + (size_t)f * n * 4, n, &r) != PCOP_OK) {
      std::fprintf(stderr, "pcop_process: %s\n", pcop_last_error(h));
      return 1;
    }
    print_frame(f, "single", r);
  }
  std::vector<pcop_frame_result> res((size_t)frames);
  std::vector<int32_t> counts((size_t)frames, n);
  if (pcop_process_batch(h, cloud.data(), (size_t)n, counts.data(), frames, res.data()) != PCOP_OK) {
    std::fprintf(stderr, "pcop_process_batch: %s\n", pcop_last_error(h));
    return 1;
  }
  for (int f = 0; f < frames; ++f) print_frame(f, "batch", res[(size_t)f]);
  pcop_destroy(h);
  return 0;
}
