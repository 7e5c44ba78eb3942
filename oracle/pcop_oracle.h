/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load or call anything in this directory.
 *
 * PARITY UNPINNED: the reference (stateSpaceRobotics/pointcloud_obstacle_processing)
 * ships no tests, golden vectors or fixtures, and its arithmetic lives in PCL /
 * FLANN / Boost.Random / Eigen, none of which is vendored, version-pinned
 * (CMakeLists.txt:13, package.xml:49-51) or installed here.  This oracle is a CPU
 * restatement of the published PCL 1.7/1.8 algorithms at the reference's call
 * sites (od.cpp = minibot_cr18/src/obstacle_detection.cpp); where PCL's result is
 * build-dependent the choice made is written beside the code ("ORACLE CHOICE").
 */
#ifndef PCOP_ORACLE_H_
#define PCOP_ORACLE_H_

#include "../include/pcop.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Stage functions.  Same meaning as the pcop_* stage entry points of include/pcop.h. */
int pcop_oracle_crop(const pcop_params* pr, const float* xyzw, int32_t n, float* out_xyzw, int32_t* kept_idx,
                     int32_t* m);
int pcop_oracle_voxel(const pcop_params* pr, const float* xyzw, int32_t m, float* out_xyzw, uint32_t* out_keys,
                      int32_t* v, uint32_t* warnings);
/* all_keys (optional, [m]) receives the per-point voxel key before sorting. */
int pcop_oracle_voxel_keys(const pcop_params* pr, const float* xyzw, int32_t m, uint32_t* all_keys);
/* distances (optional, [v]) and thr (optional) expose the SOR internals for margin reports. */
int pcop_oracle_sor(const pcop_params* pr, const float* xyzw, int32_t v, float* out_xyzw, int32_t* kept_idx,
                    int32_t* s, uint32_t* warnings, float* distances, double* thr);
int pcop_oracle_plane(const pcop_params* pr, const float* xyzw, int32_t s, float* remaining_xyzw,
                      int32_t* remaining_src_idx, int32_t* p, int32_t* n_passes, int32_t* pass_points,
                      int32_t* pass_inliers, float* pass_coeff, float* last_coeff, int32_t* inlier_idx,
                      int32_t* n_inliers, uint32_t* warnings);
int pcop_oracle_cluster(const pcop_params* pr, const float* xyzw, int32_t p, int32_t* cluster_offsets,
                        int32_t* cluster_indices, int32_t* c, int32_t* l);
int pcop_oracle_centroid_radius(const float* xyzw, int32_t p, const int32_t* cluster_offsets,
                                const int32_t* cluster_indices, int32_t c, float* obstacles);

/* Accumulator ingest (od.cpp:691-698): pcl_ros::transformPointCloud of one incoming cloud with a row-major 4x4 float
 * matrix.  PCL 1.7/1.8 pcl::transformPointCloud(cloud_in, cloud_out, Eigen::Matrix4f) evaluates, in float,
 *   out.x = m00*x + m01*y + m02*z + m03   (left to right, no FMA), likewise y and z; the 4th float is copied.
 * is_dense == 0 (Kinect clouds carry NaN holes): points with a non-finite x, y or z are copied unchanged. */
int pcop_oracle_transform(const float* xyzw, int32_t n, const float* m16, int32_t is_dense, float* out_xyzw);

/* Wire ingest (od.cpp:688-689): pcl_conversions::toPCL + pcl::fromPCLPointCloud2<PointXYZ> of a sensor_msgs/PointCloud2
 * payload: point i's FLOAT32 fields x, y, z sit at data + i*point_step + off_{x,y,z} (little endian); the PointXYZ
 * padding float is 1.0f (default-constructed point, the field copy does not touch it). */
int pcop_oracle_pointcloud2_to_xyz(const unsigned char* data, int32_t n_points, int32_t point_step, int32_t off_x,
                                   int32_t off_y, int32_t off_z, float* out_xyzw);
/* Wire egress (od.cpp:290-294): pcl::toROSMsg of a PointXYZ cloud -- the records verbatim for the (16, 0, 4, 8) layout;
 * for any other layout the three FLOAT32 fields at their offsets, other bytes zero. */
int pcop_oracle_xyz_to_pointcloud2(const float* xyzw, int32_t n_points, int32_t point_step, int32_t off_x, int32_t off_y,
                                   int32_t off_z, unsigned char* out_data);

/* Initial occupancy-grid data set (od.cpp:134-157, 175-269, sizes od.cpp:958-960), literal: per crop survivor the cell
 * found by the two while-loops (float arithmetic), int64 counts, per-row integer average, cell = 100 when
 * (float)count < (float)row_avg * (1.0f - dev_percent), else 0.  width/height as in od.cpp:958-959 (double division).
 * counts / row_avg may be NULL. */
int pcop_oracle_occupancy_dims(const pcop_params* pr, int32_t* width, int32_t* height);
int pcop_oracle_occupancy_grid(const pcop_params* pr, const float* xyzw, int32_t n, int8_t* grid_data, int64_t* counts,
                               int64_t* row_avg);

/* Shadow casting + obstacle marking on the grid (od.cpp:467-672, 817-833): per cluster of >= 2 points the members go
 * into the sensor frame (world_to_sensor16 = the "kinect2_link" <- "world" lookup of od.cpp:592), the point with the
 * smallest sensor x starts a fan of ceil(width / block_size) + 3 lines (traceShadow, cells set to grid_opacity) towards
 * the shadow end point (calculate_shadow_cast, back through sensor_to_world16); then every remaining point marks its
 * cell 100.  grid_data [height*width] is updated in place.  shadow_records (optional, [C][6]) = start_x, start_y,
 * end_x, end_y (grid cells of the first line, after the half-width shift), lines drawn, skipped flag.  The choices the
 * reference leaves open are listed beside the code ("ORACLE CHOICES"). */
int pcop_oracle_occupancy_shadows(const pcop_params* pr, const float* remaining_xyzw, int32_t n_remaining,
                                  const int32_t* cluster_offsets, const int32_t* cluster_indices, int32_t n_clusters,
                                  const float* world_to_sensor16, const float* sensor_to_world16, int8_t* grid_data,
                                  int32_t* shadow_records, uint32_t* warnings);
double pcop_oracle_det_asin(double q);
double pcop_oracle_det_tan(double x);

/* Whole pipeline; result arrays are malloc'ed, release with pcop_oracle_free_result.
 * All PCOP_OUT_* arrays are always filled. */
int pcop_oracle_process(const pcop_params* pr, const float* xyzw, int32_t n, pcop_frame_result* out);
void pcop_oracle_free_result(pcop_frame_result* r);

/* O(n^2) cross-checks using the exact float predicate (bit-authoritative for small n). */
int pcop_oracle_cluster_bruteforce(const pcop_params* pr, const float* xyzw, int32_t p, int32_t* cluster_offsets,
                                   int32_t* cluster_indices, int32_t* c, int32_t* l);
/* mean distance to the meanK nearest neighbours, brute force */
int pcop_oracle_sor_distances_bruteforce(const float* xyzw, int32_t v, int32_t meanK, float* distances);

/* Known-answer hooks. */
void pcop_oracle_rng_raw(uint32_t seed, int32_t count, uint32_t* raw, int32_t* rnd);
void pcop_oracle_draw_samples(uint32_t seed, int32_t n_points, int32_t n_samples, int32_t* samples3);
float pcop_oracle_radius2(float tolerance);
float pcop_oracle_inverse_leaf(float leaf);
/* one RANSAC segment() call: best model before refinement, refined model, counts */
int pcop_oracle_segment_once(const pcop_params* pr, const float* xyzw, int32_t n, float* ransac_coeff,
                             float* refined_coeff, int32_t* n_ransac_inliers, int32_t* n_refined_inliers,
                             int32_t* iterations);
/* deterministic elementary functions, exposed for accuracy tests */
double pcop_oracle_det_log(double x);
double pcop_oracle_det_atan2_ypos(double y, double x);
double pcop_oracle_det_sin(double x);
double pcop_oracle_det_cos(double x);
/* canonical tree sum (shared summation-order spec), exposed for tests */
double pcop_oracle_tree_sum(const double* v, int32_t n);
/* smallest eigenpair of a symmetric 3x3 (row-major 9 doubles) */
void pcop_oracle_eigen33_smallest(const double* m9, double* eval, double* evec3);

#ifdef __cplusplus
}
#endif
#endif
