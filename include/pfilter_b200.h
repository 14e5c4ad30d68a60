/* pfilter_b200 -- C ABI of the B200-native (sm_100a) PFilter hot path.
 *
 * The reference (kevrenhype/PFilter-noetic) has no FFI: its boundary is the C++ class API that the
 * ROS node loops call once per frame.  Each entry point below names the
 * reference interface it replaces (file:line under /root/reference).  The C++ classes with the
 * reference's names (headers under include/pfilter_b200/) are header-only wrappers over this ABI.
 *
 * Conventions
 *   - plain C, POD only; every function returns PF_OK (0) or a negative pf_status; never throws.
 *   - caller-owned host buffers; library-owned device state; one handle <-> one CUDA device + stream
 *     <-> one host thread (same threading contract as the reference: one caller thread per object).
 *   - points are 16 bytes.  Scan points: float4 {x, y, z, intensity}.  Map / feature points:
 *     pf_point {x, y, z, r, g, b, a} where r = "round" and g = "observe" are PFilter's persistence
 *     counters (the reference keeps them in the RGB channels of pcl::PointXYZRGB,
 *     include/odomEstimationClass.h:38).
 *   - there is NO CPU fallback: every entry point runs CUDA kernels on the handle's device or fails.
 */
#ifndef PFILTER_B200_H_
#define PFILTER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_VERSION 101

typedef enum pf_status {
    PF_OK = 0,
    PF_ERR_INVALID = -1,    /* bad argument / unsupported configuration */
    PF_ERR_CUDA = -2,       /* CUDA runtime error (see pf_last_error) */
    PF_ERR_CAPACITY = -3,   /* an input exceeded a capacity fixed at create time */
    PF_ERR_STATE = -4       /* call order violated (e.g. update before init_map) */
} pf_status;

/* lidar::Lidar, include/lidar.h:9-32 (only the fields the hot path reads). */
typedef struct pf_lidar_params {
    int32_t num_lines;      /* 16, 32 or 64 (src/laserProcessingClass.cpp:30-61) */
    double min_distance;    /* launch default 3  (launch/pfilter_kitti.launch:57-58) */
    double max_distance;    /* launch default 90 */
    double scan_period;     /* unused on the hot path */
} pf_lidar_params;

typedef struct pf_point {
    float x, y, z;
    uint8_t r, g, b, a;
} pf_point;

int pf_version(void);
const char* pf_last_error(void);          /* thread-local text of the last failure */
int pf_device_count(void);
/* pinned host memory for scan / result buffers (cudaMemcpyAsync from pageable memory is staged by the driver) */
int pf_host_alloc(void** p, uint64_t bytes);
int pf_host_free(void* p);

/* ------------------------------------------------------------------------------------------------
 * Scan packing (host side of the boundary): the float4 layout above from the forms the reference's callers hold a scan in.
 * xyzi_out is any host buffer of cap_points x 16 bytes -- normally one from pf_host_alloc, so that the packed scan is what the
 * H2D copy of pf_frame_process / pf_frame_submit reads.
 * ---------------------------------------------------------------------------------------------- */
/* sensor_msgs/PointCloud2 as the nodes receive it and pcl::fromROSMsg turns it into PointXYZI (src/laserProcessingNode.cpp:52-63):
 * offsets and sensor_msgs/PointField datatype codes (1 INT8 .. 7 FLOAT32, 8 FLOAT64) of the x, y, z, intensity fields;
 * off_intensity < 0 = no such field (intensity 0, as fromROSMsg leaves it); row_step 0 = point_step x width. */
typedef struct pf_pc2_layout {
    uint32_t point_step, row_step, width, height;
    int32_t off_x, off_y, off_z, off_intensity;
    uint8_t type_x, type_y, type_z, type_intensity;
    uint8_t is_bigendian;     /* must be 0 */
} pf_pc2_layout;
int pf_pack_pointcloud2(const uint8_t* data, uint64_t data_bytes, const pf_pc2_layout* layout, float* xyzi_out, int cap_points, int* n_points);
/* KITTI odometry velodyne/NNNNNN.bin (little-endian float32 x, y, z, reflectance): PF_ERR_CAPACITY when the file holds more points */
int pf_read_kitti_bin(const char* path, float* xyzi_out, int cap_points, int* n_points);
int pf_write_kitti_bin(const char* path, const float* xyzi, int n_points);

/* ------------------------------------------------------------------------------------------------
 * Feature extraction  --  replaces LaserProcessingClass (include/laserProcessingClass.h:32-41)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pf_extract pf_extract;

typedef struct pf_extract_config {
    int32_t max_points;        /* capacity per scan (points); rounded up to a multiple of 256 */
    int32_t max_batch;         /* scans per batched launch */
    int32_t max_ring_points;   /* capacity of one ring; sizes the per-warp shared memory (a sector of the ring) and so the
                                  occupancy of the extract kernel: set it to the sensor's points per ring plus a margin;
                                  0 = default 2304, maximum 3040 */
    int32_t surf_order;        /* order of the surf points inside a sector: 0 = ascending ring position (default, fastest), 1 = the
                                  reference's: ascending curvature, the order of its sorted walk (src/laserProcessingClass.cpp:101-104,
                                  :198-205; equal curvatures: lower ring position first).  The SETS are identical either way. */
} pf_extract_config;

/* LaserProcessingClass::init, src/laserProcessingClass.cpp:4-8 */
int pf_extract_create(const pf_lidar_params* lidar, const pf_extract_config* cfg, int device, pf_extract** out);
int pf_extract_destroy(pf_extract* h);

/* LaserProcessingClass::featureExtraction, src/laserProcessingClass.cpp:10-96 (+ :99-209).
 * xyzi: n host points (float4).  edge/surf receive the selected points (16 B each, same float4 layout) in
 * emission order ring -> sector -> {edges by descending curvature; surf by ascending ring position, or -- with
 * pf_extract_config.surf_order = 1 -- by ascending curvature as the reference emits them}; the SETS equal the reference's.
 * label (optional, n bytes): 0 = neither, 1 = edge, 2 = surf, per input index.
 * Capacities: edge needs 120 * num_lines points, surf needs n points. */
int pf_extract_run(pf_extract* h, const float* xyzi, int n, float* edge, int* n_edge, float* surf, int* n_surf,
                   uint8_t* label);

/* Batched form (extraction has no temporal dependency): scan s occupies xyzi[s*stride*4 ...] with n[s] points.
 * edge of scan s at edge[s*edge_stride*4], surf of scan s at surf[s*stride*4]; label at label[s*stride]. */
int pf_extract_run_batch(pf_extract* h, const float* xyzi, const int* n, int batch, int stride, float* edge,
                         int* n_edge, int edge_stride, float* surf, int* n_surf, uint8_t* label);

/* Device-resident form: all pointers are device pointers on the handle's device; work is enqueued on the
 * handle's stream and NOT synchronised (use pf_extract_sync).  d_label may be NULL.  stride must be a multiple of 256. */
int pf_extract_run_batch_device(pf_extract* h, const void* d_xyzi, const int* d_n, int batch, int stride,
                                void* d_edge, int* d_n_edge, int edge_stride, void* d_surf, int* d_n_surf,
                                uint8_t* d_label);
int pf_extract_sync(pf_extract* h);
void* pf_extract_stream(pf_extract* h);   /* cudaStream_t of the handle (for CUDA-event timing by the harness) */
int pf_extract_kernel_launches(pf_extract* h, uint64_t* launches);   /* kernels launched so far by this handle */

/* ------------------------------------------------------------------------------------------------
 * Odometry + persistence filter + local map  --  replaces Odom_ES_EstimationClass
 * (include/odomEstimationClass.h:140-167, src/odomEstimationClass.cpp:7-647, src/lidarOptimization.cpp)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pf_odom pf_odom;

typedef struct pf_odom_params {
    double map_resolution;     /* edge leaf; surf leaf = 2x (src/odomEstimationClass.cpp:189-190); launch default 0.4 */
    int32_t k_new;             /* PFilter parameters (launch/pfilter_kitti.launch:59-64) */
    float theta_p;
    int32_t theta_max;
    double weight_type;        /* 0 (class default), 1 observe, 2 sparsity, 12 both (src/odomEstimationClass.cpp:389-423); other values are rejected */
    int32_t max_map_points;    /* capacity of EACH local map (edge, surf); 0 = default 2M */
    int32_t max_features;      /* capacity of each per-frame feature cloud passed in; 0 = default 131072 */
} pf_odom_params;

/* Odom_ES_EstimationClass::init, src/odomEstimationClass.cpp:182-208 */
int pf_odom_create(const pf_odom_params* p, int device, pf_odom** out);
int pf_odom_destroy(pf_odom* h);

/* Feature clouds are host arrays of 16-byte points {x, y, z, <ignored>}: like pcl::copyPointCloud(XYZI -> XYZRGB)
 * at src/odomEstimationNode copy.cpp:74-80 only x, y, z are taken; counters start at r = g = b = 0, a = 255. */

/* initMapWithPoints, src/odomEstimationClass.cpp:217-222 */
int pf_odom_init_map(pf_odom* h, const float* edge, int n_edge, const float* surf, int n_surf);
/* updatePointsToMap, src/odomEstimationClass.cpp:229-282.  pose_out = [qx qy qz qw tx ty tz] (the public member
 * `odom`, include/odomEstimationClass.h:57, read by the node at src/odomEstimationNode.cpp:144-146). */
int pf_odom_update(pf_odom* h, const float* edge, int n_edge, const float* surf, int n_surf, double pose_out[7]);
int pf_odom_get_pose(pf_odom* h, double pose[7]);
/* public members laserCloudCornerMap / laserCloudSurfMap (include/odomEstimationClass.h:151-152): which = 0 edge(corner), 1 surf */
int pf_odom_map_size(pf_odom* h, int which, int* n);
int pf_odom_get_map_part(pf_odom* h, int which, pf_point* out, int cap, int* n);
/* getMap, src/odomEstimationClass.cpp:210-215: surf map followed by corner map */
int pf_odom_get_map(pf_odom* h, pf_point* out, int cap, int* n);
/* pose after every outer iteration of the last update (optimization_count entries of 7 doubles) */
int pf_odom_get_iter_poses(pf_odom* h, double* poses, int cap, int* n);
/* per-update diagnostics of the last update: down-sampled query counts and valid residual counts of the last pass */
typedef struct pf_odom_stats {
    int32_t n_edge_ds, n_surf_ds;          /* queries after VoxelGrid down-sampling */
    int32_t n_edge_res, n_surf_res;        /* residual blocks in the last outer iteration */
    int32_t map_edge, map_surf;            /* map sizes after the update */
    int32_t passes;                        /* optimization_count used */
    int32_t lm_iterations;                 /* LM step attempts in the last outer iteration */
} pf_odom_stats;
int pf_odom_get_stats(pf_odom* h, pf_odom_stats* s);
void* pf_odom_stream(pf_odom* h);
/* CUDA-event phase timing of the last update (PF_ODOM_TIMING=1 at create time): ms[0] predict + down-sample, ms[1] search-grid
 * build, ms[2] optimisation passes up to the last association, ms[3] the last pass's 5 LM evaluations, ms[4] append + map update */
int pf_odom_get_phase_ms(pf_odom* h, float ms[5]);
int pf_odom_kernel_launches(pf_odom* h, uint64_t* launches);
/* From the frame at which optimization_count has settled at 2, the launch sequence of an update is replayed as a CUDA graph (one
 * per map ping-pong buffer; PF_ODOM_GRAPH=0 turns this off).  *n = number of captures so far. */
int pf_odom_graph_captures(pf_odom* h, int* n);

/* ------------------------------------------------------------------------------------------------
 * Odom_BPF_EstimationClass (include/odomEstimationClass.h:169-205, src/odomEstimationClass.cpp:649-1306): the reference's
 * second odometry class -- the same arithmetic over three feature kinds: beam and pillar (point-to-line residuals, leaf =
 * map_resolution) and facade (point-to-plane, leaf = 2 x map_resolution).  Same handle type and accessors as above:
 * pf_odom_get_pose, pf_odom_map_size / pf_odom_get_map_part (which = 0 beam, 1 pillar, 2 facade), pf_odom_get_map (beam, pillar,
 * facade order of :683-689), pf_odom_get_iter_poses, pf_odom_get_stats (edge fields = beam, surf fields = facade; residual
 * counts = line-type and plane-type totals), pf_odom_destroy.
 * ---------------------------------------------------------------------------------------------- */
int pf_odom_bpf_create(const pf_odom_params* p, int device, pf_odom** out);                       /* init :649-681 */
int pf_odom_bpf_init_map(pf_odom* h, const float* beam, int n_beam, const float* pillar, int n_pillar, const float* facade,
                         int n_facade);                                                            /* initMapWithPoints :692-698 */
int pf_odom_bpf_update(pf_odom* h, const float* beam, int n_beam, const float* pillar, int n_pillar, const float* facade, int n_facade,
                       double pose_out[7]);                                                        /* updatePointsToMap :706-760 */

/* Device-resident hand-off: consume the outputs of the last single-scan extraction of `ex` (same device)
 * without a host round trip.  Frame 0 initialises the map, later frames update it. */
int pf_odom_process_extracted(pf_odom* h, pf_extract* ex, double pose_out[7]);
/* Whole frame: H2D scan -> extract -> (init | update) -> pose D2H.  One call per frame. */
int pf_frame_process(pf_extract* ex, pf_odom* od, const float* xyzi, int n, double pose_out[7]);
/* Same with the scan already resident on the handle's device.  pose_out == NULL: enqueue only, no host
 * synchronisation (frames can be queued back to back; errors surface at pf_odom_sync). */
int pf_frame_process_device(pf_extract* ex, pf_odom* od, const void* d_xyzi, int n, double* pose_out);
int pf_odom_sync(pf_odom* h);
int pf_odom_result_bytes(void);   /* bytes read back from the device per frame (pose, map sizes, error bits) */
/* Pipelined whole frame from a HOST scan: submit enqueues H2D + extraction + odometry and returns at once with the frame's id
 * (0 = the init frame); wait blocks until that frame is done and returns its pose.  Submitting frame k+1 before waiting for
 * frame k overlaps the upload / extraction of the next scan with the odometry of the current one (the reference overlaps them
 * by running its nodes as separate processes).  xyzi must stay valid -- and should be pinned (pf_host_alloc) -- until the
 * frame has been waited for; at most 32 frames may be outstanding. */
int pf_frame_submit(pf_extract* ex, pf_odom* od, const float* xyzi, int n, long long* frame_id);
int pf_frame_wait(pf_odom* od, long long frame_id, double pose_out[7]);
/* poses of updates first_frame .. first_frame+count-1 (frame 0 is the init frame; history depth 4096 frames) */
int pf_odom_get_pose_history(pf_odom* h, long long first_frame, int count, double* poses);

/* ------------------------------------------------------------------------------------------------
 * Stage taps (parity tests and micro-benchmarks); host buffers, synchronous.
 * ---------------------------------------------------------------------------------------------- */
/* pcl::VoxelGrid<PointXYZRGB>::filter as used by downSamplingToMap (src/odomEstimationClass.cpp:176-180, 244-245):
 * key = floor(p * (1/leaf)) - min_b, output ascending key, centroid = sum in ascending input index / n. */
int pf_voxel_downsample(int device, const pf_point* in, int n, float leaf, pf_point* out, int* n_out);
/* addPointsToMap's map maintenance (src/odomEstimationClass.cpp:606-647): CropBox [center-100, center+100],
 * rgbds(leaf) (:34-134), extractstablepoint (:7-25), r += 2 saturating (:634-646). */
int pf_map_update(int device, const pf_point* in, int n, const double center[3], float leaf, int k_new, float theta_p,
                  int theta_max, pf_point* out, int* n_out);
/* The same map maintenance in its streaming form (what pf_odom_update runs from the second update on): `sorted_map` is the
 * sorted part of a map as a previous update left it (ascending voxel key, one point per voxel), `extra` the unsorted
 * points (exceptions + the points appended by the frame).  Result = pf_map_update(sorted_map ++ extra), except that
 * centroids which left their voxel by float rounding are placed behind the sorted part: out[0, *n_sorted_out) is sorted. */
int pf_map_merge(int device, const pf_point* sorted_map, int m_sorted, const pf_point* extra, int n_extra, const double center[3],
                 float leaf, int k_new, float theta_p, int theta_max, pf_point* out, int cap_out, int* n_out, int* n_sorted_out);
/* Micro-benchmark form of pf_map_merge: inputs are uploaded once, the merge runs `reps` times on the device (the first
 * repetition is a warm-up when reps > 1); ms_total = whole pipeline per repetition, ms_stream = the streaming kernel
 * (k_mm_merge) alone, both CUDA-event timed on the library's stream.  No output copy. */
int pf_map_merge_timed(int device, const pf_point* sorted_map, int m_sorted, const pf_point* extra, int n_extra, const double center[3],
                       float leaf, int k_new, float theta_p, int theta_max, int reps, int* n_out, int* n_sorted_out, float* ms_total,
                       float* ms_stream);
/* KdTreeFLANN::nearestKSearch(k = 5) (src/odomEstimationClass.cpp:299, 447): exact, float L2_Simple distances,
 * ascending, ties by lower index.  Contract: results are exact whenever d2[5q+4] < 1.0 (the only case the
 * reference uses, :300/:451); otherwise idx[5q..] = -1 and d2 = +inf. */
int pf_knn5(int device, const pf_point* map, int m, const float* queries_xyz4, int q, int32_t* idx, float* d2);
/* Micro-benchmark form: grid build + query kernel repeated `reps` times on the device, CUDA-event timed separately. */
int pf_knn5_timed(int device, const pf_point* map, int m, const float* queries_xyz4, int q, int32_t* idx, float* d2, int reps,
                  float* ms_build, float* ms_query);
/* One association pass (addEdgeCostFactor :284-432 / addSurfCostFactor :434-578) at a given pose, weightType 0:
 * per query: flag (0 none, 1 geometric fit ok but skipped by the persistence rule, 2 residual added),
 * geometry (edge: a[3], b[3]; surf: n[3], d; 8 doubles per query, unused slots 0), query r,g after the pass;
 * map g counters are updated in place in `map` (sequential semantics). */
int pf_associate(int device, int kind /*0 edge, 1 surf*/, pf_point* map, int m, pf_point* queries, int q,
                 const double pose[7], int k_new, float theta_p, int theta_max, uint8_t* flag, double* geom8);
/* Residual + Jacobian + Huber + normal equations at a pose (src/lidarOptimization.cpp:12-78 + ceres::HuberLoss(0.1)):
 * edge residual i: p=[9i..9i+2], a=[+3..+5], b=[+6..+8]; surf residual j: p=[7j..], n=[+3..+5], d=[+6].
 * H21 = upper triangle of sum J^T J (row-major), g6 = sum J^T r, cost = 1/2 sum rho. */
int pf_eval_normal_eq(int device, const double pose[7], const double* edge9, int n_edge, const double* surf7, int n_surf,
                      double H21[21], double g6[6], double* cost);
/* The same sums by the grid-wide streaming kernel that serves the map-size sweep of BASELINE.json configs[4] (10^6 .. 5 x 10^7
 * residual blocks; pf_eval_normal_eq itself switches to it above 262144 blocks).  `reps` repetitions on the device (the first is a
 * warm-up when reps > 1); ms_kernel = mean CUDA-event time of the kernel alone. */
int pf_eval_normal_eq_timed(int device, const double pose[7], const double* edge9, int n_edge, const double* surf7, int n_surf,
                            int reps, double H21[21], double g6[6], double* cost, float* ms_kernel);
/* One ceres::Solve equivalent (src/odomEstimationClass.cpp:263-271) on fixed residuals. */
int pf_lm_solve(int device, double pose_io[7], const double* edge9, int n_edge, const double* surf7, int n_surf,
                int* iterations, double* final_cost);

/* ------------------------------------------------------------------------------------------------
 * Global map  --  replaces LaserMappingClass (include/laserMappingClass.h:32-58, src/laserMappingClass.cpp)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pf_mapping pf_mapping;

/* LaserMappingClass::init(map_resolution), src/laserMappingClass.cpp:7-32.  max_map_points: capacity of the global map
 * (0 = 16 Mi points); max_points: capacity of one input cloud (0 = 262144). */
int pf_mapping_create(double map_resolution, int max_map_points, int max_points, int device, pf_mapping** out);
int pf_mapping_destroy(pf_mapping* h);
/* updateCurrentPointsToMap(pc_in, pose_current), src/laserMappingClass.cpp:152-191.  xyzi: n host points (float4, sensor
 * frame); rt: the Eigen::Isometry3d pose as row-major 3x4 [R | t] doubles.  Points that fall outside the 5x5x5 block of
 * 50 m cells around the pose (the reference indexes a NULL / out-of-range cell there) are dropped and counted.
 * The call returns when the work is enqueued; the next call on the handle waits for it. */
int pf_mapping_update(pf_mapping* h, const float* xyzi, int n, const double rt[12]);
int pf_mapping_update_device(pf_mapping* h, const void* d_xyzi, int n, const double rt[12]);
int pf_mapping_map_size(pf_mapping* h, int* n);
/* getMap, src/laserMappingClass.cpp:196-208: every cell in (x, y, z) cell order, each cell in VoxelGrid order; the few
 * centroids that rounding pushed out of their voxel in the last update trail the array until the next update. */
int pf_mapping_get_map(pf_mapping* h, float* xyzi_out, int cap, int* n);
/* n_sorted: length of the sorted part; dropped: points dropped so far (see pf_mapping_update); launches: kernels so far */
int pf_mapping_stats(pf_mapping* h, int* n_sorted, long long* dropped, uint64_t* launches);
void* pf_mapping_stream(pf_mapping* h);

#ifdef __cplusplus
}
#endif
#endif /* PFILTER_B200_H_ */
