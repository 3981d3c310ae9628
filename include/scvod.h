/*
 * scvod.h — C-ABI of the B200-native SCV-OD dynamic-removal hot path.
 *
 * The reference (Yixin-F/DR-Using-SCV-OD) has no plugin/FFI layer: the boundary it offers is the
 * public surface of `class SSC` (reference include/ssc.h:55-104) driven by src/main.cpp:9-10.  This
 * header is the seam that sits *under* those methods: a maintainer of the reference re-implements
 * the bodies of SSC::process / segment / recognize / tracking / segDF by calling these entry points
 * (see INTEGRATION.md for the exact stubs), and `PatchWork<PointT>::estimate_ground`
 * (reference include/patchwork.h:278-398) by calling scvod_ground().
 *
 * Conventions
 *   - plain pointers and sizes only; all buffers are caller-owned HOST memory unless the name says _dev
 *   - a scan is an AoS array of float[4] = {x, y, z, intensity} (the four fields of pcl::PointXYZI the
 *     reference reads; reference include/utility.h:96-106, src/ssc.cpp:157-184)
 *   - every function returns 0 on success, <0 on error; scvod_last_error() gives the message
 *   - one context per GPU / host thread (the reference's SSC is non-reentrant: static SSC::id,
 *     reference src/ssc.cpp:28)
 *   - there is NO CPU fallback: every compute entry point fails with SCVOD_ERR_CUDA when no sm_100
 *     device is usable.
 */
#ifndef SCVOD_H_
#define SCVOD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCVOD_OK 0
#define SCVOD_ERR_ARG (-1)
#define SCVOD_ERR_CUDA (-2)
#define SCVOD_ERR_CAPACITY (-3)
#define SCVOD_ERR_STATE (-4)

/* Per-input-point outcome classes (derived from reference
 * include/patchwork.h:302-310,436,331 and src/ssc.cpp:161-172,445-466,479-499,1323-1421). */
enum scvod_point_class {
  SCVOD_PT_DROPPED_LOW = 0,     /* z < -1.8*sensor_height (patchwork.h:302-310)            */
  SCVOD_PT_DROPPED_RANGE = 1,   /* r outside (2.7, 80]    (patchwork.h:436)                */
  SCVOD_PT_DROPPED_SPARSE = 2,  /* patch with <= 10 points (patchwork.h:331)               */
  SCVOD_PT_GROUND = 3,          /* cloud_out of estimate_ground                            */
  SCVOD_PT_GATED_OUT = 4,       /* outside the SSC range/angle/azimuth window (ssc.cpp:161-172) */
  SCVOD_PT_UNCLUSTERED = 5,     /* cluster erased by bounding-box refine (ssc.cpp:445-466) */
  SCVOD_PT_STATIC = 6,          /* member of a cluster whose state != 1                    */
  SCVOD_PT_DYNAMIC = 7          /* member of a cluster with state == 1 (ssc.cpp:1324,1348) */
};

/* The subset of Utility's parameter block that the path reads (reference include/utility.h:209-236,
 * 238-240; defaults at :283-313).  Field names follow the reference (trailing '_' dropped). */
typedef struct scvod_params {
  float sensor_height;
  float min_dis, max_dis;
  float min_angle, max_angle;
  float min_azimuth, max_azimuth;
  float range_res, sector_res, azimuth_res;
  float refine_height;
  float max_z, min_z;
  float car_square;
  int32_t iteration;
  int32_t toBeClass;
  int32_t search_c;
  float intensity_diff;
  float intensity_cov;
  float occupancy;
  int32_t building, tree, car;
} scvod_params;

/* Grid dimensions exactly as SSC::SSC computes them in float (reference src/ssc.cpp:36-39). */
typedef struct scvod_grid {
  int32_t range_num, sector_num, azimuth_num, bin_num;
} scvod_grid;

typedef struct scvod_ctx scvod_ctx;

/* ---- parameter helpers --------------------------------------------------------------------- */
/* config/semantickitti.yaml and config/parkinglot.yaml of the reference (+ utility.h defaults). */
void scvod_params_semantickitti(scvod_params* p);
void scvod_params_parkinglot(scvod_params* p);
int scvod_grid_dims(const scvod_params* p, scvod_grid* g); /* ssc.cpp:36-39 */

/* ---- context ------------------------------------------------------------------------------- */
/* max_points: capacity per scan; max_batch: scans processed per launch group. */
int scvod_create(const scvod_params* p, int device, int max_points, int max_batch, scvod_ctx** out);
int scvod_destroy(scvod_ctx* ctx);
const char* scvod_last_error(void);
int scvod_num_kernel_launches(const scvod_ctx* ctx, int64_t* out); /* kernels launched so far */
/* options: "inspect" (keep per-stage cluster names for scvod_frame_point_cluster; default 1),
 * "host_threads" (threads used for the per-scan cluster bookkeeping), "replay_global" (test hook: run the cluster-name
 * replay with its union-find state in global memory, the variant very dense scans fall back to; default 0), "chain_tma"
 * (stage the plane-fit kernel's point ring with TMA bulk copies + mbarriers instead of per-lane cp.async; same results,
 * measured slower on B200, default 0). */
int scvod_set_option(scvod_ctx* ctx, const char* key, int value);
/* cumulative work counters: "scans", "points", "apri_points", "voxels", "track_pairs", "track_points", "tainted_voxels",
 * "tainted_points"; "tracked_frames" = number of frames already tracked as frame_pre_ (scvod_track resumes there); "reallocs" =
 * device / pinned buffer (re)allocations of the process so far (0 per step in steady state) */
int scvod_get_stat(scvod_ctx* ctx, const char* key, int64_t* out);

/* Run all work of this context on a caller-owned CUDA stream (e.g. torch's current stream). */
int scvod_set_stream(scvod_ctx* ctx, void* cuda_stream);
/* Process-wide per-kernel timing with CUDA events on the launching stream: enable!=0 starts (and
 * resets), the report is text lines "<kernel> <total ms> <launches>". Returns the text length. */
int scvod_kernel_timing(int enable);
int scvod_kernel_timing_report(char* buf, int cap);

/* ---- stage entry points (single scan, host buffers; replace the bodies named on each line) --- */

/* PatchWork::estimate_ground (patchwork.h:278-398) as called by SSC::extractGroudByPatchWork
 * (ssc.cpp:88-96).  Outputs are ORIGINAL point indices in the reference's output order.
 * ground_idx/nonground_idx must hold n ints each. */
int scvod_ground(scvod_ctx* ctx, const float* xyzi, int n, int32_t* ground_idx, int32_t* n_ground,
                 int32_t* nonground_idx, int32_t* n_nonground);

/* SSC::makeApriVec (ssc.cpp:155-195) + Utility polar helpers (utility.h:346-392) on an arbitrary
 * cloud.  pass[i]=1 when the point survives the three gates; the index arrays are written for every
 * point (gated-out points included, as the arithmetic is the same).  Any output may be NULL. */
int scvod_bin(scvod_ctx* ctx, const float* xyzi, int n, uint8_t* pass, int32_t* voxel_idx,
              int32_t* range_idx, int32_t* sector_idx, int32_t* azimuth_idx, float* range,
              float* angle, float* azimuth);

/* Test hook for the binning filter used inside the pipeline kernels: the filtered evaluation (approximate angles + guard band,
 * exact chain for undecided points) must return exactly scvod_bin's indices and gate outcome.  Checks n generated points
 * (xyzi == NULL; coordinates within +-extent, structured edge cases mixed in) or n caller points on the device.
 * stats7 = {points, points that took the exact chain, mismatches (0 expected), max |q_approx - q_exact| * 1e9 for the sector and
 * for the azimuth bin coordinate among the filter-decided points, and for generated points the patch-assignment filter of the
 * ground stage (pc2czm, patchwork.h:431-459): points that took the double chain, mismatches (0 expected)}. */
int scvod_bin_filter_check(scvod_ctx* ctx, const float* xyzi, int64_t n, uint32_t seed, float extent, uint64_t* stats7);

/* ---- frame pipeline (the hot path proper) ---------------------------------------------------- */

/* SSC::process + segment + recognize (ssc.cpp:224-251, 637-656, 834-895) for nscans scans in
 * one batched pass, appending nscans frames to the context's frame_set (ssc.cpp:1435-1444).
 * offsets has nscans+1 entries (point offsets into xyzi). */
int scvod_push_scans(scvod_ctx* ctx, const float* xyzi, const int64_t* offsets, int nscans);

/* Per-scan error isolation: when the last scvod_push_scans / scvod_push_scans_dev call failed because ONE scan exceeded a per-scan
 * capacity (max_points, similarity-edge table, side tables of aliased voxels), the index of that scan inside the call; -1 otherwise.
 * Frames of the failing batch (max_batch consecutive scans) are not kept; earlier batches of the same call are (scvod_num_frames
 * tells how many): drop or re-voxelise the named scan and push the remaining scans again. */
int scvod_last_failed_scan(const scvod_ctx* ctx);

/* Optional double buffering for streams of batches: start the host->device upload of the scans that a LATER scvod_push_scans call
 * will be given (same buffer and offsets).  The copy runs on a private stream and overlaps whatever the context is doing (the
 * tracking chain of the previous batch, typically); that push then finds its points on the device.  The host buffer must stay
 * valid and unchanged until that push returns; pinned memory is needed for the copy to be asynchronous. */
int scvod_prefetch_scans(scvod_ctx* ctx, const float* xyzi, const int64_t* offsets, int nscans);
/* Same as scvod_push_scans, with the scans already resident in device memory (float4 per point). */
int scvod_push_scans_dev(scvod_ctx* ctx, const void* xyzi_dev, const int64_t* offsets, int nscans);

/* SSC::tracking chain of SSC::segDF (ssc.cpp:1448-1452, 1250-1426) over all frames pushed so far
 * and not yet tracked: pairs (tracked, tracked+1) ... (n-2, n-1) with n = min(frames, nposes), in order (get "tracked_frames"
 * through scvod_get_stat).  poses: 6 floats per frame {x,y,z,roll,pitch,yaw} = the Pose fields tracking() reads (utility.h:77-93),
 * entry i belongs to frame i; the poses of EVERY pair that will be tracked by the call must be valid.  A pair is tracked once: the
 * reference would re-run the diff on a repeated call, here it is a no-op. */
int scvod_track(scvod_ctx* ctx, const float* poses6, int nposes);

/* SSC::intialization (ssc.cpp:1148-1248; dead code in the reference, its call is commented out at :1456-1470): every pushed
 * frame is diffed against the "base" frame — the LAST frame with the fewest clusters (:1153-1158) — through
 * trans_based.inverse() * trans_i, and the base-frame clusters that one transformed cluster bridges with a voxel ratio
 * >= occupancy are fused (:1208-1233); recognize() then re-types the clusters of the result (:1241).  Must run before
 * scvod_track (it reads the untracked clouds).  The initialised frame is kept in the context: pass SCVOD_INIT_FRAME as the
 * frame index of scvod_frame_counts / scvod_frame_clusters / scvod_frame_voxels (labels) to read it. */
#define SCVOD_INIT_FRAME (-1)
int scvod_initialization(scvod_ctx* ctx, const float* poses6, int nposes, int32_t* id_based);

int scvod_num_frames(const scvod_ctx* ctx);
int scvod_reset_frames(scvod_ctx* ctx); /* SSC::reset + frame_set.clear() */

/* Per-input-point outcome class (enum scvod_point_class) of frame f; cls holds n_in bytes. */
int scvod_frame_labels(scvod_ctx* ctx, int frame, uint8_t* cls, int n);

/* Labels of frames [f0,f1) concatenated in frame order into one host buffer of `cap` bytes.
 * cls == NULL only brings the device-resident label arrays up to date (no copy). */
int scvod_labels_range(scvod_ctx* ctx, int f0, int f1, uint8_t* cls, int64_t cap);

/* frame inspection — sizes: counts[0]=n_in, [1]=n_ground, [2]=n_nonground, [3]=n_apri (cloud_use),
 * [4]=n_voxels (hash_cloud.size()), [5]=clusters after CVC, [6]=after intensity refine,
 * [7]=after bounding-box refine, [8]=clusters now (after tracking mutations). */
int scvod_frame_counts(scvod_ctx* ctx, int frame, int32_t counts[9]);
int scvod_frame_ground_order(scvod_ctx* ctx, int frame, int32_t* ground_src, int32_t* nonground_src);
/* apri_vec (ssc.cpp:177-193): src = original point index of apri entry m. Any output may be NULL. */
int scvod_frame_apri(scvod_ctx* ctx, int frame, int32_t* src, int32_t* voxel_idx);
/* hash_cloud (ssc.cpp:253-289) sorted by ascending voxel_idx; center is 3 floats per voxel;
 * tri is 3 ints per voxel (range_idx, sector_idx, azimuth_idx of the first inserted point);
 * label is the voxel's cluster label now. */
int scvod_frame_voxels(scvod_ctx* ctx, int frame, int32_t* voxel_idx, int32_t* count, float* av,
                       float* cov, float* center, int32_t* tri, int32_t* label);
/* cluster name of each apri entry at stage 0 (after clusterAndCreateFrame), 1 (after
 * refineClusterByIntensity), 2 (after refineClusterByBoundingBox; -1 = erased). */
int scvod_frame_point_cluster(scvod_ctx* ctx, int frame, int stage, int32_t* name);
/* cluster_set in iteration order: bbox = {min.x,min.y,min.z,max.x,max.y,max.z}. cap = array capacity.  `type` tells car from
 * non-car: every non-car cluster is reported as params.tree (the reference splits large clusters into building / tree with
 * regionGrowing, ssc.cpp:845-856, which no label depends on; scvod_region_growing gives that split for a cluster cloud). */
int scvod_frame_clusters(scvod_ctx* ctx, int frame, int cap, int32_t* name, int32_t* type,
                         int32_t* state, int32_t* npts, int32_t* nvox, float* bbox);

/* Static submap of frames [f0,f1) in the map frame (transformCloud arithmetic with the frame's pose), written to device memory
 * for the NCCL all-gather; returns the point count.  Default = the reference's instance map: the points of every cluster that
 * is not dynamic (SSC::saveSegCloud mode 3, ssc.cpp:446-555: `*instance_map += *rgb_ptr` concatenates cluster points only), i.e.
 * class SCVOD_PT_STATIC.  scvod_set_option("submap_all_static", 1) keeps every input point whose class is not DYNAMIC instead
 * (ground, gated-out and unclustered points included: the dynamic-free scan). */
int scvod_static_submap_dev(scvod_ctx* ctx, int f0, int f1, const float* poses6, void* out_xyzi_dev,
                            int64_t cap_points, int64_t* n_points);

/* ---- GICP scan-to-map stage ------------------------------------------------------------------ */
/* The reference names this stage but holds no code for it: src/gicp.cpp:1-57 is a PCD merge tool,
 * src/ssc.cpp:1458,1467 are commented-out "TODO: gicp" lines.  Definition: docs/gicp_spec.md
 * (plane-to-plane Generalized-ICP, Gauss-Newton, fixed-radius neighbourhoods in a uniform grid). */
typedef struct scvod_gicp_params {
  float cov_radius;     /* r: neighbourhood radius of the per-point covariance (m)      */
  float max_corr_dist;  /* d_max: correspondence gate (m)                               */
  float cov_eps;        /* epsilon of the (1,1,eps) covariance regularisation           */
  float planarity;      /* a point is usable iff lambda2 <= planarity * lambda1          */
  int32_t min_neighbors;
  int32_t max_iter;
  float rot_eps, trans_eps; /* convergence thresholds on |omega| (rad) and |v| (m)      */
} scvod_gicp_params;

typedef struct scvod_gicp_result {
  float T[12];     /* source -> target, row-major 3x4                                           */
  float pose6[6];  /* {x,y,z,roll,pitch,yaw}: Utility::rotationMatrixToEulerAngles convention   */
                   /* (utility.h:488-505), i.e. pcl::getTransformation(pose6) == T              */
  double H[36];    /* normal equations of the last evaluated iteration                          */
  double b[6];
  double cost;
  int32_t iterations, n_corr, converged, n_src_valid, n_tgt_valid;
} scvod_gicp_result;

void scvod_gicp_default_params(scvod_gicp_params* p);
/* Target (map) side: grid + per-point normals, kept on the device until replaced. */
int scvod_gicp_set_target(scvod_ctx* ctx, const float* tgt_xyzi, int n, const scvod_gicp_params* p);
int scvod_gicp_set_target_dev(scvod_ctx* ctx, const void* tgt_xyzi_dev, int n, const scvod_gicp_params* p);
/* Align one source scan to the current target starting from T0 (12 floats, row-major 3x4). */
int scvod_gicp_align(scvod_ctx* ctx, const float* src_xyzi, int n, const float T0[12], scvod_gicp_result* out);
int scvod_gicp_align_dev(scvod_ctx* ctx, const void* src_xyzi_dev, int n, const float T0[12], scvod_gicp_result* out);
/* The radius-search kernel alone (docs/gicp_spec.md section 3): unit normal, validity and neighbour
 * count of every point of one cloud, in input order.  Any output may be NULL. */
int scvod_gicp_normals(scvod_ctx* ctx, const float* xyzi, int n, const scvod_gicp_params* p, float* normals3,
                       uint8_t* valid, int32_t* count);
/* pcl::getTransformation(x,y,z,roll,pitch,yaw) as 12 floats (the matrix SSC::tracking builds, ssc.cpp:1255). */
void scvod_pose_matrix(const float pose6[6], float T[12]);

/* Per-patch plane fits of the most recent ground pass of batch slot `slot`: 504 rows of 12 floats
 * {normal[3], mean[3], singular values[3], d, decision, npts}; rows of skipped patches are stale. */
int scvod_last_patch_records(scvod_ctx* ctx, int slot, float* rec504x12);
/* The device port of glibc's atan2f evaluated on the GPU (libm-parity test hook). */
int scvod_atan2f_device(scvod_ctx* ctx, const float* y, const float* x, float* out, int64_t n);

/* Host-only: the cluster bookkeeping of SSC::segment + recognize (ssc.cpp:299-393, 571-635, 437-467,
 * 834-895) on caller-provided voxel tables (the tables the GPU stages produce).  Needs no device; used
 * by the CPU test-suite.  Returns the number of clusters (<0 on error). */
int scvod_host_segment(const scvod_params* p, int V, const int32_t* vox_cnt, const int32_t* vox_root,
                       const int32_t* vox_nbr, const float* vox_bbox, int n_events, const int32_t* ev_cid,
                       int n_edges, const int32_t* edges, int32_t* name_stage0, int32_t* name_stage1,
                       int32_t* name_stage2, int32_t n_clusters[3], int cap, int32_t* cluster_name,
                       int32_t* cluster_type, int32_t* max_name);

/* The same with "tainted" voxels: voxels that hold a point whose range / sector / azimuth index is -1 (dis == min_dis,
 * y == 0 with x > 0, azimuth == min_azimuth; ssc.cpp:185-188).  Such a point hashes into a voxel that is not its own cell, so
 * the voxel's points no longer share one neighbour list and clusterAndCreateFrame is replayed point by point for them:
 * tv_cid = the tainted voxels (ascending), points of voxel i = tp_*[tv_base[i] .. tv_base[i+1]) in ascending apri index
 * (tp_m), tp_nbr = each point's own findVoxelNeighbors list (27 compact ids, -1 = absent), an event ev_cid >= V is point
 * ev_cid - V of those tables (every point of a tainted voxel is an event).  tp_stage receives [3][n_tpts] cluster names. */
int scvod_host_segment_pts(const scvod_params* p, int V, const int32_t* vox_cnt, const int32_t* vox_root,
                           const int32_t* vox_nbr, const float* vox_bbox, int n_events, const int32_t* ev_cid,
                           int n_edges, const int32_t* edges, int n_tvox, const int32_t* tv_cid, const int32_t* tv_base,
                           int n_tpts, const int32_t* tp_m, const float* tp_xyz, const int32_t* tp_nbr,
                           int32_t* name_stage0, int32_t* name_stage1, int32_t* name_stage2, int32_t* tp_stage,
                           int32_t n_clusters[3], int cap, int32_t* cluster_name, int32_t* cluster_type,
                           int32_t* cluster_npts, int32_t* cluster_nvox, int32_t* max_name);

/* ---- loader front end on the device ------------------------------------------------------------------
 * What SSC::getCloud does to a SemanticKITTI scan before process() sees it (reference src/ssc.cpp:1060-1111): points whose label
 * (low 16 bits) is 0 or 1 are dropped (:1063), intensity is multiplied by max_intensity (:1071), and the cloud goes through
 * pcl::VoxelGrid<PointXYZI> with leaf 0.08 m (:1108-1111; PCL 1.8 voxel_grid.hpp: one centroid per occupied leaf, in ascending
 * leaf index).  raw_xyzi = the .bin file contents (x, y, z, intensity in [0, 1]), labels = the .label file contents (NULL: keep
 * every point, as for PCD input).  Mask, leaf indices, output order and leaf populations are exact; a centroid is the float sum
 * of its leaf's points in ascending input index (the reference leaves that order to std::sort, which is unstable), so leaves
 * with >= 3 points may differ from the reference in the last bits.
 *   scvod_load_kitti      host buffers; out_xyzi (room for every input point) receives the scans packed one after the other and
 *                         out_offsets[nscans + 1] their boundaries: both can be handed to scvod_push_scans as they are.
 *   scvod_load_kitti_dev  device buffers; scan b is written at out + (offsets[b] - offsets[0]); counts4[4 b] = output points,
 *                         [4 b + 1] = points that passed the mask, [4 b + 2] = 1 if the leaf grid overflowed an int and the
 *                         masked scan was passed through unchanged (PCL's "leaf size is too small" path). */
int scvod_load_kitti(scvod_ctx* ctx, const float* raw_xyzi, const uint32_t* labels, const int64_t* offsets, int nscans, float leaf,
                     float max_intensity, float* out_xyzi, int64_t* out_offsets);
int scvod_load_kitti_dev(scvod_ctx* ctx, const void* raw_xyzi_dev, const uint32_t* labels_dev, const int64_t* offsets, int nscans,
                         float leaf, float max_intensity, void* out_xyzi_dev, int32_t* counts4);

/* ---- quality measures on the device ---------------------------------------------------------------------
 * Clouds are host arrays of float[4] points.  For scvod_evaluate_map the 4th float is the SemanticKITTI label stored as a float (what
 * the reference writes into PCD intensity, src/ssc.cpp:1079; low 16 bits = semantic class), dynamic_classes the moving classes
 * (DYNAMIC_CLASSES of tool/analysis.py:6 = the YAML's dynamic_label_).  gt = ground-truth map, est = estimated static map.
 * Restates evaluate() + calc_naive_preservation() (tool/analysis.py:124-194): nearest neighbour of every gt point in est; preserved
 * when closer than voxelsize * sqrt(3) / 2.  nn_index (may be NULL) receives that neighbour's index per gt point, -1 if none. */
typedef struct scvod_eval_result {
  int64_t gt_static, gt_dynamic, est_static, est_dynamic;
  int64_t preserved, static_preserved, dynamic_preserved;
  double preservation_rate, rejection_rate, f1; /* PR %, RR %, F1 (tool/analysis.py:186-188) */
  int64_t gt_per_class[8], est_per_class[8];    /* points of every dynamic class (the "R. R" table, :158-165) */
} scvod_eval_result;
int scvod_evaluate_map(scvod_ctx* ctx, const float* gt_xyzl, int64_t n_gt, const float* est_xyzl, int64_t n_est, float voxelsize,
                       const int32_t* dynamic_classes, int n_classes, scvod_eval_result* out, int32_t* nn_index);
/* evaluate() of src/evaluate.cpp:79-145.  pred = xyz + "predicted static" flag (4th float != 0, the reference's ori.g != 0), the two
 * ground-truth clouds are searched with r_hit (0.15) first and r_miss (0.1) second.  counts5 = {TP, FN (predicted static, dynamic
 * point nearby), TN, FN (predicted dynamic, static point nearby), not shown}; per_point (may be NULL) the class of every point. */
int scvod_evaluate_confusion(scvod_ctx* ctx, const float* pred_xyzs, int64_t n, const float* static_gt_xyz, int64_t ns,
                             const float* dynamic_gt_xyz, int64_t nd, float r_hit, float r_miss, int64_t counts5[5], uint8_t* per_point);

/* ---- k-NN normals, intensity calibration, region growing ---------------------------------------------
 * scvod_knn_normals          exact k nearest neighbours (1 <= k <= 16, the point itself included, sorted by distance) of every point
 *                            of a host cloud in the cloud itself, the PCA normal of the neighbourhood turned towards the origin and
 *                            its curvature lambda_min / trace: what pcl::NormalEstimation with setKSearch(k) computes.  Any output
 *                            may be NULL.  Normals are checked by tolerance (PCL's float eigen33 is not reproduced bit for bit).
 * scvod_calibrate_intensity  SSC::intensityCalibrationByCurvature (src/ssc.cpp:98-153) in place on a host cloud (float[4] points).
 * scvod_region_growing       SSC::regionGrowing (src/ssc.cpp:797-832) on a cluster cloud: *is_building = 1 when the planar segments hold
 *                            >= 20 % of the points; segment_of (may be NULL) = segment of every point, planar_points (may be NULL) =
 *                            points in segments of >= 20 points.  recognize() needs it only to tell building from tree (:845-856). */
int scvod_knn_normals(scvod_ctx* ctx, const float* xyzi, int n, int k, float* normals3, float* curvature, int32_t* neighbors);
int scvod_calibrate_intensity(scvod_ctx* ctx, float* xyzi, int n, int search_num, float max_intensity);
int scvod_region_growing(scvod_ctx* ctx, const float* xyzi, int n, int32_t* is_building, int32_t* segment_of, int32_t* planar_points);

/* ---- chain hand-off between contexts (one unbroken tracking chain over a sequence cut into chunks) ------------------------
 * SSC::segDF tracks a whole sequence as ONE chain: tracking(frame_set[i], frame_set[i+1]) for every i (ssc.cpp:1450-1452).  When the
 * sequence is cut into chunks owned by different contexts (workers of one GPU, GPUs of a box, processes), the pair that straddles a
 * cut is run by the context that owns the later chunk: tracking() only reads, of frame_pre_, the clouds of its car clusters in
 * cluster_set order and writes their state / type back (ssc.cpp:1261-1275, 1324-1397).
 *   scvod_export_tail      serialises that view of the context's LAST frame (after its own frames have been tracked) into host memory:
 *                          int32 {magic, ncars, npts, SSC::name}, ncars x int32 {name, track_id, npts, 0}, npts x float4 xyzi
 *                          (own points part by part, then the clouds carried along the chain).  buf == NULL: size query.
 *   scvod_track_from_tail  runs tracking(tail, frame 0) in the context that holds the next chunk (before its own scvod_track): frame 0 is
 *                          mutated exactly as in the unbroken chain (splits, fusions, carried clouds, track ids, SSC::name).  Writes
 *                          (state, type) of every exported car cluster to state_type[2 * ncars]; returns ncars.
 *   scvod_apply_tail_states stores those back into the exporting context's last frame (its labels are refreshed on the next read). */
int scvod_export_tail(scvod_ctx* ctx, void* buf, size_t cap, size_t* nbytes);
int scvod_track_from_tail(scvod_ctx* ctx, const void* tail, size_t nbytes, const float pose_pre6[6], const float pose_next6[6],
                          int32_t* state_type, int cap);
int scvod_apply_tail_states(scvod_ctx* ctx, const int32_t* state_type, int n);

/* ---- helpers shared by tests and the bench ---------------------------------------------------- */
/* trans_next.inverse() * trans_pre of SSC::tracking (ssc.cpp:1255-1257) as 12 floats row-major 3x4. */
void scvod_relative_pose(const float pose_next6[6], const float pose_pre6[6], float T[12]);

/* Deterministic synthetic 64-beam-style scan: rings x cols rays cast into a
 * procedural street scene. Writes up to rings*cols points; returns the count in *n and the ego
 * pose in pose6 (may be NULL). No libm calls: identical bytes on every host. */
int scvod_synth_scan(uint64_t seed, int scan_id, int rings, int cols, float* xyzi, int* n,
                     float* pose6);
/* The same scan with a SemanticKITTI-style label per point (low 16 bits: 40 road, 10 car, 50 building, 80 pole, 71 trunk,
 * 70 vegetation, 252 moving car, 1 outlier; high 16 bits: object id): ground truth for the quality measures. */
int scvod_synth_scan_labeled(uint64_t seed, int scan_id, int rings, int cols, float* xyzi, int* n, float* pose6, uint32_t* labels);

#ifdef __cplusplus
}
#endif
#endif /* SCVOD_H_ */
