// ssc.h — `class SSC` for the B200 path.
//
// The class offers the public surface of the reference's SSC (reference include/ssc.h:7-105: same base class, same public
// data members, same method names and signatures), so the reference's src/main.cpp (`SSC ssc; ssc.segDF();`,
// main.cpp:9-10) and code written against those members build unchanged.  The bodies live in host/src/ssc.cpp and
// forward to the C-ABI of libscvod_b200.so (include/scvod.h); INTEGRATION.md has the method -> entry point table.
// Declarations are grouped by what they map to in that C-ABI; the aliases below are plain typedefs, the signatures are
// the reference's.
#ifndef SSC_H_
#define SSC_H_

#include "patchwork.h"
#include "utility.h"

struct scvod_ctx;

class SSC : public Utility {
 public:
  using CloudI = pcl::PointCloud<pcl::PointXYZI>;
  using CloudIPtr = CloudI::Ptr;
  using CloudRGBPtr = pcl::PointCloud<pcl::PointXYZRGB>::Ptr;
  using VoxelMap = std::unordered_map<int, Voxel>;

  SSC();
  ~SSC();
  void allocateMemory();
  void reset();  // clears the per-scan members only (ssc.cpp:79-86); the frames stay in the library context

  // ---- the run: loaders, then segDF() = batched scvod_push_scans + scvod_track + scvod_labels_range -------------------
  void getPose();   // poses.txt (tr^-1 * cam * tr) or the synthetic generator -> pose_vec, trans_vec
  void getCloud();  // KITTI .bin/.label (+ 0.08 m voxel grid) or the synthetic generator -> cloud_vec
  void segDF();
  std::vector<CloudIPtr> cloud_vec;
  std::vector<Pose> pose_vec;
  std::vector<Eigen::Matrix4f> trans_vec;
  std::vector<Frame> frame_set;
  static int id;  // frames processed so far (ssc.cpp:28)
  int name = 0;   // next track id (ssc.h:49)

  // ---- grid of the curved voxels: scvod_grid_dims (ssc.cpp:36-39) -------------------------------------------------------
  int range_num, sector_num, azimuth_num, bin_num;

  // ---- one scan: process() = one scvod_push_scans of a single scan; the members are filled from scvod_frame_* ---------
  void process(const CloudIPtr& cloudIn_);
  CloudIPtr extractGroudByPatchWork(const CloudIPtr& cloudIn_);  // scvod_ground
  void makeApriVec(const CloudIPtr& cloud_);                     // scvod_bin
  void makeHashCloud(const std::vector<PointAPRI>& apriIn_);     // scvod_frame_voxels
  void intensityCalibrationByCurvature(CloudIPtr& cloudIn_);     // disabled in the reference (ssc.cpp:234-235): no-op
  void intensityVisualization(const CloudIPtr& cloud_);          // visualisation only: no-op
  void recordIntensity(VoxelMap& hash_);                         // text dumps of the reference (ssc.cpp:1550-1587): no-op
  boost::shared_ptr<PatchWork<pcl::PointXYZI>> PatchworkGroundSeg;
  std::vector<CloudIPtr> g_cloud_vec;  // ground clouds, one per processed scan
  CloudIPtr cloud_use;                 // points that passed the gates of makeApriVec, apri order
  std::vector<PointAPRI> apri_vec;
  VoxelMap hash_cloud;
  Frame frame_ssc;

  // ---- clusters of the scan: computed by the push already, these publish cluster_set / Voxel::label --------------------
  void segment();
  void clusterAndCreateFrame(const std::vector<PointAPRI>& apri_vec_, VoxelMap& hash_cloud_);
  void refineClusterByIntensity(Frame& frame_ssc);
  void refineClusterByBoundingBox(Frame& frame_ssc_);
  void recognize(Frame& frame_ssc_);
  std::vector<int> findVoxelNeighbors(const int& range_idx_, const int& sector_idx_, const int& azimuth_idx_, int size_);
  void mergeClusters(std::vector<int>& clusterIdxs_, const int& idx1_, const int& idx2_);
  pcl::PointXYZI getCenterOfCloud(const CloudIPtr& cloud_);
  std::pair<pcl::PointXYZI, pcl::PointXYZI> getBoundingBoxOfCloud(const CloudIPtr& cloud_);
  void getVoxelCloudFromHashCloud(VoxelMap& hashCloud_);

  // ---- frame chain: scvod_track / scvod_initialization -------------------------------------------------------------------
  void tracking(Frame& frame_pre_, Frame& frame_next_, Pose pose_pre_, Pose pose_next_);
  Frame intialization(const std::vector<Frame>& frames_, const std::vector<Pose>& poses_);
  Frame frame_based;

  // ---- outputs of the reference that are files or coloured clouds (kept as members, written by saveSegCloud) -----------
  void saveSegCloud(Frame& frame_ssc, const CloudIPtr& cloud_, const std::string& path_, int mode);
  std::string calib_save, seg_save, pcd_save, map_save, evaluate_save;
  CloudRGBPtr cloud_original, cloud_dynamic, cloud_static, instance_map;
  std::vector<CloudIPtr> eva_ori;
  CloudIPtr cloud_eva_ori, cloud_eva_static, cloud_eva_dynamic;
  ofstream ofs;

  // ---- additions of this implementation (not in the reference) ----------------------------------------------------------
  // per-input-point outcome class (enum scvod_point_class) of every frame processed by segDF()
  std::vector<std::vector<uint8_t>> point_class;
  // scan-to-map GICP (docs/gicp_spec.md): aligns cloud_ to map_ starting from guess_, returns the refined pose
  Pose gicpScanToMap(const CloudIPtr& cloud_, const CloudIPtr& map_, Pose guess_);
  scvod_ctx* context();  // the CUDA context behind this object (created on first use)

 private:
  void fillFrameFromContext(int f, const CloudIPtr& cloudIn_);
  void refreshClusters(Frame& frame_);
  void refreshClustersFrom(Frame& frame_, int scvod_frame_index);  // SCVOD_INIT_FRAME: the initialised frame
  scvod_ctx* ctx_ = nullptr;
  int ctx_points_ = 0, ctx_batch_ = 0;
  CloudIPtr last_input_;
};

#endif
