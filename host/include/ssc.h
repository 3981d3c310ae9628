// ssc.h — `class SSC` with the public surface of the reference (include/ssc.h:7-105): same base class,
// same public data members, same method names and signatures, so the reference's src/main.cpp
// (`SSC ssc; ssc.segDF();`, main.cpp:9-10) and code written against those members build unchanged.
// The method bodies live in host/src/ssc.cpp and forward to the C-ABI of libscvod_b200.so
// (include/scvod.h); see INTEGRATION.md for the method -> entry point table.
#ifndef SSC_H_
#define SSC_H_

#include "patchwork.h"
#include "utility.h"

struct scvod_ctx;

class SSC : public Utility {
 public:
  static int id;

  int range_num;
  int sector_num;
  int azimuth_num;
  int bin_num;

  std::string calib_save;
  std::string seg_save;
  std::string pcd_save;
  std::string map_save;
  std::string evaluate_save;

  std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr> cloud_vec;
  std::vector<Pose> pose_vec;
  std::vector<Eigen::Matrix4f> trans_vec;

  std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr> g_cloud_vec;

  std::vector<PointAPRI> apri_vec;
  std::unordered_map<int, Voxel> hash_cloud;
  Frame frame_ssc;

  boost::shared_ptr<PatchWork<pcl::PointXYZI>> PatchworkGroundSeg;
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_use;

  pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud_original;
  pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud_dynamic;
  pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud_static;

  pcl::PointCloud<pcl::PointXYZRGB>::Ptr instance_map;

  std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr> eva_ori;
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_eva_ori;
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_eva_static;
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_eva_dynamic;

  Frame frame_based;
  int name = 0;
  std::vector<Frame> frame_set;

  ofstream ofs;

  ~SSC();
  SSC();

  void allocateMemory();
  void reset();

  // per-scan stages
  void process(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloudIn_);
  pcl::PointCloud<pcl::PointXYZI>::Ptr extractGroudByPatchWork(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloudIn_);
  void intensityCalibrationByCurvature(pcl::PointCloud<pcl::PointXYZI>::Ptr& cloudIn_);
  void makeApriVec(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_);
  void intensityVisualization(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_);
  void makeHashCloud(const std::vector<PointAPRI>& apriIn_);

  void segment();
  void clusterAndCreateFrame(const std::vector<PointAPRI>& apri_vec_, std::unordered_map<int, Voxel>& hash_cloud_);
  std::vector<int> findVoxelNeighbors(const int& range_idx_, const int& sector_idx_, const int& azimuth_idx_, int size_);
  void mergeClusters(std::vector<int>& clusterIdxs_, const int& idx1_, const int& idx2_);
  pcl::PointXYZI getCenterOfCloud(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_);
  std::pair<pcl::PointXYZI, pcl::PointXYZI> getBoundingBoxOfCloud(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_);
  void refineClusterByBoundingBox(Frame& frame_ssc_);
  void refineClusterByIntensity(Frame& frame_ssc);
  void getVoxelCloudFromHashCloud(std::unordered_map<int, Voxel>& hashCloud_);
  void saveSegCloud(Frame& frame_ssc, const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_, const std::string& path_, int mode);

  void recognize(Frame& frame_ssc_);

  // frame chain
  Frame intialization(const std::vector<Frame>& frames_, const std::vector<Pose>& poses_);
  void tracking(Frame& frame_pre_, Frame& frame_next_, Pose pose_pre_, Pose pose_next_);

  // driver
  void getPose();
  void getCloud();
  void segDF();

  void recordIntensity(std::unordered_map<int, Voxel>& hash_);

  // ---- additions of this implementation (not in the reference) ------------------------------------------
  // per-input-point outcome class (enum scvod_point_class) of every frame processed by segDF()
  std::vector<std::vector<uint8_t>> point_class;
  // scan-to-map GICP (docs/gicp_spec.md): aligns cloud_ to map_ starting from guess_, returns the refined pose
  Pose gicpScanToMap(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_, const pcl::PointCloud<pcl::PointXYZI>::Ptr& map_, Pose guess_);
  scvod_ctx* context();  // the CUDA context behind this object (created on first use)

 private:
  void fillFrameFromContext(int f, const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloudIn_);
  void refreshClusters(Frame& frame_);
  void refreshClustersFrom(Frame& frame_, int scvod_frame_index);  // SCVOD_INIT_FRAME: the initialised frame
  scvod_ctx* ctx_ = nullptr;
  int ctx_points_ = 0, ctx_batch_ = 0;
  pcl::PointCloud<pcl::PointXYZI>::Ptr last_input_;
};

#endif
