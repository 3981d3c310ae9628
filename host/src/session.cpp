// session.cpp — definitions for the members `class Session` declares (reference include/session.h:15-21).
// The reference never defines them (its src/session.cpp is empty), so these are minimal loaders that make
// the class linkable; the scan-vs-map work BASELINE.json attributes to "session.cpp" is SSC::tracking.
#include "session.h"

Session::Session() { allocateMemory(); }
Session::~Session() {}
void Session::allocateMemory() {}

// pose file: one "x y z roll pitch yaw" line per frame
void Session::getPose(pcl::PointCloud<Pose>::Ptr& pose_, const std::string& pose_path_) {
  if (!pose_) pose_.reset(new pcl::PointCloud<Pose>());
  std::ifstream in(pose_path_);
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream is(line);
    Pose p;
    p.x = p.y = p.z = p.roll = p.pitch = p.yaw = 0.f;
    p.intensity = (float)pose_->points.size();
    p.time = 0.0;
    if (is >> p.x >> p.y >> p.z >> p.roll >> p.pitch >> p.yaw) pose_->points.push_back(p);
  }
}

// segmented session clouds: KITTI-style float[4] .bin files in in_path_, one per pose
void Session::getCloudSeg(std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr>& session_seg_, pcl::PointCloud<Pose>::Ptr& pose_,
                          const std::string& in_path_, const std::string& out_path_) {
  (void)out_path_;
  std::vector<std::string> files;
  if (fs::is_directory(in_path_))
    for (auto& e : fs::directory_iterator(in_path_)) files.push_back(e.path().string());
  std::sort(files.begin(), files.end());
  const size_t n = pose_ ? std::min(files.size(), pose_->points.size()) : files.size();
  for (size_t i = 0; i < n; ++i) {
    std::ifstream in(files[i], std::ios::binary);
    in.seekg(0, std::ios::end);
    const size_t npts = (size_t)in.tellg() / (4 * sizeof(float));
    in.seekg(0, std::ios::beg);
    std::vector<float> v(4 * npts);
    in.read((char*)v.data(), v.size() * sizeof(float));
    pcl::PointCloud<pcl::PointXYZI>::Ptr c(new pcl::PointCloud<pcl::PointXYZI>());
    c->points.resize(npts);
    for (size_t k = 0; k < npts; ++k) {
      c->points[k].x = v[4 * k];
      c->points[k].y = v[4 * k + 1];
      c->points[k].z = v[4 * k + 2];
      c->points[k].intensity = v[4 * k + 3];
    }
    session_seg_.push_back(c);
  }
}

void Session::getReloInfo(std::vector<cv::Mat>& relo_vec_, const std::string& relo_path_,
                          std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr>& build_vec_, const std::string& build_path_) {
  (void)relo_vec_;
  (void)relo_path_;
  (void)build_vec_;
  (void)build_path_;
  ROS_WARN("Session::getReloInfo has no definition in the reference; nothing to load");
}
