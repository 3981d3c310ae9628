// session.h — `class Session` with the members the reference declares (reference include/session.h:15-21).
//
// The reference defines none of them: src/session.cpp is an empty file, the header includes a common.h that does not
// exist, and no target compiles the class — there is no behaviour to reproduce.  BASELINE.json's "scan-vs-map diff in
// session.cpp" is in fact SSC::tracking / SSC::intialization (src/ssc.cpp:1250-1426, 1148-1248).
// The three loaders below are thin host glue (host/src/session.cpp) kept so that code including session.h links.
#ifndef _SESSION_H_
#define _SESSION_H_

#include "utility.h"

class Session : public Utility {
 public:
  using PosePtr = pcl::PointCloud<Pose>::Ptr;
  using CloudVec = std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr>;

  Session();
  ~Session();
  void allocateMemory();

  // key-frame poses of a stored session (one Pose per line: x y z roll pitch yaw)
  void getPose(PosePtr& pose_, const std::string& pose_path_);
  // the session's segmented clouds, one per pose, moved to the map frame and written to out_path_
  void getCloudSeg(CloudVec& session_seg_, PosePtr& pose_, const std::string& in_path_, const std::string& out_path_);
  // relocalisation descriptors + the clouds they were built from
  void getReloInfo(std::vector<cv::Mat>& relo_vec_, const std::string& relo_path_, CloudVec& build_vec_, const std::string& build_path_);
};

#endif
