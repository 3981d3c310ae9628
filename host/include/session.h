// session.h — `class Session` with the members the reference declares (include/session.h:15-21).  The
// reference defines none of them (src/session.cpp is an empty file and the class is compiled by nothing),
// so there is no behaviour to reproduce; the loaders below are thin host glue over the same C-ABI, kept so
// that code including session.h links.  BASELINE.json's "scan-vs-map diff in session.cpp" is in fact
// SSC::tracking / SSC::intialization (src/ssc.cpp:1250-1426, 1148-1248).
#ifndef _SESSION_H_
#define _SESSION_H_

#include "utility.h"

class Session : public Utility {
 public:
  Session();
  ~Session();

  void allocateMemory();
  void getPose(pcl::PointCloud<Pose>::Ptr& pose_, const std::string& pose_path_);
  void getCloudSeg(std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr>& session_seg_, pcl::PointCloud<Pose>::Ptr& pose_,
                   const std::string& in_path_, const std::string& out_path_);
  void getReloInfo(std::vector<cv::Mat>& relo_vec_, const std::string& relo_path_,
                   std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr>& build_vec_, const std::string& build_path_);
};

#endif
