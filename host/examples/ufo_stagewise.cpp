// ufo_stagewise.cpp — drives SSC exactly as the body of the reference's SSC::segDF does (src/ssc.cpp:1435-1452):
// per scan process() -> segment() -> recognize() -> frame_set.emplace_back() -> reset(), then tracking() over
// consecutive frames — one method call at a time instead of the batched segDF() fast path.  Dumps, per frame,
// the points of every cluster whose state is dynamic (what saveSegCloud colours red, src/ssc.cpp:479-499) so the
// test-suite can compare them with the oracle.  usage: ufo_stagewise _params:=file.yaml <out.bin>
#include <cstring>

#include "ssc.h"

int main(int argc, char** argv) {
  ros::init(argc, argv, "ufo_stagewise");
  const char* out_path = argc > 2 ? argv[argc - 1] : "stagewise.bin";
  try {
    SSC ssc;
    ssc.getPose();
    ssc.getCloud();
    for (auto& cloud : ssc.cloud_vec) {
      ssc.process(cloud);
      ssc.segment();
      ssc.recognize(ssc.frame_ssc);
      ssc.frame_set.emplace_back(ssc.frame_ssc);
      ssc.reset();
    }
    for (int i = 0; i + 1 < (int)ssc.frame_set.size(); i++)
      ssc.tracking(ssc.frame_set[i], ssc.frame_set[i + 1], ssc.pose_vec[i], ssc.pose_vec[i + 1]);
    std::ofstream out(out_path, std::ios::binary);
    int32_t nf = (int32_t)ssc.frame_set.size();
    out.write((const char*)&nf, 4);
    for (auto& fr : ssc.frame_set) {
      std::vector<float> dyn;
      int32_t ncl = (int32_t)fr.cluster_set.size(), ncar = 0;
      for (auto& cs : fr.cluster_set) {
        if (cs.second.type == ssc.car) ++ncar;
        if (cs.second.state != 1) continue;
        for (int m : cs.second.occupy_pts) {
          const pcl::PointXYZI& p = fr.cloud_use->points[m];
          dyn.push_back(p.x);
          dyn.push_back(p.y);
          dyn.push_back(p.z);
        }
      }
      int32_t nd = (int32_t)(dyn.size() / 3), nuse = (int32_t)fr.cloud_use->points.size(), nvox = (int32_t)fr.hash_cloud.size();
      out.write((const char*)&nd, 4);
      out.write((const char*)&ncl, 4);
      out.write((const char*)&ncar, 4);
      out.write((const char*)&nuse, 4);
      out.write((const char*)&nvox, 4);
      out.write((const char*)dyn.data(), dyn.size() * sizeof(float));
    }
  } catch (const std::exception& e) {
    ROS_ERROR("%s", e.what());
    return 2;
  }
  return 0;
}
