// ufo_main.cpp — the node entry point, equivalent to the reference's src/main.cpp:3-15 (init ROS, build one
// SSC, run segDF, spin).  `make ref_main` compiles the reference's own main.cpp against host/include
// instead, unchanged, when /root/reference is present.
#include "ssc.h"

int main(int argc, char** argv) {
  ros::init(argc, argv, "ufo");
  ros::console::set_logger_level(ROSCONSOLE_DEFAULT_NAME, ros::console::levels::Debug);
  ROS_INFO("----> ufo (scvod_b200) started");
  try {
    SSC ssc;
    ssc.segDF();
    size_t dyn = 0, tot = 0;
    for (auto& f : ssc.point_class)
      for (uint8_t c : f) {
        dyn += (c == SCVOD_PT_DYNAMIC);
        ++tot;
      }
    ROS_INFO("frames %d, points %zu, dynamic %zu", (int)ssc.frame_set.size(), tot, dyn);
  } catch (const std::exception& e) {
    ROS_ERROR("%s", e.what());
    return 2;
  }
  ros::spin();
  return 0;
}
