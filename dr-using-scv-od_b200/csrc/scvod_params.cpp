// The two parameter sets of the reference's YAML files (config/semantickitti.yaml, config/parkinglot.yaml) with the Utility()
// defaults for the keys a file omits (reference include/utility.h:283-313).  Host-only C++: linked into libscvod_b200.so and,
// with the scan generator (synth.cpp), into libscvod_synth.so, which is all the CPU reference arm of bench.py loads.
#include "../../include/scvod.h"

static void params_common(scvod_params* p) {
  // Utility() defaults (reference include/utility.h:283-313) for keys a YAML file may omit
  p->sensor_height = 2.0f;
  p->min_dis = 0.0f;
  p->max_dis = 50.0f;
  p->min_angle = 0.0f;
  p->max_angle = 360.0f;
  p->min_azimuth = -30.0f;
  p->max_azimuth = 60.0f;
  p->range_res = 0.2f;
  p->sector_res = 1.2f;
  p->azimuth_res = 2.0f;
  p->refine_height = -1.0f;
  p->max_z = 1.0f;
  p->min_z = -1.0f;
  p->car_square = 2.0f;
  p->iteration = 3;
  p->toBeClass = 1;
  p->search_c = 2;
  p->intensity_diff = 50.f;
  p->intensity_cov = 20.f;
  p->occupancy = 0.6f;
  p->building = 0;
  p->tree = 1;
  p->car = 2;
}

extern "C" void scvod_params_semantickitti(scvod_params* p) {  // reference config/semantickitti.yaml:24-53
  params_common(p);
  p->sensor_height = 1.73f;
  p->refine_height = -0.2f;
  p->max_z = 0.8f;
  p->min_z = -1.2f;
  p->car_square = 30.0f;
  p->min_dis = 1.5f;
  p->max_dis = 30.0f;
  p->min_angle = 0.0f;
  p->max_angle = 360.0f;
  p->min_azimuth = -40.0f;
  p->max_azimuth = 80.0f;
  p->range_res = 0.4f;
  p->sector_res = 1.2f;
  p->azimuth_res = 2.0f;
  p->iteration = 3;
  p->toBeClass = 10;
  p->search_c = 2;
  p->intensity_diff = 2.0f;
  p->intensity_cov = 1.0f;
  p->occupancy = 0.4f;
}

extern "C" void scvod_params_parkinglot(scvod_params* p) {  // reference config/parkinglot.yaml:23-50 (+ utility.h defaults)
  params_common(p);
  p->sensor_height = 1.83f;
  p->min_dis = 0.8f;
  p->max_dis = 40.0f;
  p->min_angle = 0.0f;
  p->max_angle = 360.0f;
  p->min_azimuth = -30.0f;
  p->max_azimuth = 60.0f;
  p->range_res = 0.4f;
  p->sector_res = 1.2f;
  p->azimuth_res = 2.0f;
  p->iteration = 3;
  p->toBeClass = 6;
  p->search_c = 2;
  p->intensity_diff = 2.0f;
  p->intensity_cov = 1.0f;
  p->occupancy = 0.8f;
}

