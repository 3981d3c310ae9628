// ssc.cpp — bodies of the `SSC` methods over the C-ABI of libscvod_b200.so (include/scvod.h).
//
// Reference bodies these replace (paths relative to the reference repository):
//   process      src/ssc.cpp:224-251  -> scvod_push_scans (ground + binning + descriptor + clustering + car rule, one pass)
//   segment      src/ssc.cpp:637-656  -> cluster tables of the same pass (scvod_frame_point_cluster / scvod_frame_clusters)
//   recognize    src/ssc.cpp:834-895  -> cluster types of the same pass
//   tracking     src/ssc.cpp:1250-1426 -> scvod_track
//   segDF        src/ssc.cpp:1428-1452 -> batched scvod_push_scans over all scans, scvod_track, scvod_labels_range
// The public containers of the class (apri_vec, hash_cloud, frame_ssc, cloud_use, g_cloud_vec, frame_set) are
// filled from the library's inspection calls so that downstream code reading them keeps working; filling them
// is bookkeeping, not a compute path — nothing here bins, fits or clusters on the CPU.
#include "ssc.h"
#include "voxel_grid.h"

#include <cstring>
#include <stdexcept>

#include "scvod.h"

int SSC::id = 0;

namespace {

void check(int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + scvod_last_error());
}

std::vector<float> pack_xyzi(const pcl::PointCloud<pcl::PointXYZI>& c) {
  std::vector<float> out((size_t)4 * std::max<size_t>(1, c.points.size()));
  for (size_t i = 0; i < c.points.size(); ++i) {
    out[4 * i] = c.points[i].x;
    out[4 * i + 1] = c.points[i].y;
    out[4 * i + 2] = c.points[i].z;
    out[4 * i + 3] = c.points[i].intensity;
  }
  return out;
}

void pose6(const Pose& p, float out[6]) {
  out[0] = p.x;
  out[1] = p.y;
  out[2] = p.z;
  out[3] = p.roll;
  out[4] = p.pitch;
  out[5] = p.yaw;
}

// general 4x4 inverse (Gauss-Jordan, double); used once per pose line for tr^-1 * cam * tr (src/ssc.cpp:967)
Eigen::Matrix4f inverse4(const Eigen::Matrix4f& m) {
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      a[i][j] = m(i, j);
      a[i][4 + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r)
      if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[piv][j]);
    double d = a[c][c];
    if (d == 0.0) return Eigen::Matrix4f::Identity();
    for (int j = 0; j < 8; ++j) a[c][j] /= d;
    for (int r = 0; r < 4; ++r) {
      if (r == c) continue;
      double f = a[r][c];
      for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
    }
  }
  Eigen::Matrix4f out;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out(i, j) = (float)a[i][4 + j];
  return out;
}

bool synthetic_source(const std::string& path, int* count, int* rings, int* cols) {
  // data_path_ "synth:<scans>[:<rings>:<cols>]" selects the deterministic generator (scvod_synth_scan)
  if (path.rfind("synth:", 0) != 0) return false;
  *count = 8;
  *rings = 64;
  *cols = 1800;
  std::sscanf(path.c_str() + 6, "%d:%d:%d", count, rings, cols);
  return true;
}

}  // namespace

SSC::~SSC() {
  if (ctx_) scvod_destroy(ctx_);
}

SSC::SSC() {
  allocateMemory();
  scvod_params p;
  scvod_params_semantickitti(&p);
  p.min_dis = min_dis;
  p.max_dis = max_dis;
  p.min_angle = min_angle;
  p.max_angle = max_angle;
  p.min_azimuth = min_azimuth;
  p.max_azimuth = max_azimuth;
  p.range_res = range_res;
  p.sector_res = sector_res;
  p.azimuth_res = azimuth_res;
  scvod_grid g;
  check(scvod_grid_dims(&p, &g), "scvod_grid_dims");  // src/ssc.cpp:36-39
  range_num = g.range_num;
  sector_num = g.sector_num;
  azimuth_num = g.azimuth_num;
  bin_num = g.bin_num;
  calib_save = out_path + calib_path;
  seg_save = out_path + seg_path;
  pcd_save = out_path + pcd_path;
  map_save = out_path + map_path;
  evaluate_save = out_path + evaluate_path;
  if (save && out_path != " " && !out_path.empty())
    for (const std::string& d : {calib_save, seg_save, pcd_save, map_save, evaluate_save}) fsmkdir(d);
  std::cout << "----  SSC INITIALIZATION (scvod_b200)  ----\n"
            << "range_res: " << range_res << " sector_res: " << sector_res << " azimuth_res: " << azimuth_res << "\n"
            << "min_dis: " << min_dis << " max_dis: " << max_dis << " range_num: " << range_num << "\n"
            << "min_angle: " << min_angle << " max_angle: " << max_angle << " sector_num: " << sector_num << "\n"
            << "min_azimuth: " << min_azimuth << " max_azimuth: " << max_azimuth << " azimuth_num: " << azimuth_num << "\n"
            << "data_path: " << data_path << "\npose_path: " << pose_path << std::endl;
}

void SSC::allocateMemory() {
  PatchworkGroundSeg.reset(new PatchWork<pcl::PointXYZI>());
  cloud_use.reset(new pcl::PointCloud<pcl::PointXYZI>());
  cloud_original.reset(new pcl::PointCloud<pcl::PointXYZRGB>());
  cloud_dynamic.reset(new pcl::PointCloud<pcl::PointXYZRGB>());
  cloud_static.reset(new pcl::PointCloud<pcl::PointXYZRGB>());
  cloud_eva_static.reset(new pcl::PointCloud<pcl::PointXYZI>());
  cloud_eva_dynamic.reset(new pcl::PointCloud<pcl::PointXYZI>());
  cloud_eva_ori.reset(new pcl::PointCloud<pcl::PointXYZI>());
  instance_map.reset(new pcl::PointCloud<pcl::PointXYZRGB>());
}

void SSC::reset() {  // src/ssc.cpp:79-86
  Frame frame_new;
  frame_ssc = frame_new;
  apri_vec.clear();
  hash_cloud.clear();
  cloud_use.reset(new pcl::PointCloud<pcl::PointXYZI>());
}

scvod_ctx* SSC::context() {
  if (ctx_) return ctx_;
  scvod_params p;
  scvod_params_semantickitti(&p);
  p.sensor_height = sensor_height;
  p.min_dis = min_dis;
  p.max_dis = max_dis;
  p.min_angle = min_angle;
  p.max_angle = max_angle;
  p.min_azimuth = min_azimuth;
  p.max_azimuth = max_azimuth;
  p.range_res = range_res;
  p.sector_res = sector_res;
  p.azimuth_res = azimuth_res;
  p.refine_height = refine_height;
  p.max_z = max_z;
  p.min_z = min_z;
  p.car_square = car_square;
  p.iteration = iteration;
  p.toBeClass = toBeClass;
  p.search_c = search_c;
  p.intensity_diff = intensity_diff;
  p.intensity_cov = intensity_cov;
  p.occupancy = occupancy;
  p.building = building;
  p.tree = tree;
  p.car = car;
  if (ctx_points_ <= 0) ctx_points_ = 1 << 18;
  if (ctx_batch_ <= 0) ctx_batch_ = 16;
  const char* dev = std::getenv("UFO_DEVICE");
  check(scvod_create(&p, dev ? std::atoi(dev) : 0, ctx_points_, ctx_batch_, &ctx_), "scvod_create");
  return ctx_;
}

// ---------------------------------------------------------------------------------------------------------
// per-scan stages
// ---------------------------------------------------------------------------------------------------------
pcl::PointCloud<pcl::PointXYZI>::Ptr SSC::extractGroudByPatchWork(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloudIn_) {
  double time_pw;
  pcl::PointCloud<pcl::PointXYZI>::Ptr g_cloud(new pcl::PointCloud<pcl::PointXYZI>());
  pcl::PointCloud<pcl::PointXYZI>::Ptr ng_cloud(new pcl::PointCloud<pcl::PointXYZI>());
  g_cloud_vec.emplace_back(g_cloud);
  PatchworkGroundSeg->set_sensor(sensor_height);
  PatchworkGroundSeg->estimate_ground(*cloudIn_, *g_cloud, *ng_cloud, time_pw);
  return ng_cloud;
}

void SSC::intensityCalibrationByCurvature(pcl::PointCloud<pcl::PointXYZI>::Ptr&) {
  // the reference's call is commented out (src/ssc.cpp:234-235): intentionally a no-op
}
void SSC::intensityVisualization(const pcl::PointCloud<pcl::PointXYZI>::Ptr&) {}  // debug colouring only (src/ssc.cpp:197-222)
void SSC::recordIntensity(std::unordered_map<int, Voxel>&) {}                      // text dumps only (src/ssc.cpp:1550-1587)
void SSC::getVoxelCloudFromHashCloud(std::unordered_map<int, Voxel>& hashCloud_) {
  frame_ssc.vox_cloud->clear();
  for (auto& v : hashCloud_) frame_ssc.vox_cloud->push_back(v.second.center);
}
void SSC::saveSegCloud(Frame&, const pcl::PointCloud<pcl::PointXYZI>::Ptr&, const std::string&, int) {}  // PCD dumps (src/ssc.cpp:469-569)

// makeApriVec on an arbitrary cloud: polar coordinates, gates and curved-voxel indices from k_bin_only.
void SSC::makeApriVec(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_) {
  const int n = (int)cloud_->points.size();
  std::vector<float> xyzi = pack_xyzi(*cloud_);
  std::vector<uint8_t> pass(std::max(1, n));
  std::vector<int32_t> vid(std::max(1, n)), ri(std::max(1, n)), si(std::max(1, n)), ei(std::max(1, n));
  std::vector<float> rg(std::max(1, n)), an(std::max(1, n)), az(std::max(1, n));
  check(scvod_bin(context(), xyzi.data(), n, pass.data(), vid.data(), ri.data(), si.data(), ei.data(), rg.data(), an.data(), az.data()), "scvod_bin");
  for (int i = 0; i < n; ++i) {
    const pcl::PointXYZI& pt = cloud_->points[i];
    if (!pass[i]) {  // src/ssc.cpp:161-172
      cloud_eva_static->points.push_back(pt);
      continue;
    }
    cloud_use->points.push_back(pt);
    frame_ssc.cloud_use->points.push_back(pt);
    PointAPRI a;
    a.x = pt.x;
    a.y = pt.y;
    a.z = pt.z;
    a.intensity = pt.intensity;
    a.range = rg[i];
    a.angle = an[i];
    a.azimuth = az[i];
    a.range_idx = ri[i];
    a.sector_idx = si[i];
    a.azimuth_idx = ei[i];
    a.voxel_idx = vid[i];
    apri_vec.emplace_back(a);
  }
}

// makeHashCloud: container bookkeeping for callers that use the stage on its own.  SSC::process fills
// hash_cloud from the GPU descriptor instead (fillFrameFromContext).
void SSC::makeHashCloud(const std::vector<PointAPRI>& apriIn_) {
  for (int i = 0; i < (int)apriIn_.size(); ++i) {
    Voxel& v = hash_cloud[apriIn_[i].voxel_idx];
    if (v.ptIdx.empty()) {
      v.range_idx = apriIn_[i].range_idx;
      v.sector_idx = apriIn_[i].sector_idx;
      v.azimuth_idx = apriIn_[i].azimuth_idx;
    }
    v.ptIdx.emplace_back(i);
    v.intensity_record.emplace_back(apriIn_[i].intensity);
  }
}

void SSC::fillFrameFromContext(int f, const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloudIn_) {
  scvod_ctx* c = context();
  int32_t cnt[9];
  check(scvod_frame_counts(c, f, cnt), "scvod_frame_counts");
  const int nG = cnt[1], nN = cnt[2], M = cnt[3], V = cnt[4];
  std::vector<int32_t> gsrc(std::max(1, nG)), nsrc(std::max(1, nN)), asrc(std::max(1, M)), avid(std::max(1, M));
  check(scvod_frame_ground_order(c, f, gsrc.data(), nsrc.data()), "scvod_frame_ground_order");
  check(scvod_frame_apri(c, f, asrc.data(), avid.data()), "scvod_frame_apri");
  pcl::PointCloud<pcl::PointXYZI>::Ptr g_cloud(new pcl::PointCloud<pcl::PointXYZI>());
  for (int i = 0; i < nG; ++i) g_cloud->points.push_back(cloudIn_->points[gsrc[i]]);
  g_cloud_vec.emplace_back(g_cloud);
  std::vector<int32_t> vvid(std::max(1, V)), vcnt(std::max(1, V)), vtri(3 * std::max(1, V)), vlab(std::max(1, V));
  std::vector<float> vav(std::max(1, V)), vcov(std::max(1, V)), vctr(3 * std::max(1, V));
  check(scvod_frame_voxels(c, f, vvid.data(), vcnt.data(), vav.data(), vcov.data(), vctr.data(), vtri.data(), vlab.data()), "scvod_frame_voxels");
  hash_cloud.clear();
  hash_cloud.reserve(V);
  for (int v = 0; v < V; ++v) {
    Voxel& vx = hash_cloud[vvid[v]];
    vx.range_idx = vtri[3 * v];
    vx.sector_idx = vtri[3 * v + 1];
    vx.azimuth_idx = vtri[3 * v + 2];
    vx.intensity_av = vav[v];
    vx.intensity_cov = vcov[v];
    vx.center.x = vctr[3 * v];
    vx.center.y = vctr[3 * v + 1];
    vx.center.z = vctr[3 * v + 2];
    vx.center.intensity = (float)vvid[v];  // src/ssc.cpp:277
    vx.label = vlab[v];
    vx.ptIdx.reserve(vcnt[v]);
  }
  cloud_use->clear();
  frame_ssc.cloud_use->clear();
  apri_vec.clear();
  apri_vec.resize(M);
  const int rs = range_num * sector_num;
  for (int m = 0; m < M; ++m) {
    const pcl::PointXYZI& pt = cloudIn_->points[asrc[m]];
    cloud_use->points.push_back(pt);
    Voxel& vx = hash_cloud[avid[m]];
    vx.ptIdx.emplace_back(m);
    vx.intensity_record.emplace_back(pt.intensity);
    PointAPRI& a = apri_vec[m];
    a.x = pt.x;
    a.y = pt.y;
    a.z = pt.z;
    a.intensity = pt.intensity;
    a.voxel_idx = avid[m];
    a.azimuth_idx = avid[m] / rs;  // exact for in-range indices; the aliased (-1) cases keep the voxel's own triple below
    a.range_idx = (avid[m] % rs) / sector_num;
    a.sector_idx = avid[m] % sector_num;
    a.range = a.angle = a.azimuth = 0.f;  // polar floats are not kept per point on the device; makeApriVec() returns them
  }
  *frame_ssc.cloud_use = *cloud_use;
  frame_ssc.scvod_frame = f;
}

void SSC::process(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloudIn_) {
  scvod_ctx* c = context();
  const int f = scvod_num_frames(c);
  std::vector<float> xyzi = pack_xyzi(*cloudIn_);
  int64_t off[2] = {0, (int64_t)cloudIn_->points.size()};
  check(scvod_push_scans(c, xyzi.data(), off, 1), "scvod_push_scans");
  last_input_ = cloudIn_;
  fillFrameFromContext(f, cloudIn_);
}

void SSC::refreshClusters(Frame& fr) {
  if (fr.scvod_frame >= 0) refreshClustersFrom(fr, fr.scvod_frame);
}

// cluster_set / voxel labels of `fr` from frame f of the library (f may be SCVOD_INIT_FRAME)
void SSC::refreshClustersFrom(Frame& fr, int f) {
  scvod_ctx* c = context();
  int32_t cnt[9];
  check(scvod_frame_counts(c, f, cnt), "scvod_frame_counts");
  const int V = cnt[4], C = cnt[8];
  std::vector<int32_t> vvid(std::max(1, V)), vlab(std::max(1, V));
  check(scvod_frame_voxels(c, f, vvid.data(), nullptr, nullptr, nullptr, nullptr, nullptr, vlab.data()), "scvod_frame_voxels");
  std::vector<int32_t> name(std::max(1, C)), type(std::max(1, C)), state(std::max(1, C)), npts(std::max(1, C)), nvox(std::max(1, C));
  std::vector<float> bbox(6 * std::max(1, C));
  check(scvod_frame_clusters(c, f, C, name.data(), type.data(), state.data(), npts.data(), nvox.data(), bbox.data()), "scvod_frame_clusters");
  std::unordered_map<int, Cluster> fresh;
  int max_name = 0;
  for (int i = 0; i < C; ++i) {  // inserted in the library's cluster_set iteration order
    Cluster cl;
    auto old = fr.cluster_set.find(name[i]);
    if (old != fr.cluster_set.end()) cl.track_id = old->second.track_id;
    cl.name = name[i];
    cl.type = type[i];
    cl.state = state[i];
    cl.bounding_box.first.x = bbox[6 * i];
    cl.bounding_box.first.y = bbox[6 * i + 1];
    cl.bounding_box.first.z = bbox[6 * i + 2];
    cl.bounding_box.second.x = bbox[6 * i + 3];
    cl.bounding_box.second.y = bbox[6 * i + 4];
    cl.bounding_box.second.z = bbox[6 * i + 5];
    fresh.insert(std::make_pair(name[i], cl));
    max_name = std::max(max_name, name[i] + 1);
  }
  for (int v = 0; v < V; ++v) {
    auto hv = fr.hash_cloud.find(vvid[v]);
    if (hv != fr.hash_cloud.end()) hv->second.label = vlab[v];
    if (vlab[v] < 0) continue;
    auto cl = fresh.find(vlab[v]);
    if (cl == fresh.end()) continue;
    cl->second.occupy_voxels.emplace_back(vvid[v]);
    if (hv != fr.hash_cloud.end()) {
      for (int m : hv->second.ptIdx) {
        cl->second.occupy_pts.emplace_back(m);
        if (m < (int)fr.cloud_use->points.size()) cl->second.cloud->points.push_back(fr.cloud_use->points[m]);
      }
    }
  }
  fr.cluster_set.swap(fresh);
  fr.max_name = max_name;
}

void SSC::segment() {  // clustering + both refinements ran inside scvod_push_scans; publish their result
  frame_ssc.hash_cloud = hash_cloud;
  refreshClusters(frame_ssc);
  for (auto& v : frame_ssc.hash_cloud) hash_cloud[v.first].label = v.second.label;
  getVoxelCloudFromHashCloud(hash_cloud);
}
void SSC::clusterAndCreateFrame(const std::vector<PointAPRI>&, std::unordered_map<int, Voxel>&) { refreshClusters(frame_ssc); }
void SSC::refineClusterByIntensity(Frame& fr) { refreshClusters(fr); }
void SSC::refineClusterByBoundingBox(Frame& fr) { refreshClusters(fr); }
void SSC::recognize(Frame& fr) { refreshClusters(fr); }
void SSC::mergeClusters(std::vector<int>& clusterIdxs_, const int& idx1_, const int& idx2_) {
  for (int& c : clusterIdxs_)
    if (c == idx1_) c = idx2_;
}

// src/ssc.cpp:395-411 (pure index arithmetic; the device twin is k_vox_nbr / k_similar_edges)
std::vector<int> SSC::findVoxelNeighbors(const int& range_idx_, const int& sector_idx_, const int& azimuth_idx_, int size_) {
  std::vector<int> out;
  if (range_idx_ > range_num * 0.6) size_ = 1;
  for (int x = range_idx_ - size_; x <= range_idx_ + size_; ++x) {
    if (x > range_num - 1 || x < 0) continue;
    for (int y = sector_idx_ - size_; y <= sector_idx_ + size_; ++y) {
      if (y > sector_num - 1 || y < 0) continue;
      for (int z = azimuth_idx_ - size_; z <= azimuth_idx_ + size_; ++z) {
        if (z > azimuth_num - 1 || z < 0) continue;
        out.emplace_back(x * sector_num + y + z * range_num * sector_num);
      }
    }
  }
  return out;
}

pcl::PointXYZI SSC::getCenterOfCloud(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_) {
  pcl::PointXYZI c;
  for (auto& p : cloud_->points) {
    c.x += p.x;
    c.y += p.y;
    c.z += p.z;
  }
  const float n = (float)std::max<size_t>(1, cloud_->points.size());
  c.x /= n;
  c.y /= n;
  c.z /= n;
  return c;
}

std::pair<pcl::PointXYZI, pcl::PointXYZI> SSC::getBoundingBoxOfCloud(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_) {
  pcl::PointXYZI lo, hi;
  lo.x = lo.y = lo.z = FLT_MAX;
  hi.x = hi.y = hi.z = -FLT_MAX;
  for (auto& p : cloud_->points) {
    lo.x = std::min(lo.x, p.x);
    lo.y = std::min(lo.y, p.y);
    lo.z = std::min(lo.z, p.z);
    hi.x = std::max(hi.x, p.x);
    hi.y = std::max(hi.y, p.y);
    hi.z = std::max(hi.z, p.z);
  }
  return std::make_pair(lo, hi);
}

// ---------------------------------------------------------------------------------------------------------
// frame chain
// ---------------------------------------------------------------------------------------------------------
void SSC::tracking(Frame& frame_pre_, Frame& frame_next_, Pose pose_pre_, Pose pose_next_) {
  scvod_ctx* c = context();
  const int a = frame_pre_.scvod_frame, b = frame_next_.scvod_frame;
  if (a < 0 || b != a + 1) {
    ROS_WARN("tracking: frames %d and %d are not consecutive frames of this SSC object", a, b);
    return;
  }
  // scvod_track tracks every pair from the first untracked frame up to b: called out of order it would run the skipped pairs
  // with whatever poses the array holds, and a pair that is already tracked is not run again.  The reference's segDF only ever
  // calls tracking(i, i + 1) for i = 0, 1, 2, ... (ssc.cpp:1450-1452); anything else is refused instead of silently differing.
  int64_t tracked = 0;
  check(scvod_get_stat(c, "tracked_frames", &tracked), "scvod_get_stat");
  if ((int64_t)a != tracked) {
    ROS_WARN("tracking: frame %d is not the next frame to be tracked (%d): pairs must be tracked once, in order", a, (int)tracked);
    return;
  }
  std::vector<float> poses((size_t)6 * (b + 1), 0.f);
  pose6(pose_pre_, &poses[6 * a]);
  pose6(pose_next_, &poses[6 * b]);
  check(scvod_track(c, poses.data(), b + 1), "scvod_track");
  refreshClusters(frame_pre_);
  refreshClusters(frame_next_);
}

// SSC::intialization (src/ssc.cpp:1148-1248; dead code in the reference, never called by segDF): the frames must be the
// frames of this object in order (they carry their library index); the diff against the base frame and the cluster
// fusion run in scvod_initialization, the returned frame is a copy of the base frame with the fused clusters.
Frame SSC::intialization(const std::vector<Frame>& frames_, const std::vector<Pose>& poses_) {
  if (frames_.empty() || poses_.size() < frames_.size()) return Frame();
  scvod_ctx* c = context();
  for (size_t i = 0; i < frames_.size(); ++i) {
    if (frames_[i].scvod_frame != (int)i) {
      ROS_WARN("intialization: frame %d is not frame %d of this SSC object", frames_[i].scvod_frame, (int)i);
      return frames_.front();
    }
  }
  std::vector<float> poses((size_t)6 * frames_.size(), 0.f);
  for (size_t i = 0; i < frames_.size(); ++i) pose6(poses_[i], &poses[6 * i]);
  int32_t id_based = -1;
  check(scvod_initialization(c, poses.data(), (int)frames_.size(), &id_based), "scvod_initialization");
  Frame frame_based = frames_[id_based];
  refreshClustersFrom(frame_based, SCVOD_INIT_FRAME);
  mapping_init = true;  // ssc.cpp:1243
  ROS_INFO("initialization: based_id: %d, initialized cluster_num: %d", frame_based.id, (int)frame_based.cluster_set.size());
  return frame_based;
}

// ---------------------------------------------------------------------------------------------------------
// driver
// ---------------------------------------------------------------------------------------------------------
void SSC::getPose() {
  int n_synth, rings, cols;
  if (synthetic_source(data_path, &n_synth, &rings, &cols)) {
    std::vector<float> scratch((size_t)4 * rings * cols);
    for (int k = 0; k < n_synth; ++k) {
      int n = 0;
      float p6[6];
      check(scvod_synth_scan(0x5C0D0000ull, k, rings, cols, scratch.data(), &n, p6), "scvod_synth_scan");
      Pose p;
      std::memset(&p, 0, sizeof(p));
      p.x = p6[0];
      p.y = p6[1];
      p.z = p6[2];
      p.roll = p6[3];
      p.pitch = p6[4];
      p.yaw = p6[5];
      pose_vec.emplace_back(p);
    }
    return;
  }
  if (is_pcd) {
    ROS_ERROR("PCD pose clouds (is_pcd_) are not supported by this host layer: convert to KITTI poses.txt");
    ros::shutdown();
    return;
  }
  // KITTI poses.txt: T_velo = tr^-1 * cam * tr (src/ssc.cpp:941-991)
  std::ifstream pose_file(pose_path);
  std::string line;
  int count = 0;
  const Eigen::Matrix4f tr_inv = inverse4(tr);
  while (std::getline(pose_file, line)) {
    if (count < start || (count - start) % skip != 0) {
      count++;
      continue;
    }
    if (count >= end) break;
    std::istringstream is(line);
    Eigen::Matrix4f cam = Eigen::Matrix4f::Identity();
    for (int i = 0; i < 12; ++i) is >> cam.d[i];
    Eigen::Matrix4f velo_to_cam = tr_inv * cam * tr;
    trans_vec.emplace_back(velo_to_cam);
    Eigen::Matrix3f rot;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rot(i, j) = velo_to_cam(i, j);
    Eigen::Vector3f rpy = rotationMatrixToEulerAngles(rot);
    Pose p;
    std::memset(&p, 0, sizeof(p));
    p.x = velo_to_cam(0, 3);
    p.y = velo_to_cam(1, 3);
    p.z = velo_to_cam(2, 3);
    p.roll = rpy[0];
    p.pitch = rpy[1];
    p.yaw = rpy[2];
    count++;
    pose_vec.emplace_back(p);
  }
  ROS_DEBUG("load pose size: %d", (int)pose_vec.size());
}

void SSC::getCloud() {
  int n_synth, rings, cols;
  if (synthetic_source(data_path, &n_synth, &rings, &cols)) {
    std::vector<float> scratch((size_t)4 * rings * cols);
    for (int k = 0; k < n_synth; ++k) {
      int n = 0;
      check(scvod_synth_scan(0x5C0D0000ull, k, rings, cols, scratch.data(), &n, nullptr), "scvod_synth_scan");
      pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>());
      cloud->points.resize(n);
      for (int i = 0; i < n; ++i) {
        cloud->points[i].x = scratch[4 * i];
        cloud->points[i].y = scratch[4 * i + 1];
        cloud->points[i].z = scratch[4 * i + 2];
        cloud->points[i].intensity = scratch[4 * i + 3];
      }
      cloud_vec.emplace_back(cloud);
    }
    return;
  }
  if (is_pcd) {
    ROS_ERROR("PCD input (is_pcd_) is not supported by this host layer");
    ros::shutdown();
    return;
  }
  // KITTI velodyne .bin (+ .label): drop labels 0/1, intensity * max_intensity (src/ssc.cpp:1041-1071), then the
  // 0.08 m pcl::VoxelGrid downsample of src/ssc.cpp:1108-1111 (restated in include/voxel_grid.h; loader code, it runs
  // on the host as in the reference).
  std::vector<std::string> bins, labels;
  for (auto& e : fs::directory_iterator(data_path)) bins.push_back(e.path().string());
  std::sort(bins.begin(), bins.end());
  if (fs::is_directory(label_path)) {
    for (auto& e : fs::directory_iterator(label_path)) labels.push_back(e.path().string());
    std::sort(labels.begin(), labels.end());
  }
  if (start < 0 || end > (int)bins.size()) {
    ROS_WARN("the start or end index set error");
    ros::shutdown();
    return;
  }
  for (int i = start; i < end; i += skip) {
    std::ifstream in_cloud(bins[i], std::ios::binary);
    in_cloud.seekg(0, std::ios::end);
    const size_t npts = (size_t)in_cloud.tellg() / (4 * sizeof(float));
    in_cloud.seekg(0, std::ios::beg);
    std::vector<float> values(4 * npts);
    in_cloud.read((char*)values.data(), values.size() * sizeof(float));
    std::vector<uint32_t> lab;
    if (i < (int)labels.size()) {
      std::ifstream in_label(labels[i], std::ios::binary);
      lab.resize(npts);
      in_label.read((char*)lab.data(), npts * sizeof(uint32_t));
    }
    pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>());
    for (size_t k = 0; k < npts; ++k) {
      if (!lab.empty() && ((lab[k] & 0xFFFF) == 0 || (lab[k] & 0xFFFF) == 1)) continue;
      pcl::PointXYZI p;
      p.x = values[4 * k];
      p.y = values[4 * k + 1];
      p.z = values[4 * k + 2];
      p.intensity = values[4 * k + 3] * max_intensity;
      cloud->points.push_back(p);
    }
    ufo::voxelGridXYZI(*cloud, 0.08f, *cloud);
    cloud_vec.emplace_back(cloud);
  }
  ROS_DEBUG("load cloud size: %d", (int)cloud_vec.size());
}

// segDF (src/ssc.cpp:1428-1452): the per-scan loop becomes batched passes (scvod_push_scans over chunks of
// scans), the tracking loop one scvod_track call; frame_set / point_class are then filled from the library.
void SSC::segDF() {
  id = start;
  getPose();
  getCloud();
  const int nf = (int)std::min(cloud_vec.size(), pose_vec.size());
  if (nf == 0) {
    ROS_WARN("segDF: no scans loaded");
    return;
  }
  size_t max_pts = 1;
  for (auto& c : cloud_vec) max_pts = std::max(max_pts, c->points.size());
  if (!ctx_) {
    ctx_points_ = 1024;
    while ((size_t)ctx_points_ < max_pts) ctx_points_ *= 2;
    ctx_batch_ = std::min(nf, 32);
  }
  scvod_ctx* c = context();
  check(scvod_set_option(c, "inspect", 1), "scvod_set_option");
  TicToc t_all;
  const int f0 = scvod_num_frames(c);
  for (int s = 0; s < nf; s += ctx_batch_) {
    const int e = std::min(nf, s + ctx_batch_);
    std::vector<int64_t> off(e - s + 1, 0);
    for (int k = s; k < e; ++k) off[k - s + 1] = off[k - s] + (int64_t)cloud_vec[k]->points.size();
    std::vector<float> flat((size_t)4 * std::max<int64_t>(1, off.back()));
    for (int k = s; k < e; ++k) {
      const auto& pts = cloud_vec[k]->points;
      float* dst = flat.data() + 4 * off[k - s];
      for (size_t i = 0; i < pts.size(); ++i) {
        dst[4 * i] = pts[i].x;
        dst[4 * i + 1] = pts[i].y;
        dst[4 * i + 2] = pts[i].z;
        dst[4 * i + 3] = pts[i].intensity;
      }
    }
    check(scvod_push_scans(c, flat.data(), off.data(), e - s), "scvod_push_scans");
  }
  std::vector<float> poses((size_t)6 * (f0 + nf), 0.f);
  for (int k = 0; k < nf; ++k) pose6(pose_vec[k], &poses[6 * (f0 + k)]);
  check(scvod_track(c, poses.data(), f0 + nf), "scvod_track");
  const double ms = t_all.toc();
  ROS_INFO("segDF: %d frames in %.1f ms (%.1f scans/s), ground + binning + descriptor + clustering + tracking on the GPU", nf, ms,
           1000.0 * nf / ms);
  // publish results in the reference's containers
  point_class.assign(nf, std::vector<uint8_t>());
  for (int k = 0; k < nf; ++k) {
    point_class[k].resize(std::max<size_t>(1, cloud_vec[k]->points.size()));
    check(scvod_frame_labels(c, f0 + k, point_class[k].data(), (int)cloud_vec[k]->points.size()), "scvod_frame_labels");
    point_class[k].resize(cloud_vec[k]->points.size());
    fillFrameFromContext(f0 + k, cloud_vec[k]);
    frame_ssc.id = id;
    frame_ssc.hash_cloud = hash_cloud;
    refreshClusters(frame_ssc);
    frame_set.emplace_back(frame_ssc);
    reset();
    id += skip;
  }
  // dynamic / static clouds of the whole run in the map frame are one call away: scvod_static_submap_dev
  const char* dump = std::getenv("UFO_DUMP_LABELS");
  if (dump && *dump) {
    std::ofstream out(dump, std::ios::binary);
    int32_t n32 = nf;
    out.write((const char*)&n32, 4);
    for (int k = 0; k < nf; ++k) {
      int32_t n = (int32_t)point_class[k].size();
      out.write((const char*)&n, 4);
      out.write((const char*)point_class[k].data(), n);
    }
  }
}

Pose SSC::gicpScanToMap(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud_, const pcl::PointCloud<pcl::PointXYZI>::Ptr& map_, Pose guess_) {
  scvod_ctx* c = context();
  std::vector<float> src = pack_xyzi(*cloud_), tgt = pack_xyzi(*map_);
  check(scvod_gicp_set_target(c, tgt.data(), (int)map_->points.size(), nullptr), "scvod_gicp_set_target");
  float g6[6], T0[12];
  pose6(guess_, g6);
  scvod_pose_matrix(g6, T0);
  scvod_gicp_result res;
  check(scvod_gicp_align(c, src.data(), (int)cloud_->points.size(), T0, &res), "scvod_gicp_align");
  Pose out = guess_;
  out.x = res.pose6[0];
  out.y = res.pose6[1];
  out.z = res.pose6[2];
  out.roll = res.pose6[3];
  out.pitch = res.pose6[4];
  out.yaw = res.pose6[5];
  return out;
}


// test hook (tests/test_host_cpp.py): the loader's voxel-grid downsample on a flat xyzi array
extern "C" int ufo_voxel_grid(const float* xyzi, int n, float leaf, float* out_xyzi, int* n_out) {
  pcl::PointCloud<pcl::PointXYZI> in, out;
  in.points.resize(n);
  for (int i = 0; i < n; ++i) {
    in.points[i].x = xyzi[4 * i];
    in.points[i].y = xyzi[4 * i + 1];
    in.points[i].z = xyzi[4 * i + 2];
    in.points[i].intensity = xyzi[4 * i + 3];
  }
  const bool ok = ufo::voxelGridXYZI(in, leaf, out);
  *n_out = (int)out.points.size();
  for (size_t i = 0; i < out.points.size(); ++i) {
    out_xyzi[4 * i] = out.points[i].x;
    out_xyzi[4 * i + 1] = out.points[i].y;
    out_xyzi[4 * i + 2] = out.points[i].z;
    out_xyzi[4 * i + 3] = out.points[i].intensity;
  }
  return ok ? 0 : 1;
}
