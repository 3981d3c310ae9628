// host_cluster.h — host-side mirror of the reference's cluster-level bookkeeping.
//
// The GPU produces everything that is per-point or per-voxel (ground labels, voxel indices, the
// occupancy descriptor, voxel adjacency, connected components, similarity edges, re-binned tracking
// hits).  What remains is the reference's *sequential* cluster logic, whose results depend on
// libstdc++ container iteration order (std::unordered_map<int, Cluster>, reference include/utility.h:180;
// SURVEY.md hard part 6).  It is kept on the host with the same containers and the same insertion
// sequence as the reference so that cluster names, fusion order and tracking decisions come out
// identical; the data it touches is O(voxels + clusters) per scan, not O(points).
#pragma once
#include <algorithm>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "../../include/scvod.h"

namespace scvod {

struct P4 {
  float x, y, z, w;
};

// Cluster (reference include/utility.h:142-162) at voxel granularity. Voxels are compact ids, which are
// monotone in voxel_idx, so every sort / lexicographic comparison gives the reference's order.
struct HCluster {
  int track_id = -1, name = -1, type = -1, state = -1;
  float bb_min[3] = {0, 0, 0}, bb_max[3] = {0, 0, 0};
  std::vector<int> occupy_voxels;
  std::vector<int> part_end;    // ends (in occupy_voxels) of the initial CVC components concatenated by fusion
  int npts = 0;                 // occupy_pts.size(); the point lists themselves stay on the device (voxel CSR)
  // transformed clouds appended by tracking (*cloud += *cluster, ssc.cpp:1382): (offset, length) ranges of the
  // device-resident output of the tracking pass that produced them
  std::vector<std::pair<int, int>> carried;
  int n_carried = 0;
  // own voxels as runs of the frame's device-resident car CSR (built when the frame is pushed, concatenated by fusion
  // like part_end): lets the tracking kernel take a car cluster as a handful of kernel arguments instead of one
  // uploaded segment per voxel.  Empty for clusters that are not cars.
  struct OwnRun {
    int csr_start, csr_len, part_base, npts;
  };
  std::vector<OwnRun> own_runs;
  // Points of "tainted" voxels (voxels that hold a point with a -1 index, ssc.cpp:185-188): such a voxel appears in
  // occupy_voxels of every cluster that owns one of its points, but only the listed subgroups (FrameClusters::subgroups)
  // are in occupy_pts.  `part` is the part (see part_end) the subgroup's points belong to.  Empty for ordinary scans.
  struct TUnit {
    int sg, part;
  };
  std::vector<TUnit> tunits;
};

// The points of one tainted voxel that clusterAndCreateFrame put into one cluster (apri indices, ascending)
struct SubGroup {
  int vox = -1;
  std::vector<int> pts;
  float bb_min[3] = {0, 0, 0}, bb_max[3] = {0, 0, 0};
  int stage_name[3] = {-1, -1, -1};  // cluster that holds it after CVC / intensity refine / bbox refine (inspection)
};

// per-scan inputs from the GPU (host copies)
struct ScanTables {
  int M = 0, V = 0, n_events = 0, n_edges = 0;
  const int32_t* vox_cnt = nullptr;   // [V]
  const int32_t* vox_root = nullptr;  // [V] CCL root (min compact id of the component)
  const int32_t* vox_nbr = nullptr;   // [V][27]
  const float* vox_bbox = nullptr;    // [V][6]
  const int32_t* ev_cid = nullptr;    // [n_events]
  const int32_t* edges = nullptr;     // [n_edges][2] (root_from, root_to)
  // cluster names replayed on the device (k_name_replay); when null the host replays them from ev_cid / vox_nbr
  const int32_t* vox_name = nullptr;    // [V]
  const int32_t* name_first = nullptr;  // [max_name + 1] first event of every name (0x7fffffff: name vanished)
  int max_name = 0;
  // Tainted voxels (a point with range / sector / azimuth index -1 hashes into a voxel that is not its own cell, ssc.cpp:185-188):
  // their points are named one by one.  tv_cid ascending; points of voxel i are tp_*[tv_base[i] .. tv_base[i+1]) in ascending
  // apri index.  An event with ev_cid >= V is point V + t of these tables.
  int n_tvox = 0, n_tpts = 0;
  const int32_t* tv_cid = nullptr;   // [n_tvox]
  const int32_t* tv_base = nullptr;  // [n_tvox + 1]
  const int32_t* tp_m = nullptr;     // [n_tpts] apri index
  const float* tp_xyz = nullptr;     // [n_tpts][4]
  const int32_t* tp_name = nullptr;  // [n_tpts] names replayed on the device (with vox_name)
  const int32_t* tp_nbr = nullptr;   // [n_tpts][27] the point's own findVoxelNeighbors list (host replay only)
};

struct FrameClusters {
  int max_name = 0;
  std::vector<int> vox_label;  // hash_cloud[v].label
  std::unordered_map<int, HCluster> cluster_set;
  int n_clusters[3] = {0, 0, 0};
  std::vector<int> vox_name_stage[3];  // per-voxel cluster name after CVC / intensity refine / bbox refine (inspection)
  // tainted voxels of the frame (ascending compact id), the subgroups of each, and all subgroups
  std::vector<int> tvox;
  std::vector<std::vector<int>> sgs_of_tvox;
  std::vector<SubGroup> subgroups;
  bool tainted(int v) const { return !tvox.empty() && std::binary_search(tvox.begin(), tvox.end(), v); }
  const std::vector<int>* subgroups_of(int v) const {
    auto it = std::lower_bound(tvox.begin(), tvox.end(), v);
    return (it != tvox.end() && *it == v) ? &sgs_of_tvox[it - tvox.begin()] : nullptr;
  }
};

// SSC::clusterAndCreateFrame + refineClusterByIntensity + refineClusterByBoundingBox + recognize
// (reference src/ssc.cpp:299-393, 571-635, 437-467, 834-895) on the GPU-produced tables.
// Returns false when the replayed partition disagrees with the GPU components (internal error).
bool segment_and_recognize(const scvod_params& p, const ScanTables& t, FrameClusters& out, bool keep_stages);

// recognize (ssc.cpp:834-895) alone: cluster types from the bounding boxes stored in the clusters
void recognize_clusters(const scvod_params& p, FrameClusters& out);

// trans_next.inverse() * trans_pre (ssc.cpp:1255-1257) — 12 floats, row-major 3x4
void relative_pose(const float pose_next[6], const float pose_pre[6], float T[12]);
// pcl::getTransformation(x,y,z,roll,pitch,yaw) as 12 floats
void pose_matrix(const float pose[6], float T[12]);

}  // namespace scvod
