// host_cluster.cpp — see host_cluster.h.
#include "host_cluster.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#ifdef SEG_PROF
#include <chrono>
#include <cstdio>
static double g_tp[8];
static inline double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define TP(i) do { double t__ = now_us(); g_tp[i] += t__ - tp_last; tp_last = t__; } while (0)
extern "C" void seg_prof_dump() { for (int i = 0; i < 8; ++i) printf("phase %d: %.1f us\n", i, g_tp[i]); }
#else
#define TP(i)
#endif

namespace scvod {

namespace {

// ---------------------------------------------------------------------------------------------
// Cluster names of SSC::clusterAndCreateFrame (reference src/ssc.cpp:299-354).
//
// The reference walks the apri points in order; point i collects the points of the <=27 voxels around
// it (findVoxelNeighbors order) and runs the oc/nc propagation of :323-351, where merging renames the
// *current* point's cluster to the neighbour's (mergeClusters, :413-419).  Names therefore depend on
// the visiting sequence.  Facts used to replay it at voxel granularity (all points of a voxel share
// one neighbour list when no index is -1):
//   * the labelled points of a voxel always belong to one set, and a voxel is in one of three
//     states: no point labelled / only its first point labelled / all points labelled;
//   * a visitor that is still unlabelled skips unlabelled voxels (they stay unlabelled), adopts the
//     set of the first labelled voxel it meets, and from then on labels or merges everything it meets;
//   * after a point that was labelled when its turn came, its voxel and all 27 neighbours are in one
//     set for good ("stable"), so later points of that voxel are no-ops.  Only the first three points
//     of a voxel can find it unstable -> the GPU sends exactly those ("events"), in point order.
// ---------------------------------------------------------------------------------------------
struct NameReplay {
  std::vector<int> parent, setname;
  std::vector<uint8_t> state, stable;
  int find(int v) {
    int r = v;
    while (parent[r] != r) r = parent[r];
    while (parent[v] != r) {
      int nx = parent[v];
      parent[v] = r;
      v = nx;
    }
    return r;
  }
};

int replay_cluster_names(const ScanTables& t, std::vector<int>& vox_name) {
  const int V = t.V;
  NameReplay u;
  u.parent.resize(V);
  u.setname.assign(V, -1);
  u.state.assign(V, 0);
  u.stable.assign(V, 0);
  for (int v = 0; v < V; ++v) u.parent[v] = v;
  int cluster_name = 4;  // ssc.cpp:300
  for (int e = 0; e < t.n_events; ++e) {
    const int W = t.ev_cid[e];
    if (u.stable[W]) continue;
    const int32_t* nb = t.vox_nbr + 27 * (size_t)W;
    const bool labelled = (u.state[W] == 2);
    int oc = labelled ? u.find(W) : -1;
    bool skipped = false;
    for (int k = 0; k < 27; ++k) {
      const int Vn = nb[k];
      if (Vn < 0) continue;
      if (u.state[Vn] == 0) {
        if (oc >= 0) {
          u.parent[Vn] = oc;  // clusterIdxs[neighbor] = oc (:338)
          u.state[Vn] = 2;
        } else {
          skipped = true;
        }
      } else {
        int r = u.find(Vn);
        if (oc < 0) {
          oc = r;  // clusterIdxs[i] = nc (:334)
        } else if (r != oc) {
          u.parent[oc] = r;  // mergeClusters(oc -> nc): the neighbour's name survives (:329)
          oc = r;
        }
        u.state[Vn] = 2;
      }
    }
    if (oc < 0) {  // a new class (:345-351)
      ++cluster_name;
      u.parent[W] = W;
      u.setname[W] = cluster_name;
      u.state[W] = 2;
      for (int k = 0; k < 27; ++k) {
        const int Vn = nb[k];
        if (Vn < 0 || Vn == W) continue;
        u.parent[Vn] = W;
        u.state[Vn] = 2;
      }
      u.stable[W] = 1;
    } else {
      if (u.state[W] == 0) {  // only this (first) point of W got the label
        u.state[W] = 1;
        u.parent[W] = oc;
      }
      u.stable[W] = skipped ? 0 : 1;
    }
  }
  vox_name.resize(V);
  for (int v = 0; v < V; ++v) vox_name[v] = u.setname[u.find(v)];
  return cluster_name;
}

inline void sample_vec(std::vector<int>& v) {  // Utility::sampleVec, utility.h:452-456
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
}
inline bool name_in(const std::vector<int>& v, int n) { return std::find(v.begin(), v.end(), n) != v.end(); }

}  // namespace

// recognize (ssc.cpp:834-895; features :723-751): car <=> footprint < car_square, min z < min_z, max z < max_z
void recognize_clusters(const scvod_params& p, FrameClusters& out) {
  for (auto& c : out.cluster_set) {
    HCluster& cl = c.second;
    double diff_x = cl.bb_max[0] - cl.bb_min[0];
    double diff_y = cl.bb_max[1] - cl.bb_min[1];
    double square = diff_x * diff_y;
    double f6 = cl.bb_max[2], f9 = cl.bb_min[2];
    if (square > p.car_square) {
      cl.type = p.tree;  // building/tree split (regionGrowing, :797-832) does not feed labels; see DESIGN.md
    } else if (f9 < p.min_z && square < p.car_square && f6 < p.max_z) {
      cl.type = p.car;
    } else {
      cl.type = p.tree;
    }
  }
}

bool segment_and_recognize(const scvod_params& p, const ScanTables& t, FrameClusters& out, bool keep_stages) {
#ifdef SEG_PROF
  double tp_last = now_us();
#endif
  const int V = t.V;
  // ---- clusterAndCreateFrame (ssc.cpp:299-393) ---------------------------------------------------
  std::vector<int> vox_name;
  int last_name;
  // cluster_pt is filled in point order, so its keys are inserted in order of each cluster's first point (:360-375)
  std::unordered_map<int, int> cluster_pt;
  if (t.vox_name) {  // names replayed on the device
    vox_name.assign(t.vox_name, t.vox_name + V);
    last_name = t.max_name;
    std::vector<std::pair<int, int>> order;  // (first event, name)
    for (int nm = 0; nm <= last_name; ++nm)
      if (t.name_first[nm] != 0x7fffffff) order.push_back(std::make_pair(t.name_first[nm], nm));
    std::sort(order.begin(), order.end());
    for (auto& o : order) cluster_pt.insert(std::make_pair(o.second, 0));
  } else {
    last_name = replay_cluster_names(t, vox_name);
    std::vector<char> seen(last_name + 2, 0);
    for (int e = 0; e < t.n_events; ++e) {
      int nm = vox_name[t.ev_cid[e]];
      if (!seen[nm]) {
        seen[nm] = 1;
        cluster_pt.insert(std::make_pair(nm, 0));
      }
    }
  }
  TP(0);
  out.max_name = last_name;  // frame_ssc.max_name = cluster_name++ (:354)
  out.vox_label = vox_name;
  out.cluster_set.clear();
  for (int v = 0; v < V; ++v)
    if (vox_name[v] < 0 || vox_name[v] > last_name) return false;

  // voxels of every cluster in ascending compact id (== sorted voxel_idx): counting sort by name
  std::vector<int> name_start(last_name + 3, 0), vox_sorted(V);
  for (int v = 0; v < V; ++v) name_start[vox_name[v] + 1]++;
  for (int nm = 0; nm <= last_name + 1; ++nm) name_start[nm + 1] += name_start[nm];
  {
    std::vector<int> cur(name_start.begin(), name_start.end() - 1);
    for (int v = 0; v < V; ++v) vox_sorted[cur[vox_name[v]]++] = v;
  }
  std::unordered_map<int, std::vector<int>> roots_of;  // cluster name -> CVC component roots it contains
  for (auto& c : cluster_pt) {  // (:377-385) same iteration order as the reference's cluster_pt
    HCluster cl;
    cl.name = c.first;
    cl.occupy_voxels.assign(vox_sorted.begin() + name_start[c.first], vox_sorted.begin() + name_start[c.first + 1]);
    cl.part_end.push_back((int)cl.occupy_voxels.size());
    int np = 0;
    for (int v : cl.occupy_voxels) np += t.vox_cnt[v];
    cl.npts = np;
    if (!cl.occupy_voxels.empty()) roots_of[c.first].push_back(t.vox_root[cl.occupy_voxels[0]]);
    out.cluster_set.insert(std::make_pair(cl.name, std::move(cl)));
  }
  out.n_clusters[0] = (int)out.cluster_set.size();
  if (keep_stages) out.vox_name_stage[0] = vox_name;
  // the replayed name partition must coincide with the GPU's connected components
  {
    std::vector<int> root_name(V, -1);
    int nroots = 0;
    for (int v = 0; v < V; ++v) {
      int r = t.vox_root[v];
      if (r < 0 || r >= V) return false;
      if (root_name[r] == -1) {
        root_name[r] = vox_name[v];
        ++nroots;
      } else if (root_name[r] != vox_name[v]) {
        return false;
      }
    }
    if (nroots != (int)out.cluster_set.size()) return false;
  }

  TP(1);
  // ---- refineClusterByIntensity (ssc.cpp:571-635) at component granularity -------------------------
  // E(K): components reached from component K by a voxel pair passing the similarity test (:588-594)
  std::unordered_map<int, std::vector<int>> comp_edges;
  for (int e = 0; e < t.n_edges; ++e) comp_edges[t.edges[2 * e]].push_back(t.edges[2 * e + 1]);
  std::vector<int>& vox_label = out.vox_label;
  int iter = p.iteration;
  while (iter) {
    std::vector<std::pair<int, const HCluster*>> clusters;
    clusters.reserve(out.cluster_set.size());
    for (auto& c : out.cluster_set) clusters.push_back(std::make_pair(c.first, &c.second));
    // sort1 (:24-26): occupy_voxels compared with >= ; the vectors are pairwise different
    std::sort(clusters.begin(), clusters.end(), [](const std::pair<int, const HCluster*>& a, const std::pair<int, const HCluster*>& b) {
      return a.second->occupy_voxels > b.second->occupy_voxels;
    });
    std::vector<int> invalid_name;
    std::unordered_map<int, std::vector<int>> fusion_map;
    for (auto& c : clusters) {
      if (name_in(invalid_name, c.first)) continue;
      std::vector<int> neighbor_name;
      auto rit = roots_of.find(c.first);
      if (rit != roots_of.end()) {
        for (int K : rit->second) {
          auto eit = comp_edges.find(K);
          if (eit == comp_edges.end()) continue;
          for (int K2 : eit->second) {
            int lab = vox_label[K2];  // hash_cloud[n].label: every voxel of a component carries the same label
            if (!name_in(invalid_name, lab)) neighbor_name.push_back(lab);
          }
        }
      }
      sample_vec(neighbor_name);
      if (neighbor_name.size() > 1) {
        invalid_name.insert(invalid_name.end(), neighbor_name.begin(), neighbor_name.end());
        fusion_map.insert(std::make_pair(c.first, neighbor_name));
      }
      sample_vec(invalid_name);
    }
    for (auto& cn : fusion_map) {  // (:613-626)
      HCluster fusion;
      std::vector<int> froots;
      for (auto& f : cn.second) {
        fusion.name = f;
        HCluster& src = out.cluster_set[f];
        int basev = (int)fusion.occupy_voxels.size();
        fusion.occupy_voxels.insert(fusion.occupy_voxels.end(), src.occupy_voxels.begin(), src.occupy_voxels.end());
        for (int pe : src.part_end) fusion.part_end.push_back(basev + pe);
        fusion.npts += src.npts;
        auto rf = roots_of.find(f);
        if (rf != roots_of.end()) {
          froots.insert(froots.end(), rf->second.begin(), rf->second.end());
          roots_of.erase(rf);
        }
        out.cluster_set.erase(f);
      }
      for (int v : fusion.occupy_voxels) vox_label[v] = fusion.name;
      roots_of[fusion.name] = froots;
      out.cluster_set.insert(std::make_pair(fusion.name, std::move(fusion)));
    }
    iter--;
  }
  out.n_clusters[1] = (int)out.cluster_set.size();
  if (keep_stages) out.vox_name_stage[1] = vox_label;

  TP(2);
  // ---- refineClusterByBoundingBox (ssc.cpp:437-467) ------------------------------------------------
  std::vector<int> erase_id;
  for (auto& c : out.cluster_set) {
    HCluster& cl = c.second;
    float lo[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, hi[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    for (int v : cl.occupy_voxels) {
      const float* bb = t.vox_bbox + 6 * (size_t)v;
      for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(lo[d], bb[d]);
        hi[d] = std::max(hi[d], bb[3 + d]);
      }
    }
    for (int d = 0; d < 3; ++d) {
      cl.bb_min[d] = lo[d];
      cl.bb_max[d] = hi[d];
    }
    float diff_z = hi[2] - lo[2];
    if (lo[2] > 0.f || ((size_t)cl.npts < (size_t)p.toBeClass) || diff_z < 0.2) erase_id.emplace_back(c.first);
  }
  for (auto& e : erase_id) {
    for (auto& v : out.cluster_set[e].occupy_voxels) vox_label[v] = -1;
    out.cluster_set.erase(e);
  }
  out.n_clusters[2] = (int)out.cluster_set.size();
  if (keep_stages) out.vox_name_stage[2] = vox_label;

  TP(3);
  recognize_clusters(p, out);
  TP(4);
  return true;
}

// pcl::getTransformation in float (PCL 1.8 common/impl/eigen.hpp), call sites ssc.cpp:1255-1256
void pose_matrix(const float q[6], float T[12]) {
  float A = std::cos(q[5]), B = std::sin(q[5]), C = std::cos(q[4]), D = std::sin(q[4]), E = std::cos(q[3]), F = std::sin(q[3]);
  float DE = D * E, DF = D * F;
  T[0] = A * C;
  T[1] = A * DF - B * E;
  T[2] = B * F + A * DE;
  T[3] = q[0];
  T[4] = B * C;
  T[5] = A * E + B * DF;
  T[6] = B * DE - A * F;
  T[7] = q[1];
  T[8] = -D;
  T[9] = C * F;
  T[10] = C * E;
  T[11] = q[2];
}

void relative_pose(const float pose_next[6], const float pose_pre[6], float T[12]) {
  float N[12], Pm[12];
  pose_matrix(pose_next, N);
  pose_matrix(pose_pre, Pm);
  // Affine3f::inverse(): cofactor inverse of the linear part, t' = -(Linv * t)
  auto m = [&](int i, int j) { return N[4 * i + j]; };
  auto cof = [&](int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
  };
  float c0[3] = {cof(0, 0), cof(1, 0), cof(2, 0)};
  float det = (c0[0] * m(0, 0) + c0[1] * m(1, 0)) + c0[2] * m(2, 0);
  float invdet = 1.f / det;
  float I[12];
  I[0] = c0[0] * invdet;
  I[1] = c0[1] * invdet;
  I[2] = c0[2] * invdet;
  I[4] = cof(0, 1) * invdet;
  I[5] = cof(1, 1) * invdet;
  I[6] = cof(2, 1) * invdet;
  I[8] = cof(0, 2) * invdet;
  I[9] = cof(1, 2) * invdet;
  I[10] = cof(2, 2) * invdet;
  for (int i = 0; i < 3; ++i) I[4 * i + 3] = -((I[4 * i] * N[3] + I[4 * i + 1] * N[7]) + I[4 * i + 2] * N[11]);
  // product of two affine transforms
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = (I[4 * i] * Pm[j] + I[4 * i + 1] * Pm[4 + j]) + I[4 * i + 2] * Pm[8 + j];
    T[4 * i + 3] = ((I[4 * i] * Pm[3] + I[4 * i + 1] * Pm[7]) + I[4 * i + 2] * Pm[11]) + I[4 * i + 3];
  }
}

}  // namespace scvod
