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

// Hybrid replay.  Nodes are the voxels [0, V) and, for the voxels that hold a point with a -1 index ("tainted", see
// ScanTables), their points one by one [V, V + n_tpts): such a voxel is seen by visitors as the sequence of its points
// (ptIdx order), and each of its points visits with its OWN findVoxelNeighbors list (which need not contain the voxel it
// hashes into, ssc.cpp:185-188,311).  Every point of a tainted voxel is an event; ordinary voxels keep the three-state logic.
int replay_cluster_names(const ScanTables& t, std::vector<int>& vox_name, std::vector<int>& tp_name) {
  const int V = t.V, N = V + t.n_tpts;
  NameReplay u;
  u.parent.resize(N);
  u.setname.assign(N, -1);
  u.state.assign(N, 0);
  u.stable.assign(N, 0);
  for (int v = 0; v < N; ++v) u.parent[v] = v;
  std::vector<int> tbase(t.n_tvox ? V : 0, -1), tcnt(t.n_tvox ? V : 0, 0);
  for (int i = 0; i < t.n_tvox; ++i) {
    tbase[t.tv_cid[i]] = t.tv_base[i];
    tcnt[t.tv_cid[i]] = t.tv_base[i + 1] - t.tv_base[i];
  }
  std::vector<int> seq;
  int cluster_name = 4;  // ssc.cpp:300
  for (int e = 0; e < t.n_events; ++e) {
    const int W = t.ev_cid[e];
    const bool is_pt = W >= V;
    if (!is_pt && u.stable[W]) continue;
    const int32_t* nb = is_pt ? t.tp_nbr + 27 * (size_t)(W - V) : t.vox_nbr + 27 * (size_t)W;
    const bool labelled = (u.state[W] == 2);
    int oc = labelled ? u.find(W) : -1;
    bool skipped = false;
    seq.clear();
    for (int k = 0; k < 27; ++k) {  // the visiting sequence: whole ordinary voxels, tainted voxels point by point
      const int Vn = nb[k];
      if (Vn < 0) continue;
      if (t.n_tvox && tbase[Vn] >= 0) {
        for (int j = 0; j < tcnt[Vn]; ++j) seq.push_back(V + tbase[Vn] + j);
      } else {
        seq.push_back(Vn);
      }
    }
    for (int Vn : seq) {
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
      for (int Vn : seq) {
        if (Vn == W) continue;
        u.parent[Vn] = W;
        u.state[Vn] = 2;
      }
      u.stable[W] = 1;
    } else if (is_pt) {
      if (u.state[W] == 0) {
        u.state[W] = 2;
        u.parent[W] = oc;
      }
    } else {
      if (u.state[W] == 0) {  // only this (first) point of W got the label
        u.state[W] = 1;
        u.parent[W] = oc;
      }
      u.stable[W] = skipped ? 0 : 1;
    }
  }
  vox_name.resize(V);
  for (int v = 0; v < V; ++v) vox_name[v] = (t.n_tvox && tbase[v] >= 0) ? -1 : u.setname[u.find(v)];
  tp_name.resize(t.n_tpts);
  for (int q = 0; q < t.n_tpts; ++q) tp_name[q] = u.setname[u.find(V + q)];
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
  const bool taint = t.n_tvox > 0;
  // ---- clusterAndCreateFrame (ssc.cpp:299-393) ---------------------------------------------------
  std::vector<int>& vox_name = out.vox_label;  // frame_ssc hash_cloud labels start out as the cluster names (:387-391)
  std::vector<int> tp_name;
  int last_name;
  // cluster_pt is filled in point order, so its keys are inserted in order of each cluster's first point (:360-375)
  std::unordered_map<int, int> cluster_pt;
  if (t.vox_name) {  // names replayed on the device
    vox_name.assign(t.vox_name, t.vox_name + V);
    if (taint) tp_name.assign(t.tp_name, t.tp_name + t.n_tpts);
    last_name = t.max_name;
    std::vector<std::pair<int, int>> order;  // (first event, name)
    for (int nm = 0; nm <= last_name; ++nm)
      if (t.name_first[nm] != 0x7fffffff) order.push_back(std::make_pair(t.name_first[nm], nm));
    std::sort(order.begin(), order.end());
    for (auto& o : order) cluster_pt.insert(std::make_pair(o.second, 0));
  } else {
    last_name = replay_cluster_names(t, vox_name, tp_name);
    std::vector<char> seen(last_name + 2, 0);
    for (int e = 0; e < t.n_events; ++e) {
      const int node = t.ev_cid[e];
      int nm = node >= V ? tp_name[node - V] : vox_name[node];
      if (nm < 0) return false;
      if (!seen[nm]) {
        seen[nm] = 1;
        cluster_pt.insert(std::make_pair(nm, 0));
      }
    }
  }
  TP(0);
  out.max_name = last_name;  // frame_ssc.max_name = cluster_name++ (:354)
  out.cluster_set.clear();
  out.tvox.clear();
  out.sgs_of_tvox.clear();
  out.subgroups.clear();
  // per-thread scratch (one context = one host thread at a time; worker threads of a batch each have their own): the tables below are
  // O(V) per scan and were re-allocated for every scan
  static thread_local std::vector<uint8_t> tflag;
  static thread_local std::vector<int> name_start, vox_sorted, cur, root_name;
  static thread_local std::vector<std::vector<int>> roots_tab;  // cluster name -> classes of voxels it contains (see roots_of below)
  if (taint) {
    tflag.assign(V, 0);
    for (int i = 0; i < t.n_tvox; ++i) {
      if (t.tv_cid[i] < 0 || t.tv_cid[i] >= V || (i && t.tv_cid[i] <= t.tv_cid[i - 1])) return false;
      tflag[t.tv_cid[i]] = 1;
    }
    for (int q = 0; q < t.n_tpts; ++q)
      if (tp_name[q] < 0 || tp_name[q] > last_name) return false;
  }
  // voxels of every cluster in ascending compact id (== sorted voxel_idx): counting sort by name (the range check of the names
  // rides on the counting pass)
  name_start.assign(last_name + 3, 0);
  vox_sorted.resize(V);
  for (int v = 0; v < V; ++v) {
    if (taint && tflag[v]) continue;
    const int nm = vox_name[v];
    if (nm < 0 || nm > last_name) return false;
    name_start[nm + 1]++;
  }
  for (int nm = 0; nm <= last_name + 1; ++nm) name_start[nm + 1] += name_start[nm];
  cur.assign(name_start.begin(), name_start.end() - 1);
  for (int v = 0; v < V; ++v)
    if (!(taint && tflag[v])) vox_sorted[cur[vox_name[v]]++] = v;
  // tainted voxels: their points grouped by the name clusterAndCreateFrame gave them
  std::unordered_map<int, std::vector<int>> sgs_of_name;  // name -> subgroups, ascending voxel
  if (taint) {
    out.tvox.assign(t.tv_cid, t.tv_cid + t.n_tvox);
    out.sgs_of_tvox.resize(t.n_tvox);
    for (int i = 0; i < t.n_tvox; ++i) {
      for (int q = t.tv_base[i]; q < t.tv_base[i + 1]; ++q) {
        int sg = -1;
        for (int k : out.sgs_of_tvox[i])
          if (out.subgroups[k].stage_name[0] == tp_name[q]) sg = k;
        if (sg < 0) {
          sg = (int)out.subgroups.size();
          out.subgroups.emplace_back();
          SubGroup& g = out.subgroups.back();
          g.vox = t.tv_cid[i];
          g.stage_name[0] = tp_name[q];
          for (int d = 0; d < 3; ++d) {
            g.bb_min[d] = 3.402823466e38f;
            g.bb_max[d] = -3.402823466e38f;
          }
          out.sgs_of_tvox[i].push_back(sg);
          sgs_of_name[tp_name[q]].push_back(sg);
        }
        SubGroup& g = out.subgroups[sg];
        g.pts.push_back(t.tp_m[q]);
        for (int d = 0; d < 3; ++d) {
          g.bb_min[d] = std::min(g.bb_min[d], t.tp_xyz[4 * (size_t)q + d]);
          g.bb_max[d] = std::max(g.bb_max[d], t.tp_xyz[4 * (size_t)q + d]);
        }
      }
    }
  }
  // cluster name -> classes of voxels it contains: roots of the CVC components of ordinary voxels, tainted voxels one by one.
  // A table indexed by name (names are <= last_name); an empty row = "no entry".  Its iteration order is never observed.
  if (roots_tab.size() < (size_t)last_name + 2) roots_tab.resize(last_name + 2);
  for (int nm = 0; nm <= last_name + 1; ++nm) roots_tab[nm].clear();
  const int roots_n = last_name + 2;
  static const std::vector<int> no_roots;
  auto roots_get = [&](int nm) -> const std::vector<int>& { return (nm >= 0 && nm < roots_n) ? roots_tab[nm] : no_roots; };
  auto& roots_of = roots_tab;
  for (auto& c : cluster_pt) {  // (:377-385) same iteration order as the reference's cluster_pt
    HCluster cl;
    cl.name = c.first;
    cl.occupy_voxels.assign(vox_sorted.begin() + name_start[c.first], vox_sorted.begin() + name_start[c.first + 1]);
    int np = 0;
    for (int v : cl.occupy_voxels) np += t.vox_cnt[v];
    if (!taint) {
      if (!cl.occupy_voxels.empty()) roots_of[c.first].push_back(t.vox_root[cl.occupy_voxels[0]]);
    } else {
      std::vector<int>& ro = roots_of[c.first];
      for (int v : cl.occupy_voxels) ro.push_back(t.vox_root[v]);
      sample_vec(ro);
      auto sit = sgs_of_name.find(c.first);
      if (sit != sgs_of_name.end()) {
        std::vector<int> tv;
        for (int sg : sit->second) {
          cl.tunits.push_back(HCluster::TUnit{sg, 0});
          np += (int)out.subgroups[sg].pts.size();
          tv.push_back(out.subgroups[sg].vox);
          ro.push_back(out.subgroups[sg].vox);
        }
        std::vector<int> merged(cl.occupy_voxels.size() + tv.size());  // sampleVec(occupy_voxels), :383
        std::merge(cl.occupy_voxels.begin(), cl.occupy_voxels.end(), tv.begin(), tv.end(), merged.begin());
        cl.occupy_voxels.swap(merged);
      }
    }
    cl.part_end.push_back((int)cl.occupy_voxels.size());
    cl.npts = np;
    out.cluster_set.insert(std::make_pair(cl.name, std::move(cl)));
  }
  std::vector<int>& vox_label = out.vox_label;  // (== vox_name: from here on the names are only read through the clusters)
  if (taint) {  // hash_cloud[v].label = c.first in cluster_set order (:387-391): a tainted voxel keeps the LAST cluster that lists it
    for (int i = 0; i < t.n_tvox; ++i) vox_label[t.tv_cid[i]] = -1;
    for (auto& c : out.cluster_set)
      for (auto& tu : c.second.tunits) vox_label[out.subgroups[tu.sg].vox] = c.first;
  }
  out.n_clusters[0] = (int)out.cluster_set.size();
  if (keep_stages) out.vox_name_stage[0] = vox_label;
  // the replayed names must be constant on the GPU's connected components of ordinary voxels (and, without tainted voxels,
  // the two partitions coincide)
  {
    root_name.assign(V, -1);
    int nroots = 0;
    for (int v = 0; v < V; ++v) {
      if (taint && tflag[v]) continue;
      int r = t.vox_root[v];
      if (r < 0 || r >= V) return false;
      if (root_name[r] == -1) {
        root_name[r] = vox_name[v];
        ++nroots;
      } else if (root_name[r] != vox_name[v]) {
        return false;
      }
    }
    if (!taint && nroots != (int)out.cluster_set.size()) return false;
  }

  TP(1);
  // ---- refineClusterByIntensity (ssc.cpp:571-635) at component granularity -------------------------
  // E(K): classes reached from class K by a voxel pair passing the similarity test (:588-594).  All voxels of a class
  // carry the same label at any time (an ordinary component lies inside one cluster and labels are written per cluster).
  std::unordered_map<int, std::vector<int>> comp_edges;
  for (int e = 0; e < t.n_edges; ++e) comp_edges[t.edges[2 * e]].push_back(t.edges[2 * e + 1]);
  int iter = p.iteration;
  while (iter) {
    std::vector<std::pair<int, const HCluster*>> clusters;
    clusters.reserve(out.cluster_set.size());
    for (auto& c : out.cluster_set) clusters.push_back(std::make_pair(c.first, &c.second));
    // sort1 (:24-26): occupy_voxels compared with >= ; the vectors are pairwise different (equal only for two clusters that
    // consist of points of the same tainted voxels: std::sort with >= on equal elements is not reproduced, they stay adjacent)
    std::sort(clusters.begin(), clusters.end(), [](const std::pair<int, const HCluster*>& a, const std::pair<int, const HCluster*>& b) {
      return a.second->occupy_voxels > b.second->occupy_voxels;
    });
    std::vector<int> invalid_name;
    std::unordered_map<int, std::vector<int>> fusion_map;
    for (auto& c : clusters) {
      if (name_in(invalid_name, c.first)) continue;
      std::vector<int> neighbor_name;
      {
        for (int K : roots_get(c.first)) {
          auto eit = comp_edges.find(K);
          if (eit == comp_edges.end()) continue;
          for (int K2 : eit->second) {
            int lab = vox_label[K2];  // hash_cloud[n].label
            if (!name_in(invalid_name, lab)) neighbor_name.push_back(lab);
          }
        }
      }
      if (neighbor_name.empty()) continue;  // (invalid_name is already sorted and unique: sampleVec would not change it)
      sample_vec(neighbor_name);
      if (neighbor_name.size() > 1) {
        invalid_name.insert(invalid_name.end(), neighbor_name.begin(), neighbor_name.end());
        fusion_map.insert(std::make_pair(c.first, neighbor_name));
        sample_vec(invalid_name);
      }
    }
    for (auto& cn : fusion_map) {  // (:613-626)
      HCluster fusion;
      std::vector<int> froots;
      for (auto& f : cn.second) {
        fusion.name = f;
        HCluster& src = out.cluster_set[f];
        int basev = (int)fusion.occupy_voxels.size();
        const int base_parts = (int)fusion.part_end.size();
        fusion.occupy_voxels.insert(fusion.occupy_voxels.end(), src.occupy_voxels.begin(), src.occupy_voxels.end());
        for (int pe : src.part_end) fusion.part_end.push_back(basev + pe);
        for (auto tu : src.tunits) fusion.tunits.push_back(HCluster::TUnit{tu.sg, base_parts + tu.part});
        fusion.npts += src.npts;
        froots.insert(froots.end(), roots_get(f).begin(), roots_get(f).end());
        if (f >= 0 && f < roots_n) roots_of[f].clear();
        out.cluster_set.erase(f);
      }
      for (int v : fusion.occupy_voxels) vox_label[v] = fusion.name;
      if (fusion.name >= 0 && fusion.name < roots_n) roots_of[fusion.name] = froots;
      out.cluster_set.insert(std::make_pair(fusion.name, std::move(fusion)));
    }
    if (fusion_map.empty()) break;  // nothing changed: every further iteration would find the same (no) fusions
    iter--;
  }
  out.n_clusters[1] = (int)out.cluster_set.size();
  if (keep_stages) {
    out.vox_name_stage[1] = vox_label;
    for (auto& c : out.cluster_set)
      for (auto& tu : c.second.tunits) out.subgroups[tu.sg].stage_name[1] = c.first;
  }

  TP(2);
  // ---- refineClusterByBoundingBox (ssc.cpp:437-467) ------------------------------------------------
  std::vector<int> erase_id;
  for (auto& c : out.cluster_set) {
    HCluster& cl = c.second;
    float lo[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, hi[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    for (int v : cl.occupy_voxels) {
      if (taint && tflag[v]) continue;  // only the cluster's own points of a tainted voxel count: below
      const float* bb = t.vox_bbox + 6 * (size_t)v;
      for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(lo[d], bb[d]);
        hi[d] = std::max(hi[d], bb[3 + d]);
      }
    }
    for (auto& tu : cl.tunits) {
      const SubGroup& g = out.subgroups[tu.sg];
      for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(lo[d], g.bb_min[d]);
        hi[d] = std::max(hi[d], g.bb_max[d]);
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
  if (keep_stages) {
    out.vox_name_stage[2] = vox_label;
    for (auto& c : out.cluster_set)
      for (auto& tu : c.second.tunits) out.subgroups[tu.sg].stage_name[2] = c.first;
  }

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
