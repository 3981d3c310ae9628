// scvod_internal.h — shared declarations between the CUDA kernels (scvod_ground.cu, scvod_voxel.cu, scvod_track.cu), the C-ABI
// layer (scvod_api.cpp) and the host-side cluster logic (host_cluster.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/scvod.h"

namespace scvod {

// PatchWork concentric zone model constants (reference include/patchwork.h:48-51,83-94,115-129).
constexpr int kNumZones = 4;
constexpr int kNumPatches = 504;  // 2*16 + 4*32 + 4*54 + 4*32
constexpr int kMinPatchPts = 10;  // num_min_pts_: patches need > 10 points (patchwork.h:331)

// Tainted voxels (a point with a -1 index hashes into a voxel that is not its own cell, ssc.cpp:185-188): per-scan capacities of
// the side tables that name such voxels point by point.  taint_cnt[scan][4] = {quirk points listed, tainted voxels, their points, error}
constexpr int kQuirkCap = 4096;   // points with a -1 index per scan
constexpr int kTvCap = 2048;      // tainted voxels per scan
constexpr int kTpCap = 16384;     // points of tainted voxels per scan
constexpr int kTaintCntStride = 4;
constexpr int kVirtualVox = 0x40000000;  // car CSR / tracking segments: "voxel" is a subgroup of a tainted voxel, points at tv_pts + (id & ~flag)

struct GridSpec {
  int range_num, sector_num, azimuth_num, bin_num;
  int key_off;    // voxel_idx + key_off >= 0 for every reachable (aliased) index, ssc.cpp:185-188
  int key_count;  // number of distinct keys
  int words;      // bitmap words per scan
};

// Per-batch device workspace. All per-point arrays are indexed by the batch-global point position
// (scan s occupies [off[s], off[s+1])); per-apri arrays reuse the same positions (M_s <= N_s).
struct BatchDev {
  int cap_points = 0;  // total points capacity of the batch
  int cap_scans = 0;
  // inputs
  float4* pts = nullptr;      // xyzi
  int64_t* off = nullptr;     // [cap_scans+1]
  // ground stage
  int16_t* patch_of = nullptr;     // per point: patch id or -1/-2 (dropped low / range)
  int32_t* patch_cnt = nullptr;    // [scans][504]
  int32_t* patch_off = nullptr;    // [scans][505] exclusive (relative to scan base)
  int32_t* patch_cur = nullptr;    // [scans][504] scatter cursors
  int32_t* sort_ctr = nullptr;     // [3][2]: count, cursor of the worklist of each upper sort tier
  int32_t* sort_list = nullptr;    // [3][cap_scans*504]: scan*504 + patch of every patch above 1024 points, by tier
  uint64_t* bucket_kv = nullptr;   // per bucket slot: (zkey<<32 | local idx)
  float4* sorted_xyz = nullptr;    // per bucket slot: the point, z-sorted inside its patch (k_patch_sort -> chain / rank)
  int32_t* sorted_idx = nullptr;   // per bucket slot after sort: local point idx
  int32_t* slot_pos = nullptr;     // per bucket slot: role<<30 | local position in ground / nonground list
  int32_t* slot_apos = nullptr;    // per bucket slot: local position in apri list or -1
  int32_t* slot_vid = nullptr;     // per bucket slot: voxel_idx (apri points)
  int16_t* slot_patch = nullptr;   // per bucket slot: patch id
  int32_t* patch_out = nullptr;    // [scans][504][8]: n_ground, n_nonground, n_apri, quirk count, |ground set|, its gate-passing part, rejected
  int32_t* patch_out_off = nullptr;  // [scans][505][3] exclusive offsets
  float* patch_dbg = nullptr;      // [scans][504][12] normal, mean, sv, d, decision, npts (debug/inspection)
  int32_t* scan_counts = nullptr;  // [scans][8]: n_ground, n_ng, n_apri, n_voxels, n_quirk, n_events, n_comp, n_edges
  // outputs of the ground stage
  int32_t* ground_src = nullptr;  // per scan base: original idx in cloud_out order
  int32_t* ng_src = nullptr;      // ... in cloud_nonground order
  uint8_t* cls = nullptr;         // per point outcome class
  // apri arrays (indexed by scan base + m)
  int32_t* apri_src = nullptr;
  int32_t* apri_vid = nullptr;
  float4* apri_xyzi = nullptr;
  int32_t* apri_cid = nullptr;   // compact voxel id
  int32_t* apri_rank = nullptr;  // rank of the point inside its voxel (ascending m)
  // voxel structures
  uint32_t* bitmap = nullptr;     // [scans][words]
  int32_t* word_rank = nullptr;   // [scans][words] exclusive popcount prefix
  int32_t* vox_cnt = nullptr;     // per scan base + cid
  int32_t* vox_off = nullptr;     // per scan base + cid (exclusive, relative to scan base); +1 slack handled by cnt
  int32_t* vox_cur = nullptr;
  int32_t* vox_pts_tmp = nullptr;  // unsorted fill
  int32_t* vox_pts = nullptr;      // CSR point ids (m), ascending inside each voxel
  int32_t* vox_vid = nullptr;
  float* vox_av = nullptr;
  float* vox_cov = nullptr;
  float* vox_center = nullptr;  // 3 per voxel
  int32_t* vox_tri = nullptr;   // 3 per voxel
  int32_t* vox_nbr = nullptr;   // 27 per voxel: compact ids in findVoxelNeighbors order, -1 = absent
  int32_t* vox_root = nullptr;  // CCL root (min compact id of the component)
  float* vox_bbox = nullptr;    // 6 per voxel (per-voxel bbox; reduced per component on the host)
  int32_t* ev_cid = nullptr;    // clustering events: compact voxel id of every apri point with rank < 3, in m order
  int32_t* edge_buf = nullptr;  // [scans][edge_cap][2] directed similar-intensity component edges (roots)
  int32_t* edge_hash = nullptr; // [scans][hash_cap] dedupe table
  int edge_cap = 0, hash_cap = 0;
  // tainted voxels (see kQuirkCap)
  int32_t* taint_cnt = nullptr;  // [scans][kTaintCntStride]
  int32_t* q_list = nullptr;     // [scans][kQuirkCap] apri index of the points with a -1 index (unordered)
  int32_t* vox_tnt = nullptr;    // per voxel: -1 ordinary, -3 ordinary with a tainted voxel among its 27 neighbours, >= 0 tainted: offset of its points in tp_*
  int32_t* vox_group = nullptr;  // per voxel: components of the voxel graph INCLUDING the links through tainted voxels (replay work partition)
  int32_t* tv_cid = nullptr;     // [scans][kTvCap] tainted voxels, ascending compact id
  int32_t* tv_base = nullptr;    // [scans][kTvCap + 1]
  int32_t* tp_m = nullptr;       // [scans][kTpCap] apri index of every point of a tainted voxel (voxel-major, ascending m)
  int32_t* tp_cid = nullptr;     // [scans][kTpCap] compact id of the point's voxel
  float4* tp_xyz = nullptr;      // [scans][kTpCap]
  int32_t* tp_name = nullptr;    // [scans][kTpCap] cluster name of the point (k_name_replay)
};

struct PackDesc {
  const int32_t* src;
  long long dst;
  int n;
  int pad;
};

struct HostParams {
  scvod_params p;
  GridSpec g;
  bool chain_tma = false;  // k_patch_chain: stage the ring with cp.async.bulk + mbarrier instead of per-lane cp.async
  int track_ctas_per_sm = 1;  // grid cap of k_track (set per launch from the number of live contexts)
};

// kernel launch wrappers (scvod_ground.cu / scvod_voxel.cu / scvod_track.cu). All asynchronous on `stream`. Return launch count.
int launch_ground(const HostParams& hp, BatchDev& d, int nscans, int total_points, void* stream);
int launch_descriptor(const HostParams& hp, BatchDev& d, int nscans, int total_points, void* stream);
int launch_cluster_prep(const HostParams& hp, BatchDev& d, int nscans, int total_points, void* stream);
int launch_bin_only(const HostParams& hp, const float4* pts_dev, int n, uint8_t* pass, int32_t* vid, int32_t* ri, int32_t* si,
                    int32_t* ei, float* range, float* angle, float* azimuth, void* stream);
// Run table of one tracking launch, passed in the kernel arguments (no upload): run r covers points [dst_off[r], dst_off[r+1]);
// src >= 0: own voxels, positions [src, src + len) of the frame's car CSR, part order = order + csr_part; src < 0: carried
// range at -1 - src of the previous launch's output, order = 0x40000000 + ordinal.
constexpr int kTrackMaxRuns = 176;
struct TrackRuns {
  int n;
  int dst_off[kTrackMaxRuns];
  int src[kTrackMaxRuns];
  int len[kTrackMaxRuns];
  int order[kTrackMaxRuns];
  int cluster[kTrackMaxRuns];
};
// tracking diff of one frame pair.  runs != nullptr: run table (above) + the frame's car CSR; else uploaded segments
// {dst_off, source (>=0 voxel of frame_pre_ / <0 carried range), cluster, order} and first_seg[b] = segment holding point 256*b
int launch_track(const HostParams& hp, const float4* own_xyzi, const int32_t* vox_off, const int32_t* vox_pts, const float4* carried,
                 const int4* segs, const int32_t* first_seg, int nseg, const TrackRuns* runs, const int32_t* csr_ptoff, const int32_t* csr_vox,
                 const int32_t* csr_part, const int32_t* tv_pts, int k, const float T12[12], const uint32_t* next_bitmap, const int32_t* next_word_rank, int ncl,
                 int vn, float4* out_xyzi, unsigned long long* first, int32_t* ctr_dev, int32_t* hit_list_dev, int32_t* out_quads_mapped,
                 int cap_quads, void* stream);
int launch_final_labels(const int64_t* off, const int32_t* scan_counts, int nscans, int max_scan_points, const int32_t* apri_src,
                        const int32_t* apri_cid, const int32_t* vcls_off, const uint8_t* vcls, uint8_t* cls, void* stream);
// classes of the points of tainted voxels: items = (batch-global apri position, class), item_scan = slot of the point's scan
int launch_label_override(const int32_t* items_dev, const int32_t* item_scan_dev, int n, const int32_t* apri_src, const int64_t* off, uint8_t* cls,
                          void* stream);
int launch_submap(const float4* pts, const uint8_t* cls, const int64_t* off, const float* Ts_dev, int first_scan, int nscans,
                  int max_scan_points, float4* out, unsigned long long* counter, long long cap, unsigned keep_mask, void* stream);
// cluster-name replay (one CTA per scan, components dealt to its warps); vox_name is indexed like the voxel arrays,
// name_first is [nscans][name_cap]
// max_nodes = max over the scans of voxels + points of tainted voxels; any_taint: some scan has tainted voxels
int launch_name_replay(const HostParams& hp, BatchDev& d, int nscans, int max_nodes, int max_events, bool force_global, bool any_taint,
                       int32_t* vox_name, int32_t* name_first, int name_cap, void* stream);
int launch_pack(const PackDesc* descs_dev, int ndesc, int max_n, int32_t* out, void* stream);
int launch_patch_filter_check(const HostParams& hp, long long n, uint32_t seed, float extent, unsigned long long* stats_dev, void* stream);
int launch_atan2f_probe(const float* y, const float* x, float* out, long long n, void* stream);
int launch_bin_filter_check(const HostParams& hp, long long n, uint32_t seed, float extent, const float4* pts_dev, unsigned long long* stats_dev,
                            void* stream);

// context accessors for the translation units that do not see struct scvod_ctx (scvod_gicp.cu)
void* ctx_stream(scvod_ctx* c);
int ctx_device(const scvod_ctx* c);
void ctx_add_launches(scvod_ctx* c, int n);
void** ctx_gicp_slot(scvod_ctx* c);                          // opaque GICP state owned by scvod_gicp.cu
void ctx_set_gicp_free(scvod_ctx* c, void (*fn)(void*));     // called by scvod_destroy
int api_fail(int code, const std::string& msg);               // sets scvod_last_error()

// RAII CUDA-event timer around a kernel launch (active only while scvod_kernel_timing is enabled)
struct LaunchTimer {
  void* st;
  int id;
  void *e0, *e1;
  bool on;
  LaunchTimer(const char* name, void* stream);
  ~LaunchTimer();
};

// per-kernel CUDA-event timing (process-wide; used by bench.py for the roofline block)
void timing_enable(bool on);
void timing_collect();
void timing_reset();
std::string timing_report();  // lines of "<kernel> <total ms> <launches>"

}  // namespace scvod
