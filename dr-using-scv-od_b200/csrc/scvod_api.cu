// scvod_api.cu — the extern "C" layer declared in include/scvod.h: context, device memory, the batched
// frame pipeline (GPU stages + host cluster bookkeeping) and the tracking chain.
//
// There is no CPU fallback anywhere in this file: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "host_cluster.h"
#include "scvod_internal.h"

using namespace scvod;

static thread_local std::string g_err;
static std::atomic<int> g_live_contexts(0);  // contexts alive in this process
// contexts that are inside a tracking chain right now: sizes the footprint of the latency-bound tracking kernel (a lone chain may
// fill the GPU for the shortest round trip; many concurrent chains take one CTA per SM each so that their kernels co-run)
static std::atomic<int> g_tracking_now(0);
struct TrackingScope {
  TrackingScope() { g_tracking_now.fetch_add(1); }
  ~TrackingScope() { g_tracking_now.fetch_sub(1); }
};
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                                             \
  do {                                                                                                       \
    cudaError_t e__ = (call);                                                                                \
    if (e__ != cudaSuccess) return fail(SCVOD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

namespace {

// host-side section timer (SCVOD_PROFILE=1 prints the accumulated breakdown at scvod_destroy): wall time and the CPU time of the
// calling thread per section (a section that waits for the GPU asleep has wall >> cpu; one that spins has wall == cpu)
static double thread_cpu_ms() {
  timespec ts;
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
struct HostProf {
  bool on = getenv("SCVOD_PROFILE") != nullptr;
  struct Acc {
    std::string name;
    double wall = 0, cpu = 0;
    long long calls = 0;
  };
  std::vector<Acc> acc;
  std::mutex mu;
  void add(const char* name, double ms, double cpu_ms = 0) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& a : acc)
      if (a.name == name) {
        a.wall += ms;
        a.cpu += cpu_ms;
        a.calls++;
        return;
      }
    Acc a;
    a.name = name;
    a.wall = ms;
    a.cpu = cpu_ms;
    a.calls = 1;
    acc.push_back(a);
  }
  void reset() {
    std::lock_guard<std::mutex> lk(mu);
    acc.clear();
  }
  void dump() {
    if (!on) return;
    for (auto& a : acc)
      fprintf(stderr, "[scvod profile] %-28s wall %10.3f ms  cpu %10.3f ms  calls %lld\n", a.name.c_str(), a.wall, a.cpu, a.calls);
  }
};
HostProf g_prof;
struct ProfScope {
  const char* name;
  std::chrono::steady_clock::time_point t0;
  double c0;
  explicit ProfScope(const char* n) : name(n), t0(std::chrono::steady_clock::now()), c0(g_prof.on ? thread_cpu_ms() : 0) {}
  ~ProfScope() {
    if (g_prof.on)
      g_prof.add(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), thread_cpu_ms() - c0);
  }
};
#define PROF_CAT2(a, b) a##b
#define PROF_CAT(a, b) PROF_CAT2(a, b)
#define PROF(name) ProfScope PROF_CAT(prof_scope__, __COUNTER__)(name)

std::atomic<long long> g_reallocs(0);  // device / pinned (re)allocations so far in this process (scvod_get_stat "reallocs")

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    g_reallocs.fetch_add(1, std::memory_order_relaxed);
    if (p) {
      cudaFree(p);
      count += count / 2 + 1024;  // geometric growth: buffers that grow step by step are not re-allocated every time
    }
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    g_reallocs.fetch_add(1, std::memory_order_relaxed);
    if (p) {
      cudaFreeHost(p);
      count += count / 2 + 1024;
    }
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMallocHost((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
};

// device data that must outlive the batch: what tracking, labelling, the submap and inspection read
struct PersistBatch {
  int nscans = 0;
  int64_t total = 0;
  std::vector<int64_t> off;
  DevBuf<float4> pts, apri_xyzi;
  DevBuf<uint8_t> cls;
  DevBuf<int32_t> apri_src, apri_cid, apri_vid, word_rank, vox_off, vox_pts, vox_cnt, ground_src, ng_src;
  DevBuf<uint32_t> bitmap;
  DevBuf<int64_t> off_dev;           // device copy of off
  DevBuf<int32_t> scan_counts_dev;   // device copy of the per-scan counters
  // car CSR of the batch, three arrays of csr_n ints: point offsets (one extra entry per frame), voxel ids, part indices;
  // car clusters reference it by runs (HCluster::own_runs), see diff_clusters
  DevBuf<int32_t> csr;
  PinBuf<int32_t> h_csr;
  int csr_n = 0;
  int tv_n = 0;  // 4th section of csr: apri indices of the subgroups of tainted voxels (FrameHost::sg_tv_off)
  int max_n = 0;
  bool labels_current = false;
  // inspection-only
  DevBuf<int32_t> vox_vid, vox_tri;
  DevBuf<float> vox_av, vox_cov, vox_center;
  void release() {
    pts.release();
    apri_xyzi.release();
    cls.release();
    apri_src.release();
    apri_cid.release();
    apri_vid.release();
    word_rank.release();
    vox_off.release();
    vox_pts.release();
    vox_cnt.release();
    ground_src.release();
    ng_src.release();
    bitmap.release();
    off_dev.release();
    scan_counts_dev.release();
    csr.release();
    h_csr.release();
    vox_vid.release();
    vox_tri.release();
    vox_av.release();
    vox_cov.release();
    vox_center.release();
  }
};

struct FrameHost {
  int csr_base = -1;  // first position of this frame in its batch's car CSR (PersistBatch::csr), -1: none
  int batch = -1, slot = -1;
  int64_t base = 0;
  int n_in = 0, n_ground = 0, n_ng = 0, n_apri = 0, n_vox = 0;
  FrameClusters fc;
  std::vector<int32_t> vox_cnt;
  std::vector<int> sg_tv_off;  // per subgroup of fc: offset of its points in the batch's tv_pts
};

}  // namespace

struct scvod_ctx {
  HostParams hp;
  int device = 0;
  int max_points = 0, max_batch = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int64_t launches = 0;
  int host_threads = 1;
  bool inspect = true;
  bool replay_global = false;  // test hook: force the global-memory variant of k_name_replay
  bool submap_all_static = false;  // scvod_static_submap_dev: every non-dynamic input point instead of the instance map (cluster points)

  BatchDev ws;  // transient workspace (pointers into the DevBufs below and into the current PersistBatch)
  DevBuf<int64_t> d_off;
  DevBuf<int16_t> d_patch_of, d_slot_patch;
  DevBuf<int32_t> d_patch_cnt, d_patch_off, d_patch_cur, d_sorted_idx, d_slot_pos, d_slot_apos, d_slot_vid, d_patch_out,
      d_patch_out_off, d_scan_counts, d_sort_ctr, d_sort_list, d_apri_rank, d_vox_cur, d_vox_pts_tmp, d_vox_nbr, d_vox_root, d_ev_cid, d_edge_buf;
  DevBuf<uint64_t> d_bucket_kv, d_edge_hash;
  DevBuf<float4> d_sorted_xyz;
  DevBuf<float> d_patch_dbg, d_vox_bbox, d_T;
  // tainted voxels (points with a -1 index, see scvod_internal.h): side tables of the batch
  DevBuf<int32_t> d_taint_cnt, d_q_list, d_vox_tnt, d_vox_group, d_tv_cid, d_tv_base, d_tp_m, d_tp_cid, d_tp_name;
  DevBuf<float4> d_tp_xyz;
  PinBuf<int32_t> h_taint_cnt;
  // pinned host mirrors of what the host logic reads per batch
  PinBuf<int32_t> h_scan_counts, h_vox_cnt, h_vox_root, h_vox_nbr, h_ev_cid, h_edge_buf;
  PinBuf<float> h_vox_bbox;
  PinBuf<int32_t> h_pack;
  DevBuf<int32_t> d_pack;
  PinBuf<PackDesc> h_desc;
  DevBuf<PackDesc> d_desc;
  DevBuf<int32_t> d_vox_name, d_name_first;
  int name_cap = 0;
  // tracking buffers: packed request (segments + own indices), ping-pong transformed clouds, hit table
  DevBuf<int32_t> d_treq, d_triples, d_track_ctr, d_track_list;
  DevBuf<unsigned long long> d_first;
  PinBuf<int32_t> h_treq, h_triples;  // segment table + its per-block index; hit quads (written by the kernel)
  DevBuf<float4> d_tout[2];
  int tout_cur = 0;  // d_tout[tout_cur] holds the carried clouds of the frame that is the next frame_pre_
  // host scratch of track_pair, kept across pairs so that the steady state allocates nothing
  struct LabelAcc {
    int lab = -1;
    uint64_t min_key = 0;
    std::vector<int> vox;
  };
  TrackRuns track_runs;
  std::vector<int> hit_start, hit_cur, hit_vox, label_order;
  std::vector<uint64_t> hit_key;
  std::vector<LabelAcc> label_accs;
  // label refresh staging
  DevBuf<int32_t> d_vcls, d_lov;
  PinBuf<int32_t> h_vcls, h_lov;  // per-voxel classes; (apri position, class, scan) of the points of tainted voxels
  DevBuf<float> d_Ts;
  PinBuf<float> h_Ts;
  DevBuf<unsigned long long> d_counter;

  std::vector<std::unique_ptr<PersistBatch>> batches;
  std::vector<std::unique_ptr<PersistBatch>> batch_pool;  // released batches kept for reuse (no cudaMalloc in steady state)
  std::vector<FrameHost> frames;
  int tracked = 0;  // frames [0, tracked) have been used as frame_pre_
  bool head_tracked = false;  // frame 0 has been frame_next_ of an imported tail (scvod_track_from_tail)
  // The tracking chain runs on its own stream; SCVOD_TRACK_PRIORITY=1 gives it the highest priority (a k_track launch is tiny and
  // latency critical - the host waits for its answer before it can decide the pair - while the per-scan kernels of the other
  // contexts are bulk work; the block scheduler then hands freed CTA slots to the chain first).
  cudaStream_t tstream = nullptr;  // == stream unless own_tstream
  bool own_tstream = false;
  cudaEvent_t tlink = nullptr;
  cudaEvent_t sync_event = nullptr;  // cudaEventBlockingSync: waiting host threads sleep instead of spinning (see wait_stream)
  // scvod_prefetch_scans: upload of the next batch on a private stream, consumed by the next scvod_push_scans of the same buffer
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t prefetch_done = nullptr;
  DevBuf<float4> d_prefetch;
  const char* prefetch_host = nullptr;  // first prefetched byte in the caller's buffer
  size_t prefetch_bytes = 0;
  FrameClusters init_fc;  // frame_based after scvod_initialization (read with frame index SCVOD_INIT_FRAME)
  int init_base = -1;
  bool have_init = false;
  int track_name = 0;  // SSC::name (ssc.h:49)
  int failed_scan = -1;  // index (in the offending push call) of the scan that made the call fail, -1: none / not attributable to one scan
  int batch_first_scan = 0;  // position of the batch being pushed inside its push call
  int64_t stat_track_points = 0, stat_track_pairs = 0, stat_scans = 0, stat_points = 0, stat_apri = 0, stat_voxels = 0, stat_tvox = 0, stat_tpts = 0;
  void* gicp = nullptr;  // GICP state (scvod_gicp.cu)
  void (*gicp_free)(void*) = nullptr;
};

// Wait for everything queued on the context's stream.  The wait sleeps on a blocking-sync event instead of spinning in
// cudaStreamSynchronize: with several contexts per GPU (one host thread each) and few host cores per GPU, a thread that
// waits for the GPU must leave its core to the threads that have cluster bookkeeping to do.
static cudaError_t wait_stream(scvod_ctx* c, cudaStream_t st) {
  if (!c->sync_event) {
    cudaError_t e = cudaEventCreateWithFlags(&c->sync_event, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(c->sync_event, st);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(c->sync_event);
}

namespace scvod {
void* ctx_stream(scvod_ctx* c) { return (void*)c->stream; }
// the tracking stream joins the context's stream on entry (everything queued so far happens before the chain) and hands back on
// exit (everything queued afterwards sees the chain's results)
struct TrackStreamScope {
  scvod_ctx* c;
  explicit TrackStreamScope(scvod_ctx* ctx) : c(ctx) {
    if (!c->own_tstream) {
      c->tstream = c->stream;  // (scvod_set_stream may have replaced the context's stream since the last chain)
      return;
    }
    cudaEventRecord(c->tlink, c->stream);
    cudaStreamWaitEvent(c->tstream, c->tlink, 0);
  }
  ~TrackStreamScope() {
    if (!c->own_tstream) return;
    cudaEventRecord(c->tlink, c->tstream);
    cudaStreamWaitEvent(c->stream, c->tlink, 0);
  }
};

int ctx_device(const scvod_ctx* c) { return c->device; }
void ctx_add_launches(scvod_ctx* c, int n) { c->launches += n; }
void** ctx_gicp_slot(scvod_ctx* c) { return &c->gicp; }
void ctx_set_gicp_free(scvod_ctx* c, void (*fn)(void*)) { c->gicp_free = fn; }
int api_fail(int code, const std::string& msg) { return fail(code, msg); }
}  // namespace scvod

// scvod_params_semantickitti / scvod_params_parkinglot live in scvod_params.cpp (also linked into libscvod_synth.so)

extern "C" int scvod_grid_dims(const scvod_params* p, scvod_grid* g) {  // reference src/ssc.cpp:36-39 (float arithmetic)
  if (!p || !g) return fail(SCVOD_ERR_ARG, "null argument");
  g->range_num = (int)std::ceil((p->max_dis - p->min_dis) / p->range_res);
  g->sector_num = (int)std::ceil((p->max_angle - p->min_angle) / p->sector_res);
  g->azimuth_num = (int)std::ceil((p->max_azimuth - p->min_azimuth) / p->azimuth_res);
  g->bin_num = g->range_num * g->sector_num * g->azimuth_num;
  return SCVOD_OK;
}

extern "C" const char* scvod_last_error(void) { return g_err.c_str(); }

extern "C" void scvod_relative_pose(const float pose_next6[6], const float pose_pre6[6], float T[12]) {
  relative_pose(pose_next6, pose_pre6, T);
}
extern "C" void scvod_pose_matrix(const float pose6[6], float T[12]) { pose_matrix(pose6, T); }

static int alloc_workspace(scvod_ctx* c) {
  const size_t P = (size_t)c->max_points * c->max_batch;  // total points per batch
  const size_t S = (size_t)c->max_batch;
  BatchDev& w = c->ws;
  w.cap_points = (int)P;
  w.cap_scans = (int)S;
  w.edge_cap = 16384;
  w.hash_cap = 65536;
  CU(c->d_off.alloc(S + 1));
  CU(c->d_patch_of.alloc(P));
  CU(c->d_slot_patch.alloc(P));
  CU(c->d_patch_cnt.alloc(S * kNumPatches));
  CU(c->d_patch_off.alloc(S * (kNumPatches + 1)));
  CU(c->d_patch_cur.alloc(S * kNumPatches));
  CU(c->d_sort_ctr.alloc(8));
  CU(c->d_sort_list.alloc(3 * S * kNumPatches));
  CU(c->d_bucket_kv.alloc(P));
  CU(c->d_sorted_xyz.alloc(P));
  CU(c->d_sorted_idx.alloc(P));
  CU(c->d_slot_pos.alloc(P));
  CU(c->d_slot_apos.alloc(P));
  CU(c->d_slot_vid.alloc(P));
  CU(c->d_patch_out.alloc(S * kNumPatches * 8));
  CU(c->d_patch_out_off.alloc(S * (kNumPatches + 1) * 3));
  CU(c->d_patch_dbg.alloc(S * kNumPatches * 12));
  CU(c->d_scan_counts.alloc(S * 8 + 8));
  CU(c->d_apri_rank.alloc(P));
  CU(c->d_vox_cur.alloc(P));
  CU(c->d_vox_pts_tmp.alloc(P));
  CU(c->d_vox_nbr.alloc(P * 27));
  CU(c->d_vox_root.alloc(P));
  CU(c->d_vox_bbox.alloc(P * 6));
  CU(c->d_ev_cid.alloc(P));
  CU(c->d_edge_buf.alloc(S * w.edge_cap * 2));
  CU(c->d_edge_hash.alloc(S * w.hash_cap));
  CU(c->d_T.alloc(16));
  CU(c->d_counter.alloc(1));
  CU(c->d_taint_cnt.alloc(S * kTaintCntStride));
  CU(c->d_q_list.alloc(S * kQuirkCap));
  CU(c->d_vox_tnt.alloc(P));
  CU(c->d_vox_group.alloc(P));
  CU(c->d_tv_cid.alloc(S * kTvCap));
  CU(c->d_tv_base.alloc(S * (kTvCap + 1)));
  CU(c->d_tp_m.alloc(S * kTpCap));
  CU(c->d_tp_cid.alloc(S * kTpCap));
  CU(c->d_tp_name.alloc(S * kTpCap));
  CU(c->d_tp_xyz.alloc(S * kTpCap));
  CU(c->h_taint_cnt.alloc(S * kTaintCntStride));
  w.taint_cnt = c->d_taint_cnt.p;
  w.q_list = c->d_q_list.p;
  w.vox_tnt = c->d_vox_tnt.p;
  w.vox_group = c->d_vox_group.p;
  w.tv_cid = c->d_tv_cid.p;
  w.tv_base = c->d_tv_base.p;
  w.tp_m = c->d_tp_m.p;
  w.tp_cid = c->d_tp_cid.p;
  w.tp_name = c->d_tp_name.p;
  w.tp_xyz = c->d_tp_xyz.p;
  c->name_cap = c->max_points + 8;
  CU(c->d_vox_name.alloc(P));
  CU(c->d_name_first.alloc(S * (size_t)c->name_cap));
  w.off = c->d_off.p;
  w.patch_of = c->d_patch_of.p;
  w.slot_patch = c->d_slot_patch.p;
  w.patch_cnt = c->d_patch_cnt.p;
  w.patch_off = c->d_patch_off.p;
  w.patch_cur = c->d_patch_cur.p;
  w.sort_ctr = c->d_sort_ctr.p;
  w.sort_list = c->d_sort_list.p;
  w.bucket_kv = c->d_bucket_kv.p;
  w.sorted_xyz = c->d_sorted_xyz.p;
  w.sorted_idx = c->d_sorted_idx.p;
  w.slot_pos = c->d_slot_pos.p;
  w.slot_apos = c->d_slot_apos.p;
  w.slot_vid = c->d_slot_vid.p;
  w.patch_out = c->d_patch_out.p;
  w.patch_out_off = c->d_patch_out_off.p;
  w.patch_dbg = c->d_patch_dbg.p;
  w.scan_counts = c->d_scan_counts.p;
  w.apri_rank = c->d_apri_rank.p;
  w.vox_cur = c->d_vox_cur.p;
  w.vox_pts_tmp = c->d_vox_pts_tmp.p;
  w.vox_nbr = c->d_vox_nbr.p;
  w.vox_root = c->d_vox_root.p;
  w.vox_bbox = c->d_vox_bbox.p;
  w.ev_cid = c->d_ev_cid.p;
  w.edge_buf = c->d_edge_buf.p;
  w.edge_hash = reinterpret_cast<int32_t*>(c->d_edge_hash.p);
  CU(c->h_scan_counts.alloc(S * 8 + 8));
  return SCVOD_OK;
}

extern "C" int scvod_create(const scvod_params* p, int device, int max_points, int max_batch, scvod_ctx** out) {
  if (!p || !out || max_points <= 0 || max_batch <= 0) return fail(SCVOD_ERR_ARG, "bad arguments to scvod_create");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return fail(SCVOD_ERR_CUDA, "no CUDA device: the SCV-OD path has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(SCVOD_ERR_ARG, "device index out of range");
  CU(cudaSetDevice(device));
  std::unique_ptr<scvod_ctx> c(new scvod_ctx());
  c->hp.p = *p;
  scvod_grid g;
  scvod_grid_dims(p, &g);
  c->hp.g.range_num = g.range_num;
  c->hp.g.sector_num = g.sector_num;
  c->hp.g.azimuth_num = g.azimuth_num;
  c->hp.g.bin_num = g.bin_num;
  c->hp.g.key_off = g.range_num * g.sector_num + g.sector_num + 1;  // all three indices == -1 (ssc.cpp:185-188)
  c->hp.g.key_count = g.bin_num + c->hp.g.key_off;
  c->hp.g.words = (c->hp.g.key_count + 31) / 32;
  c->device = device;
  c->max_points = max_points;
  c->max_batch = max_batch;
  unsigned hc = std::thread::hardware_concurrency();
  c->host_threads = hc ? (int)std::min<unsigned>(hc, 32u) : 4;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;  // numerically lower = higher priority
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // measured on one B200 with 16 contexts: a high-priority tracking stream shortens every pair (112 -> ~60 us under load, host
    // polling 10 -> 6.5 busy cores) but does not raise the throughput, which is bound by the aggregate kernel work (35.6k flat vs
    // 33.5k scans/s prioritised; 25.2k vs 24.8k on 4 cores): opt-in
    static const bool flat = !(getenv("SCVOD_TRACK_PRIORITY") && atoi(getenv("SCVOD_TRACK_PRIORITY")) != 0);
    // Without the priority the chain stays on the context's own stream: a second stream per context would double the number of
    // streams of the process, and beyond CUDA_DEVICE_MAX_CONNECTIONS (32 at most) streams share hardware queues, where a k_track
    // launch can end up behind another context's bulk upload (seen with 24 contexts: 27.4k -> 12.3k scans/s end to end).
    if (!flat) {
      CU(cudaStreamCreateWithPriority(&c->tstream, cudaStreamNonBlocking, hi));
      CU(cudaEventCreateWithFlags(&c->tlink, cudaEventDisableTiming));
      c->own_tstream = true;
    }
    (void)lo;
  }
  int rc = alloc_workspace(c.get());
  if (rc != SCVOD_OK) return rc;
  g_live_contexts.fetch_add(1);
  *out = c.release();
  return SCVOD_OK;
}

extern "C" int scvod_destroy(scvod_ctx* c) {
  if (!c) return SCVOD_OK;
  if (g_live_contexts.fetch_sub(1) == 1) g_prof.dump();  // the last context of the process prints the host-side profile
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->gicp && c->gicp_free) c->gicp_free(c->gicp);
  c->gicp = nullptr;
  for (auto& b : c->batches)
    if (b) b->release();
  for (auto& b : c->batch_pool)
    if (b) b->release();
  c->d_off.release(); c->d_patch_of.release(); c->d_slot_patch.release(); c->d_patch_cnt.release(); c->d_patch_off.release();
  c->d_patch_cur.release(); c->d_sort_ctr.release(); c->d_sort_list.release(); c->d_sorted_idx.release(); c->d_slot_pos.release(); c->d_slot_apos.release(); c->d_slot_vid.release();
  c->d_patch_out.release(); c->d_patch_out_off.release(); c->d_scan_counts.release(); c->d_apri_rank.release(); c->d_vox_cur.release();
  c->d_vox_pts_tmp.release(); c->d_vox_nbr.release(); c->d_vox_root.release(); c->d_ev_cid.release(); c->d_edge_buf.release();
  c->d_bucket_kv.release(); c->d_sorted_xyz.release(); c->d_edge_hash.release(); c->d_patch_dbg.release(); c->d_vox_bbox.release(); c->d_T.release();
  c->h_scan_counts.release(); c->h_vox_cnt.release(); c->h_vox_root.release(); c->h_vox_nbr.release(); c->h_ev_cid.release();
  c->h_edge_buf.release(); c->h_vox_bbox.release(); c->d_first.release(); c->d_triples.release();
  c->h_treq.release(); c->h_triples.release(); c->d_treq.release(); c->d_track_ctr.release(); c->d_track_list.release(); c->d_tout[0].release(); c->d_tout[1].release(); c->d_vcls.release(); c->h_vcls.release(); c->d_lov.release(); c->h_lov.release();
  c->d_Ts.release(); c->h_Ts.release();
  c->d_counter.release();
  c->d_taint_cnt.release(); c->d_q_list.release(); c->d_vox_tnt.release(); c->d_vox_group.release(); c->d_tv_cid.release(); c->d_tv_base.release();
  c->d_tp_m.release(); c->d_tp_cid.release(); c->d_tp_name.release(); c->d_tp_xyz.release(); c->h_taint_cnt.release();
  c->h_pack.release(); c->d_pack.release(); c->h_desc.release(); c->d_desc.release(); c->d_vox_name.release(); c->d_name_first.release();
  if (c->own_stream) cudaStreamDestroy(c->stream);
  if (c->own_tstream && c->tstream) {
    cudaStreamSynchronize(c->tstream);
    cudaStreamDestroy(c->tstream);
  }
  if (c->tlink) cudaEventDestroy(c->tlink);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  if (c->prefetch_done) cudaEventDestroy(c->prefetch_done);
  if (c->sync_event) cudaEventDestroy(c->sync_event);
  c->d_prefetch.release();
  delete c;
  return SCVOD_OK;
}

extern "C" int scvod_num_kernel_launches(const scvod_ctx* c, int64_t* out) {
  if (!c || !out) return fail(SCVOD_ERR_ARG, "null argument");
  *out = c->launches;
  return SCVOD_OK;
}

extern "C" int scvod_get_stat(scvod_ctx* c, const char* key, int64_t* out) {
  if (!c || !key || !out) return fail(SCVOD_ERR_ARG, "null argument");
  std::string k(key);
  if (k == "track_points") *out = c->stat_track_points;
  else if (k == "track_pairs") *out = c->stat_track_pairs;
  else if (k == "scans") *out = c->stat_scans;
  else if (k == "points") *out = c->stat_points;
  else if (k == "apri_points") *out = c->stat_apri;
  else if (k == "voxels") *out = c->stat_voxels;
  else if (k == "tainted_voxels") *out = c->stat_tvox;
  else if (k == "tainted_points") *out = c->stat_tpts;
  else if (k == "tracked_frames") *out = c->tracked;  // frames 0 .. tracked-1 have been tracked as frame_pre_ (the chain resumes at `tracked`)
  else if (k == "reallocs") *out = g_reallocs.load();  // process-wide: device / pinned buffer (re)allocations so far
  else return fail(SCVOD_ERR_ARG, "unknown stat " + k);
  return SCVOD_OK;
}

extern "C" int scvod_set_option(scvod_ctx* c, const char* key, int value) {
  if (!c || !key) return fail(SCVOD_ERR_ARG, "null argument");
  std::string k(key);
  if (k == "inspect")
    c->inspect = value != 0;
  else if (k == "host_threads")
    c->host_threads = std::max(1, value);
  else if (k == "replay_global")
    c->replay_global = value != 0;
  else if (k == "chain_tma")
    c->hp.chain_tma = value != 0;
  else if (k == "submap_all_static")
    c->submap_all_static = value != 0;
  else if (k == "profile_reset")  // SCVOD_PROFILE=1: forget what the host-side section timers have accumulated (e.g. after a warm-up)
    g_prof.reset();
  else
    return fail(SCVOD_ERR_ARG, "unknown option " + k);
  return SCVOD_OK;
}

extern "C" int scvod_set_stream(scvod_ctx* c, void* cuda_stream) {
  if (!c) return fail(SCVOD_ERR_ARG, "null ctx");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)cuda_stream;
  c->own_stream = false;
  return SCVOD_OK;
}

extern "C" int scvod_kernel_timing(int enable) {
  timing_enable(enable != 0);
  if (enable) timing_reset();
  return SCVOD_OK;
}

extern "C" int scvod_kernel_timing_report(char* buf, int cap) {
  if (!buf || cap <= 0) return fail(SCVOD_ERR_ARG, "bad buffer");
  std::string r = timing_report();
  if ((int)r.size() + 1 > cap) return fail(SCVOD_ERR_CAPACITY, "report buffer too small");
  std::memcpy(buf, r.c_str(), r.size() + 1);
  return (int)r.size();
}

// ---------------------------------------------------------------------------------------------
// batched frame pipeline
// ---------------------------------------------------------------------------------------------
static int push_batch(scvod_ctx* c, const void* xyzi, bool on_device, const int64_t* offsets, int nscans) {
  cudaStream_t st = c->stream;
  const int64_t total = offsets[nscans] - offsets[0];
  int max_n = 0;
  std::vector<int64_t> off(nscans + 1);
  for (int s = 0; s <= nscans; ++s) off[s] = offsets[s] - offsets[0];
  for (int s = 0; s < nscans; ++s) {
    int64_t n = off[s + 1] - off[s];
    if (n < 0 || n > c->max_points) {
      c->failed_scan = c->batch_first_scan + s;
      return fail(SCVOD_ERR_CAPACITY, "scan " + std::to_string(c->failed_scan) + " larger than max_points");
    }
    max_n = std::max<int>(max_n, (int)n);
  }
  if (total > (int64_t)c->ws.cap_points) return fail(SCVOD_ERR_CAPACITY, "batch larger than workspace");

  PROF("push_batch total");
  std::unique_ptr<PersistBatch> pb;
  if (!c->batch_pool.empty()) {
    pb = std::move(c->batch_pool.back());
    c->batch_pool.pop_back();
  } else {
    pb.reset(new PersistBatch());
  }
  pb->nscans = nscans;
  pb->total = total;
  pb->off = off;
  pb->max_n = max_n;
  pb->labels_current = false;
  const size_t T = (size_t)std::max<int64_t>(total, 1);
  CU(pb->pts.alloc(T));
  CU(pb->apri_xyzi.alloc(T));
  CU(pb->cls.alloc(T));
  CU(pb->apri_src.alloc(T));
  CU(pb->apri_cid.alloc(T));
  CU(pb->apri_vid.alloc(T));
  CU(pb->vox_off.alloc(T));
  CU(pb->vox_pts.alloc(T));
  CU(pb->vox_cnt.alloc(T));
  CU(pb->ground_src.alloc(T));
  CU(pb->ng_src.alloc(T));
  CU(pb->bitmap.alloc((size_t)nscans * c->hp.g.words));
  CU(pb->word_rank.alloc((size_t)nscans * c->hp.g.words));
  CU(pb->vox_vid.alloc(T));
  CU(pb->vox_tri.alloc(3 * T));
  CU(pb->vox_av.alloc(T));
  CU(pb->vox_cov.alloc(T));
  CU(pb->vox_center.alloc(3 * T));
  CU(pb->off_dev.alloc(nscans + 1));
  CU(pb->scan_counts_dev.alloc((size_t)nscans * 8));

  BatchDev& w = c->ws;
  w.pts = pb->pts.p;
  w.apri_xyzi = pb->apri_xyzi.p;
  w.cls = pb->cls.p;
  w.apri_src = pb->apri_src.p;
  w.apri_cid = pb->apri_cid.p;
  w.apri_vid = pb->apri_vid.p;
  w.vox_off = pb->vox_off.p;
  w.vox_pts = pb->vox_pts.p;
  w.vox_cnt = pb->vox_cnt.p;
  w.ground_src = pb->ground_src.p;
  w.ng_src = pb->ng_src.p;
  w.bitmap = pb->bitmap.p;
  w.word_rank = pb->word_rank.p;
  w.vox_vid = pb->vox_vid.p;
  w.vox_tri = pb->vox_tri.p;
  w.vox_av = pb->vox_av.p;
  w.vox_cov = pb->vox_cov.p;
  w.vox_center = pb->vox_center.p;

  std::chrono::steady_clock::time_point tk0 = std::chrono::steady_clock::now();
  const double tk0_cpu = g_prof.on ? thread_cpu_ms() : 0;
  CU(cudaMemcpyAsync(w.off, off.data(), sizeof(int64_t) * (nscans + 1), cudaMemcpyHostToDevice, st));
  if (total > 0) {
    const char* src = (const char*)xyzi + sizeof(float) * 4 * offsets[0];
    const size_t bytes = sizeof(float4) * (size_t)total;
    if (!on_device && c->prefetch_host && src >= c->prefetch_host && src + bytes <= c->prefetch_host + c->prefetch_bytes) {
      // already on its way (scvod_prefetch_scans): wait for the private copy stream, then a device-to-device move
      CU(cudaStreamWaitEvent(st, c->prefetch_done, 0));
      CU(cudaMemcpyAsync(w.pts, (const char*)c->d_prefetch.p + (src - c->prefetch_host), bytes, cudaMemcpyDeviceToDevice, st));
      if (src + bytes == c->prefetch_host + c->prefetch_bytes) c->prefetch_host = nullptr;  // consumed up to its end
    } else {
      CU(cudaMemcpyAsync(w.pts, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    }
  }
  CU(cudaMemsetAsync(w.scan_counts, 0, sizeof(int32_t) * ((size_t)w.cap_scans * 8 + 8), st));
  c->launches += launch_ground(c->hp, w, nscans, max_n, st);
  c->launches += launch_descriptor(c->hp, w, nscans, max_n, st);
  c->launches += launch_cluster_prep(c->hp, w, nscans, max_n, st);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_scan_counts.p, w.scan_counts, sizeof(int32_t) * ((size_t)w.cap_scans * 8 + 8), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_taint_cnt.p, w.taint_cnt, sizeof(int32_t) * (size_t)nscans * kTaintCntStride, cudaMemcpyDeviceToHost, st));
  CU(wait_stream(c, st));
  if (g_prof.on) g_prof.add("  h2d + kernels + sync", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tk0).count(), thread_cpu_ms() - tk0_cpu);
  const int32_t* sc = c->h_scan_counts.p;
  if (sc[(size_t)w.cap_scans * 8] & 1) return fail(SCVOD_ERR_CAPACITY, "a PatchWork patch holds more points than the largest fit tile");

  // per-scan table sizes -> packed host copies
  CU(cudaMemcpyAsync(pb->off_dev.p, w.off, sizeof(int64_t) * (nscans + 1), cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(pb->scan_counts_dev.p, w.scan_counts, sizeof(int32_t) * (size_t)nscans * 8, cudaMemcpyDeviceToDevice, st));
  std::vector<int64_t> vbase(nscans + 1, 0), ebase(nscans + 1, 0), gbase(nscans + 1, 0), mbase(nscans + 1, 0), obase(nscans + 1, 0);
  // tainted voxels (points with a -1 index, ssc.cpp:185-188): tv = voxels, tp = their points
  const int32_t* tc = c->h_taint_cnt.p;
  std::vector<int64_t> tvb(nscans + 1, 0), tpb(nscans + 1, 0);
  bool any_taint = false;
  for (int s = 0; s < nscans; ++s) {
    int V = sc[s * 8 + 3], E = sc[s * 8 + 5], G = sc[s * 8 + 7];
    // capacity limits are per scan: name the scan, so that the caller can drop or re-voxelise it and push the others again
    // (the frames of this batch are not kept; frames of earlier calls and of earlier batches of this call are)
    const std::string which = "scan " + std::to_string(c->batch_first_scan + s) + " of the call: ";
    if (G < 0 || G > w.edge_cap) {
      c->failed_scan = c->batch_first_scan + s;
      return fail(SCVOD_ERR_CAPACITY, which + "similarity edge table overflow");
    }
    if (tc[s * kTaintCntStride + 3]) {
      c->failed_scan = c->batch_first_scan + s;
      return fail(SCVOD_ERR_CAPACITY, which + "more points with a -1 curved-voxel index (or voxels / points aliased by them) than the side tables take");
    }
    const int ntv = tc[s * kTaintCntStride + 1], ntp = tc[s * kTaintCntStride + 2];
    if ((int64_t)V + ntp > off[s + 1] - off[s]) {
      c->failed_scan = c->batch_first_scan + s;
      return fail(SCVOD_ERR_CAPACITY, which + "too small for the replay scratch of its aliased voxels");
    }
    any_taint = any_taint || ntp > 0;
    tvb[s + 1] = tvb[s] + ntv;
    tpb[s + 1] = tpb[s] + ntp;
    vbase[s + 1] = vbase[s] + V;
    ebase[s + 1] = ebase[s] + E;
    gbase[s + 1] = gbase[s] + G;
    mbase[s + 1] = mbase[s] + sc[s * 8 + 2];
    obase[s + 1] = obase[s] + V + 1;
  }

  std::chrono::steady_clock::time_point td0 = std::chrono::steady_clock::now();
  const double td0_cpu = g_prof.on ? thread_cpu_ms() : 0;
  // cluster names: one warp per scan replays the reference's sequential naming on the device
  int max_vox = 1, max_ev = 1;
  for (int s = 0; s < nscans; ++s) {
    max_vox = std::max(max_vox, sc[s * 8 + 3] + tc[s * kTaintCntStride + 2]);  // union-find nodes: voxels + points of tainted voxels
    max_ev = std::max(max_ev, sc[s * 8 + 5]);
  }
  c->launches += launch_name_replay(c->hp, w, nscans, max_vox, max_ev, c->replay_global, any_taint, c->d_vox_name.p, c->d_name_first.p, c->name_cap, st);
  CU(cudaGetLastError());
  // one packed gather + one D2H for all per-scan tables: [cnt Vt][root Vt][name Vt][bbox 6Vt][name_first Nt][edges 2Gt][max_name S]
  const int64_t Vt = vbase[nscans], Gt = gbase[nscans];
  std::vector<int64_t> nbase(nscans + 1, 0);
  for (int s = 0; s < nscans; ++s)
    nbase[s + 1] = nbase[s] + std::min<int64_t>(c->name_cap, (int64_t)sc[s * 8 + 3] + tc[s * kTaintCntStride + 2] + 6);
  const int64_t Nt = nbase[nscans];
  const int64_t TVt = tvb[nscans], TPt = tpb[nscans];
  const int64_t o_cnt = 0, o_root = Vt, o_name = 2 * Vt, o_bbox = 3 * Vt, o_nf = 9 * Vt, o_edge = 9 * Vt + Nt, o_max = 9 * Vt + Nt + 2 * Gt;
  // [tv_cid TVt][tv_base TVt + S][tp_m TPt][tp_name TPt][tp_xyz 4 TPt] after the counters (empty without tainted voxels)
  const int64_t o_tvc = o_max + (int64_t)nscans * 8, o_tvb = o_tvc + TVt, o_tpm = o_tvb + TVt + nscans, o_tpn = o_tpm + TPt, o_tpx = o_tpn + TPt;
  const int64_t pack_ints = std::max<int64_t>(1, o_tpx + 4 * TPt);
  CU(c->h_pack.alloc(pack_ints));
  CU(c->d_pack.alloc(pack_ints));
  CU(c->h_desc.alloc((size_t)nscans * 11 + 1));
  CU(c->d_desc.alloc((size_t)nscans * 11 + 1));
  int nd = 0, max_desc_n = 1;
  auto add = [&](const void* src, int64_t dst, int n) {
    if (n <= 0) return;
    PackDesc d;
    d.src = reinterpret_cast<const int32_t*>(src);
    d.dst = dst;
    d.n = n;
    d.pad = 0;
    c->h_desc.p[nd++] = d;
    max_desc_n = std::max(max_desc_n, n);
  };
  for (int s = 0; s < nscans; ++s) {
    int V = sc[s * 8 + 3], G = sc[s * 8 + 7];
    int64_t b = off[s];
    add(w.vox_cnt + b, o_cnt + vbase[s], V);
    add(w.vox_root + b, o_root + vbase[s], V);
    add(c->d_vox_name.p + b, o_name + vbase[s], V);
    add(w.vox_bbox + b * 6, o_bbox + vbase[s] * 6, V * 6);
    add(c->d_name_first.p + (size_t)s * c->name_cap, o_nf + nbase[s], (int)(nbase[s + 1] - nbase[s]));
    add(w.edge_buf + (size_t)s * w.edge_cap * 2, o_edge + gbase[s] * 2, G * 2);
    const int ntv = (int)(tvb[s + 1] - tvb[s]), ntp = (int)(tpb[s + 1] - tpb[s]);
    if (ntp > 0) {
      add(w.tv_cid + (size_t)s * kTvCap, o_tvc + tvb[s], ntv);
      add(w.tv_base + (size_t)s * (kTvCap + 1), o_tvb + tvb[s] + s, ntv + 1);
      add(w.tp_m + (size_t)s * kTpCap, o_tpm + tpb[s], ntp);
      add(w.tp_name + (size_t)s * kTpCap, o_tpn + tpb[s], ntp);
      add(w.tp_xyz + (size_t)s * kTpCap, o_tpx + 4 * tpb[s], 4 * ntp);
    }
  }
  add(w.scan_counts, o_max, nscans * 8);  // re-read the counters: slot 6 now holds max_name
  CU(cudaMemcpyAsync(c->d_desc.p, c->h_desc.p, sizeof(PackDesc) * nd, cudaMemcpyHostToDevice, st));
  c->launches += launch_pack(c->d_desc.p, nd, max_desc_n, c->d_pack.p, st);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_pack.p, c->d_pack.p, sizeof(int32_t) * pack_ints, cudaMemcpyDeviceToHost, st));
  CU(wait_stream(c, st));
  const int32_t* hp_cnt = c->h_pack.p + o_cnt;
  const int32_t* hp_root = c->h_pack.p + o_root;
  const int32_t* hp_name = c->h_pack.p + o_name;
  const float* hp_bbox = reinterpret_cast<const float*>(c->h_pack.p + o_bbox);
  const int32_t* hp_nf = c->h_pack.p + o_nf;
  const int32_t* hp_edge = c->h_pack.p + o_edge;
  const int32_t* hp_sc2 = c->h_pack.p + o_max;
  if (g_prof.on) g_prof.add("  d2h voxel tables", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - td0).count(), thread_cpu_ms() - td0_cpu);
  std::chrono::steady_clock::time_point th0 = std::chrono::steady_clock::now();
  const double th0_cpu = g_prof.on ? thread_cpu_ms() : 0;
  // host cluster bookkeeping, one scan per task
  const int batch_id = (int)c->batches.size();
  const size_t f0 = c->frames.size();
  c->frames.resize(f0 + nscans);
  struct FramesGuard {  // an error below leaves no frame that points at a batch that was never committed
    scvod_ctx* c;
    size_t f0;
    bool ok = false;
    ~FramesGuard() {
      if (!ok) c->frames.resize(f0);
    }
  } guard{c, f0};
  std::atomic<int> next(0);
  std::atomic<int> bad(0);
  auto worker = [&]() {
    for (;;) {
      int s = next.fetch_add(1);
      if (s >= nscans) break;
      FrameHost& fr = c->frames[f0 + s];
      fr.batch = batch_id;
      fr.slot = s;
      fr.base = off[s];
      fr.n_in = (int)(off[s + 1] - off[s]);
      fr.n_ground = sc[s * 8 + 0];
      fr.n_ng = sc[s * 8 + 1];
      fr.n_apri = sc[s * 8 + 2];
      fr.n_vox = sc[s * 8 + 3];
      ScanTables t;
      t.M = fr.n_apri;
      t.V = fr.n_vox;
      t.n_events = sc[s * 8 + 5];
      t.n_edges = sc[s * 8 + 7];
      t.vox_cnt = hp_cnt + vbase[s];
      t.vox_root = hp_root + vbase[s];
      t.vox_bbox = hp_bbox + vbase[s] * 6;
      t.edges = hp_edge + gbase[s] * 2;
      t.vox_name = hp_name + vbase[s];
      t.name_first = hp_nf + nbase[s];
      t.max_name = hp_sc2[s * 8 + 6];
      t.n_tvox = (int)(tvb[s + 1] - tvb[s]);
      t.n_tpts = (int)(tpb[s + 1] - tpb[s]);
      if (t.n_tpts > 0) {
        t.tv_cid = c->h_pack.p + o_tvc + tvb[s];
        t.tv_base = c->h_pack.p + o_tvb + tvb[s] + s;
        t.tp_m = c->h_pack.p + o_tpm + tpb[s];
        t.tp_name = c->h_pack.p + o_tpn + tpb[s];
        t.tp_xyz = reinterpret_cast<const float*>(c->h_pack.p + o_tpx + 4 * tpb[s]);
      }
      fr.vox_cnt.assign(t.vox_cnt, t.vox_cnt + t.V);
      if (t.max_name + 1 > (int)(nbase[s + 1] - nbase[s])) {  // more names than the table holds: name_first would be read past its slice
        bad.store(1);
        continue;
      }
      if (!segment_and_recognize(c->hp.p, t, fr.fc, c->inspect)) bad.store(1);
    }
  };
  int nt = std::min(c->host_threads, nscans);
  if (nt <= 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < nt; ++i) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  if (g_prof.on) g_prof.add("  host segment+recognize", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - th0).count(), thread_cpu_ms() - th0_cpu);
  // car CSR of the batch (see PersistBatch::csr): the own voxels of every car cluster, in occupy_voxels order, with running
  // point offsets and part indices; one upload per batch, so that tracking needs no per-pair segment upload.  A tainted voxel
  // contributes only the subgroups the cluster owns, as "virtual voxels" whose points sit in the 4th section (tv_pts).
  if (bad.load()) {  // nothing of this batch stays queryable (FramesGuard)
    c->batch_pool.push_back(std::move(pb));
    return fail(SCVOD_ERR_STATE, "internal: replayed cluster partition differs from the GPU components");
  }
  {
    PROF("  car csr build");
    size_t csr_n = 0, tv_n = 0;
    for (int s = 0; s < nscans; ++s) {
      FrameHost& fr = c->frames[f0 + s];
      for (auto& cs : fr.fc.cluster_set)
        if (cs.second.type == c->hp.p.car) csr_n += cs.second.occupy_voxels.size() + cs.second.tunits.size();
      fr.sg_tv_off.assign(fr.fc.subgroups.size(), 0);
      for (size_t g = 0; g < fr.fc.subgroups.size(); ++g) {
        fr.sg_tv_off[g] = (int)tv_n;
        tv_n += fr.fc.subgroups[g].pts.size();
      }
    }
    csr_n += (size_t)nscans;  // one closing point offset per frame
    CU(pb->h_csr.alloc(3 * csr_n + tv_n));
    CU(pb->csr.alloc(3 * csr_n + tv_n));
    pb->csr_n = (int)csr_n;
    pb->tv_n = (int)tv_n;
    int32_t* ptoff = pb->h_csr.p;
    int32_t* cvox = pb->h_csr.p + csr_n;
    int32_t* cpart = pb->h_csr.p + 2 * csr_n;
    int32_t* tvp = pb->h_csr.p + 3 * csr_n;
    size_t pos = 0;
    for (int s = 0; s < nscans; ++s) {
      FrameHost& fr = c->frames[f0 + s];
      for (size_t g = 0; g < fr.fc.subgroups.size(); ++g)
        std::copy(fr.fc.subgroups[g].pts.begin(), fr.fc.subgroups[g].pts.end(), tvp + fr.sg_tv_off[g]);
      const bool taint = !fr.fc.tvox.empty();
      fr.csr_base = (int)pos;
      int run_pts = 0;
      for (auto& cs : fr.fc.cluster_set) {
        HCluster& cl = cs.second;
        cl.own_runs.clear();
        if (cl.type != c->hp.p.car) continue;
        HCluster::OwnRun orun;
        orun.csr_start = (int)(pos - fr.csr_base);
        orun.part_base = 0;
        const int pts0 = run_pts;
        int part = 0, vi = 0;
        for (int v : cl.occupy_voxels) {
          while (part < (int)cl.part_end.size() && vi >= cl.part_end[part]) ++part;
          ++vi;
          if (taint && fr.fc.tainted(v)) continue;
          ptoff[pos] = run_pts;
          cvox[pos] = v;
          cpart[pos] = part;
          run_pts += std::max(0, fr.vox_cnt[v]);
          ++pos;
        }
        for (auto& tu : cl.tunits) {
          ptoff[pos] = run_pts;
          cvox[pos] = kVirtualVox | fr.sg_tv_off[tu.sg];
          cpart[pos] = tu.part;
          run_pts += (int)fr.fc.subgroups[tu.sg].pts.size();
          ++pos;
        }
        orun.csr_len = (int)(pos - fr.csr_base) - orun.csr_start;
        orun.npts = run_pts - pts0;
        cl.own_runs.push_back(orun);
      }
      ptoff[pos] = run_pts;  // closing offset (the voxel / part slots of this position stay unused)
      cvox[pos] = 0;
      cpart[pos] = 0;
      ++pos;
    }
    CU(cudaMemcpyAsync(pb->csr.p, pb->h_csr.p, sizeof(int32_t) * (3 * csr_n + tv_n), cudaMemcpyHostToDevice, st));
  }
  c->stat_scans += nscans;
  c->stat_points += total;
  c->stat_apri += mbase[nscans];
  c->stat_voxels += vbase[nscans];
  c->stat_tvox += TVt;
  c->stat_tpts += TPt;
  c->batches.push_back(std::move(pb));
  guard.ok = true;
  return SCVOD_OK;
}

static int push_scans_impl(scvod_ctx* c, const void* xyzi, bool on_device, const int64_t* offsets, int nscans) {
  if (!c || !offsets || nscans < 0 || (!xyzi && nscans > 0 && offsets[nscans] > offsets[0]))
    return fail(SCVOD_ERR_ARG, "bad arguments to scvod_push_scans");
  CU(cudaSetDevice(c->device));
  c->failed_scan = -1;
  int s = 0;
  while (s < nscans) {
    int e = s;
    int64_t pts = 0;
    while (e < nscans && (e - s) < c->max_batch && pts + (offsets[e + 1] - offsets[e]) <= (int64_t)c->ws.cap_points) {
      pts += offsets[e + 1] - offsets[e];
      ++e;
    }
    if (e == s) {
      c->failed_scan = s;
      return fail(SCVOD_ERR_CAPACITY, "scan " + std::to_string(s) + " larger than the batch workspace");
    }
    c->batch_first_scan = s;
    int rc = push_batch(c, xyzi, on_device, offsets + s, e - s);
    if (rc != SCVOD_OK) return rc;
    s = e;
  }
  return SCVOD_OK;
}

extern "C" int scvod_prefetch_scans(scvod_ctx* c, const float* xyzi, const int64_t* offsets, int nscans) {
  if (!c || !offsets || nscans < 0 || (!xyzi && nscans > 0 && offsets[nscans] > offsets[0]))
    return fail(SCVOD_ERR_ARG, "bad arguments to scvod_prefetch_scans");
  CU(cudaSetDevice(c->device));
  const int64_t total = offsets[nscans] - offsets[0];
  if (total <= 0) return SCVOD_OK;
  if (!c->copy_stream) {
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->prefetch_done, cudaEventDisableTiming));
  }
  // the staging buffer may still feed the device-to-device move of the previous push: that move is on the compute stream
  CU(cudaEventRecord(c->prefetch_done, c->stream));
  CU(cudaStreamWaitEvent(c->copy_stream, c->prefetch_done, 0));
  if ((size_t)total > c->d_prefetch.n) {
    CU(cudaStreamSynchronize(c->stream));
    CU(c->d_prefetch.alloc((size_t)total));
  }
  const char* src = (const char*)xyzi + sizeof(float) * 4 * offsets[0];
  CU(cudaMemcpyAsync(c->d_prefetch.p, src, sizeof(float4) * (size_t)total, cudaMemcpyHostToDevice, c->copy_stream));
  CU(cudaEventRecord(c->prefetch_done, c->copy_stream));
  c->prefetch_host = src;
  c->prefetch_bytes = sizeof(float4) * (size_t)total;
  return SCVOD_OK;
}

extern "C" int scvod_push_scans(scvod_ctx* c, const float* xyzi, const int64_t* offsets, int nscans) {
  return push_scans_impl(c, xyzi, false, offsets, nscans);
}
extern "C" int scvod_push_scans_dev(scvod_ctx* c, const void* xyzi_dev, const int64_t* offsets, int nscans) {
  return push_scans_impl(c, xyzi_dev, true, offsets, nscans);
}

extern "C" int scvod_last_failed_scan(const scvod_ctx* c) { return c ? c->failed_scan : -1; }

extern "C" int scvod_num_frames(const scvod_ctx* c) { return c ? (int)c->frames.size() : 0; }

extern "C" int scvod_reset_frames(scvod_ctx* c) {
  if (!c) return fail(SCVOD_ERR_ARG, "null ctx");
  PROF("reset_frames total");
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& b : c->batches)
    if (b) c->batch_pool.push_back(std::move(b));
  c->batches.clear();
  c->frames.clear();
  c->tracked = 0;
  c->head_tracked = false;
  c->track_name = 0;
  c->tout_cur = 0;
  c->have_init = false;
  c->init_base = -1;
  return SCVOD_OK;
}

// ---------------------------------------------------------------------------------------------
// tracking (SSC::tracking, reference src/ssc.cpp:1250-1426)
// ---------------------------------------------------------------------------------------------
// Re-binning of the clouds of `cars` (clusters of frame `pre`) into the curved-voxel grid of frame `next` after the rigid
// transform T (transformCloud + ssc.cpp:1275-1315 / :1180-1205): one k_track launch for all clusters.  On return the distinct
// (cluster, hit voxel) pairs with their first-occurrence keys are grouped by cluster in c->hit_start / hit_key / hit_vox,
// cstart holds the offset of every cluster's transformed cloud in c->d_tout[1 - c->tout_cur].
static int diff_clusters(scvod_ctx* c, FrameHost& pre, FrameHost& next, const float T[12], const std::vector<HCluster*>& cars,
                         std::vector<size_t>& cstart, bool allow_runs) {
  const int ncl = (int)cars.size();
  const int vn = next.n_vox;
  cstart.assign(cars.size() + 1, 0);
  // per cluster: (first-occurrence key, hit voxel), grouped by cluster in c->hit_start / c->hit_key / c->hit_vox.  The cloud
  // of a cluster is [part 0][part 1]...[carried 0]... (ssc.cpp:380,617,1382) with ascending apri index inside a part,
  // which is what the 64-bit key encodes.
  c->hit_start.assign(cars.size() + 1, 0);
  const int in_buf = c->tout_cur, out_buf = 1 - c->tout_cur;
  PROF("  track: gpu round trip");
  PersistBatch& pbp = *c->batches[pre.batch];
  PersistBatch& pbn = *c->batches[next.batch];
  // ---- which points: a run table in the kernel arguments (car clusters described by runs of the frame's device CSR) ... ----
  size_t K = 0;
  size_t nruns = 0;
  bool use_runs = allow_runs && pre.csr_base >= 0;
  if (use_runs) {
    for (HCluster* cl : cars) {
      if (cl->own_runs.empty() && !cl->occupy_voxels.empty()) use_runs = false;
      nruns += cl->own_runs.size() + cl->carried.size();
    }
    if (nruns > (size_t)kTrackMaxRuns) use_runs = false;
  }
  TrackRuns& runs = c->track_runs;
  size_t si = 0;  // segments (fallback path)
  if (use_runs) {
    int r = 0;
    for (size_t i = 0; i < cars.size(); ++i) {
      cstart[i] = K;
      HCluster& cl = *cars[i];
      for (auto& orun : cl.own_runs) {
        if (orun.npts <= 0) continue;
        runs.dst_off[r] = (int)K;
        runs.src[r] = pre.csr_base + orun.csr_start;
        runs.len[r] = orun.csr_len;
        runs.order[r] = orun.part_base;
        runs.cluster[r] = (int)i;
        ++r;
        K += (size_t)orun.npts;
      }
      int ord = 0x40000000;
      for (auto& cr : cl.carried) {
        if (cr.second <= 0) continue;
        runs.dst_off[r] = (int)K;
        runs.src[r] = -1 - cr.first;
        runs.len[r] = cr.second;
        runs.order[r] = ord++;
        runs.cluster[r] = (int)i;
        ++r;
        K += (size_t)cr.second;
      }
    }
    runs.n = r;
    cstart[cars.size()] = K;
    if (r == 0) K = 0;
  } else {
    // ---- ... or one uploaded segment per voxel (clusters without runs: scvod_initialization; more runs than fit) ----
    size_t n_seg = 0;
    for (HCluster* cl : cars) n_seg += cl->occupy_voxels.size() + cl->carried.size() + cl->tunits.size();
    const bool taint = !pre.fc.tvox.empty();
    size_t k_bound = 0;  // upper bound of the number of points (the per-block index follows the segment table)
    for (HCluster* cl : cars) k_bound += (size_t)std::max(0, cl->npts) + (size_t)std::max(0, cl->n_carried);
    CU(c->h_treq.alloc(std::max<size_t>(4, n_seg * 4) + k_bound / 256 + 2));
    int32_t* seg = c->h_treq.p;
    size_t k = 0;
    for (size_t i = 0; i < cars.size(); ++i) {
      cstart[i] = k;
      HCluster& cl = *cars[i];
      int part = 0, vi = 0;
      for (int v : cl.occupy_voxels) {
        while (part < (int)cl.part_end.size() && vi >= cl.part_end[part]) ++part;
        ++vi;
        int cnt = pre.vox_cnt[v];
        if (cnt <= 0 || (taint && pre.fc.tainted(v))) continue;  // of a tainted voxel only the owned subgroups: below
        seg[4 * si] = (int)k;
        seg[4 * si + 1] = v;
        seg[4 * si + 2] = (int)i;
        seg[4 * si + 3] = part;
        ++si;
        k += cnt;
      }
      for (auto& tu : cl.tunits) {
        seg[4 * si] = (int)k;
        seg[4 * si + 1] = kVirtualVox | pre.sg_tv_off[tu.sg];
        seg[4 * si + 2] = (int)i;
        seg[4 * si + 3] = tu.part;
        ++si;
        k += pre.fc.subgroups[tu.sg].pts.size();
      }
      int ord = 0x40000000;
      for (auto& cr : cl.carried) {
        if (cr.second <= 0) continue;
        seg[4 * si] = (int)k;
        seg[4 * si + 1] = -1 - cr.first;
        seg[4 * si + 2] = (int)i;
        seg[4 * si + 3] = ord++;
        ++si;
        k += cr.second;
      }
    }
    cstart[cars.size()] = k;
    K = k;
    if (si == 0) K = 0;
  }
  c->stat_track_pairs += 1;
  if (K == 0 || vn <= 0) return SCVOD_OK;
  c->stat_track_points += (int64_t)K;
  std::chrono::steady_clock::time_point ta0 = std::chrono::steady_clock::now();
  const double ta0_cpu = g_prof.on ? thread_cpu_ms() : 0;
  const int cap_quads = (int)std::min<size_t>((size_t)ncl * vn, (size_t)1 << 20);
  size_t nblk = 0;
  if (!use_runs) {
    // per block of 256 points: the segment that holds its first point (a CTA of k_track stages only the segments of its
    // block); the index follows the segment table in the same pinned buffer, so one small H2D copy carries both
    int32_t* seg = c->h_treq.p;
    nblk = (K + 255) / 256;
    if (si * 4 + nblk + 1 > c->h_treq.n) return fail(SCVOD_ERR_STATE, "internal: cluster point counts disagree with the voxel table");
    int32_t* fs = c->h_treq.p + si * 4;
    for (size_t s = 0; s < si; ++s) {
      const size_t lo = (size_t)seg[4 * s], hi = (s + 1 < si ? (size_t)seg[4 * (s + 1)] : K) - 1;  // points [lo, hi]
      for (size_t b = (lo + 255) / 256; b <= hi / 256; ++b) fs[b] = (int32_t)s;
    }
    CU(c->d_treq.alloc(si * 4 + nblk + 1));
  }
  CU(c->d_tout[out_buf].alloc(K));
  {
    const size_t need = (size_t)ncl * vn;
    if (need > c->d_first.n) {  // (re)allocation: the table must start out "empty"; afterwards the epilogue of k_track keeps it so
      CU(c->d_first.alloc(need));
      CU(cudaMemsetAsync(c->d_first.p, 0xff, sizeof(unsigned long long) * c->d_first.n, c->tstream));
    }
  }
  CU(c->h_triples.alloc(4 + 4 * (size_t)cap_quads));
  if (!c->d_track_ctr.p) {
    CU(c->d_track_ctr.alloc(4));
    CU(cudaMemsetAsync(c->d_track_ctr.p, 0, sizeof(int32_t) * c->d_track_ctr.n, c->tstream));
  }
  CU(c->d_track_list.alloc((size_t)cap_quads));
  *reinterpret_cast<volatile int32_t*>(c->h_triples.p) = -1;
  std::atomic_thread_fence(std::memory_order_release);
  if (g_prof.on) g_prof.add("    track: buffers", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ta0).count(), thread_cpu_ms() - ta0_cpu);
  {
    PROF("    track: enqueue+wait+read");
    if (!use_runs) CU(cudaMemcpyAsync(c->d_treq.p, c->h_treq.p, sizeof(int32_t) * (si * 4 + nblk + 1), cudaMemcpyHostToDevice, c->tstream));
    // one context: the kernel may fill the GPU (shortest latency); many contexts: one CTA per SM so that their kernels co-run
    const int chains = g_tracking_now.load();
    c->hp.track_ctas_per_sm = chains <= 2 ? 16 : chains <= 6 ? 4 : 1;
    c->launches += launch_track(c->hp, pbp.apri_xyzi.p + pre.base, pbp.vox_off.p + pre.base, pbp.vox_pts.p + pre.base, c->d_tout[in_buf].p,
                                use_runs ? nullptr : reinterpret_cast<const int4*>(c->d_treq.p), use_runs ? nullptr : c->d_treq.p + si * 4,
                                (int)si, use_runs ? &runs : nullptr, pbp.csr.p, pbp.csr.p + pbp.csr_n, pbp.csr.p + 2 * pbp.csr_n,
                                pbp.csr.p + 3 * (size_t)pbp.csr_n, (int)K, T,
                                pbn.bitmap.p + (size_t)next.slot * c->hp.g.words, pbn.word_rank.p + (size_t)next.slot * c->hp.g.words, ncl, vn,
                                c->d_tout[out_buf].p, c->d_first.p, c->d_track_ctr.p, c->d_track_list.p, c->h_triples.p, cap_quads, c->tstream);
    CU(cudaGetLastError());
    // The kernel's last CTA writes the hits and then their count straight into this pinned buffer: poll the count
    // instead of paying a stream synchronisation per frame pair (stream order still protects every device buffer).
    int nt;
    static const bool sync_wait = getenv("SCVOD_TRACK_SYNC") != nullptr;  // A/B switch: stream synchronisation instead of polling
    if (sync_wait) {
      PROF("    track: wait (sync)");
      CU(cudaStreamSynchronize(c->tstream));
      nt = c->h_triples.p[0];
    } else {
      PROF("    track: wait (poll)");
      volatile int32_t* flag = c->h_triples.p;
      // Spin for a few microseconds (a lone context gets its answer with the lowest latency), then yield the core on
      // every probe: with more contexts than host cores per GPU the waiting threads must not starve the ones that have
      // cluster decisions to compute.  Every ~2000 probes make sure the stream is still alive.
      static const int spin_first = getenv("SCVOD_POLL_SPINS") ? atoi(getenv("SCVOD_POLL_SPINS")) : 200;
      static const int nap_us = getenv("SCVOD_POLL_NAP_US") ? atoi(getenv("SCVOD_POLL_NAP_US")) : 0;
      static const int nap_after = getenv("SCVOD_POLL_NAP_AFTER") ? atoi(getenv("SCVOD_POLL_NAP_AFTER")) : 32;
      int spins = 0, since_query = 0;
      while ((nt = *flag) < 0) {
        if (++since_query >= 2000) {
          since_query = 0;
          cudaError_t qe = cudaStreamQuery(c->tstream);
          if (qe == cudaSuccess) {
            nt = *flag;
            if (nt < 0) return fail(SCVOD_ERR_CUDA, "k_track finished without publishing its hit count");
            break;
          }
          if (qe != cudaErrorNotReady) return fail(SCVOD_ERR_CUDA, std::string("k_track: ") + cudaGetErrorString(qe));
        }
        if (++spins > spin_first) {
          // more waiting contexts than host cores: after a few dozen yields the thread sleeps in short naps instead, so that the
          // cores go to the threads that have cluster decisions or launches to do (a nap costs latency on this chain only)
          if (nap_us > 0 && spins > spin_first + nap_after) {
            timespec ts = {0, nap_us * 1000L};
            nanosleep(&ts, nullptr);
          } else {
            sched_yield();
          }
        } else {
#if defined(__x86_64__)
          __builtin_ia32_pause();
#endif
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (nt > cap_quads) {  // entries past the list were not reset by the kernel epilogue: wipe the table before giving up
      cudaMemsetAsync(c->d_first.p, 0xff, sizeof(unsigned long long) * c->d_first.n, c->tstream);
      cudaStreamSynchronize(c->tstream);
      return fail(SCVOD_ERR_CAPACITY, "tracking hit table overflow");
    }
    // counting sort of the quads by cluster (no order inside a cluster: the decisions only need, per next-frame
    // label, the smallest key and the set of hit voxels)
    const int32_t* quads = c->h_triples.p + 4;
    for (int t = 0; t < nt; ++t) c->hit_start[quads[4 * t] + 1]++;
    for (size_t i = 0; i < cars.size(); ++i) c->hit_start[i + 1] += c->hit_start[i];
    c->hit_key.resize(nt);
    c->hit_vox.resize(nt);
    c->hit_cur.assign(c->hit_start.begin(), c->hit_start.end() - 1);
    for (int t = 0; t < nt; ++t) {
      const int32_t* q = quads + 4 * t;
      const int pos = c->hit_cur[q[0]]++;
      c->hit_key[pos] = ((uint64_t)(uint32_t)q[2] << 32) | (uint32_t)q[3];
      c->hit_vox[pos] = q[1];
    }
  }
  return SCVOD_OK;
}

// label -> hit voxels of cluster ci (ssc.cpp:1277-1317 / :1180-1205) from the grouped hits of diff_clusters.  The reference inserts a
// label when the first point that hits one of its voxels comes by, so the insertion order (which fixes the iteration order of the
// unordered_map) is the order of the labels' smallest keys; the voxel lists are sorted afterwards (sampleVec), so their arrival
// order is irrelevant.
static void build_remap(scvod_ctx* c, size_t ci, const std::vector<int>& nlabel, std::unordered_map<int, std::vector<int>>& remap_name) {
  auto& accs = c->label_accs;  // few labels per cluster: linear search, vectors reused across clusters and pairs
  size_t na = 0;
  for (int t = c->hit_start[ci]; t < c->hit_start[ci + 1]; ++t) {
    const int v = c->hit_vox[t];
    const int lab = nlabel[v];
    if (lab == -1) continue;
    size_t k = 0;
    while (k < na && accs[k].lab != lab) ++k;
    if (k == na) {
      if (accs.size() <= na) accs.emplace_back();
      accs[na].lab = lab;
      accs[na].min_key = c->hit_key[t];
      accs[na].vox.clear();
      ++na;
    } else if (c->hit_key[t] < accs[k].min_key) {
      accs[k].min_key = c->hit_key[t];
    }
    accs[k].vox.push_back(v);
  }
  c->label_order.resize(na);
  for (size_t k = 0; k < na; ++k) c->label_order[k] = (int)k;
  std::sort(c->label_order.begin(), c->label_order.end(), [&](int x, int y) { return accs[x].min_key < accs[y].min_key; });
  for (size_t k = 0; k < na; ++k) {
    auto& acc = accs[c->label_order[k]];
    std::sort(acc.vox.begin(), acc.vox.end());  // sampleVec: already unique
    remap_name.insert(std::make_pair(acc.lab, acc.vox));
  }
}

// tracking(frame_pre_, frame_next_) for the car clusters `cars` of frame_pre_, given in cluster_set iteration order
static int track_cars(scvod_ctx* c, FrameHost& pre, FrameHost& next, const float T[12], std::vector<HCluster*>& cars) {
  const scvod_params& P = c->hp.p;
  std::vector<size_t> cstart;
  const int out_buf = 1 - c->tout_cur;
  {
    int rc = diff_clusters(c, pre, next, T, cars, cstart, true);
    if (rc) return rc;
  }

  PROF("  track: host decisions");
  std::vector<int>& nlabel = next.fc.vox_label;
  auto& nset = next.fc.cluster_set;
  for (size_t ci = 0; ci < cars.size(); ++ci) {
    HCluster& cc = *cars[ci];
    if (cc.type != P.car) continue;
    if (cc.track_id == -1) {
      cc.track_id = c->track_name;
      c->track_name++;
    }
    std::unordered_map<int, std::vector<int>> remap_name;
    build_remap(c, ci, nlabel, remap_name);
    if (remap_name.size() == 0) {
      cc.state = 1;
    } else if (remap_name.size() == 1) {
      auto it = remap_name.begin();
      float ratio = (float)it->second.size() / (float)nset[it->first].occupy_voxels.size();
      if (ratio < P.occupancy) {
        if (nset[it->first].type == P.car) {
          cc.state = 1;
        } else {
          cc.state = 0;
          cc.type = nset[it->first].type;
          HCluster cluster_new;
          cluster_new.track_id = cc.track_id;
          cluster_new.name = next.fc.max_name++;
          cluster_new.type = nset[it->first].type;
          cluster_new.occupy_voxels = it->second;
          HCluster& src = nset[it->first];
          {
            // reduceVec (utility.h:445-450) removes every occurrence of each value in turn; one pass that drops the members of the
            // (sorted, unique: sampleVec) hit list leaves the same elements in the same order, in O(n log m) instead of O(n m)
            const std::vector<int>& rem = cluster_new.occupy_voxels;
            src.occupy_voxels.erase(std::remove_if(src.occupy_voxels.begin(), src.occupy_voxels.end(),
                                                   [&](int x) { return std::binary_search(rem.begin(), rem.end(), x); }),
                                    src.occupy_voxels.end());
          }
          int np = 0, lost = 0;  // points the new cluster gets (whole voxels, :1367) / points the old one loses (reduceVec, :1370)
          for (int v : it->second) {
            nlabel[v] = cluster_new.name;
            np += next.vox_cnt[v];
            const std::vector<int>* sgs = next.fc.subgroups_of(v);
            if (!sgs) {
              lost += next.vox_cnt[v];
            } else {  // a tainted voxel: every subgroup goes to the new cluster, the old one loses those it owned
              for (int g : *sgs) cluster_new.tunits.push_back(HCluster::TUnit{g, (int)cluster_new.part_end.size()});
              for (size_t u = 0; u < src.tunits.size();) {
                if (next.fc.subgroups[src.tunits[u].sg].vox == v) {
                  lost += (int)next.fc.subgroups[src.tunits[u].sg].pts.size();
                  src.tunits.erase(src.tunits.begin() + u);
                } else {
                  ++u;
                }
              }
            }
            cluster_new.part_end.push_back((int)cluster_new.part_end.size() + 1);  // one voxel per part: ptIdx order (ssc.cpp:1367)
          }
          cluster_new.npts = np;
          src.npts -= lost;
          src.part_end.clear();  // point order of a split non-car cluster is never read again
          src.part_end.push_back((int)src.occupy_voxels.size());
          {
            const int new_name = cluster_new.name;
            nset.insert(std::make_pair(new_name, std::move(cluster_new)));  // same key, same insertion point: iteration order unchanged
          }
        }
      } else {
        if (nset[it->first].type == P.car) {
          cc.state = 0;
          HCluster& dst = nset[it->first];
          dst.track_id = cc.track_id;
          int len = (int)(cstart[ci + 1] - cstart[ci]);
          if (len > 0) {  // *cloud += *cluster (ssc.cpp:1382): the transformed cloud stays on the device
            dst.carried.push_back(std::make_pair((int)cstart[ci], len));
            dst.n_carried += len;
          }
        }
      }
    } else {
      cc.state = 0;
      HCluster cluster_new;
      cluster_new.track_id = cc.track_id;
      cluster_new.name = next.fc.max_name++;
      cluster_new.type = P.car;
      for (auto& re : remap_name) {
        if (nset[re.first].type == P.car &&
            ((float)re.second.size() / (float)nset[re.first].occupy_voxels.size()) >= P.occupancy) {
          HCluster& src = nset[re.first];
          const int basev = (int)cluster_new.occupy_voxels.size();  // addVec of pts and voxels (ssc.cpp:1409-1410): parts are kept
          cluster_new.occupy_voxels.insert(cluster_new.occupy_voxels.end(), src.occupy_voxels.begin(), src.occupy_voxels.end());
          const int base_parts = (int)cluster_new.part_end.size();
          for (int pe : src.part_end) cluster_new.part_end.push_back(basev + pe);
          for (auto orun : src.own_runs) {  // the same concatenation, as runs of the frame's car CSR
            orun.part_base += base_parts;
            cluster_new.own_runs.push_back(orun);
          }
          for (auto tu : src.tunits) cluster_new.tunits.push_back(HCluster::TUnit{tu.sg, base_parts + tu.part});
          cluster_new.npts += src.npts;
          nset.erase(re.first);
        }
      }
      for (int v : cluster_new.occupy_voxels) nlabel[v] = cluster_new.name;
      {
            const int new_name = cluster_new.name;
            nset.insert(std::make_pair(new_name, std::move(cluster_new)));  // same key, same insertion point: iteration order unchanged
          }
    }
  }
  c->tout_cur = out_buf;  // this pass's output holds the carried clouds of `next`
  c->batches[pre.batch]->labels_current = false;
  c->batches[next.batch]->labels_current = false;
  return SCVOD_OK;
}

static int track_pair(scvod_ctx* c, FrameHost& pre, FrameHost& next, const float* pose_pre, const float* pose_next) {
  PROF("track_pair total");
  float T[12];
  relative_pose(pose_next, pose_pre, T);
  // car clusters of frame_pre_ in cluster_set order (ssc.cpp:1261-1264)
  std::vector<HCluster*> cars;
  {
    PROF("  track: prepare");
    for (auto& cs : pre.fc.cluster_set)
      if (cs.second.type == c->hp.p.car) cars.push_back(&cs.second);
  }
  return track_cars(c, pre, next, T, cars);
}

extern "C" int scvod_track(scvod_ctx* c, const float* poses6, int nposes) {
  if (!c || !poses6) return fail(SCVOD_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  TrackingScope in_chain;
  TrackStreamScope on_tstream(c);
  int nf = std::min<int>((int)c->frames.size(), nposes);
  for (int i = c->tracked; i + 1 < nf; ++i) {
    int rc = track_pair(c, c->frames[i], c->frames[i + 1], poses6 + 6 * i, poses6 + 6 * (i + 1));
    if (rc) return rc;
    c->tracked = i + 1;
  }
  return SCVOD_OK;
}

// ---------------------------------------------------------------------------------------------
// Chain hand-off between contexts: tracking(k, k+1) only reads, of frame k, the clouds of its car clusters in
// cluster_set order (ssc.cpp:1261-1275) and writes their state / type back.  A sequence that is cut into chunks owned by different
// contexts (workers, GPUs, processes) therefore keeps ONE unbroken chain (ssc.cpp:1450-1452) if the tail of chunk i is exported,
// tracked into the head of chunk i+1 by that chunk's context, and the resulting states are applied to the tail.
// Tail layout (host bytes): int32 {magic, ncars, npts, SSC::name}, ncars x int32 {name, track_id, npts, 0}, npts x float4.
// ---------------------------------------------------------------------------------------------
static const int32_t kTailMagic = 0x5C7A11;

extern "C" int scvod_export_tail(scvod_ctx* c, void* buf, size_t cap, size_t* nbytes) {
  if (!c || !nbytes) return fail(SCVOD_ERR_ARG, "null argument");
  if (c->frames.empty()) return fail(SCVOD_ERR_STATE, "no frames");
  if (c->tracked + 1 != (int)c->frames.size()) return fail(SCVOD_ERR_STATE, "scvod_export_tail: track the context's own frames first");
  CU(cudaSetDevice(c->device));
  FrameHost& fr = c->frames.back();
  PersistBatch& pb = *c->batches[fr.batch];
  std::vector<HCluster*> cars;
  for (auto& cs : fr.fc.cluster_set)
    if (cs.second.type == c->hp.p.car) cars.push_back(&cs.second);
  size_t npts = 0;
  for (HCluster* cl : cars) npts += (size_t)std::max(0, cl->npts) + (size_t)std::max(0, cl->n_carried);
  const size_t need = sizeof(int32_t) * 4 * (1 + cars.size()) + sizeof(float) * 4 * npts;
  *nbytes = need;
  if (!buf) return SCVOD_OK;  // size query
  if (cap < need) return fail(SCVOD_ERR_CAPACITY, "scvod_export_tail: buffer too small");
  int32_t* hdr = (int32_t*)buf;
  hdr[0] = kTailMagic;
  hdr[1] = (int32_t)cars.size();
  hdr[2] = (int32_t)npts;
  hdr[3] = c->track_name;
  float* out = (float*)(hdr + 4 * (1 + cars.size()));
  // the frame's apri points and voxel CSR (one scan: ~1 MB) come to the host once
  std::vector<float> xyzi((size_t)std::max(1, fr.n_apri) * 4);
  std::vector<int32_t> vox_off((size_t)std::max(1, fr.n_vox)), vox_pts((size_t)std::max(1, fr.n_apri));
  CU(cudaStreamSynchronize(c->stream));
  if (fr.n_apri > 0) {
    CU(cudaMemcpy(xyzi.data(), pb.apri_xyzi.p + fr.base, sizeof(float4) * fr.n_apri, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(vox_pts.data(), pb.vox_pts.p + fr.base, sizeof(int32_t) * fr.n_apri, cudaMemcpyDeviceToHost));
  }
  if (fr.n_vox > 0) CU(cudaMemcpy(vox_off.data(), pb.vox_off.p + fr.base, sizeof(int32_t) * fr.n_vox, cudaMemcpyDeviceToHost));
  const bool taint = !fr.fc.tvox.empty();
  size_t pos = 0;
  std::vector<int> ms;
  for (size_t i = 0; i < cars.size(); ++i) {
    HCluster& cl = *cars[i];
    const size_t pos0 = pos;
    // own points: part by part (the concatenation order of fusions, ssc.cpp:617,1409), ascending apri index inside a part (:360-380)
    const int nparts = (int)cl.part_end.size();
    int vi = 0;
    for (int part = 0; part < nparts; ++part) {
      ms.clear();
      for (; vi < cl.part_end[part]; ++vi) {
        const int v = cl.occupy_voxels[vi];
        if (taint && fr.fc.tainted(v)) continue;
        for (int j = 0; j < fr.vox_cnt[v]; ++j) ms.push_back(vox_pts[vox_off[v] + j]);
      }
      for (auto& tu : cl.tunits)
        if (tu.part == part) ms.insert(ms.end(), fr.fc.subgroups[tu.sg].pts.begin(), fr.fc.subgroups[tu.sg].pts.end());
      std::sort(ms.begin(), ms.end());
      for (int m : ms) {
        std::memcpy(out + 4 * pos, xyzi.data() + 4 * (size_t)m, sizeof(float) * 4);
        ++pos;
      }
    }
    if ((int)(pos - pos0) != std::max(0, cl.npts)) return fail(SCVOD_ERR_STATE, "internal: cluster point count disagrees with its voxels");
    for (auto& cr : cl.carried) {  // clouds appended by tracking (ssc.cpp:1382), device resident
      if (cr.second <= 0) continue;
      CU(cudaMemcpy(out + 4 * pos, c->d_tout[c->tout_cur].p + cr.first, sizeof(float4) * cr.second, cudaMemcpyDeviceToHost));
      pos += cr.second;
    }
    int32_t* rec = hdr + 4 * (1 + i);
    rec[0] = cl.name;
    rec[1] = cl.track_id;
    rec[2] = (int32_t)(pos - pos0);
    rec[3] = 0;
  }
  if (pos != npts) return fail(SCVOD_ERR_STATE, "internal: exported point count");
  return SCVOD_OK;
}

extern "C" int scvod_track_from_tail(scvod_ctx* c, const void* tail, size_t nbytes, const float pose_pre6[6], const float pose_next6[6],
                                     int32_t* state_type, int cap) {
  if (!c || !tail || !pose_pre6 || !pose_next6) return fail(SCVOD_ERR_ARG, "null argument");
  if (c->frames.empty()) return fail(SCVOD_ERR_STATE, "scvod_track_from_tail: push the chunk's scans first");
  if (c->tracked != 0 || c->head_tracked) return fail(SCVOD_ERR_STATE, "scvod_track_from_tail: the head frame has already been tracked");
  const int32_t* hdr = (const int32_t*)tail;
  if (nbytes < sizeof(int32_t) * 4 || hdr[0] != kTailMagic || hdr[1] < 0 || hdr[2] < 0) return fail(SCVOD_ERR_ARG, "not a tail buffer");
  const int ncars = hdr[1], npts = hdr[2];
  if (nbytes < sizeof(int32_t) * 4 * (1 + (size_t)ncars) + sizeof(float) * 4 * (size_t)npts) return fail(SCVOD_ERR_ARG, "truncated tail buffer");
  if (state_type && cap < ncars) return fail(SCVOD_ERR_CAPACITY, "state buffer too small");
  CU(cudaSetDevice(c->device));
  c->track_name = hdr[3];  // SSC::name runs on along the chain (ssc.cpp:1267-1271)
  const float* pts = (const float*)(hdr + 4 * (1 + (size_t)ncars));
  // the imported clouds become "carried" ranges of a stand-in frame_pre_
  const int in_buf = c->tout_cur;
  CU(c->d_tout[in_buf].alloc((size_t)std::max(1, npts)));
  if (npts > 0) CU(cudaMemcpyAsync(c->d_tout[in_buf].p, pts, sizeof(float4) * (size_t)npts, cudaMemcpyHostToDevice, c->stream));  // before the scope below joins
  FrameHost& next = c->frames[0];
  FrameHost pre;
  pre.batch = next.batch;
  pre.slot = next.slot;
  pre.base = next.base;
  pre.csr_base = 0;
  std::vector<HCluster> store((size_t)ncars);
  std::vector<HCluster*> cars;
  int off = 0;
  for (int i = 0; i < ncars; ++i) {
    const int32_t* rec = hdr + 4 * (1 + (size_t)i);
    HCluster& cl = store[i];
    cl.name = rec[0];
    cl.track_id = rec[1];
    cl.type = c->hp.p.car;
    cl.npts = 0;
    if (rec[2] > 0) cl.carried.push_back(std::make_pair(off, rec[2]));
    cl.n_carried = rec[2];
    off += rec[2];
    cars.push_back(&cl);
  }
  if (off != npts) return fail(SCVOD_ERR_ARG, "tail buffer: point counts disagree");
  float T[12];
  relative_pose(pose_next6, pose_pre6, T);
  TrackingScope in_chain;
  int rc;
  {
    TrackStreamScope on_tstream(c);
    rc = track_cars(c, pre, next, T, cars);
  }
  CU(cudaStreamSynchronize(c->stream));  // the caller may free `tail` now
  if (rc) return rc;
  c->head_tracked = true;
  if (state_type)
    for (int i = 0; i < ncars; ++i) {
      state_type[2 * i] = store[i].state;
      state_type[2 * i + 1] = store[i].type;
    }
  return ncars;
}

extern "C" int scvod_apply_tail_states(scvod_ctx* c, const int32_t* state_type, int n) {
  if (!c || (!state_type && n > 0)) return fail(SCVOD_ERR_ARG, "null argument");
  if (c->frames.empty()) return fail(SCVOD_ERR_STATE, "no frames");
  FrameHost& fr = c->frames.back();
  std::vector<HCluster*> cars;
  for (auto& cs : fr.fc.cluster_set)
    if (cs.second.type == c->hp.p.car) cars.push_back(&cs.second);
  if ((int)cars.size() != n) return fail(SCVOD_ERR_STATE, "scvod_apply_tail_states: the tail frame changed since it was exported");
  for (int i = 0; i < n; ++i) {
    cars[i]->state = state_type[2 * i];
    cars[i]->type = state_type[2 * i + 1];
  }
  c->batches[fr.batch]->labels_current = false;
  return SCVOD_OK;
}

// ---------------------------------------------------------------------------------------------
// initialization (SSC::intialization, reference src/ssc.cpp:1148-1248)
// ---------------------------------------------------------------------------------------------
extern "C" int scvod_initialization(scvod_ctx* c, const float* poses6, int nposes, int32_t* id_based_out) {
  if (!c || !poses6 || !id_based_out) return fail(SCVOD_ERR_ARG, "null argument");
  if (c->tracked > 0) return fail(SCVOD_ERR_STATE, "scvod_initialization works on untracked frames: call it before scvod_track");
  CU(cudaSetDevice(c->device));
  const int nf = std::min<int>((int)c->frames.size(), nposes);
  if (nf <= 0) return fail(SCVOD_ERR_STATE, "no frames");
  const scvod_params& P = c->hp.p;
  int max_num = 999999, id_based = 0;  // the LAST frame with the fewest clusters (<=, ssc.cpp:1153-1158)
  for (int i = 0; i < nf; ++i) {
    if ((int)c->frames[i].fc.cluster_set.size() <= max_num) {
      max_num = (int)c->frames[i].fc.cluster_set.size();
      id_based = i;
    }
  }
  FrameHost& base = c->frames[id_based];
  c->init_fc = base.fc;  // Frame frame_based = frames_[id_based]: same container copy as the reference
  c->init_base = id_based;
  FrameClusters& fb = c->init_fc;
  std::vector<size_t> cstart;
  TrackStreamScope on_tstream(c);
  for (int i = 0; i < nf; ++i) {
    if (i == id_based) continue;
    FrameHost& fi = c->frames[i];
    float T[12];
    relative_pose(poses6 + 6 * id_based, poses6 + 6 * i, T);  // trans_based.inverse() * trans_i (:1172)
    std::vector<HCluster*> all;  // every cluster of frame i, in cluster_set order (:1180)
    for (auto& cs : fi.fc.cluster_set) all.push_back(&cs.second);
    int rc = diff_clusters(c, fi, base, T, all, cstart, false);
    if (rc) return rc;
    for (size_t ci = 0; ci < all.size(); ++ci) {
      std::unordered_map<int, std::vector<int>> remap_name;
      build_remap(c, ci, fb.vox_label, remap_name);
      if (remap_name.size() <= 1) continue;
      // name stays -1 when no label passes the occupancy test (utility.h:154).  bounding_box is never recomputed for the fused
      // cluster (:1208-1241): it keeps the zeros of a default pcl::PointXYZI pair, which is what recognize(frame_based) then reads
      HCluster fusion;
      std::vector<int> erase_id;
      for (auto& re : remap_name) {
        HCluster& src = fb.cluster_set[re.first];
        if (((float)re.second.size() / (float)src.occupy_voxels.size()) >= P.occupancy) {
          erase_id.emplace_back(re.first);
          fusion.name = re.first;
          const int basev = (int)fusion.occupy_voxels.size();
          fusion.occupy_voxels.insert(fusion.occupy_voxels.end(), src.occupy_voxels.begin(), src.occupy_voxels.end());
          const int base_parts = (int)fusion.part_end.size();
          for (int pe : src.part_end) fusion.part_end.push_back(basev + pe);
          for (auto tu : src.tunits) fusion.tunits.push_back(HCluster::TUnit{tu.sg, base_parts + tu.part});
          fusion.npts += src.npts;
        }
      }
      for (int e : erase_id) fb.cluster_set.erase(e);
      const int fname = fusion.name;
      const std::vector<int> fvox = fusion.occupy_voxels;
      fb.cluster_set.insert(std::make_pair(fname, std::move(fusion)));
      for (int v : fvox) fb.vox_label[v] = fname;
    }
  }
  recognize_clusters(P, fb);  // recognize(frame_based), :1241
  c->have_init = true;
  *id_based_out = id_based;
  return SCVOD_OK;
}

// ---------------------------------------------------------------------------------------------
// labels
// ---------------------------------------------------------------------------------------------
// per-voxel classes of every frame of a batch (decided on the host) -> per-point classes on the device
static int refresh_batch_labels(scvod_ctx* c, int batch) {
  PersistBatch& pb = *c->batches[batch];
  if (pb.labels_current) return SCVOD_OK;
  std::vector<FrameHost*> frs(pb.nscans, nullptr);
  for (auto& fr : c->frames)
    if (fr.batch == batch) frs[fr.slot] = &fr;
  size_t vtot = 0;
  for (FrameHost* fr : frs) vtot += fr ? ((fr->n_vox + 3) & ~3) : 0;
  const size_t words = pb.nscans + (vtot + 3) / 4 + 1;
  CU(c->h_vcls.alloc(words));
  CU(c->d_vcls.alloc(words));
  int32_t* voff = c->h_vcls.p;
  uint8_t* vc = reinterpret_cast<uint8_t*>(c->h_vcls.p + pb.nscans);
  size_t pos = 0;
  for (int s = 0; s < pb.nscans; ++s) {
    voff[s] = (int32_t)pos;
    FrameHost* fr = frs[s];
    if (!fr) continue;
    std::memset(vc + pos, SCVOD_PT_UNCLUSTERED, fr->n_vox);
    for (auto& cs : fr->fc.cluster_set) {
      uint8_t v = (cs.second.state == 1) ? SCVOD_PT_DYNAMIC : SCVOD_PT_STATIC;
      for (int vx : cs.second.occupy_voxels) vc[pos + vx] = v;
    }
    pos += (fr->n_vox + 3) & ~3;
  }
  CU(cudaMemcpyAsync(c->d_vcls.p, c->h_vcls.p, sizeof(int32_t) * words, cudaMemcpyHostToDevice, c->stream));
  c->launches += launch_final_labels(pb.off_dev.p, pb.scan_counts_dev.p, pb.nscans, pb.max_n, pb.apri_src.p, pb.apri_cid.p, c->d_vcls.p,
                                     reinterpret_cast<const uint8_t*>(c->d_vcls.p + pb.nscans), pb.cls.p, c->stream);
  CU(cudaGetLastError());
  // Points of tainted voxels follow the cluster that owns them, not their voxel: the subgroup's class is that of the LAST
  // cluster (cluster_set order, as the per-cluster emission of saveSegCloud, ssc.cpp:477-545) that lists it.
  {
    size_t nit = 0;
    for (FrameHost* fr : frs)
      if (fr)
        for (auto& g : fr->fc.subgroups) nit += g.pts.size();
    if (nit > 0) {
      CU(c->h_lov.alloc(3 * nit));
      CU(c->d_lov.alloc(3 * nit));
      int32_t* items = c->h_lov.p;
      int32_t* iscan = c->h_lov.p + 2 * nit;
      size_t k = 0;
      for (int s = 0; s < pb.nscans; ++s) {
        FrameHost* fr = frs[s];
        if (!fr || fr->fc.subgroups.empty()) continue;
        std::vector<uint8_t> sg_cls(fr->fc.subgroups.size(), (uint8_t)SCVOD_PT_UNCLUSTERED);
        for (auto& cs : fr->fc.cluster_set)
          for (auto& tu : cs.second.tunits) sg_cls[tu.sg] = (cs.second.state == 1) ? SCVOD_PT_DYNAMIC : SCVOD_PT_STATIC;
        for (size_t g = 0; g < fr->fc.subgroups.size(); ++g)
          for (int m : fr->fc.subgroups[g].pts) {
            items[2 * k] = (int32_t)(fr->base + m);
            items[2 * k + 1] = sg_cls[g];
            iscan[k] = s;
            ++k;
          }
      }
      CU(cudaMemcpyAsync(c->d_lov.p, c->h_lov.p, sizeof(int32_t) * 3 * nit, cudaMemcpyHostToDevice, c->stream));
      c->launches += launch_label_override(c->d_lov.p, c->d_lov.p + 2 * nit, (int)nit, pb.apri_src.p, pb.off_dev.p, pb.cls.p, c->stream);
      CU(cudaGetLastError());
    }
  }
  CU(wait_stream(c, c->stream));  // the staging buffers are reused by the next batch
  pb.labels_current = true;
  return SCVOD_OK;
}

extern "C" int scvod_frame_labels(scvod_ctx* c, int frame, uint8_t* cls, int n) {
  if (!c || !cls || frame < 0 || frame >= (int)c->frames.size()) return fail(SCVOD_ERR_ARG, "bad frame index");
  CU(cudaSetDevice(c->device));
  FrameHost& fr = c->frames[frame];
  if (n < fr.n_in) return fail(SCVOD_ERR_ARG, "label buffer too small");
  int rc = refresh_batch_labels(c, fr.batch);
  if (rc) return rc;
  PersistBatch& pb = *c->batches[fr.batch];
  if (fr.n_in > 0) CU(cudaMemcpyAsync(cls, pb.cls.p + fr.base, fr.n_in, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SCVOD_OK;
}

// labels of frames [f0,f1) into one host buffer (concatenated in frame order); cls == NULL: refresh only
extern "C" int scvod_labels_range(scvod_ctx* c, int f0, int f1, uint8_t* cls, int64_t cap) {
  if (!c || f0 < 0 || f1 > (int)c->frames.size() || f0 > f1) return fail(SCVOD_ERR_ARG, "bad frame range");
  PROF("labels_range total");
  CU(cudaSetDevice(c->device));
  int64_t pos = 0;
  int f = f0;
  while (f < f1) {
    FrameHost& fr = c->frames[f];
    int rc = refresh_batch_labels(c, fr.batch);
    if (rc) return rc;
    // frames of one batch are contiguous in frame order and in device memory: one copy per run
    int g = f;
    int64_t bytes = 0;
    while (g < f1 && c->frames[g].batch == fr.batch) {
      bytes += c->frames[g].n_in;
      ++g;
    }
    if (cls) {
      if (pos + bytes > cap) return fail(SCVOD_ERR_ARG, "label buffer too small");
      PersistBatch& pb = *c->batches[fr.batch];
      if (bytes > 0) CU(cudaMemcpyAsync(cls + pos, pb.cls.p + fr.base, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    pos += bytes;
    f = g;
  }
  CU(wait_stream(c, c->stream));
  return SCVOD_OK;
}

extern "C" int scvod_static_submap_dev(scvod_ctx* c, int f0, int f1, const float* poses6, void* out_xyzi_dev, int64_t cap_points,
                                       int64_t* n_points) {
  if (!c || !poses6 || !out_xyzi_dev || !n_points || f0 < 0 || f1 > (int)c->frames.size() || f0 > f1)
    return fail(SCVOD_ERR_ARG, "bad arguments to scvod_static_submap_dev");
  PROF("static_submap total");
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(c->d_counter.p, 0, sizeof(unsigned long long), c->stream));
  const int nf = f1 - f0;
  CU(c->h_Ts.alloc(std::max(1, nf) * 12));
  CU(c->d_Ts.alloc(std::max(1, nf) * 12));
  for (int f = f0; f < f1; ++f) pose_matrix(poses6 + 6 * f, c->h_Ts.p + 12 * (f - f0));
  if (nf > 0) CU(cudaMemcpyAsync(c->d_Ts.p, c->h_Ts.p, sizeof(float) * 12 * nf, cudaMemcpyHostToDevice, c->stream));
  int f = f0;
  while (f < f1) {
    FrameHost& fr = c->frames[f];
    int rc = refresh_batch_labels(c, fr.batch);
    if (rc) return rc;
    int g = f;
    while (g < f1 && c->frames[g].batch == fr.batch) ++g;
    PersistBatch& pb = *c->batches[fr.batch];
    c->launches += launch_submap(pb.pts.p, pb.cls.p, pb.off_dev.p, c->d_Ts.p + 12 * (f - f0), fr.slot, g - f, pb.max_n,
                                 (float4*)out_xyzi_dev, c->d_counter.p, cap_points,
                                 c->submap_all_static ? (0xffu & ~(1u << SCVOD_PT_DYNAMIC)) : (1u << SCVOD_PT_STATIC), c->stream);
    CU(cudaGetLastError());
    f = g;
  }
  unsigned long long cnt = 0;
  CU(cudaMemcpyAsync(&cnt, c->d_counter.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
  CU(wait_stream(c, c->stream));
  *n_points = (int64_t)std::min<unsigned long long>(cnt, (unsigned long long)cap_points);
  return cnt > (unsigned long long)cap_points ? fail(SCVOD_ERR_CAPACITY, "submap buffer too small") : SCVOD_OK;
}

// ---------------------------------------------------------------------------------------------
// inspection
// ---------------------------------------------------------------------------------------------
#define FRAME_OR_FAIL()                                                                                  \
  if (!c || frame < 0 || frame >= (int)c->frames.size()) return fail(SCVOD_ERR_ARG, "bad frame index"); \
  CU(cudaSetDevice(c->device));                                                                          \
  FrameHost& fr = c->frames[frame];                                                                      \
  PersistBatch& pb = *c->batches[fr.batch];                                                              \
  (void)pb;
// the same, but SCVOD_INIT_FRAME selects the frame produced by scvod_initialization: per-point / per-voxel data of its
// base frame, clusters and voxel labels of the initialised copy (fcl)
#define FRAME_OR_INIT_OR_FAIL()                                                                                          \
  if (c && frame == SCVOD_INIT_FRAME && !c->have_init) return fail(SCVOD_ERR_STATE, "scvod_initialization has not run"); \
  const bool is_init__ = c && frame == SCVOD_INIT_FRAME;                                                                 \
  if (is_init__) frame = c->init_base;                                                                                   \
  FRAME_OR_FAIL();                                                                                                       \
  FrameClusters& fcl = is_init__ ? c->init_fc : fr.fc;

template <typename T>
static int d2h(scvod_ctx* c, T* dst, const T* src, size_t n) {
  if (!dst || n == 0) return SCVOD_OK;
  CU(cudaMemcpyAsync(dst, src, sizeof(T) * n, cudaMemcpyDeviceToHost, c->stream));
  return SCVOD_OK;
}

extern "C" int scvod_frame_counts(scvod_ctx* c, int frame, int32_t counts[9]) {
  FRAME_OR_INIT_OR_FAIL();
  counts[0] = fr.n_in;
  counts[1] = fr.n_ground;
  counts[2] = fr.n_ng;
  counts[3] = fr.n_apri;
  counts[4] = fr.n_vox;
  counts[5] = fcl.n_clusters[0];
  counts[6] = fcl.n_clusters[1];
  counts[7] = fcl.n_clusters[2];
  counts[8] = (int)fcl.cluster_set.size();
  return SCVOD_OK;
}

extern "C" int scvod_frame_ground_order(scvod_ctx* c, int frame, int32_t* ground_src, int32_t* nonground_src) {
  FRAME_OR_FAIL();
  int rc = d2h(c, ground_src, pb.ground_src.p + fr.base, fr.n_ground);
  if (rc) return rc;
  rc = d2h(c, nonground_src, pb.ng_src.p + fr.base, fr.n_ng);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return SCVOD_OK;
}

extern "C" int scvod_frame_apri(scvod_ctx* c, int frame, int32_t* src, int32_t* voxel_idx) {
  FRAME_OR_FAIL();
  int rc = d2h(c, src, pb.apri_src.p + fr.base, fr.n_apri);
  if (rc) return rc;
  rc = d2h(c, voxel_idx, pb.apri_vid.p + fr.base, fr.n_apri);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return SCVOD_OK;
}

extern "C" int scvod_frame_voxels(scvod_ctx* c, int frame, int32_t* voxel_idx, int32_t* count, float* av, float* cov, float* center,
                                  int32_t* tri, int32_t* label) {
  FRAME_OR_INIT_OR_FAIL();
  int rc = 0;
  rc |= d2h(c, voxel_idx, pb.vox_vid.p + fr.base, fr.n_vox);
  rc |= d2h(c, count, pb.vox_cnt.p + fr.base, fr.n_vox);
  rc |= d2h(c, av, pb.vox_av.p + fr.base, fr.n_vox);
  rc |= d2h(c, cov, pb.vox_cov.p + fr.base, fr.n_vox);
  rc |= d2h(c, center, pb.vox_center.p + 3 * fr.base, (size_t)fr.n_vox * 3);
  rc |= d2h(c, tri, pb.vox_tri.p + 3 * fr.base, (size_t)fr.n_vox * 3);
  if (rc) return SCVOD_ERR_CUDA;
  CU(cudaStreamSynchronize(c->stream));
  if (label) std::memcpy(label, fcl.vox_label.data(), sizeof(int32_t) * fr.n_vox);
  return SCVOD_OK;
}

extern "C" int scvod_frame_point_cluster(scvod_ctx* c, int frame, int stage, int32_t* name) {
  FRAME_OR_FAIL();
  if (stage < 0 || stage > 2 || !name) return fail(SCVOD_ERR_ARG, "bad stage");
  if ((int)fr.fc.vox_name_stage[stage].size() != fr.n_vox) return fail(SCVOD_ERR_STATE, "stage names not kept (inspect option off)");
  std::vector<int32_t> cid(std::max(1, fr.n_apri));
  int rc = d2h(c, cid.data(), pb.apri_cid.p + fr.base, fr.n_apri);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  for (int m = 0; m < fr.n_apri; ++m) name[m] = cid[m] >= 0 ? fr.fc.vox_name_stage[stage][cid[m]] : -1;
  for (auto& g : fr.fc.subgroups)  // points of tainted voxels: the cluster of the point, not of the voxel
    for (int m : g.pts) name[m] = g.stage_name[stage];
  return SCVOD_OK;
}

extern "C" int scvod_frame_clusters(scvod_ctx* c, int frame, int cap, int32_t* name, int32_t* type, int32_t* state, int32_t* npts,
                                    int32_t* nvox, float* bbox) {
  FRAME_OR_INIT_OR_FAIL();
  int i = 0;
  for (auto& cs : fcl.cluster_set) {
    if (i >= cap) break;
    if (name) name[i] = cs.first;
    if (type) type[i] = cs.second.type;
    if (state) state[i] = cs.second.state;
    if (npts) npts[i] = cs.second.npts;
    if (nvox) nvox[i] = (int)cs.second.occupy_voxels.size();
    if (bbox)
      for (int d = 0; d < 3; ++d) {
        bbox[6 * i + d] = cs.second.bb_min[d];
        bbox[6 * i + 3 + d] = cs.second.bb_max[d];
      }
    ++i;
  }
  return (int)fcl.cluster_set.size();
}

// ---------------------------------------------------------------------------------------------
// single-stage entry points
// ---------------------------------------------------------------------------------------------
extern "C" int scvod_ground(scvod_ctx* c, const float* xyzi, int n, int32_t* ground_idx, int32_t* n_ground, int32_t* nonground_idx,
                            int32_t* n_nonground) {
  if (!c || (!xyzi && n > 0) || !ground_idx || !n_ground || !nonground_idx || !n_nonground || n < 0)
    return fail(SCVOD_ERR_ARG, "bad arguments to scvod_ground");
  if (n > c->max_points) return fail(SCVOD_ERR_CAPACITY, "scan larger than max_points");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  DevBuf<float4> pts, apri_xyzi;
  DevBuf<uint8_t> cls;
  DevBuf<int32_t> apri_src, apri_vid, gsrc, ngsrc;
  const size_t T = (size_t)std::max(n, 1);
  CU(pts.alloc(T));
  CU(apri_xyzi.alloc(T));
  CU(cls.alloc(T));
  CU(apri_src.alloc(T));
  CU(apri_vid.alloc(T));
  CU(gsrc.alloc(T));
  CU(ngsrc.alloc(T));
  BatchDev w = c->ws;
  w.pts = pts.p;
  w.apri_xyzi = apri_xyzi.p;
  w.cls = cls.p;
  w.apri_src = apri_src.p;
  w.apri_vid = apri_vid.p;
  w.ground_src = gsrc.p;
  w.ng_src = ngsrc.p;
  int64_t off[2] = {0, n};
  CU(cudaMemcpyAsync(w.off, off, sizeof(off), cudaMemcpyHostToDevice, st));
  if (n) CU(cudaMemcpyAsync(w.pts, xyzi, sizeof(float4) * n, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(w.scan_counts, 0, sizeof(int32_t) * ((size_t)w.cap_scans * 8 + 8), st));
  c->launches += launch_ground(c->hp, w, 1, n, st);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_scan_counts.p, w.scan_counts, sizeof(int32_t) * ((size_t)w.cap_scans * 8 + 8), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (c->h_scan_counts.p[(size_t)w.cap_scans * 8] & 1) return fail(SCVOD_ERR_CAPACITY, "a PatchWork patch holds more points than the largest fit tile");
  *n_ground = c->h_scan_counts.p[0];
  *n_nonground = c->h_scan_counts.p[1];
  if (*n_ground) CU(cudaMemcpyAsync(ground_idx, gsrc.p, sizeof(int32_t) * *n_ground, cudaMemcpyDeviceToHost, st));
  if (*n_nonground) CU(cudaMemcpyAsync(nonground_idx, ngsrc.p, sizeof(int32_t) * *n_nonground, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  pts.release(); apri_xyzi.release(); cls.release(); apri_src.release(); apri_vid.release(); gsrc.release(); ngsrc.release();
  return SCVOD_OK;
}

// debug view of the per-patch plane fits of the most recent ground pass of scan slot `slot`:
// 12 floats per patch: normal[3], mean[3], singular values[3], d, decision, npts (0 rows for skipped patches)
extern "C" int scvod_last_patch_records(scvod_ctx* c, int slot, float* rec504x12) {
  if (!c || !rec504x12 || slot < 0 || slot >= c->max_batch) return fail(SCVOD_ERR_ARG, "bad arguments");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy(rec504x12, c->ws.patch_dbg + (size_t)slot * kNumPatches * 12, sizeof(float) * kNumPatches * 12, cudaMemcpyDeviceToHost));
  return SCVOD_OK;
}

extern "C" int scvod_bin(scvod_ctx* c, const float* xyzi, int n, uint8_t* pass, int32_t* voxel_idx, int32_t* range_idx,
                         int32_t* sector_idx, int32_t* azimuth_idx, float* range, float* angle, float* azimuth) {
  if (!c || (!xyzi && n > 0) || n < 0) return fail(SCVOD_ERR_ARG, "bad arguments to scvod_bin");
  CU(cudaSetDevice(c->device));
  if (n == 0) return SCVOD_OK;
  cudaStream_t st = c->stream;
  DevBuf<float4> pts;
  DevBuf<uint8_t> dpass;
  DevBuf<int32_t> dvid, dri, dsi, dei;
  DevBuf<float> dr, da_, daz;
  CU(pts.alloc(n));
  CU(dpass.alloc(n));
  CU(dvid.alloc(n));
  CU(dri.alloc(n));
  CU(dsi.alloc(n));
  CU(dei.alloc(n));
  CU(dr.alloc(n));
  CU(da_.alloc(n));
  CU(daz.alloc(n));
  CU(cudaMemcpyAsync(pts.p, xyzi, sizeof(float4) * n, cudaMemcpyHostToDevice, st));
  c->launches += launch_bin_only(c->hp, pts.p, n, dpass.p, dvid.p, dri.p, dsi.p, dei.p, dr.p, da_.p, daz.p, st);
  CU(cudaGetLastError());
  int rc = 0;
  rc |= d2h(c, pass, dpass.p, n);
  rc |= d2h(c, voxel_idx, dvid.p, n);
  rc |= d2h(c, range_idx, dri.p, n);
  rc |= d2h(c, sector_idx, dsi.p, n);
  rc |= d2h(c, azimuth_idx, dei.p, n);
  rc |= d2h(c, range, dr.p, n);
  rc |= d2h(c, angle, da_.p, n);
  rc |= d2h(c, azimuth, daz.p, n);
  CU(cudaStreamSynchronize(st));
  pts.release(); dpass.release(); dvid.release(); dri.release(); dsi.release(); dei.release(); dr.release(); da_.release(); daz.release();
  return rc ? SCVOD_ERR_CUDA : SCVOD_OK;
}

// Test hook: soundness of the binning filter (dev_bin_filtered == dev_bin_point) on n generated points (xyzi == NULL: counter-based
// generator with structured edge cases, coordinates within +-extent) or on n caller points.  stats5: points, points that took the
// exact chain, mismatches, max |q_approx - q_exact| * 1e9 of the sector and of the azimuth coordinate among filter-decided points,
// then (generated points only) the same two counts for the patch-assignment filter of the ground stage (dev_patch_filtered).
extern "C" int scvod_bin_filter_check(scvod_ctx* c, const float* xyzi, int64_t n, uint32_t seed, float extent, uint64_t* stats7) {
  if (!c || !stats7 || n < 0) return fail(SCVOD_ERR_ARG, "bad arguments to scvod_bin_filter_check");
  CU(cudaSetDevice(c->device));
  DevBuf<unsigned long long> st;
  DevBuf<float4> pts;
  CU(st.alloc(8));
  CU(cudaMemsetAsync(st.p, 0, sizeof(unsigned long long) * 8, c->stream));
  if (xyzi && n > 0) {
    CU(pts.alloc(n));
    CU(cudaMemcpyAsync(pts.p, xyzi, sizeof(float4) * n, cudaMemcpyHostToDevice, c->stream));
  }
  if (n > 0) c->launches += launch_bin_filter_check(c->hp, n, seed, extent, xyzi ? pts.p : nullptr, st.p, c->stream);
  if (n > 0 && !xyzi) c->launches += launch_patch_filter_check(c->hp, n, seed, extent, st.p, c->stream);  // G1 (pc2czm) filter: stats[5], [6]
  CU(cudaGetLastError());
  unsigned long long h[7] = {0, 0, 0, 0, 0, 0, 0};
  CU(cudaMemcpyAsync(h, st.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 7; ++i) stats7[i] = h[i];
  st.release();
  pts.release();
  return SCVOD_OK;
}

// device port of glibc atan2f evaluated on the GPU (test hook for the libm-parity check)
extern "C" int scvod_atan2f_device(scvod_ctx* c, const float* y, const float* x, float* out, int64_t n) {
  if (!c || !y || !x || !out || n < 0) return fail(SCVOD_ERR_ARG, "bad arguments");
  CU(cudaSetDevice(c->device));
  if (n == 0) return SCVOD_OK;
  DevBuf<float> dy, dx, dout;
  CU(dy.alloc(n));
  CU(dx.alloc(n));
  CU(dout.alloc(n));
  CU(cudaMemcpyAsync(dy.p, y, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(dx.p, x, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
  c->launches += launch_atan2f_probe(dy.p, dx.p, dout.p, n, c->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  dy.release(); dx.release(); dout.release();
  return SCVOD_OK;
}

// Host-only hook: the cluster bookkeeping (host_cluster.cpp) on caller-provided voxel tables, without a
// GPU.  Used by the CPU test-suite to check the name replay / fusion / bounding-box / car logic against
// the oracle; the product pipeline calls segment_and_recognize() directly with GPU-produced tables.
extern "C" int scvod_host_segment_pts(const scvod_params* p, int V, const int32_t* vox_cnt, const int32_t* vox_root, const int32_t* vox_nbr,
                                      const float* vox_bbox, int n_events, const int32_t* ev_cid, int n_edges, const int32_t* edges, int n_tvox,
                                      const int32_t* tv_cid, const int32_t* tv_base, int n_tpts, const int32_t* tp_m, const float* tp_xyz,
                                      const int32_t* tp_nbr, int32_t* name_stage0, int32_t* name_stage1, int32_t* name_stage2,
                                      int32_t* tp_stage, int32_t n_clusters[3], int cap, int32_t* cluster_name, int32_t* cluster_type,
                                      int32_t* cluster_npts, int32_t* cluster_nvox, int32_t* max_name) {
  if (!p || V < 0 || n_tvox < 0 || n_tpts < 0) return fail(SCVOD_ERR_ARG, "bad arguments");
  ScanTables t;
  t.V = V;
  t.n_events = n_events;
  t.n_edges = n_edges;
  t.vox_cnt = vox_cnt;
  t.vox_root = vox_root;
  t.vox_nbr = vox_nbr;
  t.vox_bbox = vox_bbox;
  t.ev_cid = ev_cid;
  t.edges = edges;
  t.n_tvox = n_tvox;
  t.n_tpts = n_tpts;
  t.tv_cid = tv_cid;
  t.tv_base = tv_base;
  t.tp_m = tp_m;
  t.tp_xyz = tp_xyz;
  t.tp_nbr = tp_nbr;
  FrameClusters fc;
  if (!segment_and_recognize(*p, t, fc, true)) return fail(SCVOD_ERR_STATE, "replayed partition differs from the given components");
  if (name_stage0) std::memcpy(name_stage0, fc.vox_name_stage[0].data(), sizeof(int32_t) * V);
  if (name_stage1) std::memcpy(name_stage1, fc.vox_name_stage[1].data(), sizeof(int32_t) * V);
  if (name_stage2) std::memcpy(name_stage2, fc.vox_name_stage[2].data(), sizeof(int32_t) * V);
  if (tp_stage && n_tpts > 0) {  // [3][n_tpts]: cluster of every point of a tainted voxel after the three stages
    std::unordered_map<int, int> pos_of_m;
    for (int q = 0; q < n_tpts; ++q) pos_of_m[tp_m[q]] = q;
    for (auto& g : fc.subgroups)
      for (int m : g.pts)
        for (int st = 0; st < 3; ++st) tp_stage[(size_t)st * n_tpts + pos_of_m[m]] = g.stage_name[st];
  }
  if (n_clusters)
    for (int i = 0; i < 3; ++i) n_clusters[i] = fc.n_clusters[i];
  if (max_name) *max_name = fc.max_name;
  int i = 0;
  for (auto& cs : fc.cluster_set) {
    if (i >= cap) break;
    if (cluster_name) cluster_name[i] = cs.first;
    if (cluster_type) cluster_type[i] = cs.second.type;
    if (cluster_npts) cluster_npts[i] = cs.second.npts;
    if (cluster_nvox) cluster_nvox[i] = (int)cs.second.occupy_voxels.size();
    ++i;
  }
  return (int)fc.cluster_set.size();
}

extern "C" int scvod_host_segment(const scvod_params* p, int V, const int32_t* vox_cnt, const int32_t* vox_root, const int32_t* vox_nbr,
                                  const float* vox_bbox, int n_events, const int32_t* ev_cid, int n_edges, const int32_t* edges,
                                  int32_t* name_stage0, int32_t* name_stage1, int32_t* name_stage2, int32_t n_clusters[3], int cap,
                                  int32_t* cluster_name, int32_t* cluster_type, int32_t* max_name) {
  return scvod_host_segment_pts(p, V, vox_cnt, vox_root, vox_nbr, vox_bbox, n_events, ev_cid, n_edges, edges, 0, nullptr, nullptr, 0, nullptr,
                                nullptr, nullptr, name_stage0, name_stage1, name_stage2, nullptr, n_clusters, cap, cluster_name, cluster_type,
                                nullptr, nullptr, max_name);
}
