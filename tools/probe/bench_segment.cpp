// Times the product's host-side cluster bookkeeping (csrc/host_cluster.cpp segment_and_recognize, fed with device-style names) on the tables
// of one synthetic 64x1800 scan written by tools/host_segment_bench.py; prints us per scan and a hash of the resulting clusters
// (unchanged hash = unchanged results).  Build with -DSEG_PROF to get the phase breakdown.
#include "host_cluster.h"
#include <chrono>
#include <cstdio>
#include <vector>
using namespace scvod;
extern "C" void scvod_params_semantickitti(scvod_params* p);
#ifdef SEG_PROF
extern "C" void seg_prof_dump();
#endif
int main(int argc, char** argv) {
  FILE* f = fopen(argc > 2 ? argv[2] : "/tmp/scvod_segment_tables.bin", "rb");
  if (!f) { fprintf(stderr, "tables file missing: run tools/host_segment_bench.py\n"); return 2; }
  int hdr[5];
  size_t got = fread(hdr, 4, 5, f);
  int V = hdr[0], NE = hdr[1], NG = hdr[2], maxn = hdr[3], nnf = hdr[4];
  std::vector<int> cnt(V), root(V), nbr(V * 27), ev(NE), edges(NG * 2), name(V), nf(nnf);
  std::vector<float> bbox(V * 6);
  got += fread(cnt.data(), 4, V, f) + fread(root.data(), 4, V, f) + fread(nbr.data(), 4, V * 27, f) + fread(bbox.data(), 4, V * 6, f);
  got += fread(ev.data(), 4, NE, f) + fread(edges.data(), 4, NG * 2, f) + fread(name.data(), 4, V, f) + fread(nf.data(), 4, nnf, f);
  if (got != (size_t)(5 + 2 * V + 27 * V + 6 * V + NE + 2 * NG + V + nnf)) { fprintf(stderr, "short tables file\n"); return 2; }
  scvod_params p; scvod_params_semantickitti(&p);
  ScanTables t; t.V = V; t.n_events = NE; t.n_edges = NG; t.vox_cnt = cnt.data(); t.vox_root = root.data(); t.vox_nbr = nbr.data();
  t.vox_bbox = bbox.data(); t.ev_cid = ev.data(); t.edges = edges.data(); t.vox_name = name.data(); t.name_first = nf.data(); t.max_name = maxn;
  int reps = argc > 1 ? atoi(argv[1]) : 20000;
  FrameClusters out;
  bool ok = segment_and_recognize(p, t, out, false);
  unsigned long long h = 1469598103934665603ull;
  for (auto& c : out.cluster_set) { h = (h ^ (unsigned)c.first) * 1099511628211ull; h = (h ^ (unsigned)c.second.type) * 1099511628211ull; h = (h ^ (unsigned)c.second.occupy_voxels.size()) * 1099511628211ull; for (int v : c.second.occupy_voxels) h = (h ^ (unsigned)v) * 1099511628211ull; }
  for (int v : out.vox_label) h = (h ^ (unsigned)v) * 1099511628211ull;
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r) { FrameClusters o2; ok &= segment_and_recognize(p, t, o2, false); }
  double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
#ifdef SEG_PROF
  seg_prof_dump();
#endif
  printf("ok=%d clusters=%zu hash=%llx  %.2f us per scan\n", (int)ok, out.cluster_set.size(), h, us);
}
