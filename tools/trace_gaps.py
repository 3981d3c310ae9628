#!/usr/bin/env python
"""GPU timeline statistics from a chrome trace written by `bench.py --quick --trace FILE` (torch.profiler / CUPTI):
how long the GPU ran 0, 1, 2, ... kernels at once, per-kernel totals, and the per-stream queueing delay between consecutive kernels.
usage: trace_gaps.py trace.json"""
import json, sys, collections
ev = json.load(open(sys.argv[1]))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
t0 = min(e["ts"] for e in ks); t1 = max(e["ts"] + e["dur"] for e in ks)
print(f"{len(ks)} GPU activities over {(t1 - t0) / 1000:.2f} ms")
pts = []
for e in ks:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
hist = collections.Counter(); cur = 0; last = t0
for t, d in pts:
    hist[cur] += t - last; last = t; cur += d
tot = sum(hist.values())
print("concurrency histogram (share of wall time with k activities in flight):")
for k in sorted(hist):
    if hist[k] / tot > 0.002:
        print(f"  {k:3d}: {hist[k] / tot * 100:5.1f}%")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ks:
    n = e["name"].split("(")[0].replace("void ", "").replace("scvod::", "")[:40]
    agg[n][0] += 1; agg[n][1] += e["dur"]
print("per kernel: launches, total ms, avg us, sum/wall")
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"  {n:40s} {c:6d} {d / 1000:9.3f} {d / c:9.1f} {d / (t1 - t0):6.2f}")
print(f"sum of activity durations / wall = {sum(e['dur'] for e in ks) / (t1 - t0):.2f}")
