#!/usr/bin/env python
"""bench.py — scans/s of the full SCV-OD dynamic-removal path (BASELINE.json metric) on N B200s.

One "step" = one pass of the hot path over one batch of synthetic 64x1800 scans per GPU: W independent sequence chunks
of S scans each (W workers per GPU, one chunk per worker and step; default_workers(): 24 x 64 = 1536 scans with >= 16 host cores per
GPU, 16 x 64 below), every chunk going through
PatchWork ground fit -> curved-voxel binning -> occupancy descriptor -> clustering/classification ->
tracking diff over the chunk -> per-point classes -> static submap (+ one NCCL all-gather of the per-GPU
submaps per chunk when N > 1).  Scans shard across ranks (one process per GPU, independent chunks, weak scaling).

  value : inputs already resident in HBM (scvod_push_scans_dev), labels stay on the device
  e2e   : same work through the host-buffer C-ABI calls a reference maintainer would bind
          (scvod_push_scans from pinned host memory, labels copied back to the host) — copies inside
          the timed region (16 workers already keep the copy engine 73 % busy; double buffering inside a
          worker with scvod_prefetch_scans, --prefetch, was measured slower: 18.6k vs 21.7k scans/s)
  --impl reference : the reference's CPU path (the oracle restatement; the reference itself cannot be
          built in this image) on all host threads, bounded sample per step, rank 0 only.
"""
import argparse
import ctypes
import os

# Independent sequence chunks run on up to 16 streams per GPU; with the default of 8 hardware work queues, streams that share a
# queue serialise behind each other's bulk copies.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

def env_int_early(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


METRIC = "scans/sec (64-beam ~120k pts) full SCV-OD removal"
UNIT = "scans/s"
WORKLOAD = ("configs[1]: SemanticKITTI-shape streams (config/semantickitti.yaml), full PatchWork+SSC+tracking labelling; a step = W independent "
            "64-scan sequences (one per worker), each tracked as one unbroken chain")
RINGS, COLS = env_int_early("SCVOD_BENCH_RINGS", 64), env_int_early("SCVOD_BENCH_COLS", 1800)  # tuning only: the metric is quoted on 64 x 1800
SEED = 0x5C0D0000
# algorithmic (compulsory) bytes per unit for the stage each kernel dominates
ALGO_BYTES = {
    "k_patch_chain": ("ground stage (P1-P6): 32 B per input point (16N read + 16N written in reference order)", 32.0, "points"),
    "k_patch_sort_4k": ("ground stage (P1-P6): 32 B per input point", 32.0, "points"),
    "k_patch_sort_1k": ("ground stage (P1-P6): 32 B per input point", 32.0, "points"),
    "k_name_replay": ("cluster stage (C1-C2): 8 B per voxel read + 4 B per voxel + 4 B per apri point written", 16.0, "voxels_plus_apri"),
    "k_patch_assign": ("ground stage (P1-P6): 32 B per input point", 32.0, "points"),
    "k_patch_scatter": ("ground stage (P1-P6): 32 B per input point", 32.0, "points"),
    "k_emit": ("ground stage (P1-P6): 32 B per input point", 32.0, "points"),
    "k_vox_stats": ("descriptor stage (B3): 12 B per apri point", 12.0, "apri"),
    "k_track": ("diff stage (D1-D2): 16 B read + 16 B written (carried cloud) + 4 B hit per re-binned cluster point", 36.0, "track_points"),
}


def host_cores():
    """Cores this process may run on (cgroup / taskset aware), not the machine total."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bind_to_gpu_numa_node(local_rank, local_world):
    """One process per GPU: run this rank (its worker threads and, through first touch, its pinned staging buffers) on the host
    cores of its GPU's NUMA node (all of them, shared with the other ranks of that node).  Returns a short description."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
        bus = {}
        for ln in out.strip().splitlines():
            idx, b = [t.strip() for t in ln.split(",")]
            bus[int(idx)] = b.lower()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = [int(t) for t in vis.split(",")] if vis and all(t.strip().isdigit() for t in vis.split(",")) else list(range(len(bus)))

        def node_of(lr):
            b = bus[phys[lr]]  # 00000000:1b:00.0 -> sysfs name 0000:1b:00.0
            name = b[-12:] if len(b) > 12 else b
            return int(open(f"/sys/bus/pci/devices/{name}/numa_node").read())

        nodes = [node_of(lr) for lr in range(local_world)]
        node = nodes[local_rank]
        if node < 0 or len(set(nodes)) < 2:
            # no NUMA information (containers often report -1) or a single node: leave the threads to the scheduler.  Confining each rank
            # to its 1/N share of the cores was measured on the 8-GPU box (32 cores, node -1): 158k vs 178k+ scans/s unconfined
            return f"not pinned (numa nodes of the local GPUs: {sorted(set(nodes))})"
        allowed = sorted(os.sched_getaffinity(0))
        cores = allowed
        if node >= 0:
            node_cores = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                node_cores.update(range(int(a), int(b or a) + 1))
            cores = [c for c in allowed if c in node_cores] or allowed
        # the whole node, shared by the ranks whose GPUs sit on it: a strict 1/N split of the cores was measured slower (see above)
        mine = cores
        os.sched_setaffinity(0, mine)
        return f"numa node {node}: cores {mine[0]}-{mine[-1]} ({len(mine)} of {len(allowed)} allowed), shared by the node's ranks"
    except Exception as e:  # noqa: BLE001
        return f"not pinned ({type(e).__name__}: {e})"


def default_workers(cores_total, local_world):
    """Independent sequence chunks processed side by side per GPU (one context + host thread each).  Measured on one B200: with 16
    host cores per GPU 16 / 24 / 32 workers give 36.4k / 37.5k-38.0k / 38.9k scans/s device-resident and 25.9k / 27.4k end to end
    (16 -> 24); with 4 cores per GPU 4 / 8 / 12 / 16 workers give 18.6k / 22.2k / 24.5k / 26.5k.  Both arms of the bench use this."""
    per_gpu = cores_total // max(1, local_world)
    if per_gpu >= 16:
        return 24
    return max(4, min(16, 4 * per_gpu))


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def gen_scans(pkg, first_id, count, seed):
    def one(k):
        return pkg.synth_scan(seed, k, RINGS, COLS)

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        res = list(ex.map(one, range(first_id, first_id + count)))
    return [r[0] for r in res], np.stack([r[1] for r in res])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._reader, daemon=True).start()
        except Exception:
            self.proc = None

    def _reader(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        deadline = time.time() + 1.0
        while not self.lines and time.time() < deadline:  # a very short timed region: wait for the first sample
            time.sleep(0.02)
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel`, averaged over its launches in the
    committed `ncu --set full` capture of the same 64-scan workload (profiles/r02_ncu_full_summary.csv); None if absent."""
    import csv

    path = os.path.join(ROOT, "profiles", "r02_ncu_full_summary.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:]
                if r[ik].replace("void ", "").split("(")[0].split("<")[0].strip() == kernel]
        return sum(vals) / len(vals) if vals else None
    except Exception:
        return None


def flat_pool(batches_np):
    """Concatenate the pool's chunks: (points float32 [n,4], offsets int64 [chunks*S+1], poses [chunks*S,6])."""
    scans = [s for b in batches_np for s in b[0]]
    off = np.zeros(len(scans) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in scans])
    return np.ascontiguousarray(np.concatenate(scans, axis=0), np.float32), off, np.concatenate([b[1] for b in batches_np], axis=0)


def oracle_handle(params):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest

    return conftest.Oracle(params)


def cpu_chunks(params, pool, S, nchunks, nthreads):
    """The oracle (kind "port") in the GPU arm's decomposition: nchunks independent S-scan sequences, one host thread per chunk
    at a time (per-scan stages + the chunk's own serial tracking chain).  Returns scans/s."""
    flat, off, poses = pool
    orc = oracle_handle(params)
    secs, _ = orc.run_chunks(flat, off, poses, S, nchunks, nthreads=nthreads)
    orc.close()
    return nchunks * S / secs, secs


def cpu_single_chain(params, scans, poses, nthreads, max_scans):
    """One sequence the way segDF runs it (ssc.cpp:1435-1452): per-scan stages (spread over nthreads), then ONE serial tracking chain."""
    orc = oracle_handle(params)
    n = min(len(scans), max_scans)
    secs, _, _ = orc.run_sequence(scans[:n], poses[:n], nthreads=nthreads, want_labels=False)
    orc.close()
    return n / secs, n, secs


def run_reference(args):
    """CPU arm: the reference's path (oracle port; the reference needs ROS/PCL/Eigen and cannot be built here) on the host cores,
    in the SAME decomposition as the GPU arm: W independent chunks of S scans per step.  Loads libscvod_synth.so (generator +
    parameters) and the oracle only; the CUDA library is never mapped."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    pkg = entry._load_package()
    params = pkg.semantickitti_params()
    ncores = host_cores()
    S = args.scans_per_step
    W = args.workers if args.workers > 0 else default_workers(ncores, max(1, env_int("LOCAL_WORLD_SIZE", args.gpus)))  # the GPU arm's step
    # bounded sample: a step processes min(W, cores) chunks (each core runs one chunk from start to end), scaled to W chunks
    nchunks = max(1, min(W, ncores))
    pool = flat_pool([gen_scans(pkg, b * 1000, S, SEED) for b in range(min(args.pool, nchunks))])
    for _ in range(1 if args.warmup else 0):  # one untimed pass pages the code and the pool in; the CPU arm has no other warm-up state
        cpu_chunks(params, pool, S, nchunks, ncores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_chunks(params, pool, S, nchunks, ncores)
    dt = time.perf_counter() - t0
    value = nchunks * S * args.steps / dt
    chain_val, chain_n, chain_secs = cpu_single_chain(params, *gen_scans(pkg, 0, min(S, 2 * ncores), SEED), ncores, S)
    sample = (f"{nchunks} independent chunks x {S} scans per step x {args.steps} steps on {min(nchunks, ncores)} threads (one chunk per thread: per-scan "
              f"stages + its own tracking chain), i.e. {nchunks}/{W} of the GPU arm's step; oracle/scvod_oracle.cpp")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps * (W / nchunks), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "scans_per_step": W * S, "chunks_per_step": W, "scans_per_chunk": S,
                   "points_per_scan": float((pool[1][-1] - pool[1][0]) / (len(pool[1]) - 1)), "rings": RINGS, "cols": COLS,
                   "sampled_chunks_per_step": nchunks},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(nchunks, ncores), "kind": "port", "sample": sample,
                         "single_chain": {"value": chain_val, "unit": UNIT, "cores": ncores,
                                          "sample": f"ONE {chain_n}-scan sequence in {chain_secs:.2f} s: per-scan stages on {ncores} threads, then the serial "
                                                    "tracking chain (how segDF runs a sequence)"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="stream", choices=["stream", "sequence", "parkinglot_gicp", "stress"],
                    help="stream = BASELINE configs[1] (default, the headline); sequence = configs[3]; parkinglot_gicp = configs[2]; stress = configs[4] "
                         "(tools/bench_workloads.py)")
    ap.add_argument("--quick", action="store_true", help="tuning runs: skip the e2e, roofline, cpu_baseline and parity legs; prints value only")
    ap.add_argument("--concurrent-timing", action="store_true",
                    help="tuning runs (with --quick): per-kernel CUDA-event durations measured INSIDE the multi-worker run (kernels of different "
                         "workers overlap, so the durations are not additive; a kernel stretched against its single-stream time is contended)")
    ap.add_argument("--with-e2e", action="store_true", help="tuning runs (with --quick): also time the host-buffer (e2e) leg")
    ap.add_argument("--trace", default="", help="tuning runs (with --quick): after the timed region, record a CUPTI kernel timeline of 2 more steps "
                                                 "with torch.profiler and write it to this chrome-trace file (tools/trace_gaps.py reads it)")
    ap.add_argument("--skip-tracking", action="store_true", help="ablation: per-scan stages only (INVALID as a bench number)")
    ap.add_argument("--seq-scans", type=int, default=1000, help="--workload sequence: scans in the sequence")
    ap.add_argument("--scans-per-step", type=int, default=64, help="scans per sequence chunk; a step is one chunk per worker")
    ap.add_argument("--pool", type=int, default=3, help="distinct input batches rotated through (pool > L2)")
    ap.add_argument("--workers", type=int, default=0, help="independent sequence chunks processed side by side per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="scans for the cpu_baseline leg (0 = auto)")
    ap.add_argument("--gather", default="exact", choices=["exact", "padded", "off"],
                    help="N > 1: static-submap all-gather per chunk: rows = largest count of the call (default), capacity-sized (round-1 behaviour), "
                         "or none (ablation, INVALID as a bench number)")
    ap.add_argument("--submap", default="instance", choices=["instance", "all"],
                    help="what a chunk contributes to the merged map: the points of its non-dynamic clusters (reference saveSegCloud mode 3, default) or "
                         "every non-dynamic input point")
    ap.add_argument("--no-pin", action="store_true", help="N > 1: do not bind the rank to the host cores of its GPU's NUMA node")
    ap.add_argument("--prefetch", action="store_true",
                    help="e2e leg: scvod_prefetch_scans the worker's next chunk during the tracking chain (measured slower with 16 workers: "
                         "a saturated PCIe link delays the launches of every tracking chain)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "stream":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_workloads

        return bench_workloads.run(args, sys.modules[__name__])

    import torch
    import torch.distributed as dist

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SCV-OD path has no CPU fallback")
    cores_total = host_cores()  # before the rank binds itself to its share of them
    pin_note = bind_to_gpu_numa_node(local_rank, env_int("LOCAL_WORLD_SIZE", world)) if (world > 1 and not args.no_pin) else "not pinned"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    pkg = entry._load_package()
    params = pkg.semantickitti_params()
    S = args.scans_per_step
    local_world = env_int("LOCAL_WORLD_SIZE", world)
    # workers per GPU: 16 when there are at least 4 host cores per GPU, else 4 per core.  A worker that waits for the GPU sleeps
    # (blocking-sync events) or yields (tracking poll), so workers share cores: measured on one B200 restricted to 4 cores,
    # 4 / 8 / 12 / 16 workers give 18.6k / 22.2k / 24.5k / 26.5k scans/s
    W = args.workers if args.workers > 0 else default_workers(cores_total, local_world)
    # each rank owns its own sequence chunks (scan-sharding, no data-path collective before the submap merge)
    batches = []
    for b in range(args.pool):
        first = (rank * args.pool + b) * 1000
        scans, poses = gen_scans(pkg, first, S, SEED + rank)
        off = np.zeros(S + 1, np.int64)
        off[1:] = np.cumsum([len(s) for s in scans])
        flat = torch.from_numpy(np.concatenate(scans, axis=0)).pin_memory()
        batches.append({"scans": scans, "poses": poses, "off": off, "host": flat, "dev": flat.to(dev), "npts": int(off[-1])})
    max_pts = max(b["npts"] for b in batches)

    # W independent workers per GPU: each owns a context, a CUDA stream and its sequence chunks.  The tracking
    # chain of a chunk is a latency-bound host<->device ping-pong; running several chunks side by side keeps
    # the GPU busy (the same sharding that spreads chunks over GPUs, applied inside one GPU).
    class Worker:
        def __init__(self, wid):
            self.wid = wid
            self.stream = torch.cuda.Stream(device=dev)
            self.ssc = pkg.SSC(params, device=local_rank, max_points=RINGS * COLS, max_batch=S)
            self.ssc.set_option("inspect", 0)
            self.ssc.set_option("host_threads", max(1, cores_total // (W * max(1, local_world))))
            self.ssc.set_option("submap_all_static", 1 if args.submap == "all" else 0)
            self.ssc.set_stream(self.stream.cuda_stream)
            self.labels_host = torch.empty(max_pts, dtype=torch.uint8).pin_memory()
            # two send buffers per worker: the all-gather of chunk i runs while the worker already fills the other one for chunk i+1
            self.submaps = [torch.empty((max_pts, 4), dtype=torch.float32, device=dev) for _ in range(2 if world > 1 else 1)]
            # per send buffer: "its gather has been issued" (set by the comm thread) and the CUDA event recorded right after that gather;
            # the worker itself waits for both before it overwrites the buffer, so the comm thread never has to come back to it
            self.submap_issued = [threading.Event() for _ in self.submaps]
            self.submap_done = [None for _ in self.submaps]
            for ev in self.submap_issued:
                ev.set()
            self.count = 0

    workers = [Worker(w) for w in range(W)]
    par = entry._load_parallel()
    gatherer = par.SubmapGatherer(max_pts, dev) if world > 1 else None  # one NCCL all-gather of the static submaps per chunk
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    gather_count = [0]

    def step(wk, i, host_io, prefetch_next=False):
        # every worker cycles through the whole pool (so its buffers reach their steady-state sizes during warm-up) and
        # neighbouring workers are on different batches at any time
        wk.count += 1
        b = batches[(wk.wid + wk.count) % len(batches)]
        ssc = wk.ssc
        ssc.reset()
        if host_io:
            ssc.process_host_ptr(b["host"].data_ptr(), b["off"])  # finds its points on the device if they were prefetched
            if prefetch_next and args.prefetch:  # double buffering: the upload of this worker's next chunk overlaps the tracking chain of this one
                nb = batches[(wk.wid + wk.count + 1) % len(batches)]
                ssc.prefetch_host_ptr(nb["host"].data_ptr(), nb["off"])
        else:
            ssc.process_device(b["dev"].data_ptr(), b["off"])
        if not args.skip_tracking:
            ssc.tracking(b["poses"])
        if host_io:
            ssc.labels_into(0, S, wk.labels_host.data_ptr(), wk.labels_host.numel())
        else:
            ssc.refresh_labels(0, S)
        k = wk.count % len(wk.submaps)
        if world > 1:
            wk.submap_issued[k].wait()  # the gather of the submap this buffer held two chunks ago has been issued ...
            if wk.submap_done[k] is not None:
                wk.submap_done[k].synchronize()  # ... and has finished reading it
                wk.submap_done[k] = None
            wk.submap_issued[k].clear()
        return k, ssc.static_submap_device(0, S, b["poses"], wk.submaps[k].data_ptr(), max_pts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(first_step, nsteps, host_io):
        """nsteps steps spread over the workers; the per-GPU static submaps are merged by ONE NCCL all-gather
        per step, issued in step order by a single communication thread (same order on every rank)."""
        done = {}
        cv = threading.Condition()
        errors = []

        def work(wk):
            try:
                with torch.cuda.stream(wk.stream):
                    mine = list(range(first_step + wk.wid, first_step + nsteps, W))
                    for i in mine:
                        k, n_static = step(wk, i, host_io, prefetch_next=(i != mine[-1]))
                        with cv:
                            done[i] = (wk, k, n_static)
                            cv.notify_all()
            except Exception as e:  # noqa: BLE001
                with cv:
                    errors.append(e)
                    cv.notify_all()

        def comm():
            # One communication thread issues the gathers in chunk order (the same order on every rank) on its own stream and never
            # waits for a gather to finish: the worker that owns a send buffer waits for that buffer's event before reusing it.
            with torch.cuda.stream(comm_stream):
                for i in range(first_step, first_step + nsteps):
                    with cv:
                        while i not in done and not errors:
                            cv.wait()
                        if errors:
                            break
                        wk, k, n_static = done.pop(i)
                    if args.gather != "off":
                        gatherer.gather(wk.submaps[k], n_static, padded=(args.gather == "padded"))
                        gather_count[0] += 1
                    ev = torch.cuda.Event()
                    ev.record(comm_stream)
                    wk.submap_done[k] = ev
                    wk.submap_issued[k].set()
                comm_stream.synchronize()
            for wk in workers:  # an error elsewhere must not leave a worker waiting for a gather that will never be issued
                for ev in wk.submap_issued:
                    ev.set()

        threads = [threading.Thread(target=work, args=(wk,)) for wk in workers]
        if world > 1:
            threads.append(threading.Thread(target=comm))
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    def timed(host_io, with_kernel_timing):
        run_steps(0, W * max(args.warmup, len(batches)), host_io)  # every worker sees every pool batch once: buffers reach steady state
        barrier()
        if with_kernel_timing:
            pkg.kernel_timing(True)
        if os.environ.get("SCVOD_PROFILE"):
            workers[0].ssc.set_option("profile_reset", 1)  # host-side section timers: steady state only
        l0 = sum(wk.ssc.kernel_launches for wk in workers)
        a0 = workers[0].ssc.stat("reallocs")
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        c0 = time.process_time()
        run_steps(1000 * W, args.steps * W, host_io)
        for wk in workers:
            torch.cuda.current_stream().wait_stream(wk.stream)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        cpu_busy[0] = (time.process_time() - c0) / max(wall, 1e-9)  # host cores this rank kept busy (user + system), polling included
        ev = e0.elapsed_time(e1) / 1000.0
        clocks = sampler.stop() if rank == 0 else None
        rep = pkg.kernel_timing_report() if with_kernel_timing else None
        if with_kernel_timing:
            pkg.kernel_timing(False)
        secs = max(ev, wall)  # every step ends with host-side bookkeeping, so wall >= device time
        reallocs[0] += workers[0].ssc.stat("reallocs") - a0  # buffer (re)allocations inside the timed regions: 0 in steady state
        return par.max_over_ranks(secs, dev), sum(wk.ssc.kernel_launches for wk in workers) - l0, clocks, rep

    reallocs = [0]
    cpu_busy = [0.0]
    secs_dev, launches, clocks, crep = timed(False, args.concurrent_timing)
    if args.quick:
        if args.trace and rank == 0:
            from torch.profiler import ProfilerActivity, profile

            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                run_steps(5000 * W, 2 * W, False)
                torch.cuda.synchronize()
            prof.export_chrome_trace(args.trace)
        e2e_quick = None
        if args.with_e2e:
            secs_q, _, _, _ = timed(True, False)
            e2e_quick = world * W * S * args.steps / secs_q
        if rank == 0:
            if crep:
                for k, (ms, cnt) in sorted(crep.items(), key=lambda kv: -kv[1][0]):
                    print(f"  {k:26s} {ms / (W * args.steps):8.4f} ms/chunk under concurrency ({cnt // (W * args.steps)} launches/chunk)", file=sys.stderr)
            print(json.dumps({"quick": True, "value": world * W * S * args.steps / secs_dev, "unit": UNIT, "workers": W, "ms_per_step": 1000.0 * secs_dev / args.steps,
                              "e2e": e2e_quick, "skip_tracking": args.skip_tracking, "gpu_launches": int(launches), "reallocs_in_timed_region": int(reallocs[0]), "host_cores_busy": round(cpu_busy[0], 2), "host_cores": host_cores(), "n_gpus": world, "gather": args.gather if world > 1 else None,
                              "submap": args.submap, "host_binding": pin_note,
                              "gather_mb_per_chunk": (gatherer.bytes_moved / max(1, gather_count[0]) / 1e6) if gatherer else None}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        for wk in workers:
            wk.ssc.close()
        return 0
    secs_e2e, _, clocks_e2e, _ = timed(True, False)
    # roofline block: per-kernel CUDA-event durations need launches that do not overlap with other streams, so
    # they are measured in a dedicated pass of the same steps on ONE worker (events on that worker's stream)
    rep = None
    if rank == 0:
        pkg.kernel_timing(True)
        with torch.cuda.stream(workers[0].stream):
            for i in range(args.steps):
                k, _ = step(workers[0], 2000 + i, False)
                if world > 1:
                    workers[0].submap_issued[k].set()
        torch.cuda.synchronize()
        rep = pkg.kernel_timing_report()
        pkg.kernel_timing(False)
    if world > 1:
        dist.barrier()

    if rank == 0:
        value = world * W * S * args.steps / secs_dev
        e2e_value = world * W * S * args.steps / secs_e2e
        gather_calls = gather_count[0]
        avg_pts = float(np.mean([b["npts"] for b in batches])) / S
        peak, peak_src = measured_peak()
        # dominant kernel by total device time inside the timed region
        dom = max(rep.items(), key=lambda kv: kv[1][0])
        dom_name, (dom_ms, dom_cnt) = dom
        key = dom_name
        desc, bytes_per_unit, unit_kind = ALGO_BYTES.get(key, ("ground stage, 32 B per input point", 32.0, "points"))
        w0 = workers[0].ssc
        per_kind = {"points": w0.stat("points"), "apri": w0.stat("apri_points"), "track_points": w0.stat("track_points"),
                    "voxels_plus_apri": (12 * w0.stat("voxels") + 4 * w0.stat("apri_points")) / 16.0}
        steps_seen = max(1, w0.stat("scans") // S)
        units_per_launch = per_kind[unit_kind] / steps_seen * (args.steps / dom_cnt)  # units handled by one launch, on average
        achieved = bytes_per_unit * units_per_launch / (dom_ms / dom_cnt * 1e-3) / 1e9
        total_kernel_ms = sum(v[0] for v in rep.values())
        roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(dom_name), "traffic_source": "profiles/r02_ncu_full_summary.csv (ncu --set full, bytes per launch)",
                    "peak_source": peak_src, "algorithmic_bytes": desc, "avg_launch_ms": dom_ms / dom_cnt,
                    "kernel_share_of_gpu_time": dom_ms / total_kernel_ms,
                    "measured": f"CUDA events on the launching stream, dedicated single-worker pass of {args.steps} chunks inside bench.py",
                    "kernels_ms_per_chunk": {k: round(v[0] / args.steps, 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])}}
        ncores = host_cores()
        b0 = batches[0]
        # cpu_baseline: the oracle in the SAME decomposition (independent S-scan chunks, one host thread per chunk), bounded to
        # min(W, cores, pool) chunks; its labels double as the parity check of the timed workload (outside the timed region)
        in_chunks = len(batches)
        nchunks = args.cpu_sample or max(1, min(W, ncores))
        pool_np = flat_pool([(b["scans"], b["poses"]) for b in batches])
        orc = oracle_handle(params)
        cpu_secs, olab = orc.run_chunks(*pool_np, S, nchunks, nthreads=ncores, want_labels=True)
        cpu_val = nchunks * S / cpu_secs
        # the reference itself is single-threaded (OpenMP only inside transformCloud): the faithful figure, on a smaller sample
        cpu1_secs, _, _ = orc.run_sequence(b0["scans"][:6], b0["poses"][:6], nthreads=1, want_labels=False)
        cpu1_n, cpu1_val = 6, 6 / cpu1_secs
        orc.close()
        checked = min(in_chunks, nchunks)
        mism, npts_checked, frames_bad = 0, 0, 0
        wk = workers[0]
        pos = 0
        for bi in range(in_chunks):
            b = batches[bi]
            if bi < checked:
                with torch.cuda.stream(wk.stream):
                    wk.ssc.reset()
                    wk.ssc.process_host_ptr(b["host"].data_ptr(), b["off"])
                    wk.ssc.tracking(b["poses"])
                    wk.ssc.labels_into(0, S, wk.labels_host.data_ptr(), wk.labels_host.numel())
                got = wk.labels_host.numpy()[: b["npts"]]
                exp = olab[pos: pos + b["npts"]]
                diff = got != exp
                mism += int(diff.sum())
                npts_checked += b["npts"]
                frames_bad += sum(1 for f in range(S) if diff[b["off"][f]:b["off"][f + 1]].any())
            pos += b["npts"]
        parity = {"checked": f"per-point classes of {checked} of the timed 64-scan chunks (every frame, through scvod_push_scans + scvod_track) against "
                             "the oracle on the same inputs, outside the timed region", "points": npts_checked, "mismatching_points": mism,
                  "mismatching_frames": frames_bad, "bit_exact": mism == 0,
                  "note": "oracle = restatement of the reference (parity unpinned: the reference ships no vectors and cannot be built here)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * secs_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "scans_per_step": W * S, "chunks_per_step": W, "scans_per_chunk": S, "points_per_scan": avg_pts, "rings": RINGS, "cols": COLS,
                       "l2": f"inputs rotate over a pool of {args.pool} batches ({args.pool * b0['npts'] * 16 / 1e6:.0f} MB) larger than the 126 MB L2; "
                             "per-step workspace (>1 GB) is rewritten every step",
                       "workers_per_gpu": W,
                       "sharding": ("scan-sharded: independent sequence chunks per rank and per worker; one NCCL all-gather of static submaps per chunk "
                                    f"(--gather {args.gather}: {gatherer.bytes_moved / max(1, gather_calls) / 1e6:.1f} MB received per rank and chunk)"
                                    if world > 1 else "single GPU; independent sequence chunks per worker"),
                       "submap": ("instance map: points of the non-dynamic clusters (reference saveSegCloud mode 3)" if args.submap == "instance"
                                  else "every non-dynamic input point"),
                       "host_binding": pin_note},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(W * (b0["npts"] * 16 + (S + 1) * 8)), "d2h_bytes_per_step": int(W * b0["npts"]),
                    "ms_per_step": 1000.0 * secs_e2e / args.steps},
            "gpu_launches": int(launches),
            "reallocs_in_timed_region": int(reallocs[0]),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": min(nchunks, ncores), "kind": "port",
                             "sample": f"{nchunks} independent {S}-scan chunks of the same workload in {cpu_secs:.2f} s, one host thread per chunk (per-scan "
                                       "stages + the chunk's own tracking chain); oracle/scvod_oracle.cpp (reference not buildable here)",
                             "single_thread": {"value": cpu1_val, "unit": UNIT, "cores": 1,
                                               "sample": f"{cpu1_n} scans in {cpu1_secs:.2f} s (how the reference itself runs: one thread)"}},
            "parity": parity,
            "clocks": clocks,
            "clocks_e2e": clocks_e2e,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for wk in workers:
        wk.ssc.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
