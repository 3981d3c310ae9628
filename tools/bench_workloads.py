"""The other BASELINE.json configurations behind `bench.py --workload ...` (the default workload, configs[1], lives in bench.py).

  sequence         configs[3]: ONE 1000-scan sequence, scan-sharded over ranks x workers in 64-scan chunks.  Two modes are timed:
                   "exact" keeps the reference's single tracking chain (ssc.cpp:1450-1452) with the tail hand-off between chunk
                   owners (scvod_export_tail / scvod_track_from_tail; ChainLink between ranks), "cut" treats every chunk as its own
                   sequence.  The label difference between the two is reported, and the exact chain is checked against the oracle
                   across the first cut.  One NCCL all-gather of the per-rank static submaps per step.
  parkinglot_gicp  configs[2]: dense (128 x 2700) scans with config/parkinglot.yaml through the per-scan stages + one GICP
                   scan-to-map alignment each; pose compared with the CPU oracle (docs/gicp_spec.md; self-consistency only).
  stress           configs[4]: curved-voxel binning + radius-kNN (normals) kernels only on aggregated dense clouds, algorithmic
                   GB/s against the HBM peak; replicas only (every rank runs its own copy, no exchange).

Every function prints ONE JSON line on rank 0 with the keys of the bench contract.
"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RINGS, COLS = 64, 1800
SEED = 0x5C0D0000


def _oracle(params):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest

    return conftest.Oracle(params)


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class Env:
    def __init__(self, args, bench):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args, self.bench = torch, dist, args, bench
        self.rank, self.world, self.local_rank = bench.env_int("RANK", 0), bench.env_int("WORLD_SIZE", 1), bench.env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the SCV-OD path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.dev)
        self.pkg = bench.entry._load_package()
        self.par = bench.entry._load_parallel()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()

    def sum_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return t.tolist()


# ------------------------------------------------------------------------------------------------------------------
# configs[3]: one long sequence
# ------------------------------------------------------------------------------------------------------------------
def run_sequence(args, bench):
    env = Env(args, bench)
    torch, pkg, par = env.torch, env.pkg, env.par
    params = pkg.semantickitti_params()
    T, S = args.seq_scans, args.scans_per_step
    mine = par.shard_chunks(T, S, env.world, env.rank)  # consecutive chunks: rank order = sequence order
    chunks = []
    for (a, b) in mine:
        scans, poses = bench.gen_scans(pkg, a, b - a, SEED)  # scan ids a..b-1 of ONE generator sequence
        off = np.zeros(b - a + 1, np.int64)
        off[1:] = np.cumsum([len(s) for s in scans])
        flat = torch.from_numpy(np.concatenate(scans, axis=0)).pin_memory()
        chunks.append({"a": a, "b": b, "scans": scans, "poses": poses, "off": off, "host": flat, "dev": flat.to(env.dev), "npts": int(off[-1]),
                       "labels": torch.empty(int(off[-1]), dtype=torch.uint8).pin_memory()})
    npts_rank = sum(c["npts"] for c in chunks)
    ctxs = []
    for c in chunks:
        s = pkg.SSC(params, device=env.local_rank, max_points=RINGS * COLS, max_batch=S)
        s.set_option("inspect", 0)
        st = torch.cuda.Stream(device=env.dev)
        s.set_stream(st.cuda_stream)
        ctxs.append((s, st))
    cap = int(max(env.par.max_over_ranks(float(npts_rank), env.dev), 1))  # the padded all-gather needs one capacity for all ranks
    submap = torch.empty((cap, 4), dtype=torch.float32, device=env.dev)
    gatherer = par.SubmapGatherer(cap, env.dev) if env.world > 1 else None
    link = par.ChainLink(env.dev) if env.world > 1 else None

    def push(i, host_io):
        s, st = ctxs[i]
        c = chunks[i]
        with torch.cuda.stream(st):
            s.reset()
            if host_io:
                s.process_host_ptr(c["host"].data_ptr(), c["off"])
            else:
                s.process_device(c["dev"].data_ptr(), c["off"])

    def finish_chunk(i, host_io, sub_off):
        s, st = ctxs[i]
        c = chunks[i]
        with torch.cuda.stream(st):
            if host_io:
                s.labels_into(0, c["b"] - c["a"], c["labels"].data_ptr(), c["labels"].numel())
            else:
                s.refresh_labels(0, c["b"] - c["a"])
            return s.static_submap_device(0, c["b"] - c["a"], c["poses"], submap.data_ptr() + 16 * sub_off, cap - sub_off)

    def step(mode, host_io):
        n = len(ctxs)
        errs = []

        def guarded(fn, *a):
            try:
                fn(*a)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        if mode == "cut":
            def whole(i):
                push(i, host_io)
                with torch.cuda.stream(ctxs[i][1]):
                    ctxs[i][0].tracking(chunks[i]["poses"])
            th = [threading.Thread(target=guarded, args=(whole, i)) for i in range(n)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        else:
            th = [threading.Thread(target=guarded, args=(push, i, host_io)) for i in range(n)]
            for t in th:
                t.start()
            # the chain, in sequence order; the per-scan stages of later chunks run meanwhile
            for i in range(n):
                th[i].join()
                if errs:
                    break
                s, st = ctxs[i]
                with torch.cuda.stream(st):
                    if i > 0:
                        sts = s.track_from_tail(ctxs[i - 1][0].export_tail(), chunks[i - 1]["poses"][-1], chunks[i]["poses"][0])
                        ctxs[i - 1][0].apply_tail_states(sts)
                    elif env.rank > 0 and n > 0:
                        tail = link.recv_tail(env.rank - 1)
                        pose_pre = np.frombuffer(tail[-24:].tobytes(), np.float32)  # the sender appends the pose of its last frame
                        sts = s.track_from_tail(tail[:-24], pose_pre, chunks[i]["poses"][0])
                        link.send_states(sts, env.rank - 1)
                    s.tracking(chunks[i]["poses"])
            if not errs and env.rank + 1 < env.world and n > 0:
                s, st = ctxs[-1]
                with torch.cuda.stream(st):
                    tail = np.concatenate([s.export_tail(), np.frombuffer(np.ascontiguousarray(chunks[-1]["poses"][-1], np.float32).tobytes(), np.uint8)])
                link.send_tail(tail, env.rank + 1)
                s.apply_tail_states(link.recv_states(env.rank + 1))
        if errs:
            raise errs[0]
        sub_off = 0
        for i in range(n):
            sub_off += finish_chunk(i, host_io, sub_off)
        for _, st in ctxs:
            torch.cuda.current_stream().wait_stream(st)
        if gatherer is not None:
            gatherer.gather(submap, sub_off)
        return sub_off

    def timed(mode, host_io):
        for _ in range(max(3, args.warmup)):
            step(mode, host_io)
        env.barrier()
        l0 = sum(s.kernel_launches for s, _ in ctxs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(mode, host_io)
        env.barrier()
        secs = par.max_over_ranks(time.perf_counter() - t0, env.dev)
        return secs, sum(s.kernel_launches for s, _ in ctxs) - l0

    sampler = bench.ClockSampler(env.local_rank)
    if env.rank == 0:
        sampler.start()
    res = {}
    for mode in ("exact", "cut"):
        for host_io in (False, True):
            res[(mode, host_io)] = timed(mode, host_io)
    clocks = sampler.stop() if env.rank == 0 else None
    # label difference between the two modes (host copies of every frame's classes), and the oracle across the first cut
    step("exact", True)
    torch.cuda.synchronize()
    exact = [c["labels"].numpy().copy() for c in chunks]
    step("cut", True)
    torch.cuda.synchronize()
    dpts, dframes = 0, 0
    for c, ex in zip(chunks, exact):
        d = c["labels"].numpy() != ex
        dpts += int(d.sum())
        dframes += sum(1 for f in range(c["b"] - c["a"]) if d[c["off"][f]:c["off"][f + 1]].any())
    dpts, dframes, total_pts = env.sum_over_ranks([dpts, dframes, npts_rank])
    parity = None
    if env.rank == 0:
        ncheck = min(len(chunks), 2)
        if ncheck:
            sc = [s for c in chunks[:ncheck] for s in c["scans"]]
            po = np.concatenate([c["poses"] for c in chunks[:ncheck]])
            orc = _oracle(params)
            _, olab, ooff = orc.run_sequence(sc, po, nthreads=bench.host_cores())
            orc.close()
            nfr = len(sc) - (1 if (len(chunks) > ncheck or env.world > 1) else 0)  # the last checked frame is `pre` of a frame outside the check
            got = np.concatenate([e for e in exact[:ncheck]])
            mism = int((got[: ooff[nfr]] != olab[: ooff[nfr]]).sum())
            parity = {"checked": f"per-point classes of frames 0..{nfr - 1} of the exact chain (crossing {ncheck - 1} chunk cut(s)) against the oracle's single chain",
                      "points": int(ooff[nfr]), "mismatching_points": mism, "bit_exact": mism == 0}
    if env.rank == 0:
        ex_dev, launches = res[("exact", False)]
        ex_e2e, _ = res[("exact", True)]
        cut_dev, _ = res[("cut", False)]
        cut_e2e, _ = res[("cut", True)]
        line = {
            "metric": "scans/sec, ONE 1000-scan sequence end to end (BASELINE configs[3])", "value": T * args.steps / ex_dev, "unit": "scans/s",
            "n_gpus": env.world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": 1000.0 * ex_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[3]: one {T}-scan synthetic SemanticKITTI-shape sequence, {S}-scan chunks block-sharded over {env.world} rank(s) "
                                   f"({len(chunks)} chunk contexts on rank 0), ONE unbroken tracking chain (tail hand-off between chunk owners), "
                                   "one all-gather of the static submaps per step",
                       "scans_per_step": T, "scans_per_chunk": S, "rings": RINGS, "cols": COLS,
                       "l2": f"{T * 1.8:.0f} MB of input per step, > the 126 MB L2"},
            "e2e": {"value": T * args.steps / ex_e2e, "unit": "scans/s", "h2d_bytes_per_step": int(total_pts * 16), "d2h_bytes_per_step": int(total_pts),
                    "ms_per_step": 1000.0 * ex_e2e / args.steps},
            "gpu_launches": int(launches),
            "cut_mode": {"value": T * args.steps / cut_dev, "e2e": T * args.steps / cut_e2e, "unit": "scans/s",
                         "what": "every chunk tracked as its own sequence (chunks run side by side; one tracking(k, k+1) is dropped per cut)",
                         "label_difference_vs_exact": {"points": int(dpts), "of_points": int(total_pts), "frames": int(dframes), "of_frames": T,
                                                       "cuts": len(par.chunk_sequence(T, S)) - 1}},
            "parity": parity,
            "roofline": {"bound": "hbm", "achieved": 7.0e6 * T * args.steps / ex_dev / 1e9, "peak": _peak()[0], "unit": "GB/s",
                         "frac": 7.0e6 * T * args.steps / ex_dev / 1e9 / _peak()[0], "traffic": None,
                         "note": "whole path, 7.0 MB algorithmic bytes per scan (SURVEY 8d); the exact mode is bound by the serial tracking chain "
                                 "(one k_track round trip per frame pair), not by bandwidth"},
            "clocks": clocks,
        }
        print(json.dumps(line))
    for s, _ in ctxs:
        s.close()
    env.finish()
    return 0


# ------------------------------------------------------------------------------------------------------------------
# configs[2]: parkinglot dense scans + GICP scan-to-map
# ------------------------------------------------------------------------------------------------------------------
def run_parkinglot_gicp(args, bench):
    env = Env(args, bench)
    torch, pkg = env.torch, env.pkg
    params = pkg.parkinglot_params()
    rings, cols, nmap = 128, 2700, 2
    scans = [pkg.synth_scan(SEED + 5 + env.rank, k, rings=rings, cols=cols) for k in range(nmap + 3)]

    def to_world(s, pose):
        Tm = pkg.pose_matrix(pose)
        out = s.copy()
        out[:, :3] = s[:, :3] @ Tm[:, :3].T + Tm[:, 3]
        return out

    tgt = np.concatenate([to_world(s, p) for s, p in scans[:nmap]])
    srcs = scans[nmap:]
    gp = pkg.gicp_default_params()
    ssc = pkg.SSC(params, device=env.local_rank, max_points=rings * cols, max_batch=1)
    ssc.set_option("inspect", 0)
    tgt_dev = torch.from_numpy(tgt).to(env.dev)
    src_dev = [torch.from_numpy(s).to(env.dev) for s, _ in srcs]
    src_pin = [torch.from_numpy(s).pin_memory() for s, _ in srcs]
    ssc.gicp_set_target_device(tgt_dev.data_ptr(), len(tgt), gp)
    guesses = []
    for s, pose in srcs:
        pert = pose.copy()
        pert[0] += 0.15
        pert[1] -= 0.10
        pert[5] += np.deg2rad(1.5)
        guesses.append(pkg.pose_matrix(pert))
    last = {}

    def step(i, host_io):
        k = i % len(srcs)
        s, pose = srcs[k]
        off = np.array([0, len(s)], np.int64)
        ssc.reset()
        if host_io:
            ssc.process_host_ptr(src_pin[k].data_ptr(), off)
            r = ssc.gicp_align(s, guesses[k])
        else:
            ssc.process_device(src_dev[k].data_ptr(), off)
            r = ssc.gicp_align_device(src_dev[k].data_ptr(), len(s), guesses[k])
        last[k] = r
        return r

    def timed(host_io):
        for i in range(max(3, args.warmup)):
            step(i, host_io)
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ssc.kernel_launches
        t0 = time.perf_counter()
        iters = 0
        for i in range(args.steps):
            iters += step(i, host_io)["iterations"]
        env.barrier()
        return env.par.max_over_ranks(time.perf_counter() - t0, env.dev), iters, ssc.kernel_launches - l0

    sampler = bench.ClockSampler(env.local_rank)
    if env.rank == 0:
        sampler.start()
    secs_dev, iters, launches = timed(False)
    secs_e2e, _, _ = timed(True)
    clocks = sampler.stop() if env.rank == 0 else None
    # per-kernel times of one alignment
    pkg.kernel_timing(True)
    step(0, False)
    torch.cuda.synchronize()
    rep = pkg.kernel_timing_report()
    pkg.kernel_timing(False)
    if env.rank == 0:
        s, pose = srcs[0]
        orc = _oracle(params)
        t0 = time.perf_counter()
        oc = orc.gicp_align(s, tgt, guesses[0], gp)
        cpu_secs = time.perf_counter() - t0
        orc.close()
        g = step(0, False)
        dpos = float(np.abs(g["pose6"][:3] - oc["pose6"][:3]).max())
        drot = float(np.abs(g["pose6"][3:] - oc["pose6"][3:]).max())
        gk = {k: round(v[0], 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])}
        corr_ms = sum(v[0] for k, v in rep.items() if k in ("k_gicp_corr", "k_gicp_src_count", "k_gicp_src_place"))
        n_it = max(1, g["iterations"])
        peak, peak_src = _peak()
        bytes_it = 40.0 * (len(s) + len(tgt))
        line = {
            "metric": "dense scans/sec, parkinglot.yaml per-scan stages + GICP scan-to-map (BASELINE configs[2])", "value": env.world * args.steps / secs_dev,
            "unit": "scans/s", "n_gpus": env.world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": 1000.0 * secs_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[2]: parkinglot.yaml, dense {rings}x{cols} scans ({len(s)} pts), map = {nmap} dense scans ({len(tgt)} pts), "
                                   "initial guess 0.15 m / 0.10 m / 1.5 deg off; replicas (one stream of scans per rank)",
                       "l2": "inputs rotate over 3 scans; the map (>11 MB) and the scan are re-binned every iteration"},
            "e2e": {"value": env.world * args.steps / secs_e2e, "unit": "scans/s", "h2d_bytes_per_step": int(len(s) * 16 * 2), "d2h_bytes_per_step": 0,
                    "ms_per_step": 1000.0 * secs_e2e / args.steps},
            "gpu_launches": int(launches),
            "gicp": {"iterations_per_align": iters / args.steps, "align_ms_per_iteration_kernels": corr_ms / n_it,
                     "pose_delta_vs_oracle": {"max_abs_m": dpos, "max_abs_rad": drot, "tolerance": "1e-3 m / 1e-3 rad", "ok": dpos <= 1e-3 and drot <= 1e-3,
                                              "note": "self-consistency only: the reference holds no GICP code (docs/gicp_spec.md)"},
                     "kernels_ms_per_align": gk},
            "roofline": {"bound": "hbm", "kernel": "k_gicp_corr", "achieved": bytes_it * n_it / max(rep.get("k_gicp_corr", (1e9, 1))[0], 1e-9) / 1e6, "peak": peak,
                         "unit": "GB/s", "frac": bytes_it * n_it / max(rep.get("k_gicp_corr", (1e9, 1))[0], 1e-9) / 1e6 / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes": "40 B x (source + target points) per Gauss-Newton iteration (SURVEY 8d)"},
            "cpu_baseline": {"value": 1.0 / cpu_secs, "unit": "scans/s", "cores": 1, "kind": "port",
                             "sample": f"one GICP alignment of the same scan in {cpu_secs:.2f} s (oracle/gicp_oracle.cpp, double precision, one thread)"},
            "clocks": clocks,
        }
        print(json.dumps(line))
    ssc.close()
    env.finish()
    return 0


# ------------------------------------------------------------------------------------------------------------------
# configs[4]: binning + kNN kernels on aggregated dense clouds
# ------------------------------------------------------------------------------------------------------------------
def run_stress(args, bench):
    env = Env(args, bench)
    torch, pkg = env.torch, env.pkg
    params = pkg.parkinglot_params()
    peak, peak_src = _peak()

    def dense_cloud(n_target, seed):
        clouds, k = [], 0
        while sum(len(c) for c in clouds) < n_target:
            s, pose = pkg.synth_scan(seed, k, rings=128, cols=2250)
            Tm = pkg.pose_matrix(pose)
            out = s.copy()
            out[:, :3] = s[:, :3] @ Tm[:, :3].T + Tm[:, 3]
            clouds.append(out)
            k += 1
        return np.ascontiguousarray(np.concatenate(clouds)[:n_target], np.float32)

    ssc = pkg.SSC(params, device=env.local_rank, max_points=4096, max_batch=1)
    gp = pkg.gicp_default_params()
    sizes = [512_000, 2_048_000, 8_192_000]
    big = dense_cloud(max(sizes), SEED + 9 + env.rank)
    sweep = []
    headline = None
    env.barrier()
    sampler = bench.ClockSampler(env.local_rank)
    if env.rank == 0:
        sampler.start()
    t_all = time.perf_counter()
    for n in sizes:
        reps = max(3, args.steps)
        # rotate over disjoint slices of the big cloud so that consecutive repetitions do not find their input in L2
        slices = [big[o:o + n] for o in range(0, len(big) - n + 1, n)][:16] or [big[:n]]
        for i in range(3):
            ssc.makeApriVec(slices[i % len(slices)])
        pkg.kernel_timing(True)
        for i in range(reps):
            ssc.makeApriVec(slices[i % len(slices)])
        rep = pkg.kernel_timing_report()
        pkg.kernel_timing(False)
        ms = rep["k_bin_only"][0] / reps
        bytes_bin = n * (16 + 1 + 4 * 4 + 3 * 4)  # xyzi in; pass flag, four indices and three floats out (the inspection entry point)
        ent = {"kernel": "k_bin_only", "points": n, "ms": ms, "GBps": bytes_bin / ms / 1e6, "frac_of_peak": bytes_bin / ms / 1e6 / peak,
               "algorithmic_bytes": bytes_bin}
        sweep.append(ent)
        knn = None
        if n <= 2_100_000:
            cl = slices[0]
            ssc.gicp_normals(cl, gp)
            pkg.kernel_timing(True)
            kreps = max(2, min(5, args.steps))
            for i in range(kreps):
                ssc.gicp_normals(slices[i % len(slices)], gp)
            rep = pkg.kernel_timing_report()
            pkg.kernel_timing(False)
            kms = rep["k_gicp_normals"][0] / kreps
            build = sum(v[0] for k, v in rep.items() if k.startswith("k_gicp") and k != "k_gicp_normals") / kreps
            bytes_knn = n * 32
            knn = {"kernel": "k_gicp_normals (radius search + 3x3 eigen)", "points": n, "ms": kms, "grid_build_ms": build, "GBps": bytes_knn / kms / 1e6,
                   "frac_of_peak": bytes_knn / kms / 1e6 / peak, "algorithmic_bytes": bytes_knn}
            sweep.append(knn)
        if n == 512_000:
            headline = (ent, knn)
    secs = time.perf_counter() - t_all
    clocks = sampler.stop() if env.rank == 0 else None
    agg = env.sum_over_ranks([(headline[0]["algorithmic_bytes"] + headline[1]["algorithmic_bytes"]) / ((headline[0]["ms"] + headline[1]["ms"]) * 1e6)])[0]
    if env.rank == 0:
        line = {
            "metric": "algorithmic GB/s, curved-voxel binning + radius-kNN kernels on a 512k-point dense cloud (BASELINE configs[4])", "value": agg,
            "unit": "GB/s", "n_gpus": env.world, "steps": args.steps, "warmup": 3, "ms_per_step": headline[0]["ms"] + headline[1]["ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[4]: aggregated dense cloud (128x2250 scans moved into one frame), binning + radius search kernels only; replicas "
                                   "(every rank runs its own copy, no exchange); kernel time = the library's CUDA-event timers around each launch",
                       "l2": "repetitions rotate over disjoint slices of an 8.2 M-point cloud; the 512k cloud itself is 8 MB"},
            "e2e": None,
            "gpu_launches": int(ssc.kernel_launches),
            "roofline": {"bound": "hbm", "kernel": "k_bin_only @ 8.2 M points", "achieved": sweep[-1]["GBps"], "peak": peak, "unit": "GB/s",
                         "frac": sweep[-1]["frac_of_peak"], "traffic": None, "peak_source": peak_src,
                         "note": "the radius search is instruction-bound (O(points x neighbourhood candidates)), not bandwidth-bound: see frac_of_peak in sweep"},
            "sweep": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in e.items()} for e in sweep],
            "wall_s": secs,
            "clocks": clocks,
        }
        print(json.dumps(line))
    ssc.close()
    env.finish()
    return 0


def run(args, bench):
    return {"sequence": run_sequence, "parkinglot_gicp": run_parkinglot_gicp, "stress": run_stress}[args.workload](args, bench)
