"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: chunk sharding and the static-submap all-gather.

The GPU compute cannot run here; what these cover is what differs between N=1 and N>1: which rank owns which
sequence chunk, the variable-length gather (count exchange + padded all_gather), rank-order concatenation (the
reference's `*instance_map += ...` order, src/ssc.cpp:553-555) and the max-over-ranks timing reduction.
"""
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conftest


def load_parallel():
    name = "scvod_b200_parallel"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(conftest.PKG_DIR, "parallel.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def test_sharding_covers_every_chunk_once():
    par = load_parallel()
    for n_scans, chunk, world in [(1000, 64, 8), (1000, 64, 1), (7, 3, 2), (64, 64, 4), (0, 8, 2), (129, 16, 3)]:
        chunks = par.chunk_sequence(n_scans, chunk)
        assert sum(e - s for s, e in chunks) == n_scans and all(0 < e - s <= chunk for s, e in chunks)
        owned = [par.shard_chunks(n_scans, chunk, world, r) for r in range(world)]
        flat = [c for o in owned for c in o]
        assert flat == chunks  # block partition: consecutive chunks stay on one rank, rank order = sequence order
        sizes = [len(o) for o in owned]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_range(10, 2, 2)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = load_parallel()
    dev = torch.device("cpu")
    cap = 1000
    g = par.SubmapGatherer(cap, dev, granule=64)
    ok = True
    for step in range(4):
        n = 100 * (rank + 1) + 7 * step  # different valid counts per rank and per step
        sub = torch.full((cap, 4), float("nan"))
        sub[:n] = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4) + 10000 * rank + step
        merged, counts = g.gather(sub, n, padded=(step == 3))  # exact-size rows (default) and the capacity-sized ablation
        ok &= g.stride == (cap if step == 3 else -(-max(100 * world + 7 * step, 1) // 64) * 64)
        want_counts = [100 * (r + 1) + 7 * step for r in range(world)]
        ok &= counts.tolist() == want_counts
        comp = g.compact()
        ref = torch.cat([torch.arange(c * 4, dtype=torch.float32).reshape(c, 4) + 10000 * r + step for r, c in enumerate(want_counts)])
        ok &= bool(torch.equal(comp, ref))
    t = par.max_over_ranks(1.0 + rank, dev)
    ok &= t == float(world)
    try:
        g.gather(torch.zeros((cap, 4)), cap + 1)
        ok = False
    except ValueError:
        pass
    results[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_submap_gather_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


def test_single_process_gather_and_label_merge():
    par = load_parallel()
    g = par.SubmapGatherer(16, torch.device("cpu"))
    sub = torch.arange(64, dtype=torch.float32).reshape(16, 4)
    merged, counts = g.gather(sub, 5)
    assert counts.tolist() == [5] and torch.equal(g.compact(), sub[:5])
    labels = par.merge_labels_host([[np.array([1]), np.array([2])], [np.array([3])]], [(0, 0, 0), (0, 1, 1), (1, 0, 2)])
    assert [int(a[0]) for a in labels] == [1, 2, 3]


def _link_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = load_parallel()
    link = par.ChainLink()
    rng = np.random.default_rng(5)
    tail = rng.integers(0, 255, 4 * 4 * 3 + 16 * 1000, dtype=np.uint8)  # header + 2 clusters + 1000 points worth of bytes
    states = np.array([[1, 2], [0, 1]], np.int32)
    ok = True
    if rank == 0:  # owner of the earlier chunk: sends its tail, waits for the states
        link.send_tail(tail, 1)
        ok &= np.array_equal(link.recv_states(1), states)
        link.send_tail(np.zeros(0, np.uint8), 1)  # an empty tail travels too
    else:
        ok &= np.array_equal(link.recv_tail(0), tail)
        link.send_states(states, 0)
        ok &= link.recv_tail(0).size == 0
    results[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_chain_link_world_size_2_gloo():
    """Transport of the chain hand-off between ranks (the tail of rank r's last frame -> rank r + 1, states back)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_link_worker, args=(2, port, results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}
