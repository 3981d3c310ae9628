"""Multi-GPU plumbing of the SCV-OD path: scan sharding and the static-submap merge.

The path shards embarrassingly: per-scan stages are independent per scan and the tracking
diff is a chain *inside* a sequence, so whole sequence chunks are dealt to ranks (one process per GPU,
``torch.distributed``) and no data-path collective exists until the per-GPU static submaps are merged — the
reference's ``*instance_map += *rgb_ptr`` (src/ssc.cpp:553-555: plain concatenation, no dedup) becomes ONE
all-gather of the map-frame static points per step.  Nothing here computes on points: the submap tensor is
written by ``k_submap`` (scvod_static_submap_dev) straight into the send buffer.

Works with the ``nccl`` backend (CUDA tensors, NVLink/NVSwitch) and with ``gloo`` (CPU tensors; used by the
world_size-2 tests that run without a GPU).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Block partition of ``n_items`` consecutive items: rank g owns [g*n/world, (g+1)*n/world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def chunk_sequence(n_scans: int, chunk: int) -> List[Tuple[int, int]]:
    """Cuts a sequence of n_scans frames into chunks of at most ``chunk`` consecutive frames.

    Chunks are units of work (one context each).  For ONE sequence the chain is kept unbroken across them with the tail hand-off
    (``track_chunks_as_one_chain`` / ``ChainLink``); treating every chunk as its own sequence instead drops one tracking(k, k+1)
    per cut (DESIGN.md §6 has the measured label difference)."""
    if chunk <= 0:
        raise ValueError("chunk must be positive")
    return [(s, min(n_scans, s + chunk)) for s in range(0, n_scans, chunk)]


def track_chunks_as_one_chain(contexts: Sequence, poses: Sequence) -> None:
    """One unbroken tracking chain (reference src/ssc.cpp:1450-1452) over consecutive chunks that live in different contexts of
    this process: ``contexts[i]`` holds chunk i's frames (pushed, not yet tracked), ``poses[i]`` its [n_i, 6] poses.  The pair
    that straddles a cut is tracked by the later chunk's context from the exported tail of the earlier one."""
    for i, ssc in enumerate(contexts):
        if i > 0:
            st = ssc.track_from_tail(contexts[i - 1].export_tail(), poses[i - 1][-1], poses[i][0])
            contexts[i - 1].apply_tail_states(st)
        ssc.tracking(poses[i])


class ChainLink:
    """The same hand-off between ranks (one process per GPU): the tail of rank r's last chunk goes to rank r + 1, the (state, type)
    pairs come back.  Two point-to-point messages per cut (a length, then the bytes; < 4 MB: a frame's car clusters), so the
    chain stays a chain — ranks do their per-scan stages concurrently and their tracking in rank order.  Works over ``gloo`` (CPU
    tensors) and ``nccl`` (the bytes are staged through a CUDA tensor)."""

    def __init__(self, device: torch.device = torch.device("cpu"), group=None):
        self.device = device
        self.group = group

    def _send(self, arr, dst: int):
        t = torch.from_numpy(arr.view("uint8").reshape(-1).copy())
        n = torch.tensor([t.numel()], dtype=torch.int64)
        dist.send(n.to(self.device), dst, group=self.group)
        if t.numel():
            dist.send(t.to(self.device), dst, group=self.group)

    def _recv(self, src: int):
        n = torch.zeros(1, dtype=torch.int64, device=self.device)
        dist.recv(n, src, group=self.group)
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=self.device)
        if t.numel():
            dist.recv(t, src, group=self.group)
        return t.cpu().numpy()

    def send_tail(self, tail, dst: int):
        self._send(tail, dst)

    def recv_tail(self, src: int):
        return self._recv(src)

    def send_states(self, state_type, dst: int):
        import numpy as np

        self._send(np.ascontiguousarray(state_type, np.int32), dst)

    def recv_states(self, src: int):
        return self._recv(src).view("int32").reshape(-1, 2)


def shard_chunks(n_scans: int, chunk: int, world: int, rank: int) -> List[Tuple[int, int]]:
    """The chunks of a sequence owned by ``rank``: consecutive chunks stay on one GPU (block partition)."""
    chunks = chunk_sequence(n_scans, chunk)
    lo, hi = shard_range(len(chunks), world, rank)
    return chunks[lo:hi]


class SubmapGatherer:
    """All-gather of variable-length per-rank static submaps ([n_r, 4] float32 xyzi in the map frame).

    One count exchange + one ``all_gather_into_tensor`` per call; buffers are allocated once for ``cap_points`` per rank.
    The data gather moves ``rows`` = the largest count of the call (rounded up to ``granule`` rows) per rank, not the capacity
    (``padded=True`` restores the capacity-sized gather: the ablation of DESIGN.md section 6).  ``gather`` returns
    (merged buffer, counts [world]); rank r's points are ``merged[r*stride : r*stride + counts[r]]`` with ``stride`` = the rows of
    that call (``self.stride``) - ``compact`` concatenates them in rank order, which is the reference's concatenation order when
    ranks own consecutive chunks.
    """

    def __init__(self, cap_points: int, device: torch.device, group=None, granule: int = 4096):
        self.cap = int(cap_points)
        self.device = device
        self.group = group
        self.granule = max(1, int(granule))
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.merged = torch.empty((self.world * self.cap, 4), dtype=torch.float32, device=device)
        self.counts = torch.zeros(self.world, dtype=torch.int64, device=device)
        self._mine = torch.zeros(1, dtype=torch.int64, device=device)
        self.stride = self.cap
        self.bytes_moved = 0  # bytes this rank received from the data gathers so far (diagnostics)

    def gather(self, submap: torch.Tensor, n_points: int, padded: bool = False):
        if submap.shape[0] < self.cap or submap.shape[1] != 4 or submap.dtype != torch.float32:
            raise ValueError("submap must be a [>=cap, 4] float32 tensor")
        if n_points > self.cap:
            raise ValueError("submap holds more points than the gather capacity")
        self._mine.fill_(int(n_points))
        if self.world == 1:
            self.counts.copy_(self._mine)
            self.stride = self.cap
            self.merged[: n_points].copy_(submap[: n_points])
            return self.merged, self.counts
        dist.all_gather_into_tensor(self.counts, self._mine, group=self.group)
        if padded:
            rows = self.cap
        else:  # every rank derives the same row count from the gathered counts (one small device->host read per call)
            rows = int(self.counts.max().item())
            rows = min(self.cap, max(self.granule, -(-rows // self.granule) * self.granule))
        self.stride = rows
        dist.all_gather_into_tensor(self.merged[: self.world * rows], submap[:rows], group=self.group)
        self.bytes_moved += (self.world - 1) * rows * 16
        return self.merged, self.counts

    def compact(self) -> torch.Tensor:
        """Concatenation of every rank's valid points, in rank order."""
        counts = [int(c) for c in self.counts.tolist()]
        return torch.cat([self.merged[r * self.stride: r * self.stride + counts[r]] for r in range(self.world)], dim=0)


def max_over_ranks(seconds: float, device: torch.device, group=None) -> float:
    """Timing rule of the bench: a multi-GPU number is the MAX over ranks."""
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def merge_labels_host(per_rank_labels: Sequence[Sequence], order: Sequence[Tuple[int, int, int]]):
    """Re-assembles per-frame label arrays computed on different ranks into sequence order.

    ``order`` lists (rank, local_index, global_frame) triples; returns a list indexed by global frame."""
    n = max(g for _, _, g in order) + 1 if order else 0
    out = [None] * n
    for r, i, g in order:
        out[g] = per_rank_labels[r][i]
    return out
