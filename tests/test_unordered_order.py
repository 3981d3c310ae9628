"""CPU suite: the iteration order of libstdc++'s std::unordered_map<int, T> - which the reference's cluster bookkeeping makes
observable (cluster_set is iterated in that order by tracking(), src/ssc.cpp:1261, and by both refines) - follows a compact rule:

  * hash(key) = (size_t)key, bucket = hash % bucket_count;
  * a new node goes to the HEAD of its bucket's run in the singly linked node list, or to the FRONT of the whole list when its
    bucket is empty;
  * a rehash re-inserts every node, in the old iteration order, by the same rule under the new bucket count;
  * erase unlinks the node and changes nothing else; clear() keeps the bucket count;
  * the bucket count grows 1 -> 13 -> 29 -> 59 -> 127 -> 257 -> 541 -> ... when an insert would exceed the load factor 1.

The product keeps these containers on the host (csrc/host_cluster.cpp, scvod_api.cu::track_cars) precisely because of this order;
this test pins the rule against the real container (tools/probe/unordered_order.cpp, built here with g++), so that a device-side
decision table has a checked specification to follow."""
import os
import random
import subprocess

import pytest

import conftest

SRC = os.path.join(conftest.ROOT, "tools", "probe", "unordered_order.cpp")


class UnorderedIntMapOrder:
    """The rule above.  `order` is the iteration order; buckets are implicit (the nodes of a bucket are contiguous in the list)."""

    def __init__(self):
        self.order = []
        self.nb = 1

    def _place(self, lst, key, nb):
        b = (key % (1 << 64)) % nb  # (size_t)key
        for i, k in enumerate(lst):
            if (k % (1 << 64)) % nb == b:
                lst.insert(i, key)
                return
        lst.insert(0, key)

    def rehash(self, nb):
        new = []
        for k in self.order:
            self._place(new, k, nb)
        self.order, self.nb = new, nb

    def insert(self, key, nb_after):
        if key in self.order:
            return
        if nb_after != self.nb:  # the growth policy is checked separately; here the real container tells when it rehashes
            self.rehash(nb_after)
        self._place(self.order, key, self.nb)

    def erase(self, key):
        if key in self.order:
            self.order.remove(key)

    def clear(self):
        self.order = []


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("probe") / "unordered_order")
    subprocess.check_call(["g++", "-O1", "-std=gnu++17", SRC, "-o", exe])
    return exe


def run_probe(exe, ops):
    text = "\n".join("c" if o[0] == "c" else f"{o[0]} {o[1]}" for o in ops) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.splitlines()
    res = []
    for line in out:
        nb, _, keys = line.partition(":")
        res.append((int(nb), [int(k) for k in keys.split()]))
    return res


def test_bucket_growth_sequence(probe):
    res = run_probe(probe, [("i", k) for k in range(1200)])
    growth = []
    for nb, _ in res:
        if not growth or growth[-1] != nb:
            growth.append(nb)
    assert growth[:8] == [13, 29, 59, 127, 257, 541, 1109, 2357]
    # a rehash happens exactly when the element count would exceed the bucket count (max_load_factor 1)
    for n, (nb, keys) in enumerate(res, start=1):
        assert len(keys) == n and n <= nb


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_iteration_order_rule_matches_libstdcxx(probe, seed):
    rng = random.Random(seed)
    ops = []
    live = set()
    for step in range(1500):
        r = rng.random()
        if r < 0.60 or not live:
            # cluster names are small positive ints that grow; a few negative and large keys exercise the size_t cast
            k = rng.choice([rng.randrange(0, 400), rng.randrange(0, 40), -rng.randrange(1, 50), rng.randrange(10**6, 10**6 + 200)])
            ops.append(("i", k))
            live.add(k)
        elif r < 0.97:
            k = rng.choice(sorted(live)) if rng.random() < 0.8 else rng.randrange(0, 400)
            ops.append(("e", k))
            live.discard(k)
        else:
            ops.append(("c", 0))
            live.clear()
    res = run_probe(probe, ops)
    emu = UnorderedIntMapOrder()
    for (op, key), (nb, keys) in zip(ops, res):
        if op == "i":
            emu.insert(key, nb)
        elif op == "e":
            emu.erase(key)
        else:
            emu.clear()
        assert emu.nb == nb or op != "i", (op, key)
        assert emu.order == keys, (op, key, nb)


def test_the_orders_the_pipeline_observes(probe):
    """Shapes the path produces: names 5, 6, 7, ... inserted in order of first appearance (clusterAndCreateFrame), a fusion
    (erase several, insert the survivor again), a split (insert max_name++), then iteration."""
    ops = [("i", k) for k in (5, 9, 6, 14, 7, 8, 10, 11, 12, 13, 15, 16, 17, 18)]  # 14 names: crosses the 13 -> 29 rehash
    ops += [("e", 9), ("e", 14), ("e", 6), ("i", 6)]  # refineClusterByIntensity: fuse {9, 14, 6} into 6
    ops += [("i", 19), ("e", 10), ("e", 11), ("i", 20)]  # tracking: a split cluster (max_name++), then a fused car cluster
    res = run_probe(probe, ops)
    emu = UnorderedIntMapOrder()
    for (op, key), (nb, keys) in zip(ops, res):
        emu.insert(key, nb) if op == "i" else emu.erase(key)
        assert emu.order == keys
    assert res[13][0] == 29 and res[12][0] == 13
