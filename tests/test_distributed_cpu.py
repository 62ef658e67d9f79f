"""The N>1 path on CPU: world_size-2 gloo run of the sharding + packed-key + moment-sum logic
(lsqrrecipes_b200/dist.py), with the CPU oracle standing in for each rank's scoring so that no GPU
is needed.  The same partition formulas and key layout are used by engine.cu on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from lsqrrecipes_b200 import dist as ldist
from lsqrrecipes_b200 import synth


def test_shard_formulas_cover_everything():
    for count in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [ldist.hypothesis_shard(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    for n in (0, 31, 32, 33, 10_000_000):
        for world in (1, 2, 4, 8):
            spans = [ldist.point_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert all(s[0] % 32 == 0 for s in spans)


def test_upload_slices_tile_the_data():
    for n in (0, 1, 7, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            spans = [ldist.upload_slice(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert all(s[1] - s[0] <= s[2] for s in spans) and len({s[2] for s in spans}) == 1


def test_packed_key_orders_like_the_reference():
    # larger count wins; equal counts -> smaller index wins (strict '>' of RANSAC.hxx:100,245)
    assert ldist.pack_key(10, 5) > ldist.pack_key(9, 0)
    assert ldist.pack_key(10, 5) > ldist.pack_key(10, 6)
    assert ldist.unpack_key(ldist.pack_key(123, 456)) == (123, 456)
    assert ldist.pack_key(2**31 - 1, 0) < 2**63          # fits a signed all-reduce
    counts = np.array([3, 9, 9, 2])
    assert ldist.unpack_key(ldist.best_key(counts, 100)) == (9, 101)


def _worker(rank, world, port_no, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import MODELS, Oracle
    orc = Oracle("port")
    m = MODELS["plane3"]
    n, H = 3000, 600
    data, _ = synth.plane(n, seed=5)
    subs = synth.random_subsets(n, 3, H, seed=6)
    lo, hi = ldist.hypothesis_shard(H, rank, world)
    counts, params = orc.score_subsets(m, 0.5, data, subs[lo:hi])
    key = torch.tensor([ldist.best_key(counts, lo)], dtype=torch.int64)
    dist.all_reduce(key, op=dist.ReduceOp.MAX)                      # exchange step 1
    cnt, idx = ldist.unpack_key(int(key.item()))
    # every rank re-derives the winner, then reduces the moments of its point shard
    best = orc.estimate(m, 0.5, data[subs[idx]])
    b, e = ldist.point_shard(n, rank, world)
    _, mask = orc.agree(m, 0.5, best, data[b:e])
    pts = data[b:e][mask.astype(bool)]
    mom = torch.tensor(np.concatenate([[len(pts)], pts.sum(0), (pts[:, :, None] * pts[:, None, :]).reshape(len(pts), 9).sum(0)]))
    dist.all_reduce(mom, op=dist.ReduceOp.SUM)                      # exchange step 2
    # replication of the points: each rank contributes its slice, the all-gather rebuilds the whole buffer
    odd, _ = synth.plane(1001, seed=9)
    mine = torch.from_numpy(odd.copy())
    lo_u, hi_u, _ = ldist.upload_slice(len(odd), rank, world)
    mine[:lo_u] = -1.0                                               # a rank only needs to hold its own slice
    mine[hi_u:] = -1.0
    full = ldist.gather_replicated(mine, rank, world, torch.device("cpu"))
    # consensus set: every rank holds the bits of its point shard, the OR over ranks is the whole set
    part = np.zeros(n, dtype=np.uint8)
    part[b:e] = mask
    whole = ldist.full_mask(part)
    _, want_mask = orc.agree(m, 0.5, best, data)
    out_q.put((rank, cnt, idx, mom.numpy(), bool(np.array_equal(full.numpy(), odd)) and bool(np.array_equal(whole, want_mask))))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(port):
    from oracle.pyoracle import MODELS
    m = MODELS["plane3"]
    n, H = 3000, 600
    data, _ = synth.plane(n, seed=5)
    subs = synth.random_subsets(n, 3, H, seed=6)
    counts, params = port.score_subsets(m, 0.5, data, subs)
    want_idx = int(np.argmax(counts))
    _, mask = port.agree(m, 0.5, params[want_idx], data)
    pts = data[mask.astype(bool)]
    want_mom = np.concatenate([[len(pts)], pts.sum(0), (pts[:, :, None] * pts[:, None, :]).reshape(len(pts), 9).sum(0)])

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, cnt, idx, mom, replicated in results:
        assert (cnt, idx) == (int(counts[want_idx]), want_idx)
        assert np.allclose(mom, want_mom, rtol=1e-12, atol=1e-6)
        assert replicated, "all-gather of the per-rank slices must rebuild the data exactly"
