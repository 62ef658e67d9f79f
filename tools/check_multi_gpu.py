#!/usr/bin/env python
"""Run under torchrun with N >= 2 GPUs: the all-gather upload (dist.upload_replicated) must leave every rank with exactly the
data a plain host upload gives -- same fp64 counts, same consensus mask, same refit -- and sharded requests (lsqr_set_shard with
the NCCL hooks) must return what an unsharded context returns."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP64, SAMPLE_LIST, Engine, synth  # noqa: E402
from lsqrrecipes_b200.dist import full_mask, init_native, install_hooks, upload_replicated  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, n in (("plane3", 1_000_003), ("absor", 200_001), ("pivot", 50_000)):
    data, _ = synth.GENERATORS[name](n, seed=3)
    host = torch.from_numpy(data).pin_memory()
    subs = synth.random_subsets(n, Engine(name, 1.0, device=local).k, 64, seed=4)
    res = []
    for mode in ("host", "gather"):
        eng = Engine(name, synth.DELTAS[name], device=local)
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        if mode == "host":
            eng.upload(data)
        else:
            upload_replicated(eng, host, rank, world)
        r = eng.score(sampler=SAMPLE_LIST, subsets=subs, precision=FP64, want_counts=True)
        cnt = eng.consensus(r["best_params"])
        res.append((r["counts"].copy(), cnt, eng.get_mask().copy(), eng.refine().copy()))
        eng.close()
    same = np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1] and np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])
    ok = ok and same
    print(f"rank {rank} {name}: gather == host upload: {same}", flush=True)
# sharded requests (hypotheses partitioned over the ranks, refine moments summed over point shards) against the same
# requests on an unsharded context: same winner, same consensus set, refit within the fp64 summation-order tolerance.
# Two ways of sharding: "native" = the library's own NCCL communicator (lsqr_ctx_init_nccl; the upload fetches every
# world-th chunk and all-gathers the rest inside the library), "hooks" = caller-supplied torch.distributed collectives.
from lsqrrecipes_b200 import FP32  # noqa: E402
for name, n, H in (("plane3", 400_003, 200_000), ("sphere3", 200_001, 50_000), ("circle2", 100_001, 20_000)):
    data, _ = synth.GENERATORS[name](n, seed=8)
    host = torch.from_numpy(data).pin_memory()
    out = []
    for mode in ("single", "native", "hooks"):
        eng = Engine(name, synth.DELTAS[name], device=local)
        if mode == "hooks":
            eng.set_stream(torch.cuda.current_stream().cuda_stream)
            install_hooks(eng, rank, world)
            upload_replicated(eng, host, rank, world)
        elif mode == "native":
            init_native(eng, rank, world)
            eng.upload(data if rank % 2 else host.numpy())     # pageable on odd ranks, page-locked on even ones
        else:
            eng.upload(data)
        r = eng.score(count=H, precision=FP32, seed=11)
        cnt = eng.consensus(r["best_params"])
        gather = full_mask if mode != "single" else (lambda m: m)     # a rank holds the consensus bits of its own point shard
        out.append((r["best_index"], r["best_count"], r["best_params"].copy(), cnt, gather(eng.get_mask()).copy(), eng.refine().copy()))
        c = eng.ransac(0.999, precision=FP32, seed=12)
        if mode == "native":   # and the pipelined one-call form gives the same answer
            c2 = eng.compute(data, 0.999, precision=FP32, seed=12)
            assert (c2["best_index"], c2["fraction"], c2["tries"]) == (c["best_index"], c["fraction"], c["tries"]) and np.array_equal(full_mask(c2["mask"]), full_mask(c["mask"]))
        out[-1] += (c["best_index"], c["fraction"], gather(c["mask"]).copy(), c["params"].copy())
        eng.close()
    for mode, b in zip(("native", "hooks"), out[1:]):
        a = out[0]
        checks = {"best_index": a[0] == b[0], "best_count": a[1] == b[1], "best_params": np.array_equal(a[2], b[2]), "consensus": a[3] == b[3],
                  "mask": np.array_equal(a[4], b[4]), "refit": a[5].shape == b[5].shape and np.allclose(a[5], b[5], rtol=1e-9, atol=1e-9),
                  "compute_index": a[6] == b[6], "compute_fraction": a[7] == b[7], "compute_mask": np.array_equal(a[8], b[8]),
                  "compute_params": a[9].shape == b[9].shape and np.allclose(a[9], b[9], rtol=1e-6, atol=1e-6)}
        same = all(checks.values())
        if not same:
            print(f"rank {rank} {name} {mode}: differs in {[k for k, v in checks.items() if not v]}: {a[0], a[1], a[3], a[6], a[7]} vs {b[0], b[1], b[3], b[6], b[7]}\n  {a[5]}\n  {b[5]}", flush=True)
        ok = ok and same
        print(f"rank {rank} {name}: {mode} sharded == unsharded: {same}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU CHECK OK" if int(t.item()) == 1 else "MULTI-GPU CHECK MISMATCH", flush=True)
dist.destroy_process_group()
