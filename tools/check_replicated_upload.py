#!/usr/bin/env python
"""Run under torchrun with N >= 2 GPUs: the all-gather upload (dist.upload_replicated) must leave every rank with exactly the
data a plain host upload gives -- same fp64 counts, same consensus mask, same refit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP64, SAMPLE_LIST, Engine, synth  # noqa: E402
from lsqrrecipes_b200.dist import install_hooks, upload_replicated  # noqa: E402

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
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("REPLICATED UPLOAD OK" if int(t.item()) == 1 else "REPLICATED UPLOAD MISMATCH", flush=True)
dist.destroy_process_group()
