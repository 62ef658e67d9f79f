#!/usr/bin/env python
"""R&D: where does the wall time of lsqr_ransac_batch from page-locked memory go?  Compares with a plain upload of the same buffer."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP32, FP64, Engine, synth
nprob, npts = 65536, 256
name = "plane3"
base = [synth.GENERATORS[name](npts, seed=100 + i)[0] for i in range(32)]
data = np.concatenate([base[i % 32] for i in range(nprob)])
offsets = (np.arange(nprob + 1) * npts).astype(np.uint64)
pin = torch.from_numpy(data).pin_memory()
eng = Engine(name, 0.5)
def t(f, reps=5):
    f(); ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts))
print("bytes", data.nbytes / 1e6, "MB")
print("upload_ptr (pinned, ingest)      ", t(lambda: eng.upload_ptr(pin.data_ptr(), nprob * npts, 24)), "ms")
dev = torch.empty_like(pin, device="cuda")
print("torch copy_ pinned -> device     ", t(lambda: (dev.copy_(pin, non_blocking=True), torch.cuda.synchronize())), "ms")
for cap in (1, 2048):
    for prec in (FP64, FP32):
        print(f"ransac_batch pinned max_tries={cap} prec={prec}", t(lambda: eng.ransac_batch(pin.numpy(), offsets, prob=0.999, max_tries=cap, seed=3, precision=prec)), "ms")
print("ransac_batch pageable max_tries=2048 ", t(lambda: eng.ransac_batch(data, offsets, prob=0.999, max_tries=2048, seed=3)), "ms")
eng.close()
