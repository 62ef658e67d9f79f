#!/usr/bin/env python
"""R&D: wall time of RANSAC<T,S>::compute through the C ABI on N points -- upload from page-locked and from pageable host
memory, the call on resident data, and the upload alone."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP32, Engine, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "plane3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else None
data, _ = synth.GENERATORS[name](N)
stride = data.shape[1] * 8
host = torch.from_numpy(data).pin_memory()
mask_pin = torch.empty(N, dtype=torch.uint8).pin_memory().numpy()
mask_pag = np.empty(N, dtype=np.uint8)
eng = Engine(name, synth.DELTAS[name], gpus=gpus)
print(f"{name} N={N} world={eng.world}")


def med(f, reps=7):
    t = []
    for i in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f(i)
        t.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(t[1:])), t


def up_pin(i):
    eng.upload_ptr(host.data_ptr(), N, stride)
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)


def up_pag(i):
    eng.upload(data)
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)


print("upload pinned   ms", med(up_pin))
print("upload pageable ms", med(up_pag))
print("ransac resident, pinned mask ms", med(lambda i: eng.ransac(0.999, precision=FP32, seed=i, mask_out=mask_pin)))
print("ransac resident, no mask     ms", med(lambda i: eng.ransac(0.999, precision=FP32, seed=i, want_mask=False)))
r = eng.ransac(0.999, precision=FP32, seed=3, want_mask=False)
print("   tries", r["tries"], "device_ms", r["device_ms"], "fraction", r["fraction"])


def full_pin(i):
    eng.upload_ptr(host.data_ptr(), N, stride)
    eng.ransac(0.999, precision=FP32, seed=i, mask_out=mask_pin)


def full_pag(i):
    eng.upload(data)
    eng.ransac(0.999, precision=FP32, seed=i, mask_out=mask_pag)


print("upload + ransac pinned   ms", med(full_pin))
print("upload + ransac pageable ms", med(full_pag))
print("lsqr_compute pinned   ms", med(lambda i: eng.compute(host.data_ptr(), 0.999, precision=FP32, seed=i, mask_out=mask_pin, n=N, stride_bytes=stride)))
print("lsqr_compute pageable ms", med(lambda i: eng.compute(data, 0.999, precision=FP32, seed=i, mask_out=mask_pag)))
print("lsqr_compute pinned, no mask ms", med(lambda i: eng.compute(host.data_ptr(), 0.999, precision=FP32, seed=i, want_mask=False, n=N, stride_bytes=stride)))
st = eng.last_refine_stats()
print("refine kernel", st)
eng.close()
