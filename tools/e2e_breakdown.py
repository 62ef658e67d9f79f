#!/usr/bin/env python
"""R&D: wall-clock breakdown of the end-to-end step bench.py times (host buffers in, mask + parameters out)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP32, Engine, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
data, _ = synth.plane(N)
host = torch.from_numpy(data).pin_memory()
eng = Engine("plane3", 0.5)
for rep in range(5):
    t = [time.perf_counter()]
    eng.upload_ptr(host.data_ptr(), N, 24); t.append(time.perf_counter())
    r = eng.score(count=H, precision=FP32, seed=rep); t.append(time.perf_counter())
    cnt = eng.consensus(r["best_params"]); t.append(time.perf_counter())
    mask = eng.get_mask(); t.append(time.perf_counter())
    prm = eng.refine(); t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
    print(f"rep {rep}: upload {d[0]:8.2f}  score {d[1]:8.2f} (device {r['score_ms']:8.2f}, consensus {r['consensus_ms']:8.2f})  consensus {d[2]:6.2f}  get_mask {d[3]:6.2f}  refine {d[4]:6.2f}  total {sum(d):8.2f} ms", flush=True)
