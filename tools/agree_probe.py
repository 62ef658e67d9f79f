#!/usr/bin/env python
"""R&D: latency of the estimator's agree() on one datum through the C ABI (the reference's examples call it datum by datum)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import Engine, synth
for name in ("plane3", "sphere3", "pivot"):
    data, true = synth.GENERATORS[name](1000, seed=1)
    eng = Engine(name, synth.DELTAS[name])
    prm = eng.estimate(data[: eng.k])
    for n in (1, 16, 17, 256):
        eng.agree(prm, data[:n])
        t0 = time.perf_counter()
        for i in range(2000):
            eng.agree(prm, data[i % 500: i % 500 + n])
        print(f"{name} agree() on {n:3d} data: {1e6 * (time.perf_counter() - t0) / 2000:7.2f} us per call", flush=True)
    eng.close()
