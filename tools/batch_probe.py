#!/usr/bin/env python
"""R&D: where does the batched kernel's time go?  kernel ms against the cap on tries (rounds of 32)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP32, FP64, Engine, synth
nprob, npts = 65536, 256
for name in sys.argv[1:] or ["line2d", "plane3"]:
    D = synth.GENERATORS[name](8, seed=1)[0].shape[1]
    rng = np.random.default_rng(7)
    bases = [synth.GENERATORS[name](npts, seed=100 + i)[0] for i in range(32)]
    data = np.empty((nprob * npts, D))
    for i in range(nprob):
        data[i * npts:(i + 1) * npts] = bases[i % 32] + rng.uniform(-50, 50, D)
    offsets = (np.arange(nprob + 1) * npts).astype(np.uint64)
    eng = Engine(name, synth.DELTAS[name], ls_type=1)
    eng.ransac_batch(data[: 64 * npts], offsets[:65], max_tries=32, seed=1)
    for cap in (1, 32, 96, 256, 2048):
        for prec, tag in ((FP64, "fp64"), (FP32, "fp32")):
            out = eng.ransac_batch(data, offsets, exhaustive=False, prob=0.999, max_tries=cap, seed=2, precision=prec)
            print(f"{name} {tag} scoring, max_tries={cap:5d}: kernel {out['device_ms']:7.3f} ms  mean count {out['counts'].mean():6.1f}", flush=True)
    eng.close()
