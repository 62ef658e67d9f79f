#!/usr/bin/env python
"""Small pass over every model and entry point, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP32, FP64, SAMPLE_EXHAUSTIVE, Engine, MODELS, synth  # noqa: E402

big = "--big" in sys.argv
only = os.environ.get("SMOKE_MODELS", "").split()          # e.g. SMOKE_MODELS="plane3 sphere3" for the slower racecheck
for name in MODELS:
    if only and name not in only:
        continue
    n = 3000
    data, true = synth.GENERATORS[name](n, seed=5)
    for ls in ([0, 1] if (name.startswith(("circle", "sphere")) or name in ("usxw", "uscp")) else [1]):
        eng = Engine(name, synth.DELTAS[name], ls_type=ls)
        eng.upload(data)
        r64 = eng.score(count=700, precision=FP64, seed=1, want_counts=True, want_params=True)
        r32 = eng.score(count=700, precision=FP32, seed=1, want_counts=True)
        if big and name in ("plane3", "line2d", "line3", "sphere3", "pivot", "uscp", "plane8", "sphere8", "line8", "dense8"):
            eng.score(count=98304 + 128, precision=FP32, seed=2)       # constant-bank kernel
        eng.consensus(r64["best_params"])
        eng.get_mask()
        prm = eng.refine()
        out = eng.ransac(0.99, precision=FP32, seed=3)
        eng.compute(data, 0.99, precision=FP32, seed=3)                # pipelined upload + first round + refine
        eng.estimate(data[: eng.k])
        eng.agree(r64["best_params"], data[:100])
        eng.least_squares(data[:500])
        if name in ("plane3", "line2d", "sphere3", "dense5", "plane8", "sphere8", "line5", "dense2"):
            off = np.arange(0, 9) * 200
            eng.ransac_batch(data[:1600], off, max_tries=256, want_masks=True)
            eng.ransac_batch(data[:1599], np.append(off[:-1], 1599), max_tries=256, want_masks=True, precision=FP32)   # odd-sized last problem: NaN-padded pair
        small = eng
        small.upload(data[:12])
        small.ransac_exhaustive()
        eng.close()
        print(name, ls, "ok", r64["best_count"], r32["best_count"], len(prm), flush=True)
# a batch large enough for several copy pieces and launches (4 MB pieces, 32 MB launches)
npb, per = 100000, 16
bd = np.tile(synth.GENERATORS["plane3"](per, seed=9)[0], (npb, 1))
eng = Engine("plane3", 0.5)
r = eng.ransac_batch(bd, (np.arange(npb + 1) * per).astype(np.uint64), exhaustive=True, want_masks=True)
assert len(set(r["counts"].tolist())) == 1, "identical problems must give identical answers in every chunk (exhaustive mode)"
eng.close()
print("chunked batch ok", flush=True)
eng = Engine("absor", 2.0)
d, _ = synth.absolute_orientation(2000, seed=1)
eng.weighted_least_squares(d, np.ones(2000))
eng.close()
print("done")
