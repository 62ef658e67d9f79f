#!/usr/bin/env python
"""R&D / evidence: how far from the threshold can a datum be and still be decided differently by the fp32 fast mode?

For every estimator: H Philox hypotheses are scored in fp32 and in fp64 (same parameter vectors); then the fp64 count is taken
again at delta + b and delta - b.  Every fp32/fp64 count difference of a hypothesis must be explained by data whose residual
lies within b of the threshold, i.e. |c32 - c64| <= c64(delta + b) - c64(delta - b).  The smallest b (on a 2^k grid) for which
that holds for every hypothesis is the measured half-width of the fp32 band; tests/test_parity_gpu.py::_fp32_band uses twice
that.  Output: one JSON line per model.   python tools/measure_fp32_band.py [n] [H]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP32, FP64, MODELS, SAMPLE_PARAMS, Engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
only = sys.argv[3:]
for name in MODELS:
    if only and name not in only:
        continue
    data, _ = synth.GENERATORS[name](n, seed=4711)
    delta = synth.DELTAS[name]
    eng = Engine(name, delta)
    eng.upload(data)
    r32 = eng.score(count=H, seed=5, precision=FP32, want_counts=True, want_params=True)
    prm = r32["params"]
    ok = ~np.isnan(prm[:, 0])
    c32 = r32["counts"].astype(np.int64)
    c64 = eng.score(sampler=SAMPLE_PARAMS, params=prm, precision=FP64, want_counts=True)["counts"].astype(np.int64)
    diff = np.abs(c32 - c64)
    scale = float(np.abs(data).max() + 1.0)
    found = None
    for k in range(0, 24):
        b = 1e-6 * 2.0 ** k
        if b >= delta:
            break
        eng.set_estimator(delta + b)
        hi = eng.score(sampler=SAMPLE_PARAMS, params=prm, precision=FP64, want_counts=True)["counts"].astype(np.int64)
        eng.set_estimator(delta - b)
        lo = eng.score(sampler=SAMPLE_PARAMS, params=prm, precision=FP64, want_counts=True)["counts"].astype(np.int64)
        if np.all(diff[ok] <= (hi - lo)[ok]):
            found = b
            break
    eng.close()
    print(json.dumps({"model": name, "points": n, "hypotheses": int(ok.sum()), "max_count_diff": int(diff[ok].max()), "sum_count_diff": int(diff[ok].sum()),
                      "sum_counts": int(c64[ok].sum()), "band_abs": found, "coordinate_scale": scale, "band_rel_to_scale": (found / scale) if found else None,
                      "delta": delta}), flush=True)
