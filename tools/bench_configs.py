#!/usr/bin/env python
"""Measures the BASELINE.json configs that are not the headline bench line (parity-test cases there):
   configs[2] sphere RANSAC + Levenberg-Marquardt refine, 10 M points
   configs[3] absolute orientation over 1 M correspondences
   configs[4] 65,536 independent small line / plane problems (one CTA per problem)
   plus the fp64 validation mode of configs[1] on a hypothesis slice,
each next to the reference's CPU implementation (oracle/_ref) on a bounded sample.
One JSON object per line; run on the GPU box:  python tools/bench_configs.py [--quick]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP32, FP64, Engine, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def cpu_oracle():
    return pyoracle.Oracle("ref" if pyoracle.available("ref") else "port")


def scoring(name, n, H, precision, label, cpu_hyps=0):
    data, _ = synth.GENERATORS[name](n)
    delta = synth.DELTAS[name]
    eng = Engine(name, delta)
    eng.upload(data)
    eng.score(count=min(H, 4096), precision=precision, seed=1)
    r = eng.score(count=H, precision=precision, seed=2)
    rec = {"config": label, "model": name, "points": n, "hypotheses": H, "precision": "fp32" if precision == FP32 else "fp64",
           "consensus_ms": r["consensus_ms"], "score_ms": r["score_ms"], "evals_per_s": float(H) * n / (r["consensus_ms"] * 1e-3),
           "best_count": r["best_count"]}
    if cpu_hyps:
        orc = cpu_oracle()
        m = pyoracle.MODELS[name]
        subs = synth.random_subsets(n, pyoracle.INFO[m][2], cpu_hyps, seed=3)
        t0 = time.perf_counter()
        orc.score_subsets(m, delta, data, subs, want_params=False)
        dt = time.perf_counter() - t0
        rec["cpu_evals_per_s"] = float(cpu_hyps) * n / dt
        rec["cpu_cores"] = orc.num_threads()
        rec["cpu_kind"] = orc.kind
    eng.close()
    emit(**rec)
    return data, delta


def compute_e2e(name, n, ls_type, label, cpu_n=0):
    data, true = synth.GENERATORS[name](n)
    delta = synth.DELTAS[name]
    eng = Engine(name, delta, ls_type=ls_type)
    eng.upload(data)
    eng.ransac(0.999, precision=FP32, seed=1)          # warm-up
    t0 = time.perf_counter()
    eng.upload(data)
    r = eng.ransac(0.999, precision=FP32, seed=2)
    dt = time.perf_counter() - t0
    st = eng.last_refine_stats()
    rec = {"config": label, "model": name, "points": n, "compute_ms": dt * 1e3, "device_ms": r["device_ms"], "tries": r["tries"],
           "fraction": r["fraction"], "params": [float(x) for x in r["params"]], "true": [float(x) for x in true],
           "lm_iterations": st["lm_iterations"], "refine_pass_ms": st["kernel_ms"],
           "refine_pass_GBps": st["bytes"] / (st["kernel_ms"] * 1e-3) / 1e9 if st["kernel_ms"] > 0 else None}
    if cpu_n:
        orc = cpu_oracle()
        if orc.kind == "ref":
            sub = data[:cpu_n]
            t0 = time.perf_counter()
            prm, mask, frac = orc.ransac_random(pyoracle.MODELS[name], delta, sub, 0.999, ls_type=ls_type)
            rec["cpu_reference_compute_ms"] = (time.perf_counter() - t0) * 1e3
            rec["cpu_reference_points"] = cpu_n
            rec["cpu_reference_fraction"] = frac
    eng.close()
    emit(**rec)


def batched(name, nprob, npts, label, cpu_problems=0):
    m = pyoracle.MODELS[name]
    D = pyoracle.INFO[m][0]
    delta = synth.DELTAS[name]
    rng = np.random.default_rng(7)
    base, _ = synth.GENERATORS[name](npts, seed=11)
    # distinct problems: random rigid shifts of a few base problems keep generation cheap
    data = np.empty((nprob * npts, D))
    bases = [synth.GENERATORS[name](npts, seed=100 + i)[0] for i in range(32)]
    for i in range(nprob):
        data[i * npts:(i + 1) * npts] = bases[i % 32] + rng.uniform(-50, 50, D)
    offsets = (np.arange(nprob + 1) * npts).astype(np.uint64)
    eng = Engine(name, delta, ls_type=1)
    eng.ransac_batch(data[: 64 * npts], offsets[:65], exhaustive=False, prob=0.999, max_tries=2048, seed=1)
    t0 = time.perf_counter()
    out = eng.ransac_batch(data, offsets, exhaustive=False, prob=0.999, max_tries=2048, seed=2)
    dt = time.perf_counter() - t0
    rec = {"config": label, "model": name, "problems": nprob, "points_per_problem": npts, "wall_ms": dt * 1e3, "kernel_ms": out["device_ms"],
           "problems_per_s_kernel": nprob / (out["device_ms"] * 1e-3), "mean_inlier_fraction": float(out["counts"].mean() / npts)}
    if cpu_problems:
        orc = cpu_oracle()
        if orc.kind == "ref":
            t0 = time.perf_counter()
            for i in range(cpu_problems):
                orc.ransac_random(m, delta, data[i * npts:(i + 1) * npts], 0.999)
            rec["cpu_reference_problems_per_s"] = cpu_problems / (time.perf_counter() - t0)
            rec["cpu_reference_threads"] = 1
    eng.close()
    emit(**rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    q = a.quick
    N = 1_000_000 if q else 10_000_000
    scoring("plane3", N, 20_000 if q else 100_000, FP64, "configs[1] fp64 validation slice", cpu_hyps=32)
    scoring("sphere3", N, 100_000 if q else 1_000_000, FP32, "configs[2] sphere scoring", cpu_hyps=32)
    compute_e2e("sphere3", N, 1, "configs[2] sphere compute() + LM refine", cpu_n=200_000)
    compute_e2e("sphere3", N, 0, "configs[2] sphere compute(), algebraic refine")
    scoring("absor", 1_000_000, 100_000 if q else 1_000_000, FP32, "configs[3] absolute orientation scoring", cpu_hyps=64)
    compute_e2e("absor", 1_000_000, 1, "configs[3] absolute orientation compute()", cpu_n=200_000)
    compute_e2e("plane3", N, 1, "configs[1] plane compute()", cpu_n=200_000)
    # next rows of SURVEY.md 8f: dense linear systems and the cross-wire ultrasound calibration (LM over 11 parameters)
    scoring("dense6", 1_000_000, 100_000 if q else 1_000_000, FP32, "8f-2 dense linear system (n = 6) scoring", cpu_hyps=32)
    compute_e2e("dense6", 1_000_000, 1, "8f-2 dense linear system (n = 6) compute()", cpu_n=200_000)
    scoring("usxw", 200_000, 100_000 if q else 1_000_000, FP32, "8f-3 cross-wire US calibration scoring", cpu_hyps=32)
    compute_e2e("usxw", 200_000, 1, "8f-3 cross-wire US calibration compute() + LM refine (11 parameters)", cpu_n=20_000)
    nprob = 4096 if q else 65536
    batched("line2d", nprob, 256, "configs[4] batched 2D line", cpu_problems=256)
    batched("plane3", nprob, 256, "configs[4] batched 3D plane", cpu_problems=256)


if __name__ == "__main__":
    main()
