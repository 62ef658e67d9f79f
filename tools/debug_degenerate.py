#!/usr/bin/env python
"""R&D: prints the minimal subsets on which device and oracle disagree about rejection (synth.degenerate_pool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import FP64, SAMPLE_LIST, Engine, synth
from oracle.pyoracle import INFO, MODELS, Oracle
port = Oracle("port")
np.set_printoptions(precision=17, linewidth=200)
for name in sys.argv[1:] or ["absor"]:
    m = MODELS[name]; D, P, k = INFO[m]
    data = synth.degenerate_pool(name, seed=31 + m)
    subs = synth.random_subsets(len(data), k, 600, seed=7 + m)
    c_ref, p_ref = port.score_subsets(m, synth.DELTAS[name], data, subs)
    eng = Engine(name, synth.DELTAS[name]); eng.upload(data)
    r = eng.score(sampler=SAMPLE_LIST, subsets=subs, precision=FP64, want_counts=True, want_params=True)
    bad_o, bad_d = np.isnan(p_ref[:, 0]), np.isnan(r["params"][:, 0])
    idx = np.flatnonzero(bad_o != bad_d)
    print(name, "mismatches", len(idx), "n_valid device", r["n_valid"], "oracle valid", int((~bad_o).sum()))
    for h in idx[:4]:
        print(" subset", subs[h], "oracle", p_ref[h], "device", r["params"][h])
        print(data[subs[h]])
    eng.close()
