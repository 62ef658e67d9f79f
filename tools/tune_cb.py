#!/usr/bin/env python
"""R&D: times the fp32 consensus of one model with every library variant under tools/bin/variants/."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "plane3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
libs = sorted(glob.glob(os.path.join(ROOT, f"tools/bin/variants/lib_m{api.MODELS[name]}_*.so"))) + [api.lib_path()]   # tools/build_variants.sh "<id> <R> <PPI>"
pid = os.fork() if False else None
for lib in libs:
    # one subprocess per variant: ctypes cannot unload a library
    r, w = os.pipe()
    if os.fork() == 0:
        api.lib_path = lambda lib=lib: lib
        from lsqrrecipes_b200 import FP32, Engine, synth
        data, _ = synth.GENERATORS[name](n)
        eng = Engine(name, synth.DELTAS[name])
        eng.upload(data)
        eng.score(count=H, precision=FP32, seed=1)
        best = min(eng.score(count=H, precision=FP32, seed=2 + i)["consensus_ms"] for i in range(3))
        print(f"{os.path.basename(lib):28s} {name} n={n} H={H}: {best:9.3f} ms  {n * H / best / 1e9:7.3f} T evals/s  {148 * 4 * 32 * 1.965e9 / (n * H / best * 1e3):.2f} cyc/eval", flush=True)
        os._exit(0)
    os.wait()
