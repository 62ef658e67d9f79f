#!/usr/bin/env python
"""R&D: times the consensus-set + moments pass (mask_moments_kernel) with every library variant under tools/bin/variants/."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "plane3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
libs = [api.lib_path()] if os.environ.get("TUNE_MM_ONLY_SHIPPED") else sorted(glob.glob(os.path.join(ROOT, "tools/bin/variants/lib_mm_*.so"))) + [api.lib_path()]
for lib in libs:
    if os.fork() == 0:
        api.lib_path = lambda lib=lib: lib
        from lsqrrecipes_b200 import FP32, Engine, synth
        data, true = synth.GENERATORS[name](n)
        eng = Engine(name, synth.DELTAS[name])
        eng.upload(data)
        r = eng.score(count=2048, precision=FP32, seed=1)
        best = 1e9
        for i in range(6):
            eng.consensus(r["best_params"])
            st = eng.last_refine_stats()
            best = min(best, st["kernel_ms"])
        print(f"{os.path.basename(lib):24s} {name} n={n}: {1e3 * best:8.2f} us  {st['bytes'] / best / 1e6:8.1f} GB/s", flush=True)
        os._exit(0)
    os.wait()
