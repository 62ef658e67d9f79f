#!/usr/bin/env python
"""R&D: times the fp64 validation kernel with every variant under tools/bin/variants/lib_fp64r*.so (hypotheses per thread)."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsqrrecipes_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "plane3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
H = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
for lib in sorted(glob.glob(os.path.join(ROOT, "tools/bin/variants/lib_fp64r*.so"))) + [api.lib_path()]:
    if os.fork() == 0:      # one process per variant: ctypes cannot unload a library
        api.lib_path = lambda lib=lib: lib
        from lsqrrecipes_b200 import FP64, Engine, synth
        data, _ = synth.GENERATORS[name](n)
        eng = Engine(name, synth.DELTAS[name])
        eng.upload(data)
        eng.score(count=H, precision=FP64, seed=1)
        best = min(eng.score(count=H, precision=FP64, seed=2 + i)["consensus_ms"] for i in range(3))
        print(f"{os.path.basename(lib):24s} {name} n={n} H={H}: {best:9.3f} ms  {n * H / best / 1e9:7.3f} T evals/s", flush=True)
        os._exit(0)
    os.wait()
