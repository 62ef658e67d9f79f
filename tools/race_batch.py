import sys, numpy as np
sys.path.insert(0, "/root/repo")
from lsqrrecipes_b200 import Engine, synth
for name in ("sphere3", "circle2", "plane3"):
    data, _ = synth.GENERATORS[name](1600, seed=5)
    eng = Engine(name, 0.5, ls_type=1)
    r = eng.ransac_batch(data, np.arange(0, 9) * 200, max_tries=256)
    print(name, r["counts"])
    eng.close()
