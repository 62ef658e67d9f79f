import sys; sys.path.insert(0,'.')
import numpy as np
from lsqrrecipes_b200 import Engine, FP64, FP32, synth
data,_=synth.plane(4096,seed=1)
eng=Engine("plane3",0.5)
eng.upload(data)
print("score philox"); r=eng.score(count=1024,precision=FP64,seed=1); print(r["best_count"])
print("consensus"); c=eng.consensus(r["best_params"]); print(c)
print("mask"); m=eng.get_mask(); print(m.sum())
print("refine"); p=eng.refine(); print(p)
print("ransac"); out=eng.ransac(0.999,precision=FP64,seed=3); print(out["fraction"])
