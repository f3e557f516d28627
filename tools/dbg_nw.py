import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_windows import _scene
for pv in (1, 2, 3, 10):
    print("per_visit", pv)
    for nw in (1, 1, 2, 2, 4, 8):
        objs, det, marks, eng = _scene("legacy")
        hist = []
        for s in range(6):
            c = eng.run_windows(1, proposals_per_visit=pv, n_warps=nw, t0=0.03, seed=9, sweep_offset=s)
            hist.append(tuple(c[:5]))
        print(nw, hist)
