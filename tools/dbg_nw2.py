import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_windows import _scene
for nw in (1, 2):
    objs, det, marks, eng = _scene("legacy")
    c, md = eng.run_windows(1, proposals_per_visit=1, n_warps=nw, t0=0.03, seed=9, sweep_offset=0, debug=True)
    print(nw, c[:5], md)
