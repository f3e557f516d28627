import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_windows import _scene
from mpp_cnn_rs_object_detection_b200 import _lib
res = {}
for nw in (1, 2):
    objs, det, marks, eng = _scene("legacy")
    dbg = torch.zeros(8 + 4000 * 10, dtype=torch.float32, device=eng.device)
    cnt = (C.c_ulonglong * 8)()
    _lib.check(eng.lib.mpp_run_windows(eng.ctx, 1, 1, nw, 0.03, 1.0, 0.0, 9, 0, cnt, dbg.data_ptr()))
    d = dbg.cpu().numpy()
    n = int(d[1:2].view(np.int32)[0])
    tr = d[8:8 + n * 10].reshape(n, 10)
    tr = tr[np.lexsort((tr[:, 1], tr[:, 0]))]
    res[nw] = tr
    print(nw, list(cnt)[:5], n)
a, b = res[1], res[2]
print(a.shape, b.shape)
for i in range(min(len(a), len(b))):
    if not np.array_equal(a[i], b[i]):
        print("DIFF", i, a[i], b[i])
np.set_printoptions(linewidth=200, suppress=False)
print("cols: win it kernel r de la logu accept nc log_ratio")
for row in a:
    flag = "" if np.isfinite(row).all() else " <-- nonfinite"
    if abs(row[5] - row[6]) < 0.5 or flag or row[7] > 0:
        print(row, flag)
