"""Longer Monte-Carlo comparison of the device sequential chain and the parallel sweeps (GPU only; development tool)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpp_cnn_rs_object_detection_b200 import synth
from tests.gpu_util import make_engine


def batch_se(x, nb=40):
    x = np.asarray(x, dtype=np.float64)
    m = len(x) // nb
    b = x[:m * nb].reshape(nb, m).mean(1)
    return x.mean(), b.std(ddof=1) / np.sqrt(nb)


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "legacy"
    for temp, shape, nrect in ((0.3, (64, 96), 8), (0.1, (64, 96), 8), (0.15, (128, 128), 20)):
        objs, det, marks = synth.make_scene(11, shape, nrect)
        n0 = len(objs)
        eng = make_engine(cfg, det, marks, "fp32", intensity=max(1, n0))
        eng.add_objects(objs[:, :2], objs[:, 2:5])
        eng.run_chain(50000, t0=temp, seed=1)
        t = time.time()
        _, trace = eng.run_chain(2000000, t0=temp, seed=1, step_offset=50000, trace=True)
        dt = time.time() - t
        mb, sb = batch_se(trace["n_after"][::20])
        res = [f"chain {mb:.4f}+-{sb:.4f} acc {trace['accepted'].mean():.3f} ({2e6 / dt / 1e3:.0f}k steps/s)"]
        for stride, pv in ((3, 4), (3, 1), (4, 4)):
            e2 = make_engine(cfg, det, marks, "fp32", intensity=max(1, n0))
            e2.add_objects(objs[:, :2], objs[:, 2:5])
            e2.run_sweeps(1000, proposals_per_visit=pv, stride=stride, t0=temp, seed=2)
            nc = []
            for s in range(20000):
                e2.run_sweeps(1, proposals_per_visit=pv, stride=stride, t0=temp, seed=2, sweep_offset=1000 + s, read_counters=False)
                nc.append(len(e2))
            mc, sc = batch_se(nc)
            res.append(f"sweeps(stride {stride}, pv {pv}) {mc:.4f}+-{sc:.4f}")
        for nw, pv in ((4, 8), (1, 2)):
            e2 = make_engine(cfg, det, marks, "fp32", intensity=max(1, n0))
            e2.add_objects(objs[:, :2], objs[:, 2:5])
            e2.run_windows(1000, proposals_per_visit=pv, n_warps=nw, t0=temp, seed=2)
            nc = []
            for s in range(20000):
                e2.run_windows(1, proposals_per_visit=pv, n_warps=nw, t0=temp, seed=2, sweep_offset=1000 + s, read_counters=False)
                nc.append(len(e2))
            mc, sc = batch_se(nc)
            res.append(f"windows(nw {nw}, pv {pv}) {mc:.4f}+-{sc:.4f}")
        print(f"T={temp} shape={shape} n0={n0}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
