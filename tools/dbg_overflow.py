import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mpp_cnn_rs_object_detection_b200 import synth
from mpp_cnn_rs_object_detection_b200 import engine as E
dev = torch.device("cuda", 0)
h = w = 2048
objs, det, marks = synth.make_scene_torch(0, (h, w), 2600, dev)
C, H = bench.CALIB_HRCM, bench.HRC
spec = E.ModelSpec(setup="legacy", pos_threshold=C["detection_threshold"], remap_coefs=C["coefs"], remap_intercepts=C["intercepts"],
                 min_area=C["min_area"], max_area=C["max_area"], combinator="hierarchical",
                 comb_w=list(H["weights_data"]) + list(H["weights_prior"]) + list(H["data_prior_weights"]) + [0.0])
for init in ("naive", "gt", "naive"):
    for sched in ("colours", "dataflow"):
        for seed in (0, 1):
            eng = E.Engine((h, w), device=dev)
            eng.set_maps(det, marks); eng.set_model(spec)
            if init == "gt":
                eng.add_objects(objs[:, :2], objs[:, 2:5])
            else:
                eng.naive_init(C["detection_threshold"], 6.0)
            n0 = len(eng)
            eng.set_kernels(intensity=max(1, n0))
            hh, xy, mk, uid = eng.read_objects()
            cells = (xy[:, 0] // 32) * 64 + xy[:, 1] // 32
            try:
                for s in range(1):
                    c = eng.run_windows(18, 32, 8, t0=0.02, seed=seed, sweep_offset=s, schedule=sched)
                print(init, sched, seed, "ok n0", n0, "n", len(eng), "max per cell at start", np.bincount(cells).max(), c[:5])
            except Exception as ex:
                print(init, sched, seed, "FAILED at sweep", s, "n0", n0, "n", len(eng), str(ex)[:80])
            eng.close()
