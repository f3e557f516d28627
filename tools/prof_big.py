"""One mpp_run_windows call on a large scene (profiling target; development tool)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mpp_cnn_rs_object_detection_b200 import synth
from mpp_cnn_rs_object_detection_b200.engine import Engine, ModelSpec
size, nw, pv = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
objs, det, marks = synth.make_scene_torch(0, (size, size), int(round(2600 * size * size / 2048.0 ** 2)), dev)
C, H = bench.CALIB_HRCM, bench.HRC
spec = ModelSpec(setup="legacy", pos_threshold=C["detection_threshold"], remap_coefs=C["coefs"], remap_intercepts=C["intercepts"],
                 min_area=C["min_area"], max_area=C["max_area"], combinator="hierarchical",
                 comb_w=list(H["weights_data"]) + list(H["weights_prior"]) + list(H["data_prior_weights"]) + [0.0])
eng = Engine((size, size), device=dev)
eng.set_maps(det, marks); eng.set_model(spec); eng.set_kernels(intensity=max(1, len(objs)))
eng.add_objects(objs[:, :2], objs[:, 2:5])
eng.run_windows(2, pv, nw, t0=0.02, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); c = eng.run_windows(3, pv, nw, t0=0.02, seed=1, sweep_offset=2); e1.record(); torch.cuda.synchronize()
print(f"size {size} nw {nw} pv {pv}: {c[4] / e0.elapsed_time(e1) / 1e3:.1f} M proposals/s")
