"""Times mpp_run_windows on the bench scene for several (n_warps, proposals_per_visit) settings (development tool)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mpp_cnn_rs_object_detection_b200 import synth
from mpp_cnn_rs_object_detection_b200.engine import Engine, ModelSpec

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
temp = 0.02
dev = torch.device("cuda", 0)
n_rect = int(round(2600 * size * size / 2048.0 ** 2))
objs, det, marks = synth.make_scene_torch(0, (size, size), n_rect, dev)
C, H = bench.CALIB_HRCM, bench.HRC
spec = ModelSpec(setup="legacy", pos_threshold=C["detection_threshold"], remap_coefs=C["coefs"], remap_intercepts=C["intercepts"],
                 min_area=C["min_area"], max_area=C["max_area"], combinator="hierarchical",
                 comb_w=list(H["weights_data"]) + list(H["weights_prior"]) + list(H["data_prior_weights"]) + [0.0])
configs = ((1, 32, "dataflow"), (2, 32, "dataflow"), (4, 32, "dataflow"), (8, 32, "dataflow"), (2, 16, "dataflow"), (4, 16, "dataflow")) if size > 2048 else \
    ((4, 16, "colours"), (4, 32, "colours"), (1, 16, "dataflow"), (2, 16, "dataflow"), (4, 16, "dataflow"), (4, 32, "dataflow"), (8, 16, "dataflow"),
     (8, 32, "dataflow"), (8, 64, "dataflow"), (2, 32, "dataflow"))
if os.environ.get("MPP_TUNE"):
    configs = tuple((int(a.split(':')[0]), int(a.split(':')[1]), "dataflow") for a in os.environ["MPP_TUNE"].split(','))
for nw, pv, sched in configs:
    eng = Engine((size, size), device=dev)
    eng.set_maps(det, marks); eng.set_model(spec); eng.set_kernels(intensity=max(1, len(objs)))
    eng.add_objects(objs[:, :2], objs[:, 2:5])
    eng.run_windows(5, pv, nw, t0=temp, seed=1, schedule=sched)
    sweeps = 20 if size <= 2048 else 4
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_windows(sweeps, pv, nw, t0=temp, seed=1, sweep_offset=5, read_counters=False, schedule=sched)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = eng.run_windows(0, pv, nw, t0=temp)
    print(f"{sched} nw={nw} pv={pv}: {ms / sweeps * 1e3:.0f} us/sweep, attempted {c[0] / ms / 1e3:.1f} M/s, evaluated {c[4] / ms / 1e3:.1f} M/s, "
          f"acc {c[1] / max(1, c[4]):.3f}, n={len(eng)}", flush=True)
    eng.close()
