import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mpp_cnn_rs_object_detection_b200.api as api
from mpp_cnn_rs_object_detection_b200 import synth
from mpp_cnn_rs_object_detection_b200.api import rjmcmc as R, device_state as DS
from mpp_cnn_rs_object_detection_b200 import engine as E

dev = torch.device("cuda", 0)
h = w = 2048
objs, det, marks = synth.make_scene_torch(0, (h, w), 2600, dev)
det_h = torch.empty(det.shape, dtype=torch.float32, pin_memory=True).copy_(det)
marks_h = torch.empty(marks.shape, dtype=torch.float32, pin_memory=True).copy_(marks)
image = api.ImageWMaps("s", (h, w), None, det_h, [marks_h[0], marks_h[1], marks_h[2]], api.default_mappings(), ["size", "ratio", "angle"])
C, H = bench.CALIB_HRCM, bench.HRC
setup = api.LegacyEnergySetup(calibration_params={}, energy_calibration=api.LegacyEnergiesCalibration(C["detection_threshold"], list(C["coefs"]), list(C["intercepts"]), C["min_area"], C["max_area"]))
comb = api.HierarchicalEnergyCombinator(np.array(H["weights_data"]), np.array(H["weights_prior"]), np.array(H["data_prior_weights"]), 0.0, 0.0)

def timed(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {name:28s} {1e3 * (time.perf_counter() - t):8.2f} ms", flush=True)
    return r

for rep in range(3):
    print("rep", rep)
    unit, pair = setup.make_energies(image)
    layout = timed("build_layout", lambda: DS.build_layout(unit, pair))
    eng = timed("Engine()", lambda: E.Engine((h, w), device=dev))
    maps = timed("device_maps (H2D)", lambda: DS.device_maps(layout.det, layout.marks, dev, reuse=False))
    tmp = torch.empty_like(marks)
    timed("raw copy_ 1.6GB pinned->dev", lambda: tmp.copy_(marks_h, non_blocking=True))
    timed("raw copy_ slice 0.5GB", lambda: tmp[0].copy_(marks_h[0], non_blocking=True))
    print("  pinned:", marks_h.is_pinned(), marks_h[0].is_pinned(), det_h.is_pinned())
    del tmp
    timed("set_maps", lambda: eng.set_maps(maps.det, maps.marks, det_sum=maps.det_sum))
    eng.set_model(DS.apply_combinator(layout, comb)); eng.set_kernels(intensity=2376)
    n = timed("naive_init", lambda: eng.naive_init(C["detection_threshold"], 6.0))
    timed("run_windows 18 sweeps", lambda: eng.run_windows(18, 32, 8, t0=0.02, seed=rep))
    r = timed("read_objects", lambda: eng.read_objects())
    timed("engine.close", lambda: eng.close())
    rng = np.random.default_rng(rep)
    timed("sample_rjmcmc total", lambda: api.sample_rjmcmc(image, rng, 1, comb, "naive", 0.02, 1.0, 18 * 4096 * 32 - 3, setup, 1, 0.0, reuse_device_maps=False))
