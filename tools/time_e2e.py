import sys, os, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mpp_cnn_rs_object_detection_b200.api as api
from mpp_cnn_rs_object_detection_b200 import synth
dev = torch.device("cuda", 0)
h = w = 2048
objs, det, marks = synth.make_scene_torch(0, (h, w), 2600, dev)
det_h = torch.empty(det.shape, dtype=torch.float32, pin_memory=True).copy_(det)
marks_h = torch.empty(marks.shape, dtype=torch.float32, pin_memory=True).copy_(marks)
image = api.ImageWMaps("s", (h, w), None, det_h, [marks_h[0], marks_h[1], marks_h[2]], api.default_mappings(), ["size", "ratio", "angle"])
C, H = bench.CALIB_HRCM, bench.HRC
setup = api.LegacyEnergySetup(calibration_params={}, energy_calibration=api.LegacyEnergiesCalibration(C["detection_threshold"], list(C["coefs"]), list(C["intercepts"]), C["min_area"], C["max_area"]))
comb = api.HierarchicalEnergyCombinator(np.array(H["weights_data"]), np.array(H["weights_prior"]), np.array(H["data_prior_weights"]), 0.0, 0.0)
params = dict(num_samples=1, energy_combinator=comb, init_config="naive", init_temperature=0.02, alpha_t=1.0, burn_in=18 * 4096 * 32 - 3,
              energy_setup=setup, samples_interval=1, target_temperature=0.0, return_stats=True)
rng = np.random.default_rng(0)
api.sample_rjmcmc_batch([image] * 2, rng, **params)
torch.cuda.synchronize()
t = time.perf_counter(); api.sample_rjmcmc_batch([image] * 6, rng, **params); torch.cuda.synchronize()
print("per image ms", 1e3 * (time.perf_counter() - t) / 6)
pr = cProfile.Profile(); pr.enable(); api.sample_rjmcmc_batch([image] * 6, rng, **params); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
