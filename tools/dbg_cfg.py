import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpp_cnn_rs_object_detection_b200.api as api
from tests import golden_util as gu
from tests.test_gpu_configs import _val_image, _recall_precision, GOLD
cfg = json.load(open(os.path.join(GOLD, "model_mpp_log", "config.json")))
setup = api.NoCalibrationEnergySetup(**cfg["energy_setup_params"])
setup.load_calibration(os.path.join(GOLD, "model_mpp_log"))
comb = api.LogisticEnergyCombinator(weights=gu.LOG_WEIGHTS, bias=gu.LOG_BIAS, energy_names=setup.energy_names)
img, objs = _val_image(api, 2781)
p = dict(cfg["inference"]["rjmcmc_params"])
for sampler, mult in (("parallel", 6), ("sequential", 6), ("parallel", 24), ("sequential", 24)):
    for seed in (0, 1):
        out = api.sample_rjmcmc(img, np.random.default_rng(seed), 1, comb, "naive", energy_setup=setup, iter_multiplier=mult, sampler=sampler, **p)
        c = [[r.x, r.y] for r in out[-1]]
        print(sampler, mult, seed, len(c), _recall_precision(c, objs), flush=True)
# energy of the ground truth vs of the result
unit, pair = setup.make_energies(img)
gt = api.EPointsSet(img.gt_config, img.shape, unit, pair)
print("GT energy", gt.energy_graph.compute_subset(list(gt), energy_combinator=comb), "n", len(gt))
res = api.EPointsSet(list(out[-1]), img.shape, unit, pair)
print("result energy", res.energy_graph.compute_subset(list(res), energy_combinator=comb), "n", len(res))
