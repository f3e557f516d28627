"""Detection quality of the window sampler against proposals_per_visit on the validation-sized scenes (development tool)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpp_cnn_rs_object_detection_b200.api as api
from tests.test_gpu_configs import _val_image, _recall_precision, GOLD

cfg = json.load(open(os.path.join(GOLD, "model_mpp_hrcM", "config.json")))
for pid in (2781, 2789, 2794):
    try:
        img, objs = _val_image(api, pid)
    except KeyError:
        continue
    for pv in (64, 96, 128):
        rs, ps, ns = [], [], []
        for seed in range(4):
            model = api.MPPModel(cfg, model_dir=os.path.join(GOLD, "model_mpp_hrcM"))
            model.rng = np.random.default_rng(seed)
            res = model.infer_image(img, proposals_per_visit=pv)
            r, p = _recall_precision(res["detection_center"], objs)
            rs.append(r); ps.append(p); ns.append(len(res["detection_center"]))
        print(f"{pid} {img.shape} pv={pv}: recall {np.mean(rs):.3f} (min {min(rs):.3f})  precision {np.mean(ps):.3f} (min {min(ps):.3f})  found {np.mean(ns):.0f} / {len(objs)}", flush=True)
