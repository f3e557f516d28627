"""Shared helpers: load golden fixtures (tests/golden, written by oracle/gen_golden.py from the reference's
own code) and rebuild the seeded synthetic inputs they were generated from."""
import hashlib
import json
import os

import numpy as np

from mpp_cnn_rs_object_detection_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CALIB_HRCM = dict(detection_threshold=0.6464646464646465,
                  coefs=(40.61517366849468, 37.691647645515616, 35.287160617991965),
                  intercepts=(-3.4763386136080006, -2.389875684851162, -3.8677248427633804),
                  min_area=23.553573615517713, max_area=166.58205129586045)  # models_storage/mpp/mpp_hrcM/calibration.json
CALIB_LOG = dict(min_area=25.642491476585217, max_area=65.52715843309969)  # models_storage/mpp/mpp_log/calibration.json
# decoded from models_storage/mpp/mpp_log/energy_combination_model.pkl (float32 weights)
LOG_WEIGHTS = np.array([4.563805103302002, 0.19907087087631226, -0.27805596590042114, 2.1343348026275635,
                        7.730772495269775, 0.3594052493572235, 0.43729642033576965, 1.5640376806259155], dtype=np.float32)
LOG_BIAS = 0.7927545309066772
# decoded from models_storage/mpp/mpp_hrcM/energy_combination_model.pkl (float64)
HRC = dict(weights_data=(0.8, 0.2), weights_prior=(0.7058823529411764, 0.058823529411764705, 0.23529411764705882),
           data_prior_weights=(0.5, 0.5), detection_threshold=0.0, bias=0.0)


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def known_answers():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)


def checksum(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def scene_inputs(g):
    """(objects, det, marks) regenerated from the fixture's seed; asserts the stored checksum."""
    objs, det, marks = synth.make_scene(int(g["seed"]), tuple(int(v) for v in g["shape"]), int(g["n_rect"]))
    assert checksum(det, *marks) == str(g["maps_checksum"]), "synthetic inputs drifted from the golden fixture"
    return objs, det, marks
