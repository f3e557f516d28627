"""CPU tests of the inference glue (api/mpp_model.py): restricted unpickling of the reference's model files, the manual
combinator from the config JSON, config resolution, annotation conversion and cropping."""
import json
import os
import pickle
import sys
import types

import numpy as np
import pytest

import mpp_cnn_rs_object_detection_b200.api as api
from tests import golden_util as gu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _reference_style_pickle(module, cls_name, fields):
    """Pickle bytes as the reference writes them: a dataclass instance of `module.cls_name` (mpp_model.py:199-200)."""
    mod = types.ModuleType(module)
    cls = type(cls_name, (), {"__module__": module})
    setattr(mod, cls_name, cls)
    parts = module.split(".")
    created = []
    for i in range(1, len(parts) + 1):
        name = ".".join(parts[:i])
        if name not in sys.modules:
            sys.modules[name] = mod if i == len(parts) else types.ModuleType(name)
            created.append(name)
    sys.modules[module] = mod
    try:
        obj = cls()
        obj.__dict__.update(fields)
        return pickle.dumps(obj, protocol=4)
    finally:
        for name in created + [module]:
            sys.modules.pop(name, None)


def test_restricted_unpickler_loads_combinators_and_refuses_code(tmp_path):
    data = _reference_style_pickle("models.mpp.energies.combination.logistic", "LogisticEnergyCombinator",
                                   dict(weights=gu.LOG_WEIGHTS, bias=gu.LOG_BIAS, energy_names=list(api.NoCalibrationEnergySetup(ratio_prior=True).energy_names)))
    p = tmp_path / "energy_combination_model.pkl"
    p.write_bytes(data)
    m = api.load_energy_combination_model(str(p))
    assert isinstance(m, api.LogisticEnergyCombinator) and m.bias == gu.LOG_BIAS
    np.testing.assert_array_equal(m.weights, gu.LOG_WEIGHTS)
    data = _reference_style_pickle("models.mpp.energies.combination.hierarchical", "HierarchicalEnergyCombinator",
                                   dict(weights_data=np.array([.8, .2]), weights_prior=np.array(gu.HRC["weights_prior"]),
                                        data_prior_weights=np.array([.5, .5]), detection_threshold=0.0, bias=0.0))
    m = api.restricted_load(data)
    assert isinstance(m, api.HierarchicalEnergyCombinator) and list(m.weights_data) == [.8, .2]

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned",))

    with pytest.raises(pickle.UnpicklingError):
        api.restricted_load(pickle.dumps(Evil()))
    with pytest.raises(pickle.UnpicklingError):
        p.write_bytes(pickle.dumps({"not": "a model"}))
        api.load_energy_combination_model(str(p))


def test_manual_config_and_model_construction():
    cfg = json.load(open(os.path.join(GOLD, "model_mpp_hrcM", "config.json")))
    comb = api.combinator_from_manual_config(cfg)
    np.testing.assert_allclose(comb.weights_data, gu.HRC["weights_data"])
    np.testing.assert_allclose(comb.weights_prior, gu.HRC["weights_prior"], rtol=1e-15)
    np.testing.assert_allclose(comb.data_prior_weights, gu.HRC["data_prior_weights"])
    assert comb.detection_threshold == 0.0
    model = api.MPPModel(cfg, phase="val", load=True, model_dir=os.path.join(GOLD, "model_mpp_hrcM"))
    assert isinstance(model.energy_setup, api.LegacyEnergySetup)
    cal = model.energy_setup.energy_calibration
    assert abs(cal.detection_threshold - gu.CALIB_HRCM["detection_threshold"]) < 1e-15 and cal.min_area == gu.CALIB_HRCM["min_area"]
    assert tuple(cal.param_dist_remap_coefs) == gu.CALIB_HRCM["coefs"]
    assert model.config["inference"]["rjmcmc_params"] == dict(samples_interval=128, init_temperature=1, target_temperature=0.0,
                                                              alpha_t=0.999, burn_in=30000)
    cfg_log = json.load(open(os.path.join(GOLD, "model_mpp_log", "config.json")))
    with pytest.raises(FileNotFoundError):  # mpp_log has learned weights: needs its energy_combination_model.pkl
        api.MPPModel(cfg_log, model_dir=os.path.join(GOLD, "model_mpp_log"))
    with pytest.raises(NotImplementedError):
        api.MPPModel(cfg, phase="train", load=False)
    assert api.resolve_model_config_path("model_mpp_hrcM", [GOLD]).endswith(os.path.join("model_mpp_hrcM", "config.json"))
    with pytest.raises(FileNotFoundError):
        api.resolve_model_config_path("nope", [GOLD])


def test_labels_and_cropping():
    labels = {"centers": np.array([[10, 20], [300, 40]]), "parameters": np.array([[4.0, 8.0, 0.3], [5.0, 10.0, 3.5]]),
              "categories": np.array(["small-vehicle", "large-vehicle"]), "difficult": np.array([False, True])}
    rects = api.labels_to_rectangles(labels)
    assert (rects[0].x, rects[0].y, rects[0].size, rects[0].ratio) == (10, 20, 6.0, 0.5) and abs(rects[1].angle - (3.5 % np.pi)) < 1e-15
    a, b, w = api.sra_to_wla(rects[0].size, rects[0].ratio, rects[0].angle)
    assert (a, b) == (4.0, 8.0)
    det = np.arange(400 * 300, dtype=np.float32).reshape(400, 300)
    marks = [np.zeros((400, 300, 32), np.float32) for _ in range(3)]
    img = api.ImageWMaps("7", (400, 300), np.zeros((400, 300, 3)), det, marks, api.default_mappings(), ["size", "ratio", "angle"],
                         labels=labels, gt_config=rects)
    crop = api.crop_image_w_maps(img, np.array([144, 0]), 256)
    assert crop.shape == (256, 256) and crop.detection_map[0, 0] == det[144, 0] and crop.param_dist_maps[0].shape == (256, 256, 32)
    assert len(crop.gt_config) == 1 and (crop.gt_config[0].x, crop.gt_config[0].y) == (156, 40)
    assert list(crop.crop_data["tl_anchor"]) == [144, 0]
