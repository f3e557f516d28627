"""-m gpu: the two named inference configurations (BASELINE configs[0:2]: mpp_hrcM and mpp_log) on the shipped validation
images' sizes and ground-truth annotations.  The posnet / shapenet weights are not shipped with the reference
(.MISSING_LARGE_BLOBS), so the position / mark maps are synthesised from the ground truth; the
sampler then has to recover the annotated vehicles.  Also edge cases of the sampler entry points."""
import json
import os

import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _val_image(api, pid):
    from mpp_cnn_rs_object_detection_b200 import synth
    ann = np.load(os.path.join(GOLD, "val_annotations.npz"))
    shape = tuple(int(v) for v in ann[f"shape_{pid}"])
    labels = {"centers": ann[f"centers_{pid}"], "parameters": ann[f"parameters_{pid}"]}
    gt = api.labels_to_rectangles(labels)
    # keep what the mark mappings can represent (size < 32 px) and the sampler's capacity assumptions
    gt = [r for r in gt if 0 < r.size < 32 and 0 < r.ratio <= 1]
    objs = np.array([[r.x, r.y, r.size, r.ratio, r.angle] for r in gt])
    det, marks = synth.make_maps(objs, shape, seed=pid)
    img = api.ImageWMaps(f"{pid:04}", shape, None, det, marks, api.default_mappings(), ["size", "ratio", "angle"], gt_config=gt)
    return img, objs


def _recall_precision(centers, objs, tol=3):
    if len(centers) == 0:
        return 0.0, 0.0
    d = np.abs(np.asarray(centers)[:, None, :] - objs[None, :, :2]).max(-1)
    return float((d.min(0) <= tol).mean()), float((d.min(1) <= tol).mean())


@pytest.mark.parametrize("pid,budget", [(2781, 1), (2789, 1), (2794, 1), (2781, 4)])
@pytest.mark.parametrize("name", ["mpp_hrcM", "mpp_log"])
def test_named_configs_on_val_annotations(name, pid, budget):
    """budget: multiple of the shipped burn-in (30 000 steps per 256^2 patch, model_configs/mpp/*.json)."""
    import mpp_cnn_rs_object_detection_b200.api as api
    cfg = json.load(open(os.path.join(GOLD, f"model_{name}", "config.json")))
    cfg["inference"]["rjmcmc_params"]["burn_in"] = int(cfg["inference"]["rjmcmc_params"]["burn_in"]) * budget
    if name == "mpp_hrcM":
        model = api.MPPModel(cfg, model_dir=os.path.join(GOLD, "model_mpp_hrcM"))  # manual weights from the config JSON
    else:
        setup = api.NoCalibrationEnergySetup(**cfg["energy_setup_params"])
        model = api.MPPModel.__new__(api.MPPModel)
        model.config, model.rng, model.save_path, model.energy_setup = cfg, np.random.default_rng(0), None, setup
        setup.load_calibration(os.path.join(GOLD, "model_mpp_log"))
        model.energy_model = api.LogisticEnergyCombinator(weights=gu.LOG_WEIGHTS, bias=gu.LOG_BIAS, energy_names=setup.energy_names)
    # the shipped schedule: T0 = 1, alpha = 0.999 per step, 30 000 burn-in steps per 256^2 patch (iter_multiplier = #patches)
    img, objs = _val_image(api, pid)
    res = model.infer_image(img)
    recall, precision = _recall_precision(res["detection_center"], objs)
    print(f"\n{name} on {pid} {img.shape} budget x{budget}: {len(objs)} annotated, {len(res['detection_center'])} found, recall {recall:.3f}, precision {precision:.3f}")
    # the dense parking lot of image 2781 (272 vehicles on 469x753) is not fully recovered within the reference's budget by
    # EITHER sampler (parallel 0.78-0.84, device sequential chain 0.76-0.81); with 4x the budget both exceed 0.9.  The two
    # sparser images are recovered at the shipped budget.
    bar = 0.7 if (pid == 2781 and budget == 1) else 0.9
    assert recall > bar and precision > 0.9, (recall, precision, bar)
    assert len(res["detection_score"]) == len(res["detection_center"])


def test_edge_cases():
    import mpp_cnn_rs_object_detection_b200.api as api
    from mpp_cnn_rs_object_detection_b200 import synth
    from tests.gpu_util import make_engine
    # empty scene, image smaller than one window, sizes that are not multiples of 32
    for shape in ((20, 17), (33, 95), (64, 64)):
        det, marks = synth.make_maps(np.zeros((0, 5)), shape, seed=1)
        eng = make_engine("legacy", det, marks, "fp32", intensity=1.0)
        assert len(eng) == 0
        vec, comb, raw, tot = eng.energy_vectors([])
        assert vec.shape[0] == 0 and raw == 0.0 and tot == 0.0
        assert eng.delta_batch(np.zeros(0, dtype=eng.sample_proposals([]).dtype)).shape == (0,)
        c = eng.run_windows(30, 8, 4, t0=0.02, seed=1)
        assert c[0] > 0 and c[4] > 0 and c[3] == 0 or c[2] >= c[3]  # births only from an empty configuration
        assert 0 <= len(eng) <= 8  # background only: (almost) nothing to detect
        c2, tr = eng.run_chain(200, t0=0.02, seed=2, trace=True)
        assert len(tr) == 200 and c2[0] == 200
        _, xy, mk, uid = eng.read_objects()
        assert np.all((xy[:, 0] >= 0) & (xy[:, 0] < shape[0]) & (xy[:, 1] >= 0) & (xy[:, 1] < shape[1])) if len(xy) else True
    # sample_rjmcmc with no initial configuration and with an explicit list
    objs, det, marks = synth.make_scene(3, (96, 96), 12)
    c = gu.CALIB_HRCM
    setup = api.LegacyEnergySetup(calibration_params={}, energy_calibration=api.LegacyEnergiesCalibration(
        c["detection_threshold"], list(c["coefs"]), list(c["intercepts"]), c["min_area"], c["max_area"]))
    h = gu.HRC
    comb = api.HierarchicalEnergyCombinator(np.array(h["weights_data"]), np.array(h["weights_prior"]), np.array(h["data_prior_weights"]), 0.0, 0.0)
    img = api.ImageWMaps("e", det.shape, None, det, marks, api.default_mappings(), ["size", "ratio", "angle"],
                         gt_config=[api.Rectangle(int(o[0]), int(o[1]), o[2], o[3], o[4]) for o in objs])
    for init in (None, "gt", img.gt_config[:3]):
        out = api.sample_rjmcmc(img, np.random.default_rng(0), 2, comb, init, 0.02, 1.0, 12000, setup, 500, 0.0)
        assert len(out) == 2 and all(isinstance(r, api.Rectangle) for ps in out for r in ps)
        rc, pr = _recall_precision([[r.x, r.y] for r in out[-1]], objs)
        assert rc >= 0.7 and pr >= 0.7, (init if not isinstance(init, list) else "list", rc, pr)
    # batch / tiles entry points agree in kind with the single-image one
    tiles = api.sample_rjmcmc_tiles([img, img, img], np.random.default_rng(1), n_streams=2, num_samples=1, energy_combinator=comb,
                                    init_config="naive", init_temperature=0.02, alpha_t=1.0, burn_in=3000, energy_setup=setup,
                                    samples_interval=100, target_temperature=0.0)
    assert len(tiles) == 3 and all(_recall_precision([[r.x, r.y] for r in t], objs)[0] > 0.7 for t in tiles)
    batch = api.sample_rjmcmc_batch([img, img], np.random.default_rng(2), with_scores=True, num_samples=1, energy_combinator=comb,
                                    init_config="naive", init_temperature=0.02, alpha_t=1.0, burn_in=3000, energy_setup=setup,
                                    samples_interval=100, target_temperature=0.0)
    assert len(batch) == 2 and len(batch[0][1]) == len(batch[0][0][0]) and np.all(batch[0][1] > 0)
    # the pipelined batch (upload / chain / read-back of consecutive images overlapped) returns what one call per image returns
    kw = dict(num_samples=1, energy_combinator=comb, init_config="naive", init_temperature=0.02, alpha_t=1.0, burn_in=3000, energy_setup=setup,
              samples_interval=100, target_temperature=0.0)
    imgs = [img, api.ImageWMaps("f", det.shape, None, det.copy(), [m.copy() for m in marks], api.default_mappings(), ["size", "ratio", "angle"]), img]
    piped = api.sample_rjmcmc_batch(imgs, np.random.default_rng(7), return_stats=True, **kw)
    rng1 = np.random.default_rng(7)
    for (rects, stats), im in zip(piped, imgs):
        one = api.sample_rjmcmc(im, rng1, **kw)
        key = lambda r: (r.x, r.y, r.size, r.ratio, r.angle)
        assert sorted(map(key, rects[-1])) == sorted(map(key, one[-1]))
        assert stats["proposals"] > 3000 and stats["evaluated"] <= stats["proposals"] and stats["launches"] >= 1
    # capacity errors are reported, not silently ignored: 33 objects in one 32x32 cell
    from mpp_cnn_rs_object_detection_b200._lib import ERR_CELL_FULL, MPPError
    eng = make_engine("legacy", det, marks, "fp32")
    xy = np.stack([np.arange(33) % 32, np.arange(33) // 32], 1)
    with pytest.raises(MPPError) as e:
        eng.add_objects(xy, np.tile([8.0, 0.5, 0.1], (33, 1)))
    assert e.value.code == ERR_CELL_FULL
