"""Helpers for the -m gpu parity tests: build a device Engine for a golden case."""
import numpy as np

from mpp_cnn_rs_object_detection_b200 import _lib
from mpp_cnn_rs_object_detection_b200.engine import Engine, ModelSpec, classes_of_marks, pack_classes
from tests import golden_util as gu


def model_spec(cfg, combinator=True):
    if cfg == "legacy":
        c = gu.CALIB_HRCM
        spec = ModelSpec(setup="legacy", pos_threshold=c["detection_threshold"], remap_coefs=c["coefs"],
                         remap_intercepts=c["intercepts"], min_area=c["min_area"], max_area=c["max_area"])
        if combinator:
            h = gu.HRC
            spec.combinator = "hierarchical"
            spec.comb_w = list(h["weights_data"]) + list(h["weights_prior"]) + list(h["data_prior_weights"]) + [0.0]
            spec.comb_bias, spec.comb_threshold = h["bias"], h["detection_threshold"]
    else:
        c = gu.CALIB_LOG
        spec = ModelSpec(setup="nocalib", pos_threshold=0.0, min_area=c["min_area"], max_area=c["max_area"], ratio_prior=True)
        if combinator:
            spec.combinator = "logistic"
            spec.comb_w = [float(v) for v in gu.LOG_WEIGHTS]
            spec.comb_bias = gu.LOG_BIAS
    return spec


def make_engine(cfg, det, marks, precision="fp32", combinator=True, intensity=1.0):
    eng = Engine(det.shape, precision=precision)
    eng.set_maps(det, marks)
    eng.set_model(model_spec(cfg, combinator))
    eng.set_kernels(intensity=intensity)
    return eng


def proposals_from_rows(kernel, rem_xy_uid, add_rows, add_uid, delta, param_id, new_class, u):
    """Builds a PROPOSAL_DTYPE array; rem_xy_uid rows (x, y, uid) with uid < 0 for none; add_rows (x,y,size,ratio,angle)
    with NaN x for none."""
    m = len(kernel)
    p = np.zeros(m, dtype=_lib.PROPOSAL_DTYPE)
    p["kernel"] = kernel
    for i in range(m):
        if rem_xy_uid[i][2] >= 0:
            p["rem_x"][i], p["rem_y"][i], p["rem_uid"][i] = rem_xy_uid[i]
        else:
            p["rem_uid"][i] = _lib.NO_OBJECT
        a = add_rows[i]
        if not np.isnan(a[0]):
            p["add_x"][i], p["add_y"][i] = int(a[0]), int(a[1])
            p["add_size"][i], p["add_ratio"][i], p["add_angle"][i] = a[2], a[3], a[4]
            p["add_cls"][i] = pack_classes(classes_of_marks(np.array(a[2:5])))[0]
            p["add_uid"][i] = add_uid[i]
        else:
            p["add_uid"][i] = _lib.NO_OBJECT
    p["delta0"], p["delta1"] = delta[:, 0], delta[:, 1]
    p["param_id"], p["new_class"], p["u"] = param_id, new_class, u
    return p
