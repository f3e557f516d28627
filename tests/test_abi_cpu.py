"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/mpp_b200.h declares (no
compute calls without a GPU), the ctypes mirror agrees with the compiled struct sizes, and the host-side logic of the
reference-facing API (term translation, kernel probabilities, mark classes, error behaviour) works."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpp_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from mpp_cnn_rs_object_detection_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 28
    assert sorted(_lib.SYMBOLS) == declared, set(declared) ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mpp_abi_version() == 1
    assert lib.mpp_abi_struct_size(0) == ctypes.sizeof(_lib.ModelParams)
    assert lib.mpp_abi_struct_size(1) == ctypes.sizeof(_lib.KernelParams)
    assert lib.mpp_abi_struct_size(2) == _lib.PROPOSAL_DTYPE.itemsize
    assert lib.mpp_abi_struct_size(3) == _lib.STEP_RESULT_DTYPE.itemsize


def test_argument_validation_without_gpu():
    """Calls that fail validation before touching CUDA return negative status codes with a message."""
    from mpp_cnn_rs_object_detection_b200 import _lib
    lib = _lib.load()
    assert lib.mpp_ctx_create(None, 0, 16, 16, 0, None) == _lib.ERR_INVALID
    ctx = ctypes.c_void_p()
    assert lib.mpp_ctx_create(ctypes.byref(ctx), 0, -1, 16, 0, None) == _lib.ERR_INVALID
    assert b"bad arguments" in lib.mpp_last_error()
    assert lib.mpp_set_model(None, None) == _lib.ERR_INVALID
    assert lib.mpp_ctx_destroy(None) == 0


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from mpp_cnn_rs_object_detection_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine((64, 64))


def test_kernel_probabilities_and_classes():
    from mpp_cnn_rs_object_detection_b200.engine import classes_of_marks, kernel_probabilities, pack_classes, snap_to_edges
    p = kernel_probabilities()
    np.testing.assert_allclose(p * 18, [1, 1, 2, 2, 2, 4, 2, 4], rtol=1e-15)  # make_kernels.py:13-24
    w = dict(bd_weight=2, uniform_bd_weight=1, data_bd_weight=1, ms_weight=1, translation_weight=1, gaussian_translation_weight=1,
             data_translation_weight=1, transformation_weight=1, gaussian_transformation_weight=1, data_transformation_weight=1)
    np.testing.assert_allclose(kernel_probabilities(weights=w), [1 / 8] * 8)
    edges = np.linspace(0, np.pi, 33)[:-1]
    c = classes_of_marks(np.array([[0.0, 0.0, 0.0], [31.999, 0.999, np.pi], [8.0, 0.5, edges[7]], [7.999999, 0.49999, np.nextafter(edges[7], 0)]]))
    np.testing.assert_array_equal(c, [[0, 0, 0], [31, 31, 31], [8, 16, 7], [7, 15, 6]])
    with pytest.raises(ValueError):
        classes_of_marks(np.array([[-1.0, 0.5, 0.1]]))
    assert pack_classes(np.array([[1, 2, 3]]))[0] == 1 | (2 << 8) | (3 << 16)
    # float32 read-back of bin edges snaps to the exact float64 edge (same class as the reference's float64 value)
    m32 = np.stack([np.float32(3.0), np.float32(edges[5] / np.pi), np.float32(edges[9])]).astype(np.float64)[None]
    snapped = snap_to_edges(m32)
    assert snapped[0, 2] == edges[9] and classes_of_marks(snapped)[0, 2] == 9


def test_term_translation_and_rejections():
    import mpp_cnn_rs_object_detection_b200.api as api
    from mpp_cnn_rs_object_detection_b200.api.device_state import apply_combinator, build_layout
    det = np.zeros((64, 64), np.float32)
    marks = [np.full((64, 64, 32), 1 / 32, np.float32)] * 3
    img = api.ImageWMaps("i", (64, 64), None, det, marks, api.default_mappings(), ["size", "ratio", "angle"])
    legacy = api.LegacyEnergySetup(calibration_params={}, energy_calibration=api.LegacyEnergiesCalibration(0.6, [1, 2, 3], [4, 5, 6], 10, 100))
    lay = build_layout(*legacy.make_energies(img))
    assert lay.spec.setup == "legacy" and lay.names == ["PositionEnergy", "ShapeEnergy", "AreaPriorEnergy", "RectangleOverlapEnergy", "ShapeAlignmentEnergy"]
    assert lay.names_by_column() == legacy.energy_names and lay.spec.remap_coefs == [1, 2, 3] and lay.max_dist == 32
    assert legacy.detection_threshold == 0.6
    nocal = api.NoCalibrationEnergySetup(ratio_prior=True)
    nocal.energy_calibration = api.NoCalibEnergiesCalibration(20, 60)
    lay2 = build_layout(*nocal.make_energies(img))
    assert lay2.spec.setup == "nocalib" and lay2.spec.ratio_prior and lay2.names_by_column() == nocal.energy_names
    assert nocal.detection_threshold == 0.5
    comb = api.LogisticEnergyCombinator(weights=np.arange(8.0), bias=0.5, energy_names=nocal.energy_names)
    spec = apply_combinator(lay2, comb)
    assert spec.combinator == "logistic" and spec.comb_w[:8] == list(np.arange(8.0)) and spec.comb_bias == 0.5
    h = api.HierarchicalEnergyCombinator(np.array([.8, .2]), np.array([.7, .1, .2]), np.array([.5, .5]), 0.0)
    assert apply_combinator(lay, h).comb_w[:7] == [.8, .2, .7, .1, .2, .5, .5]
    with pytest.raises(ValueError):
        apply_combinator(lay2, h)  # hierarchical needs the legacy names (hierarchical.py:22-29)
    toy = build_layout([api.ConstantUnitEnergy("u", -10.0)], [api.DistanceIndicatorPairEnergy("p", 3.0, 1.0, True)])
    assert toy.spec.setup == "toy" and toy.spec.toy_pair_strict and toy.names == ["u", "p"]

    class Plug(api.UnitEnergyConstructor):
        pass

    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        build_layout([Plug("x")], [])
    with pytest.raises(NotImplementedError):
        build_layout([], [api.DistanceIndicatorPairEnergy("p", 64.0)])
    with pytest.raises(AssertionError):
        build_layout([api.ConstantUnitEnergy("a", 1.0), api.ConstantUnitEnergy("a", 2.0)], [])
    kernels, p10 = api.make_kernels(img, 1.0, use_split_merge=True)  # optional split / merge kernels (make_kernels.py:145-166)
    assert len(kernels) == 10 and isinstance(kernels[8], api.SplitKernel) and isinstance(kernels[9], api.MergeKernel)
    np.testing.assert_allclose(p10 * 24, [1, 1, 2, 2, 2, 4, 2, 4, 3, 3], rtol=1e-15)
    with pytest.raises(NotImplementedError):
        api.sample_rjmcmc(img, np.random.default_rng(0), 1, None, None, 1.0, 0.99, 10, legacy, 1, 0.0, use_split_merge=True)


def test_shapes_and_mappings_match_golden():
    import mpp_cnn_rs_object_detection_b200.api as api
    from tests import golden_util as gu
    g = gu.load("geometry.npz")
    for k, r in enumerate(g["rects"][:200]):
        rect = api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4])
        np.testing.assert_allclose(rect.poly_coord, g["corners"][k], atol=1e-9)
        np.testing.assert_allclose([rect.length, rect.width], g["length_width"][k], rtol=1e-15)
    gm = gu.load("mappings.npz")
    for i, m in enumerate(api.default_mappings()):
        np.testing.assert_array_equal(m.feature_mapping, gm[f"edges_{i}"])
        np.testing.assert_array_equal([m.value_to_class(float(v)) for v in gm[f"values_{i}"]], gm[f"classes_{i}"])
        np.testing.assert_allclose([m.clip(float(v)) for v in gm[f"clip_in_{i}"]], gm[f"clip_out_{i}"], rtol=0, atol=0)
    a, b, w = api.polygon_to_abw(api.Rectangle(10, 20, 8, .5, .3).poly_coord)
    assert abs(a - 16 / 3) < 1e-9 and abs(b - 32 / 3) < 1e-9
    p = api.Point(1, 2)
    assert p != api.Point(1, 2) and hash(p) == id(p) and list(p.get_coord()) == [1, 2]
