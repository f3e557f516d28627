"""Pins the CPU oracle (oracle/mpp_oracle.py, oracle/mpp_oracle_c.c) against golden vectors produced by the
reference's own code (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import mpp_oracle as orc
from tests import golden_util as gu


def make_oracle_scene(cfg, det, marks):
    if cfg == "legacy":
        c = gu.CALIB_HRCM
        scene = orc.OracleScene(det, marks, setup="legacy", detection_threshold=c["detection_threshold"],
                                remap_coefs=c["coefs"], remap_intercepts=c["intercepts"], min_area=c["min_area"],
                                max_area=c["max_area"])
        comb = orc.OracleHierarchical(**gu.HRC)
    else:
        c = gu.CALIB_LOG
        scene = orc.OracleScene(det, marks, setup="nocalib", detection_threshold=0.0, min_area=c["min_area"],
                                max_area=c["max_area"], ratio_prior=True)
        comb = orc.OracleLogistic(weights=gu.LOG_WEIGHTS, bias=gu.LOG_BIAS, energy_names=orc.NOCALIB_NAMES)
    return scene, comb


def test_geometry_matches_reference():
    g = gu.load("geometry.npz")
    rects = g["rects"]
    for k, r in enumerate(rects):
        np.testing.assert_allclose(orc.rect_corners(*r), g["corners"][k], rtol=0, atol=1e-9)
        assert abs(orc.poly_area(orc.rect_corners(*r)) - g["areas"][k]) <= 1e-9 * max(1.0, g["areas"][k])
        np.testing.assert_allclose(orc.rect_length_width(r[2], r[3]), g["length_width"][k], rtol=1e-15)
    objs = [orc.ORect(*r) for r in rects]
    sc = orc.OracleScene(np.full((4, 4), 0.5, np.float32), [np.full((4, 4, 32), 1 / 32, np.float32)] * 3)
    sel = np.random.default_rng(0).choice(len(g["pairs"]), 3000, replace=False)
    for k in sel:
        i, j = g["pairs"][k]
        assert abs(sc.overlap_energy(objs[i], objs[j]) - g["overlap"][k]) < 1e-9
        assert abs(sc.align_energy(objs[i], objs[j]) - g["align_rewarding"][k]) < 1e-12


def test_c_and_python_clip_agree():
    g = gu.load("geometry.npz")
    lib = orc._load_clib()
    if not lib:
        pytest.skip("C oracle not built")
    rects = g["rects"]
    sel = np.random.default_rng(1).choice(len(g["pairs"]), 1500, replace=False)
    for k in sel:
        i, j = g["pairs"][k]
        p1, p2 = orc.rect_corners(*rects[i]), orc.rect_corners(*rects[j])
        inter = orc.convex_intersection_area(p1, p2)
        e = inter / (min(orc.poly_area(p1), orc.poly_area(p2)) + 1e-6)
        assert abs(e - g["overlap"][k]) < 1e-9


def test_mappings_match_reference():
    g = gu.load("mappings.npz")
    for i in range(3):
        np.testing.assert_array_equal(orc.mapping_edges(i), g[f"edges_{i}"])
        got = [orc.value_to_class(i, float(v)) for v in g[f"values_{i}"]]
        np.testing.assert_array_equal(got, g[f"classes_{i}"])
        got = [orc.mapping_clip(i, float(v)) for v in g[f"clip_in_{i}"]]
        np.testing.assert_allclose(got, g[f"clip_out_{i}"], rtol=0, atol=0)


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_energy_vectors_and_deltas_match_reference(cfg):
    g = gu.load(f"energies_{cfg}.npz")
    _, det, marks = gu.scene_inputs(g)
    scene, comb = make_oracle_scene(cfg, det, marks)
    assert list(scene.names) == [str(s) for s in g["names"]]
    config = [orc.ORect(*r) for r in g["config"]]
    state = orc.OracleState(scene, config)
    _, mat = state.energy_matrix()
    # energy_matrix iterates cell-major; reorder to the fixture's order
    order = {o.serial: k for k, o in enumerate(state.objects())}
    mat = mat[[order[o.serial] for o in config]]
    np.testing.assert_allclose(mat, g["vectors"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(orc.brute_force_energy_matrix(scene, config), g["vectors"], rtol=1e-12, atol=1e-12)
    # raw (combinator-less) sums accumulate the float32 PositionEnergy values with np.sum in *set iteration order*
    # (energy_graph.py:133), so the reference itself is only reproducible to float32 rounding there
    assert abs(state.total_energy() - float(g["raw_total"])) < 1e-4
    assert abs(state.subset_energy(config, comb) - float(g["comb_total"])) < 1e-9
    np.testing.assert_allclose(orc.per_object_combined(scene, comb, g["vectors"]), g["per_object_comb"], atol=1e-12)
    for k, (ri, add) in enumerate(zip(g["pert_removal"], g["pert_addition"])):
        rem = [config[ri]] if ri >= 0 else []
        addl = [] if np.isnan(add[0]) else [orc.ORect(*add)]
        assert abs(state.delta(rem, addl) - g["delta_raw"][k]) < 2e-5, k
        assert abs(state.delta(rem, addl, comb) - g["delta_comb"][k]) < 1e-9, k
    # naive detection (sample_rjmcmc.py:23-35)
    naive = orc.naive_detection(scene, float(g["naive_threshold"]))
    got = np.array([o.as_tuple() for o in naive]).reshape(-1, 5)
    np.testing.assert_allclose(got, g["naive"], rtol=0, atol=1e-12)
    # kernels
    np.testing.assert_allclose(orc.kernel_probabilities(), g["p_kernels"], rtol=0, atol=0)
    kern = orc.OracleKernels(scene, intensity=float(g["intensity"]))
    for row in g["kernel_probs"]:
        kid, ri = int(row[0]), int(row[1])
        add = None if np.isnan(row[2]) else orc.ORect(*row[2:7])
        prop = orc.Proposal(kid, removal=config[ri] if ri >= 0 else None, addition=add, delta=(row[7], row[8]),
                            param_id=int(row[9]), new_class=int(row[10]))
        f, b = kern.forward_backward(prop, len(config))
        assert abs(f - row[11]) <= 1e-13 * max(1.0, abs(row[11])), (kid, f, row[11])
        assert abs(b - row[12]) <= 1e-13 * max(1.0, abs(row[12])), (kid, b, row[12])


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_replay_reproduces_reference_chain(cfg):
    g = gu.load(f"stream_{cfg}.npz")
    _, det, marks = gu.scene_inputs(g)
    scene, comb = make_oracle_scene(cfg, det, marks)
    init = [orc.ORect(*r, uid=k) for k, r in enumerate(g["init"])]
    state = orc.OracleState(scene, init)
    kern = orc.OracleKernels(scene, intensity=float(g["intensity"]))
    by_uid = {o.uid: o for o in init}
    rec = g["records"]
    temp = float(g["t0"])
    n_steps = 1500  # the pure-python oracle replays ~500 steps/s; the GPU test replays the full stream
    for row in rec[:n_steps]:
        kid, ruid, auid = int(row[0]), int(row[1]), int(row[2])
        add = None if auid < 0 else orc.ORect(*row[3:8], uid=auid)
        prop = orc.Proposal(kid, removal=by_uid[ruid] if ruid >= 0 else None, addition=add,
                            delta=(row[8], row[9]), param_id=int(row[10]), new_class=int(row[11]))
        assert temp == row[13]
        d_e, fwd, bwd, la = orc.evaluate_proposal(state, kern, comb, prop, temp)
        assert abs(d_e - row[14]) <= 1e-9, (d_e, row[14])
        assert abs(fwd - row[15]) <= 1e-13 * max(1, abs(row[15]))
        assert abs(bwd - row[16]) <= 1e-13 * max(1, abs(row[16]))
        acc = orc.accept(row[12], la)
        assert acc == bool(row[17])
        if acc:
            state.apply([prop.removal] if prop.removal else [], [add] if add else [])
            if add is not None:
                by_uid[auid] = add
        assert len(state) == int(row[18])
        if temp > 0.0:
            temp *= float(g["alpha_t"])


def test_reference_known_answers():
    """Toy-energy known answers of test/test_energy_graph.py:177-244 and test/test_interacting_points_set.py:149-272
    as re-evaluated on the current reference code."""
    ka = gu.known_answers()
    assert ka["energy_graph_compute_delta"] == [-10.0, -8.0, -10.0, 1.0, -10.0, -8.0, 0.0, 7.0]
    assert ka["epointsset_total_energy"] == [5.0, 8.0, 6.0]
    assert ka["epointsset_energy_delta"] == [-1.0, 1.0]
    assert ka["pointsset_grid_200x516_r32"] == [7, 17]


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_split_merge_kernels_match_reference(cfg):
    """Optional 2-object moves (R24): Delta-energies of list perturbations and the kernels' forward / backward probabilities."""
    g = gu.load(f"split_merge_{cfg}.npz")
    _, det, marks = gu.scene_inputs(g)
    scene, comb = make_oracle_scene(cfg, det, marks)
    config = [orc.ORect(*r) for r in g["config"]]
    state = orc.OracleState(scene, config)
    p = orc.kernel_probabilities_split_merge()
    np.testing.assert_allclose(p, g["p_kernels"], rtol=1e-15)
    lam = float(g["intensity"])
    for row in g["rows"]:
        kid, r0, r1 = int(row[0]), int(row[1]), int(row[2])
        rem = [config[k] for k in (r0, r1) if k >= 0]
        add = [orc.ORect(*row[3 + 5 * j:8 + 5 * j]) for j in range(2) if not np.isnan(row[3 + 5 * j])]
        if kid == 8:
            f, b = orc.split_forward_backward(state, p[8], p[9], lam, add, row[13:15], row[15:18])
        else:
            f, b = orc.merge_forward_backward(state, p[8], p[9], lam, rem, int(row[18]))
        assert abs(f - row[21]) <= 1e-12 * abs(row[21]) and abs(b - row[22]) <= 1e-12 * abs(row[22]), (kid, f, row[21], b, row[22])
        if rem or add:
            assert abs(state.delta(rem, add) - row[19]) < 4e-5
            assert abs(state.delta(rem, add, comb) - row[20]) < 1e-9
