"""-m gpu: the reference-facing Python API (mpp_cnn_rs_object_detection_b200.api) on the device.  The first three groups
restate the reference's own unit tests (test/test_points_set.py, test/test_energy_graph.py,
test/test_interacting_points_set.py) with the toy terms expressed as device terms; the rest checks the real terms against
the golden vectors produced by the reference."""
import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _api():
    import mpp_cnn_rs_object_detection_b200.api as api
    return api


# ------------------------------------------------------------------------------------------------ test/test_points_set.py
def test_points_set_grid_and_membership():
    api = _api()
    ps = api.PointsSet(support_shape=(200, 516), maximum_interaction_radius=32)
    assert ps._n_x == 7 and ps._n_y == 17  # test/test_points_set.py:29-41
    pts = [api.Point(0, 0), api.Point(5, 7), api.Point(199, 515), api.Point(64, 64), api.Point(64, 64)]
    for p in pts:
        ps.add(p)
    assert len(ps) == 5
    for p in pts:
        assert p in ps
        assert p in ps._find_local_point_set(p)
    assert api.Point(0, 0) not in ps  # identity, not value
    assert set(iter(ps)) == set(pts)
    ps.remove(pts[1])
    assert len(ps) == 4 and pts[1] not in ps
    with pytest.raises(KeyError):
        ps.remove(pts[1])
    with pytest.raises(AssertionError):
        ps.add(api.Point(200, 0))  # out of bounds (point_set.py:99)
    cp = ps.copy()
    cp.remove(pts[0])
    assert len(cp) == 3 and len(ps) == 4 and pts[0] in ps


def test_points_set_neighbours_vs_brute_force():
    api = _api()
    rng = np.random.default_rng(0)
    shape = (300, 260)
    ps = api.PointsSet(shape, 32)
    pts = [api.Point(int(x), int(y)) for x, y in zip(rng.integers(0, shape[0], 600), rng.integers(0, shape[1], 600))]
    for p in pts:
        ps.add(p)
    for u in pts[:40]:
        for r in (8, 32, 64):  # r=64: two cell offsets (test/test_points_set.py:222-244)
            got = ps.get_neighbors(u, r, suppress_warnings=True)
            want = {p for p in pts if p is not u and np.hypot(p.x - u.x, p.y - u.y) <= r}
            assert got == want
        pot = ps.get_potential_neighbors(u, 32)
        for p in pot:  # Chebyshev bound of the 3x3-cell block (test/test_points_set.py:135-157)
            assert abs(p.x // 32 - u.x // 32) <= 1 and abs(p.y // 32 - u.y // 32) <= 1
        want_pot = {p for p in pts if p is not u and abs(p.x // 32 - u.x // 32) <= 1 and abs(p.y // 32 - u.y // 32) <= 1}
        assert pot == want_pot
    # churn (test/test_points_set.py:160-185)
    alive = set(pts)
    for k in range(2000):
        if alive and rng.random() < 0.5:
            p = next(iter(alive))
            alive.remove(p)
            ps.remove(p)
        else:
            p = api.Point(int(rng.integers(0, shape[0])), int(rng.integers(0, shape[1])))
            alive.add(p)
            ps.add(p)
    assert len(ps) == len(alive) and set(iter(ps)) == alive
    c = ps.random_choice(np.random.default_rng(1))
    assert c in alive


# ------------------------------------------------------------------------------------------------ test/test_energy_graph.py
def _toy_graph(api):
    eg = api.EnergyGraph(unit_energies_constructors=[api.ConstantUnitEnergy(name="unit", value=-10.0)],
                         pair_energies_constructors=[api.DistanceIndicatorPairEnergy(name="pair", max_dist=1.0, value=1.0)])
    ps = api.PointsSet(support_shape=(64, 64), maximum_interaction_radius=32)
    return eg, ps


def test_energy_graph_structure():
    api = _api()
    eg, ps = _toy_graph(api)
    p1 = api.Point(10, 10)
    ps.add(p1); eg.add_point(p1, ps)
    assert p1 in eg.ue_per_point and len(eg.ue_per_point[p1]) == 1
    assert p1 in eg.pe_per_point and len(eg.pe_per_point[p1]) == 0
    p2 = api.Point(10, 11)
    ps.add(p2); eg.add_point(p2, ps)
    assert len(eg.ue_per_point[p2]) == 1 and len(eg.pe_per_point[p2]) == 1 and len(eg.pe_per_point[p1]) == 1
    assert eg.pe_per_point[p2][0].point_2 is p1
    assert eg.pe_per_point[p2][0] is eg.pe_per_point[p1][0]
    assert eg.pe_per_point[p2][0].compute() == 1.0 and eg.ue_per_point[p2][0].compute() == -10.0
    p3 = api.Point(20, 20)
    ps.add(p3); eg.add_point(p3, ps)
    assert len(eg.pe_per_point[p3]) == 0 and len(eg.pe_per_point[p2]) == 1 and len(eg.pe_per_point[p1]) == 1
    ps.remove(p2); eg.remove_point(p2)
    assert len(eg.pe_per_point[p3]) == 0 and len(eg.pe_per_point[p1]) == 0
    assert p2 not in eg.pe_per_point


def test_energy_graph_compute_delta_known_answers():
    """test/test_energy_graph.py:177-244: -10, -8, -10, +1, -10, -8, 0, +7."""
    api = _api()
    eg, ps = _toy_graph(api)
    P, B, D, T = api.Perturbation, api.BirthKernel, api.DeathKernel, api.DataDrivenTranslationKernel
    want = gu.known_answers()["energy_graph_compute_delta"]
    got = []

    def add(p):
        ps.add(p); eg.add_point(p, ps)

    p1 = api.Point(10, 10)
    got.append(eg.compute_delta(ps, P(type=B, removal=None, addition=p1))); add(p1)
    p2 = api.Point(10, 11)
    got.append(eg.compute_delta(ps, P(type=B, removal=None, addition=p2))); add(p2)
    p3 = api.Point(20, 20)
    got.append(eg.compute_delta(ps, P(type=B, removal=None, addition=p3))); add(p3)
    p32 = api.Point(10, 12)
    got.append(eg.compute_delta(ps, P(type=T, removal=p3, addition=p32)))
    ps.remove(p3); eg.remove_point(p3); add(p32)
    p4 = api.Point(5, 5)
    got.append(eg.compute_delta(ps, P(type=B, removal=None, addition=p4))); add(p4)
    p5 = api.Point(5, 6)
    got.append(eg.compute_delta(ps, P(type=B, removal=None, addition=p5))); add(p4)
    p6 = api.Point(5, 7)
    add(p6)
    p61 = api.Point(5, 8)
    got.append(eg.compute_delta(ps, P(type=T, removal=p6, addition=p61)))
    ps.remove(p6); eg.remove_point(p6); add(p61)
    got.append(eg.compute_delta(ps, P(type=D, removal=p2, addition=None)))
    assert got == want
    # total energy / subsets (test/test_energy_graph.py:94-174)
    total = eg.total_energy(ps)
    vec = eg.compute_subset(list(ps), return_vector=True)
    assert set(vec) == {"unit", "pair"} and total == sum(vec["unit"]) + sum(vec["pair"])


# ------------------------------------------------------------------------------------------------ test/test_interacting_points_set.py
def _toy_eps(api, pts, shape=(10, 10)):
    return api.EPointsSet(points=pts, support_shape=shape, unit_energies_constructors=[api.ConstantUnitEnergy(name="ue", value=1.0)],
                          pair_energies_constructors=[api.DistanceIndicatorPairEnergy(name="pe", max_dist=3, value=1.0, strict=True)])


def test_epointsset_known_answers():
    api = _api()
    ka = gu.known_answers()
    pts = [api.Point(0, 0), api.Point(0, 1), api.Point(0, 4)]
    s = _toy_eps(api, pts)
    e = [s.total_energy()]
    s.add(api.Point(0, 5))
    e.append(s.total_energy())
    e.append(_toy_eps(api, [api.Point(0, 0), api.Point(0, 1), api.Point(1, 0)]).total_energy())
    assert e == ka["epointsset_total_energy"]  # 5, 8, 6 (test/test_interacting_points_set.py:149-208)
    some = [api.Point(0, 0), api.Point(0, 1), api.Point(0, 5)]
    p0 = _toy_eps(api, some)
    pert1 = api.Perturbation(type=api.DeathKernel, removal=some[2])
    e0 = p0.total_energy()
    d1 = p0.energy_delta(pert1)
    p1 = p0.apply_perturbation(pert1)
    assert e0 == 5.0 and p1.total_energy() == e0 + d1 and len(p0) == 3 and len(p1) == 2
    pert2 = api.Perturbation(type=api.DataDrivenTranslationKernel, removal=some[2], addition=api.Point(1, 0))
    d2 = p0.energy_delta(pert2)
    p2 = p0.apply_perturbation(pert2)
    assert [d1, d2] == ka["epointsset_energy_delta"] and p2.total_energy() == e0 + d2 == 6.0
    back = p2.unapply_perturbation(pert2)
    assert back.total_energy() == e0 and some[2] in back
    # membership / iteration / errors (test/test_interacting_points_set.py:46-140, energy_point_set.py:88-116)
    for p in some:
        assert p in p0 and p in p0.points and p in p0.energy_graph.ue_per_point and p in p0.energy_graph.pe_per_point
    with pytest.raises(KeyError):
        p0.energy_delta(api.Perturbation(type=api.DeathKernel, removal=api.Point(3, 3)))
    with pytest.raises(ValueError):
        p0.papangelou(some[0])
    assert p0.papangelou(some[2], remove_u_from_point_set=True, return_energy_delta=True) == -d1
    with pytest.raises(AssertionError):
        api.EPointsSet([], (10, 10), [api.ConstantUnitEnergy(name="a", value=1.0), api.ConstantUnitEnergy(name="a", value=2.0)], [])


def test_python_plugin_terms_are_rejected_loudly():
    api = _api()

    class MyUnit(api.UnitEnergyConstructor):
        def compute(self, u):
            return 1.0

    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        api.EPointsSet([], (10, 10), [MyUnit(name="mine")], [])


# ------------------------------------------------------------------------------------------------ real terms vs golden
def _image(api, det, marks, name="img"):
    return api.ImageWMaps(name=name, shape=det.shape, image=None, detection_map=det, param_dist_maps=list(marks),
                          mappings=api.default_mappings(), param_names=["size", "ratio", "angle"])


def _setup(api, cfg):
    if cfg == "legacy":
        c = gu.CALIB_HRCM
        s = api.LegacyEnergySetup(calibration_params={}, energy_calibration=api.LegacyEnergiesCalibration(
            c["detection_threshold"], list(c["coefs"]), list(c["intercepts"]), c["min_area"], c["max_area"]))
        h = gu.HRC
        comb = api.HierarchicalEnergyCombinator(np.array(h["weights_data"]), np.array(h["weights_prior"]), np.array(h["data_prior_weights"]),
                                                h["detection_threshold"], h["bias"])
    else:
        c = gu.CALIB_LOG
        s = api.NoCalibrationEnergySetup(ratio_prior=True)
        s.energy_calibration = api.NoCalibEnergiesCalibration(c["min_area"], c["max_area"])
        comb = api.LogisticEnergyCombinator(weights=gu.LOG_WEIGHTS, bias=gu.LOG_BIAS, energy_names=s.energy_names)
    return s, comb


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_epointsset_real_terms_match_reference(cfg):
    api = _api()
    g = gu.load(f"energies_{cfg}.npz")
    _, det, marks = gu.scene_inputs(g)
    setup, comb = _setup(api, cfg)
    unit, pair = setup.make_energies(_image(api, det, marks))
    rects = [api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in g["config"]]
    eps = api.EPointsSet(rects, det.shape, unit, pair)
    vec = eps.energy_graph.compute_subset(rects, return_vector=True)
    names = [str(s) for s in g["names"]]
    assert set(vec) == set(names)
    got = np.array([vec[k] for k in names]).T
    np.testing.assert_allclose(got, g["vectors"], rtol=1e-5, atol=1e-5)
    assert abs(eps.energy_graph.compute_subset(rects, energy_combinator=comb) - float(g["comb_total"])) < 1e-3
    assert abs(eps.total_energy() - float(g["raw_total"])) < 2e-3
    # the combinator's own compute() (device) on the reference's vectors
    assert abs(comb.compute({k: list(g["vectors"][:, i]) for i, k in enumerate(names)}) - float(g["comb_total"])) < 1e-9 * max(1, abs(float(g["comb_total"])))
    for k, (ri, add) in enumerate(zip(g["pert_removal"], g["pert_addition"])):
        rem = rects[ri] if ri >= 0 else None
        a = None if np.isnan(add[0]) else api.Rectangle(int(add[0]), int(add[1]), add[2], add[3], add[4])
        p = api.Perturbation(type=api.BirthKernel, removal=rem, addition=a)
        assert abs(eps.energy_delta(p, energy_combinator=comb) - g["delta_comb"][k]) < 2e-5 + 1e-5 * abs(g["delta_comb"][k])
        assert abs(eps.energy_delta(p) - g["delta_raw"][k]) < 4e-5 + 1e-5 * abs(g["delta_raw"][k])
    assert len(eps) == len(rects)
    # batched Papangelou scores == one by one (mpp_model.py:296-304)
    objs, scores = eps.papangelou_all(energy_combinator=comb)
    one = [eps.papangelou(u, energy_combinator=comb, remove_u_from_point_set=True) for u in objs[:5]]
    np.testing.assert_allclose(scores[:5], one, rtol=1e-6)
    # naive initial configuration (sample_rjmcmc.py:23-35)
    naive = api.naive_detection(_image(api, det, marks), float(g["naive_threshold"]), energy_setup=setup)
    got_n = np.array(sorted((r.x, r.y, r.size, r.ratio, r.angle) for r in naive))
    want_n = np.array(sorted(tuple(r) for r in g["naive"]))
    np.testing.assert_allclose(got_n, want_n, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_kernel_probabilities_match_reference(cfg):
    """Kernel.forward_probability / backward_probability through the API on the reference's recorded proposals."""
    api = _api()
    g = gu.load(f"energies_{cfg}.npz")
    _, det, marks = gu.scene_inputs(g)
    setup, comb = _setup(api, cfg)
    img = _image(api, det, marks)
    unit, pair = setup.make_energies(img)
    rects = [api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in g["config"]]
    eps = api.EPointsSet(rects, det.shape, unit, pair)
    kernels, p = api.make_kernels(img, intensity=float(g["intensity"]), rng=np.random.default_rng(0))
    np.testing.assert_allclose(p, g["p_kernels"], rtol=0, atol=1e-15)
    for row in g["kernel_probs"]:
        kid, ri = int(row[0]), int(row[1])
        add = None if np.isnan(row[2]) else api.Rectangle(int(row[2]), int(row[3]), row[4], row[5], row[6])
        data = {"delta": np.array([row[7], row[8]]) if kid == 4 else float(row[7]), "param_id": int(row[9]), "new_param_class_value": int(row[10])}
        u = api.Perturbation(type=kernels[kid].__class__, removal=rects[ri] if ri >= 0 else None, addition=add, data=data)
        f, b = kernels[kid].forward_probability(eps.points, u), kernels[kid].backward_probability(eps.points, u)
        assert abs(f - row[11]) <= 3e-6 * abs(row[11]) + 1e-300, (kid, f, row[11])
        assert abs(b - row[12]) <= 3e-6 * abs(row[12]) + 1e-300, (kid, b, row[12])


def test_rjmcmc_step_and_run_and_sample_rjmcmc():
    api = _api()
    g = gu.load("energies_legacy.npz")
    truth, det, marks = gu.scene_inputs(g)
    setup, comb = _setup(api, "legacy")
    img = _image(api, det, marks)
    unit, pair = setup.make_energies(img)
    rects = [api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in g["config"]]
    # (1) step-by-step through the Kernel objects
    eps = api.EPointsSet(rects, det.shape, unit, pair)
    kernels, p = api.make_kernels(img, intensity=max(1, len(rects)), rng=np.random.default_rng(0))
    chain = api.RJMCMC(t0=0.05, kernels=kernels, p_kernels=p, initial_state=eps, stopping_condition=api.StopOnMaxIter(59),
                       rng=np.random.default_rng(3), energy_combinator=comb, alpha_t=1.0)
    n_acc = 0
    for _ in range(60):
        s = chain.step()
        n_acc += bool(s.move_accepted)
        assert s.n_points == len(eps)
    with pytest.raises(StopIteration):
        chain.step()
    assert 0 < n_acc < 60
    # every stored object is consistent with the device (energies still evaluate)
    vec = eps.energy_graph.compute_subset(list(eps), return_vector=True)
    assert len(vec["PositionEnergy"]) == len(eps)
    # (2) whole chain on the device
    eps2 = api.EPointsSet(rects, det.shape, unit, pair)
    chain2 = api.RJMCMC(t0=0.05, kernels=kernels, p_kernels=p, initial_state=eps2, stopping_condition=api.StopOnMaxIter(2000),
                        rng=np.random.default_rng(4), energy_combinator=comb, alpha_t=0.999,
                        sampling_rule=lambda step: step >= 1000 and step % 500 == 0)
    states, summaries = chain2.run()
    assert len(summaries) == 2002 and summaries[-1].iter == 2000  # max_iter + 1 steps (stopping.py:42)
    assert len(states) == 1 + 3  # snapshots at 1000, 1500, 2000
    assert abs(len(states[-1]) - len(rects)) < 0.5 * len(rects)
    # (3) the drop-in entry point, both samplers
    for sampler in ("parallel", "sequential"):
        out = api.sample_rjmcmc(image_data=img, rng=np.random.default_rng(5), num_samples=1, energy_combinator=comb, init_config="naive",
                                init_temperature=0.1, alpha_t=0.9995, burn_in=6000, energy_setup=setup, samples_interval=64,
                                target_temperature=0.0, sampler=sampler)
        assert len(out) == 1
        final = list(out[0])
        assert all(isinstance(r, api.Rectangle) for r in final)
        # annealed from T=0.1 to ~0.005 (a T0=1 schedule needs the full 30k-step budget): the configuration is close to the objects the maps were synthesised from
        assert 0.5 * len(truth) <= len(final) <= 1.6 * len(truth), (sampler, len(final), len(truth))


def test_mpp_model_infer_and_cli(tmp_path):
    """MPPModel.infer (whole image and reference-style tiling + merge) and the command line on a synthetic scene."""
    import json
    import os
    api = _api()
    from mpp_cnn_rs_object_detection_b200 import synth
    from mpp_cnn_rs_object_detection_b200.main import main as cli
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    cfg = json.load(open(os.path.join(gold, "model_mpp_hrcM", "config.json")))
    cfg["inference"]["rjmcmc_params"].update(init_temperature=0.1, alpha_t=0.9996, burn_in=12000, samples_interval=64)
    model = api.MPPModel(cfg, model_dir=os.path.join(gold, "model_mpp_hrcM"))
    truth, det, marks = synth.make_scene(31, (300, 420), 110)
    img = api.ImageWMaps("0007", det.shape, None, det, marks, api.default_mappings(), ["size", "ratio", "angle"])
    found = {}
    for tile in (False, True):
        res = model.infer([img], results_dir=str(tmp_path / f"tile{int(tile)}"), tile=tile)[0]
        centers = res["detection_center"]
        assert len(res["detection_score"]) == len(centers) == len(res["detection"]) and res["detection"].shape[1:] == (4, 2)
        d = np.abs(centers[:, None, :] - truth[None, :, :2]).max(-1)
        recall = (d.min(0) <= 3).mean()
        precision = (d.min(1) <= 3).mean()
        assert recall > 0.85 and precision > 0.85, (tile, recall, precision, len(centers), len(truth))
        assert os.path.exists(tmp_path / f"tile{int(tile)}" / "0007_results.pkl")
        found[tile] = len(centers)
        if tile:  # merge_patches: no two detections closer than 3 px survive
            dd = np.abs(centers[:, None, :] - centers[None, :, :]).astype(float)
            dist = np.hypot(dd[..., 0], dd[..., 1]) + np.eye(len(centers)) * 1e9
            assert dist.min() > 3
    assert abs(found[True] - found[False]) <= 0.2 * len(truth)
    out = cli(["-p", "infer", "-m", "mpp", "-c", os.path.join(gold, "model_mpp_hrcM", "config.json"), "--synthetic", "128x160"])
    assert len(out) == 1 and "detection_score" in out[0]
    with pytest.raises(ValueError):
        cli(["-p", "train", "-m", "mpp", "-c", os.path.join(gold, "model_mpp_hrcM", "config.json")])


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_split_merge_kernels_match_reference(cfg):
    """Optional two-object moves: Delta-energy of list perturbations and forward / backward probabilities
    against the reference's own values; a short chain with all ten kernels runs through RJMCMC.step."""
    api = _api()
    g = gu.load(f"split_merge_{cfg}.npz")
    _, det, marks = gu.scene_inputs(g)
    setup, comb = _setup(api, cfg)
    img = _image(api, det, marks)
    unit, pair = setup.make_energies(img)
    rects = [api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in g["config"]]
    eps = api.EPointsSet(rects, det.shape, unit, pair)
    kernels, p = api.make_kernels(img, intensity=float(g["intensity"]), rng=np.random.default_rng(0), use_split_merge=True)
    np.testing.assert_allclose(p, g["p_kernels"], rtol=1e-15)
    for row in g["rows"]:
        kid, r0, r1 = int(row[0]), int(row[1]), int(row[2])
        rem = [rects[k] for k in (r0, r1) if k >= 0]
        add = [api.Rectangle(int(row[3 + 5 * j]), int(row[4 + 5 * j]), row[5 + 5 * j], row[6 + 5 * j], row[7 + 5 * j]) for j in range(2)
               if not np.isnan(row[3 + 5 * j])]
        if kid == 8:
            u = api.Perturbation(api.SplitKernel, removal=rem[0] if rem else None, addition=add or None,
                                 data={"pos_delta": row[13:15], "shape_delta": row[15:18]})
        else:
            u = api.Perturbation(api.MergeKernel, removal=rem or None, addition=add[0] if add else None, data={"n_neighbors": int(row[18])})
        f, b = kernels[kid].forward_probability(eps.points, u), kernels[kid].backward_probability(eps.points, u)
        assert abs(f - row[21]) <= 1e-9 * abs(row[21]) and abs(b - row[22]) <= 1e-9 * abs(row[22]), (kid, f, row[21], b, row[22])
        if rem or add:
            assert abs(eps.energy_delta(u, energy_combinator=comb) - row[20]) < 4e-5 + 1e-5 * abs(row[20])
            assert abs(eps.energy_delta(u) - row[19]) < 8e-5 + 1e-5 * abs(row[19])
            assert len(eps) == len(rects)  # mutate-and-revert left the state untouched
    vec = eps.energy_graph.compute_subset(rects, return_vector=True)
    names = [str(s) for s in gu.load(f"energies_{cfg}.npz")["names"]]
    np.testing.assert_allclose(np.array([vec[k] for k in names]).T, gu.load(f"energies_{cfg}.npz")["vectors"], rtol=1e-5, atol=1e-5)
    chain = api.RJMCMC(t0=0.05, kernels=kernels, p_kernels=p, initial_state=eps, stopping_condition=api.StopOnMaxIter(150),
                       rng=np.random.default_rng(1), energy_combinator=comb, alpha_t=1.0)
    states, summaries = chain.run()
    kinds = {s.kernel for s in summaries[1:]}
    assert api.SplitKernel in kinds and api.MergeKernel in kinds and len(summaries) == 152
    assert summaries[-1].n_points == len(states[-1])


def test_plugin_combinators_mlp_and_manual():
    """Combinators that are not fused into the kernels (a torch MLP) go through the before / after recipe on device-computed
    vectors; the manual hierarchical combinator is fused.  Both must satisfy E(after) - E(before) == energy_delta."""
    import torch
    api = _api()
    g = gu.load("energies_nocalib.npz")
    _, det, marks = gu.scene_inputs(g)
    setup, _ = _setup(api, "nocalib")
    img = _image(api, det, marks)
    unit, pair = setup.make_energies(img)
    rects = [api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in g["config"]]
    torch.manual_seed(0)
    mlp = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.ReLU(), torch.nn.Linear(8, 1), torch.nn.Sigmoid())
    combs = [api.MLPEnergyCombinator(model=mlp, energy_names=setup.energy_names),
             api.ManualHierarchicalEnergyCombinator(weights_dict={k: 0.3 + 0.1 * i for i, k in enumerate(setup.energy_names)},
                                                    indicator_energy="PositionEnergy", detection_threshold=0.0)]
    for comb in combs:
        eps = api.EPointsSet(rects, det.shape, unit, pair)
        e0 = eps.energy_graph.compute_subset(list(eps), energy_combinator=comb)
        for k, (ri, add) in enumerate(list(zip(g["pert_removal"], g["pert_addition"]))[:25]):
            rem = rects[ri] if ri >= 0 else None
            a = None if np.isnan(add[0]) else api.Rectangle(int(add[0]), int(add[1]), add[2], add[3], add[4])
            p = api.Perturbation(type=api.BirthKernel, removal=rem, addition=a)
            d = eps.energy_delta(p, energy_combinator=comb)
            after = eps.apply_perturbation(p)
            e1 = after.energy_graph.compute_subset(list(after), energy_combinator=comb)
            assert abs((e1 - e0) - d) < 2e-4 + 1e-5 * abs(e0), (type(comb).__name__, k, e1 - e0, d)


def test_training_helpers_energy_vectors_and_kernel_walks():
    """energy_utils.compute_many_energy_vectors (energy_utils.py:69-82) on the reference's perturbed configurations against the
    reference's own vectors (tests/golden/training_helpers.npz), and the kernel walks of perturbation_sampler.py:127-169."""
    api = _api()
    g = gu.load("training_helpers.npz")
    _, det, marks = gu.scene_inputs(g)
    setup, comb = _setup(api, "legacy")
    image = _image(api, det, marks)
    image.gt_config = [api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in g["config"]]
    ue, pe = setup.make_energies(image)
    cfgs = []
    for name in ("light", "overlap", "strong"):
        flat, lens = g[f"pert_{name}"], g[f"pert_{name}_len"]
        off = 0
        for n in lens:
            cfgs.append([api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in flat[off:off + n]])
            off += n
    names = [str(s) for s in g["energy_names"]]
    assert names == setup.energy_names
    vec = api.compute_many_energy_vectors(cfgs, image, ue, pe, names)
    assert vec.shape == g["vectors"].shape

    def canon(m):  # the reference lists a configuration's objects in set-iteration order: compare as multisets of rows
        return m[np.lexsort(np.round(m, 4).T[::-1])]
    off = 0
    for c in cfgs:
        np.testing.assert_allclose(canon(vec[off:off + len(c)]), canon(g["vectors"][off:off + len(c)]), rtol=1e-5, atol=1e-5)
        off += len(c)
    one, got_names = api.compute_energy_vector(cfgs[0], ue, pe, det.shape, names, return_names=True)
    np.testing.assert_allclose(canon(one), canon(g["vectors"][:len(cfgs[0])]), rtol=1e-5, atol=1e-5)
    assert set(got_names) == set(names) and api.names_from_energies(ue + pe) == [e.name for e in ue + pe]
    assert api.compute_many_energy_vectors([], image, ue, pe, names).shape == (0, len(names))
    # kernel walks: iter_per_point * n moves from the ground truth; the aggregated perturbation turns the start into the end
    rng = np.random.default_rng(4)
    walks = api.sample_multiple_kernel_perturbations(image, n_samples=2, rng=rng, energy_setup=setup, iter_per_point=0.5)
    assert len(walks) == 2 and all(isinstance(w, api.EPointsSet) for w in walks)
    perts = api.sample_multiple_kernel_perturbations(image, n_samples=1, rng=rng, energy_setup=setup, iter_per_point=0.5,
                                                     return_perturbations=True, aggregate_pert=True)
    agg = perts[0]
    start = api.EPointsSet(image.gt_config, det.shape, ue, pe)
    kernels, p_kernels = api.make_kernels(image, intensity=1.0, rng=rng)
    end, seq = api.sample_kernel_perturbations(kernels, p_kernels, 0.5, start, rng)
    assert len(seq) == int(0.5 * len(start)) and len(start) == len(image.gt_config)
    net = api.aggregate_perturbations(seq)
    want = (set(start) - set(net.removal)) | set(net.addition)
    assert set(end) == want and isinstance(agg.removal, list) and isinstance(agg.addition, list)


def test_sample_point_2d_export():
    """utils/sampler2d.py:5-48 through the device inverse-CDF sampler: signature, shapes, distribution, without replacement."""
    import mpp_cnn_rs_object_detection_b200.api as api
    rng = np.random.default_rng(3)
    shape = (37, 53)
    dens = rng.random(shape).astype(np.float32) ** 3
    dens[5:9, :] = 0.0
    pts = api.sample_point_2d(shape, size=1, density=dens, rng=rng)
    assert pts.shape == (1, 2) and dens[pts[0, 0], pts[0, 1]] > 0
    uni = api.sample_point_2d(shape, size=5, rng=rng)
    assert uni.shape == (5, 2) and np.all((uni >= 0) & (uni < np.array(shape)))
    # many single draws: chi-square against the density, coarsened to row bands
    n = 20000
    from mpp_cnn_rs_object_detection_b200.api.sampler2d import _device_draws
    d = _device_draws(dens, shape, n, seed=11)
    assert np.all(dens[d[:, 0], d[:, 1]] > 0)
    want = dens.reshape(-1) / dens.sum()
    got = np.bincount(d[:, 0] * shape[1] + d[:, 1], minlength=shape[0] * shape[1]) / n
    band = lambda v: v.reshape(shape).sum(1)  # noqa: E731
    e, o = band(want) * n, band(got) * n
    keep = e > 5
    chi2 = float(np.sum((o[keep] - e[keep]) ** 2 / e[keep]))
    assert chi2 < 2.5 * keep.sum(), (chi2, keep.sum())
    # without replacement: all pixels of a tiny support come out exactly once
    small = np.zeros(shape, dtype=np.float32)
    small[3, 4], small[10, 11], small[20, 1] = 0.7, 0.2, 0.1
    allp = api.sample_point_2d(shape, size=3, density=small, rng=rng)
    assert sorted(map(tuple, allp.tolist())) == [(3, 4), (10, 11), (20, 1)]
    with pytest.raises(ValueError):
        api.sample_point_2d(shape, size=4, density=small, rng=rng)
    # mask semantics of the reference: density[mask] = 0
    mask = np.zeros(shape, dtype=bool)
    mask[3, 4] = True
    two = api.sample_point_2d(shape, size=2, density=small.copy(), rng=rng, mask=mask)
    assert sorted(map(tuple, two.tolist())) == [(10, 11), (20, 1)]


def test_rjmcmc_timer_interface():
    """RJMCMC.get_timings() / get_state_log() with the reference's stage names (rjmcmc.py:18-48,183-187)."""
    import mpp_cnn_rs_object_detection_b200.api as api
    from mpp_cnn_rs_object_detection_b200 import synth
    objs, det, marks = synth.make_scene(4, (64, 96), 10)
    img = api.ImageWMaps("t", (64, 96), None, det, marks, api.default_mappings(), ["size", "ratio", "angle"])
    setup = api.LegacyEnergySetup(calibration_params={}, energy_calibration=api.LegacyEnergiesCalibration(
        gu.CALIB_HRCM["detection_threshold"], list(gu.CALIB_HRCM["coefs"]), list(gu.CALIB_HRCM["intercepts"]), gu.CALIB_HRCM["min_area"], gu.CALIB_HRCM["max_area"]))
    unit, pair = setup.make_energies(img)
    pts = api.EPointsSet([api.Rectangle(int(o[0]), int(o[1]), float(o[2]), float(o[3]), float(o[4])) for o in objs], img.shape, unit, pair)
    rng = np.random.default_rng(0)
    kernels, p = api.make_kernels(img, intensity=len(objs), rng=rng)
    chain = api.RJMCMC(t0=0.1, kernels=kernels, p_kernels=p, initial_state=pts, stopping_condition=api.StopOnMaxIter(40), rng=rng, alpha_t=0.99, verbose=1)
    chain.run()
    t = chain.get_timings()
    assert isinstance(t, api.RJMCMCTimer)
    for key in ("sample_kernel", "sample_perturbation", "compute_energy", "compute_alpha", "apply_perturbation", "log", "total", "n_points"):
        assert key in t.timings and len(t.timings[key]) == 41, (key, len(t.timings.get(key, [])))
    assert all(v >= 0 for v in t.timings["total"]) and len(chain.get_state_log()) >= 1
    t.show_results()


def test_split_merge_device_draws():
    """Device draws of the split / merge kernels (mpp_sample_split_merge): structure of the perturbation, support of the
    deltas, uniform choice among the neighbours within the radius, and fwd / bwd consistency of a split with the merge that undoes it."""
    api = _api()
    g = gu.load("split_merge_legacy.npz")
    _, det, marks = gu.scene_inputs(g)
    setup, comb = _setup(api, "legacy")
    img = _image(api, det, marks)
    unit, pair = setup.make_energies(img)
    rects = [api.Rectangle(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in g["config"]]
    eps = api.EPointsSet(rects, det.shape, unit, pair)
    kernels, p = api.make_kernels(img, intensity=float(len(rects)), rng=np.random.default_rng(0), use_split_merge=True)
    split, merge = kernels[8], kernels[9]
    rng = np.random.default_rng(5)
    sig = [0.1 * 32.0, 0.1 * 1.0, 0.1 * np.pi]
    deltas = []
    for _ in range(300):
        u = split.sample_perturbation(eps.points, rng)
        q, (a0, a1) = u.removal, u.addition
        assert q in rects and len(u.addition) == 2
        d, sd = u.data["pos_delta"], u.data["shape_delta"]
        assert d[0] >= 0 and d[1] >= 0 and np.hypot(*d) <= 16 + 1e-9          # quarter disc (the reference draws uniform(0, r))
        assert a0.x == int(np.clip(q.x - d[0], 0, det.shape[0] - 1)) and a1.y == int(np.clip(q.y + d[1], 0, det.shape[1] - 1))
        assert abs(a1.size - np.clip(q.size + sd[0], 0, 32)) < 1e-6 and abs(a0.ratio - np.clip(q.ratio - sd[1], 0, 1)) < 1e-6
        assert 0 <= a0.angle < np.pi + 1e-9 and abs((a1.angle - (q.angle + sd[2])) % np.pi) < 1e-6 or abs((a1.angle - (q.angle + sd[2])) % np.pi - np.pi) < 1e-6
        f, b = split.forward_probability(eps.points, u), split.backward_probability(eps.points, u)
        assert f > 0 and b > 0
        deltas.append(np.concatenate([d, sd / np.array(sig)]))
    deltas = np.array(deltas)
    assert abs(deltas[:, 2:].mean()) < 0.15 and 0.8 < deltas[:, 2:].std() < 1.2   # standard normal shape deltas
    picks = {}
    for _ in range(400):
        u = merge.sample_perturbation(eps.points, rng)
        nn = u.data["n_neighbors"]
        if u.removal is None:
            assert nn == 0 and u.addition is None
            continue
        p0, p1 = u.removal
        assert p0 is not p1 and np.hypot(p0.x - p1.x, p0.y - p1.y) <= 16
        assert nn == len(eps.points.get_neighbors(p0, radius=16))
        assert u.addition.x == int(np.clip((p0.x + p1.x) / 2, 0, det.shape[0] - 1)) and abs(u.addition.size - (p0.size + p1.size) / 2) < 1e-6
        picks.setdefault(id(p0), {}).setdefault(id(p1), 0)
        picks[id(p0)][id(p1)] += 1
        assert merge.forward_probability(eps.points, u) == pytest.approx(p[9] / len(rects) / nn, rel=1e-12)
    assert any(len(v) > 1 for v in picks.values())   # objects with several neighbours see more than one of them drawn
