"""-m gpu: the production parallel sampler (mpp_run_windows).  Checks: the fast Delta-energy (top-2 reductions in shared
memory) against its brute-force recomputation inside the kernel; independence of the chain from the speculation depth;
integrity of the records it writes; and agreement of its stationary distribution with the sequential device chain."""
import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _scene(cfg, seed=5, shape=(192, 224), n_rect=60):
    from mpp_cnn_rs_object_detection_b200 import synth
    from tests.gpu_util import make_engine
    objs, det, marks = synth.make_scene(seed, shape, n_rect)
    eng = make_engine(cfg, det, marks, "fp32", intensity=max(1, len(objs)))
    eng.add_objects(objs[:, :2], objs[:, 2:5])
    return objs, det, marks, eng


@pytest.mark.parametrize("n_warps", [4, 0])
@pytest.mark.parametrize("cfg,n_rect,temp", [("legacy", 60, 0.03), ("nocalib", 60, 0.03), ("legacy", 170, 0.02), ("nocalib", 170, 0.01)])
def test_fast_delta_equals_brute_force(cfg, n_rect, temp, n_warps):
    """n_rect=170 on 192x224 is ~4x the benchmark density: many partners within reach, overlaps and second-best partners.
    n_warps=0 is the lane-per-proposal mode."""
    objs, det, marks, eng = _scene(cfg, n_rect=n_rect)
    cnt, maxdiff = eng.run_windows(40, proposals_per_visit=12 if n_warps else 40, n_warps=n_warps, t0=temp, seed=3, debug=True,
                                   schedule="colours" if n_rect < 100 else "dataflow")
    assert cnt[0] > 0 and cnt[4] > 0 and cnt[1] > 0
    assert maxdiff < 2e-5, maxdiff
    assert len(eng) == len(objs) + cnt[2] - cnt[3]


def test_chain_is_independent_of_speculation_depth_and_schedule():
    """The chain depends neither on the number of speculating warps nor on the schedule (colour barriers vs dataflow)."""
    finals = []
    for nw, schedule in ((1, "colours"), (2, "colours"), (4, "colours"), (8, "colours"), (1, "dataflow"), (4, "dataflow"), (8, "dataflow")):
        objs, det, marks, eng = _scene("legacy")
        cnt = eng.run_windows(25, proposals_per_visit=10, n_warps=nw, t0=0.03, seed=9, schedule=schedule)
        h, xy, mk, uid = eng.read_objects()
        order = np.lexsort((xy[:, 1], xy[:, 0]))
        finals.append((cnt[:5], xy[order], mk[order]))
    for f in finals[1:]:
        assert f[0] == finals[0][0]
        np.testing.assert_array_equal(f[1], finals[0][1])
        np.testing.assert_array_equal(f[2], finals[0][2])


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_written_records_are_consistent(cfg):
    """After many accepted moves the cached unit energies / geometry of every stored record equal a fresh insertion."""
    from tests.gpu_util import make_engine
    objs, det, marks, eng = _scene(cfg)
    eng.run_windows(60, proposals_per_visit=16, n_warps=4, t0=0.05, seed=4)
    h, xy, mk, uid = eng.read_objects()
    assert len(set(uid.tolist())) == len(uid)
    vec, comb, raw, tot = eng.energy_vectors(h)
    fresh = make_engine(cfg, det, marks, "fp32")
    h2 = fresh.add_objects(xy, mk)
    vec2, comb2, raw2, tot2 = fresh.energy_vectors(h2)
    np.testing.assert_allclose(vec, vec2, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(comb, comb2, rtol=1e-6, atol=1e-6)
    assert np.all((xy[:, 0] >= 0) & (xy[:, 0] < det.shape[0]) & (xy[:, 1] >= 0) & (xy[:, 1] < det.shape[1]))


def _batch_se(x, nb=20):
    x = np.asarray(x, dtype=np.float64)
    m = len(x) // nb
    b = x[:m * nb].reshape(nb, m).mean(1)
    return x.mean(), b.std(ddof=1) / np.sqrt(nb)


def test_lane_mode_is_schedule_independent_and_consistent():
    """Lane-per-proposal mode: same chain under both schedules; the records it writes are consistent."""
    from tests.gpu_util import make_engine
    finals = []
    for schedule in ("colours", "dataflow"):
        objs, det, marks, eng = _scene("legacy")
        cnt = eng.run_windows(25, proposals_per_visit=40, n_warps=0, t0=0.03, seed=9, schedule=schedule)
        h, xy, mk, uid = eng.read_objects()
        order = np.lexsort((xy[:, 1], xy[:, 0]))
        finals.append((cnt[:5], xy[order], mk[order]))
        assert cnt[1] > 0 and len(eng) == len(objs) + cnt[2] - cnt[3]
    assert finals[0][0] == finals[1][0]
    np.testing.assert_array_equal(finals[0][1], finals[1][1])
    np.testing.assert_array_equal(finals[0][2], finals[1][2])
    vec, comb, raw, tot = eng.energy_vectors(h)
    fresh = make_engine("legacy", det, marks, "fp32")
    vec2, comb2, _, _ = fresh.energy_vectors(fresh.add_objects(xy, mk))
    np.testing.assert_allclose(vec, vec2, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n_warps", [2, 0])
def test_stationary_distribution_matches_sequential_chain(n_warps):
    from mpp_cnn_rs_object_detection_b200 import synth
    from tests.gpu_util import make_engine
    temp = 0.3
    objs, det, marks = synth.make_scene(11, (64, 96), 8)
    eng = make_engine("legacy", det, marks, "fp32", intensity=max(1, len(objs)))
    eng.add_objects(objs[:, :2], objs[:, 2:5])
    eng.run_chain(20000, t0=temp, seed=1)
    _, trace = eng.run_chain(600000, t0=temp, seed=1, step_offset=20000, trace=True)
    mb, sb = _batch_se(trace["n_after"][::10])
    e2 = make_engine("legacy", det, marks, "fp32", intensity=max(1, len(objs)))
    e2.add_objects(objs[:, :2], objs[:, 2:5])
    e2.run_windows(300, proposals_per_visit=4, n_warps=n_warps, t0=temp, seed=2)
    nc = []
    for s in range(6000):
        e2.run_windows(1, proposals_per_visit=4, n_warps=n_warps, t0=temp, seed=2, sweep_offset=300 + s, read_counters=False)
        nc.append(len(e2))
    mc, sc = _batch_se(nc)
    print(f"\nobject count at T={temp}: device chain {mb:.3f}+-{sb:.3f} | windows(n_warps={n_warps}) {mc:.3f}+-{sc:.3f}")
    assert abs(mb - mc) < 5 * np.hypot(sb, sc), (mb, mc)


@pytest.mark.parametrize("world", [2, 3])
def test_split_scene_follows_the_single_gpu_chain(world):
    """A scene split into row bands (one device context per band, boundary objects exchanged between colour-row phases)
    ends in exactly the configuration of mpp_run_windows(schedule='colours') on one context."""
    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg, synth
    from tests.gpu_util import make_engine
    objs, det, marks = synth.make_scene(21, (352, 224), 110)
    n_sweeps, pv, nw, temp, seed = 12, 8, 4, 0.03, 17
    uid = np.arange(len(objs))
    single = make_engine("legacy", det, marks, "fp32", intensity=max(1, len(objs)))
    single.add_objects(objs[:, :2], objs[:, 2:5], uid=uid)
    for s in range(n_sweeps):
        single.run_windows(1, pv, nw, t0=temp, seed=seed, sweep_offset=s, schedule="colours", read_counters=False)
        assert single.window_grid(seed, s) == mg.grid_offset(seed, s)
    _, xy1, mk1, uid1 = single.read_objects()
    scenes = []
    for r in range(world):
        eng = make_engine("legacy", det, marks, "fp32", intensity=max(1, len(objs)))
        sc = mg.SplitScene(eng, det.shape[0], r, world, capacity=1024)
        sel = sc.select_initial(objs[:, :2])
        eng.add_objects(objs[sel, :2], objs[sel, 2:5], uid=uid[sel])
        scenes.append(sc)
    for s in range(n_sweeps):
        mg.sweep_local(scenes, pv, nw, temp, seed, s)
    parts = [sc.owned_objects() for sc in scenes]
    xy2 = np.concatenate([p[0] for p in parts]); mk2 = np.concatenate([p[1] for p in parts]); uid2 = np.concatenate([p[2] for p in parts])
    assert len(xy2) == len(xy1) and len(set(uid2.tolist())) == len(uid2)
    o1, o2 = np.argsort(uid1), np.argsort(uid2)
    np.testing.assert_array_equal(uid1[o1], uid2[o2])
    np.testing.assert_array_equal(xy1[o1], xy2[o2])
    np.testing.assert_array_equal(mk1[o1], mk2[o2])
    assert abs(len(xy1) - len(objs)) < len(objs)  # the chain did something sensible


def test_soak_speculation_depths_annealing_and_brute_force_check():
    """tools/soak.py: five scenes (sparse to dense, T from 0.02 to 1, with and without annealing inside the visit), 7 to 128
    proposals per visit: the chain, the counters and the uids are the same for 1, 4 and 8 speculating warps, the counters are
    consistent with the object count, and every Delta-energy agrees with its brute-force recomputation."""
    import os
    import runpy
    runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "soak.py"), run_name="__main__")
