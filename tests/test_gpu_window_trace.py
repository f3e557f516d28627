"""-m gpu: the production sampler pinned to a CPU oracle proposal by proposal.

mpp_run_windows (debug instantiation of the SAME kernels, 8 speculating warps, dataflow schedule -- and 1 warp under the
colour-barrier schedule) writes one mpp_window_trace record per proposal of the chain.  oracle/window_oracle.py replays the
chain on the CPU: it regenerates the Philox words, re-derives the kernel, the picked object and the drawn perturbation,
recomputes the Delta-energy with the reference's algorithm (OracleState.delta, energy_graph.py:139-225) and the Green ratio
of the window-restricted kernels from first principles, and checks the decision.  Tolerances (device = float32 + fast
intrinsics, oracle = float64): |Delta E| 1e-5 relative + 2e-5 absolute, log(bwd/fwd) 5e-4 absolute; a decision may only
differ when log u lies inside that noise band of log alpha (counted as `borderline`, bounded below)."""
import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _oracle_scene(cfg, det, marks):
    from oracle import mpp_oracle as orc, window_oracle as wo
    if cfg == "legacy":
        c = gu.CALIB_HRCM
        scene = wo.WindowScene(det, marks, setup="legacy", detection_threshold=c["detection_threshold"], remap_coefs=c["coefs"],
                               remap_intercepts=c["intercepts"], min_area=c["min_area"], max_area=c["max_area"])
        comb = orc.OracleHierarchical(**gu.HRC)
        gain = 0.5 * gu.HRC["weights_prior"][0]
    else:
        c = gu.CALIB_LOG
        scene = wo.WindowScene(det, marks, setup="nocalib", detection_threshold=0.0, min_area=c["min_area"], max_area=c["max_area"],
                               ratio_prior=True)
        comb = orc.OracleLogistic(weights=gu.LOG_WEIGHTS, bias=gu.LOG_BIAS, energy_names=orc.NOCALIB_NAMES)
        gain = 0.5 * float(gu.LOG_WEIGHTS[4])
    return scene, comb, gain


def _run(cfg, shape, n_rect, n_sweeps, per_visit, n_warps, schedule, t0, alpha_t, seed, scene_seed=5, t_target=0.0):
    from mpp_cnn_rs_object_detection_b200 import synth
    from mpp_cnn_rs_object_detection_b200.engine import classes_of_marks
    from oracle import window_oracle as wo
    from tests.gpu_util import make_engine
    objs, det, marks = synth.make_scene(scene_seed, shape, n_rect)
    eng = make_engine(cfg, det, marks, "fp32", intensity=max(1, len(objs)))
    uid = np.arange(len(objs))
    eng.add_objects(objs[:, :2], objs[:, 2:5], uid=uid)
    cls = classes_of_marks(objs[:, 2:5])
    cnt, maxdiff, trace = eng.trace_windows(n_sweeps, per_visit, n_warps=n_warps, t0=t0, alpha_t=alpha_t, t_target=t_target, seed=seed,
                                            schedule=schedule)
    scene, comb, gain = _oracle_scene(cfg, det, marks)
    oracle = wo.WindowOracle(scene, comb, [tuple(o) + (int(u), tuple(c)) for o, u, c in zip(objs, uid, cls)], intensity=max(1, len(objs)),
                             seed=seed, per_visit=per_visit, t0=t0, alpha_t=alpha_t, t_target=t_target, overlap_gain=gain)
    st = oracle.replay(trace)
    # the device counters and the final configuration agree with the replayed chain
    assert st["proposals"] == cnt[0] and st["evaluated"] == cnt[4] and st["accepted"] == cnt[1], (st, cnt)
    ks = eng.window_stats()  # the per-kernel tallies bench.py reports (mpp_window_stats)
    for key in ("evaluated_empty", "evaluated_occupied", "accepted_empty", "accepted_occupied"):
        assert ks[key] == st[key], (key, ks[key], st[key])
    assert ks["identity_accepted"] == st["identity"]
    _, xy, mk, u = eng.read_objects()
    order = np.argsort(u)
    want = oracle.configuration()
    assert len(want) == len(u)
    np.testing.assert_array_equal(u[order], want[:, 5].astype(np.uint32))
    np.testing.assert_array_equal(xy[order], want[:, :2].astype(np.int32))
    np.testing.assert_allclose(mk[order], want[:, 2:5], rtol=0, atol=1e-6)
    assert maxdiff < 2e-5
    eng.close()
    return st


@pytest.mark.parametrize("cfg,n_warps,schedule", [("legacy", 8, "dataflow"), ("nocalib", 8, "dataflow"), ("legacy", 1, "colours")])
def test_trace_replays_in_the_oracle(cfg, n_warps, schedule):
    """>= 1e5 proposals over the parametrisations; sparse scene (most windows empty: births-only mixture) at a temperature
    where births, deaths and moves are all accepted."""
    st = _run(cfg, (160, 192), 45, n_sweeps=30, per_visit=32, n_warps=n_warps, schedule=schedule, t0=0.25, alpha_t=1.0, seed=3)
    print(f"\n{cfg} nw={n_warps} {schedule}: {st}")
    assert st["proposals"] >= 38000
    assert st["empty_window"] > 0.2 * st["proposals"] and st["identity"] > 0 and st["left_window"] > 0
    assert all(a > 0 for a in st["accepted_per_kernel"]), st["accepted_per_kernel"]
    assert st["borderline"] <= 2e-3 * st["evaluated"] + 2


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_trace_dense_scene_with_annealing(cfg):
    """Dense scene (overlapping neighbours, second-best partners, full speculation) under the annealing schedule of the
    reference (temperature decaying inside a visit), down to the cold regime of the benchmark."""
    st = _run(cfg, (128, 160), 90, n_sweeps=12, per_visit=96, n_warps=8, schedule="dataflow", t0=0.2, alpha_t=0.8, seed=11,
              scene_seed=8, t_target=0.02)
    print(f"\n{cfg} dense annealed: {st}")
    assert st["proposals"] >= 30000 and st["accepted"] > 0
    assert st["borderline"] <= 2e-3 * st["evaluated"] + 2
