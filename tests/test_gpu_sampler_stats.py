"""-m gpu: the three samplers target the same distribution.  At a fixed temperature the number of objects of
  (a) the CPU oracle's sequential chain (== the reference's RJMCMC.run),
  (b) the device sequential chain (mpp_run_chain, reference kernels, Philox),
  (c) the device parallel colour sweeps (mpp_run_sweeps, cell-local kernels)
must have the same mean within Monte-Carlo error (batch means)."""
import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _batch_se(x, nb=20):
    x = np.asarray(x, dtype=np.float64)
    m = len(x) // nb
    b = x[:m * nb].reshape(nb, m).mean(1)
    return x.mean(), b.std(ddof=1) / np.sqrt(nb)


@pytest.mark.parametrize("cfg", ["legacy"])
def test_stationary_object_count_agrees(cfg):
    from mpp_cnn_rs_object_detection_b200 import synth
    from oracle import mpp_oracle as orc
    from tests.gpu_util import make_engine
    from tests.test_oracle_golden import make_oracle_scene

    temp = 0.3
    objs, det, marks = synth.make_scene(11, (64, 96), 8)
    scene, comb = make_oracle_scene(cfg, det, marks)
    n0 = len(objs)
    # (a) CPU oracle
    smp = orc.OracleSampler(scene, comb, [orc.ORect(*r) for r in objs], np.random.default_rng(0), temp, 1.0)
    smp.run(3000)
    na = []
    for _ in range(1500):
        smp.run(10)
        na.append(len(smp.state))
    ma, sa = _batch_se(na)
    # (b) device sequential chain
    eng = make_engine(cfg, det, marks, "fp32", intensity=max(1, n0))
    eng.add_objects(objs[:, :2], objs[:, 2:5])
    eng.run_chain(20000, t0=temp, seed=1)
    _, trace = eng.run_chain(400000, t0=temp, seed=1, step_offset=20000, trace=True)
    mb, sb = _batch_se(trace["n_after"][::10])
    acc_b = trace["accepted"].mean()
    # (c) device parallel sweeps
    eng2 = make_engine(cfg, det, marks, "fp32", intensity=max(1, n0))
    eng2.add_objects(objs[:, :2], objs[:, 2:5])
    eng2.run_sweeps(500, proposals_per_visit=4, stride=3, t0=temp, seed=2)
    nc = []
    tot = acc = 0
    for s in range(4000):
        c = eng2.run_sweeps(1, proposals_per_visit=4, stride=3, t0=temp, seed=2, sweep_offset=500 + s)
        tot += c[0]; acc += c[1]
        nc.append(len(eng2))
    mc, sc = _batch_se(nc)
    print(f"\nobject count at T={temp}: oracle {ma:.3f}+-{sa:.3f} | device chain {mb:.3f}+-{sb:.3f} (acc {acc_b:.3f}) | "
          f"sweeps {mc:.3f}+-{sc:.3f} (acc {acc / tot:.3f}); start {n0}")
    assert abs(ma - mb) < 5 * np.hypot(sa, sb) + 0.05, (ma, mb)
    assert abs(mb - mc) < 5 * np.hypot(sb, sc) + 0.05, (mb, mc)
