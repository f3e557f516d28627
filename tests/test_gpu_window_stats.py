"""-m gpu: the production window sampler (mpp_run_windows: 8 speculating warps, dataflow schedule) targets the same
distribution as the reference chain.  At a fixed temperature three samplers are run on the same small scene:

  (a) the CPU oracle's sequential chain (OracleSampler == RJMCMC.run, rjmcmc.py:83-181, the reference's global kernels),
      several independent chains in worker processes;
  (b) the device sequential chain with the reference's global kernels (mpp_run_chain);
  (c) the production window sampler.

Statistics: number of objects, total combined energy U(X) (what the accept rule sees), and the histogram of the objects'
size classes in four bins (a wrong mark-class density or proposal-mass factor biases this one while leaving the count
alone).  (c) must agree with (b) on count and energy and with (a) on all three within 5 standard errors -- batch means for
the device chains, across-chain spread for the oracle -- with no additive slack.  Both shipped configurations, two
temperatures."""
import multiprocessing as mp

import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu

SHAPE, N_RECT, SCENE_SEED = (64, 96), 8, 11


def _batch_se(x, nb=25):
    x = np.asarray(x, dtype=np.float64)
    m = len(x) // nb
    b = x[:m * nb].reshape(nb, m, *x.shape[1:]).mean(1)
    return x[:m * nb].mean(0), b.std(0, ddof=1) / np.sqrt(nb)


def _size_hist(size_classes):
    return np.bincount(np.asarray(size_classes, dtype=np.int64) // 8, minlength=4)[:4].astype(np.float64)


def _oracle_chain(args):
    cfg, temp, seed, burn, steps, every = args
    from mpp_cnn_rs_object_detection_b200 import synth
    from oracle import mpp_oracle as orc
    from tests.test_oracle_golden import make_oracle_scene
    objs, det, marks = synth.make_scene(SCENE_SEED, SHAPE, N_RECT)
    scene, comb = make_oracle_scene(cfg, det, marks)
    smp = orc.OracleSampler(scene, comb, [orc.ORect(*r) for r in objs], np.random.default_rng(seed), temp, 1.0)
    smp.run(burn)
    rows = []
    for _ in range(steps // every):
        smp.run(every)
        cur = smp.state.objects()
        e = smp.state.subset_energy(cur, comb) if cur else 0.0
        rows.append([len(cur), e, *_size_hist([orc.value_to_class(0, o.size) for o in cur])])
    return np.mean(np.array(rows, dtype=np.float64), axis=0)


@pytest.mark.parametrize("cfg,temp", [("legacy", 0.3), ("legacy", 0.12), ("nocalib", 0.3), ("nocalib", 0.12)])
def test_window_sampler_stationary_statistics(cfg, temp):
    from mpp_cnn_rs_object_detection_b200 import synth
    from mpp_cnn_rs_object_detection_b200.engine import classes_of_marks
    from tests.gpu_util import make_engine

    objs, det, marks = synth.make_scene(SCENE_SEED, SHAPE, N_RECT)
    n0 = len(objs)
    # (a) CPU oracle: independent chains, one process each
    n_chains = 32
    with mp.get_context("fork").Pool(min(n_chains, max(2, mp.cpu_count()))) as pool:
        res = np.array(pool.map(_oracle_chain, [(cfg, temp, 100 + k, 6000, 20000, 10) for k in range(n_chains)], chunksize=1))
    ma, sa = res.mean(0), res.std(0, ddof=1) / np.sqrt(n_chains)

    # (b) device sequential chain: count and energy (E0 + accepted Delta-energies) of every 10th step
    eng = make_engine(cfg, det, marks, "fp32", intensity=max(1, n0))
    h0 = eng.add_objects(objs[:, :2], objs[:, 2:5])
    e0 = eng.energy_vectors(h0)[3]
    _, tr0 = eng.run_chain(30000, t0=temp, seed=1, trace=True)
    e0 += float(np.sum(tr0["delta_e"] * tr0["accepted"]))
    _, tr = eng.run_chain(1500000, t0=temp, seed=1, step_offset=30000, trace=True)
    energy = e0 + np.cumsum(tr["delta_e"] * tr["accepted"])
    mb, sb = _batch_se(np.stack([tr["n_after"][::10].astype(np.float64), energy[::10]], axis=1))
    _, _, _, e_end = eng.energy_vectors(eng.read_objects()[0])
    assert abs(e_end - energy[-1]) < 1e-3 * max(1.0, abs(e_end)), "energy bookkeeping of the sequential chain drifted"

    # (c) production window sampler
    e2 = make_engine(cfg, det, marks, "fp32", intensity=max(1, n0))
    e2.add_objects(objs[:, :2], objs[:, 2:5])
    e2.run_windows(400, proposals_per_visit=16, n_warps=8, t0=temp, seed=2)
    rows = []
    for s in range(12000):
        e2.run_windows(1, proposals_per_visit=16, n_warps=8, t0=temp, seed=2, sweep_offset=400 + s, read_counters=False)
        hd, _, mk, _ = e2.read_objects()
        e = e2.energy_vectors(hd)[3] if len(hd) else 0.0
        rows.append([len(hd), e, *_size_hist(classes_of_marks(mk)[:, 0] if len(hd) else [])])
    mc, sc = _batch_se(np.array(rows, dtype=np.float64))

    names = ["count", "energy", "size<8", "size<16", "size<24", "size<32"]
    print(f"\n{cfg} T={temp}:")
    for i, nm in enumerate(names):
        b = f"{mb[i]:.4f}+-{sb[i]:.4f}" if i < 2 else "-"
        print(f"  {nm:8s} oracle {ma[i]:.4f}+-{sa[i]:.4f} | device chain {b} | windows {mc[i]:.4f}+-{sc[i]:.4f}")
    for i in range(2):
        assert abs(mb[i] - mc[i]) < 5 * np.hypot(sb[i], sc[i]), (names[i], "device chain vs windows", mb[i], mc[i], sb[i], sc[i])
        assert abs(ma[i] - mb[i]) < 5 * np.hypot(sa[i], sb[i]), (names[i], "oracle vs device chain", ma[i], mb[i], sa[i], sb[i])
    for i in range(len(names)):
        assert abs(ma[i] - mc[i]) < 5 * np.hypot(sa[i], sc[i]) + 1e-9, (names[i], "oracle vs windows", ma[i], mc[i], sa[i], sc[i])
    eng.close(); e2.close()
