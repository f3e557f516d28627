"""-m gpu parity tests: the CUDA path (through the C ABI) against the golden vectors of the reference and the
CPU oracle.  Tolerances: float32 energies within 1e-5 relative (absolute floor 1e-5 for near-zero terms, see
SURVEY.md section 7 'hard parts'); float64 instantiation within 1e-9; replay decisions identical."""
import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu

RTOL32, ATOL32 = 1e-5, 1e-5


def _load_case(cfg):
    g = gu.load(f"energies_{cfg}.npz")
    _, det, marks = gu.scene_inputs(g)
    return g, det, marks


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_energy_vectors_match_reference(cfg, precision):
    from tests.gpu_util import make_engine
    g, det, marks = _load_case(cfg)
    eng = make_engine(cfg, det, marks, precision)
    config = g["config"]
    handles = eng.add_objects(config[:, :2], config[:, 2:5])
    assert len(eng) == len(config)
    vec, comb, raw_total, comb_total = eng.energy_vectors(handles)
    rtol, atol = (RTOL32, ATOL32) if precision == "fp32" else (1e-9, 1e-9)
    if precision == "fp64" and cfg == "legacy":
        # ShapeEnergy is float32 arithmetic in the reference (sigmoid of float32 maps); expf differs from numpy's
        # float32 exp by <= 1 ulp, so this one column is only float32-exact even in the float64 instantiation
        np.testing.assert_allclose(vec[:, 1], g["vectors"][:, 1], rtol=3e-7, atol=1e-7)
        vec[:, 1] = g["vectors"][:, 1]
        atol = 1e-7
    np.testing.assert_allclose(vec, g["vectors"], rtol=rtol, atol=atol)
    np.testing.assert_allclose(comb, g["per_object_comb"], rtol=rtol, atol=atol)
    assert abs(comb_total - float(g["comb_total"])) <= rtol * abs(float(g["comb_total"])) + atol
    assert abs(raw_total - float(g["raw_total"])) <= 1e-4 + rtol * abs(float(g["raw_total"]))
    # position energies are float32 arithmetic in the reference too: bit-exact
    np.testing.assert_array_equal(vec[:, 0].astype(np.float32), g["vectors"][:, 0].astype(np.float32))
    # enumeration returns the same objects
    h2, xy, m, uid = eng.read_objects()
    assert sorted(h2.tolist()) == sorted(handles.tolist())
    order = np.argsort(h2)
    np.testing.assert_array_equal(xy[order], config[np.argsort(handles), :2].astype(np.int32))


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("combinator", [True, False])
def test_delta_batch_matches_reference(cfg, precision, combinator):
    from tests.gpu_util import make_engine, proposals_from_rows
    g, det, marks = _load_case(cfg)
    eng = make_engine(cfg, det, marks, precision, combinator=combinator)
    config = g["config"]
    uid = np.arange(len(config))
    eng.add_objects(config[:, :2], config[:, 2:5], uid=uid)
    rem = g["pert_removal"]
    m = len(rem)
    rem_rows = [(int(config[k, 0]), int(config[k, 1]), int(k)) if k >= 0 else (0, 0, -1) for k in rem]
    props = proposals_from_rows(np.zeros(m, np.int32), rem_rows, g["pert_addition"], np.arange(m) + 10000,
                                np.zeros((m, 2)), np.zeros(m, np.int32), np.zeros(m, np.int32), np.full(m, 0.5))
    got = eng.delta_batch(props)
    want = g["delta_comb"] if combinator else g["delta_raw"]
    if precision == "fp32":
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-5)
    else:
        # raw reference sums accumulate float32 position energies in set order (see test_oracle_golden)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=(1e-7 if cfg == "legacy" else 1e-9) if combinator else 2e-5)
    assert len(eng) == len(config)  # nothing was applied


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_replay_reproduces_reference_chain(cfg, precision):
    """Replays the recorded reference chain: identical accept/reject sequence and final configuration."""
    from tests.gpu_util import make_engine, proposals_from_rows
    g = gu.load(f"stream_{cfg}.npz")
    _, det, marks = gu.scene_inputs(g)
    eng = make_engine(cfg, det, marks, precision, intensity=float(g["intensity"]))
    init = g["init"]
    eng.add_objects(init[:, :2], init[:, 2:5], uid=np.arange(len(init)))
    rec = g["records"]
    m = len(rec)
    pos = {k: (int(init[k, 0]), int(init[k, 1])) for k in range(len(init))}
    rem_rows = []
    for row in rec:  # positions of removed objects come from the host mirror of accepted additions
        ruid, auid = int(row[1]), int(row[2])
        rem_rows.append(pos[ruid] + (ruid,) if ruid >= 0 else (0, 0, -1))
        if row[17] > 0:
            if auid >= 0:
                pos[auid] = (int(row[3]), int(row[4]))
    props = proposals_from_rows(rec[:, 0].astype(np.int32), rem_rows, rec[:, 3:8], rec[:, 2].astype(np.int64),
                                rec[:, 8:10], rec[:, 10].astype(np.int32), rec[:, 11].astype(np.int32), rec[:, 12])
    out = eng.replay(props, t0=float(g["t0"]), alpha_t=float(g["alpha_t"]), t_target=0.0)
    assert np.all(out["accepted"] >= 0)
    np.testing.assert_array_equal(out["temperature"], rec[:, 13])
    mism = np.nonzero(out["accepted"] != rec[:, 17].astype(np.int32))[0]
    if precision == "fp64":
        assert len(mism) == 0, f"decisions differ at steps {mism[:10]}"
    else:
        # float32 energies: a decision may only flip where the Green ratio is within float32 noise of log(u)
        first = mism[0] if len(mism) else m
        la_ref = (-rec[:first, 14] / rec[:first, 13]) + np.log(rec[:first, 16] + 1e-16) - np.log(rec[:first, 15] + 1e-16)
        np.testing.assert_allclose(out["delta_e"][:first], rec[:first, 14], rtol=1e-5, atol=2e-5)
        assert len(mism) == 0, (f"fp32 replay diverges at step {first}: log_alpha {out['log_alpha'][first]} vs "
                                f"log(u) {np.log(rec[first, 12] + 1e-16)}")
    np.testing.assert_allclose(out["fwd"], rec[:, 15], rtol=2e-6, atol=1e-300)
    np.testing.assert_allclose(out["bwd"], rec[:, 16], rtol=2e-6, atol=1e-300)
    np.testing.assert_array_equal(out["n_after"], rec[:, 18].astype(np.int32))
    tol = dict(rtol=1e-9, atol=1e-7 if cfg == "legacy" else 1e-9) if precision == "fp64" else dict(rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(out["delta_e"], rec[:, 14], **tol)
    # final configuration
    _, xy, marks_out, uid = eng.read_objects()
    final = g["final"]
    order = np.argsort(uid)
    np.testing.assert_array_equal(uid[order], final[:, 0].astype(np.uint32))
    np.testing.assert_array_equal(xy[order], final[:, 1:3].astype(np.int32))
    np.testing.assert_allclose(marks_out[order], final[:, 3:6], rtol=1e-6 if precision == "fp32" else 1e-15)


def test_birth_sampler_follows_density():
    """K5: inverse-CDF birth sampler draws pixels ~ det / sum(det) and classes ~ mark rows (chi-square)."""
    from tests.gpu_util import make_engine
    g, det, marks = _load_case("legacy")
    eng = make_engine("legacy", det, marks)
    n = 400000
    s = eng.sample_births(n, seed=123)
    h, w = det.shape
    assert s[:, 0].min() >= 0 and s[:, 0].max() < h and s[:, 1].min() >= 0 and s[:, 1].max() < w
    counts = np.bincount(s[:, 0] * w + s[:, 1], minlength=h * w).astype(np.float64)
    expect = det.reshape(-1).astype(np.float64) / det.astype(np.float64).sum() * n
    # pool pixels into 8x8 blocks so that every expected count is large
    hb, wb = h // 4, w // 4
    cb = counts.reshape(h, w)[:hb * 4, :wb * 4].reshape(hb, 4, wb, 4).sum((1, 3))
    eb = expect.reshape(h, w)[:hb * 4, :wb * 4].reshape(hb, 4, wb, 4).sum((1, 3))
    chi2 = ((cb - eb) ** 2 / eb).sum()
    dof = cb.size - 1
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), (chi2, dof)
    # mark classes at the most sampled pixel
    top = np.argmax(counts)
    sel = s[(s[:, 0] * w + s[:, 1]) == top]
    for i in range(3):
        p = marks[i][top // w, top % w].astype(np.float64)
        p /= p.sum()
        c = np.bincount(sel[:, 2 + i], minlength=32)
        assert abs(c[np.argmax(p)] / len(sel) - p.max()) < 5 * np.sqrt(p.max() * (1 - p.max()) / len(sel)) + 1e-3


def test_error_conventions():
    from mpp_cnn_rs_object_detection_b200._lib import MPPError, ERR_OUT_OF_BOUNDS, ERR_NOT_FOUND
    from tests.gpu_util import make_engine
    g, det, marks = _load_case("legacy")
    eng = make_engine("legacy", det, marks)
    with pytest.raises(MPPError) as e:
        eng.add_objects([[det.shape[0], 0]], [[8.0, 0.5, 0.1]])  # point_set.py:99 assert
    assert e.value.code == ERR_OUT_OF_BOUNDS
    h = eng.add_objects([[5, 5]], [[8.0, 0.5, 0.1]])
    eng.remove_objects(h)
    with pytest.raises(MPPError) as e:
        eng.remove_objects(h)  # KeyError in the reference (energy_point_set.py:88-100)
    assert e.value.code == ERR_NOT_FOUND
    assert len(eng) == 0


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_pair_overlap_structured_cases_match_oracle(precision):
    """mpp_pair_values (PairEnergy.compute, base_energies.py:77-80) on rectangle pairs that stress the intersection routine:
    identical, nested, collinear edges, touching from outside, touching at a corner, perpendicular, class-aligned angles,
    extreme size ratios -- against the oracle's polygon clip (RectangleOverlapEnergy, prior_energies.py:12-24)."""
    from oracle import mpp_oracle as orc
    from tests.gpu_util import make_engine
    h = w = 704
    det = np.full((h, w), 0.5, dtype=np.float32)
    marks = [np.full((h, w, 32), 1.0 / 32, dtype=np.float32) for _ in range(3)]
    eng = make_engine("nocalib", det, marks, precision)
    rng = np.random.default_rng(5)
    pi = np.pi
    # (dx, dy, sizeA, ratioA, angleA, sizeB, ratioB, angleB)
    cases = [(0, 0, 6, 0.5, 0.0, 6, 0.5, 0.0), (0, 4, 6, 0.5, 0.0, 6, 0.5, 0.0), (0, 8, 6, 0.5, 0.0, 6, 0.5, 0.0), (4, 8, 6, 0.5, 0.0, 6, 0.5, 0.0),
             (0, 1, 12, 0.75, 0.0, 3, 0.5, 0.0), (1, 0, 3, 0.5, 0.3, 12, 0.75, 0.3), (0, 0, 6, 0.5, 0.0, 6, 0.5, pi / 2), (2, 1, 8, 1.0, 0.0, 8, 1.0, pi / 4),
             (1, 1, 6, 0.5, 0.3, 6, 0.5, 0.3), (3, 2, 6, 0.5, 5 * pi / 32, 6, 0.5, 5 * pi / 32), (3, 2, 6, 0.5, 5 * pi / 32, 6, 0.5, 21 * pi / 32),
             (2, 2, 30, 0.2, 1.0, 1.0, 0.9, 2.0), (5, 5, 31, 1.0, 0.1, 0.7, 0.3, 0.2), (9, 0, 6, 0.5, 0.0, 6, 0.5, 0.0), (0, 0, 20, 1.0, 0.0, 20, 1.0, pi / 4)]
    for _ in range(200):
        cases.append((int(rng.integers(-12, 13)), int(rng.integers(-12, 13)), float(rng.uniform(1, 31)), float(rng.uniform(0.1, 1)),
                      float(rng.integers(0, 32) * pi / 32 if rng.random() < 0.5 else rng.uniform(0, pi)),
                      float(rng.uniform(1, 31)), float(rng.uniform(0.1, 1)), float(rng.integers(0, 32) * pi / 32 if rng.random() < 0.5 else rng.uniform(0, pi))))
    xy, mk = [], []
    for i, (dx, dy, sa, ra, aa, sb, rb, ab) in enumerate(cases):  # one pair per 40-px block
        cx, cy = 40 + 40 * (i % ((w - 80) // 40)), 40 + 40 * (i // ((w - 80) // 40))
        assert cx + 32 < h and cy + 32 < w, "scene too small for the structured cases"
        xy += [(cx, cy), (cx + dx, cy + dy)]
        mk += [(sa, ra, aa), (sb, rb, ab)]
    handles = eng.add_objects(np.array(xy), np.array(mk))
    got = eng.pair_values(handles[0::2], handles[1::2])[:, 0]
    want = np.array([orc.OracleScene.overlap_energy(orc.ORect(*xy[2 * i], *mk[2 * i]), orc.ORect(*xy[2 * i + 1], *mk[2 * i + 1]))
                     if (xy[2 * i][0] - xy[2 * i + 1][0]) ** 2 + (xy[2 * i][1] - xy[2 * i + 1][1]) ** 2 <= 32 ** 2 else 0.0 for i in range(len(cases))])
    if precision == "fp32":
        # float32 geometry, clipped in the frame of the thinner rectangle: 5e-6 of the smaller area down to 0.1-px half-sides
        # (tools/clip_check.cu over 4 M pairs), so the north-star bound holds for the sub-pixel cases 12 and 13 too
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
