"""-m gpu: the multi-scene / split-scene schedule (mpp_run_windows_batch, csrc/mpp_multi.cuh).

* A batch of tiles sampled in ONE launch: every tile ends in exactly the configuration mpp_run_windows gives it alone.
* One scene split into row bands, each band sampled by its own context and its own concurrently running kernel (here on one
  GPU, separate streams; on a multi-GPU box one process per band, tools/split_check.py), the cells across a band boundary read
  and written directly in the neighbour's context: the union of the bands equals the single-context chain, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sorted_state(eng, rows=None):
    _, xy, mk, uid = eng.read_objects()
    if rows is not None:
        sel = (xy[:, 0] >= rows[0]) & (xy[:, 0] < rows[1])
        xy, mk, uid = xy[sel], mk[sel], uid[sel]
    return xy, mk, uid


def _canon(xy, mk, uid):
    order = np.lexsort((uid, xy[:, 1], xy[:, 0]))
    return xy[order], mk[order], uid[order]


@pytest.mark.parametrize("cfg", ["legacy", "nocalib"])
def test_tile_batch_equals_tiles_sampled_alone(cfg):
    from mpp_cnn_rs_object_detection_b200 import synth
    from mpp_cnn_rs_object_detection_b200.engine import run_windows_batch
    from tests.gpu_util import make_engine
    shape, seed, n_sweeps, pv = (128, 160), 7, 12, 24
    tiles = [synth.make_scene(20 + k, shape, 10 + 5 * k) for k in range(6)]
    alone, batch = [], []
    cnt_alone = np.zeros(5, dtype=np.int64)
    for objs, det, marks in tiles:
        e = make_engine(cfg, det, marks, "fp32", intensity=max(1, len(objs)))
        e.add_objects(objs[:, :2], objs[:, 2:5], uid=np.arange(len(objs)))
        cnt_alone += np.array(e.run_windows(n_sweeps, pv, n_warps=8, t0=0.05, alpha_t=0.9, t_target=0.01, seed=seed)[:5])
        alone.append(_canon(*_sorted_state(e)))
        e.close()
    engines = []
    for objs, det, marks in tiles:
        e = make_engine(cfg, det, marks, "fp32", intensity=max(1, len(objs)))
        e.add_objects(objs[:, :2], objs[:, 2:5], uid=np.arange(len(objs)))
        engines.append(e)
    # two calls: the completion stamps carry over from one launch to the next
    c1 = run_windows_batch(engines, [seed] * len(engines), 5, pv, n_warps=8, t0=0.05, alpha_t=0.9, t_target=0.01)
    c2, maxdiff = run_windows_batch(engines, [seed] * len(engines), n_sweeps - 5, pv, n_warps=8, t0=0.05 * 0.9 ** 5, alpha_t=0.9, t_target=0.01,
                                    sweep_offset=5, debug=True)
    assert maxdiff < 2e-5
    np.testing.assert_array_equal(np.array(c1[:5]) + np.array(c2[:5]), cnt_alone)
    for e, want in zip(engines, alone):
        got = _canon(*_sorted_state(e))
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)
        e.close()


@pytest.mark.parametrize("bands", [2, 3])
def test_split_scene_with_peer_access_equals_single_context(bands):
    import torch
    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg, synth
    from mpp_cnn_rs_object_detection_b200.engine import run_windows_batch
    from tests.gpu_util import make_engine
    shape, seed, n_sweeps, pv = (416 * bands, 128), 5, 8, 16
    objs, det, marks = synth.make_scene(31, shape, 60 * bands)
    det_d, marks_d = torch.as_tensor(det).cuda(), torch.stack([torch.as_tensor(m) for m in marks]).cuda()
    det_sum = float(np.sum(det))
    uid = np.arange(len(objs))
    ref = make_engine("legacy", det, marks, "fp32", intensity=max(1, len(objs)))
    ref.add_objects(objs[:, :2], objs[:, 2:5], uid=uid)
    cnt_ref = ref.run_windows(n_sweeps, pv, n_warps=8, t0=0.04, seed=seed)
    want = _canon(*_sorted_state(ref))
    ref.close()

    rows = mg.row_bands(shape[0], bands)
    streams = [torch.cuda.Stream() for _ in range(bands)]
    engines = []
    for b in range(bands):
        with torch.cuda.stream(streams[b]):
            e = make_engine("legacy", det_d, marks_d, "fp32", intensity=max(1, len(objs)))
            e.set_maps(det_d, marks_d, det_sum=det_sum)
            sel = (objs[:, 0] >= rows[b][0]) & (objs[:, 0] < rows[b][1])
            e.add_objects(objs[sel, :2], objs[sel, 2:5], uid=uid[sel])
            engines.append(e)
    for b, e in enumerate(engines):
        e.split_attach_local(rows[b][0], rows[b][1], engines[b - 1] if b > 0 else None, engines[b + 1] if b < bands - 1 else None)
    torch.cuda.synchronize()
    # all bands are launched before anything synchronises: each kernel waits for its neighbours' boundary windows
    for k in range(2):
        for b, e in enumerate(engines):
            run_windows_batch([e], [seed], n_sweeps // 2, pv, n_warps=8, t0=0.04, sweep_offset=k * (n_sweeps // 2), max_ctas=48,
                              read_counters=False)
    torch.cuda.synchronize()
    parts, cnt = [], np.zeros(5, dtype=np.int64)
    for b, e in enumerate(engines):
        cnt += np.array(e.run_windows(0, pv, n_warps=8, t0=0.04)[:5])
        parts.append(_sorted_state(e, rows[b]))
        _, xy_all, _, _ = e.read_objects()
        assert np.all((xy_all[:, 0] // 32 >= rows[b][0] // 32) & (xy_all[:, 0] // 32 < -(-rows[b][1] // 32))), "a band holds only its own cell rows"
    got = _canon(*[np.concatenate([p[i] for p in parts]) for i in range(3)])
    np.testing.assert_array_equal(cnt, np.array(cnt_ref[:5]))
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    for e in engines:
        e.split_detach()
        e.close()
