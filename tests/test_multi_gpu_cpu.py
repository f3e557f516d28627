"""CPU tests (gloo, world_size 2) of the host-side multi-GPU logic: tile sharding, row bands, and the halo-exchange
protocol of a split scene driven over torch.distributed with a mock engine (no CUDA involved)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg


class MockEngine:
    """Holds objects as an [n, 8] float64 array (x, y, size, ratio, angle, cls, uid, 0).  A colour-row phase moves every
    object that lies in an owned window row of colour ci one column to the right (deterministic stand-in for sampling) and
    records the proposals it would have evaluated from the objects within 64 px (halo dependency)."""

    def __init__(self, height, width, objs):
        self.h, self.w = height, width
        self.o = np.array(objs, dtype=np.float64).reshape(-1, 8)
        self.log = []

    def window_grid(self, seed, sweep_id):
        return mg.grid_offset(seed, sweep_id)

    def run_window_rows(self, per_visit, n_warps, temperature, seed, sweep_id, ci, row_lo, row_hi):
        ox, _ = mg.grid_offset(seed, sweep_id)
        nwx = (self.h + ox + 31) // 32
        for wi in range(ci, nwx, 3):
            start = max(32 * wi - ox, 0)
            if not (row_lo <= start < row_hi):
                continue
            x0, x1 = start, min(32 * wi - ox + 32, self.h)
            inside = (self.o[:, 0] >= x0) & (self.o[:, 0] < x1)
            near = (self.o[:, 0] >= x0 - 64) & (self.o[:, 0] < x1 + 64)
            # what a window visit depends on: the uids of everything within 64 px, in canonical order
            self.log.append((sweep_id, wi, tuple(sorted(self.o[near, 6].astype(int).tolist())), float(self.o[near, 1].sum())))
            self.o[inside, 1] = (self.o[inside, 1] + 1) % self.w

    def pack_rows(self, lo, hi, capacity=8192):
        sel = (self.o[:, 0] >= lo) & (self.o[:, 0] < hi)
        return torch.as_tensor(self.o[sel].copy())

    def unpack_rows(self, lo, hi, records):
        keep = ~((self.o[:, 0] >= lo) & (self.o[:, 0] < hi))
        self.o = np.concatenate([self.o[keep], records.numpy().reshape(-1, 8)])

    def read_objects(self):
        return None, self.o[:, :2].astype(np.int32), self.o[:, 2:5], self.o[:, 6].astype(np.uint32)


def _objects(height, width, n, seed=0):
    rng = np.random.default_rng(seed)
    o = np.zeros((n, 8))
    o[:, 0] = rng.integers(0, height, n)
    o[:, 1] = rng.integers(0, width, n)
    o[:, 2:5] = rng.random((n, 3))
    o[:, 6] = np.arange(n)
    return o


def test_shard_items_and_row_bands():
    assert mg.shard_items(10, 4, 1) == [1, 5, 9]
    assert sorted(sum((mg.shard_items(256, 8, r) for r in range(8)), [])) == list(range(256))
    for h, world in ((8192, 8), (2048, 2), (300, 3), (96, 3)):
        bands = mg.row_bands(h, world)
        assert bands[0][0] == 0 and bands[-1][1] == h
        assert all(b[1] == bands[i + 1][0] for i, b in enumerate(bands[:-1]))
        assert all(b[0] % 32 == 0 and b[1] > b[0] for b in bands)
    with pytest.raises(ValueError):
        mg.row_bands(64, 3)
    for row, ox in ((0, 0), (0, 5), (64, 0), (64, 31), (70, 3)):
        g = mg.first_grid_line_at_or_after(row, ox)
        assert g >= row and (g + ox) % 32 == 0 and g - 32 < row


def _reference_run(height, width, objs, n_sweeps, seed):
    eng = MockEngine(height, width, objs)
    for s in range(n_sweeps):
        for ci in range(3):
            eng.run_window_rows(0, 0, 0.0, seed, s, ci, 0, height)
    return eng


def _worker(rank, world, port, height, width, objs, n_sweeps, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    band = mg.row_bands(height, world)[rank]
    eng = MockEngine(height, width, objs)
    scene = mg.SplitScene(eng, height, rank, world, capacity=512)
    eng.o = eng.o[scene.select_initial(eng.o[:, :2])]  # every rank starts with its band + halo only
    for s in range(n_sweeps):
        mg.sweep_dist(scene, 0, 0, 0.0, seed, s)
    xy, marks, uid = scene.owned_objects()
    out.put((rank, band, xy.tolist(), uid.tolist(), eng.log))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_split_scene_protocol_matches_single_process(world):
    height, width, n_sweeps, seed = 320, 128, 7, 5
    objs = _objects(height, width, 150)
    ref = _reference_run(height, width, objs, n_sweeps, seed)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, height, width, objs, n_sweeps, seed, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # union of the owned objects == the single-process configuration
    got = {}
    for rank, band, xy, uid, log in results:
        for (x, y), u in zip(xy, uid):
            assert band[0] <= x < band[1]
            assert u not in got
            got[u] = (x, y)
    want = {int(u): (int(x), int(y)) for x, y, u in zip(ref.o[:, 0], ref.o[:, 1], ref.o[:, 6])}
    assert got == want
    # every window visit saw exactly the neighbourhood (objects within 64 px and their columns) the single process saw
    ref_log = {(s, wi): (uids, cols) for s, wi, uids, cols in ref.log}
    seen = set()
    for rank, band, xy, uid, log in results:
        for s, wi, uids, cols in log:
            assert (s, wi) not in seen
            seen.add((s, wi))
            assert ref_log[(s, wi)] == (uids, cols), (rank, s, wi)
    assert seen == set(ref_log)


def test_band_geometry_of_the_peer_split():
    """Host-side geometry of the peer-access split: bands are whole 32-px cell rows covering the scene, map rows add 64 rows either
    side, every object is stored by exactly one rank, and undersized bands are refused."""
    import pytest
    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg

    class _Eng:  # the geometry needs no device
        pass

    for h, world in ((8192, 8), (4096, 3), (2048, 2), (1000, 2)):
        bands = mg.row_bands(h, world)
        assert bands[0][0] == 0 and bands[-1][1] == h and all(b[1] == n[0] for b, n in zip(bands, bands[1:]))
        assert all(b[0] % 32 == 0 for b in bands) and all(b[1] % 32 == 0 for b in bands[:-1])
        xy = np.stack([np.arange(h), np.zeros(h, dtype=int)], axis=1)
        owners = np.zeros(h, dtype=int)
        for r in range(world):
            sc = mg.PeerSplitScene(_Eng(), h, r, world)
            owners += sc.select_initial(xy).astype(int)
            m0, m1 = mg.PeerSplitScene.map_rows(h, r, world)
            assert m0 == max(0, sc.r0 - 64) and m1 == min(h, sc.r1 + 64)
        assert np.all(owners == 1)
    with pytest.raises(ValueError):
        mg.PeerSplitScene(_Eng(), 1000, 0, 4)   # 250-row bands: a window would reach beyond the neighbour band
    assert mg.shard_items(10, 4, 1) == [1, 5, 9] and sorted(sum((mg.shard_items(256, 8, r) for r in range(8)), [])) == list(range(256))


def test_sweep_planning_counts_the_windows_of_the_shifted_grids():
    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg
    from mpp_cnn_rs_object_detection_b200.api.rjmcmc import plan_sweeps
    shape, seed = (469, 753), 12345
    pv, n_sweeps, per_sweep = plan_sweeps(shape, seed, 500000, 96)
    total = 0
    for s in range(n_sweeps):
        ox, oy = mg.grid_offset(seed, s)
        total += ((shape[0] + ox + 31) // 32) * ((shape[1] + oy + 31) // 32) * pv
    assert pv == 96 and total >= 500000 and total - ((shape[0] + 62) // 32) * ((shape[1] + 62) // 32) * pv < 500000
    assert abs(per_sweep * n_sweeps - total) < 1e-6
    # a budget smaller than one sweep lowers the proposals per visit instead of overshooting by a whole sweep
    pv_small, n_small, _ = plan_sweeps((64, 64), 1, 40, 96)
    assert n_small >= 1 and pv_small <= 10
