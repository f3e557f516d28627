"""TEST / BENCH INFRASTRUCTURE ONLY -- times the CPU restatement of the reference sampler (oracle/mpp_oracle.py).

This is the `cpu_baseline` / `--impl reference` leg of bench.py: the reference's own production decomposition
(models/mpp/mpp_model.py:231-264: the image is cut into 256x256 patches and every patch runs an independent
sequential RJMCMC chain in a multiprocessing.Pool with os.cpu_count() workers), restated with the oracle's
sequential sampler (OracleSampler == RJMCMC.run, rjmcmc.py:83-181).  As in SURVEY.md section 8d only `.run()` is
timed (make_energies / make_kernels are excluded); throughput = sum of proposals / slowest worker's run time.

The product package never imports this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import List, Sequence, Tuple

import numpy as np

PATCH = 256  # models/mpp/mpp_model.py:231 (patch_size = 256)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _chain(args):
    (det, marks, objs, setup, calib, comb_kind, comb_args, n_steps, n_warm, seed, t0, alpha_t) = args
    from oracle import mpp_oracle as orc
    if setup == "legacy":
        scene = orc.OracleScene(det, marks, setup="legacy", detection_threshold=calib["detection_threshold"],
                                remap_coefs=calib["coefs"], remap_intercepts=calib["intercepts"],
                                min_area=calib["min_area"], max_area=calib["max_area"])
    else:
        scene = orc.OracleScene(det, marks, setup="nocalib", detection_threshold=0.0, min_area=calib["min_area"],
                                max_area=calib["max_area"], ratio_prior=True)
    comb = orc.OracleHierarchical(**comb_args) if comb_kind == "hierarchical" else orc.OracleLogistic(**comb_args)
    sampler = orc.OracleSampler(scene, comb, [orc.ORect(*r) for r in objs], np.random.default_rng(seed), t0, alpha_t)
    sampler.run(n_warm)
    a0 = sampler.n_accepted
    t = time.perf_counter()
    sampler.run(n_steps)
    dt = time.perf_counter() - t
    return n_steps, dt, sampler.n_accepted - a0, len(sampler.state)


def crop_patches(det: np.ndarray, marks: Sequence[np.ndarray], objs: np.ndarray, n_patches: int) -> List[Tuple]:
    """The first n_patches non-overlapping 256^2 patches (row-major) with the objects whose centre falls inside."""
    h, w = det.shape
    out = []
    for i in range(0, h, PATCH):
        for j in range(0, w, PATCH):
            if len(out) >= n_patches:
                return out
            i1, j1 = min(h, i + PATCH), min(w, j + PATCH)
            sel = (objs[:, 0] >= i) & (objs[:, 0] < i1) & (objs[:, 1] >= j) & (objs[:, 1] < j1)
            o = objs[sel].copy()
            o[:, 0] -= i
            o[:, 1] -= j
            out.append((np.ascontiguousarray(det[i:i1, j:j1]), [np.ascontiguousarray(m[i:i1, j:j1]) for m in marks], o))
    return out


def _single_chain(args):
    (det, marks, objs, setup, calib, comb_kind, comb_args, n_steps, n_warm, seed, t0) = args
    from oracle import mpp_oracle as orc
    t_setup = time.perf_counter()
    scene = orc.OracleScene(det, marks, setup="legacy", detection_threshold=calib["detection_threshold"],
                            remap_coefs=calib["coefs"], remap_intercepts=calib["intercepts"],
                            min_area=calib["min_area"], max_area=calib["max_area"])
    comb = orc.OracleHierarchical(**comb_args)
    sampler = orc.OracleSampler(scene, comb, [orc.ORect(*r) for r in objs], np.random.default_rng(seed), t0, 1.0)
    t_setup = time.perf_counter() - t_setup
    sampler.run(n_warm)
    a0 = sampler.n_accepted
    t = time.perf_counter()
    sampler.run(n_steps)
    dt = time.perf_counter() - t
    return n_steps, dt, sampler.n_accepted - a0, t_setup


def single_chain_whole_scene(det, marks, objs, setup, calib, comb_kind, comb_args, n_steps=300, n_warm=20, seed=0, t0=0.02):
    """ONE sequential chain over the whole scene on one core (RJMCMC.run rjmcmc.py:172-181 without the patch tiling of
    mpp_model.py:231-264): the data-driven birth draws from all H*W pixels (sampler2d.py:43).  Runs in a child process so that the
    scene-sized float maps it derives are released afterwards.  Returns (proposals, seconds, accepted, setup seconds)."""
    with mp.get_context("fork").Pool(1) as pool:
        return pool.apply(_single_chain, ((det, marks, objs, setup, calib, comb_kind, comb_args, n_steps, n_warm, seed, t0),))


def image_time_from_patch_rates(rates: Sequence[float], steps_per_patch: int, workers: int) -> float:
    """Seconds the reference decomposition needs for one image: every patch runs `steps_per_patch` steps at its measured rate,
    patches are handed to `workers` processes longest first (Pool.map hands them out in order; longest-first is the better case)."""
    loads = [0.0] * max(1, workers)
    for t in sorted((steps_per_patch / r for r in rates), reverse=True):
        k = loads.index(min(loads))
        loads[k] += t
    return max(loads)


class PatchPool:
    """A pool of worker processes, one 256^2 patch chain each (mpp_model.py:262 `_map_to_images(multiprocess=True)`)."""

    def __init__(self, patches, setup, calib, comb_kind, comb_args, workers=None, t0=0.02, alpha_t=1.0):
        self.workers = min(workers or host_cores(), len(patches))
        self.patches = patches[:self.workers]
        self.model = (setup, calib, comb_kind, comb_args)
        self.t0, self.alpha_t = t0, alpha_t
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None

    def run(self, steps_per_worker: int, warm: int = 0, seed: int = 0):
        """Returns (proposals, seconds (slowest worker), accepted, objects)."""
        setup, calib, kind, cargs = self.model
        jobs = [(d, m, o, setup, calib, kind, cargs, steps_per_worker, warm, seed + k, self.t0, self.alpha_t)
                for k, (d, m, o) in enumerate(self.patches)]
        res = self.pool.map(_chain, jobs, chunksize=1) if self.pool else [_chain(j) for j in jobs]
        return (sum(r[0] for r in res), max(r[1] for r in res), sum(r[2] for r in res), sum(r[3] for r in res))

    def run_each(self, patches, steps_per_patch: int, warm: int = 0, seed: int = 0):
        """Every patch of `patches` (any number; the pool works through them) for `steps_per_patch` steps: list of
        (proposals, seconds, accepted, objects) per patch."""
        setup, calib, kind, cargs = self.model
        jobs = [(d, m, o, setup, calib, kind, cargs, steps_per_patch, warm, seed + k, self.t0, self.alpha_t)
                for k, (d, m, o) in enumerate(patches)]
        return self.pool.map(_chain, jobs, chunksize=1) if self.pool else [_chain(j) for j in jobs]

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()
            self.pool = None
