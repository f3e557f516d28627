"""TEST INFRASTRUCTURE ONLY -- golden vectors for the optional split / merge kernels, produced by
running the UNMODIFIED reference (models/mpp/rjmcmc_sampler/kernels/split_and_merge_kernels.py) under the stubs of
oracle/ref_stubs.py.  Run in the build container:   python -m oracle.gen_golden_split_merge

Output tests/golden/split_merge_<cfg>.npz: the configuration (the one of energies_<cfg>.npz), and for every sampled
perturbation the removed indices, the added rectangles, the kernel's `data`, EPointsSet.energy_delta (raw and with the
shipped combinator) and the kernel's forward / backward probabilities."""
from __future__ import annotations

import os

import numpy as np

from oracle import gen_golden as gg
from oracle.gen_golden import BirthKernel, EPointsSet, Perturbation, Rectangle, make_kernels  # noqa: F401 (reference classes)


def gen(cfg_name: str, seed: int, shape, n_rect: int):
    objs, det, marks, image, setup, comb = gg.build_case(cfg_name, seed, shape, n_rect)
    rng = np.random.default_rng(seed + 100)
    config = gg.make_config(objs, rng, shape)  # same configuration as energies_<cfg>.npz
    ue, pe = setup.make_energies(image)
    krng = np.random.default_rng(seed + 7)
    kernels, p_kernels = make_kernels(image, intensity=max(1, len(config)), rng=krng, use_split_merge=True)
    split, merge = kernels[8], kernels[9]
    eps = EPointsSet(points=config, support_shape=shape, unit_energies_constructors=ue, pair_energies_constructors=pe)
    index = {id(p): k for k, p in enumerate(config)}
    rows = []
    for kern, kid in ((split, 8), (merge, 9)):
        got = 0
        while got < 30:
            u = kern.sample_perturbation(eps.points, krng)
            if u.removal is None and u.addition is None:
                rows.append([kid, -1, -1] + [np.nan] * 10 + [np.nan] * 5 + [float(u.data.get("n_neighbors", 0)) if u.data else 0.0,
                            0.0, 0.0, float(kern.forward_probability(eps.points, u)), float(kern.backward_probability(eps.points, u))])
                got += 1
                continue
            rem = u.removal if isinstance(u.removal, list) else [u.removal]
            add = u.addition if isinstance(u.addition, list) else [u.addition]
            d_raw = float(eps.energy_delta(u))
            d_comb = float(eps.energy_delta(u, energy_combinator=comb))
            f, b = float(kern.forward_probability(eps.points, u)), float(kern.backward_probability(eps.points, u))
            r_idx = [index[id(p)] for p in rem] + [-1] * (2 - len(rem))
            a_rows = sum(([float(p.x), float(p.y), float(p.size), float(p.ratio), float(p.angle)] for p in add), []) + [np.nan] * (5 * (2 - len(add)))
            data = list(u.data["pos_delta"]) + list(u.data["shape_delta"]) if kid == 8 else [np.nan] * 5
            rows.append([kid] + r_idx + a_rows + data + [float(u.data.get("n_neighbors", -1)) if kid == 9 else -1.0, d_raw, d_comb, f, b])
            got += 1
    assert len(eps) == len(config)
    out = os.path.join(gg.OUT, f"split_merge_{cfg_name}.npz")
    np.savez_compressed(out, seed=seed, shape=np.array(shape), n_rect=n_rect, config=np.array([gg.rect_row(r) for r in config]),
                        p_kernels=np.asarray(p_kernels, dtype=np.float64), intensity=float(max(1, len(config))),
                        rows=np.array(rows, dtype=np.float64),
                        columns=np.array(["kernel", "rem0", "rem1"] + [f"add0_{k}" for k in "xysra"] + [f"add1_{k}" for k in "xysra"] +
                                         ["pos_dx", "pos_dy", "d_size", "d_ratio", "d_angle", "n_neighbors", "delta_raw", "delta_comb", "fwd", "bwd"]),
                        maps_checksum=gg.checksum(det, *marks))
    print(out, len(rows), "perturbations")


if __name__ == "__main__":
    gen("legacy", seed=3, shape=(100, 140), n_rect=70)
    gen("nocalib", seed=4, shape=(128, 96), n_rect=60)
