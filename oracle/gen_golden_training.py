"""TEST INFRASTRUCTURE ONLY -- golden vectors for the training-time helpers, produced by running
the UNMODIFIED reference (models/mpp/perturbation_sampler.py, models/mpp/energies/energy_utils.py) under the stubs of
oracle/ref_stubs.py.  Run in the build container:   python -m oracle.gen_golden_training

Output tests/golden/training_helpers.npz: the configuration of energies_legacy.npz, the configurations returned by
sample_perturbations for three presets with a seeded Generator, aggregate_perturbations on a scripted sequence, and
compute_many_energy_vectors of the perturbed configurations."""
from __future__ import annotations

import os

import numpy as np

from oracle import gen_golden as gg
from oracle.gen_golden import Perturbation, Rectangle  # noqa: F401 (reference classes)

from models.mpp.energies.energy_utils import compute_many_energy_vectors  # noqa: E402 (reference, importable after gen_golden)
from models.mpp.perturbation_sampler import (PERTURBATION_LIGHT, PERTURBATION_MEDIUM_OVERLAP, PERTURBATION_STRONG,  # noqa: E402
                                             aggregate_perturbations, sample_perturbations)


def rows_of(cfgs):
    flat, lens = [], []
    for c in cfgs:
        lens.append(len(c))
        flat += [gg.rect_row(r) for r in c]
    return np.array(flat, dtype=np.float64).reshape(-1, 5), np.array(lens)


def gen():
    cfg_name, seed, shape, n_rect = "legacy", 3, (100, 140), 70
    objs, det, marks, image, setup, comb = gg.build_case(cfg_name, seed, shape, n_rect)
    rng = np.random.default_rng(seed + 100)
    config = gg.make_config(objs, rng, shape)
    image.gt_config = config
    out = {"config": np.array([gg.rect_row(r) for r in config]), "shape": np.array(shape), "seed": seed, "n_rect": n_rect,
           "maps_checksum": gg.checksum(det, *marks)}
    all_cfgs = []
    for name, preset in (("light", PERTURBATION_LIGHT), ("overlap", PERTURBATION_MEDIUM_OVERLAP), ("strong", PERTURBATION_STRONG)):
        prng = np.random.default_rng(seed + 11)
        cfgs = sample_perturbations(image_data=image, rng=prng, n_samples=3, **preset)
        flat, lens = rows_of(cfgs)
        out[f"pert_{name}"], out[f"pert_{name}_len"] = flat, lens
        all_cfgs += cfgs
    ue, pe = setup.make_energies(image)
    names = setup.energy_names
    vec = compute_many_energy_vectors(all_cfgs, image, ue, pe, names, multiprocess=False)
    out["vectors"], out["energy_names"] = vec, np.array(names)
    # aggregate_perturbations on a scripted sequence over objects 0..4 (a = added object ids 100+)
    a = [Rectangle(x=5 + k, y=6 + k, size=8.0, ratio=0.5, angle=0.1 * k) for k in range(3)]
    seq = [Perturbation(type=None, removal=config[0], addition=a[0]), Perturbation(type=None, removal=a[0], addition=a[1]),
           Perturbation(type=None, removal=[config[1], config[2]], addition=a[2]), Perturbation(type=None, removal=None, addition=config[1]),
           Perturbation(type=None, removal=a[2], addition=None)]
    agg = aggregate_perturbations(seq)
    idx = {id(p): k for k, p in enumerate(config)}
    idx.update({id(p): 100 + k for k, p in enumerate(a)})
    out["agg_removal"] = np.array(sorted(idx[id(p)] for p in agg.removal))
    out["agg_addition"] = np.array(sorted(idx[id(p)] for p in agg.addition))
    path = os.path.join(gg.OUT, "training_helpers.npz")
    np.savez_compressed(path, **out)
    print(path, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    gen()
