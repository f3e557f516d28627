"""Perturbed configurations for the weight-learning loops: models/mpp/perturbation_sampler.py.

`sample_perturbations` jitters a ground-truth configuration on the host (numpy draws in the reference's order, so the same
seeded Generator gives the same configurations); `sample_kernel_perturbations` walks the proposal kernels of the sampler
(device-backed, api/kernels.py) for a number of steps and can aggregate the walk into one net Perturbation."""
from __future__ import annotations

from copy import copy
from typing import List, Sequence, Tuple

import numpy as np

from .custom_types import ImageWMaps, Perturbation
from .energy_point_set import EPointsSet
from .kernels import make_kernels
from .mappings import ValueMapping
from .shapes import Rectangle

# perturbation_sampler.py:16-56
PERTURBATION_LIGHT = {"move_proba": 0.1, "param_shift_proba": [0.1, 0.1, 0.1], "position_sigma": 1, "param_sigmas": [0.02, 0.02, 0.02],
                      "point_number_sigma": 0.1, "no_addition": True}
PERTURBATION_MEDIUM = {"move_proba": 0.5, "param_shift_proba": [0.5, 0.5, 0.5], "position_sigma": 5, "param_sigmas": [0.1, 0.1, 0.1],
                       "point_number_sigma": 1.0}
PERTURBATION_HP_MEDIUM = {"move_proba": 0.8, "param_shift_proba": [0.9, 0.9, 0.9], "position_sigma": 5, "param_sigmas": [0.1, 0.1, 0.1],
                          "point_number_sigma": 1.0}
PERTURBATION_MEDIUM_OVERLAP = {"move_proba": 0.8, "param_shift_proba": [0.9, 0.9, 0.9], "position_sigma": 5, "param_sigmas": [0.1, 0.1, 0.1],
                               "point_number_sigma": 5.0, "make_overlap": 0.9}
PERTURBATION_STRONG = {"move_proba": 0.9, "param_shift_proba": [0.9, 0.9, 0.9], "position_sigma": 20, "param_sigmas": [0.5, 0.5, 0.5],
                       "point_number_sigma": 10.0}


def sample_perturbations(image_data: ImageWMaps = None, gt_rectangles: List[Rectangle] = None, rng: np.random.Generator = None,
                         image_shape: Tuple[int, int] = None, mappings: List[ValueMapping] = None, move_proba: float = None,
                         param_shift_proba: List[float] = None, position_sigma: float = None, param_sigmas: List[float] = None,
                         make_overlap: float = None, no_addition: bool = False, point_number_sigma: float = None,
                         n_samples: int = 1) -> List[List[Rectangle]]:
    """`n_samples` jittered copies of a configuration (perturbation_sampler.py:59-124): the object count is redrawn around the
    original one (objects dropped at random, or added uniformly / as copies of existing ones with probability
    `make_overlap`), then every object is moved with probability `move_proba` and each mark shifted with probability
    `param_shift_proba[i]` (cyclic marks wrap, the others are clipped)."""
    if image_data is not None:
        gt_rectangles, image_shape, mappings = image_data.gt_config, image_data.shape, image_data.mappings
    else:
        assert gt_rectangles is not None and image_shape is not None and mappings is not None
    results = []
    for _ in range(n_samples):
        new_points = [copy(p) for p in gt_rectangles]
        n0 = len(gt_rectangles)
        n1 = int(np.clip(rng.normal(n0, point_number_sigma), a_min=0, a_max=1e4))
        if no_addition:
            n1 = int(np.clip(n1, 0, n0))
        if n1 < n0:
            keep = rng.choice(range(n0), size=n1, replace=False)
            new_points = [new_points[i] for i in keep]
        elif n1 > n0:
            for _ in range(n1 - n0):
                if make_overlap is not None and rng.random() <= make_overlap:
                    new_points.append(copy(rng.choice(new_points)))
                else:
                    pos = rng.integers((0, 0), image_shape)
                    params = {name: rng.uniform(m.v_min, m.v_max) for name, m in zip(Rectangle.PARAMETERS, mappings)}
                    new_points.append(Rectangle(x=pos[0], y=pos[1], **params))
        for p in new_points:
            if rng.random() < move_proba:
                shift = rng.normal(0, position_sigma, size=2)
                p.x, p.y = (int(v) for v in np.clip((p.x + shift[0], p.y + shift[1]), (0, 0), (image_shape[0] - 1, image_shape[1] - 1)).astype(int))
            for i, (mapping, name) in enumerate(zip(mappings, Rectangle.PARAMETERS)):
                if rng.random() < param_shift_proba[i]:
                    v_min, v_max = mapping.v_min, mapping.v_max
                    value = getattr(p, name) + rng.normal(0, param_sigmas[i] * (v_max - v_min))
                    if mapping.is_cyclic:
                        value = ((value - v_min) % (v_max - v_min)) + v_min
                    setattr(p, name, float(np.clip(value, v_min, v_max)))
        results.append(new_points)
    return results


class DummyKernel:  # perturbation_sampler.py:172-173: the `type` of an aggregated perturbation
    pass


def aggregate_perturbations(perturbations: Sequence[Perturbation]) -> Perturbation:
    """Net effect of a sequence of perturbations (perturbation_sampler.py:176-211): an addition cancels an earlier removal of the
    same object and vice versa."""
    additions, removals = {}, {}  # insertion-ordered sets keyed by object identity
    for p in perturbations:
        added = p.addition if isinstance(p.addition, list) else ([] if p.addition is None else [p.addition])
        removed = p.removal if isinstance(p.removal, list) else ([] if p.removal is None else [p.removal])
        for q in added:
            if q in removals:
                del removals[q]
            else:
                additions[q] = None
        for q in removed:
            if q in additions:
                del additions[q]
            else:
                removals[q] = None
    return Perturbation(type=DummyKernel, removal=list(removals), addition=list(additions))


def sample_kernel_perturbations(kernels, p_kernels: Sequence[float], iter_per_point: float, points: EPointsSet, rng: np.random.Generator,
                                aggregate_pert: bool = False):
    """A walk of int(iter_per_point * len(points)) unconditional kernel moves from `points` (perturbation_sampler.py:154-169)."""
    assert len(kernels) == len(p_kernels)
    new_points = points.copy()
    perturbations = []
    for _ in range(int(iter_per_point * len(points))):
        kernel = kernels[int(rng.choice(len(kernels), p=p_kernels))]
        pert = kernel.sample_perturbation(x=new_points.points, rng=rng)
        perturbations.append(pert)
        new_points = new_points.apply_perturbation(pert)
    if aggregate_pert:
        perturbations = aggregate_perturbations(perturbations)
    return new_points, perturbations


def sample_multiple_kernel_perturbations(image_data: ImageWMaps, n_samples: int, rng: np.random.Generator, energy_setup, iter_per_point: float,
                                         return_perturbations: bool = False, aggregate_pert: bool = False, use_split_merge: bool = False):
    """`n_samples` independent kernel walks from the ground-truth configuration (perturbation_sampler.py:127-151)."""
    points = getattr(image_data, "gt_config_set", None)
    if points is None:
        uec, pec = energy_setup.make_energies(image_data=image_data)
        points = EPointsSet(points=image_data.gt_config, support_shape=image_data.shape, unit_energies_constructors=uec,
                            pair_energies_constructors=pec)
    kernels, p_kernels = make_kernels(image_data, intensity=1.0, rng=rng, use_split_merge=use_split_merge)
    results, perts = [], []
    for _ in range(n_samples):
        new_points, perturbations = sample_kernel_perturbations(kernels=kernels, p_kernels=p_kernels, points=points, rng=rng,
                                                                iter_per_point=iter_per_point, aggregate_pert=aggregate_pert)
        results.append(new_points)
        perts.append(perturbations)
    return perts if return_perturbations else results
