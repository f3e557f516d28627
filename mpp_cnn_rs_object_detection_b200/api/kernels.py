"""Proposal kernels of the RJMCMC sampler: models/mpp/rjmcmc_sampler/kernels/{base_kernels,transform_kernels,
make_kernels}.py.  The classes keep the reference's interface (sample_perturbation / forward_probability /
backward_probability); draws and probabilities are computed on the device (mpp_sample_proposals, mpp_proposal_probs)
with Philox randomness seeded from the numpy Generator the caller passes."""
from __future__ import annotations

from abc import abstractmethod
from copy import copy
from typing import List

import numpy as np

from .. import _lib
from ..engine import kernel_probabilities
from .custom_types import ImageWMaps, Perturbation
from .point_set import PointsSet
from .shapes import Rectangle

BASE_KERNEL_WEIGHTS = {  # make_kernels.py:13-24
    "bd_weight": 1, "uniform_bd_weight": 1, "data_bd_weight": 2, "ms_weight": 1, "translation_weight": 1,
    "gaussian_translation_weight": 1, "data_translation_weight": 2, "transformation_weight": 1,
    "gaussian_transformation_weight": 1, "data_transformation_weight": 2}


class Kernel:  # base_kernels.py:13-28
    @abstractmethod
    def sample_perturbation(self, x: PointsSet, rng: np.random.Generator) -> Perturbation:
        pass

    @abstractmethod
    def forward_probability(self, x: PointsSet, u: Perturbation) -> float:
        pass

    @abstractmethod
    def backward_probability(self, x: PointsSet, u: Perturbation) -> float:
        pass


class KernelSet:
    """Parameters shared by the eight kernels built by make_kernels (what mpp_set_kernels receives)."""

    def __init__(self, intensity: float, p_kernels: np.ndarray, translation_sigma: float = 2.0, max_delta: int = 8,
                 transform_sigma: float = 0.1):
        self.intensity = float(np.sum(intensity))  # base_kernels.py:49
        self.p_kernels = np.asarray(p_kernels, dtype=np.float64)
        self.translation_sigma, self.max_delta, self.transform_sigma = translation_sigma, max_delta, transform_sigma

    def bind(self, x: PointsSet):
        st = x._state
        if st.maps is None:
            raise ValueError("the proposal kernels need a points set built with the map-driven energy terms")
        st.use_kernels(self.intensity, self.p_kernels, self.translation_sigma, self.max_delta, self.transform_sigma)
        return st


class DeviceKernel(Kernel):
    KERNEL_ID = -1

    def __init__(self, kset: KernelSet, p_kernel: float):
        self._kset = kset
        self._p = float(p_kernel)

    @property
    def p_kernel(self) -> float:
        return self._p

    def sample_perturbation(self, x: PointsSet, rng: np.random.Generator) -> Perturbation:
        st = self._kset.bind(x)
        rec = st.engine.sample_proposals([self.KERNEL_ID], seed=int(rng.integers(0, 2 ** 62)))
        return self._to_perturbation(st, rec)

    def _to_perturbation(self, st, rec) -> Perturbation:
        r = rec[0]
        removal = st.by_uid.get(int(r["rem_uid"])) if int(r["rem_uid"]) != _lib.NO_OBJECT else None
        addition = None
        if int(r["add_uid"]) != _lib.NO_OBJECT:
            if removal is not None:
                addition = copy(removal)  # moves create a new object with a new identity (transform_kernels.py:33,84,141,183)
                addition.x, addition.y = int(r["add_x"]), int(r["add_y"])
                if int(r["kernel"]) in (6, 7):
                    setattr(addition, Rectangle.PARAMETERS[int(r["param_id"])],
                            float((r["add_size"], r["add_ratio"], r["add_angle"])[int(r["param_id"])]))
            else:
                addition = Rectangle(int(r["add_x"]), int(r["add_y"]), float(r["add_size"]), float(r["add_ratio"]), float(r["add_angle"]))
        data = {"record": rec, "delta": np.array([r["delta0"], r["delta1"]]) if int(r["kernel"]) == 4 else float(r["delta0"]),
                "param_id": int(r["param_id"]), "new_param_class_value": int(r["new_class"])}
        return Perturbation(self.__class__, removal=removal, addition=addition, data=data)

    def _record(self, st, u: Perturbation) -> np.ndarray:
        d = u.data or {}
        delta = d.get("delta", 0.0)
        delta = (float(delta[0]), float(delta[1])) if np.ndim(delta) else (float(delta), 0.0)
        return st.proposal_record(u.removal, u.addition, kernel=self.KERNEL_ID, delta=delta, param_id=int(d.get("param_id", 0)),
                                  new_class=int(d.get("new_param_class_value", 0)))

    def _probs(self, x: PointsSet, u: Perturbation):
        assert u.type == self.__class__
        st = self._kset.bind(x)
        d = u.data if u.data is not None else {}
        key = ("probs", len(st))
        if key not in d:
            d[key] = st.engine.proposal_probs(self._record(st, u))[0]
            u.data = d
        return d[key]

    def forward_probability(self, x: PointsSet, u: Perturbation) -> float:
        return float(self._probs(x, u)[0])

    def backward_probability(self, x: PointsSet, u: Perturbation) -> float:
        return float(self._probs(x, u)[1])


class BirthKernel(DeviceKernel):  # base_kernels.py:31-71
    def __init__(self, kset: KernelSet = None, p_kernel: float = 0.0, data_driven: bool = False):
        super().__init__(kset, p_kernel)
        self.KERNEL_ID = 2 if data_driven else 0
        self.p_birth = self.p_death = float(p_kernel)

    def __repr__(self):
        return "BirthKernel"


class DeathKernel(DeviceKernel):  # base_kernels.py:74-122
    def __init__(self, kset: KernelSet = None, p_kernel: float = 0.0, data_driven: bool = False):
        super().__init__(kset, p_kernel)
        self.KERNEL_ID = 3 if data_driven else 1
        self.p_birth = self.p_death = float(p_kernel)

    def __repr__(self):
        return "DeathKernel"


class GaussianTranslationKernel(DeviceKernel):  # transform_kernels.py:17-58
    KERNEL_ID = 4


class DataDrivenTranslationKernel(DeviceKernel):  # transform_kernels.py:61-116
    KERNEL_ID = 5


class GaussianShapeTransformKernel(DeviceKernel):  # transform_kernels.py:119-159
    KERNEL_ID = 6


class DataDrivenShapeTransformKernel(DeviceKernel):  # transform_kernels.py:162-225
    KERNEL_ID = 7


# ---------------------------------------------------------------------------------------------- optional split / merge
class SplitSampler:  # split_and_merge_kernels.py:14-36
    def __init__(self, pos_radius: float, shape_sigmas: List[float], mappings):
        self.pos_radius, self.shape_sigmas, self.mappings = pos_radius, shape_sigmas, mappings
        self.scaled_shaped_sigmas = [s * m.range for m, s in zip(mappings, shape_sigmas)]
        self.n_params = len(Rectangle.PARAMETERS)

    def sample(self, rng: np.random.Generator):
        pos = rng.uniform((0, 0), self.pos_radius)
        while np.linalg.norm(pos) > self.pos_radius:
            pos = rng.uniform((0, 0), self.pos_radius)
        return pos, rng.normal((0,) * self.n_params, self.scaled_shaped_sigmas)

    def pdf(self, pos_deltas, shape_deltas) -> float:
        p_pos = 1 / (np.pi * self.pos_radius * self.pos_radius)
        p_shape = [np.exp(-(d / s) ** 2 / 2) / (np.sqrt(2 * np.pi) * s) for d, s in zip(shape_deltas, self.scaled_shaped_sigmas)]
        return float(p_pos * np.prod(p_shape))


class _SplitMergeBase(Kernel):
    """The optional two-object moves (use_split_merge=False in both shipped configurations).  Draws follow the reference's
    numpy calls; neighbourhood counts come from the device index and the Delta-energy of the two-object perturbation from
    the device (EnergyGraph._delta_multi).  They run through the step-by-step RJMCMC loop only."""

    def __init__(self, p_split: float, p_merge: float, split_sampler: SplitSampler, support_shape, intensity: float, merge_radius: float):
        self.p_split, self.p_merge, self.split_sampler = p_split, p_merge, split_sampler
        self.shape, self.intensity, self.radius = support_shape, intensity, merge_radius
        assert self.radius == self.split_sampler.pos_radius


class SplitKernel(_SplitMergeBase):  # split_and_merge_kernels.py:39-107
    def sample_perturbation(self, x: PointsSet, rng: np.random.Generator) -> Perturbation:
        if len(x) == 0:
            return Perturbation(self.__class__)
        p = x.random_choice(rng)
        pos_delta, shape_delta = self.split_sampler.sample(rng)
        new = []
        for sgn in (-1, +1):
            marks = {a: m.clip(getattr(p, a) + sgn * d) for a, d, m in zip(Rectangle.PARAMETERS, shape_delta, self.split_sampler.mappings)}
            new.append(Rectangle(x=int(np.clip(p.x + sgn * pos_delta[0], 0, self.shape[0] - 1)),
                                 y=int(np.clip(p.y + sgn * pos_delta[1], 0, self.shape[1] - 1)), **marks))
        return Perturbation(self.__class__, addition=new, removal=p, data={"pos_delta": pos_delta, "shape_delta": shape_delta})

    def forward_probability(self, x: PointsSet, u: Perturbation) -> float:
        assert u.type == self.__class__
        n = len(x)
        if n == 0:
            return self.p_kernel
        return self.p_kernel * ((1 / n) * self.split_sampler.pdf(u.data["pos_delta"], u.data["shape_delta"])) / self.intensity

    def backward_probability(self, x: PointsSet, u: Perturbation) -> float:
        assert u.type == self.__class__
        n = len(x) + 1
        if n <= 1:
            return self.p_merge
        nb0 = len(x.get_potential_neighbors(u.addition[0], radius=self.radius)) + 1
        nb1 = len(x.get_potential_neighbors(u.addition[1], radius=self.radius)) + 1
        return self.p_merge * ((1 / n) * (1 / nb0) + (1 / n) * (1 / nb1))

    @property
    def p_kernel(self) -> float:
        return self.p_split


class MergeKernel(_SplitMergeBase):  # split_and_merge_kernels.py:110-178
    def sample_perturbation(self, x: PointsSet, rng: np.random.Generator) -> Perturbation:
        if len(x) <= 1:
            return Perturbation(self.__class__)
        p0 = x.random_choice(rng)
        neighbors = sorted(x.get_neighbors(p0, radius=self.radius), key=lambda q: (q.x, q.y, x._state.uid_of.get(q, 0)))
        data = {"n_neighbors": len(neighbors)}
        if not neighbors:
            return Perturbation(self.__class__, data=data)
        p1 = neighbors[int(rng.integers(0, len(neighbors)))]
        marks = {a: m.clip((getattr(p0, a) + getattr(p1, a)) / 2) for a, m in zip(Rectangle.PARAMETERS, self.split_sampler.mappings)}
        p_new = Rectangle(x=int(np.clip((p0.x + p1.x) / 2, 0, self.shape[0] - 1)), y=int(np.clip((p0.y + p1.y) / 2, 0, self.shape[0] - 1)),
                          **marks)  # the reference clips y with shape[0] too (split_and_merge_kernels.py:143)
        return Perturbation(self.__class__, addition=p_new, removal=[p0, p1], data=data)

    def forward_probability(self, x: PointsSet, u: Perturbation) -> float:
        assert u.type == self.__class__
        n = len(x)
        if n <= 1 or u.data["n_neighbors"] == 0:
            return self.p_kernel
        return self.p_kernel * ((1 / n) * (1 / u.data["n_neighbors"]))

    def backward_probability(self, x: PointsSet, u: Perturbation) -> float:
        assert u.type == self.__class__
        n = len(x) - 1
        if n == 0 or u.removal is None:
            return self.p_split
        p0, p1 = u.removal[0], u.removal[1]
        pos_delta = [(p0.x - p1.x) / 2, (p0.y - p1.y) / 2]
        shape_delta = [(getattr(p0, a) - getattr(p1, a)) / 2 for a in Rectangle.PARAMETERS]
        return self.p_split * ((1 / n) * self.split_sampler.pdf(pos_delta, shape_delta)) / self.intensity

    @property
    def p_kernel(self) -> float:
        return self.p_merge


def make_kernels(image_data: ImageWMaps, intensity: float, rng: np.random.Generator = None, use_split_merge: bool = False,
                 kernel_weights=None):
    """The eight kernels in the reference's order with their choice probabilities (make_kernels.py:50-177):
    [UniformBirth, UniformDeath, DataBirth, DataDeath, GaussianTranslation, DataTranslation, GaussianTransform,
    DataTransform], p = [1/18, 1/18, 1/9, 1/9, 1/9, 2/9, 1/9, 2/9] for the default weights; with use_split_merge two more
    (SplitKernel, MergeKernel, radius 16, shape sigmas 0.1: make_kernels.py:145-161)."""
    p = kernel_probabilities(use_split_merge=use_split_merge, weights=kernel_weights)
    kset = KernelSet(intensity=intensity, p_kernels=p[:8])
    kernels: List[Kernel] = [
        BirthKernel(kset, p[0], data_driven=False), DeathKernel(kset, p[1], data_driven=False),
        BirthKernel(kset, p[2], data_driven=True), DeathKernel(kset, p[3], data_driven=True),
        GaussianTranslationKernel(kset, p[4]), DataDrivenTranslationKernel(kset, p[5]),
        GaussianShapeTransformKernel(kset, p[6]), DataDrivenShapeTransformKernel(kset, p[7])]
    if use_split_merge:
        sampler = SplitSampler(pos_radius=16, shape_sigmas=[0.1, 0.1, 0.1], mappings=image_data.mappings)
        args = dict(p_split=p[8], p_merge=p[9], split_sampler=sampler, support_shape=tuple(image_data.shape[:2]), intensity=intensity, merge_radius=16)
        kernels += [SplitKernel(**args), MergeKernel(**args)]
    return kernels, p
