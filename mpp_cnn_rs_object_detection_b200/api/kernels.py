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
class SplitSampler:  # split_and_merge_kernels.py:14-36 (parameters only: the draws and the density run on the device)
    def __init__(self, pos_radius: float, shape_sigmas: List[float], mappings):
        self.pos_radius, self.shape_sigmas, self.mappings = float(pos_radius), [float(v) for v in shape_sigmas], mappings
        self.scaled_shaped_sigmas = [s * m.range for m, s in zip(mappings, shape_sigmas)]
        self.n_params = len(Rectangle.PARAMETERS)


class _SplitMergeBase(Kernel):
    """The optional two-object moves (use_split_merge=False in both shipped configurations): device kernel ids 8 (split) and 9
    (merge).  Draws (mpp_sample_split_merge: Philox seeded from the numpy Generator) and forward / backward densities
    (mpp_split_merge_probs) run on the device; the Delta-energy of the two-object perturbation telescopes into single-object
    Delta-energies on the device (EnergyGraph._delta_multi).  They run through the step-by-step RJMCMC loop only."""
    KIND = -1

    def __init__(self, p_split: float, p_merge: float, split_sampler: SplitSampler, support_shape, intensity: float, merge_radius: float):
        self.p_split, self.p_merge, self.split_sampler = p_split, p_merge, split_sampler
        self.shape, self.intensity, self.radius = support_shape, intensity, merge_radius
        assert self.radius == self.split_sampler.pos_radius

    def _engine(self, x: PointsSet):
        st = x._state
        st.use_kernels(self.intensity)
        return st.engine

    def _record(self, u: Perturbation) -> np.ndarray:
        """The C-ABI form of a Perturbation of this kernel (what the densities need: the additions' positions, the shape deltas,
        whether a second removal exists, the neighbour count)."""
        rec = np.zeros(1, dtype=_lib.SPLIT_MERGE_DTYPE)
        rec["kind"] = self.KIND
        rec["rem_uid"][0] = (0xFFFFFFFF, 0xFFFFFFFF)
        data = u.data or {}
        rem = [] if u.removal is None else (list(u.removal) if isinstance(u.removal, (list, tuple)) else [u.removal])
        add = [] if u.addition is None else (list(u.addition) if isinstance(u.addition, (list, tuple)) else [u.addition])
        for k, q in enumerate(rem[:2]):
            rec["rem_uid"][0][k] = 0  # present (the densities only look at presence)
            rec["rem_x"][0][k], rec["rem_y"][0][k] = q.x, q.y
        rec["n_add"] = len(add)
        for k, a in enumerate(add[:2]):
            rec["add_x"][0][k], rec["add_y"][0][k] = a.x, a.y
            rec["add_size"][0][k], rec["add_ratio"][0][k], rec["add_angle"][0][k] = a.size, a.ratio, a.angle
        rec["n_neighbors"] = int(data.get("n_neighbors", -1))
        if self.KIND == 8 and "shape_delta" in data:
            rec["pos_delta"][0][:] = np.asarray(data["pos_delta"], dtype=np.float64)
            rec["shape_delta"][0][:] = np.asarray(data["shape_delta"], dtype=np.float64)
        if self.KIND == 9 and len(rem) == 2:  # split_and_merge_kernels.py:170-172
            rec["pos_delta"][0][:] = [(rem[0].x - rem[1].x) / 2, (rem[0].y - rem[1].y) / 2]
            rec["shape_delta"][0][:] = [(getattr(rem[0], a) - getattr(rem[1], a)) / 2 for a in Rectangle.PARAMETERS]
        return rec

    def _probs(self, x: PointsSet, u: Perturbation):
        assert u.type == self.__class__
        return self._engine(x).split_merge_probs(self._record(u), self.p_split, self.p_merge, self.radius, self.split_sampler.shape_sigmas)

    def forward_probability(self, x: PointsSet, u: Perturbation) -> float:
        return self._probs(x, u)[0]

    def backward_probability(self, x: PointsSet, u: Perturbation) -> float:
        return self._probs(x, u)[1]

    def _draw(self, x: PointsSet, rng: np.random.Generator) -> np.ndarray:
        return self._engine(x).sample_split_merge(self.KIND, self.radius, self.split_sampler.shape_sigmas, seed=int(rng.integers(0, 2 ** 62)))[0]


class SplitKernel(_SplitMergeBase):  # split_and_merge_kernels.py:39-107
    KIND = 8

    def sample_perturbation(self, x: PointsSet, rng: np.random.Generator) -> Perturbation:
        if len(x) == 0:
            return Perturbation(self.__class__)
        r = self._draw(x, rng)
        p = x._state.by_uid[int(r["rem_uid"][0])]
        new = [Rectangle(x=int(r["add_x"][k]), y=int(r["add_y"][k]), size=float(r["add_size"][k]), ratio=float(r["add_ratio"][k]),
                         angle=float(r["add_angle"][k])) for k in range(2)]
        return Perturbation(self.__class__, addition=new, removal=p, data={"pos_delta": np.array(r["pos_delta"]), "shape_delta": np.array(r["shape_delta"])})

    @property
    def p_kernel(self) -> float:
        return self.p_split


class MergeKernel(_SplitMergeBase):  # split_and_merge_kernels.py:110-178
    KIND = 9

    def sample_perturbation(self, x: PointsSet, rng: np.random.Generator) -> Perturbation:
        if len(x) <= 1:
            return Perturbation(self.__class__)
        r = self._draw(x, rng)
        data = {"n_neighbors": int(r["n_neighbors"])}
        if int(r["n_add"]) == 0:
            return Perturbation(self.__class__, data=data)
        p0, p1 = x._state.by_uid[int(r["rem_uid"][0])], x._state.by_uid[int(r["rem_uid"][1])]
        p_new = Rectangle(x=int(r["add_x"][0]), y=int(r["add_y"][0]), size=float(r["add_size"][0]), ratio=float(r["add_ratio"][0]),
                          angle=float(r["add_angle"][0]))
        return Perturbation(self.__class__, addition=p_new, removal=[p0, p1], data=data)

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
