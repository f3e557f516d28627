"""Proposal kernels of the RJMCMC sampler: models/mpp/rjmcmc_sampler/kernels/{base_kernels,transform_kernels,
make_kernels}.py.  The classes keep the reference's interface (sample_perturbation / forward_probability /
backward_probability); draws and probabilities are computed on the device (mpp_sample_proposals, mpp_proposal_probs)
with Philox randomness seeded from the numpy Generator the caller passes."""
from __future__ import annotations

from abc import abstractmethod
from copy import copy
from typing import List, Optional

import numpy as np

from .. import _lib
from ..engine import kernel_probabilities
from .custom_types import ImageWMaps, Perturbation
from .point_set import PointsSet
from .shapes import Point, Rectangle

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


def make_kernels(image_data: ImageWMaps, intensity: float, rng: np.random.Generator = None, use_split_merge: bool = False,
                 kernel_weights=None):
    """The eight kernels in the reference's order with their choice probabilities (make_kernels.py:50-177):
    [UniformBirth, UniformDeath, DataBirth, DataDeath, GaussianTranslation, DataTranslation, GaussianTransform,
    DataTransform], p = [1/18, 1/18, 1/9, 1/9, 1/9, 2/9, 1/9, 2/9] for the default weights."""
    if use_split_merge:
        raise NotImplementedError("split / merge kernels (split_and_merge_kernels.py) are not built yet")
    p = kernel_probabilities(weights=kernel_weights)
    kset = KernelSet(intensity=intensity, p_kernels=p)
    kernels: List[Kernel] = [
        BirthKernel(kset, p[0], data_driven=False), DeathKernel(kset, p[1], data_driven=False),
        BirthKernel(kset, p[2], data_driven=True), DeathKernel(kset, p[3], data_driven=True),
        GaussianTranslationKernel(kset, p[4]), DataDrivenTranslationKernel(kset, p[5]),
        GaussianShapeTransformKernel(kset, p[6]), DataDrivenShapeTransformKernel(kset, p[7])]
    return kernels, p
