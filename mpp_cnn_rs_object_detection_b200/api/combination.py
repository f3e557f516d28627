"""Inference-side energy combinators: models/mpp/energies/combination/{hierarchical,logistic}.py.

`compute(vectors)` keeps the reference signature (dict name -> list of per-object values -> float) and is evaluated on
the device (mpp_combine).  EPointsSet / EnergyGraph recognise these classes and fuse the combinator into the
Delta-energy kernels instead of calling `compute`."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from ..engine import LEGACY_NAMES, NOCALIB_NAMES, ModelSpec, combine_on_device
from .custom_types import ConfigurationEnergyVector, EnergyCombinationModel


def _matrix(vectors: ConfigurationEnergyVector, names: List[str]) -> np.ndarray:
    cols = [np.asarray(vectors[k], dtype=np.float64).reshape(-1) if k in vectors else None for k in names]
    n = max([len(c) for c in cols if c is not None] + [0])
    return np.stack([c if c is not None else np.zeros(n) for c in cols], axis=1) if n else np.zeros((0, len(names)))


@dataclass
class HierarchicalEnergyCombinator(EnergyCombinationModel):  # hierarchical.py:13-32
    weights_data: np.ndarray
    weights_prior: np.ndarray
    data_prior_weights: np.ndarray
    detection_threshold: float
    bias: float = 0.0

    def device_params(self, names: List[str]):
        if list(names[:5]) != LEGACY_NAMES:
            raise ValueError("HierarchicalEnergyCombinator needs the legacy term names (hierarchical.py:22-29)")
        w = [float(v) for v in self.weights_data] + [float(v) for v in self.weights_prior] + \
            [float(v) for v in self.data_prior_weights] + [0.0]
        return "hierarchical", w, float(self.bias), float(self.detection_threshold)

    def compute(self, vectors: ConfigurationEnergyVector) -> float:
        kind, w, b, t = self.device_params(LEGACY_NAMES)
        spec = ModelSpec(setup="legacy", combinator=kind, comb_w=w, comb_bias=b, comb_threshold=t)
        return combine_on_device(spec, _matrix(vectors, LEGACY_NAMES))[1]


@dataclass
class ManualHierarchicalEnergyCombinator(EnergyCombinationModel):  # hierarchical.py:35-48
    weights_dict: Dict[str, float]
    indicator_energy: str
    detection_threshold: float = 0.0

    def device_params(self, names: List[str]):
        if names[0] != self.indicator_energy:
            raise ValueError("the indicator energy must be the position term (first term of the setup)")
        missing = [k for k in self.weights_dict if k not in names]
        if missing:
            raise KeyError(missing[0])
        w = [float(self.weights_dict.get(k, 0.0)) for k in names] + [0.0] * (8 - len(names))
        return "manual", w, 0.0, float(self.detection_threshold)

    def compute(self, vectors: ConfigurationEnergyVector) -> float:
        names = list(vectors.keys())
        names = LEGACY_NAMES if set(names) <= set(LEGACY_NAMES) else [k for k in NOCALIB_NAMES if k in names]
        kind, w, b, t = self.device_params(list(names))
        spec = ModelSpec(setup="legacy" if names == LEGACY_NAMES else "nocalib", ratio_prior=len(names) == 8, combinator=kind,
                         comb_w=w, comb_bias=b, comb_threshold=t)
        return combine_on_device(spec, _matrix(vectors, list(names)))[1]


@dataclass
class LogisticEnergyCombinator(EnergyCombinationModel):  # logistic.py:14-26
    weights: np.ndarray
    bias: float
    energy_names: List[str]

    def device_params(self, names: List[str]):
        if list(names) != list(self.energy_names):
            raise ValueError(f"LogisticEnergyCombinator was trained on {self.energy_names}, the setup provides {names}")
        w = [float(v) for v in np.asarray(self.weights).reshape(-1)] + [0.0] * (8 - len(names))
        return "logistic", w, float(self.bias), 0.0

    def compute(self, vectors: ConfigurationEnergyVector) -> float:
        names = list(self.energy_names)
        m = _matrix(vectors, names)
        if len(m) == 0:
            return 0.0
        kind, w, b, t = self.device_params(names)
        setup = "legacy" if names == LEGACY_NAMES else "nocalib"
        spec = ModelSpec(setup=setup, ratio_prior=len(names) == 8, combinator=kind, comb_w=w, comb_bias=b, comb_threshold=t)
        return combine_on_device(spec, m)[1]


@dataclass
class MLPEnergyCombinator(EnergyCombinationModel):  # combination/mlp.py:13-27
    """A torch module over the per-object energy vectors: sum(2*model(v) - 1) (or sum(model(v)) when raw_energy).  It is a
    plug-in combinator: the vectors are computed by the CUDA kernels, the module is evaluated by PyTorch on the same device,
    and Delta-energies go through the reference's before / after recipe (EnergyGraph._delta_plugin) instead of the fused
    kernels -- usable with the step-by-step chain, not with the parallel sampler."""
    model: "object"
    energy_names: List[str]
    raw_energy: bool = False

    def compute(self, vectors: ConfigurationEnergyVector) -> float:
        import torch
        if len(vectors[self.energy_names[0]]) == 0:
            return 0.0
        dev = torch.device("cuda", torch.cuda.current_device())
        x = torch.as_tensor(np.stack([np.asarray(vectors[k], dtype=np.float32) for k in self.energy_names], axis=-1), device=dev)
        with torch.no_grad():
            y = self.model.to(dev).forward(x)
            return float((2 * y - 1).sum().item()) if not self.raw_energy else float(y.sum().item())
