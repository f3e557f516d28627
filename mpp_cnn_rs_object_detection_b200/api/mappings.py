"""Mark value <-> class-bin mapping: models/shape_net/mappings.py:10-157 (ValueMapping only; 32 bins per mark)."""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class ValueMapping:
    n_classes: int
    v_min: float
    v_max: float
    is_cyclic: bool = False

    def __post_init__(self):
        self.feature_mapping = np.linspace(self.v_min, self.v_max, num=self.n_classes + 1)[:-1]  # lower edges (:17)

    def get_step(self) -> float:
        return float(np.mean(np.diff(self.feature_mapping)))

    def get_range(self) -> float:
        return self.v_max - self.v_min

    @property
    def range(self) -> float:
        return self.v_max - self.v_min

    def clip(self, value: float) -> float:  # mappings.py:39-43
        if not self.is_cyclic:
            return float(np.clip(value, self.v_min, self.v_max))
        return ((value - self.v_min) % self.range) + self.v_min

    def value_to_class(self, value):
        """max{c : value >= edge_c} (mappings.py:45-61); raises ValueError below v_min like the reference's np.max([])."""
        if not (np.all(self.v_min <= value) and np.all(value <= self.v_max)):
            logging.warning(f"feature value {value} out of range [{self.v_min:.2f},{self.v_max:.2f}]")
        c = np.searchsorted(self.feature_mapping, value, side="right") - 1
        if np.any(c < 0):
            raise ValueError(f"value {value} below v_min={self.v_min}")
        return c if isinstance(value, np.ndarray) else int(c)

    def class_to_value(self, class_id):  # mappings.py:63-74
        if hasattr(class_id, "cpu"):
            class_id = class_id.cpu().detach().numpy()
        return self.feature_mapping[class_id]


def default_mappings() -> List[ValueMapping]:
    """size [0,32), ratio [0,1), angle [0,pi) cyclic (models/shape_net/shape_net_model.py:80-85)."""
    return [ValueMapping(32, 0, 32), ValueMapping(32, 0, 1), ValueMapping(32, 0, np.pi, is_cyclic=True)]


def output_vector_to_value(output_vector, mappings: List[ValueMapping]):
    """argmax over the class axis (axis 1) of (B,C) or (B,C,H,W) arrays -> mark values (mappings.py:145-157)."""
    out = []
    for arr, mapping in zip(output_vector, mappings):
        if len(arr.shape) not in (2, 4):
            raise ValueError
        out.append(mapping.class_to_value(np.argmax(arr, axis=1)))
    return out
