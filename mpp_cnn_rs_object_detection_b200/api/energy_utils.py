"""Batches of energy vectors for the weight-learning loops: models/mpp/energies/energy_utils.py:38-82.

A configuration's (n_objects, n_terms) matrix is one launch of mpp_energy_vectors; the reference maps configurations over a
process pool (`multiprocess=True`), here they are evaluated one after the other on the device (the flag is accepted and
ignored) and the device buffers of the image's maps are shared by all of them."""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np

from .custom_types import ImageWMaps
from .energies import PairEnergyConstructor, UnitEnergyConstructor
from .energy_point_set import EPointsSet
from .shapes import Rectangle


def names_from_energies(energies: Sequence[Union[PairEnergyConstructor, UnitEnergyConstructor]]) -> List[str]:
    """energy_utils.py:38-43."""
    return [e.name for e in energies]


def compute_energy_vector(points: Union[List[Rectangle], EPointsSet], unit_energies: List[UnitEnergyConstructor],
                          pair_energies: List[PairEnergyConstructor], support_shape, energy_names: List[str], return_names: bool = False):
    """(n_objects, len(energy_names)) matrix of per-object term values, pair kinds reduced over partners (energy_utils.py:48-66)."""
    if not isinstance(points, EPointsSet):
        points = EPointsSet(points=list(points), support_shape=support_shape, unit_energies_constructors=unit_energies,
                            pair_energies_constructors=pair_energies)
    per_type = points.energy_graph.compute_subset(subset=points, return_vector=True)
    vector = np.array([per_type[k] for k in energy_names], dtype=float).T
    if len(points) == 0:
        vector = vector.reshape(0, len(energy_names))
    if return_names:
        return vector, list(per_type.keys())
    return vector


def compute_many_energy_vectors(configurations: List[List[Rectangle]], image_config: ImageWMaps, ue: List[UnitEnergyConstructor],
                                pe: List[PairEnergyConstructor], energy_names: List[str], multiprocess: bool = True) -> np.ndarray:
    """Vectors of several configurations of one image, concatenated along the object axis (energy_utils.py:69-82)."""
    shape = tuple(image_config.detection_map.shape[:2])
    out = [compute_energy_vector(cfg, unit_energies=ue, pair_energies=pe, support_shape=shape, energy_names=energy_names) for cfg in configurations]
    if not out:
        return np.zeros((0, len(energy_names)))
    return np.concatenate(out, axis=0)
