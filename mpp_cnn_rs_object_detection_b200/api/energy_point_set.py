"""EPointsSet: points set + energy graph facade (models/mpp/point_set/energy_point_set.py:18-166) over one device
context.  Papangelou intensity, Delta-energies and apply/unapply of perturbations."""
from __future__ import annotations

import logging
from typing import Iterable, List, Tuple

import numpy as np

from .custom_types import EnergyCombinationModel, Perturbation
from .device_state import DeviceState, build_layout
from .energies import PairEnergyConstructor, UnitEnergyConstructor
from .energy_graph import EnergyGraph, _as_list
from .point_set import PointsSet
from .shapes import Point


class EPointsSet:

    def __init__(self, points: Iterable[Point], support_shape: Tuple[int, int],
                 unit_energies_constructors: List[UnitEnergyConstructor],
                 pair_energies_constructors: List[PairEnergyConstructor], debug=False, precision: str = "fp32",
                 _state: DeviceState = None, reuse_device_maps: bool = True, _device_maps=None):
        self.debug = debug
        if len(pair_energies_constructors) > 0:  # energy_point_set.py:33-36
            self.maximum_interaction_radius = np.max([pec.max_dist for pec in pair_energies_constructors])
        else:
            self.maximum_interaction_radius = 0
        if _state is None:
            layout = build_layout(unit_energies_constructors, pair_energies_constructors)  # asserts unique names (:25-29)
            _state = DeviceState(support_shape, layout, precision=precision, reuse_maps=reuse_device_maps, maps=_device_maps)
            _state.add_many(list(points))
        self._state = _state
        self.points: PointsSet = PointsSet(support_shape=support_shape, maximum_interaction_radius=self.maximum_interaction_radius,
                                           _state=_state)
        self.energy_graph = EnergyGraph(unit_energies_constructors, pair_energies_constructors, _state=_state)
        self.energy_graph._layout = _state.layout

    def __copy__(self):
        st = self._state.copy()
        return EPointsSet([], self.points.support_shape, self.energy_graph.ue_constructors, self.energy_graph.pe_constructors,
                          debug=self.debug, _state=st)

    def copy(self):
        return self.__copy__()

    def __len__(self):
        return len(self.points)

    def __contains__(self, u: Point):
        return u in self.points

    def __iter__(self):
        return self.points.__iter__()

    def add(self, u: Point):
        self.points.add(u)
        self.energy_graph._members[u] = None

    def remove(self, u: Point):
        self.points.remove(u)
        self.energy_graph._members.pop(u, None)

    def total_energy(self, force_update=False) -> float:
        return self.energy_graph.total_energy(points_set=self.points)

    def energy_delta(self, p: Perturbation, energy_combinator: EnergyCombinationModel = None):
        try:
            return self.energy_graph.compute_delta(self.points, pert=p, energy_combinator=energy_combinator)
        except KeyError as e:  # energy_point_set.py:88-100
            self.energy_graph.check_integrity()
            for r in _as_list(p.removal):
                if r not in self.points:
                    logging.error(f"point to remove {r} not in points set {self}")
            raise e

    def papangelou(self, u: Point, energy_combinator: EnergyCombinationModel = None, remove_u_from_point_set: bool = False,
                   return_energy_delta: bool = False):
        """exp(-Delta E of adding u); for a member of the set (remove_u_from_point_set=True): exp(+Delta E of removing it)
        (energy_point_set.py:102-116)."""
        from .kernels import BirthKernel
        if u in self.points:
            if not remove_u_from_point_set:
                print(f"point {u} is already in current set, cannot compute papangelou conditional intensity")
                raise ValueError
            delta = -self.energy_delta(Perturbation(type=BirthKernel, removal=u, addition=None), energy_combinator=energy_combinator)
        else:
            delta = self.energy_delta(Perturbation(type=BirthKernel, removal=None, addition=u), energy_combinator=energy_combinator)
        if return_energy_delta:
            return delta
        return np.exp(-delta)

    def papangelou_all(self, energy_combinator: EnergyCombinationModel = None, return_energy_delta: bool = False):
        """Batched papangelou(u, remove_u_from_point_set=True) of every member, in iteration order: one kernel launch
        instead of len(self) Python calls (the scoring loop of models/mpp/mpp_model.py:296-304)."""
        st = self._state
        objs = st.objects()
        if not objs:
            return objs, np.zeros(0)
        st.use_combinator(energy_combinator)
        rec = np.concatenate([st.proposal_record(u, None) for u in objs])
        delta = -st.engine.delta_batch(rec)
        return objs, (delta if return_energy_delta else np.exp(-delta))

    def apply_perturbation(self, p: Perturbation, inplace=False) -> "EPointsSet":
        new_x = self if inplace else self.copy()  # energy_point_set.py:118-154: removals first, then additions
        for r in _as_list(p.removal):
            new_x.remove(r)
        for a in _as_list(p.addition):
            new_x.add(a)
        return new_x

    def unapply_perturbation(self, p: Perturbation) -> "EPointsSet":
        new_x = self.copy()
        for a in _as_list(p.addition):
            new_x.remove(a)
        for r in _as_list(p.removal):
            new_x.add(r)
        return new_x
