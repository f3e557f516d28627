"""EnergyGraph: per-object energy vectors, subset energies and the Delta-energy of a perturbation
(models/mpp/point_set/energy_graph.py:20-291), evaluated by the CUDA kernels (mpp_energy_vectors, mpp_delta_batch).

The reference keeps Python lists of UnitEnergy / PairEnergy instances per object; here the graph is implicit in the
device cell lists, and `ue_per_point` / `pe_per_point` are read-only views built on demand from device queries."""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional, Set, Union

from .custom_types import EnergyCombinationModel, Perturbation
from .device_state import DeviceState, build_layout
from .energies import PairEnergy, PairEnergyConstructor, UnitEnergy, UnitEnergyConstructor
from .point_set import PointsSet
from .shapes import Point


def _as_list(x) -> List[Point]:
    if x is None:
        return []
    return list(x) if type(x) is list else [x]


class _PerPointView:
    """dict-like view: `u in view`, `view[u]` -> list of UnitEnergy / PairEnergy, `.keys()`, `len(view)`."""

    def __init__(self, graph: "EnergyGraph", pair: bool):
        self._ref, self._pair = weakref.ref(graph), pair  # weak: no reference cycle, device state is released promptly

    @property
    def _g(self) -> "EnergyGraph":
        return self._ref()

    def __contains__(self, u):
        return u in self._g._members

    def __len__(self):
        return len(self._g._members)

    def keys(self):
        return list(self._g._members)

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, u):
        if u not in self._g._members:
            raise KeyError(u)
        return self._g._pair_list(u) if self._pair else self._g._unit_list(u)


class EnergyGraph:
    def __init__(self, unit_energies_constructors: List[UnitEnergyConstructor],
                 pair_energies_constructors: List[PairEnergyConstructor], _state: DeviceState = None):
        self.ue_constructors = unit_energies_constructors
        self.pe_constructors = pair_energies_constructors
        self._layout = build_layout(unit_energies_constructors, pair_energies_constructors)  # asserts unique names (:37-44)
        self.max_interaction_dist = max([p.max_dist for p in pair_energies_constructors], default=1)  # :27-30
        self.energies_keys = [c.name for c in unit_energies_constructors] + [c.name for c in pair_energies_constructors]
        self._state: Optional[DeviceState] = _state
        self._members: Dict[Point, None] = {} if _state is None else dict.fromkeys(_state.handle_of)
        self._pair_cache: Dict[tuple, PairEnergy] = {}
        self.ue_per_point = _PerPointView(self, pair=False)
        self.pe_per_point = _PerPointView(self, pair=True)

    # ---- binding to the points set that owns the device context
    def _bind(self, points_set: Union[PointsSet, "DeviceState"]):
        st = points_set._state if isinstance(points_set, PointsSet) else points_set
        if self._state is st and st.layout is self._layout:
            return st
        self._state = st
        if st.layout is not self._layout:
            st.rebind_layout(self._layout)
        return st

    # ---- structure
    def add_point(self, u: Point, points_set: PointsSet):
        """energy_graph.py:46-77.  The object must already be in `points_set` (the reference's calling convention:
        `ps.add(u); eg.add_point(u, ps)`); the pair structure is implicit in the device cell lists."""
        st = self._bind(points_set)
        if u not in st:
            st.add(u)
        self._members[u] = None

    def remove_point(self, u: Point):
        """energy_graph.py:79-90."""
        if u not in self._members:
            raise KeyError(u)
        del self._members[u]
        self._pair_cache = {k: v for k, v in self._pair_cache.items() if v.point_1 is not u and v.point_2 is not u}

    def __copy__(self):
        new = EnergyGraph(self.ue_constructors, self.pe_constructors)
        new._layout = self._layout
        new._state = self._state
        new._members = dict(self._members)
        return new

    def copy(self):
        return self.__copy__()

    def get_interacting_points(self, u: Point) -> Set[Point]:
        return {pe.get_other_point(u) for pe in self.pe_per_point[u]}

    def check_integrity(self):
        st = self._state
        if st is None:
            return
        for u in self._members:
            assert u in st.handle_of, f"{u} is in the energy graph but not in the points set"

    # ---- views
    def _vector_of(self, u: Point) -> Dict[str, float]:
        st = self._state
        vec, _, _, _ = st.engine.energy_vectors(st.handles([u]))
        return {name: float(vec[0, col]) for name, col in self._layout.columns}

    def _unit_list(self, u: Point) -> List[UnitEnergy]:
        return [UnitEnergy(point=u, constructor=c, _fetch=lambda ue: self._vector_of(ue.point)[ue.constructor.name])
                for c in self.ue_constructors]

    def _pair_value(self, pe: PairEnergy) -> float:
        st = self._state
        out = st.engine.pair_values(st.handles([pe.point_1]), st.handles([pe.point_2]))[0]
        col = 0 if (self._layout.spec.setup == "toy" or type(pe.constructor).__name__ == "RectangleOverlapEnergy") else 1
        return float(out[col])

    def _pair_list(self, u: Point) -> List[PairEnergy]:
        st = self._state
        out: List[PairEnergy] = []
        for c in self.pe_constructors:
            handles = st.engine.query_neighbors(int(u.x), int(u.y), float(c.max_dist), euclidean=True, exclude=st.handle_of[u])
            for h in sorted(int(v) for v in handles):
                v = st.obj_of[h]
                if v not in self._members:
                    continue
                key = (id(c),) + tuple(sorted((st.uid_of[u], st.uid_of[v])))
                pe = self._pair_cache.get(key)
                if pe is None:
                    pe = PairEnergy(point_1=u, point_2=v, constructor=c, _fetch=self._pair_value)
                    self._pair_cache[key] = pe
                out.append(pe)
        return out

    # ---- energies
    def total_energy(self, points_set: PointsSet, force_update=False) -> float:
        """Raw sum over every object and term; the combinator is ignored (energy_graph.py:105-106)."""
        return self.compute_subset(subset=points_set)

    def compute_subset(self, subset: Union[Set[Point], List[Point], PointsSet], force_update: bool = False,
                       energy_combinator: EnergyCombinationModel = None, return_vector=False):
        """energy_graph.py:108-137: per-object vectors (pair kinds reduced over partners, absent -> 0), then the
        combinator (or the raw sum)."""
        st = self._state
        objs = list(subset)
        if st is None or len(objs) == 0:
            vectors = {k: [] for k in self.energies_keys}
            if return_vector:
                return vectors
            return 0.0 if energy_combinator is None else energy_combinator.compute(vectors)
        fused = energy_combinator is None or hasattr(energy_combinator, "device_params")
        st.use_combinator(energy_combinator if fused else None)
        vec, _, raw_total, comb_total = st.engine.energy_vectors(st.handles(objs))
        if return_vector:
            return {name: [float(v) for v in vec[:, col]] for name, col in self._layout.columns}
        if energy_combinator is None:
            return raw_total
        if fused:
            return comb_total
        return energy_combinator.compute({name: [float(v) for v in vec[:, col]] for name, col in self._layout.columns})

    def compute_delta(self, points_set: PointsSet, pert: Perturbation, energy_combinator: EnergyCombinationModel = None):
        """Energy difference new - old of applying `pert` (energy_graph.py:139-225).  Nothing is modified."""
        st = self._bind(points_set)
        removed, added = _as_list(pert.removal), _as_list(pert.addition)
        for r in removed:
            if r not in st.handle_of:
                raise KeyError(r)
        if not removed and not added:
            return 0.0
        fused = energy_combinator is None or hasattr(energy_combinator, "device_params")
        if not fused:
            return self._delta_plugin(st, removed, added, energy_combinator)
        st.use_combinator(energy_combinator)
        if len(removed) <= 1 and len(added) <= 1:
            rec = st.proposal_record(removed[0] if removed else None, added[0] if added else None)
            return float(st.engine.delta_batch(rec)[0])
        return self._delta_multi(st, removed, added)

    def _delta_multi(self, st: DeviceState, removed, added) -> float:
        """Perturbations with several removals / additions (split & merge kernels): the energy is a function of the state, so
        the difference telescopes into single-object Delta-energies evaluated on the device while the state is stepped through
        the intermediate configurations; the state is restored afterwards (mutate-and-revert, like energy_graph.py:191-223)."""
        total, undo = 0.0, []
        try:
            for q in removed:
                total += float(st.engine.delta_batch(st.proposal_record(q, None))[0])
                st.remove(q)
                undo.append(("add", q))
            for a in added:
                total += float(st.engine.delta_batch(st.proposal_record(None, a))[0])
                st.add(a)
                undo.append(("remove", a))
        finally:
            for op, u in reversed(undo):
                (st.add if op == "add" else st.remove)(u)
        return total

    def _delta_plugin(self, st: DeviceState, removed, added, combinator) -> float:
        """Plug-in (Python) combinator: the reference's own recipe on device-computed vectors -- 3x3-cell neighbourhoods of
        the changed objects before and after (energy_graph.py:156-225)."""
        r = self.max_interaction_dist
        ps = PointsSet(st.support_shape, 32, _state=st)
        conn: Dict[Point, None] = {}
        for a in added:
            for p in ps.get_potential_neighbors(a, r):
                if p not in removed:
                    conn[p] = None
        for q in removed:
            for p in ps.get_potential_neighbors(q, r):
                if p not in added:
                    conn[p] = None
        unchanged = list(conn)
        e0 = self.compute_subset(unchanged + [q for q in removed if q not in conn], energy_combinator=combinator)
        for q in removed:
            st.remove(q)
        for a in added:
            st.add(a)
        try:
            e1 = self.compute_subset(unchanged + [a for a in added if a not in conn], energy_combinator=combinator)
        finally:
            for a in added:
                st.remove(a)
            for q in removed:
                st.add(q)
        return e1 - e0
