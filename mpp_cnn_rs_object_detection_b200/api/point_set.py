"""PointsSet: the interacting points set's spatial index (models/mpp/point_set/point_set.py:45-188), backed by the
device cell lists (32-px cells, MPP_CELL_SIZE) instead of a Python list of sets."""
from __future__ import annotations

from typing import List, Set, Tuple

import numpy as np

from .. import _lib
from .device_state import MAX_DEVICE_RADIUS, DeviceState, build_layout
from .shapes import Point

MIN_SPATIAL_RES = 32  # point_set.py:9


class PointsSet:
    def __init__(self, support_shape: Tuple[int, int], maximum_interaction_radius: int, _state: DeviceState = None):
        self._spatial_resolution = max(maximum_interaction_radius, MIN_SPATIAL_RES)  # point_set.py:58
        if self._spatial_resolution > MAX_DEVICE_RADIUS:
            raise NotImplementedError(f"maximum_interaction_radius {maximum_interaction_radius} > {MAX_DEVICE_RADIUS} px: the device "
                                      f"grid has fixed 32-px cells")
        self.support_shape = support_shape
        self._n_x = int(np.ceil(support_shape[0] / self._spatial_resolution))  # point_set.py:60-61
        self._n_y = int(np.ceil(support_shape[1] / self._spatial_resolution))
        self._state = _state if _state is not None else DeviceState(support_shape, build_layout([], []))

    # ---- container protocol
    def __iter__(self):
        return iter(self._state.objects())

    def __len__(self):
        return len(self._state)

    def __contains__(self, u: Point):
        self._cell_index(u)
        return u in self._state

    def __copy__(self):
        return PointsSet(self.support_shape, self._spatial_resolution, _state=self._state.copy())

    def copy(self) -> "PointsSet":
        return self.__copy__()

    def _cell_index(self, u: Point) -> int:
        i = int(u.x) // self._spatial_resolution
        j = int(u.y) // self._spatial_resolution
        assert i < self._n_x and j < self._n_y  # Point out of bounds (point_set.py:99)
        return j + i * self._n_y

    @property
    def _local_sets(self) -> List[Set[Point]]:
        """Host view of the device cells (the reference's tests read this private list)."""
        sets: List[Set[Point]] = [set() for _ in range(self._n_x * self._n_y)]
        for h, u in self._state.obj_of.items():
            sets[h >> 5].add(u)
        return sets

    def _find_local_point_set(self, u: Point) -> Set[Point]:
        return self._local_sets[self._cell_index(u)]

    def get_subsets(self):
        return self._local_sets

    def add(self, u: Point):
        self._cell_index(u)
        self._state.add(u)

    def remove(self, u: Point):
        self._state.remove(u)

    # ---- neighbourhoods (device query)
    def _query(self, u: Point, radius: float, exclude_itself: bool, euclidean: bool) -> Set[Point]:
        st = self._state
        excl = st.handle_of.get(u, _lib.NO_OBJECT) if exclude_itself else _lib.NO_OBJECT
        handles = st.engine.query_neighbors(int(u.x), int(u.y), float(radius), euclidean=euclidean, exclude=excl)
        return {st.obj_of[int(h)] for h in handles}

    def get_potential_neighbors(self, u: Point, radius: float, exclude_itself=True, suppress_warnings=False) -> Set[Point]:
        """All objects of the cells within ceil(radius / cell) offsets of u's cell (point_set.py:111-145)."""
        if int(np.ceil(radius / self._spatial_resolution)) > 1 and not suppress_warnings:
            print("[PointsSet] getting neighbors further than the specified maximum interaction radius")
        return self._query(u, radius, exclude_itself, euclidean=False)

    def get_neighbors(self, u: Point, radius: float, exclude_itself=True, suppress_warnings=False):
        """... filtered by euclidean centre distance <= radius (point_set.py:147-149)."""
        if int(np.ceil(radius / self._spatial_resolution)) > 1 and not suppress_warnings:
            print("[PointsSet] getting neighbors further than the specified maximum interaction radius")
        return self._query(u, radius, exclude_itself, euclidean=True)

    def _get_i_th_point(self, i: int):
        objs = self._state.objects()
        if not 0 <= i < len(objs):
            raise IndexError
        return objs[i]

    def random_choice(self, rng: np.random.Generator):
        """Uniform draw: rng.integers(0, n) then the i-th object in cell-major order (point_set.py:176-185)."""
        n = len(self)
        return self._get_i_th_point(int(rng.integers(0, n)))
