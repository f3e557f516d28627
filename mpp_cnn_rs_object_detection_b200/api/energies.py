"""Energy-term constructors: models/mpp/energies/{base,data,prior}_energies.py.

In the reference a constructor *computes* its term in Python.  Here a constructor is a declarative description that
EPointsSet / EnergyGraph translate into the device energy model (mpp_set_model, include/mpp_b200.h); the values are
computed by the CUDA kernels.  Arbitrary Python subclasses of UnitEnergyConstructor / PairEnergyConstructor cannot run
on the device and are rejected with a clear error when a points set is built from them; the toy terms of the
reference's own tests are available as device terms (ConstantUnitEnergy, DistanceIndicatorPairEnergy)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

from .mappings import ValueMapping
from .shapes import Point


@dataclass(eq=False)
class UnitEnergyConstructor:  # base_energies.py:9-26
    name: str

    def instanciate(self, u: Point):
        return UnitEnergy(point=u, constructor=self)

    def __hash__(self):
        return id(self)


@dataclass(eq=False)
class PairEnergyConstructor:  # base_energies.py:40-67
    name: str
    max_dist: float

    def instanciate(self, u_1: Point, u_2: Point):
        return PairEnergy(point_1=u_1, point_2=u_2, constructor=self)

    def __hash__(self):
        return id(self)


@dataclass(eq=False)
class UnitEnergy:  # base_energies.py:29-36 -- the value is fetched from the device by the owning graph
    point: Point
    constructor: UnitEnergyConstructor
    value: float = None
    _fetch: Any = field(default=None, repr=False)

    def compute(self, lazy=True) -> float:
        if self.value is None or not lazy:
            self.value = self._fetch(self)
        return self.value


@dataclass(eq=False)
class PairEnergy:  # base_energies.py:70-83
    point_1: Point
    point_2: Point
    constructor: PairEnergyConstructor
    value: float = None
    _fetch: Any = field(default=None, repr=False)

    def compute(self, lazy=True) -> float:
        if self.value is None or not lazy:
            self.value = self._fetch(self)
        return self.value

    def get_other_point(self, current: Point):
        return self.point_1 if self.point_2 is current else self.point_2


# ---------------------------------------------------------------------------------------------- data terms
@dataclass(eq=False)
class PositionEnergy(UnitEnergyConstructor):
    """-2 * (detection_map[x, y] - threshold)   (data_energies.py:13-24)."""
    detection_map: Any = None
    threshold: float = 0.0


@dataclass(eq=False)
class ShapeEnergy(UnitEnergyConstructor):
    """mean_i M_i[x, y, class_i(mark_i)]   (data_energies.py:28-45).

    Either give the raw mark distributions + the calibration (`param_dist_maps`, `remap_coefs`, `remap_intercepts`:
    M_i = -2 sigmoid(c_i P_i + b_i) + 1 is then evaluated on the fly by the kernels, energy_setup_legacy.py:142-147),
    or, as in the reference, pre-computed energy maps in `parameter_energy_map` (used as they are)."""
    parameter_energy_map: Optional[List[Any]] = None
    mappings: List[ValueMapping] = None
    param_names: List[str] = None
    param_dist_maps: Optional[List[Any]] = None
    remap_coefs: Optional[Sequence[float]] = None
    remap_intercepts: Optional[Sequence[float]] = None


@dataclass(eq=False)
class SingleMarkEnergy(UnitEnergyConstructor):
    """M[x, y, class(mark)]   (data_energies.py:49-64).  `param_dist_map` (raw distribution, energy = -P,
    energy_setup_no_calibration.py:71) or `parameter_energy_map` (pre-computed, used as it is)."""
    parameter_energy_map: Any = None
    mapping: ValueMapping = None
    param_name: str = None
    param_dist_map: Any = None


# ---------------------------------------------------------------------------------------------- prior terms
@dataclass(eq=False)
class RectangleOverlapEnergy(PairEnergyConstructor):
    """area(P1 n P2) / (min(area1, area2) + 1e-6), max over partners   (prior_energies.py:12-24)."""


@dataclass(eq=False)
class ShapeAlignmentEnergy(PairEnergyConstructor):
    """1 - |cos(angle1 - angle2)| - [rewarding]; min over partners if rewarding else max   (prior_energies.py:28-50)."""
    rewarding: bool = True
    angle_param_name: str = "angle"


@dataclass(eq=False)
class AreaPriorEnergy(UnitEnergyConstructor):
    """max(0, min_area - A, A - max_area)   (prior_energies.py:54-67)."""
    min_area: float = 0.0
    max_area: float = 1e30


@dataclass(eq=False)
class RatioPriorEnergy(UnitEnergyConstructor):
    """|target_ratio - ratio|   (prior_energies.py:71-78)."""
    target_ratio: float = 0.5


# ---------------------------------------------------------------------------------------------- toy terms
@dataclass(eq=False)
class ConstantUnitEnergy(UnitEnergyConstructor):
    """Device version of the reference tests' toy unit term (test/test_energy_graph.py:15-23: returns a constant)."""
    value: float = 0.0


@dataclass(eq=False)
class DistanceIndicatorPairEnergy(PairEnergyConstructor):
    """Device version of the reference tests' toy pair term: `value` if distance <= max_dist (strict=False,
    test/test_energy_graph.py:26-35) or < max_dist (strict=True, test/test_interacting_points_set.py:32-43), else 0;
    max-reduced over partners."""
    value: float = 1.0
    strict: bool = False


EnergyConstructor = Any
