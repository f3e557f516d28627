"""Plain data carriers of the sampler API: models/mpp/custom_types/{perturbation,rjmcmc,image_w_maps,energy}.py."""
from __future__ import annotations

from abc import abstractmethod
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Type, Union

from .mappings import ValueMapping
from .shapes import Point, Rectangle

ConfigurationEnergyVector = Dict[str, List[float]]
PointEnergyVector = Dict[str, float]


class EnergyCombinationModel:  # custom_types/energy.py:8-11
    @abstractmethod
    def compute(self, vectors: ConfigurationEnergyVector) -> float:
        pass


@dataclass
class Perturbation:  # custom_types/perturbation.py:8-12
    type: Type
    removal: Union[None, Point, List[Point]] = None
    addition: Union[None, Point, List[Point]] = None
    data: Optional[Dict[str, Any]] = None


@dataclass
class RJMCMCStateSummary:  # custom_types/rjmcmc.py:6-14
    iter: int
    n_points: int
    temperature: float = None
    energy: Union[None, float] = None
    kernel: Union[None, Type] = None
    move_accepted: Union[None, bool] = None
    alpha: Union[None, float] = None
    initial_energy: Union[None, float] = None
    proposed_energy: Union[None, float] = None


@dataclass
class ImageWMaps:  # custom_types/image_w_maps.py:12-22
    name: str
    shape: Tuple[int, int]
    image: Any
    detection_map: Any            # (H,W) float32: numpy array or torch tensor (already on the device: used in place)
    param_dist_maps: List[Any]    # 3 x (H,W,32) float32
    mappings: List[ValueMapping]
    param_names: List[str]
    labels: Dict[str, Any] = None
    gt_config: List[Rectangle] = None
    gt_config_set: Any = None
    crop_data: Dict = None
