"""Energy setups: what terms exist for an image and with which calibration.
models/mpp/energies/energy_utils.py:15-37 (EnergySetup), energy_setups/energy_setup_legacy.py:22-147 (mpp_hrcM),
energy_setups/energy_setup_no_calibration.py:20-140 (mpp_log).  `calibrate` (training-time) is out of scope: the
calibration ships as calibration.json."""
from __future__ import annotations

import json
import os
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Dict, List, Tuple

from .custom_types import ImageWMaps
from .energies import (AreaPriorEnergy, PairEnergyConstructor, PositionEnergy, RatioPriorEnergy, RectangleOverlapEnergy,
                       ShapeAlignmentEnergy, ShapeEnergy, SingleMarkEnergy, UnitEnergyConstructor)


class EnergySetup(ABC):
    @property
    @abstractmethod
    def energy_names(self) -> List[str]:
        raise NotImplementedError

    @abstractmethod
    def make_energies(self, image_data: ImageWMaps) -> Tuple[List[UnitEnergyConstructor], List[PairEnergyConstructor]]:
        raise NotImplementedError

    def calibrate(self, image_configs, rng, save_path: str = None):
        raise NotImplementedError("calibration is a training-time step of the reference (models/mpp/calibration); "
                                  "load the shipped calibration.json with load_calibration()")

    @abstractmethod
    def load_calibration(self, save_dir: str):
        raise NotImplementedError

    @property
    @abstractmethod
    def detection_threshold(self) -> float:
        raise NotImplementedError


@dataclass
class LegacyEnergiesCalibration:  # energy_setup_legacy.py:22-31
    detection_threshold: float
    param_dist_remap_coefs: List[float]
    param_dist_remap_intercepts: List[float]
    min_area: float
    max_area: float


@dataclass
class LegacyEnergySetup(EnergySetup):  # energy_setup_legacy.py:34-140
    calibration_params: Dict[str, Any] = None
    rewarding_priors: bool = True
    energy_calibration: LegacyEnergiesCalibration = None

    NAMES = ["PositionEnergy", "ShapeEnergy", "RectangleOverlapEnergy", "ShapeAlignmentEnergy", "AreaPriorEnergy"]

    @property
    def energy_names(self) -> List[str]:
        return self.NAMES.copy()

    def make_energies(self, image_data: ImageWMaps):
        cal = self.energy_calibration
        position = PositionEnergy(name=self.NAMES[0], detection_map=image_data.detection_map, threshold=cal.detection_threshold)
        # the remapped maps -2*sigmoid(c*P+b)+1 are not materialised: the kernels evaluate them at the gathered entries
        shape = ShapeEnergy(name=self.NAMES[1], mappings=image_data.mappings, param_names=image_data.param_names,
                            param_dist_maps=image_data.param_dist_maps, remap_coefs=cal.param_dist_remap_coefs,
                            remap_intercepts=cal.param_dist_remap_intercepts)
        overlap = RectangleOverlapEnergy(name=self.NAMES[2], max_dist=32)
        align = ShapeAlignmentEnergy(name=self.NAMES[3], max_dist=16, rewarding=self.rewarding_priors)
        area = AreaPriorEnergy(name=self.NAMES[4], min_area=cal.min_area, max_area=cal.max_area)
        return [position, shape, area], [overlap, align]

    def load_calibration(self, save_dir: str):
        with open(os.path.join(save_dir, "calibration.json"), "r") as f:
            d = json.load(f)
        self.energy_calibration = LegacyEnergiesCalibration(
            detection_threshold=d["detection_threshold"], param_dist_remap_coefs=d["param_dist_remap_coefs"],
            param_dist_remap_intercepts=d["param_dist_remap_intercepts"], min_area=d["min_area"], max_area=d["max_area"])

    @property
    def detection_threshold(self) -> float:
        return self.energy_calibration.detection_threshold


@dataclass
class NoCalibEnergiesCalibration:  # energy_setup_no_calibration.py:20-28
    min_area: float
    max_area: float
    param_dist_remap_coefs: List[float] = None
    param_dist_remap_intercepts: List[float] = None


class NoCalibrationEnergySetup(EnergySetup):  # energy_setup_no_calibration.py:31-140

    def __init__(self, rewarding_priors: bool = True, ratio_prior: bool = False, calib_marks: bool = False):
        if calib_marks:
            raise NotImplementedError("calib_marks=True (calibrated single-mark energies) is not used by the shipped models")
        self.energy_calibration: NoCalibEnergiesCalibration = None
        self.rewarding_priors = rewarding_priors
        self.ratio_prior = ratio_prior
        self.calib_marks = calib_marks
        self.NAMES = ["PositionEnergy", "SizeEnergy", "RatioEnergy", "AngleEnergy", "OverlapPriorEnergy",
                      "AlignmentPriorEnergy", "AreaPriorEnergy"]
        if self.ratio_prior:
            self.NAMES.append("RatioPriorEnergy")

    @property
    def energy_names(self) -> List[str]:
        return self.NAMES.copy()

    def make_energies(self, image_data: ImageWMaps):
        cal = self.energy_calibration
        position = PositionEnergy(name=self.NAMES[0], detection_map=image_data.detection_map, threshold=0)
        marks = [SingleMarkEnergy(name=self.NAMES[i + 1], mapping=image_data.mappings[i], param_name=p,
                                  param_dist_map=image_data.param_dist_maps[i])
                 for i, p in enumerate(["size", "ratio", "angle"])]
        overlap = RectangleOverlapEnergy(name=self.NAMES[4], max_dist=32)
        align = ShapeAlignmentEnergy(name=self.NAMES[5], max_dist=16, rewarding=self.rewarding_priors)
        area = AreaPriorEnergy(name=self.NAMES[6], min_area=cal.min_area, max_area=cal.max_area)
        unit = [position] + marks + [area]
        if self.ratio_prior:
            unit.append(RatioPriorEnergy(name=self.NAMES[7], target_ratio=0.5))
        return unit, [overlap, align]

    def load_calibration(self, save_dir: str):
        with open(os.path.join(save_dir, "calibration.json"), "r") as f:
            d = json.load(f)
        self.energy_calibration = NoCalibEnergiesCalibration(
            min_area=d["min_area"], max_area=d["max_area"], param_dist_remap_coefs=d.get("param_dist_remap_coefs"),
            param_dist_remap_intercepts=d.get("param_dist_remap_intercepts"))

    @property
    def detection_threshold(self) -> float:  # energy_setup_no_calibration.py: naive init thresholds the raw map at 0.5
        return 0.5
