"""Host-side mirror of the reference's Python interface for the MPP sampling path: same class
and function names, argument meaning and error behaviour as models/mpp, base/shapes and models/shape_net/mappings in the
reference; all computation goes through the C ABI of include/mpp_b200.h to the CUDA kernels."""
from .combination import (HierarchicalEnergyCombinator, LogisticEnergyCombinator, ManualHierarchicalEnergyCombinator,  # noqa: F401
                          MLPEnergyCombinator)
from .custom_types import EnergyCombinationModel, ImageWMaps, Perturbation, RJMCMCStateSummary  # noqa: F401
from .energies import (AreaPriorEnergy, ConstantUnitEnergy, DistanceIndicatorPairEnergy, PairEnergy, PairEnergyConstructor,  # noqa: F401
                       PositionEnergy, RatioPriorEnergy, RectangleOverlapEnergy, ShapeAlignmentEnergy, ShapeEnergy,
                       SingleMarkEnergy, UnitEnergy, UnitEnergyConstructor)
from .energy_graph import EnergyGraph  # noqa: F401
from .energy_utils import compute_energy_vector, compute_many_energy_vectors, names_from_energies  # noqa: F401
from .energy_point_set import EPointsSet  # noqa: F401
from .energy_setups import EnergySetup, LegacyEnergiesCalibration, LegacyEnergySetup, NoCalibEnergiesCalibration, NoCalibrationEnergySetup  # noqa: F401
from .kernels import (BirthKernel, DataDrivenShapeTransformKernel, DataDrivenTranslationKernel, DeathKernel,  # noqa: F401
                      GaussianShapeTransformKernel, GaussianTranslationKernel, Kernel, MergeKernel, SplitKernel, SplitSampler,
                      make_kernels)
from .mappings import ValueMapping, default_mappings, output_vector_to_value  # noqa: F401
from .perturbation_sampler import (PERTURBATION_HP_MEDIUM, PERTURBATION_LIGHT, PERTURBATION_MEDIUM, PERTURBATION_MEDIUM_OVERLAP,  # noqa: F401
                                   PERTURBATION_STRONG, aggregate_perturbations, sample_kernel_perturbations,
                                   sample_multiple_kernel_perturbations, sample_perturbations)
from .point_set import PointsSet  # noqa: F401
from .rjmcmc import RJMCMC, RJMCMCTimer, StopOnMaxIter, naive_detection, nms_distance, sample_rjmcmc, sample_rjmcmc_batch, sample_rjmcmc_tiles  # noqa: F401
from .sampler2d import sample_point_2d  # noqa: F401
from .shapes import Point, Rectangle, polygon_to_abw, rect_to_poly, rotation_matrix, sra_to_wla, wla_to_sra  # noqa: F401
from .mpp_model import (MPPModel, combinator_from_manual_config, crop_image_w_maps, labels_to_rectangles, load_energy_combination_model,  # noqa: F401
                        load_image_w_maps, merge_patches, resolve_model_config_path, restricted_load)
