"""MPP inference glue: models/mpp/mpp_model.py:43-399 (MPPModel) and models/mpp/data_loaders.py:30-161, 252-260
(load_image_w_maps, crop_image_w_maps, merge_patches, labels_to_rectangles), on the device sampler.

Differences from the reference, all on the host side of the hot path:
  * `MPPModel.infer` samples each image as ONE scene (no 256^2 tiling / process pool / merge): the proposal budget of
    the reference's tiling is kept through `iter_multiplier = number of patches`; `infer(..., tile=True)` reproduces the
    tile -> sample -> merge_patches flow for comparison.
  * the model files are read with a restricted unpickler (only numpy arrays and the combinator / mapping dataclasses);
    a 'manual' config builds the combinator from the JSON weights like MPPModel.train (mpp_model.py:161-183).
  * DOTA export / AP evaluation need the un-vendored DOTA_devkit and stay out of scope; results are returned and written
    as `{id}_results.pkl` with the reference's keys (mpp_model.py:356-366)."""
from __future__ import annotations

import io
import logging
import os
import pickle
import re
import time
from copy import copy
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

from .combination import HierarchicalEnergyCombinator, LogisticEnergyCombinator, ManualHierarchicalEnergyCombinator
from .custom_types import EnergyCombinationModel, ImageWMaps
from .energy_point_set import EPointsSet
from .energy_setups import EnergySetup, LegacyEnergySetup, NoCalibrationEnergySetup
from .mappings import ValueMapping, default_mappings
from .rjmcmc import sample_rjmcmc
from .shapes import Rectangle, rect_to_poly, sra_to_wla, wla_to_sra

PARAM_NAMES = ["size", "ratio", "angle"]

# ------------------------------------------------------------------------------------------------ restricted unpickling
_ALLOWED = {
    ("models.mpp.energies.combination.hierarchical", "HierarchicalEnergyCombinator"): HierarchicalEnergyCombinator,
    ("models.mpp.energies.combination.hierarchical", "ManualHierarchicalEnergyCombinator"): ManualHierarchicalEnergyCombinator,
    ("models.mpp.energies.combination.logistic", "LogisticEnergyCombinator"): LogisticEnergyCombinator,
    ("models.shape_net.mappings", "ValueMapping"): ValueMapping,
}


class _RestrictedUnpickler(pickle.Unpickler):
    """Loads the reference's model / inference pickles without executing anything else than numpy array reconstruction
    and the construction of the whitelisted dataclasses (mapped onto this package's classes)."""

    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return _ALLOWED[(module, name)]
        if module in ("numpy.core.multiarray", "numpy._core.multiarray") and name in ("_reconstruct", "scalar"):
            import numpy.core.multiarray as ma
            return getattr(ma, name)
        if module == "numpy" and name in ("ndarray", "dtype"):
            return getattr(np, name)
        if module in ("builtins",) and name in ("list", "dict", "tuple", "set", "float", "int", "str", "bool"):
            return getattr(__import__("builtins"), name)
        raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name}")


def restricted_load(path_or_bytes) -> Any:
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    return _RestrictedUnpickler(io.BytesIO(data)).load()


def load_energy_combination_model(path: str) -> EnergyCombinationModel:
    """models_storage/mpp/<name>/energy_combination_model.pkl (mpp_model.py:87-91)."""
    model = restricted_load(path)
    if not isinstance(model, (HierarchicalEnergyCombinator, ManualHierarchicalEnergyCombinator, LogisticEnergyCombinator)):
        raise pickle.UnpicklingError(f"{path} does not hold a supported energy combination model")
    return model


def combinator_from_manual_config(config: Dict[str, Any]) -> EnergyCombinationModel:
    """The 'manual' train mode (mpp_model.py:161-183): weights straight from the config JSON."""
    w = config["manual"]
    if config.get("energy_setup") in (None, "legacy"):
        def normalize(a):
            a = np.array(a, dtype=np.float64)
            return a / np.linalg.norm(a, ord=1)
        return HierarchicalEnergyCombinator(
            weights_data=normalize([w[k] for k in ("PositionEnergy", "ShapeEnergy")]),
            weights_prior=normalize([w[k] for k in ("RectangleOverlapEnergy", "ShapeAlignmentEnergy", "AreaPriorEnergy")]),
            data_prior_weights=normalize([w[k] for k in ("Data", "Prior")]), detection_threshold=w.get("threshold", 0.0))
    return ManualHierarchicalEnergyCombinator(weights_dict=w.get("weights"), indicator_energy=w.get("indicator_energy"),
                                              detection_threshold=w.get("threshold"))


# ------------------------------------------------------------------------------------------------ data_loaders.py
def labels_to_rectangles(labels, param_names: List[str] = None) -> List[Rectangle]:
    """(a, b, angle) annotations -> Rectangle(size, ratio, angle mod pi) (data_loaders.py:252-260)."""
    out = []
    for c, p in zip(labels["centers"], labels["parameters"]):
        s, r, a = wla_to_sra(p[0], p[1], p[2])
        out.append(Rectangle(int(c[0]), int(c[1]), size=s, ratio=r, angle=a % np.pi))
    return out


def load_image_w_maps(patch_id, data_dir: str, posnet_dir: str, shapenet_dir: str) -> ImageWMaps:
    """data_loaders.py:30-71 with explicit directories: <data_dir>/{images,annotations}/{id:04}.{png,pkl},
    <posnet_dir>/{id:04}_results.pkl['detection_map'], <shapenet_dir>/{id:04}_results.pkl['output','mappings']."""
    patch_id = int(patch_id)
    import cv2
    image = cv2.imread(os.path.join(data_dir, "images", f"{patch_id:04}.png"), cv2.IMREAD_COLOR)
    if image is None:
        raise FileNotFoundError(os.path.join(data_dir, "images", f"{patch_id:04}.png"))
    image = image[:, :, ::-1].astype(np.float32) / 255.0
    labels = restricted_load(os.path.join(data_dir, "annotations", f"{patch_id:04}.pkl"))
    detection_map = restricted_load(os.path.join(posnet_dir, f"{patch_id:04}_results.pkl"))["detection_map"]
    shapenet = restricted_load(os.path.join(shapenet_dir, f"{patch_id:04}_results.pkl"))
    param_dist_maps = [np.ascontiguousarray(np.moveaxis(p[0], 0, -1)) for p in shapenet["output"]]  # (1,32,H,W) -> (H,W,32)
    return ImageWMaps(image=image, name=f"{patch_id:04}", shape=image.shape[:2], detection_map=detection_map,
                      param_dist_maps=param_dist_maps, mappings=shapenet.get("mappings", default_mappings()), param_names=PARAM_NAMES,
                      labels=labels, gt_config=labels_to_rectangles(labels, PARAM_NAMES))


def crop_image_w_maps(image_data: ImageWMaps, tl_anchor: np.ndarray, patch_size: int) -> ImageWMaps:
    """data_loaders.py:74-119."""
    s = np.s_[tl_anchor[0]:tl_anchor[0] + patch_size, tl_anchor[1]:tl_anchor[1] + patch_size]
    det = image_data.detection_map[s]
    marks = [p[s] for p in image_data.param_dist_maps]
    image = None if image_data.image is None else image_data.image[s]
    labels, gt = None, None
    if image_data.labels is not None:
        keep = [j for j, c in enumerate(image_data.labels["centers"])
                if np.all(np.asarray(c) - tl_anchor >= 0) and np.all(np.asarray(c) - tl_anchor < np.array(det.shape[:2]))]
        labels = {k: np.array([image_data.labels[k][j] for j in keep]) for k in ("parameters", "categories", "difficult") if k in image_data.labels}
        labels["centers"] = np.array([np.asarray(image_data.labels["centers"][j]) - tl_anchor for j in keep])
        gt = labels_to_rectangles(labels, PARAM_NAMES)
    return ImageWMaps(image=image, name=image_data.name, shape=tuple(det.shape[:2]), detection_map=det, param_dist_maps=marks,
                      mappings=image_data.mappings, param_names=PARAM_NAMES, labels=labels, gt_config=gt,
                      crop_data={"tl_anchor": np.asarray(tl_anchor)})


def merge_patches(patches: List[ImageWMaps], results: List[List[Rectangle]], original_image: ImageWMaps,
                  energy_model: EnergyCombinationModel, method: str, energy_setup: EnergySetup, **kwargs):
    """data_loaders.py:122-161: all patch detections in one full-image EPointsSet; clusters closer than `distance` keep
    their best Papangelou score.  Neighbour queries and scores run on the device."""
    assert method in ["distance"], "only the 'distance' merge of the reference's inference path is built"
    uec, pec = energy_setup.make_energies(image_data=original_image)
    moved = []
    for patch, result in zip(patches, results):
        anchor = patch.crop_data["tl_anchor"]
        for r in result:
            new_r = copy(r)
            new_r.x, new_r.y = int(new_r.x + anchor[0]), int(new_r.y + anchor[1])
            moved.append(new_r)
    agg = EPointsSet(points=moved, support_shape=original_image.shape, unit_energies_constructors=uec, pair_energies_constructors=pec)
    distance = kwargs["distance"]
    to_remove = set()
    for p in list(agg):
        if p in to_remove:
            continue
        neigh = list(agg.points.get_neighbors(p, radius=distance, exclude_itself=False) - to_remove)
        if not neigh:
            continue
        scores = [agg.papangelou(q, energy_combinator=energy_model, remove_u_from_point_set=True) for q in neigh]
        best = neigh[int(np.argmax(scores))]
        to_remove |= set(neigh)
        to_remove -= {best}
    logging.info(f"merge removing {len(to_remove)} point(s)")
    for p in to_remove:
        agg.remove(p)
    return agg.points


# ------------------------------------------------------------------------------------------------ mpp_model.py
class MPPModel:
    """Inference facade with the reference's constructor / infer flow (mpp_model.py:43-99, 202-370)."""

    def __init__(self, config: Dict[str, Any], phase: str = "val", overwrite=False, load=True, dataset: str = None,
                 model_dir: Optional[str] = None):
        assert phase in ["val", "train"]
        if phase == "train" or not load:
            raise NotImplementedError("training / calibration of the MPP model is out of scope: load a trained model")
        self.config = config
        self.rng = np.random.default_rng(0)
        self.save_path = model_dir
        setup_conf = config.get("energy_setup") or "legacy"
        params = config.get("energy_setup_params") or {}
        if setup_conf == "legacy":
            self.energy_setup: EnergySetup = LegacyEnergySetup(calibration_params=config.get("calibration", {}).get("params", {}))
        elif setup_conf == "no-calibration":
            self.energy_setup = NoCalibrationEnergySetup(**params)
        else:
            raise ValueError("energy_setup must be one of : 'legacy', 'no-calibration' (the 'contrast' setup is not built)")
        self.energy_model: Optional[EnergyCombinationModel] = None
        if model_dir is not None:
            self.energy_setup.load_calibration(model_dir)
            pkl = os.path.join(model_dir, "energy_combination_model.pkl")
            if os.path.exists(pkl):
                self.energy_model = load_energy_combination_model(pkl)
        if self.energy_model is None and "manual" in config:
            self.energy_model = combinator_from_manual_config(config)
        if self.energy_model is None:
            raise FileNotFoundError("no energy_combination_model.pkl and no 'manual' weights in the config")

    def infer_image(self, image_data: ImageWMaps, tile: bool = False, **sampler_options) -> Dict[str, Any]:
        """One image: sampling (+ merge when tile=True) and Papangelou scoring (mpp_model.py:228-304, 356-366)."""
        params = dict(self.config["inference"]["rjmcmc_params"])
        shape = tuple(image_data.shape[:2])
        patch_size = 256
        nx, ny = int(np.ceil(shape[0] / patch_size)), int(np.ceil(shape[1] / patch_size))
        start = time.perf_counter()
        if not tile:
            mult = params.pop("iter_multiplier", None) or 1
            out = sample_rjmcmc(image_data=image_data, rng=self.rng, num_samples=1, energy_combinator=self.energy_model,
                                init_config="naive", energy_setup=self.energy_setup, iter_multiplier=mult * max(1, nx * ny),
                                **params, **sampler_options)
            result = EPointsSet([], shape, *self.energy_setup.make_energies(image_data), _state=out[-1]._state)
        else:
            anchors_x = np.linspace(0, shape[0] - patch_size, max(1, nx), dtype=int)  # mpp_model.py:236-237
            anchors_y = np.linspace(0, shape[1] - patch_size, max(1, ny), dtype=int)
            patches = [crop_image_w_maps(image_data, np.array([xa, ya]), patch_size) for xa in anchors_x for ya in anchors_y]
            results = [list(sample_rjmcmc(image_data=p, rng=self.rng, num_samples=1, energy_combinator=self.energy_model,
                                          init_config="naive", energy_setup=self.energy_setup, **params, **sampler_options)[-1])
                       for p in patches]
            merged = merge_patches(patches, results, image_data, self.energy_model, "distance", self.energy_setup, distance=3)
            result = EPointsSet([], shape, *self.energy_setup.make_energies(image_data), _state=merged._state)
        objs, scores = result.papangelou_all(energy_combinator=self.energy_model)
        logging.info(f"image {image_data.name}: {len(objs)} objects in {time.perf_counter() - start:.2f}s")
        pred_params = [sra_to_wla(p.size, p.ratio, p.angle) for p in objs]
        pred_centers = np.array([[p.x, p.y] for p in objs]).reshape(-1, 2)
        polys = np.array([rect_to_poly(c, p[0], p[1], p[2]) for c, p in zip(pred_centers, pred_params)]).reshape(-1, 4, 2)
        max_score = self.config["inference"].get("max_score") or 4.0
        return {"detection": polys, "detection_points": result.points, "detection_type": "poly", "detection_center": pred_centers,
                "detection_score": [float(s) for s in scores], "detection_score_01": np.asarray(scores) / max_score,
                "detection_params": pred_params, "mappings": image_data.mappings}

    def infer(self, images: Iterable[ImageWMaps], results_dir: Optional[str] = None, overwrite: bool = True, tile: bool = False,
              **sampler_options) -> List[Dict[str, Any]]:
        """mpp_model.py:202-370 over already loaded images (use load_image_w_maps for the reference's on-disk format)."""
        out = []
        for image_data in images:
            target = None if results_dir is None else os.path.join(results_dir, f"{image_data.name}_results.pkl")
            if target is not None and os.path.exists(target) and not overwrite:
                print(f"{os.path.basename(target)} exists, skipping")
                continue
            res = self.infer_image(image_data, tile=tile, **sampler_options)
            if target is not None:
                os.makedirs(results_dir, exist_ok=True)
                with open(target, "wb") as f:
                    pickle.dump({k: (list(v) if k == "detection_points" else v) for k, v in res.items()}, f)
            out.append(res)
        return out


def resolve_model_config_path(name_or_path: str, search_dirs: Iterable[str] = ()) -> str:
    """A path to a config .json, or a model name looked up as <dir>/<name>/config.json or <dir>/*/<name>.json
    (utils/data.py:114-132)."""
    if os.path.exists(name_or_path):
        return name_or_path
    import glob
    for d in search_dirs:
        for pat in (os.path.join(d, name_or_path, "config.json"), os.path.join(d, "*", name_or_path, "config.json"),
                    os.path.join(d, "*", name_or_path + ".json"), os.path.join(d, name_or_path + ".json")):
            hits = glob.glob(pat)
            if hits:
                return hits[-1]
    print(f"no model with name (or config with path) {name_or_path}")
    raise FileNotFoundError(name_or_path)


def image_ids(data_dir: str) -> List[int]:
    id_re = re.compile(r"([0-9]+).*\.png")
    return sorted(int(id_re.match(f).group(1)) for f in os.listdir(os.path.join(data_dir, "images")) if id_re.match(f))
