"""Command line of the MPP path, with the reference's flags (main.py:10-105):

    python -m mpp_cnn_rs_object_detection_b200.main -p infer -m mpp -c mpp_hrcM [-d DATASET] [-o]
        --models-dir models_storage/mpp --data-root data_sample --inference-root inference_results

`-c` is a config .json or a model name resolved under --models-dir (utils/data.py:114-132).  Images, annotations and the
posnet / shapenet result pickles are read from the reference's directory layout (data_loaders.py:30-71).  `--synthetic HxW`
replaces the dataset by one synthetic scene (no files needed).  Only `-m mpp` with `-p infer` is in scope."""
from __future__ import annotations

import argparse
import json
import os
import sys


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-m", "--model", help="model to use (only 'mpp' is built)")
    ap.add_argument("-d", "--dataset", help="dataset to use, defaults to the one specified in config")
    ap.add_argument("-p", "--procedure", help="procedure to execute (only 'infer' is built)")
    ap.add_argument("-c", "--config", help="model config file, or a model name")
    ap.add_argument("-o", "--overwrite", action="store_true", help="overwrite existing results")
    ap.add_argument("-r", "--resume", action="store_true")
    ap.add_argument("--models-dir", default="models_storage/mpp", help="where <name>/config.json, calibration.json, *.pkl live")
    ap.add_argument("--data-root", default="data_sample", help="<data-root>/<dataset>/<subset>/{images,annotations}")
    ap.add_argument("--inference-root", default="inference_results", help="<root>/<dataset>/<subset>/<model>/{id}_results.pkl")
    ap.add_argument("--subset", default="val")
    ap.add_argument("--synthetic", default=None, help="HxW: run on one synthetic scene instead of a dataset")
    ap.add_argument("--tile", action="store_true", help="reference flow: 256^2 patches + merge instead of whole-image sampling")
    args = ap.parse_args(argv)
    if args.model != "mpp":
        raise ValueError(f"model {args.model!r}: only the MPP sampling path (-m mpp) is built")
    if args.procedure not in ("infer",):
        raise ValueError(f"procedure {args.procedure!r}: only -p infer is built for the MPP path")

    from .api import mpp_model as mm
    config_file = mm.resolve_model_config_path(args.config, search_dirs=[args.models_dir, os.path.dirname(args.models_dir) or "."])
    with open(config_file) as f:
        config = json.load(f)
    if args.dataset:
        config["dataset"]["dataset"] = args.dataset
    model_dir = os.path.dirname(config_file) if os.path.exists(os.path.join(os.path.dirname(config_file), "calibration.json")) \
        else os.path.join(args.models_dir, config.get("model_name", ""))
    model = mm.MPPModel(config, phase="val", load=True, model_dir=model_dir)
    print("infering on dataset")
    if args.synthetic:
        from . import synth
        from .api import ImageWMaps, default_mappings
        h, w = (int(v) for v in args.synthetic.lower().split("x"))
        objs, det, marks = synth.make_scene(0, (h, w), max(1, int(2600 * h * w / 2048 ** 2)))
        images = [ImageWMaps("synthetic", (h, w), None, det, marks, default_mappings(), mm.PARAM_NAMES)]
        results_dir = None
    else:
        ds = config["dataset"]
        data_dir = os.path.join(args.data_root, ds["dataset"], args.subset)
        pos_dir = os.path.join(args.inference_root, ds["dataset"], args.subset, ds["position_model"])
        shp_dir = os.path.join(args.inference_root, ds["dataset"], args.subset, ds["shape_model"])
        images = (mm.load_image_w_maps(i, data_dir, pos_dir, shp_dir) for i in mm.image_ids(data_dir))
        results_dir = os.path.join(args.inference_root, ds["dataset"], args.subset, config.get("model_name", "mpp"))
    results = model.infer(images, results_dir=results_dir, overwrite=args.overwrite or results_dir is None, tile=args.tile)
    for r in results:
        print(f"{len(r['detection_score'])} objects, scores in [{min(r['detection_score'], default=0):.3f}, {max(r['detection_score'], default=0):.3f}]")
    print("done !")
    return results


if __name__ == "__main__":
    main(sys.argv[1:])
