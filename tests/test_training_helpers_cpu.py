"""Host-side training-time helpers against outputs of the reference's own code
(tests/golden/training_helpers.npz, written by oracle/gen_golden_training.py): sample_perturbations consumes the seeded
Generator in the reference's order, aggregate_perturbations nets a sequence of perturbations."""
import numpy as np

from tests import golden_util as gu


def _image(api, g):
    cfg = [api.Rectangle(x=int(r[0]), y=int(r[1]), size=float(r[2]), ratio=float(r[3]), angle=float(r[4])) for r in g["config"]]
    shape = tuple(int(v) for v in g["shape"])
    return api.ImageWMaps("t", shape, None, None, None, api.default_mappings(), ["size", "ratio", "angle"], gt_config=cfg), cfg


def test_sample_perturbations_reproduces_reference_draws():
    import mpp_cnn_rs_object_detection_b200.api as api
    g = gu.load("training_helpers.npz")
    image, _ = _image(api, g)
    for name, preset in (("light", api.PERTURBATION_LIGHT), ("overlap", api.PERTURBATION_MEDIUM_OVERLAP), ("strong", api.PERTURBATION_STRONG)):
        rng = np.random.default_rng(int(g["seed"]) + 11)
        cfgs = api.sample_perturbations(image_data=image, rng=rng, n_samples=3, **preset)
        assert [len(c) for c in cfgs] == g[f"pert_{name}_len"].tolist()
        flat = np.array([[r.x, r.y, r.size, r.ratio, r.angle] for c in cfgs for r in c], dtype=np.float64).reshape(-1, 5)
        np.testing.assert_allclose(flat, g[f"pert_{name}"], rtol=0, atol=1e-12)
    # the ground truth is not modified
    assert [(r.x, r.y) for r in image.gt_config] == [(int(r[0]), int(r[1])) for r in g["config"]]


def test_aggregate_perturbations_matches_reference():
    import mpp_cnn_rs_object_detection_b200.api as api
    g = gu.load("training_helpers.npz")
    _, config = _image(api, g)
    a = [api.Rectangle(x=5 + k, y=6 + k, size=8.0, ratio=0.5, angle=0.1 * k) for k in range(3)]
    P = api.Perturbation
    seq = [P(type=None, removal=config[0], addition=a[0]), P(type=None, removal=a[0], addition=a[1]),
           P(type=None, removal=[config[1], config[2]], addition=a[2]), P(type=None, removal=None, addition=config[1]),
           P(type=None, removal=a[2], addition=None)]
    agg = api.aggregate_perturbations(seq)
    idx = {id(p): k for k, p in enumerate(config)}
    idx.update({id(p): 100 + k for k, p in enumerate(a)})
    assert sorted(idx[id(p)] for p in agg.removal) == g["agg_removal"].tolist()
    assert sorted(idx[id(p)] for p in agg.addition) == g["agg_addition"].tolist()
