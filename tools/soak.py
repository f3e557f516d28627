"""Soak of the window sampler with the brute-force Delta-energy check on (development tool): several seeds, scene densities,
proposals per visit, speculation depths and temperatures; prints the largest |fast - brute-force| difference and checks the
counters and that the chain is the same for every speculation depth."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mpp_cnn_rs_object_detection_b200 import synth
from mpp_cnn_rs_object_detection_b200.engine import Engine, ModelSpec

dev = torch.device("cuda", 0)
C, H = bench.CALIB_HRCM, bench.HRC
spec = ModelSpec(setup="legacy", pos_threshold=C["detection_threshold"], remap_coefs=C["coefs"], remap_intercepts=C["intercepts"],
                 min_area=C["min_area"], max_area=C["max_area"], combinator="hierarchical",
                 comb_w=list(H["weights_data"]) + list(H["weights_prior"]) + list(H["data_prior_weights"]) + [0.0])
worst = 0.0
for seed, size, n_rect, temp, alpha in ((1, 256, 60, 0.02, 1.0), (2, 256, 160, 0.05, 0.7), (3, 384, 100, 0.3, 1.0), (4, 200, 40, 1.0, 0.2), (5, 512, 400, 0.02, 1.0)):
    objs, det, marks = synth.make_scene_torch(seed, (size, size + 37), n_rect, dev)
    for pv in (7, 37, 96, 128):
        finals = []
        for nw in (1, 4, 8):
            eng = Engine((size, size + 37), device=dev)
            eng.set_maps(det, marks); eng.set_model(spec); eng.set_kernels(intensity=max(1, len(objs)))
            eng.add_objects(objs[:, :2], objs[:, 2:5], uid=np.arange(len(objs)))  # explicit uids: the device would number them in arrival order
            cnt, maxdiff = eng.run_windows(6, pv, nw, t0=temp, alpha_t=alpha, t_target=0.001, seed=seed * 11, debug=True)
            worst = max(worst, maxdiff)
            _, xy, mk, uid = eng.read_objects()
            order = np.lexsort((uid, xy[:, 1], xy[:, 0]))
            finals.append((tuple(cnt[:5]), xy[order].tobytes(), mk[order].tobytes(), uid[order].tobytes()))
            assert cnt[4] <= cnt[0] and cnt[1] <= cnt[4] and len(eng) == len(objs) + cnt[2] - cnt[3], (cnt, len(eng), len(objs))
            eng.close()
        assert all(f == finals[0] for f in finals[1:]), ("chain depends on n_warps", seed, pv, [f[0] for f in finals])
        print(f"seed {seed} size {size} pv {pv}: counters {finals[0][0]} maxdiff so far {worst:.2e}", flush=True)
assert worst < 2e-4, worst
print("soak ok, worst |fast - brute-force Delta E| =", worst)
