"""Per-visit phase timing of the window sampler (development tool).  Needs an instrumented build of the library:

    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DMPP_TRACE \
         -o tools/_trace_build.so mpp_cnn_rs_object_detection_b200/csrc/mpp_b200.cu
    MPP_B200_DEBUG=1 MPP_B200_DEBUG_LIB=tools/_trace_build.so python tools/visit_timers.py

(the library override is only honoured with MPP_B200_DEBUG=1 and for paths inside the repository)."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mpp_cnn_rs_object_detection_b200 import synth, _lib
from mpp_cnn_rs_object_detection_b200.engine import Engine, ModelSpec
size = 2048
dev = torch.device("cuda", 0)
objs, det, marks = synth.make_scene_torch(0, (size, size), 2600, dev)
Cc, H = bench.CALIB_HRCM, bench.HRC
spec = ModelSpec(setup="legacy", pos_threshold=Cc["detection_threshold"], remap_coefs=Cc["coefs"], remap_intercepts=Cc["intercepts"],
                 min_area=Cc["min_area"], max_area=Cc["max_area"], combinator="hierarchical",
                 comb_w=list(H["weights_data"]) + list(H["weights_prior"]) + list(H["data_prior_weights"]) + [0.0])
for nw, pv in ((8, 96),):
    eng = Engine((size, size), device=dev)
    eng.set_maps(det, marks); eng.set_model(spec); eng.set_kernels(intensity=max(1, len(objs)))
    eng.add_objects(objs[:, :2], objs[:, 2:5])
    eng.run_windows(3, pv, nw, t0=0.02, seed=1)
    dbg = torch.zeros(8 + 4000 * 14, dtype=torch.float32, device=dev)
    cnt = (C.c_ulonglong * 8)()
    _lib.check(eng.lib.mpp_run_windows(eng.ctx, 1, pv, nw, 1, 0.02, 1.0, 0.0, 1, 3, cnt, dbg.data_ptr()))
    d = dbg.cpu().numpy()
    n = min(4000, int(d[1:2].view(np.int32)[0]))
    tr = d[8:8 + n * 14].reshape(n, 14)
    print(f"nw={nw} pv={pv}: visits traced {n}")
    occ = tr[:, 1] > 0
    for label, sel in (("empty windows", ~occ), ("occupied windows", occ), ("windows with >= 3 objects", tr[:, 1] >= 3)):
        t = tr[sel]
        if len(t):
            print(f"  [{label}: {len(t)} visits] total us mean {t[:, 10].mean() / 1900:.1f} p90 {np.percentile(t[:, 10], 90) / 1900:.1f} | staging (A..E) "
                  f"{(t[:, 2] + t[:, 3] + t[:, 4] + t[:, 5]).mean() / 1900:.1f} | eval {t[:, 6].mean() / 1900:.1f} | commit {t[:, 7].mean() / 1900:.1f} | rounds {t[:, 8].mean():.1f} | "
                  f"accepted {t[:, 9].mean():.1f} | staged n {t[:, 0].mean():.1f} | n_win {t[:, 1].mean():.2f}")
    for name, col in (("staged n", 0), ("n_win", 1), ("A heads us", 2), ("B rank us", 3), ("C recs us", 4), ("D+E us", 5), ("eval us", 6), ("commit us", 7),
                      ("rounds", 8), ("acc", 9), ("total us", 10), ("evaluated", 11), ("D+E pure us", 12), ("wait predraw us", 13)):
        v = tr[:, col] / 1900.0 if "us" in name else tr[:, col]
        print(f"  {name:12s} mean {v.mean():8.2f}  p50 {np.median(v):8.2f}  p90 {np.percentile(v, 90):8.2f}  max {v.max():8.2f}")

    raw = (C.c_ulonglong * 64)()
    _lib.check(eng.lib.mpp_window_stats(eng.ctx, raw))
    v = [int(x) for x in raw]
    names = ["uniform_birth", "uniform_death", "data_birth", "data_death", "gaussian_translation", "data_translation", "gaussian_transform", "data_transform"]
    if v[38]:
        print(f"  commits {v[38]}: removal + new entry {v[35] / v[38] / 1900:.2f} us | update of the neighbours' reductions {v[36] / v[38] / 1900:.2f} us | "
              f"new object's reductions {v[37] / v[38] / 1900:.2f} us")
    print("  per-kernel evaluation time of the generic rounds (us, instrumented build): draw part | Delta-energy part | count")
    for k in range(8):
        if v[56 + k]:
            print(f"    {names[k]:22s} {v[40 + k] / v[56 + k] / 1900:6.2f} | {v[48 + k] / v[56 + k] / 1900:6.2f} | {v[56 + k]}")
