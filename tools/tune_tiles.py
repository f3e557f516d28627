"""Throughput of a batch of independent 512^2 tiles (BASELINE configs[4]) on one GPU: sequential vs concurrent streams."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mpp_cnn_rs_object_detection_b200 import synth
from mpp_cnn_rs_object_detection_b200.engine import Engine, ModelSpec

n_tiles, size, temp = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 512, 0.02
dev = torch.device("cuda", 0)
C, H = bench.CALIB_HRCM, bench.HRC
spec = ModelSpec(setup="legacy", pos_threshold=C["detection_threshold"], remap_coefs=C["coefs"], remap_intercepts=C["intercepts"],
                 min_area=C["min_area"], max_area=C["max_area"], combinator="hierarchical",
                 comb_w=list(H["weights_data"]) + list(H["weights_prior"]) + list(H["data_prior_weights"]) + [0.0])
tiles = [synth.make_scene_torch(100 + i, (size, size), 162, dev) for i in range(n_tiles)]
NW = int(os.environ.get('NW', '8'))
for n_streams in (8, 32):
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    engines = []
    for i, (objs, det, marks) in enumerate(tiles):
        with torch.cuda.stream(streams[i % n_streams]):
            e = Engine((size, size), device=dev)
            e.set_maps(det, marks); e.set_model(spec); e.set_kernels(intensity=max(1, len(objs)))
            e.add_objects(objs[:, :2], objs[:, 2:5])
            engines.append(e)
    torch.cuda.synchronize()
    sweeps, pv, nw = 12, 96, NW
    for rep in range(2):
        t = time.perf_counter()
        for i, e in enumerate(engines):
            with torch.cuda.stream(streams[i % n_streams]):
                e.run_windows(sweeps, pv, nw, t0=temp, seed=i, sweep_offset=rep * sweeps, read_counters=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    ev = sum(e.run_windows(0, pv, nw, t0=temp)[4] for e in engines) / 2
    print(f"{n_tiles} tiles, {n_streams} streams: {dt * 1e3:.1f} ms per batch, {ev / dt / 1e6:.1f} M proposals/s", flush=True)
    for e in engines:
        e.close()
