"""Bit-identity of a scene split across GPUs (one process per GPU, peer access over NVLink) with the single-GPU chain.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/split_check.py [--size 2048x1024]

Every rank builds its band of the same seeded scene (band-local maps), attaches its neighbours' contexts through CUDA IPC and
runs the persistent dataflow kernel over its band; rank 0 also runs mpp_run_windows on the whole scene on its own GPU.  The
union of the bands' objects (positions, marks, uids) and the summed counters must equal the single-GPU run exactly."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="2048x1024")
    ap.add_argument("--n-rect", type=int, default=0)
    ap.add_argument("--sweeps", type=int, default=6)
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--per-visit", type=int, default=48)
    ap.add_argument("--temperature", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg, synth
    from mpp_cnn_rs_object_detection_b200.engine import Engine
    from tests.gpu_util import model_spec

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    h, w = (int(v) for v in args.size.split("x"))
    n_rect = args.n_rect or int(round(2600 * h * w / (2048.0 * 2048.0)))
    m0, m1 = mg.PeerSplitScene.map_rows(h, rank, world)
    objs, det, marks = synth.make_scene_band_torch(args.seed, (h, w), n_rect, device, row0=m0, rows=m1 - m0)
    det_sum = float(np.sum(det))
    uid = np.arange(len(objs))
    eng = Engine((h, w), device=device)
    eng.set_maps_band(torch.as_tensor(det[m0:m1]).to(device), marks, m0, det_sum)
    eng.set_model(model_spec("legacy"))
    eng.set_kernels(intensity=max(1, len(objs)))
    scene = mg.PeerSplitScene(eng, h, rank, world)
    sel = scene.select_initial(objs[:, :2])
    eng.add_objects(objs[sel, :2], objs[sel, 2:5], uid=uid[sel])
    scene.attach_dist()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for k in range(args.calls):
        scene.run(args.sweeps, args.per_visit, 8, args.temperature, args.seed, sweep_offset=k * args.sweeps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    cnt = np.array(eng.run_windows(0, args.per_visit, 8, t0=args.temperature)[:5], dtype=np.int64)
    xy, mk, u = scene.owned_objects()
    parts = [None] * world
    if world > 1:
        dist.all_gather_object(parts, (xy, mk, u, cnt))
    else:
        parts = [(xy, mk, u, cnt)]
    ok = True
    if rank == 0:
        got_xy, got_mk, got_u = (np.concatenate([p[i] for p in parts]) for i in range(3))
        got_cnt = sum(p[3] for p in parts)
        _, _, marks_full = synth.make_scene_band_torch(args.seed, (h, w), n_rect, device, objs=objs, det=det)
        ref = Engine((h, w), device=device)
        ref.set_maps(torch.as_tensor(det).to(device), marks_full, det_sum=det_sum)
        ref.set_model(model_spec("legacy"))
        ref.set_kernels(intensity=max(1, len(objs)))
        ref.add_objects(objs[:, :2], objs[:, 2:5], uid=uid)
        ref_cnt = np.zeros(5, dtype=np.int64)
        for k in range(args.calls):
            ref_cnt += np.array(ref.run_windows(args.sweeps, args.per_visit, 8, t0=args.temperature, seed=args.seed,
                                                sweep_offset=k * args.sweeps)[:5], dtype=np.int64)
        _, rxy, rmk, ru = ref.read_objects()
        o1, o2 = np.lexsort((got_u, got_xy[:, 1], got_xy[:, 0])), np.lexsort((ru, rxy[:, 1], rxy[:, 0]))
        same = (len(got_u) == len(ru) and np.array_equal(got_xy[o1], rxy[o2]) and np.array_equal(got_mk[o1], rmk[o2]) and
                np.array_equal(got_u[o1], ru[o2]) and np.array_equal(got_cnt, ref_cnt))
        ok = bool(same)
        print(f"split_check: {h}x{w} scene, {len(objs)} objects at start, {world} rank(s), {args.calls} x {args.sweeps} sweeps x {args.per_visit} "
              f"proposals per window visit at T={args.temperature}")
        print(f"  split : counters {got_cnt.tolist()}  objects at end {len(got_u)}  ({dt * 1e3:.1f} ms wall for all calls, max band)")
        print(f"  single: counters {ref_cnt.tolist()}  objects at end {len(ru)}")
        print("  RESULT: " + ("IDENTICAL (positions, marks, uids, counters)" if ok else "MISMATCH"))
    scene.detach()
    eng.close()
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=device)
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
