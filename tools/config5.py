"""BASELINE config #5: the end-to-end per-frame pipeline on N synthetic KITTI-shaped frames sharded over the ranks, one allgather of the
observation records, then one graph build + one linearisation of the global camera-object graph.

    python tools/config5.py [--frames 10000]                                  (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/config5.py

Every rank pushes its share of the frames through the gray-frame entry (host buffers: frames up every step, Canny + distance transform +
proposal kernels on the GPU, D contexts in flight), appends the per-box observation records (csb_detect_observations_device) to one device
buffer, and the ranks exchange that buffer with ONE all_gather (NCCL).  Rank 0 then assembles the graph on the host (cube_slam_wu_b200/graph.py,
the recipe of main_obj.cpp:738-803), hands it to csb_ba_set_graph and linearises it (replicas only: SURVEY.md 8e).  The frames are the
bench's 64 distinct synthetic frames per rank, reused step after step; rank 0 prints one JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(n_frames_total=10000, depth=6, ctx=None, keep=False):
    import torch
    import torch.distributed as dist
    import bench
    import cube_slam_wu_b200 as csb
    import helpers as H
    from cube_slam_wu_b200 import graph

    rank = int(os.environ.get("RANK", "0")) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    dev = torch.cuda.current_device()
    params = csb.DetectParams.default()
    batch = bench.build_batch(rank)
    F, BPF = bench.FRAMES_PER_GPU, bench.BOXES_PER_FRAME
    frames, boxes, lines, tasks, n_tasks, maps, n_map = H.gpu_inputs(csb, batch, params)
    tb, boxes = bench.pinned(boxes); tl, lines = bench.pinned(lines)
    tg, gray = bench.pinned(np.concatenate([im.ravel() for im in batch["images"]]).astype(np.uint8))
    n_boxes = boxes.shape[0]
    S = (n_frames_total + F * world - 1) // (F * world)  # steps per rank
    ctxs = [csb.Context(dev) for _ in range(depth)]
    L = csb.lib()
    obs = torch.zeros(S, n_boxes * 16, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def frames_stage(n_steps, out):
        for s in range(n_steps):
            c = ctxs[s % depth]
            if s >= depth:
                c.synchronize()  # the context's previous step has consumed its staging buffers
            c.detect_upload_gray(frames, boxes, lines, tasks, n_tasks, gray, params)
            c.detect_run(timed=False)
            rc = L.csb_detect_observations_device(c._h, C.c_void_p(out[s].data_ptr()))
            assert rc == 0
        for c in ctxs:
            c.synchronize()

    frames_stage(min(S, 2 * depth), obs)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    frames_stage(S, obs)
    t_frames = time.perf_counter() - t0
    t0 = time.perf_counter()
    if world > 1:
        allobs = torch.zeros(world, S * n_boxes * 16, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allobs.view(-1), obs.view(-1))
    else:
        allobs = obs.view(1, -1)
    rec = allobs.cpu().numpy().reshape(world, S, n_boxes, 16).copy()
    t_gather = time.perf_counter() - t0
    tt = torch.tensor([t_frames, t_gather], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_frames, t_gather = float(tt[0]), float(tt[1])
    for c in ctxs:
        c.close()
    out = None
    if rank == 0:
        t0 = time.perf_counter()
        # global frame / landmark indices (graph.globalise_records): every rank reuses its own 64 frames; data association is not part of the
        # reference (its data set has one object): every 2D box of a rank's frames is its own landmark, observed again at every pass
        flat, n_landmarks = graph.globalise_records(rec, F, n_boxes)
        poses = []
        for r in range(world):
            b = batch if r == 0 else bench.build_batch(r)
            poses.append(np.array([graph.pose7_from_matrix(T) for T in b["T"]]))
        cams_wc = np.concatenate([np.tile(poses[r], (S, 1)) for r in range(world)])
        g = graph.assemble_graph(flat, cams_wc, n_landmarks)
        t_assemble = time.perf_counter() - t0
        own = ctx is None
        if own:
            ctx = csb.Context(dev)
        t0 = time.perf_counter()
        ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=g["ec"], ep=g["ep"], eo=g["eo"])
        ctx.ba_upload_estimates(g["cams7"], g["cubes10"])
        ctx.synchronize()
        t_set = time.perf_counter() - t0
        n_edges = len(g["ec"][0]) + (len(g["eo"][0]) if g["eo"] is not None else 0)
        ctx_stream_sync = ctx.synchronize
        lin_ms = {}
        for mode, name in ((False, "numeric"), (True, "analytic")):
            ctx.ba_set_jacobian_mode(mode)
            for _ in range(2):
                ctx.ba_run()
            ctx_stream_sync()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                ctx.ba_run()
            ctx_stream_sync()
            lin_ms[name] = 1e3 * (time.perf_counter() - t0) / reps
        ctx.ba_set_jacobian_mode(False)
        res = ctx.ba_download(jacobians=False)
        if own:
            ctx.close()
        n_fr = world * S * F
        out = {"config": "config#5: %d synthetic KITTI-shaped frames (64 distinct per rank, reused) x %d boxes over %d GPU(s); one allgather of %d observation records; one graph build + linearisation"
                         % (n_fr, BPF, world, world * S * n_boxes),
               "n_gpus": world, "frames": n_fr, "steps_per_rank": S, "contexts_in_flight": depth,
               "frames_stage_s": t_frames, "frames_per_s": n_fr / t_frames,
               "allgather_and_d2h_s": t_gather, "allgather_bytes_per_rank": int(S * n_boxes * 128),
               "graph": {"cameras": int(len(g["cams7"])), "landmarks": int(g["landmark_seen"].sum()), "edges_cuboid": int(len(g["ec"][0])),
                         "edges_odometry": int(len(g["eo"][0])) if g["eo"] is not None else 0, "host_assembly_s": t_assemble, "set_graph_s": t_set},
               "linearise_ms": lin_ms, "edges_per_s": {k: n_edges / (v * 1e-3) for k, v in lin_ms.items()},
               "implied_gb_per_s": {k: bench.ALGO_BYTES_PER_EDGE * n_edges / (v * 1e-3) / 1e9 for k, v in lin_ms.items()},
               "chi2": float(res["chi2"][0]),
               "total_s": t_frames + t_gather + t_assemble + t_set + lin_ms["numeric"] * 1e-3,
               "end_to_end_frames_per_s": n_fr / (t_frames + t_gather + t_assemble + t_set + lin_ms["numeric"] * 1e-3)}
        if keep:  # tests: the assembled graph and the device linearisation
            out["_graph"], out["_lin"], out["_records"] = g, res, rec
    return out


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--depth", type=int, default=6)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run(a.frames, a.depth)
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
