"""Host-side drivers above the C ABI that bench.py, tools/ and the tests share: input packing for the batch entry points and BASELINE
config #5 (frames sharded over the ranks -> one allgather of the observation records -> graph -> linearisation).

Nothing here computes on the path: the frames go through csb_detect_upload_gray / csb_detect_run / csb_detect_observations_device, the graph
through csb_ba_set_graph / csb_ba_run; numpy only packs buffers and assembles the graph (cube_slam_wu_b200/graph.py)."""
import ctypes as C
import os
import time

import numpy as np

from . import graph, synth

FRAMES_PER_GPU = 64      # BASELINE config #2 / #5: frames per step and rank
BOXES_PER_FRAME = 8
ALGO_BYTES_PER_EDGE = 1856.0   # EdgeSE3Cuboid, fused (Jacobian not materialised)


def pack_inputs(csb, batch, params, with_maps=True):
    """synth.make_kitti_batch() dict -> (frames, boxes, lines, tasks, n_tasks, maps, n_map_floats) as the csb_detect_* entry points take them.
    maps: the caller-computed distance maps (cv2, exactly the reference's calls box_proposal_detail.cpp:320-327) or None."""
    frames = csb.make_frames(batch["K"], batch["T"], batch["img_w"], batch["img_h"], batch["box_ranges"], batch["line_ranges"])
    boxes = np.ascontiguousarray(batch["boxes"], np.float64).reshape(-1, 5)
    lines = np.ascontiguousarray(batch["lines"], np.float64).reshape(-1, 4)
    tasks, n_tasks, n_map = csb.detect_plan(frames, boxes, params)
    maps = synth.dist_maps_for_tasks(batch["images"], tasks, n_tasks, n_map) if with_maps else None
    return frames, boxes, lines, tasks, n_tasks, maps, n_map


def gray_of(batch):
    return np.ascontiguousarray(np.concatenate([im.ravel() for im in batch["images"]]).astype(np.uint8))


def pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def bind_to_gpu_numa_node(device_index):
    """Run this process (and so first-touch its pinned staging buffers) on the CPUs NVML lists as local to the GPU.  Returns the CPU list or
    None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def perturb_poses(cams_wc7, sigma_t, sigma_rot, seed):
    """Camera-pose ESTIMATES for the graph: the true pose times a small random motion (translation sigma_t metres, rotation sigma_rot
    radians per axis), what odometry drift leaves an online caller with; pose 0 (the fixed vertex) stays exact."""
    rng = np.random.default_rng(seed)
    n = len(cams_wc7)
    dt = rng.normal(0.0, sigma_t, (n, 3))
    w = rng.normal(0.0, sigma_rot, (n, 3))
    ang = np.linalg.norm(w, axis=1, keepdims=True)
    ax = w / np.maximum(ang, 1e-300)
    dq = np.concatenate([ax * np.sin(0.5 * ang), np.cos(0.5 * ang)], axis=1)
    d = np.concatenate([dt, dq], axis=1)
    d[0] = [0, 0, 0, 0, 0, 0, 1]
    return graph.se3_mul(np.asarray(cams_wc7, np.float64), d)


def run_config5(n_frames_total=10000, depth=6, ctx=None, keep=False, pose_noise=(0.02, 0.005), seed=20260925):
    """BASELINE config #5.  Every rank pushes its share of the frames through the gray-frame entry (pinned host buffers, `depth` contexts in
    flight), appends the observation records of every step to ONE device buffer, the ranks exchange that buffer with ONE all_gather (NCCL)
    after the last step, and rank 0 assembles the camera-object graph (graph.py: the recipe of main_obj.cpp:738-803), hands it to
    csb_ba_set_graph and linearises it (replicas only: SURVEY.md 8e).

    The frames are the bench's 64 distinct synthetic frames per rank, pushed again pass after pass; every 2D box of a rank's frames is its
    own landmark (data association is not part of the reference: its data set has one object), observed once per pass.  The graph's camera
    vertices start from ESTIMATES -- the true pose of the frame times a random motion that differs from pass to pass (`pose_noise`: sigma of
    the translation in metres and of the rotation in radians; the first camera is fixed and exact) -- while the odometry edges carry the true
    relative motion, so the re-observations of a landmark disagree, chi2 > 0 and the right-hand side is not zero.
    Rank 0 returns the result dict, the other ranks None."""
    import torch
    import torch.distributed as dist
    import cube_slam_wu_b200 as csb

    ddp = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if ddp else 0
    world = dist.get_world_size() if ddp else 1
    dev = torch.cuda.current_device()
    params = csb.DetectParams.default()
    F, BPF = FRAMES_PER_GPU, BOXES_PER_FRAME
    batch = synth.make_kitti_batch(F, boxes_per_frame=BPF, seed=seed + rank)
    frames, boxes, lines, tasks, n_tasks, _, _ = pack_inputs(csb, batch, params, with_maps=False)
    tb, boxes = pinned(boxes); tl, lines = pinned(lines)
    tg, gray = pinned(gray_of(batch))
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
    # every rank's camera poses (input metadata of the synthetic frames: 64 x 7 doubles) go to every rank up front; the exchange also
    # brings up the NCCL communicator, so that the timed all_gather below measures the collective and not its lazy initialisation
    poses_mine = torch.from_numpy(np.array([graph.pose7_from_matrix(T) for T in batch["T"]])).cuda()
    if world > 1:
        poses_all_t = torch.empty(world, F, 7, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(poses_all_t.view(-1), poses_mine.view(-1))
        # the receive buffer of the record exchange exists before the timed collective, and one exchange of that size has run on it
        # (NCCL sets its channels / protocol up per message size on first use: a one-off of the communicator, not of the collective)
        allobs = torch.empty(world, S * n_boxes * 16, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allobs.view(-1), obs.view(-1))
        dist.barrier()
    else:
        poses_all_t = poses_mine.view(1, F, 7)
    poses_all = poses_all_t.cpu().numpy()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    frames_stage(S, obs)
    t_frames = time.perf_counter() - t0
    # the one collective of the path: every rank's records to every rank (graph assembly is replicated or, as here, done by rank 0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if world > 1:
        dist.all_gather_into_tensor(allobs.view(-1), obs.view(-1))
    else:
        allobs = obs.view(1, -1)
    ev1.record()
    torch.cuda.synchronize()
    t_allgather = ev0.elapsed_time(ev1) * 1e-3
    t0 = time.perf_counter()
    rec = allobs.cpu().numpy().reshape(world, S, n_boxes, 16).copy()
    t_d2h = time.perf_counter() - t0
    tt = torch.tensor([t_frames, t_allgather, t_d2h], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_frames, t_allgather, t_d2h = float(tt[0]), float(tt[1]), float(tt[2])
    for c in ctxs:
        c.close()
    out = None
    if rank == 0:
        t0 = time.perf_counter()
        flat, n_landmarks = graph.globalise_records(rec, F, n_boxes)
        cams_true = np.concatenate([np.tile(poses_all[r], (S, 1)) for r in range(world)])
        cams_est = perturb_poses(cams_true, pose_noise[0], pose_noise[1], seed + 77) if pose_noise else None
        g = graph.assemble_graph(flat, cams_true, n_landmarks, cams_est_wc7=cams_est)
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
        lin_ms, lin = {}, {}
        for mode, name in ((False, "numeric"), (True, "analytic")):
            ctx.ba_set_jacobian_mode(mode)
            for _ in range(2):
                ctx.ba_run()
            ctx.synchronize()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                ctx.ba_run()
            ctx.synchronize()
            lin_ms[name] = 1e3 * (time.perf_counter() - t0) / reps
            if keep or name == "numeric":
                lin[name] = ctx.ba_download(jacobians=False)
        ctx.ba_set_jacobian_mode(False)
        res = lin["numeric"]
        if own:
            ctx.close()
        n_fr = world * S * F
        t_total = t_frames + t_allgather + t_d2h + t_assemble + t_set + lin_ms["numeric"] * 1e-3
        out = {"config": "config#5: %d synthetic KITTI-shaped frames (64 distinct per rank, pushed pass after pass) x %d boxes over %d GPU(s); ONE allgather of %d observation records; "
                         "one graph build + linearisation (camera estimates: true pose x random motion, sigma %g m / %g rad)"
                         % (n_fr, BPF, world, world * S * n_boxes, pose_noise[0] if pose_noise else 0, pose_noise[1] if pose_noise else 0),
               "n_gpus": world, "frames": n_fr, "steps_per_rank": S, "contexts_in_flight": depth,
               "frames_stage_s": t_frames, "frames_per_s": n_fr / t_frames,
               "allgather_s": t_allgather, "allgather_bytes_per_rank": int(S * n_boxes * 128), "allgather_bytes_total": int(world * S * n_boxes * 128),
               "records_d2h_s": t_d2h,
               "graph": {"cameras": int(len(g["cams7"])), "landmarks": int(g["landmark_seen"].sum()), "edges_cuboid": int(len(g["ec"][0])),
                         "edges_odometry": int(len(g["eo"][0])) if g["eo"] is not None else 0, "host_assembly_s": t_assemble, "set_graph_s": t_set},
               "linearise_ms": lin_ms, "edges_per_s": {k: n_edges / (v * 1e-3) for k, v in lin_ms.items()},
               "implied_gb_per_s": {k: ALGO_BYTES_PER_EDGE * n_edges / (v * 1e-3) / 1e9 for k, v in lin_ms.items()},
               "chi2": float(res["chi2"][0]), "b_cam_max": float(np.abs(res["b_cam"]).max()), "b_cube_max": float(np.abs(res["b_cube"]).max()),
               "total_s": t_total, "end_to_end_frames_per_s": n_fr / t_total}
        if keep:  # tests: the assembled graph and the device linearisations
            out["_graph"], out["_lin"], out["_lin_analytic"], out["_records"] = g, res, lin.get("analytic"), rec
    return out
