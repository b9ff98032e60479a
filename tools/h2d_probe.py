"""Development probe: where does the host->device time of csb_detect_upload go?  (run on a GPU box)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import pipeline, synth

torch.cuda.set_device(0)
s = torch.cuda.Stream()
ctx = csb.Context(0, stream=s.cuda_stream)
params = csb.DetectParams.default()
batch = synth.make_kitti_batch(64, boxes_per_frame=8, seed=20260925)
frames, boxes, lines, tasks, n_tasks, maps, n_map = pipeline.pack_inputs(csb, batch, params)
tm = torch.from_numpy(np.ascontiguousarray(maps)).pin_memory(); maps = tm.numpy()
dm = torch.empty(maps.size, dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def ev_time(fn, n=10, do_flush=True):
    out = []
    for i in range(n + 3):
        if do_flush:
            flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(s); fn(); b.record(s)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if i >= 3:
            out.append((a.elapsed_time(b), 1e3 * (t1 - t0)))
    o = np.median(np.array(out), axis=0)
    return "gpu %.3f ms  host %.3f ms" % (o[0], o[1])

def torch_copy(chunks):
    n = maps.size
    with torch.cuda.stream(s):
        for k in range(chunks):
            a, b = n * k // chunks, n * (k + 1) // chunks
            dm[a:b].copy_(tm[a:b], non_blocking=True)

for fl in (True, False):
    print("flush", fl)
    print("  torch copy 1 chunk   ", ev_time(lambda: torch_copy(1), do_flush=fl))
    print("  torch copy 16 chunks ", ev_time(lambda: torch_copy(16), do_flush=fl))
    print("  csb upload (maps)    ", ev_time(lambda: ctx.detect_upload(frames, boxes, lines, tasks, n_tasks, maps, n_map, params), do_flush=fl))
    print("  csb batch            ", ev_time(lambda: ctx.detect_batch(frames, boxes, lines, tasks, n_tasks, maps, n_map, params, want_stats=False), do_flush=fl))
