"""Development probe: bench.py's `value` (resident, 6 contexts) and `e2e` (gray frames in pinned host memory, 6 contexts) alone, for A/B
timing of library builds (CSB_LIB=path/to/lib.so python tools/pipelined_time.py [steps])."""
import ctypes as C
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import pipeline, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
D = 6
pipeline.bind_to_gpu_numa_node(0)
params = csb.DetectParams.default()
batch = synth.make_kitti_batch(64, boxes_per_frame=8, seed=20260925)
frames, boxes, lines, tasks, n_tasks, maps, n_map = pipeline.pack_inputs(csb, batch, params)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
tb, tl, tm = pin(boxes), pin(lines), pin(maps)
boxes, lines, maps = tb.numpy(), tl.numpy(), tm.numpy()
tg = pin(np.concatenate([im.ravel() for im in batch["images"]]).astype(np.uint8)); gray = tg.numpy()
L = csb.lib()
streams = [torch.cuda.Stream() for _ in range(D)]
ctxs = [csb.Context(0, stream=s.cuda_stream) for s in streams]
for c in ctxs:
    c.detect_upload(frames, boxes, lines, tasks, n_tasks, maps, n_map, params)
obs = torch.zeros(D, boxes.shape[0] * 16, dtype=torch.float64, device="cuda")
def step_d(i):
    k = i % D
    with torch.cuda.stream(streams[k]):
        ctxs[k].detect_run(timed=False)
        assert L.csb_detect_observations_device(ctxs[k]._h, C.c_void_p(obs[k].data_ptr())) == 0
res = []
for rep in range(3):
    for i in range(2 * D):
        step_d(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]:
        s.wait_event(e0)
    for i in range(steps):
        step_d(i)
    for s in streams[1:]:
        e = torch.cuda.Event(); e.record(s); streams[0].wait_event(e)
    e1.record(streams[0])
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / steps)
print("value: ms per step %s" % " ".join("%.4f" % r for r in res))
for c in ctxs:
    c.close()
pc = [csb.Context(0) for _ in range(D)]
def run(n):
    for i in range(n):
        if i >= D:
            pc[i % D].detect_download()
        pc[i % D].detect_upload_gray(frames, boxes, lines, tasks, n_tasks, gray, params)
        pc[i % D].detect_run(timed=False)
    for i in range(max(n - D, 0), n):
        pc[i % D].detect_download()
run(15)
res = []
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    res.append(1e3 * (time.perf_counter() - t0) / steps)
print("e2e (gray, pipelined): ms per step %s" % " ".join("%.4f" % r for r in res))
