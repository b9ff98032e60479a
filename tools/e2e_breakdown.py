"""Host-side breakdown of the end-to-end call (development tool; run on a GPU box).

Times upload / run / download of the maps path and the gray path separately (each followed by a synchronize), and the one-call
csb_detect_batch / csb_detect_batch_gray entries, and prints the per-kernel event times of the gray path.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import cube_slam_wu_b200 as csb
    from cube_slam_wu_b200 import pipeline, synth

    torch.cuda.set_device(0)
    ctx = csb.Context(0)
    params = csb.DetectParams.default()
    batch = synth.make_kitti_batch(64, boxes_per_frame=8, seed=20260925)
    frames, boxes, lines, tasks, n_tasks, maps, n_map = pipeline.pack_inputs(csb, batch, params)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    tb, boxes = pinned(boxes); tl, lines = pinned(lines); tm, maps = pinned(maps)
    tg, gray = pinned(np.concatenate([im.ravel() for im in batch["images"]]).astype(np.uint8))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timeit(fn, n=20):
        ts = []
        for i in range(n + 3):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter(); fn(); ctx.synchronize() if hasattr(ctx, "synchronize") else torch.cuda.synchronize()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(time.perf_counter() - t0)
        return 1e3 * float(np.median(ts))

    print("plan only            %.3f ms" % timeit(lambda: csb.detect_plan(frames, boxes, params) if hasattr(csb, "detect_plan") else None))
    print("maps: upload         %.3f ms" % timeit(lambda: ctx.detect_upload(frames, boxes, lines, tasks, n_tasks, maps, n_map, params)))
    print("maps: run            %.3f ms" % timeit(lambda: ctx.detect_run(timed=True)))
    print("maps: download       %.3f ms" % timeit(lambda: ctx.detect_download()))
    print("maps: batch          %.3f ms" % timeit(lambda: ctx.detect_batch(frames, boxes, lines, tasks, n_tasks, maps, n_map, params, want_stats=True)))
    print("maps: batch nostats  %.3f ms" % timeit(lambda: ctx.detect_batch(frames, boxes, lines, tasks, n_tasks, maps, n_map, params, want_stats=False)))
    print("gray: upload         %.3f ms" % timeit(lambda: ctx.detect_upload_gray(frames, boxes, lines, tasks, n_tasks, gray, params)))
    print("gray: run            %.3f ms" % timeit(lambda: ctx.detect_run(timed=True)))
    print("gray: download       %.3f ms" % timeit(lambda: ctx.detect_download()))
    _, _, st = ctx.detect_download()
    print("gray kernels: distmap %.3f prep %.3f score %.3f select %.3f recover %.3f rank %.3f  scored %d smem-tasks %d" % (
        st.gpu_ms_distmap, st.gpu_ms_prep, st.gpu_ms_score, st.gpu_ms_select, st.gpu_ms_recover, st.gpu_ms_rank, st.n_scored, st.n_tasks_smem_map))
    print("gray: batch          %.3f ms" % timeit(lambda: ctx.detect_batch_gray(frames, boxes, lines, tasks, n_tasks, gray, params, want_stats=True)))
    print("gray: batch nostats  %.3f ms" % timeit(lambda: ctx.detect_batch_gray(frames, boxes, lines, tasks, n_tasks, gray, params, want_stats=False)))
    # pure copies for scale
    dm = torch.empty(maps.size, dtype=torch.float32, device="cuda")
    dg = torch.empty(gray.size, dtype=torch.uint8, device="cuda")
    print("H2D maps alone       %.3f ms (%d MB)" % (timeit(lambda: dm.copy_(tm, non_blocking=True)), maps.nbytes >> 20))
    print("H2D gray alone       %.3f ms (%d MB)" % (timeit(lambda: dg.copy_(tg, non_blocking=True)), gray.nbytes >> 20))


if __name__ == "__main__":
    main()
