"""Turns the ncu captures of a round into the tracked summaries under profiles/ (run here, on the .ncu-rep / .csv files gpurun brought back).

  python tools/ncu_summary.py launches gpurun_out/r02_launches.csv profiles/r02_launches.md "bench.py --steps 2 --warmup 3"
  python tools/ncu_summary.py full gpurun_out/r02_detect.ncu-rep profiles/r02_ncu_detect.md
"""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(src, dst, what):
    rows = list(csv.reader(open(src, errors="replace")))
    hi = [i for i, x in enumerate(rows) if "Kernel Name" in x][0]
    h = rows[hi]
    kn, mn, mv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    d = defaultdict(lambda: [0, 0.0, 0.0])
    for x in rows[hi + 1:]:
        if len(x) <= mv:
            continue
        n = x[kn].split("(")[0].replace("void ", "").replace("csb::", "")
        v = float(x[mv].replace(",", ""))
        if x[mn] == "gpu__time_duration.sum":
            d[n][0] += 1; d[n][1] += v
        else:
            d[n][2] += v
    tot = sum(v[1] for v in d.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list: `%s`\n\n`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` "
                "(per-launch times are cold-cache and serialised: the SHARE of a kernel is what compares with bench.py's CUDA-event times).\n\n" % what)
        f.write("| kernel | launches | total us | us / launch | share | DRAM MB / launch |\n|---|---|---|---|---|---|\n")
        for n, (c, t, b) in sorted(d.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f %% | %.2f |\n" % (n, c, t / 1e3, t / 1e3 / c, 100 * t / tot, b / c / 1e6))
    print("wrote", dst)


KEYS = [("gpu__time_duration.sum", "time us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp instr"), ("sm__inst_executed.avg.per_cycle_active", "IPC"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"), ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"), ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected")]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    kn = h.index("Kernel Name")
    idx = [(h.index(k) if k in h else None, lab) for k, lab in KEYS]
    seen = {}
    for v in rows[2:]:
        n = v[kn].split("(")[0].replace("void ", "").replace("csb::", "")
        if n not in seen:
            seen[n] = v
    with open(dst, "w") as f:
        f.write("# `ncu --set full --clock-control none` summary (%s)\n\nOne launch per kernel (the first captured); stall columns = warps stalled per issue-active cycle.\n\n" % src.split("/")[-1])
        f.write("| metric | " + " | ".join("`%s`" % n for n in seen) + " |\n|---|" + "---|" * len(seen) + "\n")
        for i, lab in idx:
            if i is None:
                continue
            cells = []
            for n, v in seen.items():
                x = v[i].replace(",", "")
                try:
                    fv = float(x)
                    if "MB" in lab:
                        u = units[i]
                        fv = fv * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                    if lab == "time us" and units[i] == "ns":
                        fv /= 1e3
                    if lab == "time us" and units[i] == "ms":
                        fv *= 1e3
                    cells.append("%.4g" % fv)
                except ValueError:
                    cells.append(x)
            f.write("| %s | " % lab + " | ".join(cells) + " |\n")
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
