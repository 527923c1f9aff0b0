#!/usr/bin/env python
"""Turns ncu output into the markdown summaries kept under profiles/.

    python tools/ncu_summarise.py full  gpurun_out/final_chain.ncu-rep  > profiles/rNN_final_chain_ncu_summary.md
    python tools/ncu_summarise.py list  gpurun_out/launches.csv         > profiles/rNN_launches_summary.md
    python tools/ncu_summarise.py traffic gpurun_out/final_chain.ncu-rep > profiles/traffic.json

`full` reads an `ncu --set full` report (one launch per kernel), `list` the CSV of the
`--metrics gpu__time_duration.sum --clock-control none` pass, `traffic` writes the DRAM bytes of the range-filter launch
that bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import subprocess
import sys

FRAME_PX = 1920 * 1080                # the bench workload; blockIdx.z = frame, so grid z = frames per launch


def px_of(grid):
    return int(grid.strip("() ").split(",")[2]) * FRAME_PX

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def short(name):
    name = name.replace("void ", "").replace("dmc::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.split("(")[0]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def full(rep):
    head, units, rows = raw_rows(rep)
    ik, ig = head.index("Kernel Name"), head.index("Grid Size")
    names = [short(r[ik]) for r in rows]
    print("# ncu summary -- `ncu --set full --clock-control none --import-source on`, one launch per kernel\n")
    print("One launch = one frame group = %d frames of 1920x1080 (%.1f Mpx). Cold-cache, serialised by the profiler: compare shares, not absolutes.\n" % (px_of(rows[0][ig]) // FRAME_PX, px_of(rows[0][ig]) / 1e6))
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    print("| grid | " + " | ".join(r[ig] for r in rows) + " |")
    for m in METRICS:
        if m not in head: continue
        i = head.index(m)
        print("| `%s` [%s] | " % (m, units[i]) + " | ".join(r[i] for r in rows) + " |")
    i = head.index("smsp__inst_executed.sum")
    print("\nLane-instructions per pixel (smsp__inst_executed x 32 / pixels of the launch): " +
          ", ".join("%s %.0f" % (n.split("<")[0], float(r[i].replace(",", "")) * 32 / px_of(r[ig])) for n, r in zip(names, rows)) + ".")


def traffic(rep):
    head, units, rows = raw_rows(rep)
    ik = head.index("Kernel Name"); ir, iw = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum"); ig = head.index("Grid Size")
    for r in rows:
        if "bwrf8u_h2_kernel" in r[ik]:
            total = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            inst = float(r[head.index("smsp__inst_executed.sum")].replace(",", ""))
            print(json.dumps({"_doc": "dram__bytes_read.sum + dram__bytes_write.sum of one %s launch (%d frames of 1920x1080), ncu --set full" % (short(r[ik]), px_of(r[ig]) // FRAME_PX),
                              "range_filter_dram_bytes_per_launch": int(total), "frames_per_launch": px_of(r[ig]) // FRAME_PX,
                              "algorithmic_bytes_per_launch": 2 * px_of(r[ig]),
                              "lane_instructions_per_pixel": round(inst * 32 / px_of(r[ig]), 1),
                              "issue_active_pct": round(float(r[head.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]), 1),
                              "pipe_alu_pct": round(float(r[head.index("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")]), 1),
                              "pipe_fma_pct": round(float(r[head.index("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")]), 1)}, indent=1))
            return
    raise SystemExit("no range-filter launch in the report")


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    head = rows[0]; ik, iv, im = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Name")
    iu = head.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum": continue
        v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
        a = agg.setdefault(short(r[ik]), [0, 0.0]); a[0] += 1; a[1] += v
    ours = {k: v for k, v in agg.items() if any(t in k for t in ("median8u", "gauss8u", "minmax8u", "bwrf8u"))}
    tot = sum(v[1] for v in ours.values())
    print("# launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n")
    print("Per-launch times are cold-cache and serialised by the profiler: compare the SHARES with `stage_ms_per_step` / `roofline.share_of_step` of the bench line, not the absolutes.\n")
    print("| kernel | launches | mean us/launch | share of chain time |\n|---|---|---|---|")
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.3f |" % (k, n, t / n, t / tot))
    others = {k: v for k, v in agg.items() if k not in ours}
    if others:
        print("\nOther launches in the capture (input synthesis and the parity gate, outside the timed region): " +
              ", ".join("`%s` x%d" % (k[:48], v[0]) for k, v in sorted(others.items(), key=lambda kv: -kv[1][1])[:6]) + ".")


def multi(reps):
    """several single-launch reports (tools/profile_kernels.sh) side by side; pixels of a launch are not derivable from the
    grid here (different tilings), so the table keeps to the profiler's own figures"""
    cols = []
    for rep in reps:
        head, units, rows = raw_rows(rep)
        r = rows[0]; cols.append((rep.split("/")[-1].replace(".ncu-rep", "").replace("r02_", ""), head, units, r))
    print("# ncu summaries of single launches -- `ncu --set full --clock-control none --import-source on -c 1` (tools/profile_kernels.sh)\n")
    print("1080p Kinect fixture (tests/golden/kinect_desk_q50.png tiled 3x3); cold-cache, one launch each: read ratios, not absolutes.\n")
    print("| metric | " + " | ".join(c[0] for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    print("| kernel | " + " | ".join("`%s`" % short(c[3][c[1].index("Kernel Name")])[:60] for c in cols) + " |")
    print("| grid x block | " + " | ".join("%s x %s" % (c[3][c[1].index("Grid Size")], c[3][c[1].index("Block Size")]) for c in cols) + " |")
    extra = ["launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio",
             "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
    for m in METRICS + extra:
        vals = []
        for _, head, units, r in cols:
            if m in head:
                i = head.index(m); v = r[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                vals.append("%s %s" % (v, units[i]) if units[i] not in ("", "%") else v + ("%" if units[i] == "%" else ""))
            else:
                vals.append("-")
        print("| `%s` | " % m + " | ".join(vals) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "multi":
        multi(sys.argv[2:])
    else:
        {"full": full, "list": launch_list, "traffic": traffic}[sys.argv[1]](sys.argv[2])
