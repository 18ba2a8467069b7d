"""Turn gpurun_out/*.ncu-rep / launches csv into small text summaries under profiles/ (run in the dev container)."""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def raw_summary(rep, out_name, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, out_name), "w") as f:
        f.write(f"# {title}\n# source: {os.path.basename(rep)} (ncu --set full --clock-control none), one block per profiled launch\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n== {name[:110]}\n")
            for h, u, v in zip(hdr, units, r):
                if h in KEYS or any(h.startswith(k) for k in ("sm__inst_executed_pipe_tensor",)):
                    f.write(f"{h:85s} {v:>18s} {u}\n")
            try:
                dr = float(r[hdr.index("dram__bytes_read.sum")]); dw = float(r[hdr.index("dram__bytes_write.sum")])
                ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
                f.write(f"(dram read {dr} {ur} + write {dw} {uw})\n")
            except Exception:
                pass


def launch_list(csv_path, out_name, title):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += float(r[iv].replace(",", ""))
        order.append((name, float(r[iv].replace(",", ""))))
    tot = sum(v[1] for v in agg.values())
    unit = rows[1][hdr.index("Metric Unit")]
    with open(os.path.join(OUT, out_name), "w") as f:
        f.write(f"# {title}\n# source: {os.path.basename(csv_path)}; per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write(f"# total {tot:.1f} {unit} over {len(order)} launches\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t / tot * 100:6.2f}%  {t:14.1f} {unit}  x{c:<4d} {name}\n")


if __name__ == "__main__":
    g = os.path.join(ROOT, "gpurun_out")
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    for rep, out, title in ((f"{g}/prof_mf_{tag}c.ncu-rep", f"{tag}_mf_train_ncu.txt", "mf_train_kernel<16>, ml1m shape K=5, 10 epochs"),
                            (f"{g}/prof_owner_final.ncu-rep", f"{tag}_mf_owner_ncu.txt",
                             "mf_owner_kernel<16,cached>, ml1m shape K=5, 10 epochs (70 steps)"),
                            (f"{g}/prof_sched_final.ncu-rep", f"{tag}_mf_owner_schedule_ncu.txt",
                             "owner_schedule_kernel, ml1m shape K=5, 51-epoch window"),
                            (f"{g}/prof_ot_{tag}a.ncu-rep", f"{tag}_ot_ncu.txt", "OT grouping kernels, n=1M k=32 d=64")):
        if os.path.exists(rep):
            raw_summary(rep, out, title)
    if os.path.exists(f"{g}/launches_bench.csv"):
        launch_list(f"{g}/launches_bench.csv", f"{tag}_bench_launches.txt", "python bench.py --steps 2 --warmup 3 --no-cpu (N=1)")

