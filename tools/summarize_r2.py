"""profiles/r2_*: small text summaries of the round-2 ncu captures in gpurun_out/ (run in the dev container).
    python tools/summarize_r2.py"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, OUT = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarize_rep(name, title, command, out=None, mode="w"):
    rep = os.path.join(GO, name + ".ncu-rep")
    if not os.path.exists(rep):
        return
    hdr, units, rows = raw(rep)
    ki = hdr.index("Kernel Name")
    import re
    out = out or re.sub(r"^r2[a-z0-9]*_", "r2_", name) + "_ncu.txt"
    with open(os.path.join(OUT, out), mode) as f:
        f.write(f"# {title}\n# {command}\n# ncu --set full --clock-control none (cold caches, kernel replay): shares and ratios, not bench times\n")
        for r in rows:
            f.write(f"\n== {r[ki][:150]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:88s} {units[i]:14s} {r[i]}\n")


def summarize_launches(name, title):
    path = os.path.join(GO, name)
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    r = list(csv.reader(lines))
    hdr = r[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.OrderedDict()
    for x in r[1:]:
        try:
            v = float(x[vi].replace(",", ""))
        except Exception:
            continue
        t = tot.setdefault(x[ki][:110], [0, 0.0])
        t[0] += 1
        t[1] += v
    S = sum(v for _, v in tot.values())
    with open(os.path.join(OUT, "r2_bench_launches.txt"), "w") as f:
        f.write(f"# {title}\n# per-launch times are cold-cache and serialised: compare SHARES\n# total {S:.0f} ns over "
                f"{sum(c for c, _ in tot.values())} launches\n")
        for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * v / S:6.2f}% {v / c / 1000:10.1f} us/launch x{c:4d}  {k}\n")


if __name__ == "__main__":
    summarize_launches(sys.argv[1] if len(sys.argv) > 1 else "r2b_launches_bench.csv",
                       "python bench.py --steps 2 --warmup 3 --no-cpu --no-extra (N=1), ncu --metrics gpu__time_duration.sum")
    summarize_rep("r2_runs_c4small", "mf_runs_kernel<128>, one GPU's share of C4 (8 shards, d=128), 40 steps",
                  "python tools/run_config.py --config c4small --mode runs --max-steps 40")
    summarize_rep("r2_lazy_c4small", "mf_train_lazy_kernel<128> (round-1 schedule), same workload, 40 steps",
                  "python tools/run_config.py --config c4small --mode lazy --max-steps 40")
    summarize_rep("r2c_sched_fast", "owner_schedule_fast_kernel (ballot-loop binning: REJECTED, slower than match.any), C2",
                  "python bench.py --steps 1 --warmup 3 --no-cpu --no-extra")
    summarize_rep("r2e_cost", "cost_tma_kernel<64,1> (serial per-tile chain), n=1M k=8 d=64",
                  "python tools/prof_ot.py --n 1000000 --k 8 --d 64 --reps 1 --iters 2")
    summarize_rep("r2d_ot32", "cost_tma_kernel + colsum_kernel<32> BEFORE the full-wave grid fix, n=1M k=32 d=64",
                  "python tools/prof_ot.py --n 1000000 --k 32 --d 64 --reps 1 --iters 2")
    summarize_rep("r2f_ot", "cost_ws_kernel<64> (FIRST warp-specialised version: shared-memory atomics in the lo pass) + colsum_kernel<8>, n=1M k=8 d=64",
                  "python tools/prof_ot.py --n 1000000 --k 8 --d 64 --reps 1 --iters 2")
    summarize_rep("r2i_sched", "owner_schedule_kernel (general pre-pass, start of the session), C2 step",
                  "python bench.py --steps 1 --warmup 3 --no-cpu --no-extra", out="r2_sched_ncu.txt")
    summarize_rep("r2k_sched", "owner_schedule_tab_kernel<4> (short-epoch pre-pass), C2 step",
                  "python bench.py --steps 1 --warmup 3 --no-cpu --no-extra", out="r2_sched_ncu.txt", mode="a")
    summarize_rep("r2z_cost", "cost_ws_kernel<64> (final), n=10M k=8 d=64",
                  "python tools/prof_ot.py --n 10000000 --k 8 --d 64 --reps 1 --iters 2", out="r2_cost_final_ncu.txt")
    summarize_rep("r2z_owner", "mf_owner_kernel<16> + owner_schedule_tab_kernel<4> (final), C2 step",
                  "python bench.py --steps 1 --warmup 3 --no-cpu --no-extra", out="r2_owner_final_ncu.txt")
