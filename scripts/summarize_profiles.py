"""Turn the ncu outputs brought back in gpurun_out/ into the small CSV summaries committed under profiles/.
usage: summarize_profiles.py launches <ncu --csv launch list> <out.csv>
       summarize_profiles.py raw <ncu-rep> <out.csv>            (needs ncu on PATH; key metrics per kernel)"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(src, dst):
    with open(src) as f:
        rows = csv.DictReader([l for l in f if not l.startswith("==")])
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]
            k = r["Kernel Name"].split("(")[0]
            agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "avg_us", "share_pct"])
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, v[0], f"{v[1]:.4f}", f"{1e3 * v[1] / v[0]:.2f}", f"{100 * v[1] / tot:.2f}"])
        w.writerow(["TOTAL", sum(v[0] for v in agg.values()), f"{tot:.4f}", "", "100"])


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "value", "unit"])
        for r in rows[2:]:
            for k in KEYS:
                if k in idx:
                    w.writerow([r[idx["Kernel Name"]].split("(")[0], k, r[idx[k]], units[idx[k]]])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
