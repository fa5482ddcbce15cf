"""Turn ncu outputs into the committed summaries under profiles/.

  python tools/profile_summary.py launches <ncu --csv launch list> <out.csv> "<command line>"
  python tools/profile_summary.py traffic <ncu --csv launch list with dram__bytes_read.sum,dram__bytes_write.sum> <steps in the run> <config> <traffic.json>
                                                                             (DRAM bytes per step of each kernel group -> profiles/traffic.json, read by bench.py)
  python tools/profile_summary.py full <report.ncu-rep | raw.csv> <out.csv>   (one row block per kernel; a .ncu-rep needs ncu on
                                                                             PATH, a `--page raw --csv` export made on the box does not)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(src, dst, cmd):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")
        v = float(r[ci["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ci["Metric Unit"]], 1e-6)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary -- `{cmd}`\n")
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes).\n")
        f.write(f"# total kernel time {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches (warm-up + timed step + job set-up)\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k[:70]},{v[0]},{v[1]:.3f},{v[1] / tot:.4f}\n")


GROUPS = {"gj_fused": ("gj_fused_kernel",), "boot_sparse": ("boot_sparse_", "bs_weights_transpose"), "boot_mma": ("boot_mma",),
          "gom": ("gom_score", "gom_rowsums", "gom_aggregates", "gom_centroid", "gom_gather", "genome_rowsums", "gom_quantize", "gom_rowmax"),
          "shared_count_tile": ("shared_count_tile",), "lk_nn_chain": ("lk_nn_chain",)}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(src, steps, config, dst):
    import json
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ci = {h: i for i, h in enumerate(rows[start])}
    tot = collections.defaultdict(float)
    cnt = collections.defaultdict(int)
    for r in rows[start + 1:]:
        if len(r) < len(ci) or not r[ci["Metric Name"]].startswith("dram__bytes_"):
            continue
        k = r[ci["Kernel Name"]].replace("void ", "")
        v = float(r[ci["Metric Value"]].replace(",", "")) * UNIT.get(r[ci["Metric Unit"]], 1.0)
        for g, pats in GROUPS.items():
            if any(k.startswith(pt) for pt in pats):
                tot[g] += v
                if r[ci["Metric Name"]] == "dram__bytes_read.sum":
                    cnt[g] += 1
    try:
        out = json.load(open(dst))
    except Exception:
        out = {}
    for g in tot:
        out[f"{g}:{config}"] = {"dram_bytes_per_step": tot[g] / steps, "launches_per_step": cnt[g] / steps,
                                "source": f"{src.split('/')[-1]}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, {steps} step(s)"}
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)


def full(rep, dst):
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, {rep.split('/')[-1]}: selected metrics per captured launch\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            f.write(f"kernel,{d.get('Kernel Name', '?')[:90].replace(',', ';')}\n")
            f.write(f"grid,{d.get('Grid Size', '').replace(',', ' ')},block,{d.get('Block Size', '').replace(',', ' ')}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k},{d[k]},{u.get(k, '')}\n")
            f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5])
    else:
        full(sys.argv[2], sys.argv[3])
