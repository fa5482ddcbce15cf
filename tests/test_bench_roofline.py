"""CPU: bench.py's roofline arithmetic (make_roofline) against the algorithmic work figures DESIGN.md section 4 states, written
out here independently, and against the committed bench line of the final build (profiles/r02_bench_cfg3_final.json): the
fractions a bench run prints are algorithmic work / CUDA-event time / measured peak, nothing else."""
import json
import os

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAKS = {"hbm_gbs": 6554.6, "bf16_tflops": 1644.9, "bf16_tflops_sustained": 1395.5, "sm_max_mhz": 1965.0}


def final_line():
    with open(os.path.join(ROOT, "profiles", "r02_bench_cfg3_final.json")) as fh:
        lines = [json.loads(x) for x in fh if x.startswith("{")]
    return lines[-1]


def times_of(line, steps, groups):
    out = {k: (0.0, 0) for k in ("tail", "pairs", "gom", "base")}
    for k in groups:
        e = line["roofline"]["kernels"][k]
        cnt = e["launches_per_step"] * steps
        out[k] = (e["avg_launch_ms"] * cnt, cnt)
    return out


def test_roofline_of_the_committed_bench_line_is_reproduced():
    line = final_line()
    cfg, steps = line["config"], line["steps"]
    n, p, g, nrep = cfg["n_genomes"], cfg["n_pphmms"], cfg["n_goms"], cfg["n_bootstrap"]
    assert (n, p, g, nrep) == (10000, 30000, 200, 1000)
    region_ms = line["ms_per_step"] * steps
    times = times_of(line, steps, ("tail", "pairs", "base"))
    r = bench.make_roofline(times, "sparse", 0, [0] * 8, steps, n, p, g, nrep, region_ms, PEAKS, "measured", "PG", None, "cfg3")
    pairs = n * (n + 1) // 2
    # fused tail: 2 issue slots per (pair, GOM) cell and replicate against 148 SMs x 128 lanes x clock
    sec = times["tail"][0] * 1e-3 / steps
    want = 2.0 * pairs * g * nrep / sec / (148 * 128 * 1965.0e6)
    tail = r["kernels"]["tail"]
    assert tail["bound"] == "issue" and tail["frac"] == pytest.approx(want, rel=1e-12)
    assert tail["frac"] == pytest.approx(line["roofline"]["kernels"]["tail"]["frac"], rel=1e-9)
    # its bytes: 8 B x pairs of pair sums read + 8 N^2 written + 4 N G of GOM table, per replicate
    hb = nrep * (8.0 * pairs + 8.0 * n * n + 4.0 * n * g) / sec / 1e9
    assert tail["also"]["achieved"] == pytest.approx(hb, rel=1e-12)
    assert tail["also"]["frac"] == pytest.approx(line["roofline"]["kernels"]["tail"]["also"]["frac"], rel=1e-9)
    # replicate pair sums: the fp64 workspace written, 8 B x pairs x replicates
    sec_pairs = times["pairs"][0] * 1e-3 / steps
    pr = r["kernels"]["pairs"]
    assert pr["bound"] == "hbm" and pr["achieved"] == pytest.approx(8.0 * pairs * nrep / sec_pairs / 1e9, rel=1e-12)
    assert pr["frac"] == pytest.approx(line["roofline"]["kernels"]["pairs"]["frac"], rel=1e-9)
    # the headline roofline is the group with the largest share of the timed region, and the shares are of THAT region
    assert r["kernel"] == tail["kernel"] and r["frac"] == tail["frac"]
    assert tail["share_of_step"] == pytest.approx(times["tail"][0] / region_ms, rel=1e-12)
    assert 0.5 < line["roofline"]["frac"] < 0.7 and line["roofline"]["bound"] == "issue"
    # measured DRAM traffic per launch comes from the committed ncu launch list, not from a literal
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["gj_fused:cfg3"]
    assert tail["traffic"] == pytest.approx(t["dram_bytes_per_step"] / tail["launches_per_step"], rel=1e-12)
    assert tail["traffic"] / (hb * 1e9 * sec / tail["launches_per_step"]) == pytest.approx(1.0, abs=0.05)   # no wasted re-reads


def test_roofline_tensor_core_path_counts_algorithmic_flops():
    """Dense tables: 2 x pairs x P x replicates flops against the SUSTAINED bf16 rate; the 4 digit planes the kernel issues are
    reported beside it, not as the fraction."""
    n, p, g, nrep, steps = 1000, 5000, 50, 100, 2
    times = {"tail": (0.0, 0), "gom": (0.0, 0), "base": (0.0, 0), "pairs": (10.0, 2)}
    r = bench.make_roofline(times, "mma", 0, [0] * 8, steps, n, p, g, nrep, 20.0, PEAKS, "measured", "P", None, "cfg2")
    flops = 2.0 * (n * (n + 1) // 2) * p * nrep
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert r["achieved"] == pytest.approx(flops / 5e-3 / 1e12, rel=1e-12) and r["peak"] == 1395.5
    assert r["also"]["achieved"] == pytest.approx(4 * r["achieved"], rel=1e-12)


def test_gom_roofline_formula():
    n, p, g, nrep, steps = 500, 2000, 20, 64, 1
    gstats = [1000, 5.0e6, 40000, 0, 0, 7.0e6, 0, 0]           # points, sum n_g^2, non-zeros, -, -, stored M entries
    times = {"tail": (0.0, 0), "pairs": (0.0, 0), "base": (0.0, 0), "gom": (4.0, 2)}
    r = bench.make_roofline(times, "sparse", 0, gstats, steps, n, p, g, nrep, 8.0, PEAKS, "measured", "PG", None, "none")
    dfma = nrep * (5.0e6 + 7.0e6 + 40.0 * n * g)
    assert r["bound"] == "fp64 pipe" and r["achieved"] == pytest.approx(dfma / 4e-3, rel=1e-12)
    assert r["peak"] == pytest.approx(148 * 64 * 1965.0e6) and r["traffic"] is None
