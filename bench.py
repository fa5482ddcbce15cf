#!/usr/bin/env python
"""Benchmark of the pairwise genome-similarity hot path (BASELINE.json metric:
genome-pair similarities/sec incl. bootstraps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2|cfg3|cfg4|cfg5] [--density D] [--scheme P|PG]
                    [--impl reference]

Workloads (BASELINE.json configs):
  cfg3 (default) 10 000 genomes x 30 000 PPHMMs, G = 200, scheme PG, base matrix + 1000 PL1 bootstrap replicates
                 (GOM table rebuilt per replicate, app/src/pl1_graphs.py:78-107), density 1 %
  cfg2           1 000 x 5 000, G = 50, base + 100 replicates
  cfg5           pipeline_ii update: 20 000 reference + 5 000 query genomes x 30 000 PPHMMs, 200 GOMs from the reference
                 rows, sorted resample indices, real location table (app/src/virus_classification.py:290-315), base + 10
  cfg4           presence/absence popcount path, 50 000 x 30 000 (app/utils/shared_pphmm_graphs.py:59-97), no bootstrap
  --density 1.0  the dense worst case (cfg2/cfg3): the replicate pair sums then run on the tensor cores (tcgen05);
                 scheme P unless --scheme says otherwise (a dense location table is outside the GOM kernel's range)
One "step" = the whole job.  pairs/s = N(N+1)/2 * (1+B) / time.

  value : inputs resident in HBM, outputs written to HBM (device-timed, CUDA events on the stream the kernels run
          on, max over ranks)
  e2e   : the same job through the public host API with numpy tables in host memory, H2D and D2H inside the timed
          region.  For the bootstrap-loop workloads the product of a replicate is its dendrogram
          (app/src/pl1_graphs.py:108-117 keeps nothing else), so e2e delivers one Newick string per replicate
          (gravity2_b200.engine.Job.for_each_tree: UPGMA on the device); for the pipeline II workload (cfg5) the product is the
          classification epilogue's input -- sampled class cells and query-row maxima of every similarity matrix
          (Job.for_each_replicate_device: matrices stay in HBM); `e2e_matrices` is the variant that ships every matrix to
          host memory.
  parity_check : cells of the LAST timed step's device results (base matrix, replicate matrices still in the device
          ring, their GOM tables) against the CPU oracle's single-cell restatement of the reference
          (oracle/sampled_check.py); a mismatch makes the run fail (exit code 3).
  --impl reference : the reference's numpy pair expressions and dcor-based GOM scoring (CPU oracle port of
          app/utils/similarity_matrix_constructor.py:167-177 and parallel_gom_sig_generator.py:30-37) on all host
          cores, on a bounded row sample that keeps the workload's genomes-per-taxon ratio: exactly K timed passes
          after W warm-up passes, the sample's row count sized by a calibration pass so that the run ends within ~5 min.
  roofline.traffic : measured DRAM bytes per launch (profiles/traffic.json: ncu dram__bytes_* of the same bench command).
  N > 1 (torchrun): `value` from one rank per GPU; `e2e` from rank 0 driving all GPUs through gravity2_b200.MultiEngine.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from gravity2_b200 import synth  # noqa: E402

METRIC = "genome_pair_similarities_per_sec_incl_bootstraps"
UNIT = "pair-similarities/s"


def load_peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


def load_traffic():
    """Measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel group over one step of a config, from
    the committed ncu launch list of the same bench command (profiles/traffic.json, written by `tools/profile_summary.py
    traffic`).  .get(key, launches_per_step) gives bytes per launch, like `achieved`."""
    try:
        raw = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
    except Exception:
        raw = {}

    class _T:
        def get(self, key, launches_per_step=None):
            e = raw.get(key)
            if not e:
                return None
            return e["dram_bytes_per_step"] / (launches_per_step or e.get("launches_per_step") or 1.0)
    return _T()


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [s for s, p in zip(sm, pw) if p > 250.0] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- workloads
def build_workload(args):
    cfg = dict(synth.CONFIGS[args.config])
    if args.boot is not None:
        cfg["b"] = args.boot
    if args.config == "cfg5" and args.boot is None:
        cfg["b"] = 10
    if args.density is not None:
        cfg["density"] = args.density
    cfg["pl2"] = args.config == "cfg5"
    cfg["n_ref"] = 20000 if cfg["pl2"] else cfg["n"]
    cfg["scheme"] = args.scheme or ("P" if cfg["density"] >= 0.5 and cfg["p"] > 10000 else "PG")
    return cfg


def make_tables(args, cfg):
    seed = synth.SEED_BASE + int(args.config[-1])
    if cfg["density"] >= 0.5:
        return synth.make_tables(cfg["n"], cfg["p"], cfg["g"], cfg["density"], seed)
    if cfg["n"] * cfg["p"] > 2e7:          # hit-by-hit draw of the same distribution: seconds instead of minutes
        return synth.make_tables_sparse(cfg["n"], cfg["p"], cfg["g"], cfg["density"], seed, with_loc=args.config != "cfg4")
    return synth.make_tables(cfg["n"], cfg["p"], cfg["g"], cfg["density"], seed)


def workload_config(args, cfg, world):
    if args.config == "cfg4":
        return {"workload": f"cfg4: presence/absence popcount path, {cfg['n']} genomes x {cfg['p']} PPHMMs bit-packed, all-pairs shared-PPHMM counts",
                "n_genomes": cfg["n"], "n_pphmms": cfg["p"], "density": cfg["density"],
                "l2_policy": "the N x N int32 output (10 GB) exceeds L2; the packed table (190 MB) is L2-resident by design",
                "sharding": f"{world} rank(s): upper-triangular tile blocks, packed blocks all-gathered over NCCL"}
    loop = "PL2 loop body, sorted indices, real location table, GOMDB from the 20000 reference rows" if cfg["pl2"] \
        else "PL1 loop body incl. GOM rebuild"
    return {"workload": f"{args.config}: synthetic {cfg['n']} genomes x {cfg['p']} PPHMMs, G={cfg['g']} GOMs, scheme {cfg['scheme']}, "
                        f"base + {cfg['b']} bootstrap replicates ({loop})",
            "n_genomes": cfg["n"], "n_pphmms": cfg["p"], "n_goms": cfg["g"], "n_bootstrap": cfg["b"],
            "density": cfg["density"], "scheme": cfg["scheme"], "resample": "reference MT19937 indices (host) -> int32 weights",
            "l2_policy": "working set >> L2: the per-replicate pair-sum workspaces and output ring exceed 126 MB",
            "sharding": f"{world} rank(s): base matrix by upper-triangular tile blocks + NCCL all-gather; replicates by index"}


# --------------------------------------------------------------------------- reference arm
_G = {}     # tables shared with forked pool workers (set before the pool is created)


def _ref_rows(args):
    """rows [i0, i1) of the pair loop, similarity_matrix_constructor.py:113-177 (P and G expressions)."""
    i0, i1 = args
    sig, gom = _G["sig"], _G["gom"]
    n = sig.shape[0]
    outp = np.zeros((i1 - i0, n))
    outg = np.zeros((i1 - i0, n))
    for i in range(i0, i1):
        for j in range(i, n):
            a, b = sig[i], sig[j]
            mx = np.sum(np.maximum(a, b))
            outp[i - i0, j] = np.sum(np.minimum(a, b)) / mx if not mx == 0 else 0
            if gom is not None:
                a, b = gom[i], gom[j]
                mx = np.sum(np.maximum(a, b))
                outg[i - i0, j] = np.sum(np.minimum(a, b)) / mx if not mx == 0 else 0
    return outp, outg


def _ref_gom(task):
    """generate_gom_sigs of one GOM (parallel_gom_sig_generator.py:10-16) for genome rows [i0, i1)"""
    g, i0, i1 = task
    from oracle import gravity_oracle as O
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.generate_gom_sigs(_G["db"][g], _G["loc"][i0:i1])


def _ref_popc_rows(args):
    """shared_pphmm_ratio's inner loops (app/utils/shared_pphmm_graphs.py:63-76) for rows [i0, i1): Python, as shipped"""
    i0, i1 = args
    arr = _G["arr"]
    out = []
    for a in arr[i0:i1]:
        row = []
        for b in arr:
            hits = 0
            for x, y in zip(a, b):
                if x != 0 and y != 0:
                    hits += 1
            row.append(hits)
        out.append(row)
    return out


def reference_job(cores, sig, loc, taxo_ref, gom_ids, idx_list, scheme, pl2):
    """Base matrix + one replicate per idx in idx_list on the CPU oracle port, all host cores;
    returns pairs*(1+B).  A fresh fork pool per stage, as the reference spawns a Pool per GOM table
    (parallel_gom_sig_generator.py:18)."""
    import multiprocessing as mp
    from oracle import gravity_oracle as O
    n = sig.shape[0]
    n_ref = len(taxo_ref)
    ctxmp = mp.get_context("fork")

    def one(sig_t, loc_t):
        gom = None
        if "G" in scheme:
            _G.update(loc=loc_t, db=O.gomdb(taxo_ref, loc_t[:n_ref], gom_ids))
            # The reference's pool works GOM by GOM (parallel_gom_sig_generator.py:18-24): at the workload's G = 200 that keeps
            # every core busy, in a bounded sample with a handful of GOMs it would not -- so the sample's genome rows are cut
            # in chunks as well (same calls on the same rows, all cores busy as at full scale).
            nch = max(1, -(-2 * cores // max(1, len(gom_ids))))
            rb = np.unique(np.linspace(0, n, nch + 1).astype(int))
            tasks = [(g, int(rb[k]), int(rb[k + 1])) for g in gom_ids for k in range(len(rb) - 1)]
            with ctxmp.Pool(cores) as pool:
                parts = pool.map(_ref_gom, tasks, chunksize=1)
            per = len(rb) - 1
            cols = [np.concatenate([np.asarray(parts[gi * per + k]).reshape(-1) for k in range(per)]) for gi in range(len(gom_ids))]
            gom = np.column_stack(cols) if cols else np.zeros((n, 0))
        _G.update(sig=sig_t, gom=gom)
        bounds = np.unique(np.linspace(0, n, 8 * cores + 1).astype(int))     # balance the triangle: many small row blocks
        with ctxmp.Pool(cores) as pool:
            parts = pool.map(_ref_rows, [(bounds[k], bounds[k + 1]) for k in range(len(bounds) - 1)], chunksize=1)
        P = np.vstack([p for p, _ in parts]); Gm = np.vstack([g for _, g in parts])
        P = np.triu(P) + np.triu(P, 1).T; Gm = np.triu(Gm) + np.triu(Gm, 1).T
        S = O._combine(P, Gm, np.zeros((n, n)), None, scheme, 1.0)
        return O.dist_from_sim(S)

    one(sig, loc)
    for idx in idx_list:
        bs = np.ascontiguousarray(sig[:, idx])
        one(bs, np.ascontiguousarray(loc[:, idx]) if pl2 else bs)     # PL1: signature table doubles as location table (Q3)
    return n * (n + 1) // 2 * (1 + len(idx_list))


def reference_sample(args, cfg):
    """Bounded sample of the workload for the CPU legs: ref_rows genomes at full P, with as many taxa as keep the
    workload's genomes-per-taxon ratio (members per GOM set both the dcor point count and its dimension)."""
    n_s = min(cfg["n"], args.ref_rows)
    g_s = max(1, min(cfg["g"], int(round(n_s * cfg["g"] / cfg["n"]))))
    seed = synth.SEED_BASE + int(args.config[-1])
    gen = synth.make_tables if cfg["density"] >= 0.5 or n_s * cfg["p"] <= 2e7 else synth.make_tables_sparse
    t = gen(n_s, cfg["p"], g_s, cfg["density"], seed)
    n_ref_s = max(g_s, int(round(n_s * cfg["n_ref"] / cfg["n"])))
    return t, n_s, g_s, n_ref_s


def time_reference(args, cfg, n_boot, steps, warm):
    cores = os.cpu_count() or 1
    if args.config == "cfg4":
        import multiprocessing as mp
        n_s = min(cfg["n"], args.ref_rows_popc)
        t = synth.make_tables_sparse(n_s, cfg["p"], cfg["g"], cfg["density"], synth.SEED_BASE + 4, with_loc=False)
        _G["arr"] = t.sig
        bounds = np.unique(np.linspace(0, n_s, 4 * cores + 1).astype(int))
        for _ in range(int(warm)):
            with mp.get_context("fork").Pool(cores) as pool:
                pool.map(_ref_popc_rows, [(bounds[k], bounds[k + 1]) for k in range(len(bounds) - 1)], chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            with mp.get_context("fork").Pool(cores) as pool:
                pool.map(_ref_popc_rows, [(bounds[k], bounds[k + 1]) for k in range(len(bounds) - 1)], chunksize=1)
        dt = (time.perf_counter() - t0) / steps
        units = n_s * (n_s + 1) // 2      # the reference visits all N^2 ordered pairs; the metric counts unordered ones
        return units / dt, dt, cores, (f"{n_s} of {cfg['n']} genomes at full P={cfg['p']}: shared_pphmm_ratio's triple Python loop "
                                       f"(shared_pphmm_graphs.py:63-76) over a {cores}-process pool; {dt:.1f} s per pass")
    t, n_s, g_s, n_ref_s = reference_sample(args, cfg)
    idx = synth.reference_resample_indices(cfg["p"], n_boot, seed=args.seed, sort=cfg["pl2"])
    for _ in range(int(warm)):
        reference_job(cores, t.sig, t.loc, t.taxo[:n_ref_s], t.gom_ids, idx, cfg["scheme"], cfg["pl2"])
    t0 = time.perf_counter()
    units = 0
    for _ in range(steps):
        units += reference_job(cores, t.sig, t.loc, t.taxo[:n_ref_s], t.gom_ids, idx, cfg["scheme"], cfg["pl2"])
    dt = time.perf_counter() - t0
    sample = (f"{n_s} of {cfg['n']} genomes at full P={cfg['p']}, G={g_s} (same genomes-per-taxon ratio as the workload), "
              f"base + {n_boot} of {cfg['b']} replicates ({'PL2' if cfg['pl2'] else 'PL1'} body); numpy pair expressions of "
              f"similarity_matrix_constructor.py:167-177 + dcor GOM scoring (tasks = GOM x genome chunk, so that every core works as at "
              f"G = {cfg['g']}) over a {cores}-process pool; {dt / steps:.1f} s per step")
    return units / dt, dt / steps, cores, sample


def run_reference(args, cfg):
    """`--impl reference`: EXACTLY K timed steps after W warm-up steps, each one pass of the reference's CPU path over a bounded
    sample of the workload; the sample's row count is sized from a calibration pass so that the K + W passes end within
    REF_BUDGET_S seconds whatever the host (the cost of a pass grows with the square of the rows: pairs, and genome x GOM
    scores at a fixed genomes-per-taxon ratio)."""
    steps = max(1, args.steps if args.steps is not None else 2)
    warm = max(0, args.warmup if args.warmup is not None else 1)
    budget = float(os.environ.get("GV2_REF_BUDGET_S", "280")) / (steps + warm)
    key = "ref_rows_popc" if args.config == "cfg4" else "ref_rows"
    full, floor = getattr(args, key), (16 if args.config == "cfg4" else 32)
    cal_rows = min(full, 48 if args.config != "cfg4" else 24)
    setattr(args, key, cal_rows)
    _, cal_dt, _, _ = time_reference(args, cfg, min(cfg["b"], args.ref_boot), 1, 0)
    rows = int(cal_rows * (max(budget - 0.3, 0.05) / max(cal_dt - 0.3, 0.02)) ** 0.5)      # ~0.3 s of pool start-up per pass
    rows = max(floor, min(full, min(cfg["n"], rows)))
    setattr(args, key, rows)
    val, dt, cores, sample = time_reference(args, cfg, min(cfg["b"], args.ref_boot), steps, warm)
    sample += f"; rows sized for {steps} + {warm} passes in ~{budget * (steps + warm):.0f} s"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def measure_cpu_baseline(args, cfg):
    val, dt, cores, sample = time_reference(args, cfg, min(cfg["b"], 1), 1, 0)
    out = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if args.config != "cfg4" and args.as_shipped_rows > 0:
        out["as_shipped"] = measure_as_shipped(args, cfg)
    return out


def measure_as_shipped(args, cfg):
    """the reference as it runs today -- ONE process, the literal pair loop
    (similarity_matrix_constructor.py:112-226 restated in oracle.similarity_mat) and the GOM table column by column --
    on a small row sample; reported beside the all-core figure, not used as the comparator."""
    import warnings
    from oracle import gravity_oracle as O
    n_s = min(cfg["n"], args.as_shipped_rows)
    g_s = max(1, min(cfg["g"], int(round(n_s * cfg["g"] / cfg["n"]))))
    t = synth.make_tables_sparse(n_s, cfg["p"], g_s, cfg["density"], synth.SEED_BASE + 9)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gom = O.gom_table(t.loc, O.gomdb(t.taxo, t.loc, t.gom_ids), t.gom_ids) if "G" in cfg["scheme"] else None
        O.dist_from_sim(O.similarity_mat(t.sig.copy(), gom, t.loc, None, 0.0, 0, cfg["scheme"], 1.0))
    dt = time.perf_counter() - t0
    return {"value": n_s * (n_s + 1) // 2 / dt, "unit": UNIT, "cores": 1,
            "sample": f"{n_s} genomes at full P={cfg['p']}, G={g_s}, base matrix only, single process; {dt:.1f} s "
                      f"(the CSV re-reads of similarity_matrix_constructor.py:19-110 are not charged)"}


# -------------------------------------------------------------------------------- our arm
_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's real stdout; everything else the run prints (library banners such as
    NCCL's version line, warnings) was redirected to stderr by main()."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                              # fd 1 -> stderr for native libraries and child processes
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--config", default=os.environ.get("GV2_BENCH_CONFIG", "cfg3"), choices=sorted(synth.CONFIGS))
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--boot", type=int, default=None, help="override the number of bootstrap replicates")
    ap.add_argument("--density", type=float, default=None, help="override the table density (1.0 = dense worst case)")
    ap.add_argument("--scheme", default=None, choices=["P", "PG", "G"])
    ap.add_argument("--seed", type=int, default=20241019)
    ap.add_argument("--ref-rows", type=int, default=120)
    ap.add_argument("--ref-rows-popc", type=int, default=64)
    ap.add_argument("--ref-boot", type=int, default=1)
    ap.add_argument("--as-shipped-rows", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--e2e-matrices", action="store_true", help="also time the variant that ships every matrix to the host")
    args = ap.parse_args()
    cfg = build_workload(args)
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, cfg)
        return
    if args.config == "cfg4":
        run_popcount(args, cfg)
    else:
        run_job(args, cfg)


CPU_GROUP = None


def init_dist(local_rank, world):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # its banner goes to stdout, ahead of the JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # a host-side group: ranks that wait for rank 0's single-process multi-GPU leg must not spin in a NCCL kernel on the
        # GPUs that leg is measured on
        global CPU_GROUP
        import datetime
        CPU_GROUP = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=60))
    return torch, dist


def run_job(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = measure_cpu_baseline(args, cfg)       # before CUDA init: it forks worker pools

    torch, dist = init_dist(local_rank, world)
    from gravity2_b200 import _lib, gom_membership, sharding

    big = cfg["n"] * cfg["n"] * (1 + cfg["b"]) > 2e10          # cfg3-sized jobs: a step takes seconds
    steps = args.steps if args.steps is not None else (3 if big else 5)
    warmup = args.warmup if args.warmup is not None else 3
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    ctx = C.c_void_p()
    rc = lib.gv2_create_on_stream(local_rank, C.c_void_p(stream.cuda_stream), C.byref(ctx))
    if rc:
        raise RuntimeError(lib.gv2_last_error(None).decode())

    def ck(rc):
        _lib.check(ctx, rc)

    # ---- synthetic tables (every rank generates the same tables) and reference-stream weights
    t = make_tables(args, cfg)
    n, p, g, B = cfg["n"], cfg["p"], len(t.gom_ids), cfg["b"]
    scheme, pl2, n_ref = cfg["scheme"], cfg["pl2"], cfg["n_ref"]
    idx = synth.reference_resample_indices(p, B, seed=args.seed, sort=pl2)
    weights = np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32) if B else np.zeros((0, p), np.int32)
    gor = gom_membership(t.taxo[:n_ref], t.gom_ids)
    pairs = n * (n + 1) // 2

    # ---- device-resident job (PL1: the signature table doubles as the location table, quirk Q3)
    d_sig = torch.from_numpy(t.sig).cuda()
    d_loc = torch.from_numpy(t.loc).cuda() if (pl2 and "G" in scheme) else d_sig
    job = C.c_void_p()
    gor_p = gor.ctypes.data_as(C.c_void_p)
    ck(lib.gv2_job_create(ctx, C.c_void_p(d_sig.data_ptr()), C.c_void_p(d_loc.data_ptr()) if "G" in scheme else None, 1, n, p,
                          gor_p if "G" in scheme else None, n_ref, g if "G" in scheme else 0,
                          _lib.SCHEMES[scheme], 1.0, _lib.OUT_DISTANCE, C.byref(job)))
    my_reps = sharding.replicate_indices(B, rank, world)
    d_w = torch.from_numpy(np.ascontiguousarray(weights[my_reps])).cuda() if my_reps else None
    ntiles = int(lib.gv2_num_tiles(n))
    per = sharding.tiles_per_rank(ntiles, world)
    t0_, t1_ = sharding.tile_range(ntiles, rank, world)
    tile_elems = _lib.TILE * _lib.TILE
    d_base = torch.empty((n, n), dtype=torch.float64, device="cuda")
    d_pack = torch.zeros((per, tile_elems), dtype=torch.float64, device="cuda") if world > 1 else None
    # replicate output ring: as many replicates as fit in ~8 GB (device-only mode: replicate b lives in slot b % ring)
    ring = max(1, min(len(my_reps) or 1, int((8 << 30) // (n * n * 8))))
    d_out = torch.empty((ring, n, n), dtype=torch.float64, device="cuda")

    def base_step():
        if world > 1:
            ck(lib.gv2_job_base(job, t0_, t1_, 1, C.c_void_p(d_pack.data_ptr())))
            d_gath = sharding.all_gather_tiles(d_pack, world)
            for r in range(world):
                a, b = sharding.tile_range(ntiles, r, world)
                if b > a:
                    ck(lib.gv2_dev_unpack_tiles(ctx, n, a, b, C.c_void_p(d_gath[r * per].data_ptr()), C.c_void_p(d_base.data_ptr())))
        else:
            ck(lib.gv2_job_base(job, 0, ntiles, 0, C.c_void_p(d_base.data_ptr())))

    def step():
        base_step()
        if my_reps:      # all of this rank's replicates in one call; results stream through the device ring
            ck(lib.gv2_job_replicates_stream(job, C.c_void_p(d_w.data_ptr()), len(my_reps), C.c_void_p(d_out.data_ptr()), ring,
                                             None, None, None))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = int(lib.gv2_launch_count(ctx))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    lib.gv2_timing_enable(ctx, 1)           # CUDA events around every launch group of the hot kernels, same stream
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = int(lib.gv2_launch_count(ctx)) - l0
    times = {}
    for name, kind in (("pairs", _lib.TIME_PAIRSUMS), ("tail", _lib.TIME_TAIL), ("gom", _lib.TIME_GOM), ("base", _lib.TIME_BASE)):
        k_ms, k_n = C.c_double(0.0), C.c_int64(0)
        ck(lib.gv2_timing_read_kind(ctx, kind, C.byref(k_ms), C.byref(k_n)))
        times[name] = (k_ms.value, k_n.value)
    lib.gv2_timing_enable(ctx, 0)
    path = _lib.PATH_NAMES.get(int(lib.gv2_job_replicate_path(job)), "?")
    cells = int(lib.gv2_job_copresent_cells(job))
    gstats = (C.c_int64 * 8)()
    lib.gv2_job_gom_stats(job, gstats)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        ll = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        launches = int(ll.item())
    ms_per_step = ms / steps
    value = pairs * (1 + B) / (ms_per_step * 1e-3)

    # ---- base-only time (cfg5's headline in SURVEY 8d is the B = 0 job)
    base_ms = None
    if rank == 0 or world > 1:
        sync_all()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        base_step()
        b1.record(stream)
        sync_all()
        base_ms = b0.elapsed_time(b1)

    # ---- parity of the last step's device results against the oracle (sampled cells); multi-rank: the all-gathered
    # base matrix against this rank's own full computation on sampled tiles
    parity = None
    if not args.no_parity:
        parity = parity_check(lib, ctx, ck, job, t, cfg, weights, my_reps, d_w, d_base, d_out, ring, gor, rank, world, torch, dist)

    peaks, peak_kind = load_peaks()
    roofline = None
    if rank == 0:
        roofline = make_roofline(times, path, cells, list(gstats), steps, n, p, g, len(my_reps), ms, peaks, peak_kind, scheme,
                                 clocks, args.config)
    # ---- end to end through the public host API: release the device-resident job first (its workspaces hold ~100 GB)
    lib.gv2_job_destroy(job)
    job = None
    del d_out, d_base, d_sig, d_loc
    torch.cuda.empty_cache()
    e2e = e2e_mat = e2e_cond = None
    dd = dist if world > 1 else None
    if not args.no_e2e and world > 1:
        # N GPUs through the public API = ONE caller process driving all of them (gravity2_b200.MultiEngine: one host thread
        # per GPU inside the library).  Rank 0 is that caller; the other ranks have released their GPUs and wait.
        sync_all()
        if rank == 0:
            e2e = measure_e2e_multi(t, cfg, gor, weights, pairs, world)
        dist.barrier(group=CPU_GROUP)
    elif not args.no_e2e:
        if B > 0 and not pl2:
            e2e = measure_e2e_trees(t, cfg, gor, weights, my_reps, pairs, world, rank, dd, torch)
            if args.e2e_matrices:
                e2e_mat = measure_e2e_matrices(t, cfg, gor, weights, my_reps, pairs, world, rank, dd, torch, steps=1)
                e2e_cond = measure_e2e_matrices(t, cfg, gor, weights, my_reps, pairs, world, rank, dd, torch, steps=1, condensed=True)
        elif pl2 and B > 0:
            e2e = measure_e2e_classification(t, cfg, gor, weights, my_reps, pairs, world, rank, dd, torch)
            e2e_mat = measure_e2e_matrices(t, cfg, gor, weights, my_reps, pairs, world, rank, dd, torch, steps=1)
        else:
            e2e = measure_e2e_matrices(t, cfg, gor, weights, my_reps, pairs, world, rank, dd, torch, steps=max(1, min(steps, 2)))
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": {"sparse": "48-bit fixed-point min x u8 multiplicities -> int64 (exact)",
                          "mma": "u32 fixed-point min -> u8 digit planes x u8 multiplicities -> int32 (exact)",
                          "fp32": "f32 min, f32 slab sums -> f64"}.get(path, path) +
                         ("; GOM Jaccard u32 fixed-point min / u32 sums (exact); GOM dCor f64; epilogue f64" if "G" in scheme else "; epilogue f64"),
                "data": "synthetic",
                "config": workload_config(args, cfg, world), "clocks": clocks, "gpu_launches": launches,
                "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity_check": parity,
                "base_only_ms": base_ms}
        if e2e_mat:
            line["e2e_matrices"] = e2e_mat
        if e2e_cond:
            line["e2e_condensed"] = e2e_cond
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if parity and not parity.get("ok", False):
        sys.exit(3)


def parity_check(lib, ctx, ck, job, t, cfg, weights, my_reps, d_w, d_base, d_out, ring, gor, rank, world, torch, dist):
    """Sampled cells of the last timed step against the CPU oracle (oracle/sampled_check.py): GOM table cells rel 1e-9,
    matrix cells abs 1e-6 (north_star tolerance).  The replicate matrices checked are the ones still in the device ring
    (device-only streaming keeps replicate b in slot b % ring), their GOM tables are recomputed by gv2_job_gom_table."""
    from oracle import sampled_check as SC
    n, p, scheme, pl2, n_ref = cfg["n"], cfg["p"], cfg["scheme"], cfg["pl2"], cfg["n_ref"]
    rng = np.random.default_rng(1234 + rank)
    loc_tab = t.loc if pl2 else t.sig
    taxo_ref = np.asarray(t.taxo[:n_ref])
    g = len(t.gom_ids)
    members = {}
    out = {"ok": True, "gom_cells": 0, "matrix_cells": 0, "worst_gom_rel": 0.0, "worst_matrix_abs": 0.0, "failures": []}
    want_g = "G" in scheme

    def member_rows(k):
        if k not in members:
            members[k] = np.where(taxo_ref == t.gom_ids[k])[0]
        return members[k]

    def gom_rows(b_local):
        if not want_g:
            return None
        buf = torch.empty((n, g), dtype=torch.float64, device="cuda")
        wp = None if b_local is None else C.c_void_p(d_w[b_local].data_ptr())
        ck(lib.gv2_job_gom_table(job, wp, 1, C.c_void_p(buf.data_ptr())))
        torch.cuda.synchronize()
        return buf.cpu().numpy()

    def check_matrix(tag, dev_matrix, gt, w_row, k_pairs, k_gom):
        taxon_all = np.asarray(list(taxo_ref) + ["?"] * (n - n_ref))
        prs = SC.sample_pairs(rng, n, taxon_all, k_pairs)
        ii = torch.tensor([a for a, _ in prs], device="cuda"); jj = torch.tensor([b for _, b in prs], device="cuda")
        got = dev_matrix[ii, jj].cpu().numpy()
        got_t = dev_matrix[jj, ii].cpu().numpy()
        for (i, j), v, vt in zip(prs, got, got_t):
            ref = SC.pair_cell(t.sig, None if gt is None else gt[i], None if gt is None else gt[j], i, j, w_row, scheme, 1.0, True)
            err = abs(v - ref)
            out["matrix_cells"] += 1
            out["worst_matrix_abs"] = max(out["worst_matrix_abs"], float(err))
            if not (err <= 1e-6 and v == vt):
                out["ok"] = False
                out["failures"].append(f"{tag} D[{i},{j}] device {v!r} oracle {ref!r}")
        for _ in range(k_gom if want_g else 0):
            v = int(rng.integers(0, n))
            k = int(gor[v]) if (v < n_ref and rng.random() < 0.4) else int(rng.integers(0, g))
            ref = SC.gom_cell(loc_tab, member_rows(k), v, w_row)
            dv = float(gt[v, k])
            ok = (np.isnan(ref) and np.isnan(dv)) or abs(dv - ref) <= 1e-9 * abs(ref) + 1e-12
            out["gom_cells"] += 1
            if not np.isnan(ref):
                out["worst_gom_rel"] = max(out["worst_gom_rel"], float(abs(dv - ref) / max(abs(ref), 1e-300)))
            if not ok:
                out["ok"] = False
                out["failures"].append(f"{tag} GOM[{v},{k}] device {dv!r} oracle {ref!r}")

    if rank == 0:
        check_matrix("base", d_base, gom_rows(None), None, 24, 3)
    nloc = len(my_reps)
    for b_local in sorted({nloc - 1, nloc - 1 - ring // 2, nloc - ring} & set(range(max(0, nloc - ring), nloc))):
        check_matrix(f"replicate {my_reps[b_local]}", d_out[b_local % ring], gom_rows(b_local), weights[my_reps[b_local]], 10, 1)
    if world > 1:
        # the all-gathered base matrix against this rank's OWN computation of sampled tiles owned by other ranks
        ntiles = int(lib.gv2_num_tiles(n))
        nt = (n + 127) // 128
        chk = torch.empty((128 * 128,), dtype=torch.float64, device="cuda")
        from gravity2_b200 import sharding
        bad = 0
        for tix in rng.integers(0, ntiles, 6):
            tix = int(tix)
            ck(lib.gv2_job_base(job, tix, tix + 1, 1, C.c_void_p(chk.data_ptr())))
            bi, bj = sharding.tile_coords(tix, nt)
            r1, c1 = min(n, (bi + 1) * 128), min(n, (bj + 1) * 128)
            mine = chk.view(128, 128)[: r1 - bi * 128, : c1 - bj * 128]
            ref = d_base[bi * 128:r1, bj * 128:c1]
            if bi == bj:                      # diagonal tiles: the packed form holds the upper triangle's values
                mine, ref = torch.triu(mine), torch.triu(ref)
            if not torch.equal(mine, ref):
                bad += 1
        flag = torch.tensor([bad, 0 if out["ok"] else 1], dtype=torch.int64, device="cuda")
        dist.all_reduce(flag)
        out["allgather_tiles_mismatching"] = int(flag[0].item())
        out["ranks_failing"] = int(flag[1].item())
        if flag[0].item() or flag[1].item():
            out["ok"] = False
    out["failures"] = out["failures"][:5]
    out["tolerance"] = "GOM cells rel 1e-9; matrix cells abs 1e-6 and exactly symmetric; all-gathered tiles bit-identical"
    return out


def make_roofline(times, path, cells, gstats, steps, n, p, g, nrep, region_ms, peaks, peak_kind, scheme, clocks, config):
    """Roofline of the dominant kernel group (largest share of the timed region) + a line for every timed group.
    Durations are CUDA-event times taken inside the timed region on the launching stream (gv2_timing_read_kind).
    Each line states the BINDING resource as `bound` with achieved / peak / frac on it; the other resource is
    supporting evidence under `also`.  Algorithmic work per step, pairs = N(N+1)/2:
      tail  (gj_fused kernels, one launch per sub-chunk of replicates): pairs x G cells of the (min, +) semiring at 2
            issue slots per cell (one min on the ALU pipe + one add on the FMA pipe): bound by the FP32/INT issue
            rate; bytes: READ 8 B x pairs (PPHMM pair sums) + 4 N G (GOM table), WRITE 8 N^2;
      pairs (replicate pair sums): sparse path: WRITES 8 B x pairs x replicates of workspace (HBM bound);
            tensor-core path: ALGORITHMIC flops 2 x pairs x P x replicates against the measured sustained bf16 rate
            (the kernel issues 4 uint8 digit planes per value: 4 x that in tensor operations, reported separately);
      gom   (GOM dCor scoring): fp64 pipe: sum_g n_g^2 DFMA per replicate for the GOM-side row sums + one DFMA per
            stored M entry per replicate for the pair sums; bytes: the distance cache once per 32 replicates."""
    pairs = n * (n + 1) // 2
    hbm = peaks.get("hbm_gbs", 6550.0)
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    issue_peak = 148 * 128 * sm_mhz * 1e6                 # thread-instructions/s over all SMSPs
    fp64_peak = 148 * 64 * sm_mhz * 1e6                   # DFMA/s (64 per clock and SM on sm_100)
    traffic = load_traffic()
    lines = {}
    ms, cnt = times["tail"]
    if cnt:
        sec = ms * 1e-3 / steps
        bytes_step = nrep * (8.0 * pairs * (1 if "P" in scheme else 0) + 8.0 * n * n + 4.0 * n * g)
        slots = nrep * float(pairs) * g * 2
        per_launch = round(nrep / max(cnt / steps, 1))
        lines["tail"] = {"kernel": "gj_fused kernel (GOM-signature (min,+) pair sums on 128x64 tiles + clean/combine/clamp/1-S/mirror epilogue)",
                         "bound": "issue", "achieved": slots / sec, "peak": issue_peak, "unit": "thread-instr/s",
                         "frac": slots / sec / issue_peak,
                         "bound_detail": "FP32/INT issue slots: 2 per (pair, GOM) cell -- one min (ALU pipe) + one add (FMA pipe); "
                                         "148 SMs x 128 lanes x sm_max_mhz",
                         "also": {"bound": "hbm", "achieved": bytes_step / sec / 1e9, "peak": hbm, "unit": "GB/s",
                                  "frac": bytes_step / sec / 1e9 / hbm},
                         "traffic": traffic.get(f"gj_fused:{config}", cnt / steps),
                         "launches_per_step": cnt / steps, "avg_launch_ms": ms / cnt, "share_of_step": ms / region_ms}
    ms, cnt = times["pairs"]
    if cnt:
        sec = ms * 1e-3 / steps
        per_launch = round(nrep / max(cnt / steps, 1))
        if path == "mma":
            bf16 = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0))
            flops = 2.0 * float(pairs) * p * nrep
            lines["pairs"] = {"kernel": "boot_mma_kernel (tcgen05.mma kind::i8, uint8 digit planes x uint8 multiplicities, TMEM int32 accumulators)",
                              "bound": "tensor", "achieved": flops / sec / 1e12, "peak": bf16, "unit": "TFLOP/s",
                              "frac": flops / sec / 1e12 / bf16,
                              "bound_detail": "ALGORITHMIC flops 2 x pairs x P x replicates against bf16_tflops_sustained",
                              "also": {"bound": "tensor ops issued (4 digit planes per value)", "achieved": 4 * flops / sec / 1e12,
                                       "peak": 2.0 * bf16, "unit": "int8 TOP/s", "frac": 4 * flops / sec / 1e12 / (2.0 * bf16)},
                              "traffic": traffic.get(f"boot_mma:{config}", cnt / steps)}
        else:
            bytes_step = 8.0 * pairs * nrep
            lines["pairs"] = {"kernel": f"replicate pair sums, path {path} (pair lists x uint8 multiplicities, exact integer sums -> fp64 workspace)",
                              "bound": "hbm", "achieved": bytes_step / sec / 1e9, "peak": hbm, "unit": "GB/s",
                              "frac": bytes_step / sec / 1e9 / hbm,
                              "bound_detail": "algorithmic bytes = the fp64 pair-sum workspace it writes (8 B x pairs x replicates)",
                              "traffic": traffic.get(f"boot_sparse:{config}", cnt / steps),
                              "copresent_cells": int(cells), "cell_replicates_per_s": cells * nrep / sec if cells else None}
        lines["pairs"].update({"launches_per_step": cnt / steps, "avg_launch_ms": ms / cnt, "share_of_step": ms / region_ms})
    ms, cnt = times["gom"]
    if cnt:
        sec = ms * 1e-3 / steps
        sum_ng2, m_entries, npt, nnz = gstats[1], gstats[5], gstats[0], gstats[2]
        dfma = float(nrep) * (sum_ng2 + m_entries + 40.0 * n * g)          # + ~40 fp64 operations to finish a score
        chunks = cnt / steps
        bytes_step = chunks * (8.0 * sum_ng2 + 8.0 * m_entries + 32 * 8.0 * (2 * nnz + 2 * npt)) + nrep * 8.0 * n * g
        lines["gom"] = {"kernel": "GOM dCor scoring (gom_rowsums / gom_aggregates / genome_rowsums / gom_score kernels), 32 replicates per chunk",
                        "bound": "fp64 pipe", "achieved": dfma / sec, "peak": fp64_peak, "unit": "DFMA/s", "frac": dfma / sec / fp64_peak,
                        "bound_detail": "sum_g n_g^2 (GOM row sums) + stored M entries (pair sums) + ~40 per score, per replicate; "
                                        "148 SMs x 64 DFMA/clk x sm_max_mhz",
                        "also": {"bound": "hbm", "achieved": bytes_step / sec / 1e9, "peak": hbm, "unit": "GB/s",
                                 "frac": bytes_step / sec / 1e9 / hbm},
                        "scores_per_s": float(n) * g * nrep / sec,
                        "traffic": traffic.get(f"gom:{config}", chunks),
                        "launches_per_step": chunks, "avg_launch_ms": ms / cnt, "share_of_step": ms / region_ms}
    ms, cnt = times["base"]
    if cnt:
        lines["base"] = {"kernel": "base (un-resampled) matrix", "launches_per_step": cnt / steps, "avg_launch_ms": ms / cnt,
                         "share_of_step": ms / region_ms}
    if not lines:
        return None
    dominant = max((k for k in lines if k != "base"), key=lambda k: lines[k]["share_of_step"], default=None)
    out = dict(lines[dominant]) if dominant else {}
    out["peak_kind"] = f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs / bf16_tflops_sustained / sm_max_mhz)"
    out["kernels"] = {k: {kk: v[kk] for kk in ("kernel", "bound", "achieved", "peak", "unit", "frac", "also", "traffic", "share_of_step",
                                               "avg_launch_ms", "launches_per_step") if kk in v}
                      for k, v in lines.items()}
    return out


def _engine_and_weights(weights, my_reps, p):
    import gravity2_b200
    eng = gravity2_b200.Engine(int(os.environ.get("LOCAL_RANK", "0")))
    w_mine = np.ascontiguousarray(weights[my_reps]) if my_reps else np.zeros((0, p), np.int32)
    return gravity2_b200, eng, w_mine


def measure_e2e_matrices(t, cfg, gor, weights, my_reps, pairs, world, rank, dist, torch, steps, condensed=False):
    """Whole job through the package's public host API: numpy tables in host memory -> gravity2_b200.engine.Job (one
    H2D upload of the tables per step) -> base matrix as a numpy array -> every replicate matrix delivered in host
    memory (pinned ring, zero-copy view) to a consumer that reads it.  H2D and D2H inside the timed region."""
    from gravity2_b200.engine import Job
    n, p, g, B, scheme, pl2 = cfg["n"], cfg["p"], len(t.gom_ids), cfg["b"], cfg["scheme"], cfg["pl2"]
    G, eng, w_mine = _engine_and_weights(weights, my_reps, p)
    acc = [0.0, 0]

    def consume(b, m):              # the device->host read of the step's result: touch every delivered matrix
        acc[0] += (float(m[0]) + float(m[-1]) + float(m[len(m) // 2])) if m.ndim == 1 else (float(m[0, -1]) + float(m[-1, 0]) + float(m[n // 2, n // 3]))
        acc[1] += 1

    def job_once():
        job = Job(eng, t.sig, t.loc if pl2 else t.sig, gor if "G" in scheme else None, g, scheme, 1.0, distance=True)
        try:
            if rank == 0:
                consume(-1, job.base())
            job.for_each_replicate(w_mine, consume, condensed=condensed)
        finally:
            job.close()

    job_once()          # warm-up
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    acc[1] = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        job_once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    assert acc[1] == steps * (len(my_reps) + (1 if rank == 0 else 0)), "consumer did not see every matrix"
    if dist is not None:
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    h2d = n * p * 8 * (2 if pl2 and "G" in scheme else 1) + w_mine.nbytes + gor.nbytes
    d2h = (1 if rank == 0 else 0) * n * n * 8 + len(my_reps) * (n * (n - 1) // 2 if condensed else n * n) * 8
    eng.close()
    return {"value": pairs * (1 + B) / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": dt * 1e3,
            "format": "condensed upper triangle (scipy squareform order: what dist_mat_to_tree.py:8-9 feeds linkage)" if condensed
                      else "full (N, N) matrices",
            "api": "gravity2_b200.engine.Job(sig, loc, ...).base() + .for_each_replicate(weights, consumer"
                   + (", condensed=True" if condensed else "") + "): numpy in, numpy views of a pinned host ring out (per-rank bytes)"}


def measure_e2e_classification(t, cfg, gor, weights, my_reps, pairs, world, rank, dist, torch):
    """Pipeline II keeps, of every replicate's similarity matrix, the sampled intra / inter class cells behind the cut-offs and
    the row maxima of the query block (app/src/virus_classification.py:318-344, classification_utils.py:10-68): numpy tables
    in -> gravity2_b200.engine.Job -> per replicate those cells and maxima out, the matrices never leave HBM
    (Job.base_device / for_each_replicate_device + DeviceMatrix.gather / rowmax: what bootstrap_classification runs on).
    The reference's host-side list sampling (np.random.choice, seconds per replicate) and SVC fits are not on this path:
    ONE cell set is drawn before the timed region and read from every matrix."""
    from gravity2_b200.engine import Job
    from gravity2_b200.classification import draw_class_cells, max_similarity
    n, p, g, B, scheme = cfg["n"], cfg["p"], len(t.gom_ids), cfg["b"], cfg["scheme"]
    n_ref = cfg["n_ref"]
    G, eng, w_mine = _engine_and_weights(weights, my_reps, p)
    taxo_ref = np.asarray(t.taxo[:n_ref])
    state = np.random.get_state()
    np.random.seed(7)
    cells = draw_class_cells(taxo_ref, 10000)                  # payload default N_PairwiseSimilarityScores (api_classes.py:82)
    np.random.set_state(state)
    rows = np.concatenate([np.concatenate((c[0], c[2])) for c in cells.values()]).astype(np.int64)
    cols = np.concatenate([np.concatenate((c[1], c[3])) for c in cells.values()]).astype(np.int64)
    acc = [0.0, 0]

    def consume(b, S):
        mx, _ = max_similarity(S, taxo_ref, row_begin=n_ref, n_ref=n_ref)
        vals = S.gather(rows, cols)
        acc[0] += float(vals[0]) + float(vals[-1]) + float(mx[0])
        acc[1] += 1

    def job_once():
        job = Job(eng, t.sig, t.loc, gor if "G" in scheme else None, g, scheme, 1.0, distance=False)
        try:
            if rank == 0:
                job.base_device(lambda S: consume(-1, S))
            job.for_each_replicate_device(w_mine, consume)
        finally:
            job.close()

    job_once()          # warm-up
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    acc[1] = 0
    t0 = time.perf_counter()
    job_once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    assert acc[1] == len(my_reps) + (1 if rank == 0 else 0), "consumer did not see every matrix"
    per_matrix = len(rows) * 8 + (n - n_ref) * 16
    return {"value": pairs * (1 + B) / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(2 * n * p * 8 + w_mine.nbytes + gor.nbytes + (len(my_reps) + 1) * len(rows) * 16),
            "d2h_bytes_per_step": int((len(my_reps) + 1) * per_matrix),
            "result": f"per matrix {len(rows)} sampled class cells + the {n - n_ref} row maxima / argmaxima of the query block "
                      "(what virus_classification.py:318-344 reads of it)",
            "api": "gravity2_b200.engine.Job(sig, loc, ...).base_device + .for_each_replicate_device(weights, consumer) with "
                   "DeviceMatrix.gather / max_similarity: numpy in, cells and maxima out, matrices stay in HBM (per-rank bytes)"}


def measure_e2e_trees(t, cfg, gor, weights, my_reps, pairs, world, rank, dist, torch):
    """The bootstrap loop's own product (app/src/pl1_graphs.py:108-117 keeps only the dendrogram of a replicate): numpy
    tables in -> Job.for_each_tree -> one Newick string per replicate out, the distance matrices clustered in HBM (device
    UPGMA, scipy's nn_chain average linkage bit for bit) and never shipped.  H2D of the tables and D2H of the linkage rows
    inside the timed region.  One step (the job takes seconds)."""
    from gravity2_b200.engine import Job
    n, p, g, B, scheme = cfg["n"], cfg["p"], len(t.gom_ids), cfg["b"], cfg["scheme"]
    G, eng, w_mine = _engine_and_weights(weights, my_reps, p)
    names = [f"{x}_{i}" for i, x in enumerate(t.taxo)]
    acc = [0, 0]

    def consume(b, s):
        acc[0] += len(s)
        acc[1] += 1

    # the un-resampled tree rides along as the all-ones "replicate" of rank 0 (same kernels, clustered with the others)
    w_all = np.concatenate([np.ones((1, p), np.int32), w_mine]) if rank == 0 else w_mine

    def job_once():
        job = Job(eng, t.sig, t.sig, gor if "G" in scheme else None, g, scheme, 1.0, distance=True)
        try:
            job.for_each_tree(w_all, names, consume)
        finally:
            job.close()

    job_once()          # warm-up
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    acc[1] = 0
    t0 = time.perf_counter()
    job_once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert acc[1] == len(my_reps) + (1 if rank == 0 else 0), "consumer did not see every tree"
    if dist is not None:
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    eng.close()
    return {"value": pairs * (1 + B) / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(n * p * 8 + w_all.nbytes + gor.nbytes),
            "d2h_bytes_per_step": int(len(w_all) * (n - 1) * 32),
            "result": "one Newick string per replicate + the base tree (what app/src/pl1_graphs.py:108-117 keeps)",
            "api": "gravity2_b200.engine.Job(sig, loc, ...).for_each_tree(weights, leaf_names, consumer): numpy in, one Newick string "
                   "per replicate (and for the un-resampled table, as the all-ones replicate) out; UPGMA on the device; per-rank bytes"}


def measure_e2e_multi(t, cfg, gor, weights, pairs, world):
    """The whole job through the public host API on ALL `world` GPUs from one process: MultiJob.base() + .for_each_tree (the
    bootstrap loops' product) or, for the pipeline II workload, .for_each_replicate (matrices to host memory)."""
    import gravity2_b200 as G
    n, p, g, B, scheme, pl2 = cfg["n"], cfg["p"], len(t.gom_ids), cfg["b"], cfg["scheme"], cfg["pl2"]
    meng = G.MultiEngine(list(range(world)))
    names = [f"{x}_{i}" for i, x in enumerate(t.taxo)]
    acc = [0]
    trees = B > 0 and not pl2

    def job_once():
        job = G.MultiJob(meng, t.sig, t.loc if pl2 else t.sig, gor if "G" in scheme else None, g, scheme, 1.0, distance=True)
        try:
            if trees:
                w_all = np.concatenate([np.ones((1, p), np.int32), weights])
                job.for_each_tree(w_all, names, lambda b, s: acc.__setitem__(0, acc[0] + 1))
            else:
                base = job.base()
                acc[0] += 1 + int(base[0, -1] < 0)
                job.for_each_replicate(weights, lambda b, m: acc.__setitem__(0, acc[0] + 1 + int(m[0, -1] < 0)))
        finally:
            job.close()

    job_once()          # warm-up
    acc[0] = 0
    t0 = time.perf_counter()
    job_once()
    dt = time.perf_counter() - t0
    assert acc[0] == B + 1, "consumer did not see every replicate"
    nccl = meng.uses_nccl
    meng.close()
    d2h = (B + 1) * (n - 1) * 32 if trees else (B + 1) * n * n * 8
    return {"value": pairs * (1 + B) / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(n * p * 8 * (2 if pl2 and "G" in scheme else 1) + weights.nbytes + gor.nbytes),
            "d2h_bytes_per_step": int(d2h), "nccl": nccl,
            "result": "one Newick string per replicate + the base tree" if trees else "every replicate matrix in host memory",
            "api": f"gravity2_b200.MultiEngine({world} GPUs) / MultiJob: ONE caller process, one host thread per GPU inside libgravity_b200 "
                   "(tables uploaded once + NVLink broadcast, replicates by index, base tiles all-gathered)"}


# --------------------------------------------------------------------- cfg4: presence / absence popcount path
def run_popcount(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, dt, cores, sample = time_reference(args, cfg, 0, 1, 0)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    torch, dist = init_dist(local_rank, world)
    from gravity2_b200 import _lib, sharding
    steps = args.steps if args.steps is not None else 5
    warmup = args.warmup if args.warmup is not None else 3
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    ctx = C.c_void_p()
    rc = lib.gv2_create_on_stream(local_rank, C.c_void_p(stream.cuda_stream), C.byref(ctx))
    if rc:
        raise RuntimeError(lib.gv2_last_error(None).decode())

    def ck(rc):
        _lib.check(ctx, rc)

    t = make_tables(args, cfg)
    n, p = cfg["n"], cfg["p"]
    pairs = n * (n + 1) // 2
    n_pad = (n + 127) // 128 * 128
    words = ((p + 63) // 64 + 15) // 16 * 16
    d_bits = torch.zeros((n_pad, words), dtype=torch.int64, device="cuda")
    d_nnz = torch.zeros((n,), dtype=torch.int32, device="cuda")
    rows = 4096                                             # the float64 table goes through the packer in row chunks
    for r0 in range(0, n, rows):
        r1 = min(n, r0 + rows)
        chunk = torch.from_numpy(t.sig[r0:r1]).cuda()
        ck(lib.gv2_dev_pack_presence(ctx, C.c_void_p(chunk.data_ptr()), r1 - r0, p, p, C.c_void_p(d_bits[r0].data_ptr()),
                                     (r1 - r0 + 127) // 128 * 128 if r1 == n else rows, words, C.c_void_p(d_nnz[r0:].data_ptr())))
        torch.cuda.synchronize()
        del chunk
    ntiles = int(lib.gv2_num_tiles(n))
    per = sharding.tiles_per_rank(ntiles, world)
    t0_, t1_ = sharding.tile_range(ntiles, rank, world)
    tile_elems = 128 * 128
    d_inter = torch.empty((n, n), dtype=torch.int32, device="cuda")
    d_pack = torch.zeros((per, tile_elems), dtype=torch.int32, device="cuda") if world > 1 else None

    def step():
        if world > 1:
            ck(lib.gv2_dev_shared_counts_packed(ctx, C.c_void_p(d_bits.data_ptr()), n, n_pad, words, t0_, t1_, C.c_void_p(d_pack.data_ptr())))
            d_gath = sharding.all_gather_tiles(d_pack, world)
            for r in range(world):
                a, b = sharding.tile_range(ntiles, r, world)
                if b > a:
                    ck(lib.gv2_dev_unpack_tiles_i32(ctx, n, a, b, C.c_void_p(d_gath[r * per].data_ptr()), C.c_void_p(d_inter.data_ptr())))
        else:
            ck(lib.gv2_dev_shared_counts(ctx, C.c_void_p(d_bits.data_ptr()), n, n_pad, words, 0, ntiles, C.c_void_p(d_inter.data_ptr())))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = int(lib.gv2_launch_count(ctx))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = int(lib.gv2_launch_count(ctx)) - l0
    clocks = sampler.stop() if rank == 0 else None
    # the tile kernel alone (roofline): this rank's tile range, CUDA events on the launching stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(steps):
        if world > 1:
            ck(lib.gv2_dev_shared_counts_packed(ctx, C.c_void_p(d_bits.data_ptr()), n, n_pad, words, t0_, t1_, C.c_void_p(d_pack.data_ptr())))
        else:
            ck(lib.gv2_dev_shared_counts(ctx, C.c_void_p(d_bits.data_ptr()), n, n_pad, words, 0, ntiles, C.c_void_p(d_inter.data_ptr())))
    k1.record(stream)
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / steps
    popc_peak = C.c_double(0.0)
    ck(lib.gv2_dev_popc_peak(ctx, 4096, C.byref(popc_peak)))
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        ll = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        launches = int(ll.item())
    ms_per_step = ms / steps
    # ---- parity: bit-exact counts on sampled rows (numpy restatement of shared_pphmm_graphs.py:63-76 on the same table)
    parity = None
    if not args.no_parity:
        rng = np.random.default_rng(5 + rank)
        rows_s = [0, n - 1] + [int(x) for x in rng.integers(0, n, 6)]
        nz = t.sig != 0
        bad = 0
        for r in rows_s:
            ref = (nz & nz[r]).sum(axis=1).astype(np.int32)
            got = d_inter[r].cpu().numpy()
            got_c = d_inter[:, r].cpu().numpy()
            bad += int(not (np.array_equal(got, ref) and np.array_equal(got_c, ref)))
        nnz_ok = bool(np.array_equal(d_nnz.cpu().numpy(), nz.sum(axis=1).astype(np.int32)))
        flag = torch.tensor([bad + (0 if nnz_ok else 1)], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(flag)
        parity = {"ok": int(flag.item()) == 0, "rows_checked_per_rank": len(rows_s), "cells": len(rows_s) * n * 2,
                  "tolerance": "bit-exact int32 counts (rows and mirrored columns) and per-genome hit counts"}
    peaks, peak_kind = load_peaks()
    e2e = None
    if not args.no_e2e and world == 1:
        e2e = measure_e2e_popcount(t, n, p, pairs, torch)
    if rank == 0:
        my_tiles = t1_ - t0_
        popc_ops = float(my_tiles) * 128 * 128 * (2 * words)            # 32-bit AND + POPC + add per pair and word
        sec = kernel_ms * 1e-3
        roofline = {"kernel": "shared_count_tile_kernel (128x128 pair tiles, 32-bit AND + POPC + IADD per word)",
                    "bound": "popc pipe", "achieved": popc_ops / sec, "peak": popc_peak.value, "unit": "POPC/s",
                    "frac": popc_ops / sec / popc_peak.value if popc_peak.value else None,
                    "peak_kind": "measured in this run: register-resident LOP3 + POPC + IADD chains, 8 per thread, full grid (gv2_dev_popc_peak)",
                    "also": {"bound": "hbm", "achieved": (4.0 * n * n / world + n_pad * words * 8) / sec / 1e9, "peak": peaks.get("hbm_gbs"),
                             "unit": "GB/s", "frac": (4.0 * n * n / world + n_pad * words * 8) / sec / 1e9 / peaks.get("hbm_gbs", 6550.0)},
                    "traffic": load_traffic().get("shared_count_tile:cfg4", 1), "avg_launch_ms": kernel_ms,
                    "tiles_per_launch": my_tiles, "share_of_step": kernel_ms / ms_per_step}
        line = {"metric": METRIC, "value": pairs / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32 bit words: AND + POPC -> int32 (exact)", "data": "synthetic",
                "config": workload_config(args, cfg, world), "clocks": clocks, "gpu_launches": launches,
                "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity_check": parity}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if parity and not parity["ok"]:
        sys.exit(3)


def measure_e2e_popcount(t, n, p, pairs, torch):
    """shared_pphmm_counts(table) through the public host API: float64 table in host memory in, int32 (N, N) counts and
    per-genome hit counts in host memory out."""
    import gravity2_b200
    eng = gravity2_b200.Engine(int(os.environ.get("LOCAL_RANK", "0")))
    inter, nnz = eng.shared_counts(t.sig[:256])       # warm-up (library load, context)
    t0 = time.perf_counter()
    inter, nnz = eng.shared_counts(t.sig)
    dt = time.perf_counter() - t0
    assert inter.shape == (n, n) and int(inter[0, 0]) == int(nnz[0])
    eng.close()
    return {"value": pairs / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(n * p * 8),
            "d2h_bytes_per_step": int(n * n * 4 + n * 4),
            "api": "gravity2_b200.Engine.shared_counts(table) (what shared_pphmm_ratio / shared_norm_pphmm_ratio call): "
                   "float64 numpy table in, int32 numpy counts out"}


if __name__ == "__main__":
    main()
