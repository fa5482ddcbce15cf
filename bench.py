#!/usr/bin/env python
"""Benchmark of the pairwise genome-similarity hot path (BASELINE.json metric:
genome-pair similarities/sec incl. bootstraps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2|cfg3|...] [--impl reference]

One "step" = the whole job on synthetic tables of the named shape: the base PG similarity
matrix (GOM table + Jaccard + combine) plus B bootstrap replicate distance matrices (GOM table
rebuilt per replicate, as app/src/pl1_graphs.py:78-107 does).  pairs/s = N(N+1)/2 * (1+B) / time.

  value : inputs resident in HBM, outputs written to HBM (device-timed, CUDA events on the
          stream the kernels run on, max over ranks)
  e2e   : the same job through the public host API (gravity2_b200.engine.Job) with numpy tables in host
          memory: H2D of the tables and D2H of every replicate matrix inside the timed region
  e2e_trees : the same job with the bootstrap loops' real product as the result: one Newick string per
          replicate, the distance matrices clustered on the device (UPGMA) and never shipped
  --impl reference : the reference's numpy pair expressions and dcor-based GOM scoring (CPU oracle
          port of app/utils/similarity_matrix_constructor.py:167-177 and
          parallel_gom_sig_generator.py:30-37) on all host cores, on a bounded row sample.
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


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
            a, b = gom[i], gom[j]
            mx = np.sum(np.maximum(a, b))
            outg[i - i0, j] = np.sum(np.minimum(a, b)) / mx if not mx == 0 else 0
    return outp, outg


def _ref_gom(g):
    from oracle import gravity_oracle as O
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.generate_gom_sigs(_G["db"][g], _G["loc"])


def reference_job(cores, sig, loc, taxo, gom_ids, idx_list):
    """Base matrix + one replicate per idx in idx_list on the CPU oracle port, all host cores;
    returns pairs*(1+B).  A fresh fork pool per stage, as the reference spawns a Pool per GOM table
    (parallel_gom_sig_generator.py:18)."""
    import multiprocessing as mp
    from oracle import gravity_oracle as O
    n = sig.shape[0]
    ctxmp = mp.get_context("fork")

    def one(sig_t, loc_t):
        _G.update(sig=sig_t, loc=loc_t, db=O.gomdb(taxo, loc_t, gom_ids))
        with ctxmp.Pool(cores) as pool:
            cols = pool.map(_ref_gom, list(gom_ids))
        gom = np.column_stack(cols) if cols else np.zeros((n, 0))
        _G["gom"] = gom
        # balance the triangle: many small row blocks
        bounds = np.unique(np.linspace(0, n, 8 * cores + 1).astype(int))
        with ctxmp.Pool(cores) as pool:
            parts = pool.map(_ref_rows, [(bounds[k], bounds[k + 1]) for k in range(len(bounds) - 1)], chunksize=1)
        P = np.vstack([p for p, _ in parts]); Gm = np.vstack([g for _, g in parts])
        P = np.triu(P) + np.triu(P, 1).T; Gm = np.triu(Gm) + np.triu(Gm, 1).T
        S = O._combine(P, Gm, np.zeros((n, n)), None, "PG", 1.0)
        return O.dist_from_sim(S)

    one(sig, loc)
    for idx in idx_list:
        b = np.ascontiguousarray(sig[:, idx])
        one(b, b)                                        # PL1: signature table doubles as location table (Q3)
    return n * (n + 1) // 2 * (1 + len(idx_list))


def run_reference(args, cfg):
    cores = os.cpu_count() or 1
    n_s = min(cfg["n"], args.ref_rows)
    b_s = min(cfg["b"], args.ref_boot)
    t = synth.make_tables(n_s, cfg["p"], min(cfg["g"], max(1, n_s // 4)), cfg["density"], synth.SEED_BASE + int(args.config[-1]))
    idx = synth.reference_resample_indices(cfg["p"], b_s, seed=args.seed)
    sample = f"{n_s} of {cfg['n']} genomes at full P={cfg['p']}, G={len(t.gom_ids)}, base + {b_s} of {cfg['b']} replicates (PL1 body)"
    k = max(8, n_s // 8)
    for _ in range(min(1, args.warmup if args.warmup is not None else 1)):
        reference_job(cores, t.sig[:k], t.loc[:k], t.taxo[:k], t.gom_ids, idx[:1])
    t0 = time.perf_counter()
    units = 0
    steps = min(args.steps or 1, 2)          # each step is a bounded sample; keep the whole run to minutes
    for _ in range(steps):
        units += reference_job(cores, t.sig, t.loc, t.taxo, t.gom_ids, idx)
    dt = time.perf_counter() - t0
    val = units / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup if args.warmup is not None else 1, "ms_per_step": dt / steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, cfg, world):
    return {"workload": f"{args.config}: synthetic {cfg['n']} genomes x {cfg['p']} PPHMMs, G={cfg['g']} GOMs, scheme PG, "
                        f"base + {cfg['b']} bootstrap replicates (PL1 loop body incl. GOM rebuild)",
            "n_genomes": cfg["n"], "n_pphmms": cfg["p"], "n_goms": cfg["g"], "n_bootstrap": cfg["b"],
            "density": cfg["density"], "scheme": "PG", "resample": "reference MT19937 indices (host) -> int32 weights",
            "l2_policy": "working set >> L2: the per-replicate pair-sum workspaces and output ring exceed 126 MB",
            "sharding": f"{world} rank(s): base matrix by upper-triangular tile blocks + NCCL all-gather; replicates by index"}


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
    ap.add_argument("--config", default=os.environ.get("GV2_BENCH_CONFIG", "cfg3"))
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--boot", type=int, default=None, help="override the number of bootstrap replicates")
    ap.add_argument("--seed", type=int, default=20241019)
    ap.add_argument("--ref-rows", type=int, default=192)
    ap.add_argument("--ref-boot", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    cfg = dict(synth.CONFIGS[args.config])
    if args.boot is not None:
        cfg["b"] = args.boot
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, cfg)
        return

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = measure_cpu_baseline(args, cfg)       # before CUDA init: it forks worker pools

    import torch
    import torch.distributed as dist
    from gravity2_b200 import _lib, gom_membership, sharding

    big = cfg["n"] * cfg["n"] * (1 + cfg["b"]) > 2e10          # cfg3-sized jobs: a step takes seconds
    steps = args.steps if args.steps is not None else (3 if big else 5)
    warmup = args.warmup if args.warmup is not None else 3
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # its banner goes to stdout, ahead of the JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    ctx = C.c_void_p()
    rc = lib.gv2_create_on_stream(local_rank, C.c_void_p(stream.cuda_stream), C.byref(ctx))
    if rc:
        raise RuntimeError(lib.gv2_last_error(None).decode())

    def ck(rc):
        _lib.check(ctx, rc)

    # ---- synthetic tables (every rank generates the same tables) and reference-stream weights
    t = synth.make_config(args.config)
    n, p, g, B = cfg["n"], cfg["p"], len(t.gom_ids), cfg["b"]
    idx = synth.reference_resample_indices(p, B, seed=args.seed)
    weights = np.stack([np.bincount(i, minlength=p) for i in idx]).astype(np.int32) if B else np.zeros((0, p), np.int32)
    gor = gom_membership(t.taxo, t.gom_ids)
    pairs = n * (n + 1) // 2

    # ---- device-resident job (PL1: the signature table doubles as the location table, quirk Q3)
    d_sig = torch.from_numpy(t.sig).cuda()
    job = C.c_void_p()
    gor_p = gor.ctypes.data_as(C.c_void_p)
    ck(lib.gv2_job_create(ctx, C.c_void_p(d_sig.data_ptr()), C.c_void_p(d_sig.data_ptr()), 1, n, p, gor_p, n, g,
                          _lib.SCHEMES["PG"], 1.0, _lib.OUT_DISTANCE, C.byref(job)))
    my_reps = sharding.replicate_indices(B, rank, world)
    d_w = torch.from_numpy(np.ascontiguousarray(weights[my_reps])).cuda() if my_reps else None
    ntiles = int(lib.gv2_num_tiles(n))
    per = sharding.tiles_per_rank(ntiles, world)
    t0_, t1_ = sharding.tile_range(ntiles, rank, world)
    tile_elems = _lib.TILE * _lib.TILE
    d_base = torch.empty((n, n), dtype=torch.float64, device="cuda")
    d_pack = torch.zeros((per, tile_elems), dtype=torch.float64, device="cuda") if world > 1 else None
    # replicate output ring: as many replicates as fit in ~8 GB
    ring = max(1, min(len(my_reps) or 1, int((8 << 30) // (n * n * 8))))
    d_out = torch.empty((ring, n, n), dtype=torch.float64, device="cuda")

    def step():
        if world > 1:
            ck(lib.gv2_job_base(job, t0_, t1_, 1, C.c_void_p(d_pack.data_ptr())))
            d_gath = sharding.all_gather_tiles(d_pack, world)
            for r in range(world):
                a, b = sharding.tile_range(ntiles, r, world)
                if b > a:
                    ck(lib.gv2_dev_unpack_tiles(ctx, n, a, b, C.c_void_p(d_gath[r * per].data_ptr()), C.c_void_p(d_base.data_ptr())))
        else:
            ck(lib.gv2_job_base(job, 0, ntiles, 0, C.c_void_p(d_base.data_ptr())))
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
    lib.gv2_timing_enable(ctx, 1)           # CUDA events around every launch of the dominant kernel, same stream
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = int(lib.gv2_launch_count(ctx)) - l0
    times = {}
    for name, kind in (("pairs", _lib.TIME_PAIRSUMS), ("tail", _lib.TIME_TAIL), ("gom", _lib.TIME_GOM)):
        k_ms, k_n = C.c_double(0.0), C.c_int64(0)
        ck(lib.gv2_timing_read_kind(ctx, kind, C.byref(k_ms), C.byref(k_n)))
        times[name] = (k_ms.value, k_n.value)
    lib.gv2_timing_enable(ctx, 0)
    path = _lib.PATH_NAMES.get(int(lib.gv2_job_replicate_path(job)), "?")
    cells = int(lib.gv2_job_copresent_cells(job))
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

    # ---- dominant kernel timed alone (roofline), rank 0
    roofline = None
    e2e = None
    peaks, peak_kind = load_peaks()
    if rank == 0:
        roofline = make_roofline(times, path, cells, steps, n, p, g, len(my_reps), ms, peaks, peak_kind)
    # ---- end to end through the public host API: release the device-resident job first (its workspaces hold ~100 GB)
    lib.gv2_job_destroy(job)
    job = None
    del d_out, d_base, d_sig
    torch.cuda.empty_cache()
    if not args.no_e2e:
        e2e = measure_e2e(lib, ctx, ck, t, gor, weights, my_reps, n, p, g, B, pairs, world, rank, dist if world > 1 else None, torch,
                          steps=max(1, min(steps, 2)))
    e2e_trees = None
    if not args.no_e2e and B > 0:
        e2e_trees = measure_e2e_trees(t, gor, weights, my_reps, n, p, g, B, pairs, world, rank, dist if world > 1 else None, torch)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": {"sparse": "48-bit fixed-point min x u8 multiplicities -> int64 (exact)",
                          "mma": "u32 fixed-point min -> u8 digit planes x u8 multiplicities -> int32 (exact)",
                          "fp32": "f32 min, f32 slab sums -> f64"}.get(path, path) +
                         "; GOM Jaccard f32 min / f32 64-column slabs -> f64; GOM dCor f64; epilogue f64",
                "data": "synthetic",
                "config": workload_config(args, cfg, world), "clocks": clocks, "gpu_launches": launches,
                "e2e": e2e, "e2e_trees": e2e_trees, "roofline": roofline, "cpu_baseline": cpu_baseline}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def make_roofline(times, path, cells, steps, n, p, g, nrep, region_ms, peaks, peak_kind):
    """Roofline of the dominant kernel (largest share of the timed region) + a line for every timed kernel group.
    Durations are CUDA-event times taken inside the timed region on the launching stream (gv2_timing_read_kind).
    Algorithmic work per step, pairs = N(N+1)/2:
      tail  (gj_fused_kernel, one launch per sub-chunk of replicates): per replicate it READS the PPHMM pair sums
            (8 B x pairs) and the fp32 GOM table (4 N G) and WRITES the N x N float64 matrix (8 N^2); its arithmetic is
            pairs x G cells of the (min, +) semiring at 2 issue slots per cell (FMNMX on the ALU pipe + FADD): bounded by
            the FP32/ALU issue rate, not by HBM -- both fractions are reported;
      pairs (replicate pair sums): tensor-core path: pairs x P x replicates x 4 planes x 2 ops against 2 x the measured
            bf16 rate; sparse path: WRITES 8 B x pairs x replicates of workspace (+ CSR and multiplicities, small);
      gom   (GOM dCor scoring, 5 kernels per 32-replicate chunk): N x G scores per replicate, reported as scores/s."""
    pairs = n * (n + 1) // 2
    hbm = peaks.get("hbm_gbs", 6550.0)
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    fp32_issue = 148 * 128 * sm_mhz * 1e6                 # thread-instructions/s over all SMSPs
    lines = {}
    ms, cnt = times["tail"]
    if cnt:
        sec = ms * 1e-3 / steps
        bytes_step = nrep * (8.0 * pairs + 8.0 * n * n + 4.0 * n * g)
        slots = nrep * float(pairs) * g * 2
        lines["tail"] = {"kernel": "gj_fused_kernel (GOM-signature (min,+) pair sums on 128x64 tiles + clean/combine/clamp/1-S/mirror epilogue)",
                         "bound": "hbm", "achieved": bytes_step / sec / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": bytes_step / sec / 1e9 / hbm,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch (10 replicates at cfg3), ncu --set full,
                         # profiles/r01_top_kernels_ncu_full.csv; algorithmic bytes of the same launch: 1.21e10
                         "traffic": 1.2491e10 if (n, g, round(nrep / max(cnt / steps, 1))) == (10000, 200, 10) else None,
                         "pipe": {"bound": "fp32/alu issue (2 slots per cell: FMNMX + FADD)", "achieved": slots / sec, "peak": fp32_issue,
                                  "unit": "thread-instr/s", "frac": slots / sec / fp32_issue},
                         "launches_per_step": cnt / steps, "avg_launch_ms": ms / cnt, "share_of_step": ms / region_ms,
                         "note": "compute bound by design: the G-column min/+ reduction costs 2 issue slots per (pair, GOM) cell; the HBM "
                                 "fraction is what is left of the N x N float64 write + pair-sum read per replicate"}
    ms, cnt = times["pairs"]
    if cnt:
        sec = ms * 1e-3 / steps
        if path == "mma":
            bf16 = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0))
            ops = float(pairs) * p * nrep * 4 * 2
            lines["pairs"] = {"kernel": "boot_mma_kernel (tcgen05.mma kind::i8, uint8 digit planes x uint8 multiplicities, TMEM int32 accumulators)",
                              "bound": "tensor", "achieved": ops / sec / 1e12, "peak": 2.0 * bf16, "unit": "TFLOP/s",
                              "frac": ops / sec / 1e12 / (2.0 * bf16), "traffic": None,
                              "peak_kind": f"{peak_kind}: 2 x bf16_tflops_sustained ({bf16}) = dense int8 TOP/s"}
        else:
            bytes_step = 8.0 * pairs * nrep
            lines["pairs"] = {"kernel": f"replicate pair sums, path {path} (boot_sparse_kernel: CSR walk, 48-bit fixed point, exact integer sums)"
                              if path == "sparse" else f"replicate pair sums, path {path}",
                              "bound": "hbm", "achieved": bytes_step / sec / 1e9, "peak": hbm, "unit": "GB/s",
                              "frac": bytes_step / sec / 1e9 / hbm, "traffic": None,
                              "copresent_cells": int(cells), "cell_replicates_per_s": cells * nrep / sec if cells else None,
                              "note": "latency bound on the L2 reads of the transposed multiplicities (one byte per list entry and replicate); "
                                      "algorithmic bytes = the fp64 pair-sum workspace it writes"}
        lines["pairs"].update({"launches_per_step": cnt / steps, "avg_launch_ms": ms / cnt, "share_of_step": ms / region_ms})
    ms, cnt = times["gom"]
    if cnt:
        sec = ms * 1e-3 / steps
        lines["gom"] = {"kernel": "GOM dCor scoring (gom_rowsums / genome_rowsums / gom_aggregates / gom_score_light / gom_score_heavy)",
                        "bound": "latency / L2 (fp64)", "achieved": float(n) * g * nrep / sec, "unit": "dCor scores/s",
                        "launches_per_step": cnt / steps, "avg_launch_ms": ms / cnt, "share_of_step": ms / region_ms}
    if not lines:
        return None
    dominant = max((k for k in lines if k != "gom"), key=lambda k: lines[k]["share_of_step"], default=None)
    out = dict(lines[dominant]) if dominant else {}
    out["peak_kind"] = out.get("peak_kind", f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs / sm_max_mhz)")
    out["kernels"] = {k: {kk: v[kk] for kk in ("kernel", "share_of_step", "avg_launch_ms", "launches_per_step", "frac") if kk in v}
                      for k, v in lines.items()}
    return out


def measure_e2e(lib, ctx, ck, t, gor, weights, my_reps, n, p, g, B, pairs, world, rank, dist, torch, steps):
    """Whole job through the package's public host API (what a pipeline_i / pipeline_ii call site uses):
    numpy tables in host memory -> gravity2_b200.engine.Job (one H2D upload of the table per step) -> base
    matrix as a numpy array -> every replicate matrix delivered in host memory (pinned ring, zero-copy
    view) to a consumer that reads it.  H2D and D2H are inside the timed region, every step."""
    import gravity2_b200
    from gravity2_b200.engine import Job
    eng = gravity2_b200.Engine(int(os.environ.get("LOCAL_RANK", "0")))
    sig = t.sig
    w_mine = np.ascontiguousarray(weights[my_reps]) if my_reps else np.zeros((0, p), np.int32)
    acc = [0.0, 0]

    def consume(b, m):              # the device->host read of the step's result: touch every delivered matrix
        acc[0] += float(m[0, -1]) + float(m[-1, 0]) + float(m[n // 2, n // 3])
        acc[1] += 1

    def job_once():
        job = Job(eng, sig, sig, gor, g, "PG", 1.0, distance=True)      # PL1: sig doubles as loc (Q3)
        try:
            if rank == 0:
                consume(-1, job.base())
            job.for_each_replicate(w_mine, consume)
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
    h2d = n * p * 8 + w_mine.nbytes + gor.nbytes
    d2h = ((1 if rank == 0 else 0) + len(my_reps)) * n * n * 8
    eng.close()
    return {"value": pairs * (1 + B) / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": dt * 1e3,
            "api": "gravity2_b200.engine.Job(sig, loc, ...).base() + .for_each_replicate(weights, consumer): numpy in, "
                   "numpy views of a pinned host ring out (per-rank bytes)"}


def measure_e2e_trees(t, gor, weights, my_reps, n, p, g, B, pairs, world, rank, dist, torch):
    """The same job with the bootstrap loops' real product as the result (app/src/pl1_graphs.py:108-117 keeps only the
    dendrogram of a replicate): numpy tables in -> Job.for_each_tree -> one Newick string per replicate out, the distance
    matrices clustered in HBM (device UPGMA, scipy's nn_chain average linkage bit for bit) and never shipped.  H2D of the
    tables and D2H of the linkage rows inside the timed region.  One step (the job takes seconds)."""
    import gravity2_b200
    from gravity2_b200.engine import Job
    eng = gravity2_b200.Engine(int(os.environ.get("LOCAL_RANK", "0")))
    names = [f"{x}_{i}" for i, x in enumerate(t.taxo)]
    w_mine = np.ascontiguousarray(weights[my_reps]) if my_reps else np.zeros((0, p), np.int32)
    acc = [0, 0]

    def consume(b, s):
        acc[0] += len(s)
        acc[1] += 1

    def job_once():
        job = Job(eng, t.sig, t.sig, gor, g, "PG", 1.0, distance=True)
        try:
            if rank == 0:
                consume(-1, gravity2_b200.DistMat2Tree(job.base(), names, "average", False, engine=eng))
            job.for_each_tree(w_mine, names, consume)
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
            "h2d_bytes_per_step": int(n * p * 8 + w_mine.nbytes + gor.nbytes + (n * n * 8 if rank == 0 else 0)),
            "d2h_bytes_per_step": int(len(my_reps) * (n - 1) * 32 + (n * n * 8 + (n - 1) * 32 if rank == 0 else 0)),
            "api": "gravity2_b200.engine.Job(sig, loc, ...).for_each_tree(weights, leaf_names, consumer) + DistMat2Tree(base): numpy in, "
                   "one Newick string per replicate out (UPGMA on the device; per-rank bytes)"}


def measure_cpu_baseline(args, cfg):
    cores = os.cpu_count() or 1
    n_s = min(cfg["n"], args.ref_rows)
    b_s = min(cfg["b"], 1)
    t = synth.make_tables(n_s, cfg["p"], min(cfg["g"], max(1, n_s // 4)), cfg["density"], synth.SEED_BASE + int(args.config[-1]))
    idx = synth.reference_resample_indices(cfg["p"], b_s, seed=args.seed)
    t0 = time.perf_counter()
    units = reference_job(cores, t.sig, t.loc, t.taxo, t.gom_ids, idx)
    dt = time.perf_counter() - t0
    return {"value": units / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_s} of {cfg['n']} genomes at full P={cfg['p']}, G={len(t.gom_ids)}, base + {b_s} replicate(s); "
                      f"numpy pair expressions of similarity_matrix_constructor.py:167-177 + dcor GOM scoring over a "
                      f"{cores}-process pool; {dt:.1f} s"}


if __name__ == "__main__":
    main()
