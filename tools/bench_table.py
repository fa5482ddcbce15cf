"""Markdown table of bench.py JSON lines: python tools/bench_table.py label=file.json ... > profiles/rNN_bench_lines.md"""
import json
import sys

rows = []
for arg in sys.argv[1:]:
    label, path = arg.split("=", 1)
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        rows.append(f"| {label} | (no line: {e}) | | | | | |")
        continue
    e2e = d.get("e2e") or {}
    r = d.get("roofline") or {}
    cb = d.get("cpu_baseline") or {}
    pc = d.get("parity_check") or {}
    kern = "; ".join(f"{k} {v.get('avg_launch_ms', 0):.2f} ms x {v.get('launches_per_step', 0):.0f}"
                     + (f" ({v['bound']} {v['frac']:.3f})" if v.get("frac") is not None else "")
                     for k, v in (r.get("kernels") or {}).items())
    rows.append(f"| {label} | {d.get('n_gpus')} | {d['value']:.4g} | {d['ms_per_step']:.1f} | "
                f"{e2e.get('value', 0):.4g} ({e2e.get('ms_per_step', 0):.0f} ms) | "
                f"{r.get('bound', '')} {r.get('frac', 0):.3f} | {kern} | {cb.get('value', 0):.4g} ({cb.get('cores', '')} cores) | "
                f"{pc.get('ok', '')} |")
print("| line | GPUs | value (" + "pair-similarities/s" + ") | ms/step | e2e value (time) | roofline of the dominant kernel | kernel groups: avg launch x launches/step (bound, fraction) | CPU port | parity_check |")
print("|---|---|---|---|---|---|---|---|---|")
print("\n".join(rows))
