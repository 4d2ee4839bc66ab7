#!/usr/bin/env python
"""gpurun_out/checks_*.json + parity_table.json (written by tests/run_gpu_checks.py on the B200) -> profiles/r2_parity_table.md:
every check group with its worst error against the ONE tolerance table of tests/checks.py, and the per-tensor gradient errors
of the full-size (benchmarked-shape) cases.   python tools/make_parity_md.py gpurun_out/checks_r2f.json gpurun_out/parity_table.json"""
import json
import sys

checks = json.load(open(sys.argv[1]))
table = json.load(open(sys.argv[2]))
out = ["# Round 2 parity table (B200, `python tests/run_gpu_checks.py`, TF32 mode)", "",
       f"{len(checks['results'])} checks, {checks['failed']} failed.  Tolerance classes are defined once at the top of `tests/checks.py` "
       "(TOL_F32 2e-5, TOL_TC 1e-3, TOL_TC2 2e-3, TOL_NET 3e-3, KINK_TOL 2.5e-2, NET_GRAD 1e-2 global / 5e-2 per tensor; integers exact).", "",
       "## Worst error per check group", "", "| group | checks | worst err / tol (check) |", "|---|---:|---|"]
groups = {}
for r in checks["results"]:
    groups.setdefault(r["group"], []).append(r)
for g, rows in groups.items():
    rows = [r for r in rows if r["err"] is not None and r["tol"]]
    worst = max(rows, key=lambda r: r["err"] / r["tol"]) if rows else None
    out.append(f"| {g} | {len(groups[g])} | " + (f"{worst['err']:.2e} / {worst['tol']:.1e} ({worst['name'][:90]}) |" if worst else "- |"))
for tag, t in table.items():
    out += ["", f"## {tag}: per-tensor gradient error of the graph-replayed step vs the CPU oracle (rel-L2) and vs the reference's |grad|", ""]
    if "losses" in t:
        out += ["| loss | ours | oracle | reference (golden) |", "|---|---:|---:|---:|"]
        out += [f"| {k} | {v['ours']:.6g} | {v['oracle']:.6g} | {v['reference']:.6g} |" for k, v in t["losses"].items()]
        out += ["", "forward tensors (rel-L2 vs oracle): " + ", ".join(f"{k} {v:.2e}" for k, v in t["forward"].items()), ""]
    if "loss" in t:
        out += [f"loss: ours {t['loss']['ours']:.6g}, oracle {t['loss']['oracle']:.6g}, reference {t['loss']['reference']:.6g}", ""]
    fl = t.get("dispatch_flops", {})
    tot = sum(fl.values()) or 1.0
    out += ["contraction flops by kernel family: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in fl.items() if v), ""]
    for key in ("grad_d", "grad_g", "grad"):
        if key not in t:
            continue
        rows = sorted(t[key].items(), key=lambda kv: -kv[1]["rel_l2"])
        import statistics
        rels = [v["rel_l2"] for _, v in rows]
        out += [f"### {key}: {len(rows)} tensors, median rel-L2 {statistics.median(rels):.2e}, 90th pct {sorted(rels)[int(0.9 * len(rels))]:.2e}, max {rels[0]:.2e}", "",
                "| tensor (10 worst) | rel-L2 vs oracle | ours / reference norm |", "|---|---:|---:|"]
        out += [f"| {n} | {v['rel_l2']:.2e} | {v['norm'] / (v['norm_reference'] + 1e-30):.5f} |" for n, v in rows[:10]]
        out.append("")
open("profiles/r2_parity_table.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
