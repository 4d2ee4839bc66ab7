#!/usr/bin/env python
"""dram bytes per launch of the kernels of an `ncu --set full` capture -> profiles/r2_ncu_traffic.json (read by bench.py).

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv
  python tools/ncu_traffic.py gpurun_out/prof_raw.csv [key=kernel-name-substring ...]

Each `key=substr` selects the launches whose kernel name contains `substr`; the heaviest one (largest duration) is
reported under `key`; `key=substr@n` takes the n-th matching launch (0-based, launch order) instead.  Metrics: dram__bytes_read.sum + dram__bytes_write.sum, gpu__time_duration.sum, tensor-pipe and
DRAM utilisation."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1], errors="ignore") if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    sel = dict(a.split("=", 1) for a in sys.argv[2:]) or {"gemm_tma_kernel": "gemm_tma_kernel"}
    unit = lambda m: units[col[m]] if m in col else ""
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    out = {}
    for key, sub in sel.items():
        best = None
        nth = None
        if "@" in sub:
            sub, nth = sub.rsplit("@", 1)
            nth = int(nth)
        seen = 0
        for r in rows[2:]:
            if sub not in r[col["Kernel Name"]]:
                continue
            dur = fnum(r[col["gpu__time_duration.sum"]]) * scale.get(unit("gpu__time_duration.sum"), 1.0)
            if nth is not None:
                if seen == nth:
                    best = (dur, r)
                seen += 1
            elif best is None or dur > best[0]:
                best = (dur, r)
        if best is None:
            continue
        dur, r = best
        rd = fnum(r[col["dram__bytes_read.sum"]]) * scale.get(unit("dram__bytes_read.sum"), 1)
        wr = fnum(r[col["dram__bytes_write.sum"]]) * scale.get(unit("dram__bytes_write.sum"), 1)
        rec = dict(kernel=r[col["Kernel Name"]][:120], duration_us=dur, dram_bytes_per_launch=rd + wr, dram_read=rd, dram_write=wr)
        for m in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                  "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                  "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"):
            if m in col:
                rec[m] = fnum(r[col[m]])
        out[key] = rec
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(out)
    json.dump(old, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
