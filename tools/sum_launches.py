"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) per kernel name.
  python tools/sum_launches.py gpurun_out/launches.csv [top_n]"""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1], errors="ignore") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"^void ", "", name)
    name = name.split("<unnamed>::")[-1] if "<unnamed>::" in name else name
    v = float(r[vi].replace(",", "")); u = r[ui]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(u, 1)
    agg[name][0] += 1; agg[name][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}% {n:6d}x  {k[:110]}")
