#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/profile_select.py 2>&1 | tail -5 | tee gpurun_out/r2h_select.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_select_launches.csv python tools/profile_select.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2h_select_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; k = H.index("Kernel Name"); v = H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= v: continue
    name = r[k].split("(")[0][-60:]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[v].replace(",", ""))
for n, (c, t) in agg.items():
    print(f"{n:62s} x{c:4d}  total {t/1e3:9.1f} us  avg {t/c/1e3:8.2f} us")
PY
