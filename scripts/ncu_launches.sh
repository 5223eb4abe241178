#!/bin/bash
# launch list of one bench step, kernels issued eagerly (--eager: same kernels as the graphed default, countable per step)
# (cold-cache, serialised: compare SHARES) -> gpurun_out/launches_$TAG.csv + gpurun_out/launch_summary_$TAG.txt
# usage: scripts/ncu_launches.sh TAG [extra bench.py flags, e.g. --workload B]
TAG=${1:-r2}
shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --eager --steps 1 --warmup 3 --no-cpu-baseline --no-extras "$@" > gpurun_out/ncu_bench_$TAG.log 2>&1
python - <<PY
import csv, collections, sys
rows = list(csv.reader(open("gpurun_out/launches_$TAG.csv", errors="ignore")))
hdr = None; agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    k = d["Kernel Name"].split("(")[0][:70]
    agg[k][0] += 1; agg[k][1] += v
NS = max(1, agg.get("vds::adamw_kernel", [1])[0])   # one AdamW launch per step: the number of (identical) steps profiled
for k in agg: agg[k][1] /= NS; agg[k][0] /= NS
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/launch_summary_$TAG.txt", "w") as f:
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"{t/1e3:9.3f} ms {100*t/tot:5.1f}% n={n:7.1f} avg={t/n:8.1f} us  {k}"
        print(line); f.write(line + "\n")
    f.write(f"total {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches / step ({NS} steps profiled)\n")
print("total ms", tot / 1e3)
PY
