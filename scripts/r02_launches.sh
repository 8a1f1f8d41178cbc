#!/bin/bash
# ncu launch list (per-kernel durations) of a short bench run: scripts/r02_launches.sh <tag> <bench args...>
mkdir -p gpurun_out
T=$1; shift
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/${T}_launches_bench.log 2>&1
echo "ncu exit $?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${T}_launches.csv")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)  # -> us
    k = r[ki][:60]
    a = tot.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%4d total %10.1f us  avg %9.1f us" % (k, n, t, t / n))
PY
