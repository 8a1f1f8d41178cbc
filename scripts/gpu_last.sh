#!/bin/bash
# closing check: GPU tests, default bench line, fresh capture of the x-pass inverse kernel
TAG=${1:-r01h}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -2 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print("default ms/step", d["ms_per_step"], "value %.4g"%d["value"], "e2e", d["e2e"]["ms_per_step"], d["roofline"], {k:round(v["ms_per_step"],4) for k,v in d["stages"].items()})
PY
mkdir -p /tmp/rep
k=xpass_inverse
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -f -o /tmp/rep/${TAG}_full_$k python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_full_$k.log 2>&1
ncu -i /tmp/rep/${TAG}_full_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_$k.csv 2>/dev/null
ncu -i /tmp/rep/${TAG}_full_$k.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source_$k.csv.gz
ls -la gpurun_out/${TAG}_*
