#!/bin/bash
# A/B of an environment knob on the C4 bench: scripts/gpu_ab.sh VAR val1 val2 ...
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v python bench.py --workload c4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("$VAR=$v", "ms/step %.3f"%d["ms_per_step"], {k:round(x["ms_per_step"],3) for k,x in d["stages"].items()})
PY
done
