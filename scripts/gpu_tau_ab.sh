#!/bin/bash
# layout A/B on the C4 bench (tile-major hot-path fields vs CFGPU_SERIAL_LAYOUT=1), then GPU tests.
mkdir -p gpurun_out
run() {  # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload c4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/tau_$tag.json 2> gpurun_out/tau_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/tau_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "ms/step %.3f"%d["ms_per_step"], {k:round(x["ms_per_step"],3) for k,x in d["stages"].items()})
except Exception as e:
    print("$tag ERR", e); print(open("gpurun_out/tau_$tag.err").read()[-800:])
PY
}
run tile X=1
run serial CFGPU_SERIAL_LAYOUT=1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
