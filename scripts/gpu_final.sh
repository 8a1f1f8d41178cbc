#!/bin/bash
# Evidence run for profiles/: GPU tests, default bench line (+ c1/c2/c5), reference arm, ncu launch list + full captures.
TAG=${1:-r01f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
for w in c1 c2 c5; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python - <<PY
import json
for w in ("default","c1","c2","c5","reference"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, "ms/step", d.get("ms_per_step"), "value %.4g"%d["value"], "e2e", d.get("e2e") and d["e2e"].get("ms_per_step"), d.get("roofline"), {k:round(v["ms_per_step"],4) for k,v in d.get("stages",{}).items()})
    except Exception as e: print(w, "ERR", e)
PY
scripts/gpu_profile.sh $TAG > gpurun_out/${TAG}_profile.log 2>&1
tail -3 gpurun_out/${TAG}_profile.log
