#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for w in c4 c1; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -c 600 gpurun_out/bench_$w.err
done
python - <<'PY'
import json
for w in ("c4","c1"):
    try:
        d=json.load(open("gpurun_out/bench_%s.json"%w))
        print(w, "ms/step", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],4) for k,v in d["stages"].items()}, d["cpu_baseline"])
    except Exception as e: print(w, "ERR", e)
PY
if [ -n "$NCU_KERNEL" ]; then
ncu --set full --clock-control none --import-source on -k regex:$NCU_KERNEL -s 2 -c 2 -f -o gpurun_out/prof_${NCU_TAG:-k} python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${NCU_TAG:-k}.log 2>&1
fi
