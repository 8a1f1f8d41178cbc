#!/bin/bash
# N-GPU bench lines (N = $1) + the 1-GPU line for comparison
N=${1:-2}
mkdir -p gpurun_out
python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err
tail -c 1500 gpurun_out/bench_c4_n$N.err
python - <<PY
import json
for n in (1,$N):
    try:
        d=json.loads(open("gpurun_out/bench_c4_n%d.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["stages"].items()})
    except Exception as e: print(n, "ERR", e)
PY
