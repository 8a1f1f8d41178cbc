#!/bin/bash
# one strong-scaling line of the C4 workload at N = $1 GPUs
N=${1:-4}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
print($N, "ms/step %.3f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["stages"].items()})
PY
