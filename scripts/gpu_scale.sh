#!/bin/bash
# strong-scaling lines of the C4 workload at N = 1, 2, 4, 8 (one box)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  tail -c 300 gpurun_out/scale_n$N.err | grep -i "error\|Traceback" 
done
python - <<PY
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/scale_n%d.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step %.3f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["stages"].items()})
    except Exception as e: print(n, "ERR", e)
PY
