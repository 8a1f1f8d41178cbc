#!/bin/bash
# N-GPU round: GPU parity tests (the NCCL slab test runs with >= 2 GPUs), bench with the multi-GPU parity block in the
# exchange modes:  scripts/r02_gpu2.sh <tag> <ngpus> [modes...]
mkdir -p gpurun_out
T=${1:-r02a}; N=${2:-2}; shift; shift
MODES=${@:-push fused}
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
if [ -z "$SKIP_PYTEST" ]; then
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu_${N}gpus.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu_${N}gpus.log
tail -4 gpurun_out/${T}_pytest_gpu_${N}gpus.log
fi
for m in $MODES; do
  EXTRA=""
  if [ "$m" = "staged" ]; then export CFGPU_NO_PEER=1; else unset CFGPU_NO_PEER; export CFGPU_PEER_MODE=$m; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $BENCH_EXTRA > gpurun_out/${T}_bench_n${N}_$m.json 2> gpurun_out/${T}_bench_n${N}_$m.err; echo "bench n$N $m exit $?"
  tail -n 4 gpurun_out/${T}_bench_n${N}_$m.err
done
unset CFGPU_NO_PEER CFGPU_PEER_MODE
python scripts/print_bench.py gpurun_out/${T}_bench_n${N}_*.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_n${N}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "parity:", json.dumps(d.get("multi_gpu_parity")))
    except Exception as e:
        print(f, "ERR", e)
PY
