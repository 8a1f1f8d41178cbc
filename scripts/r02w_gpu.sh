#!/bin/bash
N=${1:-8}
T=${2:-r02s}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29560+N)) bench.py --gpus $N --steps 10 --warmup 3 --scaling weak --no-e2e > gpurun_out/${T}_bench_n${N}_weak.json 2> gpurun_out/${T}_bench_n${N}_weak.err; echo "weak n$N exit $?"
python scripts/print_bench.py gpurun_out/${T}_bench_n${N}_weak.json 2>&1 | cut -c1-1200
grep -v "^\[W\|^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/${T}_bench_n${N}_weak.err | tail -n 4 | cut -c1-300
