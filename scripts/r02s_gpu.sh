#!/bin/bash
# scaling lines of the C4 workload at N = $1 GPUs: strong, and weak when $2 = weak
N=${1:-8}
T=${3:-r02s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${T}_gpus_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_n${N}_strong.json 2> gpurun_out/${T}_bench_n${N}_strong.err; echo "strong n$N exit $?"
if [ "$2" = "weak" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29560+N)) bench.py --gpus $N --steps 20 --warmup 5 --scaling weak --no-e2e > gpurun_out/${T}_bench_n${N}_weak.json 2> gpurun_out/${T}_bench_n${N}_weak.err; echo "weak n$N exit $?"
fi
python scripts/print_bench.py gpurun_out/${T}_bench_n${N}_*.json 2>&1 | cut -c1-1200
tail -n 3 gpurun_out/${T}_bench_n${N}_strong.err
