#!/bin/bash
# round 2, first GPU call: GPU parity tests on a 2-GPU box (the NCCL slab test runs), bench on 2 GPUs with the
# multi-GPU parity block, bench on 1 GPU
mkdir -p gpurun_out
T=${1:-r02a}
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu_2gpus.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu_2gpus.log
tail -4 gpurun_out/${T}_pytest_gpu_2gpus.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 exit $?"
tail -c 1500 gpurun_out/${T}_bench_n2.json; tail -5 gpurun_out/${T}_bench_n2.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 exit $?"
tail -c 1500 gpurun_out/${T}_bench_n1.json; tail -5 gpurun_out/${T}_bench_n1.err
