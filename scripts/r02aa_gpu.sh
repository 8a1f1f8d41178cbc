#!/bin/bash
# tau setup with the staged factor write-out: setup time, tau-solver / stepper parity, launch list of a CNRK2 step
mkdir -p gpurun_out
export T=${1:-r02aa}
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -k "tau or stepper or sbdf or cnrk or variable or helmholtz or golden" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench exit $?"
python - <<'P'
import json
l=json.loads(open('gpurun_out/%s_bench_c4.json' % __import__('os').environ.get('T','r02aa')).read().strip().splitlines()[-1])
print(l['ms_per_step'], l.get('tau_setup'))
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_cnrk2.csv python bench.py --workload c4 --stepper cnrk2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1; echo "launch list exit $?"
grep -E "tau_factor|tau_profiles" gpurun_out/${T}_launches_cnrk2.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | head
