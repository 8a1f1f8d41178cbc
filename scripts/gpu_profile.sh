#!/bin/bash
# ncu evidence for profiles/: launch list of two steps + one full capture per hot kernel (C4 workload, 1 GPU).
# The .ncu-rep files stay on the box (too large); their raw metric pages and source pages come back as csv(.gz).
TAG=${1:-r01}
KEEP=${2:-none}
mkdir -p gpurun_out /tmp/rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
for k in ${KERNELS:-zpass tau_solve ygemm xpass_inverse xpass_forward}; do
  SKIP=2; if [ "$k" = "tau_solve" ]; then SKIP=6; fi   # skip the six 4-term SMRK2 start-up solves: capture the SBDF3 kernel
  ncu --set full --clock-control none --import-source on -k regex:$k -s $SKIP -c 2 -f -o /tmp/rep/${TAG}_full_$k python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_full_$k.log 2>&1
  ncu -i /tmp/rep/${TAG}_full_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_$k.csv 2>/dev/null
  ncu -i /tmp/rep/${TAG}_full_$k.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_source_$k.csv.gz
  if [ "$k" = "$KEEP" ]; then cp /tmp/rep/${TAG}_full_$k.ncu-rep gpurun_out/; fi
done
ls -la gpurun_out | tail -20
