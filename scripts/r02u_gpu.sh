#!/bin/bash
# in-place forward x-pass: parity at a long-x grid + A/B timing at Nx = 512 / 1024 / 2048 on one GPU (weak-scaling shapes use 1/N of the y planes; here whole grids of smaller Ny)
mkdir -p gpurun_out
python - <<'PY'
import os, sys, time, json
sys.path.insert(0, ".")
import numpy as np
import channelflow_b200 as cf
from tests import parity
res = {}
for Nx in (512, 1024, 2048):
    for ip in ("0", "1"):
        os.environ["CF_XPF_INPLACE"] = ip
        import subprocess
        out = subprocess.run([sys.executable, "-c", """
import sys, json; sys.path.insert(0, '.')
import numpy as np
import channelflow_b200 as cf
from tests import parity
lib = parity.gpu_lib()
cfg = dict(parity.C1, Nx=%d, Ny=65, Nz=512, Lx=4*np.pi*%d/512, Lz=2*np.pi)
u = cf.randomfield(lib, cfg['Nx'], cfg['Ny'], cfg['Nz'], cfg['Lx'], cfg['Lz'], seed=1, magn=0.2, smooth=0.4)
fl = dict(cfg['flags']); fl['dt'] = 0.002; fl['nu'] = 1/4000.
d = cf.DNS(u, cf.make_flags(**fl)); d.advance(5)
lib.profile_enable(True); lib.profile_read(reset=True)
lib.timer_start(); d.advance(10); ms = lib.timer_stop()
st, calls = lib.profile_read(reset=True)
uu, _ = d.get()
print(json.dumps(dict(ms_per_step=ms/10, fwd_x=st[3]/10, inv_x=st[1]/10, norm=uu.l2norm())))
""" % (Nx, Nx)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            res[(Nx, ip)] = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:
            res[(Nx, ip)] = out.stderr[-400:]
        print(Nx, "inplace=" + ip, res[(Nx, ip)], flush=True)
PY
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02u_bench_c4.json 2> gpurun_out/r02u_bench_c4.err; echo "bench exit $?"
python scripts/print_bench.py gpurun_out/r02u_bench_c4.json; tail -2 gpurun_out/r02u_bench_c4.err
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02u_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02u_pytest_gpu.log; tail -4 gpurun_out/r02u_pytest_gpu.log
