"""world_size-2 test of the multi-GPU slab decomposition on the CPU: the CUDA sources compiled for the emulator,
collectives supplied by torch.distributed (gloo).  The NCCL path itself is exercised on the GPU box (bench.py --gpus N)."""
import os
import subprocess
import sys

import pytest

from oracle import refcf
from tests import parity

pytestmark = pytest.mark.skipif(not refcf.available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("stepper", ["sbdf3", "cnrk2"])
def test_slab_decomposition_two_ranks(stepper):
    parity.emu_lib()  # build once, before the ranks start
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(parity.ROOT, "tests", "mp_slab_worker.py"), stepper]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, env=env, cwd=parity.ROOT)
    out = r.stdout.decode()
    assert r.returncode == 0, out[-4000:]
    assert "rank 0:" in out and "rank 1:" in out, out[-2000:]
