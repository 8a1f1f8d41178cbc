"""Worker of the world_size-2 CPU test of the slab decomposition (launched by torch.distributed.run, gloo backend).

Each rank loads the CPU-emulation build of the CUDA sources, installs host-side collectives (gloo) through
cfgpu_comm_init_external, advances the same initial field and checks its own kx rows -- and the all-gathered field --
against the single-process oracle (the compiled reference)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import channelflow_b200 as cf  # noqa: E402
from oracle import refcf  # noqa: E402
from tests import parity  # noqa: E402


def main():
    os.environ["CFGPU_DEVICE"] = "0"  # the emulator has one "device"; on a GPU box the default is LOCAL_RANK
    use_gpu = os.environ.get("CF_WORKER_BACKEND") == "nccl"  # GPU box: the product library over NCCL
    if use_gpu:
        os.environ["CFGPU_DEVICE"] = os.environ.get("LOCAL_RANK", "0")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if use_gpu:
        lib = parity.gpu_lib()
        ids = [lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        lib.comm_init_nccl(rank, world, ids[0])
    else:
        lib = parity.emu_lib() if rank == 0 else None
        dist.barrier()
        if lib is None:
            lib = parity.emu_lib()

    def exchange(sends, recvs):
        reqs = []
        for peer, buf in recvs:
            if buf.size:
                reqs.append(dist.irecv(torch.from_numpy(buf), src=int(peer)))
        for peer, buf in sends:
            if buf.size:
                reqs.append(dist.isend(torch.from_numpy(buf), dst=int(peer)))
        for r in reqs:
            r.wait()

    def allreduce(buf, op):
        dist.all_reduce(torch.from_numpy(buf), op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM)

    if not use_gpu:
        lib.comm_init_external(rank, world, exchange, allreduce)

    cfg = dict(parity.C1, Nx=16, Ny=17, Nz=12) if not use_gpu else dict(parity.C1, Nx=48, Ny=49, Nz=32)
    stepper = sys.argv[1] if len(sys.argv) > 1 else "sbdf3"
    fl = dict(cfg["flags"], timestepping=stepper)
    ur = parity.ref_random(cfg, 1)
    rd = refcf.RefDNS(ur, refcf.make_flags(**fl))
    gd = cf.DNS(parity.to_gpu(lib, ur), cf.make_flags(**fl))
    out = {"cfl0": abs(gd.cfl() - rd.cfl()) / abs(rd.cfl())}
    rd.advance(4)
    gd.advance(4)
    u1, _ = rd.get()
    u2, _ = gd.get()
    out["norm"] = abs(u2.l2norm() - u1.l2norm()) / u1.l2norm()
    out["cfl"] = abs(gd.cfl() - rd.cfl()) / abs(rd.cfl())
    Kx, Kz = cfg["Nx"] // 3 - 1, cfg["Nz"] // 3 - 1
    x0, x1, y0, y1 = lib.comm_ranges(2 * Kx + 1, cfg["Ny"], rank)
    mine = u2.get().view(np.complex128)
    ref = u1.data.view(np.complex128)
    rows = [(m if m <= Kx else m - (2 * Kx + 1)) % cfg["Nx"] for m in range(x0, x1)]
    out["own_rows"] = float(np.abs(mine[:, :, rows, :Kz + 1] - ref[:, :, rows, :Kz + 1]).max() / np.abs(ref).max())
    u2.allgather()
    u2.set_padded(False)  # full download (a padded download only fetches this rank's rows)
    full = u2.get().view(np.complex128)
    allrows = [(m if m <= Kx else m - (2 * Kx + 1)) % cfg["Nx"] for m in range(2 * Kx + 1)]
    out["gathered"] = float(np.abs(full[:, :, allrows, :Kz + 1] - ref[:, :, allrows, :Kz + 1]).max() / np.abs(ref).max())
    out["ranges"] = (x0, x1, y0, y1)
    bad = [k for k, v in out.items() if k != "ranges" and not v < 1e-12]
    print("rank %d: %s" % (rank, out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
