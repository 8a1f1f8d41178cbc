"""Generates tests/golden/mp_slab_48x49x32.npz: the multi-GPU parity case of bench.py --gpus N and of
tests/mp_slab_worker.py (48x49x32, plane Couette, SBDF3, rotational, 4 steps), computed by the oracle = the unmodified
reference compiled in oracle/_ref.  Run once in the build container (the oracle does not have to exist where the
fixture is consumed; bench.py never executes oracle/ for this check).

Stored: the retained (de-aliased) box of the initial field and of the field after 4 steps as complex arrays
[3][Ny][2Kx+1][Kz+1] (kx rows in the order 0..Kx, -Kx..-1), CFL before and after, L2Norm after, the flags.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcf  # noqa: E402
from tests import parity  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG = dict(parity.C1, Nx=48, Ny=49, Nz=32)
NSTEPS = 4


def box(arr, Nx, Nz):
    Kx, Kz = Nx // 3 - 1, Nz // 3 - 1
    c = arr.view(np.complex128)
    rows = [(m if m <= Kx else m - (2 * Kx + 1)) % Nx for m in range(2 * Kx + 1)]
    return np.ascontiguousarray(c[:, :, rows, :Kz + 1])


def main():
    ur = parity.ref_random(CFG, 1)
    u0 = box(ur.data.copy(), CFG["Nx"], CFG["Nz"])
    rd = refcf.RefDNS(ur, refcf.make_flags(**CFG["flags"]))
    cfl0 = rd.cfl()
    rd.advance(NSTEPS)
    u1, _ = rd.get()
    np.savez_compressed(os.path.join(OUT, "mp_slab_48x49x32.npz"), u0=u0, u4=box(u1.data, CFG["Nx"], CFG["Nz"]),
                        cfl0=cfl0, cfl4=rd.cfl(), norm4=u1.l2norm(), nsteps=NSTEPS,
                        Nx=CFG["Nx"], Ny=CFG["Ny"], Nz=CFG["Nz"], Lx=CFG["Lx"], Lz=CFG["Lz"], a=CFG["a"], b=CFG["b"],
                        flags=np.array(repr(sorted(CFG["flags"].items()))))
    print("wrote mp_slab_48x49x32.npz", u0.shape, cfl0, rd.cfl(), u1.l2norm())


if __name__ == "__main__":
    main()
