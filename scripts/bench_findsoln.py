#!/usr/bin/env python
"""C3 of BASELINE.json: Newton-Krylov-hookstep search (GMRES + hookstep, one DNS integration per Krylov vector) for the
stored plane-Couette solution at Re 400 on the 48x49x48 grid, everything device resident (host/devicesearch.cpp).
The initial guess is tests/golden/eq.npz (24x33x24) zero-padded in spectral space to 48x49x48 and scaled by 1.001.
Prints one JSON line: DNS integrations/s, time steps/s, Newton convergence history."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import channelflow_b200 as cf  # noqa: E402
from tests import parity  # noqa: E402


def main():
    lib = parity.gpu_lib()
    _, ur = parity.load_eq(lib)
    Nx, Ny, Nz = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (48, 49, 48)))
    c = ur.data.view(np.complex128)                      # [3][Ny][Nx][Mz]
    big = np.zeros((3, Ny, Nx, Nz // 2 + 1), np.complex128)
    Kx, Kz = ur.Nx // 3 - 1, ur.Nz // 3 - 1
    big[:, :ur.Ny, :Kx + 1, :Kz + 1] = c[:, :, :Kx + 1, :Kz + 1]
    big[:, :ur.Ny, Nx - Kx:, :Kz + 1] = c[:, :, ur.Nx - Kx:, :Kz + 1]
    ug = cf.FlowField(lib, Nx, Ny, Nz, 3, ur.Lx, ur.Lz, ur.a, ur.b).set(big.view(np.float64), padded=True)
    ug.scale(1.001)
    fl = dict(parity.C1["flags"])
    l0 = lib.launch_count()
    t0 = time.perf_counter()
    r = cf.hookstep_search(ug, cf.make_flags(**fl), 10.0, 0.03125, sigma=parity.EQ_SIGMA, Nnewton=8, epsSearch=1e-11, xrelative=True)
    sec = time.perf_counter() - t0
    steps = r["fevals"] * r["steps_per_eval"]
    print(json.dumps({"metric": "findsoln_dns_integrations_per_s", "value": r["fevals"] / sec, "unit": "integrations/s (T = 10, %d steps each)" % r["steps_per_eval"],
                      "time_steps_per_s": steps / sec, "grid": [Nx, Ny, Nz], "seconds": sec, "newton_steps": r["newton_steps"], "fevals": r["fevals"],
                      "gmres_iterations": r["gmres_iterations"], "residual_history": r["history"], "converged_below_1e-10": r["residual"] <= 1e-10,
                      "ax": r["ax"], "gpu_launches": lib.launch_count() - l0, "data": "tests/golden/eq.npz padded to the C3 grid, x 1.001"}))


if __name__ == "__main__":
    main()
