"""Generates the fixtures under tests/golden/ from the reference's own test data (run once in the build container;
/root/reference does not exist on the GPU box).

  golden_pair.npz : tests/data/uinit.nc -> ufinal.nc of timeIntegrationTest.cpp (48x35x48, ASBL, SBDF3, dt=1/40,
                    440 steps, tolerance 1e-13): physical-space velocity on the I/O grid + grid attributes.
  eq.npz          : tests/data/eq.nc (24x33x24 equilibrium used by findsolnTest.cpp)
  os_eig.npz      : Orr-Sommerfeld eigen data tests/data/os_{ueig,peig}10_65.asc, os_omega10_65.cmplx
  eq.nc           : tests/data/eq.nc unchanged (a NetCDF-4 file as stock Channelflow writes it)
  couette_ref.txt : what the reference's examples/couette.cpp prints (compiled reference, oracle/_ref/bin/couette)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nc4mini  # noqa: E402

REF = "/root/reference/tests/data"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    a0, d0 = nc4mini.read_nc(os.path.join(REF, "uinit.nc"))
    a1, d1 = nc4mini.read_nc(os.path.join(REF, "ufinal.nc"))
    assert a0 == a1
    np.savez_compressed(os.path.join(OUT, "golden_pair.npz"), uinit=d0, ufinal=d1, **{k: np.array(v) for k, v in a0.items()})
    ae, de = nc4mini.read_nc(os.path.join(REF, "eq.nc"))
    np.savez_compressed(os.path.join(OUT, "eq.npz"), u=de, **{k: np.array(v) for k, v in ae.items()})

    def read_cplx_asc(path):
        rows = [l.split() for l in open(path) if l.strip() and not l.startswith("%")]
        return np.array([[float(x) for x in r] for r in rows])
    ueig = read_cplx_asc(os.path.join(REF, "os_ueig10_65.asc"))
    peig = read_cplx_asc(os.path.join(REF, "os_peig10_65.asc"))
    om = open(os.path.join(REF, "os_omega10_65.cmplx")).read().replace("(", " ").replace(")", " ").replace(",", " ").split()
    np.savez_compressed(os.path.join(OUT, "os_eig.npz"), ueig=ueig, peig=peig, omega=np.array([float(om[0]), float(om[1])]))
    # the file itself, byte for byte, as input of the C++ NetCDF-4 reader test (host/ncfile.cpp)
    import shutil
    shutil.copyfile(os.path.join(REF, "eq.nc"), os.path.join(OUT, "eq.nc"))
    # stdout of the reference's examples/couette.cpp (oracle/_ref/bin/couette, `make -C oracle reftests`): t, CFL, L2Norm(u), ...
    # every time unit of a 30-unit run, 6 significant digits
    import subprocess, tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "couette")
    if os.path.exists(exe):
        with tempfile.TemporaryDirectory() as td:
            out = subprocess.run([exe], cwd=td, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        open(os.path.join(OUT, "couette_ref.txt"), "w").write(out)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
