"""CPU tests of the oracle itself (the unmodified reference compiled against the FFTW/Eigen shims): it must
reproduce the reference's own golden vectors and known-answer tests before anything is compared against it."""
import os
import subprocess

import numpy as np
import pytest

from oracle import refcf
from tests import parity

pytestmark = pytest.mark.skipif(not refcf.available(), reason="oracle/_ref not built (make -C oracle)")

REFBIN = os.path.join(parity.ROOT, "oracle", "_ref", "bin")


@pytest.mark.parametrize("prog", ["tridiagTest", "chebyTest", "helmholtzTest", "tausolverTest", "poissonTest", "laminarTest"])
def test_reference_unit_programs_pass_on_shim(prog, tmp_path):
    """The reference's own test programs (tests/*.cpp, own tolerances 1e-12..1e-9) linked against our FFTW shim."""
    exe = os.path.join(REFBIN, prog)
    if not os.path.exists(exe):
        pytest.skip("reference test binaries not built (make -C oracle reftests)")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, cwd=str(tmp_path))  # some write *.asc dumps
    assert r.returncode == 0, r.stdout.decode()[-2000:]


def test_golden_pair_first_unit():
    """timeIntegrationTest.cpp fixture: the loader reproduces |uinit|, div u = 0, no-slip (full 440 steps: slow test)."""
    g = np.load(os.path.join(parity.GOLDEN, "golden_pair.npz"))
    u = refcf.RefField(int(g["Nx"]), int(g["Ny"]), int(g["Nz"]), 3, float(g["Lx"]), float(g["Lz"]), float(g["a"]),
                       float(g["b"])).load_padded_physical(g["uinit"])
    assert abs(u.l2norm() - 0.2214413597357669) < 1e-13
    assert u.divnorm() < 1e-14 and u.bcnorm() < 1e-14


@pytest.mark.slow
def test_golden_pair_full():
    """uinit -> 440 SBDF3 steps with the compiled reference + shim -> ufinal, tolerance 1e-13 as in the reference."""
    if not os.environ.get("CF_SLOW"):
        pytest.skip("set CF_SLOW=1 (takes ~70 s); last run: L2Dist = 1.9e-14")
    g = np.load(os.path.join(parity.GOLDEN, "golden_pair.npz"))
    u = refcf.RefField(int(g["Nx"]), int(g["Ny"]), int(g["Nz"]), 3, float(g["Lx"]), float(g["Lz"]), float(g["a"]),
                       float(g["b"])).load_padded_physical(g["uinit"])
    v = u.like().load_padded_physical(g["ufinal"])
    dns = refcf.RefDNS(u, refcf.make_flags(nu=1 / 400, Vsuck=1 / 400, dt=1 / 40, baseflow="suction"))
    for _ in range(11):
        dns.cfl()
        dns.advance(40)
    uf, _ = dns.get()
    assert v.l2dist(uf) <= 1e-13


def test_cheby_roundtrip_shim():
    c = np.random.default_rng(0).standard_normal(33)
    p = refcf.cheby_make_physical(c)
    y = np.cos(np.pi * np.arange(33) / 32)
    direct = np.polynomial.chebyshev.chebval(y, c)
    assert np.abs(p - direct).max() < 1e-13
    assert np.abs(refcf.cheby_make_spectral(p) - c).max() < 1e-14
