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


# ------------------------------------------------------------------------------------------------ NumPy restatement
SMALLG = dict(parity.C1, Nx=12, Ny=17, Nz=12)


def test_np_oracle_helmholtz_and_tausolver():
    """oracle/np_oracle.py against the compiled reference: HelmholtzSolver and TauSolver, mode by mode."""
    from oracle import np_oracle as npo
    rng = np.random.default_rng(7)
    N, a, b, nu = 17, -1.0, 1.0, 1 / 400.0
    f = rng.standard_normal(N)
    for lam in (0.7, 91.6):
        h = npo.Helmholtz(N, a, b, lam, nu).solve(f, 0.3, -0.2)
        assert np.abs(h - refcf.helmholtz(N, a, b, lam, nu, f, 0.3, -0.2)).max() < 1e-13
    for kx, kz in ((0, 0), (1, 0), (0, 2), (-2, 1)):
        R = [rng.standard_normal(N) + 1j * rng.standard_normal(N) for _ in range(3)]
        if kx == 0 and kz == 0:
            R = [r.real + 0j for r in R]
        lam = 91.6 + 4 * np.pi ** 2 * nu * ((kx / 5.5) ** 2 + (kz / 2.5) ** 2)
        mine = npo.TauSolver(kx, kz, 5.5, 2.5, a, b, lam, nu, N).solve(*R)
        ref = refcf.tausolve(kx, kz, 5.5, 2.5, a, b, lam, nu, N, *R)
        for m_, r_ in zip(mine, ref):
            assert np.abs(m_ - r_).max() < 1e-12, (kx, kz)


def test_np_oracle_transforms_nonlinear_and_step():
    """NumPy restatement of makePhysical/makeSpectral, the rotational term and one SBDF1 step vs the compiled reference."""
    from oracle import np_oracle as npo
    cfg = SMALLG
    ur = parity.ref_random(cfg, 11)
    c = ur.cdata.copy()
    fl = dict(cfg["flags"], timestepping="sbdf1")
    U, W = refcf.base_profiles(ur, refcf.make_flags(**fl))
    # transforms
    phys = npo.to_physical(c, cfg["Nz"])
    up = ur.like(); up.data[...] = ur.data; up.set_state(*ur.state()); up.make_physical()
    assert np.abs(phys - up.data[..., :cfg["Nz"]]).max() < 1e-13
    assert np.abs(npo.to_spectral(phys) - c).max() < 1e-14
    # rotational term
    fobj = refcf.nonlinear(ur, refcf.make_flags(**fl))  # keep the field alive: .cdata is a view of its memory
    fr = fobj.cdata.copy()
    fn = npo.rotational_nl(c, U, W, cfg["Lx"], cfg["Lz"], cfg["a"], cfg["b"])
    assert np.abs(fn - fr).max() < 1e-13 * max(1.0, np.abs(fr).max())
    # one SBDF1 step
    rd = refcf.RefDNS(ur, refcf.make_flags(**fl))
    rd.advance(1)
    u1, q1 = rd.get()
    u1c, q1c = u1.cdata.copy(), q1.cdata.copy()
    un, qn = npo.sbdf1_step(c, fl["dt"], fl["nu"], U, W, cfg["Lx"], cfg["Lz"], cfg["a"], cfg["b"])
    Kx, Kz = cfg["Nx"] // 3 - 1, cfg["Nz"] // 3 - 1
    keep = [m for m in range(cfg["Nx"]) if abs(npo.kx_of(m, cfg["Nx"])) <= Kx]
    assert np.abs(un[:, :, keep, :Kz + 1] - u1c[:, :, keep, :Kz + 1]).max() < 1e-12 * np.abs(u1c).max()
    assert np.abs(qn[:, keep, :Kz + 1] - q1c[0][:, keep, :Kz + 1]).max() < 1e-11 * max(1.0, np.abs(q1c).max())
