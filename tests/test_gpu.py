"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI / host classes, against the oracle
(the unmodified reference compiled in oracle/_ref, shipped prebuilt to the GPU box) and the committed golden
fixtures.  Tolerances are the ones BASELINE.json's north_star states: relative L2 <= 1e-12 after one step,
<= 1e-9 after 100 steps, golden pair L2Dist <= 1e-13 (timeIntegrationTest.cpp:42)."""
import numpy as np
import pytest

import channelflow_b200 as cf
from oracle import refcf
from tests import parity

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refcf.available(), reason="oracle/_ref missing")]

SMALL = dict(parity.C1, Nx=16, Ny=17, Nz=12)
ODD = dict(parity.C1, Nx=12, Ny=21, Nz=18, Lx=5.5, Lz=2.5)
MID = dict(parity.C1, Nx=48, Ny=49, Nz=48, Lx=2 * np.pi / 1.14, Lz=2 * np.pi / 2.5)
BIG = dict(parity.C1, Nx=96, Ny=97, Nz=64)


@pytest.fixture(scope="module")
def lib():
    return parity.gpu_lib()


def test_native_library_is_cuda(lib):
    assert b"sm_100a" in lib.gpu.L.cfgpu_version()
    assert not lib.gpu.missing_symbols()


@pytest.mark.parametrize("cfg", [SMALL, ODD, parity.C1, MID, BIG])
def test_transforms(lib, cfg):
    r = parity.transforms(lib, cfg)
    assert max(r.values()) < 1e-13, r


@pytest.mark.parametrize("cfg", [SMALL, MID])
def test_norms(lib, cfg):
    r = parity.norms(lib, cfg)
    assert max(r.values()) < 1e-13, r


@pytest.mark.parametrize("over", [dict(), dict(Vsuck=0.0025, baseflow="suction"), dict(rotation=0.1), dict(dealiasing="none")])
@pytest.mark.parametrize("cfg", [SMALL, MID])
def test_nonlinear(lib, cfg, over):
    r = parity.nonlinear(lib, cfg, **over)
    assert r["nonlinear"] < 1e-13, r


@pytest.mark.parametrize("nl", ["conv", "div", "skew", "alt", "linear"])
@pytest.mark.parametrize("over", [dict(), dict(rotation=0.1, Vsuck=0.0025, baseflow="suction"), dict(dealiasing="none")])
def test_nonlinear_methods(lib, nl, over):
    """convection / divergence / skew-symmetric / alternating / linearized-about-profile terms (nse.cpp:12-91)."""
    if nl == "linear" and over.get("rotation"):
        pytest.skip("reference leaves f in the physical state for LinearAboutProfile with rotation != 0 (nse.cpp:17-25): undefined")
    r = parity.nonlinear(lib, SMALL, nonlinearity=nl, **over)
    assert r["nonlinear"] < 1e-12, r


POW2 = dict(parity.C1, Nx=16, Ny=17, Nz=16) if "emu" in __file__ else dict(parity.C1, Nx=64, Ny=49, Nz=64)


@pytest.mark.parametrize("nl", ["conv", "div", "skew", "alt"])
@pytest.mark.parametrize("over", [dict(), dict(rotation=0.1, Vsuck=0.0025, baseflow="suction"), dict(dealiasing="none")])
def test_nonlinear_methods_fused(lib, nl, over):
    """Power-of-two Nz: the fused compact-pencil pipeline for the convection / divergence / skew-symmetric forms."""
    r = parity.nonlinear(lib, POW2, nonlinearity=nl, **over)
    assert r["nonlinear"] < 1e-12, r


def test_dns_skew_bulkv_fused(lib):
    """BASELINE configs[1] in small: plane Poiseuille, fixed flux, skew-symmetric form."""
    r = parity.dns_steps(lib, POW2, checkpoints=(1, 4), nonlinearity="skew", constraint="bulkv", Ubulk=2.0 / 3, ulowerwall=0.0,
                         uupperwall=0.0, nu=1 / 1800.0)
    assert r[1] < 1e-12 and r[4] < 1e-12 and r["dPdx"] < 1e-11, r


@pytest.mark.parametrize("nl", ["skew", "conv", "div", "alt", "linear"])
def test_dns_nonlinear_methods(lib, nl):
    r = parity.dns_steps(lib, SMALL, checkpoints=(1, 4), nonlinearity=nl)
    assert r[1] < 1e-12 and r[4] < 1e-12, r


@pytest.mark.parametrize("cfg", [SMALL, parity.C1])
def test_tausolve_modes(lib, cfg):
    r = parity.tausolve_modes(lib, cfg)
    assert r["tau_abs_err"] <= 1e-13 * max(r["scale"], 1.0), r


@pytest.mark.parametrize("Ny", [65, 97, 129, 257, 385, 513])
def test_tausolve_lane_block_sizes(lib, Ny):
    r = parity.tausolve_modes(lib, dict(parity.C1, Nx=12, Ny=Ny, Nz=12))
    assert r["tau_abs_err"] <= 1e-13 * max(r["scale"], 1.0), r


@pytest.mark.parametrize("stepper", ["sbdf3", "sbdf1", "sbdf2", "sbdf4", "cnfe1", "cnab2", "smrk2", "cnrk2"])
def test_dns_steppers(lib, stepper):
    r = parity.dns_steps(lib, SMALL, checkpoints=(1, 5), timestepping=stepper)
    assert r[1] < 1e-12 and r[5] < 1e-12 and r["cfl0"] < 1e-12, r


def test_dns_bulk_velocity(lib):
    r = parity.dns_steps(lib, ODD, checkpoints=(1, 4), constraint="bulkv", Ubulk=2.0 / 3, ulowerwall=0.0, uupperwall=0.0,
                         nu=1 / 1800.0)
    assert r[1] < 1e-12 and r[4] < 1e-12 and r["dPdx"] < 1e-11, r


def test_c1_one_step_and_100_steps(lib):
    """BASELINE.json configs[0]: plane Couette Re=400, 32x33x32, SBDF3, rotational, dealiased, dt=0.02."""
    r = parity.dns_steps(lib, parity.C1, checkpoints=(1, 100))
    assert r[1] <= 1e-12, r
    assert r[100] <= 1e-9, r
    assert r["div"][0] < 1e-12 and r["div"][1] < 1e-12, r


def test_mid_grid_steps(lib):
    r = parity.dns_steps(lib, MID, checkpoints=(1, 10), nu=1 / 400, Vsuck=1 / 400, dt=1 / 40, baseflow="suction")
    assert r[1] <= 1e-12 and r[10] <= 1e-11, r


def test_golden_pair(lib):
    """tests/data/uinit.nc -> 440 SBDF3 steps -> ufinal.nc, L2Dist <= 1e-13 (the reference's own regression test)."""
    r, _, _ = parity.golden_pair(lib)
    assert r["l2dist_to_ufinal"] <= 1e-13, r


def test_full_size_band_limited_consistency(lib):
    """Size-independent property at the bench grid (512x257x512): a field that only excites the modes retained by a
    coarse 48x257x48 grid must, after one step on the fine grid, agree on those modes with the reference's step on
    the coarse grid (the dealiased quadratic term is exact on both)."""
    coarse = dict(parity.C1, Nx=48, Ny=257, Nz=48)
    fine = dict(coarse, Nx=512, Nz=512)
    ur = parity.ref_random(coarse, seed=11)
    fl = dict(coarse["flags"], timestepping="sbdf1")
    rd = refcf.RefDNS(ur, refcf.make_flags(**fl))
    rd.advance(1)
    u1, _ = rd.get()
    # embed the coarse spectrum into the fine grid
    Mzc, Mzf = coarse["Nz"] // 2 + 1, fine["Nz"] // 2 + 1
    big = np.zeros((3, 257, fine["Nx"], Mzf), dtype=np.complex128)
    Kx, Kz = coarse["Nx"] // 3 - 1, coarse["Nz"] // 3 - 1
    src = ur.cdata
    for kx in range(-Kx, Kx + 1):
        big[:, :, kx % fine["Nx"], :Kz + 1] = src[:, :, kx % coarse["Nx"], :Kz + 1]
    ug = cf.FlowField(lib, fine["Nx"], 257, fine["Nz"], 3, fine["Lx"], fine["Lz"]).set(big.view(np.float64), padded=True)
    gd = cf.DNS(ug, cf.make_flags(**fl))
    gd.advance(1)
    u2, _ = gd.get()
    out = u2.get().view(np.complex128)
    ref = u1.cdata
    num = den = 0.0
    for kx in range(-Kx, Kx + 1):
        d = out[:, :, kx % fine["Nx"], :Kz + 1] - ref[:, :, kx % coarse["Nx"], :Kz + 1]
        num += float(np.sum(np.abs(d) ** 2))
        den += float(np.sum(np.abs(ref[:, :, kx % coarse["Nx"], :Kz + 1]) ** 2))
    assert np.sqrt(num / den) < 1e-11, np.sqrt(num / den)


def test_full_size_roundtrip(lib):
    """makePhysical -> makeSpectral is the identity at 512x257x512 (idempotence)."""
    rng = np.random.default_rng(0)
    u = cf.FlowField(lib, 512, 257, 512, 1, 4 * np.pi, 2 * np.pi)
    a = np.zeros(u.shape)
    c = a.view(np.complex128)
    c[:, :40, :20, :20] = rng.standard_normal((1, 40, 20, 20)) + 1j * rng.standard_normal((1, 40, 20, 20))
    c[:, :, 0, 0] = c[:, :, 0, 0].real
    c[:, :, 1:20, 0] = 0  # keep the kz=0 plane Hermitian without building conjugates
    u.set(a)
    u.make_physical()
    u.make_spectral()
    b = u.get()
    assert parity.rel_l2(b, a) < 1e-13


def test_slab_decomposition_nccl():
    """2-GPU run of the slab decomposition over NCCL against the single-process oracle (skipped with < 2 GPUs)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(parity.ROOT, "tests", "mp_slab_worker.py"), "sbdf3"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, cwd=parity.ROOT,
                       env=dict(os.environ, CF_WORKER_BACKEND="nccl"))
    assert r.returncode == 0, r.stdout.decode()[-4000:]


def test_device_state_vectors(lib):
    """cfgpu_field2vector / vector2field kernels bit-identical (pack) / 1e-14 (unpack) to the reference's loops
    (flowfield.cpp:4481-4752); device dot / norm / axpy against NumPy."""
    r = parity.device_vectors(lib, MID)
    assert r["pack_max_abs"] == 0.0 and r["unpack_rel"] < 1e-14, r
    assert r["dot_rel"] < 1e-13 and r["norm_rel"] < 1e-14, r
    assert max(r["axpy_max_abs"], r["axpby_max_abs"], r["scale_max_abs"]) < 1e-15, r


def test_field2vector_roundtrip(lib):
    ur = parity.ref_random(MID, 9)
    ug = parity.to_gpu(lib, ur)
    xr, xg = ur.to_vector(), ug.to_vector()
    assert xg.shape == xr.shape and np.abs(xg - xr).max() == 0.0
    y = xr + 1e-3 * np.random.default_rng(3).standard_normal(xr.shape)
    vr, vg = ur.like().from_vector(y), ug.like().from_vector(y)
    assert parity.rel_l2(vg.get(), vr.data) < 1e-14


@pytest.mark.parametrize("wname", ["c5", "mixed_radix"])
def test_full_size_properties(lib, wname):
    """BASELINE-size grid (C5, 256x129x256) and a large mixed-radix grid (384x161x192 = 3*2^7 x 161 x 3*2^6, the non
    power-of-two FFT kernels and a 6-per-lane tau solver): size-independent properties instead of the (slow) CPU oracle --
    transform round trip is the identity; SBDF steps of a solenoidal no-slip field stay solenoidal and no-slip
    (vector2field(field2vector(u)) == u, reference flowfield.cpp:4754-4758); L2Norm is invariant under the round trip."""
    import bench
    w = bench.WORKLOADS["c5"] if wname == "c5" else dict(bench.WORKLOADS["c5"], Nx=384, Ny=161, Nz=192)
    u0 = bench.synthetic_field(w)
    ug = cf.FlowField(lib, w["Nx"], w["Ny"], w["Nz"], 3, w["Lx"], w["Lz"]).set(u0, padded=True)
    n0 = ug.l2norm()
    v = ug.copy()
    v.make_physical(); v.make_spectral()
    assert ug.l2dist(v) <= 1e-13 * n0
    dns = cf.DNS(ug, cf.make_flags(**bench.flags_kw(w)))
    dns.advance(3)
    u1, _ = dns.get()
    x = u1.to_vector()
    u2 = u1.like().from_vector(x)
    assert u1.l2dist(u2) <= 1e-12 * u1.l2norm(), (u1.l2dist(u2), u1.l2norm())
    assert abs(u1.l2norm() - n0) < 0.05 * n0


@pytest.mark.parametrize("stepper,tol", [("sbdf3", 2e-6), ("cnrk2", 5e-6), ("sbdf2", 2e-4), ("cnab2", 4e-5)])
def test_orr_sommerfeld_known_answer(lib, stepper, tol):
    """tests/dnsOrrsommTest.cpp with the reference's own per-scheme tolerances (velocity part), full T = 13."""
    r = parity.orr_sommerfeld(lib, timestepping=stepper)
    assert r["err"] < tol, r


def test_zero_and_parabola_known_answers(lib):
    cfg = dict(parity.C1, Nx=8, Ny=17, Nz=8)
    z = cf.FlowField(lib, cfg["Nx"], cfg["Ny"], cfg["Nz"], 3, cfg["Lx"], cfg["Lz"])
    d = cf.DNS(z, cf.make_flags(**cfg["flags"]))
    d.advance(5)
    assert d.get()[0].l2norm() == 0.0
    par = np.zeros(z.shape)
    par[0, 0, 0, 0], par[0, 2, 0, 0] = 0.5, -0.5
    p = z.like().set(par)
    nu = 1.0 / 400
    d = cf.DNS(p, cf.make_flags(nu=nu, dt=0.02, baseflow="zero", constraint="gradp", dPdx=-2 * nu, ulowerwall=0.0, uupperwall=0.0,
                                 dealiasing="none"))
    d.advance(10)
    assert d.get()[0].l2dist(p) < 1e-13


@pytest.mark.parametrize("junk", [False, True])
def test_tile_layout_is_transparent(lib, junk):
    """Tile-major hot-path fields vs CFGPU_SERIAL_LAYOUT=1: identical bits; aliased-mode content of an un-padded initial
    field survives the steps untouched (NSE::solve writes retained modes only, nse.cpp:566-572)."""
    r = parity.layout_equivalence(lib, MID, nsteps=4, junk=junk)
    assert r["u_equal"] and r["q_equal"] and r["junk_kept"] and r["moved"] > 0, r
    r = parity.layout_equivalence(lib, parity.C1, nsteps=3, junk=junk, timestepping="cnab2")
    assert r["u_equal"] and r["q_equal"] and r["junk_kept"], r


@pytest.mark.parametrize("cfgname", ["C1", "GOLD"])
def test_cuda_graph_replay_is_bit_identical(lib, cfgname):
    """Launch-bound grids replay `order` SBDF steps as one CUDA graph (host/dnsalgo.cpp): identical bits to eager launches."""
    cfg = parity.C1 if cfgname == "C1" else dict(parity.C1, Nx=48, Ny=35, Nz=48, Lx=2 * np.pi / 1.14, Lz=2 * np.pi / 2.5)
    r = parity.graph_equivalence(lib, cfg)
    assert r["u_identical"] and r["q_identical"] and r["cfl_identical"], r
    print("graph replay:", {k: v for k, v in r.items() if "sec" in k or "launch" in k})


def test_field_symmetry_ops(lib):
    for cfg in (dict(parity.C1, Nx=16, Ny=17, Nz=12), dict(parity.C1, Nx=12, Ny=21, Nz=18, Lx=5.5, Lz=2.5), parity.C1):
        assert parity.symmetry_ops(lib, cfg) < 1e-14


def test_findsoln_reconverges_to_the_stored_solution(lib):
    """North-star gate: the Newton-Krylov-hookstep search (device vectors, one DNS integration per Krylov vector) converges
    from the perturbed guess of tests/findsolnTest.cpp to residual <= 1e-10, and lands on the stored solution (the
    reference test's own tolerance for that distance is 1e-5, findsolnTest.cpp:129)."""
    r = parity.findsoln_eq(lib)
    print("findsoln:", {k: v for k, v in r.items()})
    assert r["residual"] <= 1e-10, r
    assert r["l2dist_to_stored"] < 1e-5, r
    assert max(r["div_bc"]) < 1e-10, r


def test_c2_variable_dt_loop(lib):
    """BASELINE configs[1]: plane Poiseuille at fixed flux, skew-symmetric form, variable dt -- the TimeStep::adjust ->
    DNS::reset_dt loop (every change rebuilds the tau operators and restarts SBDF3 with its SMRK2 initial steps) at
    128x97x128 against the compiled reference."""
    cfg = dict(parity.C1, Nx=128, Ny=97, Nz=128, Lx=2 * np.pi, Lz=np.pi, magn=0.3)
    r = parity.variable_dt_loop(lib, cfg, nintervals=4, dT=0.1, dt0=0.004, nonlinearity="skew", constraint="bulkv", Ubulk=2.0 / 3,
                                ulowerwall=0.0, uupperwall=0.0, nu=1 / 1800.0)
    print("variable dt:", r)
    assert r["changes"] >= 1 and r["steps"] >= 20, r
    assert max(r["cfl_rel"]) < 1e-10 and r["u_rel"] < 1e-10 and r["dPdx"] < 1e-10, r


def test_netcdf4_field_reader(lib):
    r = parity.netcdf_reader(lib)
    assert r["padded"] and r["rel"] < 1e-14, r


def test_findsoln_program(lib, tmp_path):
    """channelflow_b200/bin/findsoln_b200 (the reference's findsoln options, fixed-T subset): reads the reference's own NetCDF
    file, re-converges the travelling wave with the x phase shift as an unknown, writes ubest / sigmabest."""
    import os
    import shutil
    import subprocess
    exe = os.path.join(parity.ROOT, "channelflow_b200", "bin", "findsoln_b200")
    if not os.path.exists(exe):
        pytest.skip("findsoln_b200 not built")
    shutil.copyfile(os.path.join(parity.GOLDEN, "eq.nc"), str(tmp_path / "eq.nc"))
    open(str(tmp_path / "sigma.asc"), "w").write("1 1 1 1 0.28168880386692519 0\n")
    r = subprocess.run([exe, "-eqb", "-xrel", "-T", "10", "-dt", "0.03125", "-vdt", "false", "-R", "400", "-sigma", "sigma.asc", "-Nn", "5",
                        "-es", "1e-12", "eq"], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    last = [l for l in r.stdout.splitlines() if "L2Norm(G) ==" in l][-1]
    print(last)
    assert "converged" in last, r.stdout[-1500:]
    assert os.path.exists(str(tmp_path / "ubest.ff")) and os.path.exists(str(tmp_path / "sigmabest.asc"))
    hist = [float(x) for x in open(str(tmp_path / "convergence.asc")).read().split("\n")[1:] if x.strip()]
    assert hist[-1] < 1e-12 and hist[-1] < 1e-4 * hist[0], hist


def test_poincare_section_plane(lib):
    """DNSPoincare::advanceToSection with a PlaneIntersection condition against the restatement on the compiled reference
    (tests/parity.py:poincare_section).  Tolerances: crossing time 1e-10, crossing field 1e-10 relative, |h| <= 1e-13."""
    r = parity.poincare_section(lib, parity.C1, kind="plane", nstride=5, maxstrides=40)
    print("poincare plane:", r)
    assert r["found"] and r["ref_found"] and r["strides"] == r["ref_strides"] and r["sign"] == r["ref_sign"] == -1, r
    assert r["dt_cross"] < 1e-10 and r["u_rel"] < 1e-10 and abs(r["h"]) < 1e-13 and abs(r["h_by_ref"]) < 1e-12, r


def test_poincare_section_drag_dissipation(lib):
    """The I - D = 0 section (DragDissipation: wallshear - dissipation on the device) crossed downwards at t ~ 5 on a 16x17x16 box."""
    cfg = dict(parity.C1); cfg.update(Nx=16, Ny=17, Nz=16)
    r = parity.poincare_section(lib, cfg, kind="drag", nstride=5, maxstrides=80, crosssign=-1)
    print("poincare I-D:", r)
    assert r["found"] and r["ref_found"] and r["strides"] == r["ref_strides"] > 40 and r["sign"] == -1, r
    assert r["dt_cross"] < 1e-8 and r["u_rel"] < 1e-9 and abs(r["h"]) < 1e-12, r


def test_dns_symmetry_map_equivariance(lib):
    """DNS::operator*= maps the multistep history as well as the state (6 + 6 SBDF3 steps, sigma = rotation about z + half-box
    shift): the mapped run equals sigma of the unmapped one to round-off; mapping the state alone does not."""
    r = parity.dns_equivariance(lib, parity.C1)
    assert r["mapped"] < 1e-12 and r["state_only"] > 1e3 * max(r["mapped"], 1e-14), r


def test_netcdf_field_writer(lib, tmp_path):
    """FlowField::save("x.nc"): reference schema in the classic container, checked with scipy's independent reader against
    the values stock Channelflow wrote (1e-14 absolute), re-read by this package (1e-14), full-grid and CDF-1 variants."""
    r = parity.netcdf_writer(lib, tmp_path)
    assert r["dims"] == {"X": 16, "Y": 33, "Z": 16} and r["var_dims"] == ("Z", "Y", "X") and r["title"] == "FlowField", r
    assert r["vars"] == ["X", "Y", "Z", "Velocity_X", "Velocity_Y", "Velocity_Z"], r
    assert (int(r["attrs"]["Nx"]), int(r["attrs"]["Ny"]), int(r["attrs"]["Nz"])) == (24, 33, 24) and float(r["attrs"]["a"]) == -1.0, r
    assert r["grid_err"] < 1e-14 and r["values_abs"] < 1e-14 and r["reread_rel"] < 1e-14 and r["reread_padded"], r
    assert r["full_rel"] < 1e-14 and not r["full_padded"] and r["full_dims"] == {"X": 32, "Y": 33, "Z": 32}, r
    assert r["scipy_rel"] < 1e-14, r


# profile lengths: 2(Ny-1) smooth in 2, 3, 5 -> half-length FFT kernel (csrc/yfft.cu), radix mixes 2/3/4/5/8 and the short
# edge cases; Ny = 15, 23 -> the DMMA contraction (2(Ny-1) has the factor 7 / 11)
Y_LENGTHS = [5, 7, 9, 11, 13, 15, 21, 23, 25, 31, 41, 49, 51, 61, 65, 97, 101]


@pytest.mark.parametrize("Ny", Y_LENGTHS)
def test_y_transform_lengths(lib, Ny):
    cfg = dict(parity.C1); cfg.update(Nx=8, Ny=Ny, Nz=8)
    r = parity.transforms(lib, cfg)
    assert max(r.values()) < 2e-14, r


@pytest.mark.parametrize("Ny", [9, 21, 31, 41, 61])
def test_nonlinear_y_lengths(lib, Ny):
    """the rotational term needs u_y, w_y: the derivative unit of the y-transform (suffix sums + transform) at other lengths"""
    cfg = dict(parity.C1); cfg.update(Nx=12, Ny=Ny, Nz=12)
    r = parity.nonlinear(lib, cfg)
    assert r["nonlinear"] < 2e-14, r


@pytest.mark.parametrize("nl", ["div", "skew"])
@pytest.mark.parametrize("Ny", [15, 33])
def test_nonlinear_forms_second_input(lib, nl, Ny):
    """forward y-transform with a second input (d/dy of u_i v added in coefficient space) through the fused pipeline: Ny = 15
    runs the DMMA contraction with its derivative matrices (2(Ny-1) = 28 has the factor 7), Ny = 33 the two-pass FFT kernel"""
    cfg = dict(parity.C1); cfg.update(Nx=16, Ny=Ny, Nz=16)
    r = parity.nonlinear(lib, cfg, nonlinearity=nl)
    assert r["nonlinear"] < 1e-12, r
