"""CPU tests (-m "not gpu"): the CUDA sources compiled for the fiber emulator (tests/emu) are run against the
oracle on small grids -- kernel logic (indexing, shared-memory staging, barriers, DMMA fragment layout) and the C++
host classes are exercised here; numbers measured on a GPU come only from the -m gpu tests."""
import numpy as np
import pytest

import channelflow_b200 as cf
from oracle import refcf
from tests import parity

pytestmark = pytest.mark.skipif(not refcf.available(), reason="oracle/_ref not built")

SMALL = dict(parity.C1, Nx=16, Ny=17, Nz=12)
ODD = dict(parity.C1, Nx=12, Ny=21, Nz=18, Lx=5.5, Lz=2.5, a=-1.0, b=1.0)


@pytest.fixture(scope="module")
def lib():
    return parity.emu_lib()


def test_cabi_symbols_product_library():
    """libcfgpu.so (the nvcc build) loads without a GPU and exports every symbol include/cfgpu.h declares."""
    import os
    if not os.path.exists(cf.LIB_GPU):
        pytest.skip("libcfgpu.so not built yet (python __graft_entry__.py)")
    assert cf.check_symbols()
    # ... and the list checked is exactly what the header declares
    import re
    hdr = open(os.path.join(parity.ROOT, "include", "cfgpu.h")).read()
    declared = set(re.findall(r"\b(cfgpu_[A-Za-z0-9_]+)\s*\(", hdr)) - {"cfgpu_exchange_fn", "cfgpu_allreduce_fn"}
    assert declared == set(cf.CFGPU_SYMBOLS), declared ^ set(cf.CFGPU_SYMBOLS)
    lib = cf.GpuLib()
    assert all(hasattr(lib.L, s_) for s_ in declared)


def test_product_library_has_no_cpu_fallback():
    import os
    if not os.path.exists(cf.LIB_GPU):
        pytest.skip("libcfgpu.so not built yet")
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cf.CfgpuError):
        cf.Context(cf.GpuLib())


@pytest.mark.parametrize("cfg", [SMALL, ODD])
def test_transforms(lib, cfg):
    r = parity.transforms(lib, cfg)
    assert max(r.values()) < 1e-14, r


def test_norms(lib):
    r = parity.norms(lib, SMALL)
    assert max(r.values()) < 1e-14, r


@pytest.mark.parametrize("over", [dict(), dict(Vsuck=0.0025, baseflow="suction"), dict(rotation=0.1), dict(dealiasing="none")])
def test_nonlinear(lib, over):
    r = parity.nonlinear(lib, SMALL, **over)
    assert r["nonlinear"] < 1e-14, r


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


@pytest.mark.parametrize("cfg", [SMALL, ODD])
def test_tausolve_modes(lib, cfg):
    """Every retained mode against the reference TauSolver; the right-hand sides are O(1), so 1e-14 absolute is a few
    ulps (the warp-parallel scans change the summation order of the reference's recurrences, nothing else)."""
    r = parity.tausolve_modes(lib, cfg)
    assert r["tau_abs_err"] <= 1e-14 * max(r["scale"], 1.0), r


@pytest.mark.parametrize("Ny", [65, 97, 129, 257, 385])
def test_tausolve_lane_block_sizes(lib, Ny):
    """Long profiles on a tiny (kx,kz) box: exercises every lane block size E of the warp-parallel column solver."""
    r = parity.tausolve_modes(lib, dict(parity.C1, Nx=6, Ny=Ny, Nz=6))
    assert r["tau_abs_err"] <= 1e-13 * max(r["scale"], 1.0), r


@pytest.mark.parametrize("stepper", ["sbdf3", "sbdf1", "sbdf2", "sbdf4", "cnfe1", "cnab2", "smrk2", "cnrk2"])
def test_dns_steppers(lib, stepper):
    r = parity.dns_steps(lib, SMALL, checkpoints=(1, 5), timestepping=stepper)
    assert r[1] < 1e-12 and r[5] < 1e-12 and r["cfl0"] < 1e-13, r


def test_dns_bulk_velocity_constraint(lib):
    r = parity.dns_steps(lib, SMALL, checkpoints=(1, 4), constraint="bulkv", Ubulk=0.0)
    assert r[1] < 1e-12 and r[4] < 1e-12 and r["dPdx"] < 1e-12, r


def test_dns_poiseuille_bulkv(lib):
    r = parity.dns_steps(lib, ODD, checkpoints=(1, 4), constraint="bulkv", Ubulk=2.0 / 3, ulowerwall=0.0, uupperwall=0.0,
                         nu=1 / 1800.0)
    assert r[1] < 1e-12 and r[4] < 1e-12 and r["dPdx"] < 1e-11, r


def test_dns_suction_golden_flags(lib):
    r = parity.dns_steps(lib, ODD, checkpoints=(1, 4), nu=1 / 400, Vsuck=1 / 400, dt=1 / 40, baseflow="suction")
    assert r[1] < 1e-12 and r[4] < 1e-12, r


def test_dns_c1_one_step(lib):
    """north_star gate on configs[0] (32x33x32): relative L2 <= 1e-12 after one step."""
    r = parity.dns_steps(lib, parity.C1, checkpoints=(1, 3))
    assert r[1] < 1e-12 and r[3] < 1e-12, r
    assert r["div"][0] < 1e-12 and r["div"][1] < 1e-12, r


def test_ff_file_roundtrip(lib, tmp_path):
    ur = parity.ref_random(SMALL, 7)
    ug = parity.to_gpu(lib, ur)
    ug.zero_padded_modes()  # the padded .ff format stores the retained modes only
    ug.save(str(tmp_path / "u"))
    h = lib.L.cf_field_load(str(tmp_path / "u").encode())
    v = cf.FlowField(lib, ug.Nx, ug.Ny, ug.Nz, 3, ug.Lx, ug.Lz, handle=h)
    assert np.abs(v.get() - ug.get()).max() == 0.0


def test_element_access_mirror(lib):
    ur = parity.ref_random(SMALL, 8)
    ug = parity.to_gpu(lib, ur)
    c = ug.cmplx(1, 3, 2, 0)
    assert c == ur.cdata[0, 3, 1, 2]
    ug.set_cmplx(1, 3, 2, 0, c + 1.0)
    ug.make_physical_y(); ug.make_spectral_y()
    assert abs(ug.cmplx(1, 3, 2, 0) - (c + 1.0)) < 1e-14


def test_timestep_logic(lib):
    """TimeStep::adjust / adjust_for_T (reference tests/gtest/timesteptest.cpp semantics)."""
    L = lib.L
    ts = L.cf_timestep_create(0.03125, 0.001, 0.2, 1.0, 0.4, 0.6, 1)
    assert L.cf_timestep_n(ts) == 32 and abs(L.cf_timestep_dt(ts) - 1 / 32) < 1e-16
    assert L.cf_timestep_adjust(ts, 0.5) == 0
    assert L.cf_timestep_adjust(ts, 1.0) == 1 and L.cf_timestep_n(ts) == 64
    assert L.cf_timestep_adjust(ts, 0.1) == 1 and L.cf_timestep_n(ts) == 13
    L.cf_timestep_adjust_for_T(ts, 10.5)
    assert L.cf_timestep_N(ts) == 11 and abs(L.cf_timestep_dT(ts) * 11 - 10.5) < 1e-13
    L.cf_timestep_free(ts)


def test_laminar_profiles(lib):
    for kw in (dict(ulowerwall=-1, uupperwall=1), dict(constraint="bulkv", Ubulk=2 / 3.0), dict(Vsuck=0.0025, baseflow="suction"),
               dict(Vsuck=1e-5, ulowerwall=0, uupperwall=1), dict(dPdx=-0.01)):
        a = cf.laminar_profile(lib, cf.make_flags(**kw), -1.0, 1.0, 33)
        b = refcf.laminar_profile(refcf.make_flags(**kw), -1.0, 1.0, 33)
        assert np.abs(a - b).max() < 1e-14, kw


def test_device_state_vectors(lib):
    """cfgpu_field2vector / vector2field kernels bit-identical (pack) / 1e-14 (unpack) to the reference's loops
    (flowfield.cpp:4481-4752); device dot / norm / axpy against NumPy."""
    r = parity.device_vectors(lib, SMALL)
    assert r["pack_max_abs"] == 0.0 and r["unpack_rel"] < 1e-14, r
    assert r["dot_rel"] < 1e-13 and r["norm_rel"] < 1e-14, r
    assert max(r["axpy_max_abs"], r["axpby_max_abs"], r["scale_max_abs"]) < 1e-15, r


def test_field2vector_roundtrip(lib):
    """field2vector / vector2field against the reference's (flowfield.cpp:4448-4752): same vector, same rebuilt field."""
    ur = parity.ref_random(SMALL, 9)
    ug = parity.to_gpu(lib, ur)
    xr = ur.to_vector()
    xg = ug.to_vector()
    assert xg.shape == xr.shape and np.abs(xg - xr).max() == 0.0
    rng = np.random.default_rng(3)
    y = xr + 1e-3 * rng.standard_normal(xr.shape)
    vr = ur.like().from_vector(y)
    vg = ug.like().from_vector(y)
    assert parity.rel_l2(vg.get(), vr.data) < 1e-14
    assert vg.padded()


def test_orr_sommerfeld_known_answer(lib):
    """tests/dnsOrrsommTest.cpp (golden eigenfunction + eigenvalue fixtures), shortened to T = 3 on the CPU emulator."""
    r = parity.orr_sommerfeld(lib, T1=3.0)
    assert r["err"] < 2e-6 and r["err"] < 1e-3 * r["norm"], r


def test_zero_and_parabola_known_answers(lib):
    """tests/dnsZeroTest.cpp / dnsParabolaTest.cpp: u = 0 stays 0 about the laminar base flow; the parabola 1 - y^2 carried
    as a fluctuation about a zero base flow with dP/dx = -2 nu is steady (1e-13)."""
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


def test_step_vs_numpy_restatement(lib):
    """One SBDF1 step of the CUDA path against the second oracle tier (oracle/np_oracle.py, pinned to the compiled
    reference in tests/test_oracle.py): independent of the compiled reference library."""
    from oracle import np_oracle as npo
    cfg = dict(parity.C1, Nx=12, Ny=17, Nz=12)
    ur = parity.ref_random(cfg, 11)
    c = ur.cdata.copy()
    fl = dict(cfg["flags"], timestepping="sbdf1")
    gd = cf.DNS(parity.to_gpu(lib, ur), cf.make_flags(**fl))
    U, W = cf.base_profiles(parity.to_gpu(lib, ur), cf.make_flags(**fl))
    gd.advance(1)
    u2, q2 = gd.get()
    un, qn = npo.sbdf1_step(c, fl["dt"], fl["nu"], U, W, cfg["Lx"], cfg["Lz"], cfg["a"], cfg["b"])
    mine = u2.get().view(np.complex128)
    assert np.abs(mine - un).max() < 1e-12 * np.abs(un).max()


def test_host_mirror_after_state_change(lib):
    """Element access after the field changed on the device: a de-aliased spectral field downloads its retained box only,
    so what the host mirror held for an earlier state must not show through in the aliased modes."""
    ur = parity.ref_random(SMALL, 12)
    ug = parity.to_gpu(lib, ur, padded=False)
    mx_alias = SMALL["Nx"] // 2
    ug.set_cmplx(mx_alias, 3, 0, 0, 1.0 + 2.0j)            # host mirror now holds an aliased-mode value
    assert ug.cmplx(mx_alias, 3, 0, 0) == 1.0 + 2.0j
    keep = ug.cmplx(1, 3, 1, 0)
    ug.zero_padded_modes()                                 # device-side; the field is now flagged de-aliased
    assert ug.padded()
    assert ug.cmplx(mx_alias, 3, 0, 0) == 0.0              # box download: the stale mirror entry must be gone
    assert ug.cmplx(1, 3, 1, 0) == keep
    c = ug.get().view(np.complex128)
    assert c[0, 3, mx_alias, 0] == 0.0                      # serial layout [i][ny][nx][mz]


@pytest.mark.parametrize("junk", [False, True])
def test_tile_layout_is_transparent(lib, junk):
    """Tile-major hot-path fields vs CFGPU_SERIAL_LAYOUT=1: identical bits; aliased-mode content of an un-padded initial
    field survives the steps untouched (NSE::solve writes retained modes only, nse.cpp:566-572)."""
    r = parity.layout_equivalence(lib, SMALL, nsteps=4, junk=junk)
    assert r["u_equal"] and r["q_equal"] and r["junk_kept"] and r["moved"] > 0, r
    r = parity.layout_equivalence(lib, parity.C1, nsteps=3, junk=junk, timestepping="cnab2")
    assert r["u_equal"] and r["q_equal"] and r["junk_kept"], r


def test_field_symmetry_ops(lib):
    assert parity.symmetry_ops(lib, SMALL, parity.SYMMETRIES[:6]) < 1e-14
    assert parity.symmetry_ops(lib, ODD, parity.SYMMETRIES[5:]) < 1e-14


def test_hookstep_search_mechanics(lib):
    """Newton-GMRES-hookstep on the device vectors: near the laminar state (G(0) = 0) one Newton step from a 5-vector Krylov
    space must reduce |G| exactly as its linear model predicts (finite-difference Jacobian, Arnoldi, SVD of the Hessenberg
    problem).  The full search on the stored equilibrium runs on the GPU (test_gpu.py)."""
    cfg = dict(parity.C1, Nx=8, Ny=13, Nz=8)
    ug = parity.to_gpu(lib, parity.ref_random(cfg, 3, magn=1e-3))
    r = cf.hookstep_search(ug, cf.make_flags(**cfg["flags"]), 0.25, 0.03125, Nnewton=1, Ngmres=5, epsSearch=1e-12, delta=0.1)
    assert r["newton_steps"] == 1 and r["fevals"] == 7 and r["gmres_iterations"] == 5, r
    assert r["history"][1] < 0.8 * r["history"][0], r


def test_variable_dt_loop(lib):
    """TimeStep::adjust -> DNS::reset_dt loop against the reference (small grid; the C2-sized run is in test_gpu.py)."""
    r = parity.variable_dt_loop(lib, dict(SMALL, magn=0.5), nintervals=3, dT=0.1, dt0=0.0125, nonlinearity="skew", constraint="bulkv",
                                Ubulk=2.0 / 3, ulowerwall=0.0, uupperwall=0.0, nu=1 / 1800.0)
    assert r["changes"] >= 1, r
    assert max(r["cfl_rel"]) < 1e-11 and r["u_rel"] < 1e-11 and r["dPdx"] < 1e-11, r


def test_netcdf4_field_reader(lib):
    r = parity.netcdf_reader(lib)
    assert r["padded"] and r["rel"] < 1e-14, r


def test_poincare_section_plane(lib):
    """DNSPoincare::advanceToSection (PlaneIntersection) against the restatement on the compiled reference, 16x17x16."""
    cfg = dict(parity.C1); cfg.update(Nx=16, Ny=17, Nz=16)
    r = parity.poincare_section(lib, cfg, kind="plane", nstride=4, maxstrides=30)
    assert r["found"] and r["ref_found"] and r["strides"] == r["ref_strides"] and r["sign"] == r["ref_sign"] == -1, r
    assert r["dt_cross"] < 1e-10 and r["u_rel"] < 1e-10 and abs(r["h"]) < 1e-13 and abs(r["h_by_ref"]) < 1e-12, r


def test_dns_symmetry_map_equivariance(lib):
    """DNS::operator*= maps the multistep history as well as the state (4 + 3 SBDF3 steps here, 6 + 6 on the GPU; sigma = rotation about z + half-box
    shift): the mapped run equals sigma of the unmapped one to round-off; mapping the state alone does not."""
    cfg = dict(parity.C1); cfg.update(Nx=16, Ny=17, Nz=16)
    r = parity.dns_equivariance(lib, cfg, n1=4, n2=3)
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
@pytest.mark.parametrize("Ny", [15, 17])
def test_nonlinear_forms_second_input(lib, nl, Ny):
    """forward y-transform with a second input (d/dy of u_i v added in coefficient space) through the fused pipeline: Ny = 15
    runs the DMMA contraction with its derivative matrices (2(Ny-1) = 28 has the factor 7), Ny = 17 (33 on the GPU) the two-pass
    FFT kernel"""
    cfg = dict(parity.C1); cfg.update(Nx=8, Ny=Ny, Nz=16)   # (the fused pipeline needs a power-of-two Nz >= 16)
    r = parity.nonlinear(lib, cfg, nonlinearity=nl)
    assert r["nonlinear"] < 1e-12, r
