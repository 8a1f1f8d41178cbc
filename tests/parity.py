"""Parity scenarios shared by the CPU-emulation tests (-m "not gpu") and the GPU tests (-m gpu).

Every scenario runs the same inputs through (a) the oracle = the unmodified Channelflow reference compiled in
oracle/_ref (oracle/refcf.py) and (b) this package's libraries, and returns error measures.  `lib` is a
channelflow_b200.HostLib: the real CUDA build on the GPU box, or the emulation build (tests/_emu) on the CPU.
"""
import os

import numpy as np

import channelflow_b200 as cf
from oracle import refcf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "_emu")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# BASELINE.json configs[0]: plane Couette Re=400, 32x33x32, Lx=2pi, Lz=pi, SBDF3, rotational, dealiased, dt=0.02
C1 = dict(Nx=32, Ny=33, Nz=32, Lx=2 * np.pi, Lz=np.pi, a=-1.0, b=1.0,
          flags=dict(nu=1.0 / 400, dt=0.02, ulowerwall=-1.0, uupperwall=1.0, baseflow="laminar", constraint="gradp",
                     timestepping="sbdf3", initstepping="smrk2", nonlinearity="rot", dealiasing="xz"))


def emu_lib():
    import subprocess
    subprocess.check_call([os.path.join(ROOT, "tests", "emu", "build_emu.sh")], stdout=subprocess.DEVNULL)
    host = os.path.join(EMU_DIR, "libchflow_b200_emu.so")
    srcs = [os.path.join(ROOT, "channelflow_b200", "host", f) for f in os.listdir(os.path.join(ROOT, "channelflow_b200", "host"))
            if f.endswith(".cpp")]
    hdrs = []
    for d, _, fs in os.walk(os.path.join(ROOT, "channelflow_b200", "host")):
        hdrs += [os.path.join(d, f) for f in fs if f.endswith(".h")]
    newest = max(os.path.getmtime(p) for p in srcs + hdrs + [os.path.join(EMU_DIR, "libcfgpu_emu.so")])
    if not os.path.exists(host) or os.path.getmtime(host) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I",
                               os.path.join(ROOT, "channelflow_b200", "host"), "-o", host] + srcs +
                              ["-L", EMU_DIR, "-lcfgpu_emu", "-Wl,-rpath,$ORIGIN", "-Wl,-Bsymbolic"])
    return cf.HostLib(host, os.path.join(EMU_DIR, "libcfgpu_emu.so"))


def gpu_lib():
    return cf.HostLib()


def ref_random(cfg, seed=1, magn=0.2, smooth=0.4):
    """tools/randomfield.cpp rule (serial drand48) through the compiled reference."""
    return refcf.RefField(cfg["Nx"], cfg["Ny"], cfg["Nz"], 3, cfg["Lx"], cfg["Lz"], cfg["a"], cfg["b"]).randomfield(seed, magn, smooth)


def to_gpu(lib, rf, padded=True):
    st = rf.state()
    return cf.FlowField(lib, rf.Nx, rf.Ny, rf.Nz, rf.Nd, rf.Lx, rf.Lz, rf.a, rf.b).set(rf.data, st[0], st[1], padded=padded)


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def physical_mask(shape, Nz):
    m = np.ones(shape, bool)
    m[..., Nz:] = False  # the two padding reals per z-line are unspecified after c2r
    return m


# ------------------------------------------------------------------------------------------------ scenarios
def transforms(lib, cfg, seed=2):
    ur = ref_random(cfg, seed)
    u0 = ur.data.copy()
    ug = to_gpu(lib, ur)
    out = {}
    ur.make_physical_y(); ug.make_physical_y()
    out["physical_y"] = rel_l2(ug.get(), ur.data)
    ur.make_physical_xz(); ug.make_physical_xz()
    m = physical_mask(ur.data.shape, cfg["Nz"])
    out["physical"] = rel_l2(ug.get()[m], ur.data[m])
    ur.make_spectral_xz(); ug.make_spectral_xz()
    out["spectral_xz"] = rel_l2(ug.get(), ur.data)
    ur.make_spectral_y(); ug.make_spectral_y()
    out["roundtrip_vs_ref"] = rel_l2(ug.get(), ur.data)
    out["roundtrip_vs_input"] = rel_l2(ug.get(), u0)
    return out


def norms(lib, cfg):
    ur, vr = ref_random(cfg, 3), ref_random(cfg, 4, magn=0.1)
    ug, vg = to_gpu(lib, ur), to_gpu(lib, vr)
    out = {"l2norm": abs(ug.l2norm() - ur.l2norm()) / ur.l2norm(),
           "l2dist": abs(ug.l2dist(vg) - ur.l2dist(vr)) / ur.l2dist(vr),
           "l2ip": abs(ug.l2ip(vg) - ur.l2ip(vr)) / abs(ur.l2norm() * vr.l2norm())}
    out["l2norm3d"] = abs(ug.l2norm3d() - ur.l2norm3d()) / ur.l2norm3d()
    ur.set_padded(False); vr.set_padded(False); ug.set_padded(False); vg.set_padded(False)
    out["l2norm3d_unpadded"] = abs(ug.l2norm3d() - ur.l2norm3d()) / ur.l2norm3d()
    out["l2norm_unpadded"] = abs(ug.l2norm() - ur.l2norm()) / ur.l2norm()
    return out


def nonlinear(lib, cfg, seed=1, **flag_over):
    fl = dict(cfg["flags"]); fl.update(flag_over)
    ur = ref_random(cfg, seed)
    fr = refcf.nonlinear(ur, refcf.make_flags(**fl))
    fg = cf.nonlinear(to_gpu(lib, ur), cf.make_flags(**fl))
    a, b = fg.get().copy(), fr.data.copy()
    if fl.get("dealiasing", "xz") in ("none", "y"):
        # without xz-dealiasing the reference also fills the Nyquist modes kx = Nx/2, kz = Nz/2 of f; NSE::solve never
        # reads them (nse.cpp:498 skips kxmax/kzmax), and this implementation does not produce them
        for arr in (a, b):
            c = arr.view(np.complex128)
            c[:, :, cfg["Nx"] // 2, :] = 0
            c[:, :, :, cfg["Nz"] // 2] = 0
    return {"nonlinear": rel_l2(a, b), "scale": float(np.abs(fr.data).max())}


def dns_steps(lib, cfg, checkpoints=(1, 10), seed=1, u0=None, **flag_over):
    """Advance the same initial field with the reference DNS and with ours; relative L2 error of u at checkpoints."""
    fl = dict(cfg["flags"]); fl.update(flag_over)
    ur = u0 if u0 is not None else ref_random(cfg, seed)
    rd = refcf.RefDNS(ur, refcf.make_flags(**fl))
    gd = cf.DNS(to_gpu(lib, ur), cf.make_flags(**fl))
    out, done = {"cfl0": abs(gd.cfl() - rd.cfl()) / max(abs(rd.cfl()), 1e-300)}, 0
    for n in checkpoints:
        rd.advance(n - done); gd.advance(n - done); done = n
        u1, q1 = rd.get(); u2, q2 = gd.get()
        out[n] = rel_l2(u2.get(), u1.data)
        out["q%d" % n] = rel_l2(q2.get(), q1.data)
    out["dPdx"] = abs(gd.dPdx() - rd.dPdx())
    out["cfl_end"] = abs(gd.cfl() - rd.cfl()) / max(abs(rd.cfl()), 1e-300)
    out["div"] = u2_div(u2)
    return out


def u2_div(ug):
    """divergence + wall BC norm of a device field, evaluated by the reference's own divNorm/bcNorm."""
    r = refcf.RefField(ug.Nx, ug.Ny, ug.Nz, ug.Nd, ug.Lx, ug.Lz, ug.a, ug.b)
    r.data[...] = ug.get()
    r.set_padded(True)
    return (r.divnorm(), r.bcnorm())


def tausolve_modes(lib, cfg, lam_t=11.0 / 6 / 0.02, nu=1.0 / 400, seed=5):
    """C-ABI cfgpu_nse_solve against the reference TauSolver, mode by mode."""
    ctx = cf.Context(lib.gpu)
    Nx, Ny, Nz, Lx, Lz, a, b = (cfg[k] for k in ("Nx", "Ny", "Nz", "Lx", "Lz", "a", "b"))
    nse = cf.Nse(ctx, Nx, Ny, Nz, Lx, Lz, a, b, Ubase=np.zeros(Ny), Wbase=np.zeros(Ny), nu=nu)
    nse.reset_lambda([lam_t])
    R = ref_random(cfg, seed, magn=1.0)
    Rg = ctx.field(Nx, Ny, Nz, 3, Lx, Lz, a, b).upload(R.data, padded=True)
    uo, qo = Rg.like(), Rg.like(Nd=1)
    nse.solve(0, [1.0], [Rg], uo, qo)
    du, dq, Rc = uo.download().view(np.complex128), qo.download().view(np.complex128), R.cdata
    Kx, Kz = Nx // 3 - 1, Nz // 3 - 1
    err, scale = 0.0, 0.0
    for kx in range(-Kx, Kx + 1):
        for kz in range(0, Kz + 1):
            mx = kx % Nx
            lam = lam_t + 4 * np.pi ** 2 * nu * ((kx / Lx) ** 2 + (kz / Lz) ** 2)
            sol = refcf.tausolve(kx, kz, Lx, Lz, a, b, lam, nu, Ny, *(Rc[i, :, mx, kz].copy() for i in range(3)))
            if kx == 0 and kz == 0:
                sol = tuple(x.real + 0j for x in sol)
            mine = (du[0, :, mx, kz], du[1, :, mx, kz], du[2, :, mx, kz], dq[0, :, mx, kz])
            for s_, m_ in zip(sol, mine):
                err = max(err, float(np.abs(s_ - m_).max()))
                scale = max(scale, float(np.abs(s_).max()))
    return {"tau_abs_err": err, "scale": scale}


def golden_pair(lib, nsteps=440):
    """tests/timeIntegrationTest.cpp: uinit -> 440 SBDF3 steps (11 x advance(40)) -> compare with ufinal (L2Dist)."""
    g = np.load(os.path.join(GOLDEN, "golden_pair.npz"))
    geo = dict(Nx=int(g["Nx"]), Ny=int(g["Ny"]), Nz=int(g["Nz"]), Lx=float(g["Lx"]), Lz=float(g["Lz"]), a=float(g["a"]), b=float(g["b"]))
    ur = refcf.RefField(geo["Nx"], geo["Ny"], geo["Nz"], 3, geo["Lx"], geo["Lz"], geo["a"], geo["b"]).load_padded_physical(g["uinit"])
    vr = ur.like().load_padded_physical(g["ufinal"])
    fl = dict(nu=1.0 / 400, Vsuck=1.0 / 400, dt=1.0 / 40, baseflow="suction", constraint="gradp", dPdx=0.0)
    gd = cf.DNS(to_gpu(lib, ur), cf.make_flags(**fl))
    done = 0
    while done < nsteps:
        n = min(40, nsteps - done)
        gd.cfl()
        gd.advance(n)
        done += n
    u2, _ = gd.get()
    out = {"l2dist_to_ufinal": to_gpu(lib, vr).l2dist(u2), "norm_final": u2.l2norm(), "norm_ufinal": vr.l2norm()}
    return out, ur, u2


def device_vectors(lib, cfg):
    """Device state vectors: field2vector / vector2field kernels against the reference's host loops, and the Krylov
    algebra (dot / norm / axpy / axpby / scale) against NumPy."""
    ur = ref_random(cfg, 9)
    ug = to_gpu(lib, ur)
    xr = ur.to_vector()
    n = xr.size
    x = cf.DeviceVector(lib, n).from_field(ug)
    out = {"pack_max_abs": float(np.abs(x.get() - xr).max())}
    rng = np.random.default_rng(3)
    y_h = xr + 1e-3 * rng.standard_normal(n)
    y = cf.DeviceVector(lib, n).set(y_h)
    vr = ur.like().from_vector(y_h)
    vg = y.to_field(ug.like())
    out["unpack_rel"] = rel_l2(vg.get(), vr.data)
    out["dot_rel"] = abs(x.dot(y) - float(xr @ y_h)) / abs(float(xr @ y_h))
    out["norm_rel"] = abs(y.norm() - float(np.linalg.norm(y_h))) / float(np.linalg.norm(y_h))
    y.axpy(-0.75, x)
    out["axpy_max_abs"] = float(np.abs(y.get() - (y_h - 0.75 * xr)).max())
    y.axpby(2.0, x, 0.5)
    ref = 2.0 * xr + 0.5 * (y_h - 0.75 * xr)
    out["axpby_max_abs"] = float(np.abs(y.get() - ref).max())
    y.scale(-3.0)
    out["scale_max_abs"] = float(np.abs(y.get() + 3.0 * ref).max())
    return out


def smoke(cf_module=None):
    """One small DNS step of the hot path on cuda:0, checked against the oracle (used by __graft_entry__.smoke)."""
    lib = gpu_lib()
    cfg = dict(C1, Nx=16, Ny=17, Nz=16)
    r = dns_steps(lib, cfg, checkpoints=(1, 3))
    assert r[1] < 1e-12 and r[3] < 1e-12, r
    print("smoke ok:", r, "kernel launches:", lib.launch_count())
    return r


def orr_sommerfeld(lib, timestepping="sbdf3", nonlinearity="rot", T1=13.0, use_ref=False):
    """tests/dnsOrrsommTest.cpp, velocity part: a small-amplitude Orr-Sommerfeld eigenfunction (kx=1, kz=0, Re=7500, 4x65x4,
    golden fixture tests/data/os_ueig10_65.asc) on the parabolic base flow must evolve as exp(-i omega t) with the
    tabulated eigenvalue; returns the accumulated L2Dist(u_numerical, u_linear) over the checkpoints (reference
    tolerance for the sum of velocity and pressure errors with SBDF3: 2e-6)."""
    g = np.load(os.path.join(GOLDEN, "os_eig.npz"))
    Nx, Ny, Nz, Lx, Lz, a, b = 4, 65, 4, 2 * np.pi, np.pi, -1.0, 1.0
    nu, dt, N, scale = 1.0 / 7500.0, 0.02, 50, 1e-4
    omega = complex(g["omega"][0], g["omega"][1])
    lam = -1j * omega
    e = g["ueig"]
    prof = [refcf.cheby_make_spectral(e[:, 2 * i].copy()) + 1j * refcf.cheby_make_spectral(e[:, 2 * i + 1].copy()) for i in range(3)]
    Mz = Nz // 2 + 1

    def field_of(amp):
        c = np.zeros((3, Ny, Nx, Mz), dtype=np.complex128)
        for i in range(3):
            c[i, :, 1, 0] = amp * prof[i]
            c[i, :, Nx - 1, 0] = np.conj(amp * prof[i])
        return c.view(np.float64).reshape(3, Ny, Nx, 2 * Mz)

    def mk(arr):
        if use_ref:
            r = refcf.RefField(Nx, Ny, Nz, 3, Lx, Lz, a, b)
            r.data[...] = arr
            return r
        return cf.FlowField(lib, Nx, Ny, Nz, 3, Lx, Lz, a, b).set(arr)

    par = np.zeros((3, Ny, Nx, 2 * Mz))
    par[0, 0, 0, 0], par[0, 2, 0, 0] = 0.5, -0.5  # 1 - y^2 = T0/2 - T2/2
    c0 = scale * mk(par).l2norm() / mk(field_of(1.0)).l2norm()
    fl = dict(nu=nu, dt=dt, ulowerwall=0.0, uupperwall=0.0, baseflow="parabolic", constraint="bulkv", Ubulk=2.0 / 3, dPdx=-2 * nu,
              timestepping=timestepping, initstepping="cnrk2", nonlinearity=nonlinearity, dealiasing="none")
    un = mk(field_of(c0))
    dns = (refcf.RefDNS(un, refcf.make_flags(**fl)) if use_ref else cf.DNS(un, cf.make_flags(**fl)))
    err, t, amp = 0.0, 0.0, c0
    while t <= T1 + 1e-12:
        u, _ = dns.get()
        err += u.l2dist(mk(field_of(amp)))
        amp *= np.exp(lam * N * dt)
        dns.advance(N)
        t += N * dt
    return {"err": err, "norm": mk(field_of(c0)).l2norm()}


def layout_equivalence(lib, cfg, nsteps=4, junk=True, **flag_over):
    """The tile-major layout of the hot-path fields is an internal representation: a DNS run with it and one with
    CFGPU_SERIAL_LAYOUT=1 must give the same bits.  With `junk`, the initial field carries non-zero aliased modes and is
    not flagged padded: NSE::solve writes retained modes only (nse.cpp:566-572), so the junk must survive the steps
    untouched in both layouts (it never enters the de-aliased nonlinear term of this implementation)."""
    import os
    fl = dict(cfg["flags"]); fl.update(flag_over)
    ur = ref_random(cfg, 7)
    u0 = ur.data.copy()
    st = ur.state()
    Nx, Nz = cfg["Nx"], cfg["Nz"]
    Kx, Kz = Nx // 3 - 1, Nz // 3 - 1
    c0 = u0.view(np.complex128)
    alias = np.zeros(c0.shape, bool)
    alias[:, :, Kx + 1:Nx - Kx, :] = True
    alias[..., Kz + 1:] = True
    if junk:
        rng = np.random.default_rng(3)
        c0[alias] = 1e-3 * (rng.standard_normal(int(alias.sum())) + 1j * rng.standard_normal(int(alias.sum())))
    res = {}
    for mode in ("tile", "serial"):
        if mode == "serial":
            os.environ["CFGPU_SERIAL_LAYOUT"] = "1"
        try:
            ug = cf.FlowField(lib, ur.Nx, ur.Ny, ur.Nz, 3, ur.Lx, ur.Lz, ur.a, ur.b).set(u0, st[0], st[1], padded=not junk)
            gd = cf.DNS(ug, cf.make_flags(**fl))
            gd.advance(nsteps)
            u, q = gd.get()
            res[mode] = (u.get().copy(), q.get().copy())
        finally:
            os.environ.pop("CFGPU_SERIAL_LAYOUT", None)
    ut, us = res["tile"][0].view(np.complex128), res["serial"][0].view(np.complex128)
    return {"u_equal": bool(np.array_equal(res["tile"][0], res["serial"][0])),
            "q_equal": bool(np.array_equal(res["tile"][1], res["serial"][1])),
            # (without junk the field is uploaded as its retained box: the aliased modes are exactly zero on the device)
            "junk_kept": bool(np.array_equal(ut[alias], c0[alias] if junk else 0 * c0[alias])) and
                         bool(np.array_equal(us[alias], c0[alias] if junk else 0 * c0[alias])),
            "moved": float(np.abs(ut[~alias] - c0[~alias]).max())}


def graph_equivalence(lib, cfg, nsteps=100, **flag_over):
    """CUDA-graph replay of the SBDF step sequence (MultistepDNS::advance, CFGPU_GRAPH) against the eager launches:
    same kernels, same arguments, same order => the same bits; also returns the launch counts and the wall times."""
    import os
    import time
    fl = dict(cfg["flags"]); fl.update(flag_over)
    ur = ref_random(cfg, 11)
    res = {}
    for mode in ("0", "1"):
        os.environ["CFGPU_GRAPH"] = mode
        try:
            gd = cf.DNS(to_gpu(lib, ur), cf.make_flags(**fl))
            gd.advance(4)  # initial steps of the start-up scheme + first eager steps
            ctx_l0 = lib.launch_count() if hasattr(lib, "launch_count") else 0
            t0 = time.perf_counter()
            gd.advance(nsteps)
            u, q = gd.get()
            a = u.get().copy()
            dt = time.perf_counter() - t0
            res[mode] = dict(u=a, q=q.get().copy(), cfl=gd.cfl(), sec=dt, launches=(lib.launch_count() if hasattr(lib, "launch_count") else 0) - ctx_l0)
        finally:
            os.environ.pop("CFGPU_GRAPH", None)
    return {"u_identical": bool(np.array_equal(res["0"]["u"], res["1"]["u"])), "q_identical": bool(np.array_equal(res["0"]["q"], res["1"]["q"])),
            "cfl_identical": res["0"]["cfl"] == res["1"]["cfl"], "sec_eager": res["0"]["sec"], "sec_graph": res["1"]["sec"],
            "launches_eager": res["0"]["launches"], "launches_graph": res["1"]["launches"]}


SYMMETRIES = ((1, 1, 1, 1, 0.3, 0.1), (1, -1, 1, 1, 0.2, 0.0), (1, 1, -1, 1, 0.0, 0.0), (1, 1, 1, -1, 0.0, 0.37), (-1, -1, -1, -1, 0.5, 0.5),
              (1, -1, -1, 1, 0.5, 0.0), (-1, 1, 1, -1, 0.25, 0.5), (1, -1, 1, -1, 0.1, 0.2))


def symmetry_ops(lib, cfg, syms=SYMMETRIES):
    """FlowField *= FieldSymmetry on the device against flowfield.cpp:1274-1433, padded (de-aliased box) and full spectra."""
    worst = 0.0
    for padded in (True, False):
        for sym in syms:
            ur = ref_random(cfg, 3)
            if not padded:
                rng = np.random.default_rng(1)
                ur.data[...] = 1e-2 * rng.standard_normal(ur.data.shape)
                ur.set_padded(False)
                ur.set_state(0, 0)
                ur.make_spectral()   # a consistent full spectrum (all modes, Nyquist rows included)
            ug = to_gpu(lib, ur, padded=padded)
            ur.symmetry(*sym)
            ug.symmetry(*sym)
            worst = max(worst, rel_l2(ug.get(), ur.data))
    return worst


def load_eq(lib, scale=1.0):
    """tests/golden/eq.npz (reference tests/data/eq.nc, 24x33x24): the solution findsolnTest.cpp searches for, on the device."""
    g = np.load(os.path.join(GOLDEN, "eq.npz"))
    geo = dict(Nx=int(g["Nx"]), Ny=int(g["Ny"]), Nz=int(g["Nz"]), Lx=float(g["Lx"]), Lz=float(g["Lz"]), a=float(g["a"]), b=float(g["b"]))
    ur = refcf.RefField(geo["Nx"], geo["Ny"], geo["Nz"], 3, geo["Lx"], geo["Lz"], geo["a"], geo["b"]).load_padded_physical(g["u"])
    ur.make_spectral()
    ug = to_gpu(lib, ur)
    if scale != 1.0:
        ug.scale(scale)
    return ug, ur


EQ_SIGMA = (1, 1, 1, 1, 0.28168880386692519, 0.0)   # reference tests/data/sigmabest.asc


def findsoln_eq(lib, Nnewton=6, epsSearch=1e-11):
    """tests/findsolnTest.cpp (`findsoln -eqb -xrel -T 10 -sigma sigmabest`): perturb the stored travelling wave by 1.001 and
    let the Newton-Krylov-hookstep search pull it back (plane Couette Re 400, the x phase shift is an unknown of the search).
    Returns the search record and the distance to the stored field."""
    import time
    ug, ur = load_eq(lib, 1.001)
    fl = dict(C1["flags"])
    t0 = time.perf_counter()
    r = cf.hookstep_search(ug, cf.make_flags(**fl), 10.0, 0.03125, sigma=EQ_SIGMA, Nnewton=Nnewton, epsSearch=epsSearch, xrelative=True)
    r["seconds"] = time.perf_counter() - t0
    r["dist_to_stored"] = rel_l2(ug.get(), ur.data) * float(np.linalg.norm(ur.data.ravel())) / max(float(np.linalg.norm(ug.get().ravel())), 1e-300)
    sol = refcf.RefField(ur.Nx, ur.Ny, ur.Nz, 3, ur.Lx, ur.Lz, ur.a, ur.b)
    sol.data[...] = ug.get()
    sol.set_padded(True)
    r["l2dist_to_stored"] = sol.l2dist(ur)
    r["div_bc"] = (sol.divnorm(), sol.bcnorm())
    return r


def variable_dt_loop(lib, cfg, nintervals=4, dT=0.25, dt0=0.02, CFLmin=0.4, CFLmax=0.6, **flag_over):
    """The variable-time-step loop of simulateflow (programs/simulateflow.cpp:117-143; BASELINE configs[1]):
        CFL = dns.CFL(u); dns.advance(fields, dt.n()); if (dt.adjust(CFL)) dns.reset_dt(dt);
    on the reference and on the device with the same TimeStep object (this package's, fed with the device CFL).  Returns the
    relative CFL mismatch per interval, the number of dt changes, and the final field error."""
    fl = dict(cfg["flags"]); fl.update(flag_over); fl["dt"] = dt0
    ur = ref_random(cfg, 5, magn=float(cfg.get("magn", 0.2)))
    rd = refcf.RefDNS(ur, refcf.make_flags(**fl))
    gd = cf.DNS(to_gpu(lib, ur), cf.make_flags(**fl))
    L = lib.L
    ts = L.cf_timestep_create(dt0, 1e-4, 0.2, dT, CFLmin, CFLmax, 1)
    out = {"cfl_rel": [], "dt": [], "changes": 0, "steps": 0}
    for _ in range(nintervals):
        cg, cr = gd.cfl(), rd.cfl()
        out["cfl_rel"].append(abs(cg - cr) / max(abs(cr), 1e-300))
        n = L.cf_timestep_n(ts)
        rd.advance(n); gd.advance(n)
        out["steps"] += n
        out["dt"].append(L.cf_timestep_dt(ts))
        if L.cf_timestep_adjust(ts, cg):
            out["changes"] += 1
            dt = L.cf_timestep_dt(ts)
            rd.reset_dt(dt); gd.reset_dt(dt)
    L.cf_timestep_free(ts)
    u1, q1 = rd.get(); u2, q2 = gd.get()
    out["u_rel"] = rel_l2(u2.get(), u1.data)
    out["q_rel"] = rel_l2(q2.get(), q1.data)
    out["dPdx"] = abs(gd.dPdx() - rd.dPdx())
    return out


def netcdf_reader(lib):
    """FlowField("file.nc") through the library-free NetCDF-4 reader (host/ncfile.cpp) on a file written by stock Channelflow
    (tests/golden/eq.nc = reference tests/data/eq.nc: de-aliased I/O grid 16x33x16 of a 24x33x24 field) against the
    reference's own embedding of the same values (addPaddedModes, flowfield.cpp:2991-3190)."""
    h = lib.L.cf_field_load(os.path.join(GOLDEN, "eq.nc").encode())
    _, ur = load_eq(lib)
    v = cf.FlowField(lib, ur.Nx, ur.Ny, ur.Nz, 3, ur.Lx, ur.Lz, ur.a, ur.b, handle=h)
    return {"rel": rel_l2(v.get(), ur.data), "padded": v.padded()}


def poincare_section(lib, cfg, kind="plane", nstride=5, maxstrides=40, frac=0.97, seed=7, crosssign=0, **flag_over):
    """DNSPoincare::advanceToSection (host/poincare.cpp) against a restatement of the documented algorithm (dns.cpp:520-700)
    that uses the compiled reference for everything it is built from: the reference's DNS for the coarse strides and the
    step-by-step second pass, the reference's L2IP / wallshear / dissipation for h, numpy for the quadratic interpolant in
    time and the Newton iteration.  (The reference's own advanceToSection advances a copy of its arguments, dns.cpp:526-528,
    and therefore never sees a crossing; see channelflow/dns.h in this package.)  Returns times, h and the field mismatch."""
    fl = dict(cfg["flags"]); fl.update(flag_over)
    dt = fl["dt"]
    u0 = ref_random(cfg, seed, magn=float(cfg.get("magn", 0.2)))
    if kind == "plane":
        e = u0.copy()
        c = frac * u0.l2ip(e)
        h = lambda f: f.l2ip(e) - c  # noqa: E731
    else:
        # I - D of the total velocity: zero base flow, the laminar profile y carried by the field itself
        fl["baseflow"] = "zero"
        u0.cdata[0, 1, 0, 0] += 1.0
        h = lambda f: f.wallshear() - f.dissipation()  # noqa: E731
    # ---- device
    ug = to_gpu(lib, u0); qg = ug.like(Nd=1)
    if kind == "plane":
        eg = to_gpu(lib, u0); sg = to_gpu(lib, u0); sg.scale(frac)
        res = cf.poincare_search(ug, qg, cf.make_flags(**fl), nstride, maxstrides, ustar=sg, estar=eg, crosssign=crosssign)
    else:
        res = cf.poincare_search(ug, qg, cf.make_flags(**fl), nstride, maxstrides, crosssign=crosssign)
    out = {"found": res["found"], "t": res["t"], "h": res["h"], "sign": res["sign"], "strides": res["strides"]}
    # ---- restatement on the reference
    def lagr(xs, x):
        w = np.ones(len(xs))
        for i in range(len(xs)):
            for j in range(len(xs)):
                if j != i:
                    w[i] *= (x - xs[j]) / (xs[i] - xs[j])
        return w
    rd = refcf.RefDNS(u0, refcf.make_flags(**fl))
    tmp = u0.copy()
    def h_of(arr):
        tmp.data[...] = arr
        return h(tmp)
    ref = {"found": False, "strides": 0}
    us, qs = u0.copy(), u0.like(Nd=1)
    for stride in range(maxstrides):
        rd.advance(nstride)
        ue, qe = rd.get()
        ref["strides"] += 1
        h0, h1 = h(us), h(ue)
        up, down = h0 < 0 <= h1, h0 > 0 >= h1
        if (crosssign > 0 and up) or (crosssign < 0 and down) or (crosssign == 0 and (up or down)):
            tstart = rd.time() - dt * nstride
            fl2 = dict(fl); fl2["t0"] = tstart
            fd = refcf.RefDNS(us, refcf.make_flags(**fl2), q=qs)
            ts, hs, ua = [tstart], [h0], [us.data.copy()]
            for k in range(nstride + 2):
                fd.advance(1)
                uk, _ = fd.get()
                ts.insert(0, tstart + (k + 1) * dt); hs.insert(0, h(uk)); ua.insert(0, uk.data.copy())
                ts, hs, ua = ts[:3], hs[:3], ua[:3]
                if len(ts) == 3 and ((hs[2] < 0 <= hs[0]) or (hs[2] > 0 >= hs[0])):
                    s = float(np.dot(lagr(hs, 0.0), ts))
                    for it in range(6):
                        v = sum(w * a for w, a in zip(lagr(ts, s), ua))
                        g = h_of(v)
                        if abs(g) < 0.5e-13 or it == 5:
                            break
                        vd = sum(w * a for w, a in zip(lagr(ts, s + 1e-9 * s), ua))
                        s -= g / ((h_of(vd) - g) / (1e-9 * s))
                    ref.update(found=True, t=s, h=g, sign=1 if h0 < 0 else -1, u=v)
                    break
            break
        us, qs = ue, qe
    out["ref_found"] = ref["found"]; out["ref_strides"] = ref["strides"]
    if ref["found"] and res["found"]:
        out["ref_t"] = ref["t"]; out["ref_sign"] = ref["sign"]; out["ref_h"] = ref["h"]
        out["dt_cross"] = abs(ref["t"] - res["t"])
        out["u_rel"] = rel_l2(res["ucrossing"].get(), ref["u"])
        # h of the device crossing field measured by the reference
        out["h_by_ref"] = h_of(np.asarray(res["ucrossing"].get()).reshape(tmp.data.shape))
    return out


def dns_equivariance(lib, cfg, sym=(1, -1, -1, 1, 0.5, 0.0), n1=6, n2=6, seed=3, **flag_over):
    """DNS::operator*= (dns.cpp:175-180, dnsalgo.cpp:264-272): for a symmetry sigma of plane Couette flow, mapping the running
    multistep DNS (state and history) after n1 steps and taking n2 more must equal sigma of the unmapped run after n1 + n2
    steps.  Also returns what happens when only the state is mapped (the history matters)."""
    fl = dict(cfg["flags"]); fl.update(flag_over)
    u0 = to_gpu(lib, ref_random(cfg, seed, magn=float(cfg.get("magn", 0.2))))
    a = cf.DNS(u0, cf.make_flags(**fl)); a.advance(n1 + n2)
    ua, _ = a.get(); ua.symmetry(*sym)
    b = cf.DNS(u0, cf.make_flags(**fl)); b.advance(n1); b.symmetry(*sym); b.advance(n2)
    ub, _ = b.get()
    c = cf.DNS(u0, cf.make_flags(**fl)); c.advance(n1)
    uc, qc = c.get(); uc.symmetry(*sym); c.set(uc, None); c.advance(n2)   # state only: the history is stale
    uc, _ = c.get()
    return {"mapped": rel_l2(ub.get(), ua.get()), "state_only": rel_l2(uc.get(), ua.get())}


def netcdf_writer(lib, tmpdir):
    """FlowField::writeNetCDF (host/ncfile.cpp: the reference's dimensions, variables and attributes in the classic CDF-2
    container).  The field of the reference's own eq.nc is loaded and written back; scipy.io.netcdf_file -- an independent
    implementation of the format -- must find the reference's schema and the values stock Channelflow wrote (eq.npz holds
    them), and this package's reader must return the same field from both files.  Also the full-grid (unpadded) case."""
    from scipy.io import netcdf_file
    g = np.load(os.path.join(GOLDEN, "eq.npz"))
    src = os.path.join(GOLDEN, "eq.nc")
    h = lib.L.cf_field_load(src.encode())
    ug, ur = load_eq(lib)
    v = cf.FlowField(lib, ur.Nx, ur.Ny, ur.Nz, 3, ur.Lx, ur.Lz, ur.a, ur.b, handle=h)
    out = os.path.join(str(tmpdir), "eq_out.nc")
    v.save(out)
    res = {}
    with netcdf_file(out, "r", mmap=False) as nc:
        res["dims"] = {k: int(n) for k, n in nc.dimensions.items()}
        res["vars"] = list(nc.variables.keys())
        res["attrs"] = {k: getattr(nc, k) for k in ("Nx", "Ny", "Nz", "Lx", "Lz", "a", "b")}
        res["title"] = nc.title.decode() if isinstance(nc.title, bytes) else str(nc.title)
        vals = np.stack([np.array(nc.variables[n][:]) for n in ("Velocity_X", "Velocity_Y", "Velocity_Z")])
        res["var_dims"] = nc.variables["Velocity_X"].dimensions
        res["grid_err"] = max(float(np.abs(np.array(nc.variables["X"][:]) - np.arange(res["dims"]["X"]) * ur.Lx / res["dims"]["X"]).max()),
                              float(np.abs(np.array(nc.variables["Y"][:]) - np.cos(np.pi * np.arange(ur.Ny) / (ur.Ny - 1))).max()))
    res["values_abs"] = float(np.abs(vals - np.asarray(g["u"]).reshape(vals.shape)).max())
    h2 = lib.L.cf_field_load(out.encode())
    v2 = cf.FlowField(lib, ur.Nx, ur.Ny, ur.Nz, 3, ur.Lx, ur.Lz, ur.a, ur.b, handle=h2)
    res["reread_rel"] = rel_l2(v2.get(), v.get())
    res["reread_padded"] = v2.padded()
    # full grid: an unpadded field keeps every mode
    w = to_gpu(lib, ref_random(C1, 4), padded=False)
    out2 = os.path.join(str(tmpdir), "full.nc")
    w.save(out2)
    h3 = lib.L.cf_field_load(out2.encode())
    w2 = cf.FlowField(lib, C1["Nx"], C1["Ny"], C1["Nz"], 3, C1["Lx"], C1["Lz"], C1["a"], C1["b"], handle=h3)
    res["full_rel"] = rel_l2(w2.get(), w.get())
    res["full_padded"] = w2.padded()
    with netcdf_file(out2, "r", mmap=False) as nc:
        res["full_dims"] = {k: int(n) for k, n in nc.dimensions.items()}
    # a classic file written by scipy (CDF-1 container) with the same schema is readable too
    out3 = os.path.join(str(tmpdir), "scipy.nc")
    with netcdf_file(out3, "w", version=1) as nc:
        for k in ("Nx", "Ny", "Nz"):
            setattr(nc, k, np.int32(getattr(ur, k)))
        for k in ("Lx", "Lz", "a", "b"):
            setattr(nc, k, np.float64(getattr(ur, k)))
        nz, ny, nx = vals.shape[1:]
        nc.createDimension("X", nx); nc.createDimension("Y", ny); nc.createDimension("Z", nz)
        for n, ln in (("X", nx), ("Y", ny), ("Z", nz)):
            nc.createVariable(n, "d", (n,))[:] = np.zeros(ln)
        for i, n in enumerate(("Velocity_X", "Velocity_Y", "Velocity_Z")):
            nc.createVariable(n, "d", ("Z", "Y", "X"))[:] = np.asarray(g["u"]).reshape(vals.shape)[i]
    h4 = lib.L.cf_field_load(out3.encode())
    v4 = cf.FlowField(lib, ur.Nx, ur.Ny, ur.Nz, 3, ur.Lx, ur.Lz, ur.a, ur.b, handle=h4)
    res["scipy_rel"] = rel_l2(v4.get(), ur.data)
    return res
