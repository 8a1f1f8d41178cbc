"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref/libchflow_ref.so, i.e. the *unmodified* Channelflow
reference (compiled by oracle/Makefile from /root/reference) behind the C driver oracle/ref_driver.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module; the product package
channelflow_b200 never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libchflow_ref.so")

# enum values, channelflow/dnsflags.h:24-41 and cfbasics/mathdefs.h (fieldstate)
PHYSICAL, SPECTRAL = 0, 1
BASEFLOW = dict(zero=0, linear=1, parabolic=2, laminar=3, suction=4, arbitrary=5)
CONSTRAINT = dict(gradp=0, bulkv=1)
STEPPER = dict(cnfe1=0, cnab2=1, cnrk2=2, smrk2=3, sbdf1=4, sbdf2=5, sbdf3=6, sbdf4=7)
NONLIN = dict(rot=0, conv=1, div=2, skew=3, alt=4, alt_=5, linear=6)
DEALIAS = dict(none=0, xz=1, y=2, xyz=3)


class RefFlags(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("nu", "dPdx", "dPdz", "Ubulk", "Wbulk", "ulowerwall", "uupperwall", "wlowerwall", "wupperwall",
                 "Vsuck", "rotation", "t0", "dt")] + \
               [(n, C.c_int) for n in
                ("baseflow", "constraint", "timestepping", "initstepping", "nonlinearity", "dealiasing",
                 "taucorrection")]


def make_flags(nu=0.0025, dPdx=0.0, dPdz=0.0, Ubulk=0.0, Wbulk=0.0, ulowerwall=0.0, uupperwall=0.0, wlowerwall=0.0,
               wupperwall=0.0, Vsuck=0.0, rotation=0.0, t0=0.0, dt=0.03125, baseflow="laminar", constraint="gradp",
               timestepping="sbdf3", initstepping="smrk2", nonlinearity="rot", dealiasing="xz", taucorrection=True):
    """Same defaults as DNSFlags::DNSFlags (dnsflags.h:84-94)."""
    f = RefFlags()
    f.nu, f.dPdx, f.dPdz, f.Ubulk, f.Wbulk = nu, dPdx, dPdz, Ubulk, Wbulk
    f.ulowerwall, f.uupperwall, f.wlowerwall, f.wupperwall = ulowerwall, uupperwall, wlowerwall, wupperwall
    f.Vsuck, f.rotation, f.t0, f.dt = Vsuck, rotation, t0, dt
    f.baseflow = BASEFLOW[baseflow]
    f.constraint = CONSTRAINT[constraint]
    f.timestepping = STEPPER[timestepping]
    f.initstepping = STEPPER[initstepping]
    f.nonlinearity = NONLIN[nonlinearity]
    f.dealiasing = DEALIAS[dealiasing]
    f.taucorrection = 1 if taucorrection else 0
    return f


_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, d, i, dp = C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double)
        L.ref_field_create.restype = vp
        L.ref_field_create.argtypes = [i, i, i, i, d, d, d, d]
        L.ref_field_data.restype = dp
        L.ref_field_nloc.restype = C.c_long
        for n in ("ref_field_free", "ref_field_data", "ref_field_nloc", "ref_field_zero", "ref_make_physical",
                  "ref_make_spectral", "ref_make_physical_y", "ref_make_spectral_y", "ref_make_physical_xz",
                  "ref_make_spectral_xz", "ref_zero_padded_modes", "ref_l2norm", "ref_divnorm", "ref_bcnorm",
                  "ref_field2vector_size", "ref_field_padded", "ref_dns_free", "ref_dns_cfl", "ref_dns_time",
                  "ref_dns_dPdx", "ref_dns_Ubulk"):
            getattr(L, n).argtypes = [vp]
        for n in ("ref_l2norm", "ref_l2dist", "ref_l2ip", "ref_divnorm", "ref_bcnorm", "ref_l2norm2", "ref_dns_cfl",
                  "ref_dns_time", "ref_dns_dPdx", "ref_dns_Ubulk"):
            getattr(L, n).restype = d
        for n in ("ref_wallshear", "ref_dissipation"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = d
        L.ref_l2norm2.argtypes = [vp, i]
        L.ref_l2dist.argtypes = [vp, vp]
        L.ref_l2ip.argtypes = [vp, vp]
        L.ref_field_set_state.argtypes = [vp, i, i]
        L.ref_field_get_state.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
        L.ref_field_set_padded.argtypes = [vp, i]
        L.ref_field_copy.argtypes = [vp, vp]
        L.ref_field_symmetry.argtypes = [vp, i, i, i, i, d, d]
        L.ref_randomfield.argtypes = [vp, i, d, d, i]
        L.ref_load_padded_physical.argtypes = [vp, dp, i, i]
        L.ref_field2vector.argtypes = [vp, dp]
        L.ref_vector2field.argtypes = [dp, C.c_long, vp]
        L.ref_nonlinear.argtypes = [vp, vp, C.POINTER(RefFlags)]
        L.ref_base_profiles.argtypes = [vp, C.POINTER(RefFlags), dp, dp]
        L.ref_dns_create.restype = vp
        L.ref_dns_create.argtypes = [vp, vp, C.POINTER(RefFlags)]
        L.ref_dns_advance.argtypes = [vp, i]
        L.ref_dns_get.argtypes = [vp, vp, vp]
        L.ref_dns_set.argtypes = [vp, vp, vp]
        L.ref_dns_reset_dt.argtypes = [vp, d]
        L.ref_tausolve.argtypes = [i, i, d, d, d, d, d, d, i, i, dp, dp, dp, dp, dp, dp, dp]
        L.ref_tausolve_bulk.argtypes = [d, d, d, d, d, d, i, i, dp, dp, dp, d, d, dp, dp, dp]
        L.ref_helmholtz.argtypes = [i, d, d, d, d, dp, d, d, dp]
        L.ref_cheby_make_physical.argtypes = [i, dp]
        L.ref_cheby_make_spectral.argtypes = [i, dp]
        L.ref_laminar_profile.argtypes = [C.POINTER(RefFlags), d, d, i, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefField:
    """chflow::FlowField of the reference.  `.data` is a numpy view of its storage shaped [Nd][Ny][Nx][Nzpad]
    (reference serial layout, flowfield.h:370-385); `.cdata` the complex alias [Nd][Ny][Nx][Mz]."""

    def __init__(self, Nx, Ny, Nz, Nd, Lx, Lz, a=-1.0, b=1.0):
        self.Nx, self.Ny, self.Nz, self.Nd, self.Lx, self.Lz, self.a, self.b = Nx, Ny, Nz, Nd, Lx, Lz, a, b
        self.Mz = Nz // 2 + 1
        self.h = lib().ref_field_create(Nx, Ny, Nz, Nd, Lx, Lz, a, b)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_field_free(self.h)
            self.h = None

    def like(self, Nd=None):
        return RefField(self.Nx, self.Ny, self.Nz, self.Nd if Nd is None else Nd, self.Lx, self.Lz, self.a, self.b)

    @property
    def data(self):
        n = lib().ref_field_nloc(self.h)
        p = lib().ref_field_data(self.h)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.Nd, self.Ny, self.Nx, 2 * self.Mz)

    @property
    def cdata(self):
        return self.data.view(np.complex128)

    def set_state(self, xz, y):
        lib().ref_field_set_state(self.h, xz, y)

    def state(self):
        a, b = C.c_int(), C.c_int()
        lib().ref_field_get_state(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_padded(self, p):
        lib().ref_field_set_padded(self.h, 1 if p else 0)

    def copy(self):
        o = self.like()
        lib().ref_field_copy(o.h, self.h)
        return o

    def randomfield(self, seed=1, magn=0.2, smooth=0.4, meanflow=False):
        lib().ref_randomfield(self.h, seed, magn, smooth, 1 if meanflow else 0)
        return self

    def load_padded_physical(self, var):
        var = np.ascontiguousarray(var, dtype=np.float64)
        Nd, Nz_io, Ny, Nx_io = var.shape
        assert Nd == self.Nd and Ny == self.Ny
        lib().ref_load_padded_physical(self.h, _dp(var), Nx_io, Nz_io)
        return self

    def make_physical(self): lib().ref_make_physical(self.h)
    def make_spectral(self): lib().ref_make_spectral(self.h)
    def make_physical_y(self): lib().ref_make_physical_y(self.h)
    def make_spectral_y(self): lib().ref_make_spectral_y(self.h)
    def make_physical_xz(self): lib().ref_make_physical_xz(self.h)
    def make_spectral_xz(self): lib().ref_make_spectral_xz(self.h)
    def zero_padded_modes(self): lib().ref_zero_padded_modes(self.h)
    def symmetry(self, s, sx, sy, sz, ax, az): lib().ref_field_symmetry(self.h, s, sx, sy, sz, ax, az)
    def l2norm(self): return lib().ref_l2norm(self.h)
    def l2norm3d(self):
        lib().ref_l2norm3d.restype = C.c_double
        lib().ref_l2norm3d.argtypes = [C.c_void_p]
        return lib().ref_l2norm3d(self.h)

    def l2dist(self, o): return lib().ref_l2dist(self.h, o.h)
    def l2ip(self, o): return lib().ref_l2ip(self.h, o.h)
    def divnorm(self): return lib().ref_divnorm(self.h)
    def wallshear(self): return lib().ref_wallshear(self.h)
    def dissipation(self): return lib().ref_dissipation(self.h)
    def bcnorm(self): return lib().ref_bcnorm(self.h)

    def to_vector(self):
        n = lib().ref_field2vector_size(self.h)
        x = np.zeros(n)
        lib().ref_field2vector(self.h, _dp(x))
        return x

    def from_vector(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib().ref_vector2field(_dp(x), x.size, self.h)
        return self


class RefDNS:
    """chflow::DNS of the reference (dns.cpp:22-163) owning copies of (u, q)."""

    def __init__(self, u, flags, q=None):
        self.u_geom = u
        if q is None:
            q = u.like(Nd=1)
        self.flags = flags
        self.h = lib().ref_dns_create(u.h, q.h, C.byref(flags))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_dns_free(self.h)
            self.h = None

    def advance(self, n):
        lib().ref_dns_advance(self.h, n)

    def get(self):
        u = self.u_geom.like()
        q = self.u_geom.like(Nd=1)
        lib().ref_dns_get(self.h, u.h, q.h)
        return u, q

    def cfl(self): return lib().ref_dns_cfl(self.h)
    def reset_dt(self, dt): lib().ref_dns_reset_dt(self.h, dt)
    def time(self): return lib().ref_dns_time(self.h)
    def dPdx(self): return lib().ref_dns_dPdx(self.h)
    def Ubulk(self): return lib().ref_dns_Ubulk(self.h)


def nonlinear(u, flags):
    f = u.like()
    lib().ref_nonlinear(u.h, f.h, C.byref(flags))
    return f


def base_profiles(u, flags):
    U, W = np.zeros(u.Ny), np.zeros(u.Ny)
    lib().ref_base_profiles(u.h, C.byref(flags), _dp(U), _dp(W))
    return U, W


def tausolve(kx, kz, Lx, Lz, a, b, lam, nu, N, Rx, Ry, Rz, taucorr=True):
    """Rx,Ry,Rz complex length-N. Returns u,v,w,P complex."""
    arrs = []
    for R in (Rx, Ry, Rz):
        arrs += [np.ascontiguousarray(R.real, dtype=np.float64), np.ascontiguousarray(R.imag, dtype=np.float64)]
    out = np.zeros((8, N))
    lib().ref_tausolve(kx, kz, Lx, Lz, a, b, lam, nu, N, 1 if taucorr else 0, *[_dp(x) for x in arrs], _dp(out))
    return tuple(out[2 * c] + 1j * out[2 * c + 1] for c in range(4))


def tausolve_bulk(Lx, Lz, a, b, lam, nu, N, Rx, Ry, Rz, umean, wmean, taucorr=True):
    arrs = [np.ascontiguousarray(np.real(R), dtype=np.float64) for R in (Rx, Ry, Rz)]
    out = np.zeros((8, N))
    dpx, dpz = C.c_double(0), C.c_double(0)
    lib().ref_tausolve_bulk(Lx, Lz, a, b, lam, nu, N, 1 if taucorr else 0, *[_dp(x) for x in arrs], umean, wmean,
                            _dp(out), C.byref(dpx), C.byref(dpz))
    return tuple(out[2 * c] + 1j * out[2 * c + 1] for c in range(4)) + (dpx.value, dpz.value)


def helmholtz(N, a, b, lam, nu, f, ua=0.0, ub=0.0):
    f = np.ascontiguousarray(f, dtype=np.float64)
    u = np.zeros(N)
    lib().ref_helmholtz(N, a, b, lam, nu, _dp(f), ua, ub, _dp(u))
    return u


def cheby_make_physical(c):
    c = np.array(c, dtype=np.float64)
    lib().ref_cheby_make_physical(c.size, _dp(c))
    return c


def cheby_make_spectral(c):
    c = np.array(c, dtype=np.float64)
    lib().ref_cheby_make_spectral(c.size, _dp(c))
    return c


def laminar_profile(flags, a, b, Ny):
    U = np.zeros(Ny)
    lib().ref_laminar_profile(C.byref(flags), a, b, Ny, _dp(U))
    return U
