"""channelflow_b200 -- B200 (sm_100a) implementation of Channelflow's DNS time-step hot path.

This Python module is only a thin ctypes view of the two native libraries (it exists for the tests and bench.py):

  libcfgpu.so        CUDA kernels + the C-ABI of include/cfgpu.h
  libchflow_b200.so  C++ host classes mirroring the reference API (chflow::FlowField / DNS / DNSFlags / NSE ...)
                     on top of that C-ABI, plus a flat C driver for them (host/capi.cpp)

There is no CPU fallback: if the CUDA library is missing or no GPU is visible, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_GPU = os.path.join(_HERE, "libcfgpu.so")
LIB_HOST = os.path.join(_HERE, "libchflow_b200.so")

PHYSICAL, SPECTRAL = 0, 1

# every symbol include/cfgpu.h declares
CFGPU_SYMBOLS = """cfgpu_last_error cfgpu_version cfgpu_init cfgpu_finalize cfgpu_sync cfgpu_launch_count cfgpu_timer_start
cfgpu_timer_stop cfgpu_profile_enable cfgpu_profile_read cfgpu_graph_begin cfgpu_graph_end cfgpu_graph_launch cfgpu_graph_abort cfgpu_graph_destroy cfgpu_field_create cfgpu_field_destroy
cfgpu_host_alloc cfgpu_host_free cfgpu_field_upload cfgpu_field_download cfgpu_field_upload_padded cfgpu_field_download_padded cfgpu_field_copy cfgpu_field_swap cfgpu_field_zero cfgpu_field_set_state
cfgpu_field_get_state cfgpu_field_set_padded cfgpu_field_get_padded cfgpu_field_device_ptr cfgpu_field_axpby
cfgpu_field_scale cfgpu_field_get_profile cfgpu_field_add_profile cfgpu_field_zero_padded_modes
cfgpu_field_make_physical_y cfgpu_field_make_spectral_y cfgpu_field_make_physical_xz cfgpu_field_make_spectral_xz
cfgpu_field_make_physical cfgpu_field_make_spectral cfgpu_l2norm2 cfgpu_l2norm2_3d cfgpu_l2dist2 cfgpu_l2ip cfgpu_nse_create
cfgpu_nse_destroy cfgpu_nse_set_constraint cfgpu_nse_reset_lambda cfgpu_nse_nonlinear cfgpu_nse_solve
cfgpu_nse_linear cfgpu_nse_cflfactor cfgpu_nse_get_dPd cfgpu_comm_unique_id cfgpu_comm_init_nccl
cfgpu_comm_init_external cfgpu_comm_rank cfgpu_comm_ranges cfgpu_field_allgather
cfgpu_field_copy_component cfgpu_l2form_box cfgpu_bcnorm2 cfgpu_field_diffop cfgpu_field_pointwise
cfgpu_helmholtz_solve cfgpu_tridiag cfgpu_tausolve_mode cfgpu_poisson_solve cfgpu_pressure_neumann cfgpu_field_symmetry cfgpu_chebyform
cfgpu_vec_create cfgpu_vec_destroy cfgpu_vec_size cfgpu_vec_upload cfgpu_vec_download cfgpu_vec_copy cfgpu_vec_zero cfgpu_vec_dot
cfgpu_vec_nrm2 cfgpu_vec_axpy cfgpu_vec_axpby cfgpu_vec_scal cfgpu_field2vector_size cfgpu_field2vector cfgpu_vector2field""".split()


class CfgpuError(RuntimeError):
    pass


class NseConfig(C.Structure):
    _fields_ = [("nu", C.c_double), ("Vsuck", C.c_double), ("rotation", C.c_double), ("nonlinearity", C.c_int),
                ("dealias_xz", C.c_int), ("dealias_y", C.c_int), ("taucorrection", C.c_int), ("constraint", C.c_int),
                ("dPdxRef", C.c_double), ("dPdzRef", C.c_double), ("UbulkRef_minus_base", C.c_double),
                ("WbulkRef_minus_base", C.c_double)]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class GpuLib:
    """ctypes view of the C-ABI (include/cfgpu.h)."""

    def __init__(self, path=None):
        path = path or LIB_GPU
        if not os.path.exists(path):
            raise CfgpuError("CUDA library %s is missing: run `python __graft_entry__.py` (there is no CPU fallback)" % path)
        self.path = path
        self.L = L = C.CDLL(path)  # RTLD_LOCAL: the test-only emulation build may live in the same process
        vp, d, i, dpt = C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double)
        L.cfgpu_last_error.restype = C.c_char_p
        L.cfgpu_version.restype = C.c_char_p
        L.cfgpu_init.argtypes = [i, C.POINTER(vp)]
        L.cfgpu_finalize.argtypes = [vp]
        L.cfgpu_sync.argtypes = [vp]
        L.cfgpu_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.cfgpu_timer_start.argtypes = [vp]
        L.cfgpu_timer_stop.argtypes = [vp, dpt]
        L.cfgpu_graph_begin.argtypes = [vp]
        L.cfgpu_graph_end.argtypes = [vp, C.POINTER(i)]
        L.cfgpu_graph_launch.argtypes = [vp, i]
        L.cfgpu_graph_abort.argtypes = [vp]
        L.cfgpu_graph_destroy.argtypes = [vp, i]
        L.cfgpu_field_create.argtypes = [vp, i, i, i, i, d, d, d, d, C.POINTER(vp)]
        for n in ("destroy", "zero", "zero_padded_modes", "make_physical_y", "make_spectral_y", "make_physical_xz",
                  "make_spectral_xz", "make_physical", "make_spectral"):
            getattr(L, "cfgpu_field_" + n).argtypes = [vp]
        L.cfgpu_field_upload.argtypes = [vp, dpt, i, i]
        L.cfgpu_field_download.argtypes = [vp, dpt]
        L.cfgpu_field_copy.argtypes = [vp, vp]
        L.cfgpu_field_swap.argtypes = [vp, vp]
        L.cfgpu_field_set_state.argtypes = [vp, i, i]
        L.cfgpu_field_get_state.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
        L.cfgpu_field_set_padded.argtypes = [vp, i]
        L.cfgpu_field_get_padded.argtypes = [vp, C.POINTER(i)]
        L.cfgpu_field_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_longlong)]
        L.cfgpu_field_axpby.argtypes = [vp, d, vp, d, vp]
        L.cfgpu_field_scale.argtypes = [vp, d]
        L.cfgpu_field_get_profile.argtypes = [vp, i, i, i, dpt]
        L.cfgpu_field_add_profile.argtypes = [vp, i, i, i, dpt, d]
        L.cfgpu_l2norm2.argtypes = [vp, i, dpt]
        L.cfgpu_l2dist2.argtypes = [vp, vp, i, dpt]
        L.cfgpu_l2ip.argtypes = [vp, vp, i, dpt]
        L.cfgpu_nse_create.argtypes = [vp, i, i, i, d, d, d, d, C.POINTER(NseConfig), dpt, dpt, C.POINTER(vp)]
        L.cfgpu_nse_destroy.argtypes = [vp]
        L.cfgpu_nse_set_constraint.argtypes = [vp, i, d, d, d, d]
        L.cfgpu_nse_reset_lambda.argtypes = [vp, dpt, i]
        L.cfgpu_nse_nonlinear.argtypes = [vp, vp, vp]
        L.cfgpu_nse_solve.argtypes = [vp, i, i, dpt, C.POINTER(vp), vp, vp]
        L.cfgpu_nse_linear.argtypes = [vp, vp, vp, vp]
        L.cfgpu_nse_cflfactor.argtypes = [vp, vp, dpt]
        L.cfgpu_nse_get_dPd.argtypes = [vp, dpt, dpt]
        L.cfgpu_poisson_solve.argtypes = [vp, vp, vp]
        L.cfgpu_field_symmetry.argtypes = [vp, i, i, i, i, d, d]
        L.cfgpu_chebyform.argtypes = [vp, vp, i, i, dpt]
        L.cfgpu_pressure_neumann.argtypes = [vp, vp, vp, d]
        L.cfgpu_helmholtz_solve.argtypes = [vp, i, d, d, d, d, i, dpt, dpt, dpt, dpt]
        L.cfgpu_tridiag.argtypes = [vp, i, i, dpt, dpt, dpt, dpt, i, i, i]
        L.cfgpu_tausolve_mode.argtypes = [vp, i, i, i, d, d, d, d, d, d, i, i, d, d, dpt, dpt, dpt]

    def check(self, status):
        if status != 0:
            raise CfgpuError(self.L.cfgpu_last_error().decode())

    def missing_symbols(self):
        return [s for s in CFGPU_SYMBOLS if not hasattr(self.L, s)]


class Context:
    def __init__(self, lib=None, device=0):
        self.lib = lib or GpuLib()
        self.h = C.c_void_p()
        self.lib.check(self.lib.L.cfgpu_init(device, C.byref(self.h)))

    def sync(self):
        self.lib.check(self.lib.L.cfgpu_sync(self.h))

    def launch_count(self):
        n = C.c_longlong()
        self.lib.check(self.lib.L.cfgpu_launch_count(self.h, C.byref(n)))
        return n.value

    def timer_start(self):
        self.lib.check(self.lib.L.cfgpu_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self.lib.check(self.lib.L.cfgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def field(self, Nx, Ny, Nz, Nd, Lx, Lz, a=-1.0, b=1.0):
        return Field(self, Nx, Ny, Nz, Nd, Lx, Lz, a, b)

    def helmholtz_solve(self, a, b, lam, nu, f, ua, ub):
        """HelmholtzSolver::solve for the rows of f [ncols][N] (real), Dirichlet data ua[ncols], ub[ncols]."""
        import numpy as np
        f = np.ascontiguousarray(np.atleast_2d(f), dtype=np.float64)
        ua = np.ascontiguousarray(np.broadcast_to(ua, f.shape[:1]), dtype=np.float64)
        ub = np.ascontiguousarray(np.broadcast_to(ub, f.shape[:1]), dtype=np.float64)
        u = np.empty_like(f)
        dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        self.lib.check(self.lib.L.cfgpu_helmholtz_solve(self.h, f.shape[1], a, b, lam, nu, f.shape[0], dp(f), dp(ua), dp(ub), dp(u)))
        return u

    def tausolve_mode(self, kx, kz, Lx, Lz, a, b, lam, nu, Rx, Ry, Rz, taucorr=True, bulk=None):
        """TauSolver::solve for one Fourier mode: complex profiles in, (u, v, w, P[, dPdx, dPdz]) out."""
        import numpy as np
        R = np.ascontiguousarray(np.stack([Rx, Ry, Rz]), dtype=np.complex128)
        out = np.empty((4, R.shape[1]), np.complex128)
        dPd = np.zeros(2)
        dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
        um, wm = bulk if bulk is not None else (0.0, 0.0)
        self.lib.check(self.lib.L.cfgpu_tausolve_mode(self.h, R.shape[1], kx, kz, Lx, Lz, a, b, lam, nu, 1 if taucorr else 0,
                                                      0 if bulk is None else 1, um, wm, dp(R), dp(out), dp(dPd)))
        return (out[0], out[1], out[2], out[3]) + ((dPd[0], dPd[1]) if bulk is not None else ())


class Field:
    """Device-resident FlowField storage (reference serial layout [Nd][Ny][Nx][Nzpad])."""

    def __init__(self, ctx, Nx, Ny, Nz, Nd, Lx, Lz, a=-1.0, b=1.0):
        self.ctx, self.lib = ctx, ctx.lib
        self.Nx, self.Ny, self.Nz, self.Nd, self.Lx, self.Lz, self.a, self.b = Nx, Ny, Nz, Nd, Lx, Lz, a, b
        self.Mz = Nz // 2 + 1
        self.h = C.c_void_p()
        self.lib.check(self.lib.L.cfgpu_field_create(ctx.h, Nx, Ny, Nz, Nd, Lx, Lz, a, b, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                self.lib.L.cfgpu_field_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def shape(self):
        return (self.Nd, self.Ny, self.Nx, 2 * self.Mz)

    def like(self, Nd=None):
        return Field(self.ctx, self.Nx, self.Ny, self.Nz, self.Nd if Nd is None else Nd, self.Lx, self.Lz, self.a, self.b)

    def upload(self, arr, xz=SPECTRAL, y=SPECTRAL, padded=None):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.size == int(np.prod(self.shape)), (arr.shape, self.shape)
        self.lib.check(self.lib.L.cfgpu_field_upload(self.h, _dp(arr), xz, y))
        if padded is not None:
            self.set_padded(padded)
        return self

    def download(self):
        out = np.empty(self.shape, dtype=np.float64)
        self.lib.check(self.lib.L.cfgpu_field_download(self.h, _dp(out)))
        return out

    def set_padded(self, p):
        self.lib.check(self.lib.L.cfgpu_field_set_padded(self.h, 1 if p else 0))

    def state(self):
        a, b = C.c_int(), C.c_int()
        self.lib.check(self.lib.L.cfgpu_field_get_state(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def copy_from(self, o): self.lib.check(self.lib.L.cfgpu_field_copy(self.h, o.h))
    def swap(self, o): self.lib.check(self.lib.L.cfgpu_field_swap(self.h, o.h))
    def zero(self): self.lib.check(self.lib.L.cfgpu_field_zero(self.h))
    def axpby(self, a, x, b=0.0, z=None): self.lib.check(self.lib.L.cfgpu_field_axpby(self.h, a, x.h, b, z.h if z else None))
    def scale(self, s): self.lib.check(self.lib.L.cfgpu_field_scale(self.h, s))
    def zero_padded_modes(self): self.lib.check(self.lib.L.cfgpu_field_zero_padded_modes(self.h))
    def make_physical(self): self.lib.check(self.lib.L.cfgpu_field_make_physical(self.h))
    def make_spectral(self): self.lib.check(self.lib.L.cfgpu_field_make_spectral(self.h))
    def make_physical_y(self): self.lib.check(self.lib.L.cfgpu_field_make_physical_y(self.h))
    def make_spectral_y(self): self.lib.check(self.lib.L.cfgpu_field_make_spectral_y(self.h))
    def make_physical_xz(self): self.lib.check(self.lib.L.cfgpu_field_make_physical_xz(self.h))
    def make_spectral_xz(self): self.lib.check(self.lib.L.cfgpu_field_make_spectral_xz(self.h))

    def get_profile(self, mx, mz, i):
        out = np.empty(2 * self.Ny)
        self.lib.check(self.lib.L.cfgpu_field_get_profile(self.h, mx, mz, i, _dp(out)))
        return out[0::2] + 1j * out[1::2]

    def add_profile(self, mx, mz, i, prof, scale=1.0):
        p = np.empty(2 * self.Ny)
        p[0::2], p[1::2] = np.real(prof), np.imag(prof)
        self.lib.check(self.lib.L.cfgpu_field_add_profile(self.h, mx, mz, i, _dp(p), scale))

    def l2norm2(self, normalize=True):
        v = C.c_double()
        self.lib.check(self.lib.L.cfgpu_l2norm2(self.h, 1 if normalize else 0, C.byref(v)))
        return v.value

    def l2norm(self, normalize=True):
        return float(np.sqrt(self.l2norm2(normalize)))

    def l2dist(self, o, normalize=True):
        v = C.c_double()
        self.lib.check(self.lib.L.cfgpu_l2dist2(self.h, o.h, 1 if normalize else 0, C.byref(v)))
        return float(np.sqrt(v.value))

    def l2ip(self, o, normalize=True):
        v = C.c_double()
        self.lib.check(self.lib.L.cfgpu_l2ip(self.h, o.h, 1 if normalize else 0, C.byref(v)))
        return v.value


class Nse:
    """Device NSE operator (cfgpu_nse_*): nonlinear term, batched tau solve, linear term, CFL."""

    def __init__(self, ctx, Nx, Ny, Nz, Lx, Lz, a, b, Ubase=None, Wbase=None, nu=0.0025, Vsuck=0.0, rotation=0.0,
                 nonlinearity=0, dealias_xz=True, dealias_y=False, taucorrection=True, constraint=0, dPdxRef=0.0,
                 dPdzRef=0.0, UbulkRef_minus_base=0.0, WbulkRef_minus_base=0.0):
        self.ctx, self.lib = ctx, ctx.lib
        cfg = NseConfig(nu, Vsuck, rotation, nonlinearity, int(dealias_xz), int(dealias_y), int(taucorrection), constraint,
                        dPdxRef, dPdzRef, UbulkRef_minus_base, WbulkRef_minus_base)
        U = np.ascontiguousarray(Ubase, dtype=np.float64) if Ubase is not None else None
        W = np.ascontiguousarray(Wbase, dtype=np.float64) if Wbase is not None else None
        self.h = C.c_void_p()
        self.lib.check(self.lib.L.cfgpu_nse_create(ctx.h, Nx, Ny, Nz, Lx, Lz, a, b, C.byref(cfg),
                                                   _dp(U) if U is not None else None,
                                                   _dp(W) if W is not None else None, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                self.lib.L.cfgpu_nse_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset_lambda(self, lambda_t):
        lt = np.ascontiguousarray(lambda_t, dtype=np.float64)
        self.lib.check(self.lib.L.cfgpu_nse_reset_lambda(self.h, _dp(lt), lt.size))

    def nonlinear(self, u, f):
        self.lib.check(self.lib.L.cfgpu_nse_nonlinear(self.h, u.h, f.h))

    def solve(self, s, coefs, terms, uout, qout):
        c = np.ascontiguousarray(coefs, dtype=np.float64)
        arr = (C.c_void_p * len(terms))(*[t.h for t in terms])
        self.lib.check(self.lib.L.cfgpu_nse_solve(self.h, s, len(terms), _dp(c), arr, uout.h, qout.h))

    def linear(self, u, q, L):
        self.lib.check(self.lib.L.cfgpu_nse_linear(self.h, u.h, q.h, L.h))

    def cflfactor(self, u):
        v = C.c_double()
        self.lib.check(self.lib.L.cfgpu_nse_cflfactor(self.h, u.h, C.byref(v)))
        return v.value

    def get_dPd(self):
        a, b = C.c_double(), C.c_double()
        self.lib.check(self.lib.L.cfgpu_nse_get_dPd(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


def check_symbols(path=None):
    """The C-ABI library loads and exports every symbol include/cfgpu.h declares (no GPU needed)."""
    lib = GpuLib(path)
    miss = lib.missing_symbols()
    if miss:
        raise CfgpuError("libcfgpu.so is missing symbols: " + ", ".join(miss))
    return True


# =====================================================================================================================
# Host classes (libchflow_b200.so): chflow::FlowField / NSE / DNS over the C-ABI, driven through host/capi.cpp
# =====================================================================================================================
BASEFLOW = dict(zero=0, linear=1, parabolic=2, laminar=3, suction=4, arbitrary=5)
CONSTRAINT = dict(gradp=0, bulkv=1)
STEPPER = dict(cnfe1=0, cnab2=1, cnrk2=2, smrk2=3, sbdf1=4, sbdf2=5, sbdf3=6, sbdf4=7)
NONLIN = dict(rot=0, conv=1, div=2, skew=3, alt=4, alt_=5, linear=6)
DEALIAS = dict(none=0, xz=1, y=2, xyz=3)


class Flags(C.Structure):
    """DNSFlags subset passed to the C drivers (same layout as oracle/ref_driver.cpp's RefFlags)."""
    _fields_ = [(n, C.c_double) for n in
                ("nu", "dPdx", "dPdz", "Ubulk", "Wbulk", "ulowerwall", "uupperwall", "wlowerwall", "wupperwall",
                 "Vsuck", "rotation", "t0", "dt")] + \
               [(n, C.c_int) for n in
                ("baseflow", "constraint", "timestepping", "initstepping", "nonlinearity", "dealiasing",
                 "taucorrection")]


def make_flags(nu=0.0025, dPdx=0.0, dPdz=0.0, Ubulk=0.0, Wbulk=0.0, ulowerwall=0.0, uupperwall=0.0, wlowerwall=0.0,
               wupperwall=0.0, Vsuck=0.0, rotation=0.0, t0=0.0, dt=0.03125, baseflow="laminar", constraint="gradp",
               timestepping="sbdf3", initstepping="smrk2", nonlinearity="rot", dealiasing="xz", taucorrection=True):
    """Defaults of DNSFlags::DNSFlags (reference dnsflags.h:84-94)."""
    f = Flags()
    f.nu, f.dPdx, f.dPdz, f.Ubulk, f.Wbulk = nu, dPdx, dPdz, Ubulk, Wbulk
    f.ulowerwall, f.uupperwall, f.wlowerwall, f.wupperwall = ulowerwall, uupperwall, wlowerwall, wupperwall
    f.Vsuck, f.rotation, f.t0, f.dt = Vsuck, rotation, t0, dt
    f.baseflow, f.constraint = BASEFLOW[baseflow], CONSTRAINT[constraint]
    f.timestepping, f.initstepping = STEPPER[timestepping], STEPPER[initstepping]
    f.nonlinearity, f.dealiasing = NONLIN[nonlinearity], DEALIAS[dealiasing]
    f.taucorrection = 1 if taucorrection else 0
    return f


class HostLib:
    """ctypes view of libchflow_b200.so (host/capi.cpp).  `gpu_path` must be the CUDA library it was linked to."""

    def __init__(self, path=None, gpu_path=None):
        path = path or LIB_HOST
        self.gpu = GpuLib(gpu_path)  # fails loudly if the CUDA library is missing
        if not os.path.exists(path):
            raise CfgpuError("host library %s is missing: run `python __graft_entry__.py`" % path)
        self.L = L = C.CDLL(path)
        vp, d, i, dpt = C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double)
        L.cf_field_create.restype = vp
        L.cf_field_create.argtypes = [i, i, i, i, d, d, d, d]
        L.cf_field_load.restype = vp
        L.cf_field_load.argtypes = [C.c_char_p]
        L.cf_field_save.argtypes = [vp, C.c_char_p]
        L.cf_field_nloc.restype = C.c_long
        for n in ("cf_field_free", "cf_field_nloc", "cf_field_zero", "cf_make_physical", "cf_make_spectral",
                  "cf_make_physical_y", "cf_make_spectral_y", "cf_make_physical_xz", "cf_make_spectral_xz",
                  "cf_zero_padded_modes", "cf_l2norm", "cf_field_padded", "cf_dns_free", "cf_dns_cfl", "cf_dns_time",
                  "cf_dns_dPdx", "cf_dns_Ubulk", "cf_timestep_free", "cf_timestep_n", "cf_timestep_N", "cf_timestep_dt",
                  "cf_timestep_dT", "cf_timestep_CFL"):
            getattr(L, n).argtypes = [vp]
        for n in ("cf_l2norm", "cf_l2dist", "cf_l2ip", "cf_l2norm2", "cf_dns_cfl", "cf_dns_time", "cf_dns_dPdx",
                  "cf_dns_Ubulk", "cf_timer_stop", "cf_timestep_dt", "cf_timestep_dT", "cf_timestep_CFL", "cf_cmplx_get"):
            getattr(L, n).restype = d
        L.cf_comm_unique_id.argtypes = [vp]
        L.cf_comm_init_nccl.argtypes = [i, i, vp]
        L.cf_comm_ranges.argtypes = [i, i, i, C.POINTER(i)]
        L.cf_field_allgather.argtypes = [vp]
        L.cf_field_upload.argtypes = [vp, dpt]
        L.cf_field_download.argtypes = [vp, dpt]
        L.cf_field_set_state.argtypes = [vp, i, i]
        L.cf_field_get_state.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
        L.cf_field_set_padded.argtypes = [vp, i]
        L.cf_field_copy.argtypes = [vp, vp]
        L.cf_field_symmetry.argtypes = [vp, i, i, i, i, d, d]
        L.cf_randomfield.argtypes = [vp, i, d, d, i]
        L.cf_hookstep_search.argtypes = [vp, C.POINTER(Flags), d, d, dpt, dpt, dpt, i]
        L.cf_dns_symmetry.argtypes = [vp, i, i, i, i, d, d]
        L.cf_poincare_search.argtypes = [vp, vp, C.POINTER(Flags), i, vp, vp, i, i, i, d, d, dpt, vp, vp]
        L.cf_cmplx_get.argtypes = [vp, i, i, i, i, i]
        L.cf_cmplx_set.argtypes = [vp, i, i, i, i, d, d]
        L.cf_l2norm2.argtypes = [vp, i]
        L.cf_l2dist.argtypes = [vp, vp]
        L.cf_l2ip.argtypes = [vp, vp]
        L.cf_field_axpby.argtypes = [vp, d, vp, d, vp]
        L.cf_field_scale.argtypes = [vp, d]
        L.cf_nonlinear.argtypes = [vp, vp, C.POINTER(Flags)]
        L.cf_base_profiles.argtypes = [vp, C.POINTER(Flags), dpt, dpt]
        L.cf_dns_create.restype = vp
        L.cf_dns_create.argtypes = [vp, vp, C.POINTER(Flags)]
        L.cf_dns_advance.argtypes = [vp, i]
        L.cf_dns_get.argtypes = [vp, vp, vp]
        L.cf_dns_set.argtypes = [vp, vp, vp]
        L.cf_dns_reset_dt.argtypes = [vp, d]
        L.cf_launch_count.restype = C.c_longlong
        L.cf_laminar_profile.argtypes = [C.POINTER(Flags), d, d, i, dpt]
        L.cf_timestep_create.restype = vp
        L.cf_timestep_create.argtypes = [d, d, d, d, d, d, i]
        L.cf_timestep_adjust.argtypes = [vp, d]
        L.cf_timestep_adjust_for_T.argtypes = [vp, d]

    def sync(self): self.L.cf_sync()

    # ---- multi-GPU (one process per GPU)
    def comm_init_nccl(self, rank, nranks, id_bytes):
        """id_bytes: the 128-byte NCCL unique id created by rank 0 (comm_unique_id) and broadcast by the launcher."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(id_bytes))
        self.L.cf_comm_init_nccl(rank, nranks, buf)

    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        if self.L.cf_comm_unique_id(buf) != 0:
            raise CfgpuError(self.gpu.L.cfgpu_last_error().decode())
        return bytes(buf)

    def comm_init_external(self, rank, nranks, exchange, allreduce):
        """Host-supplied collectives (CPU tests): exchange(peers, sendbufs, recvbufs) / allreduce(buf, op) on numpy views."""
        EX = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong),
                         C.POINTER(C.c_void_p), C.POINTER(C.c_longlong))
        AR = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int)

        def _ex(user, n, peer, sp, sb, rp, rb):
            try:
                sends = [(peer[i], np.ctypeslib.as_array((C.c_ubyte * sb[i]).from_address(sp[i])) if sb[i] else np.empty(0, np.uint8)) for i in range(n)]
                recvs = [(peer[i], np.ctypeslib.as_array((C.c_ubyte * rb[i]).from_address(rp[i])) if rb[i] else np.empty(0, np.uint8)) for i in range(n)]
                exchange(sends, recvs)
                return 0
            except Exception:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                return 1

        def _ar(user, buf, n, op):
            try:
                allreduce(np.ctypeslib.as_array(buf, shape=(n,)), op)
                return 0
            except Exception:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                return 1

        self._cb = (EX(_ex), AR(_ar))  # keep alive
        self.L.cf_comm_init_external(rank, nranks, self._cb[0], self._cb[1])

    def comm_ranges(self, nmx, Ny, rank):
        r = (C.c_int * 4)()
        self.L.cf_comm_ranges(nmx, Ny, rank, r)
        return tuple(r)

    def profile_enable(self, on=True): self.L.cf_profile_enable(1 if on else 0)

    def profile_read(self, reset=True):
        ms = (C.c_double * 9)()
        calls = (C.c_longlong * 9)()
        self.L.cf_profile_read(ms, calls, 1 if reset else 0)
        return list(ms), list(calls)

    def launch_count(self): return self.L.cf_launch_count()
    def timer_start(self): self.L.cf_timer_start()
    def timer_stop(self): return self.L.cf_timer_stop()


class FlowField:
    """chflow::FlowField of this package (device resident)."""

    def __init__(self, lib, Nx, Ny, Nz, Nd, Lx, Lz, a=-1.0, b=1.0, handle=None):
        self.lib = lib
        self.Nx, self.Ny, self.Nz, self.Nd, self.Lx, self.Lz, self.a, self.b = Nx, Ny, Nz, Nd, Lx, Lz, a, b
        self.Mz = Nz // 2 + 1
        self.h = handle if handle is not None else lib.L.cf_field_create(Nx, Ny, Nz, Nd, Lx, Lz, a, b)

    def __del__(self):
        try:
            if self.h:
                self.lib.L.cf_field_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def shape(self):
        return (self.Nd, self.Ny, self.Nx, 2 * self.Mz)

    def like(self, Nd=None):
        return FlowField(self.lib, self.Nx, self.Ny, self.Nz, self.Nd if Nd is None else Nd, self.Lx, self.Lz, self.a, self.b)

    def set(self, arr, xz=SPECTRAL, y=SPECTRAL, padded=None):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.size == int(np.prod(self.shape))
        self.lib.L.cf_field_set_state(self.h, xz, y)
        self.lib.L.cf_field_upload(self.h, _dp(arr))
        if padded is not None:
            self.lib.L.cf_field_set_padded(self.h, 1 if padded else 0)
        return self

    def get(self):
        out = np.zeros(self.shape)  # de-aliased spectral fields only transfer their retained box
        self.lib.L.cf_field_download(self.h, _dp(out))
        return out

    def copy(self):
        o = self.like()
        self.lib.L.cf_field_copy(o.h, self.h)
        return o

    def state(self):
        a, b = C.c_int(), C.c_int()
        self.lib.L.cf_field_get_state(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_padded(self, p): self.lib.L.cf_field_set_padded(self.h, 1 if p else 0)
    def padded(self): return bool(self.lib.L.cf_field_padded(self.h))
    def make_physical(self): self.lib.L.cf_make_physical(self.h)
    def make_spectral(self): self.lib.L.cf_make_spectral(self.h)
    def make_physical_y(self): self.lib.L.cf_make_physical_y(self.h)
    def make_spectral_y(self): self.lib.L.cf_make_spectral_y(self.h)
    def make_physical_xz(self): self.lib.L.cf_make_physical_xz(self.h)
    def make_spectral_xz(self): self.lib.L.cf_make_spectral_xz(self.h)
    def zero_padded_modes(self): self.lib.L.cf_zero_padded_modes(self.h)
    def symmetry(self, s, sx, sy, sz, ax, az): self.lib.L.cf_field_symmetry(self.h, s, sx, sy, sz, ax, az)
    def l2norm(self): return self.lib.L.cf_l2norm(self.h)

    def l2norm3d(self):
        self.lib.L.cf_l2norm3d.restype = C.c_double
        self.lib.L.cf_l2norm3d.argtypes = [C.c_void_p]
        return self.lib.L.cf_l2norm3d(self.h)
    def l2dist(self, o): return self.lib.L.cf_l2dist(self.h, o.h)
    def l2ip(self, o): return self.lib.L.cf_l2ip(self.h, o.h)
    def cmplx(self, mx, my, mz, i): return complex(self.lib.L.cf_cmplx_get(self.h, mx, my, mz, i, 0), self.lib.L.cf_cmplx_get(self.h, mx, my, mz, i, 1))
    def set_cmplx(self, mx, my, mz, i, v): self.lib.L.cf_cmplx_set(self.h, mx, my, mz, i, v.real, v.imag)
    def allgather(self): self.lib.L.cf_field_allgather(self.h)

    def to_vector(self):
        """field2vector (reference flowfield.cpp:4481-4563)."""
        self.lib.L.cf_field2vector_size.argtypes = [C.c_void_p]
        n = self.lib.L.cf_field2vector_size(self.h)
        x = np.zeros(n)
        self.lib.L.cf_field2vector.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self.lib.L.cf_field2vector(self.h, _dp(x))
        return x

    def from_vector(self, x):
        """vector2field (reference flowfield.cpp:4565-4752)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.lib.L.cf_vector2field.argtypes = [C.POINTER(C.c_double), C.c_void_p]
        self.lib.L.cf_vector2field(_dp(x), self.h)
        return self
    def save(self, filebase): self.lib.L.cf_field_save(self.h, filebase.encode())
    def axpby(self, a, x, b=0.0, z=None): self.lib.L.cf_field_axpby(self.h, a, x.h, b, z.h if z is not None else None)
    def scale(self, s): self.lib.L.cf_field_scale(self.h, s)


def randomfield(lib, Nx, Ny, Nz, Lx, Lz, a=-1.0, b=1.0, seed=1, magn=0.2, smooth=0.4, meanflow=False):
    """The reference's `randomfield` rule (tools/randomfield.cpp:49-64) generated by this package's own FlowField."""
    u = FlowField(lib, Nx, Ny, Nz, 3, Lx, Lz, a, b)
    lib.L.cf_randomfield(u.h, int(seed), float(magn), float(smooth), 1 if meanflow else 0)
    return u


def hookstep_search(u, flags, T, dt, sigma=(1, 1, 1, 1, 0.0, 0.0), epsSearch=1e-13, epsGMRES=1e-3, epsDx=1e-7, delta=0.01, Nnewton=20,
                    Ngmres=120, Nhook=20, Tnormalize=False, verbose=False, xrelative=False, zrelative=False):
    """Newton-Krylov-hookstep search for sigma f^T(u) - u = 0 on the device (host/devicesearch.cpp); u (a FlowField) is the
    initial guess and is overwritten by the result.  Returns a dict with the convergence record."""
    import numpy as np
    sg = np.array(sigma, dtype=np.float64)
    par = np.array([epsSearch, epsGMRES, epsDx, delta, Nnewton, Ngmres, Nhook, 1.0 if Tnormalize else 0.0, 1.0 if verbose else 0.0,
                    1.0 if xrelative else 0.0, 1.0 if zrelative else 0.0])
    out = np.full(6 + Nnewton + 1, -1.0)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    u.lib.L.cf_hookstep_search(u.h, C.byref(flags), float(T), float(dt), dp(sg), dp(par), dp(out), len(out))
    hist = [float(v) for v in out[6:] if v >= 0]
    return dict(converged=bool(out[0]), newton_steps=int(out[1]), fevals=int(out[2]), gmres_iterations=int(out[3]), residual=float(out[4]),
                steps_per_eval=float(out[5]), history=hist, ax=float(sg[4]), az=float(sg[5]))


def poincare_search(u, q, flags, nSteps, maxstrides, ustar=None, estar=None, crosssign=0, Tmin=0.0, epsilon=1e-13):
    """DNSPoincare::advanceToSection in strides of nSteps until the section h(u) = 0 is crossed (host/poincare.cpp).
    h = (u, estar) - (ustar, estar) when estar is given, wallshear - dissipation otherwise.  u, q are advanced in place.
    Returns a dict (found, t, h, sign, strides, hcurrent, ucrossing, pcrossing)."""
    import numpy as np
    out = np.zeros(6)
    uc, pc = u.like(), q.like()
    kind = 1 if estar is not None else 0
    u.lib.L.cf_poincare_search(u.h, q.h, C.byref(flags), kind, ustar.h if kind else None, estar.h if kind else None, int(nSteps),
                               int(maxstrides), int(crosssign), float(Tmin), float(epsilon),
                               out.ctypes.data_as(C.POINTER(C.c_double)), uc.h, pc.h)
    return dict(found=bool(out[0]), t=float(out[1]), h=float(out[2]), sign=int(out[3]), strides=int(out[4]), hcurrent=float(out[5]),
                ucrossing=uc, pcrossing=pc)


class DNS:
    """chflow::DNS of this package (reference dns.cpp:22-163 semantics) owning copies of (u, q)."""

    def __init__(self, u, flags, q=None):
        self.lib, self.geom = u.lib, u
        if q is None:
            q = u.like(Nd=1)
        self.flags = flags
        self.h = self.lib.L.cf_dns_create(u.h, q.h, C.byref(flags))

    def __del__(self):
        try:
            if self.h:
                self.lib.L.cf_dns_free(self.h)
                self.h = None
        except Exception:
            pass

    def advance(self, n): self.lib.L.cf_dns_advance(self.h, n)

    def get(self):
        u, q = self.geom.like(), self.geom.like(Nd=1)
        self.lib.L.cf_dns_get(self.h, u.h, q.h)
        return u, q

    def set(self, u=None, q=None): self.lib.L.cf_dns_set(self.h, u.h if u else None, q.h if q else None)
    def symmetry(self, s, sx, sy, sz, ax, az): self.lib.L.cf_dns_symmetry(self.h, s, sx, sy, sz, ax, az)
    def cfl(self): return self.lib.L.cf_dns_cfl(self.h)
    def reset_dt(self, dt): self.lib.L.cf_dns_reset_dt(self.h, dt)
    def time(self): return self.lib.L.cf_dns_time(self.h)
    def dPdx(self): return self.lib.L.cf_dns_dPdx(self.h)
    def Ubulk(self): return self.lib.L.cf_dns_Ubulk(self.h)


class DeviceVector:
    """Device-resident state vector (cfgpu_vec): dot / norm / axpy on the GPU (nsolver's Krylov algebra)."""

    def __init__(self, lib, n):
        self.lib = lib
        L = lib.L
        L.cf_vec_create.restype = C.c_void_p
        L.cf_vec_create.argtypes = [C.c_long]
        L.cf_vec_size.restype = C.c_long
        for n_ in ("cf_vec_free", "cf_vec_size", "cf_vec_norm"):
            getattr(L, n_).argtypes = [C.c_void_p]
        L.cf_vec_dot.restype = L.cf_vec_norm.restype = C.c_double
        L.cf_vec_dot.argtypes = [C.c_void_p, C.c_void_p]
        L.cf_vec_upload.argtypes = L.cf_vec_download.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.cf_vec_axpy.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.cf_vec_axpby.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double]
        L.cf_vec_scale.argtypes = [C.c_void_p, C.c_double]
        L.cf_field2vector_dev.argtypes = L.cf_vector2field_dev.argtypes = [C.c_void_p, C.c_void_p]
        self.h = L.cf_vec_create(n)

    def __del__(self):
        try:
            if self.h:
                self.lib.L.cf_vec_free(self.h)
                self.h = None
        except Exception:
            pass

    def size(self): return self.lib.L.cf_vec_size(self.h)

    def set(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.size()
        self.lib.L.cf_vec_upload(self.h, _dp(x))
        return self

    def get(self):
        x = np.empty(self.size())
        self.lib.L.cf_vec_download(self.h, _dp(x))
        return x

    def dot(self, o): return self.lib.L.cf_vec_dot(self.h, o.h)
    def norm(self): return self.lib.L.cf_vec_norm(self.h)
    def axpy(self, a, x): self.lib.L.cf_vec_axpy(self.h, a, x.h)
    def axpby(self, a, x, b): self.lib.L.cf_vec_axpby(self.h, a, x.h, b)
    def scale(self, s): self.lib.L.cf_vec_scale(self.h, s)
    def from_field(self, u): self.lib.L.cf_field2vector_dev(u.h, self.h); return self
    def to_field(self, u): self.lib.L.cf_vector2field_dev(self.h, u.h); return u


def nonlinear(u, flags):
    f = u.like()
    u.lib.L.cf_nonlinear(u.h, f.h, C.byref(flags))
    return f


def base_profiles(u, flags):
    U, W = np.zeros(u.Ny), np.zeros(u.Ny)
    u.lib.L.cf_base_profiles(u.h, C.byref(flags), _dp(U), _dp(W))
    return U, W


def laminar_profile(lib, flags, a, b, Ny):
    U = np.zeros(Ny)
    lib.L.cf_laminar_profile(C.byref(flags), a, b, Ny, _dp(U))
    return U
