"""TEST INFRASTRUCTURE -- NumPy restatement of Channelflow's DNS time-step path (second oracle tier).

The primary oracle is the unmodified reference compiled in oracle/_ref (oracle/refcf.py).  This file restates the same
algorithm independently in NumPy, function by function with the reference file:line each follows, so that the path
is pinned twice: tests/test_oracle.py checks this restatement against the compiled reference (and through it against
the reference's golden vectors), and the GPU/emulation tests can be checked against either.  Small grids only
(pure-Python loops over modes).  Nothing outside tests/ may import it.

Layout conventions are the reference's serial ones: complex spectral arrays c[i, my, mx, mz] with mx in FFT order,
mz = 0..Nz/2 (flowfield.h:370-402).
"""
import numpy as np


# ----------------------------------------------------------------------------------------------- Chebyshev calculus
def cheb_to_physical(c):
    """ChebyCoeff::makePhysical / FlowField::makePhysical_y (chebyshev.cpp:1256-1264, flowfield.cpp:1939-1987):
    u(y_j) = sum_n c_n cos(pi j n/(N-1)), along axis 0."""
    N = c.shape[0]
    j = np.arange(N)
    return np.tensordot(np.cos(np.pi * np.outer(j, j) / (N - 1)), c, axes=(1, 0))


def cheb_to_spectral(u):
    """FlowField::makeSpectral_y (flowfield.cpp:1888-1937): REDFT00 then 1/(N-1), halves at both ends; axis 0."""
    N = u.shape[0]
    Nb = N - 1
    j = np.arange(N)
    g = np.full(N, 2.0); g[0] = g[Nb] = 1.0
    X = np.tensordot(np.cos(np.pi * np.outer(j, j) / Nb) * g[None, :], u, axes=(1, 0))
    w = np.full(N, 1.0 / Nb); w[0] = w[Nb] = 0.5 / Nb
    return X * w.reshape((N,) + (1,) * (u.ndim - 1))


def cheb_diff(u, a=-1.0, b=1.0):
    """diff (chebyshev.cpp:672-697): d[N-1]=0, d[N-2]=(4/L)(N-1)u[N-1], d[n]=d[n+2]+(4/L)(n+1)u[n+1], d[0]*=1/2; axis 0."""
    N = u.shape[0]
    d = np.zeros_like(u)
    s = 4.0 / (b - a)
    if N >= 2:
        d[N - 2] = s * (N - 1) * u[N - 1]
    for n in range(N - 3, -1, -1):
        d[n] = d[n + 2] + s * (n + 1) * u[n + 1]
    d[0] = d[0] * 0.5
    return d


def eval_b(u):
    return u.sum(axis=0)                      # chebyshev.cpp:418-430: T_n(1) = 1


def eval_a(u):
    sg = np.where(np.arange(u.shape[0]) % 2 == 0, 1.0, -1.0)
    return np.tensordot(sg, u, axes=(0, 0))   # chebyshev.cpp:405-416: T_n(-1) = (-1)^n


def cheb_mean(u):
    """ChebyCoeff::mean (chebyshev.cpp:505-512)."""
    m = u[0].copy() if isinstance(u[0], np.ndarray) else u[0]
    for n in range(2, u.shape[0], 2):
        m = m - u[n] / (n * n - 1)
    return m


# ----------------------------------------------------------------------------------------------- Helmholtz / tau
def _c(n, Nb):
    return 2 if (n == 0 or n == Nb) else 1


def _beta(n, Nb):
    return 0 if n > Nb - 2 else 1


class Helmholtz:
    """HelmholtzSolver (helmholtz.cpp:18-95): nu u'' - lambda u = f, u(a)=ua, u(b)=ub, C&H eq. 5.1.24, one bordered
    tridiagonal system per parity, UL-factored without pivoting (bandedtridiag.cpp:212-277)."""

    def __init__(self, N, a, b, lam, nu=1.0):
        self.N, self.Nb, self.a, self.b, self.lam, self.nu = N, N - 1, a, b, lam, nu
        nus = nu / ((b - a) / 2) ** 2
        Nb = self.Nb
        self.par = []
        for p in (0, 1):
            M = Nb // 2 + 1 if p == 0 else Nb // 2
            lo, dg, up, band = np.zeros(M), np.zeros(M), np.zeros(M), np.ones(M)
            Blo, Bdg, Bup = np.zeros(M), np.zeros(M), np.zeros(M)
            for i in range(1, M):
                n = 2 * i + p
                lo[i] = -(_c(n - 2, Nb) * lam) / (4 * n * (n - 1))
                dg[i] = nus + (_beta(n, Nb) * lam) / (2 * (n * n - 1))
                if _beta(n + 2, Nb):
                    up[i] = -lam / (4 * n * (n + 1))
                    Bup[i] = 1.0 / (4 * n * (n + 1))
                Blo[i] = _c(n - 2, Nb) / (4 * n * (n - 1))
                Bdg[i] = -_beta(n, Nb) / (2 * (n * n - 1))
            # ULdecomp (bandedtridiag.cpp:212-229); band(0) is diag(0)
            dg[0] = band[0]
            for k in range(M - 1, 1, -1):
                up[k - 1] /= dg[k]
                dg[k - 1] -= lo[k] * up[k - 1]
                band[k] /= dg[k]
                band[k - 1] -= lo[k] * band[k]
            if M > 1:
                band[1] /= dg[1]
                band[0] -= lo[1] * band[1]
            dg[0] = band[0]
            self.par.append((M, lo, dg, up, band, Blo, Bdg, Bup))

    def solve(self, f, ua=0.0, ub=0.0):
        """HelmholtzSolver::solve (helmholtz.cpp:79-95) for a real or complex profile f[N]."""
        u = np.zeros(self.N, dtype=f.dtype)
        for p, (M, lo, dg, up, band, Blo, Bdg, Bup) in enumerate(self.par):
            fp = f[p::2]
            g = np.zeros(M, dtype=f.dtype)
            for i in range(1, M):            # multiplyStrided (bandedtridiag.cpp:315-333)
                g[i] = Blo[i] * fp[i - 1] + Bdg[i] * fp[i] + (Bup[i] * fp[i + 1] if i + 1 < M else 0.0)
            g[0] = (ub + ua) / 2 if p == 0 else (ub - ua) / 2
            for i in range(M - 2, 0, -1):    # ULsolveStrided (bandedtridiag.cpp:232-277)
                g[i] -= up[i] * g[i + 1]
            for j in range(1, M):
                g[0] -= band[j] * g[j]
            g[0] /= dg[0]
            for i in range(1, M):
                g[i] = (g[i] - lo[i] * g[i - 1]) / dg[i]
            u[p::2] = g
        return u

    def solve_mean(self, f, umean, ua=0.0, ub=0.0):
        """Mean-constrained solve (helmholtz.cpp:158-213); returns (u, mu)."""
        uf = self.solve(f, ua, ub)
        c = np.zeros(self.N); c[0] = self.nu
        uc = self.solve(c, 0.0, 0.0)
        mu = self.nu * (umean - cheb_mean(uf)) / cheb_mean(uc)
        g = f.copy(); g[0] += mu
        return self.solve(g, ua, ub), mu


def _n_func(k, Nb):
    return Nb - 1 if k == 0 else (0 if k == Nb else (2 * (Nb - 1) if k % 2 == 0 else 2 * Nb))


class TauSolver:
    """TauSolver (tausolver.cpp:81-251, 347-402): Kleiser-Schumann influence-matrix solve of one Fourier mode."""

    def __init__(self, kx, kz, Lx, Lz, a, b, lam, nu, N, taucorr=True):
        self.kx, self.kz, self.N, self.Nb, self.a, self.b, self.lam, self.nu, self.taucorr = kx, kz, N, N - 1, a, b, lam, nu, taucorr
        self.kxx, self.kzz = 2 * np.pi * kx / Lx, 2 * np.pi * kz / Lz
        kappa2 = 4 * np.pi ** 2 * ((kx / Lx) ** 2 + (kz / Lz) ** 2)
        self.HP, self.HV = Helmholtz(N, a, b, kappa2), Helmholtz(N, a, b, lam, nu)
        z = np.zeros(N)
        d = lambda x: cheb_diff(x, a, b)  # noqa: E731
        self.Pp = self.HP.solve(z, 0.0, 1.0); self.vp = self.HV.solve(d(self.Pp))
        self.Pm = self.HP.solve(z, 1.0, 0.0); self.vm = self.HV.solve(d(self.Pm))
        A, B = eval_b(d(self.vp)), eval_b(d(self.vm))
        C, D = eval_a(d(self.vp)), eval_a(d(self.vm))
        with np.errstate(all="ignore"):
            disc = A * D - B * C
            self.i00, self.i01, self.i10, self.i11 = D / disc, -B / disc, -C / disc, A / disc
        rhs = np.array([2.0 / (b - a) * _n_func(i, self.Nb) for i in range(N)])
        self.P0 = self.HP.solve(rhs)
        dP0 = d(self.P0)
        self.v0 = self.HV.solve(dP0)
        if kx != 0 or kz != 0:
            self.P0, self.v0 = self._influence(self.P0, self.v0)
        v0yy = d(d(self.v0))
        Nb = self.Nb
        self.s0Nb1 = lam * self.v0[Nb - 1] + dP0[Nb - 1] - nu * v0yy[Nb - 1]
        self.s0Nb = lam * self.v0[Nb] + dP0[Nb] - nu * v0yy[Nb]

    def _influence(self, P, v):
        """influenceCorrection (tausolver.cpp:178-191)."""
        t = cheb_diff(v, self.a, self.b)
        vb, va = eval_b(t), eval_a(t)
        dp = -self.i00 * vb - self.i01 * va
        dm = -self.i10 * vb - self.i11 * va
        return P + (dp * self.Pp + dm * self.Pm), v + (dp * self.vp + dm * self.vm)

    def _P_and_v(self, r, Ry):
        """solve_P_and_v (tausolver.cpp:193-251) for one real profile pair."""
        a, b, Nb = self.a, self.b, self.Nb
        P = self.HP.solve(r)
        if self.kx == 0 and self.kz == 0:
            return P, np.zeros(self.N)
        v = self.HV.solve(cheb_diff(P, a, b) - Ry)
        P, v = self._influence(P, v)
        if not self.taucorr:
            return P, v
        vyy = cheb_diff(cheb_diff(v, a, b), a, b)
        Py = cheb_diff(P, a, b)
        s1Nb = self.lam * v[Nb] - self.nu * vyy[Nb] - Ry[Nb] + Py[Nb]
        s1Nb1 = self.lam * v[Nb - 1] - self.nu * vyy[Nb - 1] - Ry[Nb - 1] + Py[Nb - 1]
        sNb, sNb1 = s1Nb / (1.0 - self.s0Nb), s1Nb1 / (1.0 - self.s0Nb1)
        ev = np.arange(self.N) % 2 == 0
        return P + np.where(ev, sNb1, sNb) * self.P0, v + np.where(ev, sNb, sNb1) * self.v0

    def solve(self, Rx, Ry, Rz):
        """TauSolver::solve (tausolver.cpp:347-402): complex profiles in, (u, v, w, P) out."""
        a, b = self.a, self.b
        Ryy = cheb_diff(Ry, a, b)
        r_re = Ryy.real - self.kxx * Rx.imag - self.kzz * Rz.imag
        r_im = Ryy.imag + self.kxx * Rx.real + self.kzz * Rz.real
        Pr, vr = self._P_and_v(r_re, Ry.real)
        Pi, vi = self._P_and_v(r_im, Ry.imag)
        P, v = Pr + 1j * Pi, vr + 1j * vi
        fu, fw = 1j * self.kxx * P - Rx, 1j * self.kzz * P - Rz
        u = self.HV.solve(fu.real) + 1j * self.HV.solve(fu.imag)
        w = self.HV.solve(fw.real) + 1j * self.HV.solve(fw.imag)
        return u, v, w, P


# ----------------------------------------------------------------------------------------------- transforms, NL
def kx_of(mx, Nx):
    return mx if mx <= Nx // 2 else mx - Nx


def to_physical(c, Nz):
    """makePhysical = y then xz (flowfield.cpp:1989-1997): c[..., my, mx, mz] -> real [..., ny, nx, nz]."""
    p = cheb_to_physical(np.moveaxis(c, -3, 0))
    p = np.moveaxis(p, 0, -3)
    Nx = c.shape[-2]
    return np.fft.irfft2(p, s=(Nx, Nz), axes=(-2, -1)) * (Nx * Nz)   # unnormalised c2r (flowfield.cpp:1870-1886)


def to_spectral(r):
    """makeSpectral = xz (r2c, then 1/(Nx Nz), flowfield.cpp:1850-1868) then y."""
    Nx, Nz = r.shape[-2], r.shape[-1]
    c = np.fft.rfft2(r, axes=(-2, -1)) / (Nx * Nz)
    s = cheb_to_spectral(np.moveaxis(c, -3, 0))
    return np.moveaxis(s, 0, -3)


def curl(c, Lx, Lz, a, b):
    """curl (diffops.cpp:2229-2334) of a spectral 3-vector c[3, my, mx, mz]."""
    _, Ny, Nx, Mz = c.shape
    Nz = 2 * (Mz - 1)
    kx = np.array([kx_of(m, Nx) for m in range(Nx)], dtype=float)
    kz = np.arange(Mz, dtype=float)
    kx[kx == Nx // 2] = 0.0   # zero_last_mode (flowfield.h:593)
    kz[kz == Nz // 2] = 0.0
    Dx = (2j * np.pi * kx / Lx)[None, :, None]
    Dz = (2j * np.pi * kz / Lz)[None, None, :]
    dy = lambda f: cheb_diff(f, a, b)  # noqa: E731
    u, v, w = c
    return np.stack([dy(w) - Dz * v, Dz * u - Dx * w, Dx * v - dy(u)])


def zero_padded(c):
    """zeroPaddedModes (flowfield.cpp:2235-2255)."""
    Nx, Mz = c.shape[-2], c.shape[-1]
    Nz = 2 * (Mz - 1)
    Kx, Kz = Nx // 3 - 1, Nz // 3 - 1
    out = c.copy()
    for mx in range(Nx):
        if abs(kx_of(mx, Nx)) > Kx:
            out[..., mx, :] = 0
    out[..., Kz + 1:] = 0
    return out


def rotational_nl(c, Ubase, Wbase, Lx, Lz, a, b, Vsuck=0.0, dealias=True):
    """NSE::nonlinear, rotational form (nse.cpp:12-91, 383-391; diffops.cpp:2852-2881): f = (curl u_tot) x u_tot."""
    Nz = 2 * (c.shape[-1] - 1)
    t = c.copy()
    t[0, :, 0, 0] += Ubase
    t[2, :, 0, 0] += Wbase
    t[1, 0, 0, 0] -= Vsuck
    om = curl(t, Lx, Lz, a, b)
    up, op = to_physical(t, Nz), to_physical(om, Nz)
    f = to_spectral(np.cross(op, up, axis=0))
    return zero_padded(f) if dealias else f


def nse_solve(rhs, lam_t, nu, Lx, Lz, a, b, Ubase=None, Wbase=None, dPdx=0.0, dPdz=0.0, dealias=True, taucorr=True):
    """NSE::solve (nse.cpp:479-575), pressure-gradient constraint: per retained mode solve
    nu u'' - lambda u - grad q = -R, div u = 0.  rhs[3, my, mx, mz] complex; returns (u[3,...], q[...])."""
    _, Ny, Nx, Mz = rhs.shape
    Nz = 2 * (Mz - 1)
    Kx = Nx // 3 - 1 if dealias else Nx // 2 - 1
    Kz = Nz // 3 - 1 if dealias else Nz // 2 - 1
    u, q = np.zeros_like(rhs), np.zeros(rhs.shape[1:], dtype=complex)
    for mx in range(Nx):
        kx = kx_of(mx, Nx)
        if abs(kx) > Kx:
            continue
        for kz in range(Kz + 1):
            lam = lam_t + 4 * np.pi ** 2 * nu * ((kx / Lx) ** 2 + (kz / Lz) ** 2)   # nse.cpp:688-701
            R = [rhs[i, :, mx, kz].copy() for i in range(3)]
            if kx == 0 and kz == 0:
                R = [r.real + 0j for r in R]
                if Ubase is not None:
                    R[0] += nu * cheb_diff(cheb_diff(Ubase, a, b), a, b)
                if Wbase is not None:
                    R[2] += nu * cheb_diff(cheb_diff(Wbase, a, b), a, b)
                R[0][0] -= dPdx
                R[2][0] -= dPdz
            ts = TauSolver(kx, kz, Lx, Lz, a, b, lam, nu, Ny, taucorr)
            uu, vv, ww, pp = ts.solve(*R)
            if kx == 0 and kz == 0:
                uu, vv, ww, pp = uu.real + 0j, vv.real + 0j, ww.real + 0j, pp.real + 0j
            u[0, :, mx, kz], u[1, :, mx, kz], u[2, :, mx, kz], q[:, mx, kz] = uu, vv, ww, pp
    return u, q


def sbdf1_step(c, dt, nu, Ubase, Wbase, Lx, Lz, a, b, dPdx=0.0):
    """One SBDF1 (= CNFE1) step of MultistepDNS::advance (dnsalgo.cpp:195-260): rhs = u/dt - f(u), lambda_t = 1/dt."""
    f = rotational_nl(c, Ubase, Wbase, Lx, Lz, a, b)
    return nse_solve(c / dt - f, 1.0 / dt, nu, Lx, Lz, a, b, Ubase, Wbase, dPdx=dPdx)
