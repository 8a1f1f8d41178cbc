// Random initial data and small FlowField utilities that work through the host mirror.
//
// randomUprofile / randomVprofile / randomProfile, FlowField::addPerturbation(s) / perturb: the construction rule of
// tools/randomfield.cpp -- Gaussian Chebyshev coefficients with geometric decay drawn from the libc drand48 stream,
// corrected to no-slip walls and zero divergence (reference diffops.cpp:969-1243, flowfield.cpp:2060-2177).  The ORDER of
// the random draws is part of the rule (same seed => same field as the reference, which the bench and the parity tests
// rely on); the per-mode profiles are O(Ny) host work, the closing makePhysical/makeSpectral round trip runs on the GPU.
#include <fstream>
#include <iomanip>

#include "channelflow/diffops.h"
#include "channelflow/flowfield.h"

namespace chflow {

// u(y): N gaussian coefficients mag * decay^n, then T0, T1 adjusted so that u(a) = u(b) = 0
void randomUprofile(ComplexChebyCoeff& u, Real mag, Real decay) {
    const int N = u.length();
    u.setState(Spectral);
    for (int n = 0; n < N; ++n, mag *= decay) u.set(n, mag * randomComplex());
    const Complex ub = u.eval_b(), ua = u.eval_a();
    u.sub(0, (ub + ua) / 2.0);
    u.sub(1, (ub - ua) / 2.0);
}

// remove s0 T0 + .. + s3 T3 so that v = v' = 0 at both walls; the 4x4 system [T_n(-1); T_n(1); T_n'(-1); T_n'(1)] s =
// (a, b, c, d) on [-1,1] has the closed-form inverse used here
static void clamp_both_walls(ComplexChebyCoeff& v) {
    const ComplexChebyCoeff vy = diff(v);
    const Complex a = v.eval_a(), b = v.eval_b(), c = vy.eval_a(), d = vy.eval_b();
    v.sub(0, 0.5 * (a + b) + 0.125 * (c - d));
    v.sub(1, 0.5625 * (b - a) - 0.0625 * (c + d));
    v.sub(2, 0.125 * (d - c));
    v.sub(3, 0.0625 * (a - b + c + d));
}
// v(y): coefficients 4 .. N-3 gaussian with decay, clamped (v = v' = 0 at the walls); done on [-1,1] and mapped back
void randomVprofile(ComplexChebyCoeff& v, Real mag, Real decay) {
    const Real ya = v.a(), yb = v.b();
    v.setBounds(-1, 1);
    v.setState(Spectral);
    const int N = v.length();
    for (int n = 0; n < N; ++n) v.set(n, 0.0);
    for (int n = 4; n < N - 2; ++n, mag *= decay) v.set(n, mag * randomComplex());
    clamp_both_walls(v);
    v.setBounds(ya, yb);
}
// single Chebyshev mode with a random phase
void chebyUprofile(ComplexChebyCoeff& u, int n, Real decay) {
    u.setToZero();
    u.setState(Spectral);
    const Real theta = randomReal(0, 2 * pi);
    u.set(n, (cos(theta) + I * sin(theta)) * std::pow(decay, n));
    const Complex ub = u.eval_b(), ua = u.eval_a();
    u.sub(0, (ub + ua) / 2.0);
    u.sub(1, (ub - ua) / 2.0);
}
void chebyVprofile(ComplexChebyCoeff& v, int n, Real decay) {
    v.setToZero();
    v.setState(Spectral);
    const Real ya = v.a(), yb = v.b();
    v.setBounds(-1, 1);
    const Real theta = randomReal(0, 2 * pi);
    v.set(n, (cos(theta) + I * sin(theta)) * std::pow(decay, n));
    clamp_both_walls(v);
    v.setBounds(ya, yb);
}

// one Fourier mode (kx,kz) of a divergence-free, no-slip field.  The draw order below is the reference's
// (diffops.cpp:1136-1205): mean mode w then u; otherwise v first, then the free component(s).
void randomProfile(ComplexChebyCoeff& u, ComplexChebyCoeff& v, ComplexChebyCoeff& w, int kx, int kz, Real Lx, Real Lz, Real mag,
                   Real decay) {
    u.setState(Spectral);
    v.setState(Spectral);
    w.setState(Spectral);
    if (kx == 0 && kz == 0) {  // mean mode: real u, w, no v
        randomUprofile(w, mag, decay);
        w.im.setToZero();
        randomUprofile(u, mag, decay);
        u.im.setToZero();
        v.setToZero();
        return;
    }
    randomVprofile(v, mag, decay);
    if (kx == 0) {  // w from continuity: i kzz w = -v'
        const ComplexChebyCoeff vy = diff(v);
        randomUprofile(u, mag, decay);
        u.im.setToZero();
        w = vy;
        w *= -Lz / ((2 * pi * kz) * I);
    } else if (kz == 0) {
        const ComplexChebyCoeff vy = diff(v);
        randomUprofile(w, mag, decay);
        w.im.setToZero();
        u = vy;
        u *= -Lx / ((2 * pi * kx) * I);
    } else {
        // general mode: two independent solenoidal pieces, (u0 random, w0 from continuity) and (w1 = 0, u1 from continuity),
        // each with its own v.  (The first v drawn above only advances the random stream, as in the reference.)
        ComplexChebyCoeff v0(v), v1(v);
        randomVprofile(v0, mag, decay);
        randomVprofile(v1, mag, decay);
        const ComplexChebyCoeff v0y = diff(v0), v1y = diff(v1);
        ComplexChebyCoeff u0(v.numModes(), v.a(), v.b(), Spectral), w1(v.numModes(), v.a(), v.b(), Spectral);
        randomUprofile(u0, mag, decay);
        ComplexChebyCoeff w0(u0);
        w0 *= (2 * pi * kx / Lx) * I;
        { ComplexChebyCoeff t(v0y); t += w0; w0 = t; }  // v0' + i kxx u0 (same summation order as the reference)
        w0 *= -Lz / ((2 * pi * kz) * I);
        ComplexChebyCoeff u1(w1);
        u1 *= (2 * pi * kz / Lz) * I;
        { ComplexChebyCoeff t(v1y); t += u1; u1 = t; }
        u1 *= -Lx / ((2 * pi * kx) * I);
        u = u0; v = v0; w = w0;
        u += u1; v += v1; w += w1;
    }
}

void FlowField::addPerturbation(int kx, int kz, Real mag, Real decay) {
    assertState(Spectral, Spectral);
    if (mag == 0.0) return;
    ComplexChebyCoeff u(Ny_, a_, b_, Spectral), v(Ny_, a_, b_, Spectral), w(Ny_, a_, b_, Spectral);
    randomProfile(u, v, w, kx, kz, Lx_, Lz_, mag, decay);
    const Real k = 2 * pi * sqrt(kx * kx / (Lx_ * Lx_) + kz * kz / (Lz_ * Lz_));
    const Real damp = pow(decay, k);
    u *= damp; v *= damp; w *= damp;
    const int m_x = mx(kx), m_z = mz(kz);
    for (int ny = 0; ny < Ny_; ++ny) {
        cmplx(m_x, ny, m_z, 0) += u[ny];
        cmplx(m_x, ny, m_z, 1) += v[ny];
        cmplx(m_x, ny, m_z, 2) += w[ny];
    }
    if (kz == 0 && kx != 0) {
        // conjugate partner on the kz = 0 plane.  REFERENCE QUIRK kept (flowfield.cpp:2092-2103): the partner is added
        // once per vector component, i.e. Nd times; the closing c2r/r2c round trip of addPerturbations symmetrises the plane
        const int m_xm = mx(-kx);
        for (int rep = 0; rep < Nd_; ++rep)
            for (int ny = 0; ny < Ny_; ++ny) {
                cmplx(m_xm, ny, 0, 0) += conj(u[ny]);
                cmplx(m_xm, ny, 0, 1) += conj(v[ny]);
                cmplx(m_xm, ny, 0, 2) += conj(w[ny]);
            }
    }
}
void FlowField::addPerturbation1D(int kx, int kz, Real mag, Real decay) {
    assertState(Spectral, Spectral);
    if (mag == 0.0) return;
    ComplexChebyCoeff u(Ny_, a_, b_, Spectral), vd(Ny_, a_, b_, Spectral), wd(Ny_, a_, b_, Spectral);
    randomProfile(u, vd, wd, kx, kz, Lx_, Lz_, mag, decay);
    u *= pow(decay, 2 * pi * sqrt(kx * kx / (Lx_ * Lx_) + kz * kz / (Lz_ * Lz_)));
    const int m_x = mx(kx), m_z = mz(kz);
    for (int ny = 0; ny < Ny_; ++ny) cmplx(m_x, ny, m_z, 0) += u[ny];
    if (kz == 0 && kx != 0) {
        const int m_xm = mx(-kx);
        for (int rep = 0; rep < Nd_; ++rep)
            for (int ny = 0; ny < Ny_; ++ny) cmplx(m_xm, ny, 0, 0) += conj(u[ny]);
    }
}
void FlowField::addPerturbations(int Kx, int Kz, Real mag, Real decay, bool meanflow) {
    assertState(Spectral, Spectral);
    if (mag == 0.0) return;
    const int Kxmin = Greater(-Kx, padded() ? kxminDealiased() : kxmin());
    const int Kxmax = lesser(Kx, padded() ? kxmaxDealiased() : kxmax() - 1);
    const int Kzmax = lesser(Kz, padded() ? kzmaxDealiased() : kzmax() - 1);
    for (int kx = Kxmin; kx <= Kxmax; ++kx)
        for (int kz = 0; kz <= Kzmax; ++kz) {
            if (!meanflow && kx == 0 && kz == 0) continue;
            const Real norm = std::pow(decay, std::abs(2 * pi * kx / Lx_) + std::abs(2 * pi * kz / Lz_));
            if (Nd_ > 2) addPerturbation(kx, kz, mag * norm, decay);
            else addPerturbation1D(kx, kz, mag * norm, decay);
        }
    makePhysical();
    makeSpectral();
}
void FlowField::addPerturbations(Real, Real, bool) { assertState(Spectral, Spectral); }  // (a no-op in the reference as well)
void FlowField::perturb(Real mag, Real decay, bool meanflow) {
    addPerturbations(padded() ? kxmaxDealiased() : kxmax(), padded() ? kzmaxDealiased() : kzmax(), mag, decay, meanflow);
}

// Spectral interpolation onto this grid (flowfield.cpp:691-792): Fourier modes kxmin <= kx < kxmax, kz <= kzmax common to both
// grids are carried over (the upper bound in kx is exclusive, as in the reference), Chebyshev coefficients are copied when
// this grid has at least as many (zero-extended) and re-sampled through ChebyCoeff::interpolate otherwise; when this grid is
// wider in x the source's kx = kxmax row also feeds the -kxmax row.  A grid-change utility (I/O class): runs on the host mirror.
void FlowField::interpolate(FlowField f) {
    FlowField& g = *this;
    if (f.Lx() != g.Lx() || f.Lz() != g.Lz() || f.a() != g.a() || f.b() != g.b() || f.Nd() != g.Nd())
        cferror("FlowField::interpolate(const FlowField& f) error:\nFlowField doesn't match argument f geometrically.\n");
    const fieldstate gxz = g.xzstate(), gy = g.ystate();
    const int fNy = f.Ny(), gNy = g.Ny();
    f.makeSpectral();
    g.setState(Spectral, Spectral);
    g.setToZero();
    const int kxhi = lesser(f.kxmax(), g.kxmax()), kzhi = lesser(f.kzmax(), g.kzmax()), kxlo = Greater(f.kxmin(), g.kxmin());
    ComplexChebyCoeff fprof(fNy, a_, b_, Spectral), gprof(gNy, a_, b_, Spectral);
    auto carry = [&](int i, int fmx, int fmz, int gmx, int gmz) {
        if (fNy <= gNy) {
            for (int ny = 0; ny < fNy; ++ny) g.cmplx(gmx, ny, gmz, i) = f.cmplx(fmx, ny, fmz, i);
        } else {
            for (int ny = 0; ny < fNy; ++ny) fprof.set(ny, f.cmplx(fmx, ny, fmz, i));
            gprof.interpolate(fprof);
            for (int ny = 0; ny < gNy; ++ny) g.cmplx(gmx, ny, gmz, i) = gprof[ny];
        }
    };
    for (int i = 0; i < Nd_; ++i) {
        for (int kx = kxlo; kx < kxhi; ++kx)
            for (int kz = 0; kz <= kzhi; ++kz) carry(i, f.mx(kx), f.mz(kz), g.mx(kx), g.mz(kz));
        if (g.Nx() > f.Nx())
            for (int kz = f.kzmin(); kz <= f.kzmax(); ++kz) carry(i, f.mx(f.kxmax()), f.mz(kz), g.mx(-f.kxmax()), g.mz(kz));
    }
    g.makeState(gxz, gy);
}

// ------------------------------------------------------------------------------------------------ small utilities
FlowField FlowField::operator[](int i) const {
    assert(i >= 0 && i < Nd_);
    FlowField ui(Nx_, Ny_, Nz_, 1, Lx_, Lz_, a_, b_, cfmpi_, xzstate_, ystate_);
    ui.setComponent(0, *this, i);
    ui.setPadded(padded_);
    return ui;
}
void FlowField::setComponent(int i, const FlowField& src, int j) {
    assert(geomCongruent(src));
    cfgpu_check(cfgpu_field_copy_component(device_mut(), i, src.device(), j), "cfgpu_field_copy_component");
}
Complex FlowField::Dx(int mx_, int n) const {
    const int k = kx(mx_);
    const Complex d(0.0, 2 * pi * k / Lx_ * ((k == kxmax() && (n % 2 == 1)) ? 0 : 1));
    Complex r(1.0, 0.0);
    for (int q = 0; q < n; ++q) r *= d;
    return r;
}
Complex FlowField::Dz(int mz_, int n) const {
    const int k = kz(mz_);
    const Complex d(0.0, 2 * pi * k / Lz_ * ((k == kzmax() && (n % 2 == 1)) ? 0 : 1));
    Complex r(1.0, 0.0);
    for (int q = 0; q < n; ++q) r *= d;
    return r;
}
Vector FlowField::xgridpts() const { Vector p(Nx_); for (int n = 0; n < Nx_; ++n) p[n] = x(n); return p; }
Vector FlowField::ygridpts() const { Vector p(Ny_); for (int n = 0; n < Ny_; ++n) p[n] = y(n); return p; }
Vector FlowField::zgridpts() const { Vector p(Nz_); for (int n = 0; n < Nz_; ++n) p[n] = z(n); return p; }

FlowField& FlowField::operator+=(const Real& a) { cmplx(0, 0, 0, 0) += Complex(a, 0.0); return *this; }
FlowField& FlowField::operator-=(const Real& a) { cmplx(0, 0, 0, 0) -= Complex(a, 0.0); return *this; }
static void add_cprofile(FlowField& u, const ComplexChebyCoeff& U, Real s) {
    std::vector<Real> buf(2 * (size_t)u.Ny(), 0.0);
    for (int n = 0; n < u.Ny() && n < U.length(); ++n) { buf[2 * n] = U.re[n]; buf[2 * n + 1] = U.im[n]; }
    cfgpu_check(cfgpu_field_add_profile(u.device_mut(), 0, 0, 0, buf.data(), s), "cfgpu_field_add_profile");
}
FlowField& FlowField::operator+=(const ComplexChebyCoeff& U) { add_cprofile(*this, U, 1.0); return *this; }
FlowField& FlowField::operator-=(const ComplexChebyCoeff& U) { add_cprofile(*this, U, -1.0); return *this; }

BasisFunc FlowField::profile(int mx_, int mz_) const {
    BasisFunc f(Nd_, Ny_, kx(mx_), kz(mz_), Lx_, Lz_, a_, b_, ystate_);
    for (int i = 0; i < Nd_; ++i) f[i] = profile(mx_, mz_, i);
    return f;
}
bool FlowField::congruent(const BasisFunc& phi) const {
    return Ny_ == phi.Ny() && Lx_ == phi.Lx() && Lz_ == phi.Lz() && a_ == phi.a() && b_ == phi.b() && ystate_ == phi.state();
}
static void add_basisfunc(FlowField& u, const BasisFunc& U, Real s) {
    if (std::abs(U.kx()) > u.kxmax() || U.kz() < 0 || U.kz() > u.kzmax()) return;  // kz < 0 modes are the conjugates of stored ones
    std::vector<Real> buf(2 * (size_t)u.Ny());
    cfgpu_field d = u.device_mut();
    for (int i = 0; i < u.Nd() && i < U.Nd(); ++i) {
        for (int n = 0; n < u.Ny(); ++n) { buf[2 * n] = U[i].re[n]; buf[2 * n + 1] = U[i].im[n]; }
        cfgpu_check(cfgpu_field_add_profile(d, u.mx(U.kx()), u.mz(U.kz()), i, buf.data(), s), "cfgpu_field_add_profile");
    }
}
FlowField& FlowField::operator+=(const BasisFunc& U) { add_basisfunc(*this, U, 1.0); return *this; }
FlowField& FlowField::operator-=(const BasisFunc& U) { add_basisfunc(*this, U, -1.0); return *this; }

Real FlowField::energy(bool normalize) const { return L2Norm2(*this, normalize); }
Real FlowField::energy(int mx_, int mz_, bool normalize) const {
    Real e = 0.0;
    for (int i = 0; i < Nd_; ++i) e += L2Norm2(profile(mx_, mz_, i), normalize);
    if (!normalize) e *= Lx_ * Lz_;
    return e;
}
// new box lengths for the same coefficients: u, w scaled with the box so that the field stays divergence free, then the
// norm is restored (flowfield.cpp:669-688)
void FlowField::rescale(Real Lx, Real Lz) {
    assertState(Spectral, Spectral);
    const Real e = L2Norm(*this);
    if (Nd_ == 3) {
        FlowField c0 = (*this)[0], c2 = (*this)[2];
        c0 *= Lx / Lx_;
        c2 *= Lz / Lz_;
        setComponent(0, c0, 0);
        setComponent(2, c2, 0);
    }
    // same data on the new box: rebuild the device object with the new lengths
    std::vector<Real> tmp((size_t)Nloc());
    cfgpu_check(cfgpu_field_download(device(), tmp.data()), "cfgpu_field_download");
    const bool pad = padded_;
    const int Nx = Nx_, Ny = Ny_, Nz = Nz_, Nd = Nd_;
    const Real a = a_, b = b_;
    resize(Nx, Ny, Nz, Nd, Lx, Lz, a, b, cfmpi_);
    setState(Spectral, Spectral);
    cfgpu_check(cfgpu_field_upload(device_overwrite(), tmp.data(), CFGPU_SPECTRAL, CFGPU_SPECTRAL), "cfgpu_field_upload");
    setPadded(pad);
    const Real e2 = L2Norm(*this);
    if (e2 > 0) (*this) *= e / e2;
}

// ------------------------------------------------------------------------------------------------ ascii output
void FlowField::saveProfile(int mx_, int mz_, const std::string& filebase) const {
    ChebyTransform t(Ny_);
    saveProfile(mx_, mz_, filebase, t);
}
void FlowField::saveProfile(int mx_, int mz_, const std::string& filebase, const ChebyTransform& t) const {
    assert(xzstate_ == Spectral);
    std::ofstream os((filebase + (Nd_ == 3 ? ".bf" : ".asc")).c_str());
    os << std::setprecision(REAL_DIGITS);
    std::vector<ComplexChebyCoeff> f(Nd_);
    for (int i = 0; i < Nd_; ++i) {
        f[i] = profile(mx_, mz_, i);
        if (ystate_ == Spectral) f[i].makePhysical(t);
    }
    for (int ny = 0; ny < Ny_; ++ny) {
        for (int i = 0; i < Nd_; ++i) os << f[i].re[ny] << ' ' << f[i].im[ny] << ' ';
        os << '\n';
    }
}
void FlowField::asciiSave(const std::string& filebase) const {
    std::ofstream os(appendSuffix(filebase, ".asc").c_str());
    os << std::scientific << std::setprecision(REAL_DIGITS);
    const int w = REAL_IOWIDTH;
    os << "% Channelflow FlowField data\n% xzstate == " << xzstate_ << "\n%  ystate == " << ystate_ << '\n';
    os << "% Nx Ny Nz Nd == " << Nx_ << ' ' << Ny_ << ' ' << Nz_ << ' ' << Nd_ << " gridpoints\n";
    os << "% Mx My Mz Nd == " << Mx() << ' ' << My() << ' ' << Mz() << ' ' << Nd_ << " spectral modes\n";
    os << "% Lx Lz == " << std::setw(w) << Lx_ << ' ' << std::setw(w) << Lz_ << '\n';
    os << "% Lx Lz == " << std::setw(w) << Lx_ << ' ' << std::setw(w) << Lz_ << '\n';
    os << "% a  b  == " << std::setw(w) << a_ << ' ' << std::setw(w) << b_ << '\n';
    os << "% loop order:\n%   for (int i=0; i<Nd; ++i)\n%     for(long ny=0; ny<Ny; ++ny) // note: Ny == My\n";
    if (xzstate_ == Physical) {
        os << "%       for (int nx=0; nx<Nx; ++nx)\n%         for (int nz=0; nz<Nz; ++nz)\n%           os << f(nx, ny, nz, i) << newline;\n";
        for (int i = 0; i < Nd_; ++i)
            for (int ny = 0; ny < Ny_; ++ny)
                for (int nx = 0; nx < Nx_; ++nx)
                    for (int nz = 0; nz < Nz_; ++nz) os << std::setw(w) << (*this)(nx, ny, nz, i) << '\n';
    } else {
        os << "%       for (int mx=0; mx<Mx; ++mx)\n%         for (int mz=0; mz<Mz; ++mz)\n"
              "%           os << Re(f.cmplx(mx, ny, mz, i) << ' ' << Im(f.cmplx(mx, ny, mz, i) << newline;\n";
        for (int i = 0; i < Nd_; ++i)
            for (int ny = 0; ny < Ny_; ++ny)
                for (int m = 0; m < Mx(); ++m)
                    for (int mzz = 0; mzz < Mz(); ++mzz) {
                        const Complex c = cmplx(m, ny, mzz, i);
                        os << std::setw(w) << c.real() << ' ' << std::setw(w) << c.imag() << '\n';
                    }
    }
}
// |u_{kx,kz}| summed over y (or at one ny) in a kx-by-kz table (flowfield.cpp:2392-2470)
void FlowField::saveSpectrum(const std::string& filebase, int i, int ny, bool kxorder, bool showpadding) const {
    assert(xzstate_ == Spectral);
    std::ofstream os(appendSuffix(filebase, ".asc").c_str());
    os << std::scientific << std::setprecision(REAL_DIGITS);
    const int Kxmin = (padded_ && !showpadding) ? kxminDealiased() : kxmin(), Kxmax = (padded_ && !showpadding) ? kxmaxDealiased() : kxmax();
    const int Kzmax = (padded_ && !showpadding) ? kzmaxDealiased() : kzmax();
    auto row = [&](int m) {
        for (int kz_ = 0; kz_ <= Kzmax; ++kz_) {
            Real s = 0.0;
            if (ny < 0) { const ComplexChebyCoeff p = profile(m, mz(kz_), i); s = L2Norm(p); }
            else s = std::abs(cmplx(m, ny, mz(kz_), i));
            os << s << ' ';
        }
        os << '\n';
    };
    if (kxorder) for (int k = Kxmin; k <= Kxmax; ++k) row(mx(k));
    else for (int m = 0; m < Mx(); ++m) if (kx(m) >= Kxmin && kx(m) <= Kxmax) row(m);
}
void FlowField::saveSpectrum(const std::string& filebase, bool kxorder, bool showpadding) const {
    assert(xzstate_ == Spectral && ystate_ == Spectral);
    std::ofstream os(appendSuffix(filebase, ".asc").c_str());
    os << std::scientific << std::setprecision(REAL_DIGITS);
    const int Kxmin = (padded_ && !showpadding) ? kxminDealiased() : kxmin(), Kxmax = (padded_ && !showpadding) ? kxmaxDealiased() : kxmax();
    const int Kzmax = (padded_ && !showpadding) ? kzmaxDealiased() : kzmax();
    auto row = [&](int m) {
        for (int kz_ = 0; kz_ <= Kzmax; ++kz_) os << sqrt(energy(m, mz(kz_))) << ' ';
        os << '\n';
    };
    if (kxorder) for (int k = Kxmin; k <= Kxmax; ++k) row(mx(k));
    else for (int m = 0; m < Mx(); ++m) if (kx(m) >= Kxmin && kx(m) <= Kxmax) row(m);
}
// flowfield.cpp:2598-2650: the suffix picks the format; without one the reference writes NetCDF when it was built with the
// library and .ff otherwise -- this build keeps the native .ff as the default
void FlowField::save(const std::string& filebase, std::vector<std::string> component_names) const {
    if (hasSuffix(filebase, ".asc")) asciiSave(filebase);
    else if (hasSuffix(filebase, ".nc")) writeNetCDF(filebase, component_names);
    else if (hasSuffix(filebase, ".h5"))
        cferror("FlowField::save(filename) error : can't save to HDF5 file because HDF5 libraries are not installed. filename == " + filebase);
    else binarySave(filebase);
}

}  // namespace chflow
