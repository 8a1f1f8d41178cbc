// Host Chebyshev expansions (see channelflow/chebyshev.h).  Formulas follow the reference's chebyshev.cpp: transforms
// :262-302 (REDFT00 conventions), eval_a/eval_b :405-430, slopes :432-456, Clenshaw eval :458-503, mean :505-512, file
// forms :516-561, integrate :637-664, diff :672-697, norms :758-947.
#include "channelflow/chebyshev.h"

#include <fstream>
#include <iomanip>

namespace chflow {

// ------------------------------------------------------------------------------------------------ small functions
Real chebyIP(int m, int n) {
    if ((m + n) % 2 == 1) return 0.0;
    const Real e = 1.0, p = m, q = n;
    return (e - p * p - q * q) / ((e + p - q) * (e - p + q) * (e + p + q) * (e - p - q));
}
Real legendre(int n, Real x) {  // three-term recurrence
    Real p = 1.0, q = 0.0;
    for (int m = 0; m < n; ++m) {
        const Real r = q;
        q = p;
        p = ((2 * m + 1) * x * q - m * r) / (m + 1);
    }
    return p;
}
Real chebyshev(int n, Real x) { return std::cos(n * std::acos(x)); }

void legendre(int n, ChebyCoeff& u, ChebyTransform& trans, bool normalize) {
    const int N = u.N();
    u.setState(Physical);
    for (int q = 0; q < N; ++q) u[q] = legendre(n, std::cos(q * pi / (N - 1)));
    u.makeSpectral(trans);
    if (normalize) u *= std::sqrt((double)(2 * n + 1));
}

// nodes and weights of N-point Gauss-Legendre quadrature on [a,b] (Newton iteration on P_N from the Chebyshev guess)
void gaussLegendreQuadrature(int N, Real a, Real b, Vector& x, Vector& w) {
    x = Vector(N);
    w = Vector(N);
    const Real mid = 0.5 * (b + a), rad = 0.5 * (b - a);
    for (int m = 0; m < (N + 1) / 2; ++m) {
        Real z = std::cos(pi * (m + 0.75) / (N + 0.5)), dp = 1.0;
        for (int it = 0; it < 100; ++it) {
            Real p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < N; ++j) {
                const Real p3 = p2;
                p2 = p1;
                p1 = ((2 * j + 1) * z * p2 - j * p3) / (j + 1);
            }
            dp = N * (z * p1 - p2) / (z * z - 1.0);
            const Real z1 = z;
            z = z1 - p1 / dp;
            if (std::fabs(z - z1) < 1e-17) break;
        }
        x[m] = mid - rad * z;
        x[N - 1 - m] = mid + rad * z;
        w[m] = w[N - 1 - m] = 2.0 * rad / ((1.0 - z * z) * dp * dp);
    }
}

Vector chebypoints(int N, Real a, Real b) {
    Vector y(N);
    const Real piN = pi / (N - 1), rad = (b - a) / 2, mid = (b + a) / 2;
    for (int j = 0; j < N; ++j) y[j] = mid + rad * std::cos(piN * j);
    return y;
}

// ------------------------------------------------------------------------------------------------ transform
static long double cospi_frac(long num, long den) {  // cos(pi num/den) with exact argument reduction
    const long double PIl = 3.141592653589793238462643383279502884L;
    num %= 2 * den;
    if (num < 0) num += 2 * den;
    if (num > den) num = 2 * den - num;
    if (2 * num == den) return 0.0L;
    if (2 * num > den) return -cosl(PIl * (long double)(den - num) / (long double)den);
    return cosl(PIl * (long double)num / (long double)den);
}
ChebyTransform::ChebyTransform(int N, uint flags) : N_(N), flags_(flags) {
    if (N >= 2) {
        cos_ = std::make_shared<std::vector<Real>>(2 * (size_t)(N - 1));
        for (int k = 0; k < 2 * (N - 1); ++k) (*cos_)[k] = (Real)cospi_frac(k, N - 1);
    }
}
void ChebyTransform::inverse(std::vector<Real>& x) const {
    const int N = N_, Nb = N - 1, P = 2 * Nb;
    if (N < 2) return;
    std::vector<Real> u(N);
    for (int j = 0; j < N; ++j) {
        long double s = 0.0L;
        for (int n = 0; n < N; ++n) s += (long double)x[n] * (long double)(*cos_)[(size_t)((long)j * n % P)];
        u[j] = (Real)s;
    }
    x = u;
}
void ChebyTransform::forward(std::vector<Real>& x) const {
    const int N = N_, Nb = N - 1, P = 2 * Nb;
    if (N < 2) return;
    std::vector<Real> c(N);
    for (int n = 0; n < N; ++n) {
        long double s = 0.0L;
        for (int j = 0; j < N; ++j)
            s += (long double)x[j] * (long double)(*cos_)[(size_t)((long)j * n % P)] * ((j == 0 || j == Nb) ? 1.0L : 2.0L);
        c[n] = (Real)(s * ((n == 0 || n == Nb) ? 0.5L : 1.0L) / (long double)Nb);
    }
    x = c;
}

// ------------------------------------------------------------------------------------------------ ChebyCoeff
ChebyCoeff::ChebyCoeff() : Vector(), a_(0), b_(0), state_(Spectral) {}
ChebyCoeff::ChebyCoeff(int N, Real a, Real b, fieldstate s) : Vector(N), a_(a), b_(b), state_(s) {}
ChebyCoeff::ChebyCoeff(const Vector& v, Real a, Real b, fieldstate s) : Vector(v), a_(a), b_(b), state_(s) {}
ChebyCoeff::ChebyCoeff(int N, const ChebyCoeff& g) : Vector(N), a_(g.a_), b_(g.b_), state_(g.state_) {
    const int M = lesser(N, g.N());
    for (int i = 0; i < M; ++i) data_[i] = g.data_[i];
}
ChebyCoeff::~ChebyCoeff() {}

// ascii form: "% N a b state" then one value per line
ChebyCoeff::ChebyCoeff(const std::string& filebase) : Vector(0), a_(0), b_(0), state_(Spectral) {
    std::ifstream is;
    const std::string filename = ifstreamOpen(is, filebase, ".asc");
    if (!is.good()) cferror("ChebyCoeff::ChebyCoeff(filebase) : can't open file " + filename);
    char c = 0;
    int N = 0;
    is >> c;
    if (c != '%') cferror("ChebyCoeff::ChebyCoeff(filebase): bad header in file " + filename);
    is >> N >> a_ >> b_ >> state_;
    data_.assign(N, 0.0);
    for (auto& x : data_) is >> x;
    makeSpectral();
}
void ChebyCoeff::save(const std::string& filebase, fieldstate savestate) const {
    const fieldstate orig = state_;
    ChebyCoeff& self = const_cast<ChebyCoeff&>(*this);
    self.makeState(savestate);
    std::ofstream os(appendSuffix(filebase, ".asc").c_str());
    os << std::scientific << std::setprecision(REAL_DIGITS);
    os << "% " << data_.size() << ' ' << a_ << ' ' << b_ << ' ' << state_ << '\n';
    for (Real x : data_) os << std::setw(REAL_IOWIDTH) << x << '\n';
    os.close();
    self.makeState(orig);
}
void ChebyCoeff::binaryDump(std::ostream& os) const {
    write(os, (int)data_.size());
    write(os, a_);
    write(os, b_);
    write(os, state_);
    for (Real x : data_) write(os, x);
}
void ChebyCoeff::binaryLoad(std::istream& is) {
    if (!is.good()) cferror("ChebyCoeff::binaryLoad(istream) : input error");
    int N = 0;
    read(is, N);
    read(is, a_);
    read(is, b_);
    read(is, state_);
    data_.assign(N, 0.0);
    for (auto& x : data_) {
        if (!is.good()) cferror("ChebyCoeff::binaryLoad(istream) : input error");
        read(is, x);
    }
}
void ChebyCoeff::reconfig(const ChebyCoeff& f) {
    data_.assign(f.data_.size(), 0.0);
    a_ = f.a_; b_ = f.b_; state_ = f.state_;
}
void ChebyCoeff::randomize(Real magn, Real decay, BC aBC, BC bBC) {
    const fieldstate start = state_;
    state_ = Spectral;
    const size_t N = data_.size();
    Real m = magn;
    for (size_t n = 0; n < N; ++n, m *= decay) data_[n] = m * randomReal(-1, 1);
    for (int pass = 0; pass < 2; ++pass) {  // a second pass polishes the boundary values
        if (N == 1 && (aBC == Diri || bBC == Diri)) data_[0] = 0.0;
        else if (N >= 2 && aBC == Diri && bBC == Diri) {
            data_[1] -= 0.5 * (eval_b() - eval_a());
            data_[0] -= 0.5 * (eval_b() + eval_a());
        } else if (N >= 2 && aBC == Diri) data_[0] -= eval_a();
        else if (N >= 2 && bBC == Diri) data_[0] -= eval_b();
    }
    makeState(start);
}
void ChebyCoeff::setBounds(Real a, Real b) { a_ = a; b_ = b; }
void ChebyCoeff::setState(fieldstate s) { state_ = s; }
void ChebyCoeff::setToZero() { std::fill(data_.begin(), data_.end(), 0.0); }
void ChebyCoeff::fill(const ChebyCoeff& g) {
    const int M = lesser(length(), g.length());
    for (int i = 0; i < M; ++i) data_[i] = g.data_[i];
    for (int i = M; i < length(); ++i) data_[i] = 0.0;
}
void ChebyCoeff::interpolate(const ChebyCoeff& g) {
    state_ = Physical;
    const Real piN = pi / (data_.size() - 1), rad = (b_ - a_) / 2, mid = (b_ + a_) / 2;
    for (size_t n = 0; n < data_.size(); ++n) data_[n] = g.eval(mid + rad * std::cos(n * piN));
    makeSpectral();
}
void ChebyCoeff::reflect(const ChebyCoeff& g, parity p) {
    state_ = Physical;
    const int N = (int)data_.size();
    const Real piN = pi / (N - 1), rad = (b_ - a_) / 2, mid = (b_ + a_) / 2;
    const int sign = p == Odd ? -1 : 1;
    for (int n = 0; n < N / 2; ++n) {
        const Real v = g.eval(mid + rad * std::cos(n * piN));
        data_[n] = v;
        data_[N - 1 - n] = sign * v;
    }
    makeSpectral();
    for (int n = 2 * N / 3; n < N; ++n) data_[n] = 0.0;
    makePhysical();
}

Real ChebyCoeff::eval_b() const {
    if (data_.empty()) return 0;
    if (state_ == Physical) return data_[0];
    Real s = 0.0;
    for (int n = length() - 1; n >= 0; --n) s += data_[n];
    return s;
}
Real ChebyCoeff::eval_a() const {
    if (data_.empty()) return 0;
    if (state_ == Physical) return data_[length() - 1];
    Real s = 0.0;
    for (int n = length() - 1; n >= 0; --n) s += data_[n] * ((n % 2 == 0) ? 1 : -1);
    return s;
}
Real ChebyCoeff::slope_a() const {  // T_n'(-1) = (-1)^(n+1) n^2
    const int N = length();
    Real s = 0.0;
    for (int n = 0; n + 1 < N; n += 2) s += -(Real)n * n * data_[n] + (Real)(n + 1) * (n + 1) * data_[n + 1];
    if (N % 2 == 1) s -= (Real)(N - 1) * (N - 1) * data_[N - 1];
    return 2 * s / (b_ - a_);
}
Real ChebyCoeff::slope_b() const {  // T_n'(1) = n^2
    Real s = 0.0;
    for (int n = length() - 1; n >= 0; --n) s += (Real)n * n * data_[n];
    return 2 * s / (b_ - a_);
}
static Real clenshaw(const std::vector<Real>& c, Real a, Real b, Real x) {
    const int N = (int)c.size();
    if (N == 0) return 0;
    const Real y = (2 * x - a - b) / (b - a), y2 = 2 * y;
    Real d = 0.0, dd = 0.0;
    for (int j = N - 1; j > 0; --j) {
        const Real sv = d;
        d = y2 * d - dd + c[j];
        dd = sv;
    }
    return y * d - dd + c[0];
}
Real ChebyCoeff::eval(Real x) const { return clenshaw(data_, a_, b_, x); }
void ChebyCoeff::eval(const Vector& x, ChebyCoeff& g) const {
    const int N = x.length();
    if (g.length() != N) g.resize(N);
    g.setBounds(a_, b_);
    g.setState(Physical);
    for (int i = 0; i < N; ++i) g[i] = clenshaw(data_, a_, b_, x[i]);
}
ChebyCoeff ChebyCoeff::eval(const Vector& x) const {
    ChebyCoeff g((int)data_.size(), a_, b_, Physical);
    eval(x, g);
    return g;
}
Real ChebyCoeff::mean() const {
    if (data_.empty()) return 0.0;
    Real s = data_[0];
    for (size_t n = 2; n < data_.size(); n += 2) s -= data_[n] / (Real)(n * n - 1);
    return s;
}

ChebyCoeff& ChebyCoeff::operator*=(Real c) { for (auto& x : data_) x *= c; return *this; }
ChebyCoeff& ChebyCoeff::operator+=(const ChebyCoeff& g) { assert(congruent(g)); for (size_t i = 0; i < data_.size(); ++i) data_[i] += g.data_[i]; return *this; }
ChebyCoeff& ChebyCoeff::operator-=(const ChebyCoeff& g) { assert(congruent(g)); for (size_t i = 0; i < data_.size(); ++i) data_[i] -= g.data_[i]; return *this; }
ChebyCoeff& ChebyCoeff::operator*=(const ChebyCoeff& g) {
    assert(g.state_ == Physical && state_ == Physical);
    for (size_t i = 0; i < data_.size(); ++i) data_[i] *= g.data_[i];
    return *this;
}

void ChebyCoeff::chebyfft(const ChebyTransform& t) {
    assert(t.N() == N());
    if (N() >= 2) t.forward(data_);
    state_ = Spectral;
}
void ChebyCoeff::ichebyfft(const ChebyTransform& t) {
    assert(t.N() == N());
    if (N() >= 2) t.inverse(data_);
    state_ = Physical;
}
void ChebyCoeff::makeSpectral(const ChebyTransform& t) { if (state_ == Physical) chebyfft(t); }
void ChebyCoeff::makePhysical(const ChebyTransform& t) { if (state_ == Spectral) ichebyfft(t); }
void ChebyCoeff::makeState(fieldstate s, const ChebyTransform& t) { if (s == Physical) makePhysical(t); else makeSpectral(t); }
void ChebyCoeff::chebyfft() { ChebyTransform t(N()); chebyfft(t); }
void ChebyCoeff::ichebyfft() { ChebyTransform t(N()); ichebyfft(t); }
void ChebyCoeff::makeSpectral() { if (state_ == Physical) chebyfft(); }
void ChebyCoeff::makePhysical() { if (state_ == Spectral) ichebyfft(); }
void ChebyCoeff::makeState(fieldstate s) { if (s == Physical) makePhysical(); else makeSpectral(); }

bool ChebyCoeff::congruent(const ChebyCoeff& g) const { return g.data_.size() == data_.size() && g.a_ == a_ && g.b_ == b_ && g.state_ == state_; }
void swap(ChebyCoeff& f, ChebyCoeff& g) {
    f.data_.swap(g.data_);
    std::swap(f.a_, g.a_); std::swap(f.b_, g.b_); std::swap(f.state_, g.state_);
}

ChebyCoeff operator*(Real c, const ChebyCoeff& g) { ChebyCoeff r(g); r *= c; return r; }
ChebyCoeff operator+(const ChebyCoeff& f, const ChebyCoeff& g) { ChebyCoeff r(f); r += g; return r; }
ChebyCoeff operator-(const ChebyCoeff& f, const ChebyCoeff& g) { ChebyCoeff r(f); r -= g; return r; }
bool operator==(const ChebyCoeff& f, const ChebyCoeff& g) {
    if (!f.congruent(g)) return false;
    for (int i = 0; i < f.N(); ++i)
        if (f[i] != g[i]) return false;
    return true;
}
bool operator!=(const ChebyCoeff& f, const ChebyCoeff& g) { return !(f == g); }

// ------------------------------------------------------------------------------------------------ calculus
void diff(const ChebyCoeff& u, ChebyCoeff& dudy) {
    if (dudy.numModes() != u.numModes()) dudy.resize(u.numModes());
    dudy.setBounds(u.a(), u.b());
    dudy.setState(Spectral);
    const int Nb = u.numModes() - 1;
    if (Nb == -1) return;
    if (Nb == 0) { dudy[0] = 0.0; return; }
    const Real scale = 4.0 / u.L();
    dudy[Nb] = 0.0;
    dudy[Nb - 1] = scale * Nb * u[Nb];
    for (int n = Nb - 2; n >= 0; --n) dudy[n] = dudy[n + 2] + scale * (n + 1) * u[n + 1];
    dudy[0] *= 0.5;
}
ChebyCoeff diff(const ChebyCoeff& u) {
    ChebyCoeff d(u.numModes(), u.a(), u.b(), Spectral);
    diff(u, d);
    return d;
}
void diff2(const ChebyCoeff& u, ChebyCoeff& d2) { ChebyCoeff t; diff(u, t); diff(t, d2); }
void diff2(const ChebyCoeff& u, ChebyCoeff& d2, ChebyCoeff& tmp) { diff(u, tmp); diff(tmp, d2); }
ChebyCoeff diff2(const ChebyCoeff& u) { return diff(diff(u)); }
void diff(const ChebyCoeff& f, ChebyCoeff& df, int n) {
    df = f;
    ChebyCoeff t;
    for (int k = 0; k < n; ++k) { diff(df, t); swap(df, t); }
}
ChebyCoeff diff(const ChebyCoeff& f, int n) { ChebyCoeff d; diff(f, d, n); return d; }

void integrate(const ChebyCoeff& dudy, ChebyCoeff& u) {
    const int N = dudy.numModes();
    if (u.numModes() != N) u.resize(N);
    u.setBounds(dudy.a(), dudy.b());
    u.setState(Spectral);
    const Real h2 = (dudy.b() - dudy.a()) / 2;
    switch (N) {
        case 0: break;
        case 1: u[0] = 0.0; break;
        case 2: u[0] = 0; u[1] = h2 * dudy[0]; break;
        default:
            u[1] = h2 * (dudy[0] - dudy[2] / 2);
            for (int n = 2; n < N - 1; ++n) u[n] = h2 * (dudy[n - 1] - dudy[n + 1]) / (2 * n);
            u[N - 1] = h2 * dudy[N - 2] / (2 * (N - 1));
            u[0] -= u.mean();  // the constant is free: zero mean
    }
}
ChebyCoeff integrate(const ChebyCoeff& dudy) {
    ChebyCoeff u(dudy.length(), dudy.a(), dudy.b(), Spectral);
    integrate(dudy, u);
    return u;
}

// ------------------------------------------------------------------------------------------------ norms
Real L2InnerProduct(const ChebyCoeff& u, const ChebyCoeff& v, bool normalize) {
    const int N = u.numModes();
    Real sum = 0.0;
    const Real e = 1.0;
    for (int m = N - 1; m >= 0; --m) {
        const Real um = u[m];
        Real psum = 0.0;
        for (int n = m % 2; n < N; n += 2)
            psum += um * v[n] * (e - m * m - n * n) / ((e + m - n) * (e - m + n) * (e + m + n) * (e - m - n));
        sum += psum;
    }
    if (!normalize) sum *= u.b() - u.a();
    return sum;
}
Real L2Norm2(const ChebyCoeff& u, bool normalize) { return L2InnerProduct(u, u, normalize); }
Real L2Norm(const ChebyCoeff& u, bool normalize) { return std::sqrt(L2Norm2(u, normalize)); }
Real L2Dist2(const ChebyCoeff& u, const ChebyCoeff& v, bool normalize) { return L2Norm2(u - v, normalize); }
Real L2Dist(const ChebyCoeff& u, const ChebyCoeff& v, bool normalize) { return std::sqrt(L2Dist2(u, v, normalize)); }
Real chebyInnerProduct(const ChebyCoeff& u, const ChebyCoeff& v, bool normalize) {
    const int N = u.numModes();
    Real sum = 0.0;
    for (int m = N - 1; m > 0; --m) sum += u[m] * v[m];
    if (N > 0) sum += 2 * u[0] * v[0];  // T_0 carries a factor 2
    if (!normalize) sum *= u.b() - u.a();
    return sum * pi / 2;
}
Real chebyNorm2(const ChebyCoeff& u, bool normalize) { return chebyInnerProduct(u, u, normalize); }
Real chebyDist2(const ChebyCoeff& u, const ChebyCoeff& v, bool normalize) { return chebyNorm2(u - v, normalize); }
Real chebyNorm(const ChebyCoeff& u, bool normalize) { return std::sqrt(chebyNorm2(u, normalize)); }
Real chebyDist(const ChebyCoeff& u, const ChebyCoeff& v, bool normalize) { return std::sqrt(chebyDist2(u, v, normalize)); }
Real LinfNorm(const ChebyCoeff& f) {
    ChebyCoeff g(f);
    g.makePhysical();
    Real m = 0.0;
    for (int i = 0; i < g.N(); ++i) m = Greater(std::fabs(g[i]), m);
    return m;
}
Real LinfDist(const ChebyCoeff& f, const ChebyCoeff& g) {
    ChebyCoeff p(f), q(g);
    p.makePhysical();
    q.makePhysical();
    Real m = 0.0;
    for (int i = 0; i < p.N(); ++i) m = Greater(std::fabs(p[i] - q[i]), m);
    return m;
}
static Real integral_of_abs(ChebyCoeff g, bool normalize) {  // g physical on entry
    for (int n = 0; n < g.N(); ++n) g[n] = std::fabs(g[n]);
    g.makeSpectral();
    const ChebyCoeff G = integrate(g);
    Real r = G.eval_b() - G.eval_a();
    if (normalize) r /= g.b() - g.a();
    return r;
}
Real L1Norm(const ChebyCoeff& f, bool normalize) {
    ChebyCoeff g(f);
    g.makePhysical();
    return integral_of_abs(g, normalize);
}
Real L1Dist(const ChebyCoeff& f, const ChebyCoeff& g, bool normalize) {
    ChebyCoeff p(f), q(g);
    p.makePhysical();
    q.makePhysical();
    for (int n = 0; n < p.N(); ++n) p[n] -= q[n];
    return integral_of_abs(p, normalize);
}
Real norm2(const ChebyCoeff& u, NormType n, bool normalize) { return n == Uniform ? L2Norm2(u, normalize) : chebyNorm2(u, normalize); }
Real dist2(const ChebyCoeff& u, const ChebyCoeff& v, NormType n, bool normalize) { return n == Uniform ? L2Dist2(u, v, normalize) : chebyDist2(u, v, normalize); }
Real norm(const ChebyCoeff& u, NormType n, bool normalize) { return n == Uniform ? L2Norm(u, normalize) : chebyNorm(u, normalize); }
Real dist(const ChebyCoeff& u, const ChebyCoeff& v, NormType n, bool normalize) { return n == Uniform ? L2Dist(u, v, normalize) : chebyDist(u, v, normalize); }
Real innerProduct(const ChebyCoeff& u, const ChebyCoeff& v, NormType n, bool normalize) {
    return n == Uniform ? L2InnerProduct(u, v, normalize) : chebyInnerProduct(u, v, normalize);
}

// ------------------------------------------------------------------------------------------------ ComplexChebyCoeff
ComplexChebyCoeff::ComplexChebyCoeff() {}
ComplexChebyCoeff::ComplexChebyCoeff(int N, Real a, Real b, fieldstate s) : re(N, a, b, s), im(N, a, b, s) {}
ComplexChebyCoeff::ComplexChebyCoeff(int N, const ComplexChebyCoeff& f) : re(N, f.re), im(N, f.im) {}
ComplexChebyCoeff::ComplexChebyCoeff(const ChebyCoeff& r, const ChebyCoeff& i) : re(r), im(i) {}
// ascii form: "% N a b state" then "re im" per line
ComplexChebyCoeff::ComplexChebyCoeff(const std::string& filebase) {
    std::ifstream is;
    const std::string filename = ifstreamOpen(is, filebase, ".asc");
    if (!is.good()) cferror("ComplexChebyCoeff::ComplexChebyCoeff(filebase) : can't open file " + filename);
    char c = 0;
    int N = 0;
    Real a = 0, b = 0;
    fieldstate s = Spectral;
    is >> c;
    if (c != '%') cferror("ComplexChebyCoeff::ComplexChebyCoeff(filebase): bad header in file " + filename);
    is >> N >> a >> b >> s;
    re = ChebyCoeff(N, a, b, s);
    im = ChebyCoeff(N, a, b, s);
    for (int i = 0; i < N; ++i) is >> re[i] >> im[i];
    makeSpectral();
}
void ComplexChebyCoeff::save(const std::string& filebase, fieldstate savestate) const {
    ComplexChebyCoeff t(*this);
    t.makeState(savestate);
    std::ofstream os(appendSuffix(filebase, ".asc").c_str());
    os << std::scientific << std::setprecision(REAL_DIGITS);
    os << "% " << t.length() << ' ' << t.a() << ' ' << t.b() << ' ' << t.state() << '\n';
    for (int i = 0; i < t.length(); ++i) os << std::setw(REAL_IOWIDTH) << t.re[i] << ' ' << std::setw(REAL_IOWIDTH) << t.im[i] << '\n';
}
void ComplexChebyCoeff::binaryDump(std::ostream& os) const { re.binaryDump(os); im.binaryDump(os); }
void ComplexChebyCoeff::binaryLoad(std::istream& is) { re.binaryLoad(is); im.binaryLoad(is); }
void ComplexChebyCoeff::reconfig(const ComplexChebyCoeff& f) { re.reconfig(f.re); im.reconfig(f.im); }
void ComplexChebyCoeff::resize(int N) { re.resize(N); im.resize(N); }
void ComplexChebyCoeff::randomize(Real magn, Real decay, BC aBC, BC bBC) { re.randomize(magn, decay, aBC, bBC); im.randomize(magn, decay, aBC, bBC); }
void ComplexChebyCoeff::setToZero() { re.setToZero(); im.setToZero(); }
void ComplexChebyCoeff::setBounds(Real a, Real b) { re.setBounds(a, b); im.setBounds(a, b); }
void ComplexChebyCoeff::setState(fieldstate s) { re.setState(s); im.setState(s); }
void ComplexChebyCoeff::fill(const ComplexChebyCoeff& g) { re.fill(g.re); im.fill(g.im); }
void ComplexChebyCoeff::interpolate(const ComplexChebyCoeff& g) { re.interpolate(g.re); im.interpolate(g.im); }
void ComplexChebyCoeff::reflect(const ComplexChebyCoeff& g, parity p) { re.reflect(g.re, p); im.reflect(g.im, p); }
Complex ComplexChebyCoeff::eval_a() const { return Complex(re.eval_a(), im.eval_a()); }
Complex ComplexChebyCoeff::eval_b() const { return Complex(re.eval_b(), im.eval_b()); }
Complex ComplexChebyCoeff::eval(Real x) const { return Complex(re.eval(x), im.eval(x)); }
Complex ComplexChebyCoeff::slope_a() const { return Complex(re.slope_a(), im.slope_a()); }
Complex ComplexChebyCoeff::slope_b() const { return Complex(re.slope_b(), im.slope_b()); }
Complex ComplexChebyCoeff::mean() const { return Complex(re.mean(), im.mean()); }
ComplexChebyCoeff& ComplexChebyCoeff::operator+=(const ComplexChebyCoeff& f) { re += f.re; im += f.im; return *this; }
ComplexChebyCoeff& ComplexChebyCoeff::operator-=(const ComplexChebyCoeff& f) { re -= f.re; im -= f.im; return *this; }
ComplexChebyCoeff& ComplexChebyCoeff::operator*=(Real c) { re *= c; im *= c; return *this; }
ComplexChebyCoeff& ComplexChebyCoeff::operator*=(Complex c) {
    for (int n = 0; n < length(); ++n) set(n, (*this)[n] * c);
    return *this;
}
ComplexChebyCoeff& ComplexChebyCoeff::operator*=(const ComplexChebyCoeff& g) {
    assert(state() == Physical && g.state() == Physical);
    for (int n = 0; n < length(); ++n) set(n, (*this)[n] * g[n]);
    return *this;
}
void ComplexChebyCoeff::conjugate() { im *= -1.0; }
bool ComplexChebyCoeff::congruent(const ComplexChebyCoeff& g) const { return re.congruent(g.re) && im.congruent(g.im); }
void ComplexChebyCoeff::chebyfft() { ChebyTransform t(N()); chebyfft(t); }
void ComplexChebyCoeff::ichebyfft() { ChebyTransform t(N()); ichebyfft(t); }
void ComplexChebyCoeff::makeSpectral() { ChebyTransform t(N()); makeSpectral(t); }
void ComplexChebyCoeff::makePhysical() { ChebyTransform t(N()); makePhysical(t); }
void ComplexChebyCoeff::makeState(fieldstate s) { ChebyTransform t(N()); makeState(s, t); }
void ComplexChebyCoeff::chebyfft(const ChebyTransform& t) { re.chebyfft(t); im.chebyfft(t); }
void ComplexChebyCoeff::ichebyfft(const ChebyTransform& t) { re.ichebyfft(t); im.ichebyfft(t); }
void ComplexChebyCoeff::makeSpectral(const ChebyTransform& t) { re.makeSpectral(t); im.makeSpectral(t); }
void ComplexChebyCoeff::makePhysical(const ChebyTransform& t) { re.makePhysical(t); im.makePhysical(t); }
void ComplexChebyCoeff::makeState(fieldstate s, const ChebyTransform& t) { re.makeState(s, t); im.makeState(s, t); }
void swap(ComplexChebyCoeff& f, ComplexChebyCoeff& g) { swap(f.re, g.re); swap(f.im, g.im); }

ComplexChebyCoeff operator*(Real c, const ComplexChebyCoeff& g) { ComplexChebyCoeff r(g); r *= c; return r; }
ComplexChebyCoeff operator+(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g) { ComplexChebyCoeff r(f); r += g; return r; }
ComplexChebyCoeff operator-(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g) { ComplexChebyCoeff r(f); r -= g; return r; }
bool operator==(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g) { return f.re == g.re && f.im == g.im; }
bool operator!=(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g) { return !(f == g); }
void diff(const ComplexChebyCoeff& f, ComplexChebyCoeff& df) { diff(f.re, df.re); diff(f.im, df.im); }
void diff2(const ComplexChebyCoeff& f, ComplexChebyCoeff& d2f) { diff2(f.re, d2f.re); diff2(f.im, d2f.im); }
void diff2(const ComplexChebyCoeff& f, ComplexChebyCoeff& d2f, ComplexChebyCoeff& tmp) { diff2(f.re, d2f.re, tmp.re); diff2(f.im, d2f.im, tmp.im); }
void diff(const ComplexChebyCoeff& f, ComplexChebyCoeff& df, int n) { diff(f.re, df.re, n); diff(f.im, df.im, n); }
ComplexChebyCoeff diff(const ComplexChebyCoeff& f) { ComplexChebyCoeff d(f.numModes(), f.a(), f.b(), Spectral); diff(f, d); return d; }
ComplexChebyCoeff diff2(const ComplexChebyCoeff& f) { ComplexChebyCoeff d(f.numModes(), f.a(), f.b(), Spectral); diff2(f, d); return d; }
ComplexChebyCoeff diff(const ComplexChebyCoeff& f, int n) { ComplexChebyCoeff d(f.numModes(), f.a(), f.b(), Spectral); diff(f, d, n); return d; }
void integrate(const ComplexChebyCoeff& df, ComplexChebyCoeff& f) { integrate(df.re, f.re); integrate(df.im, f.im); }
ComplexChebyCoeff integrate(const ComplexChebyCoeff& df) { ComplexChebyCoeff f(df.numModes(), df.a(), df.b(), Spectral); integrate(df, f); return f; }
std::ostream& operator<<(std::ostream& os, const ComplexChebyCoeff& f) {
    for (int i = 0; i < f.length(); ++i) os << '(' << f.re[i] << ", " << f.im[i] << ")\n";
    return os;
}

Real L2Norm2(const ComplexChebyCoeff& u, bool normalize) { return L2Norm2(u.re, normalize) + L2Norm2(u.im, normalize); }
Real L2Dist2(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, bool normalize) { return L2Dist2(u.re, v.re, normalize) + L2Dist2(u.im, v.im, normalize); }
Real L2Norm(const ComplexChebyCoeff& u, bool normalize) { return std::sqrt(L2Norm2(u, normalize)); }
Real L2Dist(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, bool normalize) { return std::sqrt(L2Dist2(u, v, normalize)); }
// <u, v> = integral of u conj(v)
Complex L2InnerProduct(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, bool normalize) {
    return Complex(L2InnerProduct(u.re, v.re, normalize) + L2InnerProduct(u.im, v.im, normalize),
                   L2InnerProduct(u.im, v.re, normalize) - L2InnerProduct(u.re, v.im, normalize));
}
Real chebyNorm2(const ComplexChebyCoeff& u, bool normalize) { return chebyNorm2(u.re, normalize) + chebyNorm2(u.im, normalize); }
Real chebyDist2(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, bool normalize) { return chebyDist2(u.re, v.re, normalize) + chebyDist2(u.im, v.im, normalize); }
Real chebyNorm(const ComplexChebyCoeff& u, bool normalize) { return std::sqrt(chebyNorm2(u, normalize)); }
Real chebyDist(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, bool normalize) { return std::sqrt(chebyDist2(u, v, normalize)); }
Complex chebyInnerProduct(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, bool normalize) {
    return Complex(chebyInnerProduct(u.re, v.re, normalize) + chebyInnerProduct(u.im, v.im, normalize),
                   chebyInnerProduct(u.im, v.re, normalize) - chebyInnerProduct(u.re, v.im, normalize));
}
Real norm2(const ComplexChebyCoeff& u, NormType n, bool normalize) { return n == Uniform ? L2Norm2(u, normalize) : chebyNorm2(u, normalize); }
Real dist2(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, NormType n, bool normalize) { return n == Uniform ? L2Dist2(u, v, normalize) : chebyDist2(u, v, normalize); }
Real norm(const ComplexChebyCoeff& u, NormType n, bool normalize) { return n == Uniform ? L2Norm(u, normalize) : chebyNorm(u, normalize); }
Real dist(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, NormType n, bool normalize) { return n == Uniform ? L2Dist(u, v, normalize) : chebyDist(u, v, normalize); }
Complex innerProduct(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, NormType n, bool normalize) {
    return n == Uniform ? L2InnerProduct(u, v, normalize) : chebyInnerProduct(u, v, normalize);
}
Real L1Norm(const ComplexChebyCoeff& u, bool normalize) { return L1Norm(u.re, normalize) + L1Norm(u.im, normalize); }
Real L1Dist(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, bool normalize) { return L1Dist(u.re, v.re, normalize) + L1Dist(u.im, v.im, normalize); }
Real LinfNorm(const ComplexChebyCoeff& u) { return LinfNorm(u.re) + LinfNorm(u.im); }
Real LinfDist(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v) { return LinfDist(u.re, v.re) + LinfDist(u.im, v.im); }

}  // namespace chflow
