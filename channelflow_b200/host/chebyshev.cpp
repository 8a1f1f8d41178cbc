// Host Chebyshev utilities; formulas follow the reference's chebyshev.cpp (eval_a/eval_b :405-430, mean :505-512,
// integrate :637-664, diff :672-697, L2Norm2 :758-773, L2InnerProduct :785-802, transforms :262-302).
#include "channelflow/chebyshev.h"

namespace chflow {

Real ChebyCoeff::eval_b() const {
    if (data_.empty()) return 0;
    if (state_ == Physical) return data_[0];
    Real sum = 0.0;
    for (int n = length() - 1; n >= 0; --n) sum += data_[n];
    return sum;
}
Real ChebyCoeff::eval_a() const {
    if (data_.empty()) return 0;
    if (state_ == Physical) return data_[length() - 1];
    Real sum = 0.0;
    for (int n = length() - 1; n >= 0; --n) sum += data_[n] * ((n % 2 == 0) ? 1 : -1);
    return sum;
}
Real ChebyCoeff::mean() const {
    Real sum = data_[0];
    for (unsigned n = 2; n < data_.size(); n += 2) sum -= data_[n] / (n * n - 1);
    return sum;
}

static long double cospi_frac(long num, long den) {
    const long double PIl = 3.141592653589793238462643383279502884L;
    num %= 2 * den;
    if (num < 0) num += 2 * den;
    if (num > den) num = 2 * den - num;
    if (2 * num == den) return 0.0L;
    if (2 * num > den) return -cosl(PIl * (long double)(den - num) / (long double)den);
    return cosl(PIl * (long double)num / (long double)den);
}

void ChebyCoeff::makePhysical() {
    if (state_ == Physical) return;
    const int N = length(), Nb = N - 1;
    if (N >= 2) {
        std::vector<Real> u(N);
        for (int j = 0; j < N; ++j) {
            long double s = 0.0L;
            for (int n = 0; n < N; ++n) s += (long double)data_[n] * cospi_frac((long)j * n, Nb);
            u[j] = (Real)s;
        }
        data_ = u;
    }
    state_ = Physical;
}
void ChebyCoeff::makeSpectral() {
    if (state_ == Spectral) return;
    const int N = length(), Nb = N - 1;
    if (N >= 2) {
        std::vector<Real> c(N);
        for (int n = 0; n < N; ++n) {
            long double s = 0.0L;
            for (int j = 0; j < N; ++j) s += (long double)data_[j] * cospi_frac((long)j * n, Nb) * ((j == 0 || j == Nb) ? 1.0L : 2.0L);
            c[n] = (Real)(s * ((n == 0 || n == Nb) ? 0.5L : 1.0L) / (long double)Nb);
        }
        data_ = c;
    }
    state_ = Spectral;
}

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
void diff2(const ChebyCoeff& u, ChebyCoeff& d2) {
    ChebyCoeff d = diff(u);
    diff(d, d2);
}
ChebyCoeff diff2(const ChebyCoeff& u) { return diff(diff(u)); }

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
            u[0] -= u.mean();
    }
}
ChebyCoeff integrate(const ChebyCoeff& dudy) {
    ChebyCoeff u(dudy.length(), dudy.a(), dudy.b(), Spectral);
    integrate(dudy, u);
    return u;
}

std::vector<Real> chebypoints(int N, Real a, Real b) {
    std::vector<Real> y(N);
    const Real piN = pi / (N - 1);
    const Real radius = (b - a) / 2, center = (b + a) / 2;
    for (int j = 0; j < N; ++j) y[j] = center + radius * cos(piN * j);
    return y;
}

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
Real L2Norm(const ChebyCoeff& u, bool normalize) { return sqrt(L2Norm2(u, normalize)); }

}  // namespace chflow
