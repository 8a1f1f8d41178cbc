// chflow::Vector -- the plain real vector of the Channelflow API (reference cfbasics/cfvector.h): base class of
// ChebyCoeff, argument of BandedTridiag.  Host memory (std::vector): these are O(Ny) objects.
#ifndef CFB200_CFVECTOR_H
#define CFB200_CFVECTOR_H
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <vector>

#include "cfbasics/cfbasics.h"
#include "cfbasics/mathdefs.h"

namespace chflow {

class Vector {
   public:
    explicit Vector(int N = 0) : data_(N, 0.0) {}
    explicit Vector(const std::string& filebase);  // reads the .asc form written by save()
    void resize(int N) { data_.resize(N); }
    void setToZero() { std::fill(data_.begin(), data_.end(), 0.0); }

    Real& operator[](int i) { assert(i >= 0 && (size_t)i < data_.size()); return data_[i]; }
    Real operator[](int i) const { assert(i >= 0 && (size_t)i < data_.size()); return data_[i]; }
    Real& operator()(int i) { return (*this)[i]; }
    Real operator()(int i) const { return (*this)[i]; }

    Vector& operator*=(Real c) { for (auto& x : data_) x *= c; return *this; }
    Vector& operator/=(Real c) { return *this *= 1.0 / c; }
    Vector& operator+=(Real c) { for (auto& x : data_) x += c; return *this; }
    Vector& operator-=(Real c) { for (auto& x : data_) x -= c; return *this; }
    Vector& operator+=(const Vector& a) { assert(a.length() == length()); for (size_t i = 0; i < data_.size(); ++i) data_[i] += a.data_[i]; return *this; }
    Vector& operator-=(const Vector& a) { assert(a.length() == length()); for (size_t i = 0; i < data_.size(); ++i) data_[i] -= a.data_[i]; return *this; }
    Vector& dottimes(const Vector& a) { assert(a.length() == length()); for (size_t i = 0; i < data_.size(); ++i) data_[i] *= a.data_[i]; return *this; }
    Vector& dotdivide(const Vector& a) { assert(a.length() == length()); for (size_t i = 0; i < data_.size(); ++i) data_[i] /= a.data_[i]; return *this; }
    Vector& abs() { for (auto& x : data_) x = std::fabs(x); return *this; }

    Vector subvector(int offset, int N) const {
        Vector s(N);
        for (int i = 0; i < N; ++i) s[i] = data_[i + offset];
        return s;
    }
    Vector modularSubvector(int offset, int N) const {
        Vector s(N);
        for (int i = 0; i < N; ++i) s[i] = data_[(i + offset) % data_.size()];
        return s;
    }
    int length() const { return (int)data_.size(); }
    const Real* pointer() const { return data_.data(); }
    Real* pointer() { return data_.data(); }
    void save(const std::string& filebase) const {
        std::ofstream os((filebase + ".asc").c_str());
        os << std::scientific << std::setprecision(REAL_DIGITS) << "% " << data_.size() << " 1\n";
        for (Real x : data_) os << std::setw(REAL_IOWIDTH) << x << '\n';
    }

   protected:
    std::vector<Real> data_;
};

inline Vector::Vector(const std::string& filebase) {
    std::ifstream is;
    const std::string filename = ifstreamOpen(is, filebase, ".asc");
    if (!is) cferror("Vector::Vector(filebase) : can't open file " + filebase + " or " + filebase + ".asc");
    char c = 0;
    int M = 0, N = 0;
    is >> c;
    if (c != '%') cferror("Vector(filebase): bad header in file " + filename);
    is >> M >> N;
    data_.assign(M > N ? M : N, 0.0);
    for (auto& x : data_) is >> x;
}

inline void assign(Vector& u, int uistart, int uistride, int uiend, Vector& v, int vistart, int vistride, int) {
    for (int ui = uistart, vi = vistart; ui < uiend; ui += uistride, vi += vistride) u[ui] = v[vi];
}
inline Vector operator*(Real c, const Vector& v) { Vector u(v); u *= c; return u; }
inline Vector operator+(const Vector& u, const Vector& v) { Vector w(u); w += v; return w; }
inline Vector operator-(const Vector& u, const Vector& v) { Vector w(u); w -= v; return w; }
inline Real operator*(const Vector& u, const Vector& v) {
    assert(u.length() == v.length());
    Real s = 0.0;
    for (int i = 0; i < u.length(); ++i) s += u[i] * v[i];
    return s;
}
inline bool operator==(const Vector& u, const Vector& v) {
    if (u.length() != v.length()) return false;
    for (int i = 0; i < u.length(); ++i)
        if (u[i] != v[i]) return false;
    return true;
}
inline Vector dottimes(const Vector& u, const Vector& v) { Vector w(u); w.dottimes(v); return w; }
inline Vector dotdivide(const Vector& u, const Vector& v) { Vector w(u); w.dotdivide(v); return w; }
inline Vector vabs(const Vector& v) { Vector r(v); r.abs(); return r; }
inline Real L1Norm(const Vector& v) { Real s = 0.0; for (int i = 0; i < v.length(); ++i) s += std::fabs(v[i]); return s; }
inline Real L2Norm2(const Vector& v) { Real s = 0.0; for (int i = 0; i < v.length(); ++i) s += square(v[i]); return s; }
inline Real L2Norm(const Vector& v) { return std::sqrt(L2Norm2(v)); }
inline Real LinfNorm(const Vector& v) { Real m = 0.0; for (int i = 0; i < v.length(); ++i) m = Greater(std::fabs(v[i]), m); return m; }
inline Real L1Dist(const Vector& u, const Vector& v) { Real s = 0.0; for (int i = 0; i < v.length(); ++i) s += std::fabs(u[i] - v[i]); return s; }
inline Real L2Dist2(const Vector& u, const Vector& v) { Real s = 0.0; for (int i = 0; i < v.length(); ++i) s += square(u[i] - v[i]); return s; }
inline Real L2Dist(const Vector& u, const Vector& v) { return std::sqrt(L2Dist2(u, v)); }
inline Real LinfDist(const Vector& u, const Vector& v) { Real m = 0.0; for (int i = 0; i < v.length(); ++i) m = Greater(std::fabs(u[i] - v[i]), m); return m; }
inline Real mean(const Vector& v) { Real s = 0.0; for (int i = 0; i < v.length(); ++i) s += v[i]; return s / v.length(); }
inline int maxElemIndex(const Vector& v) {
    int idx = 0;
    Real m = 0.0;
    for (int i = 0; i < v.length(); ++i)
        if (std::fabs(v[i]) > m) { m = std::fabs(v[i]); idx = i; }
    return idx;
}
inline std::ostream& operator<<(std::ostream& os, const Vector& a) {
    const char sep = a.length() < 10 ? ' ' : '\n';
    for (int i = 0; i < a.length(); ++i) os << a[i] << sep;
    return os;
}

}  // namespace chflow
#endif
