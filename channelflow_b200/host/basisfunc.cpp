// BasisFunc (see channelflow/basisfunc.h); file forms as in the reference's basisfunc.cpp:43-135.
#include "channelflow/basisfunc.h"

#include <fstream>
#include <iomanip>

namespace chflow {

BasisFunc::BasisFunc() : Nd_(0), Ny_(0), kx_(0), kz_(0), Lx_(0), Lz_(0), a_(0), b_(0), state_(Spectral) {}
BasisFunc::BasisFunc(int Nd, int Ny, int kx, int kz, Real Lx, Real Lz, Real a, Real b, fieldstate s)
    : Nd_(Nd), Ny_(Ny), kx_(kx), kz_(kz), Lx_(Lx), Lz_(Lz), a_(a), b_(b), state_(s), u_(Nd, ComplexChebyCoeff(Ny, a, b, s)) {}
BasisFunc::BasisFunc(int Ny, int kx, int kz, Real Lx, Real Lz, Real a, Real b, fieldstate s) : BasisFunc(3, Ny, kx, kz, Lx, Lz, a, b, s) {}
BasisFunc::BasisFunc(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, const ComplexChebyCoeff& w, int kx, int kz, Real Lx, Real Lz)
    : Nd_(3), Ny_(u.numModes()), kx_(kx), kz_(kz), Lx_(Lx), Lz_(Lz), a_(u.a()), b_(u.b()), state_(u.state()), u_{u, v, w} {}
BasisFunc::BasisFunc(const std::string& filebase) : BasisFunc() {
    const std::string filename = filebase + ".asc";
    std::ifstream is(filename.c_str());
    if (!is.good()) cferror("BasisFunc::BasisFunc(filebase) : can't open file " + filename);
    char c = 0;
    is >> c;
    if (c != '%') cferror("BasisFunc::BasisFunc(filebase): bad header in file " + filename);
    is >> Nd_ >> Ny_ >> kx_ >> kz_ >> Lx_ >> Lz_ >> a_ >> b_ >> state_;
    u_.assign(Nd_, ComplexChebyCoeff(Ny_, a_, b_, state_));
    for (int ny = 0; ny < Ny_; ++ny)
        for (int n = 0; n < Nd_; ++n) {
            Real r = 0, i = 0;
            is >> r >> i;
            u_[n].set(ny, Complex(r, i));
        }
}
void BasisFunc::save(const std::string& filebase, fieldstate savestate) const {
    BasisFunc t(*this);
    t.makeState(savestate);
    std::ofstream os((filebase + ".asc").c_str());
    os << std::scientific << std::setprecision(REAL_DIGITS);
    os << "% " << Nd_ << ' ' << Ny_ << ' ' << kx_ << ' ' << kz_ << ' ' << Lx_ << ' ' << Lz_ << ' ' << a_ << ' ' << b_ << ' ' << t.state_ << '\n';
    for (int ny = 0; ny < Ny_; ++ny) {
        for (int n = 0; n < Nd_; ++n) os << t.u_[n].re[ny] << ' ' << t.u_[n].im[ny] << ' ';
        os << '\n';
    }
}
void BasisFunc::binaryDump(std::ostream& os) const {
    write(os, Nd_); write(os, Ny_); write(os, kx_); write(os, kz_);
    write(os, Lx_); write(os, Lz_); write(os, a_); write(os, b_); write(os, state_);
    for (const auto& p : u_) p.binaryDump(os);
}
void BasisFunc::binaryLoad(std::istream& is) {
    if (!is.good()) cferror("BasisFunc::binaryLoad(istream& is) : input error");
    read(is, Nd_); read(is, Ny_); read(is, kx_); read(is, kz_);
    read(is, Lx_); read(is, Lz_); read(is, a_); read(is, b_); read(is, state_);
    u_.assign(Nd_, ComplexChebyCoeff());
    for (auto& p : u_) p.binaryLoad(is);
}
void BasisFunc::reconfig(const BasisFunc& f) { *this = f; setToZero(); }
void BasisFunc::resize(int Ny) { Ny_ = Ny; for (auto& p : u_) p.resize(Ny); }
void BasisFunc::setBounds(Real Lx, Real Lz, Real a, Real b) { Lx_ = Lx; Lz_ = Lz; a_ = a; b_ = b; for (auto& p : u_) p.setBounds(a, b); }
void BasisFunc::setState(fieldstate s) { state_ = s; for (auto& p : u_) p.setState(s); }
void BasisFunc::setToZero() { for (auto& p : u_) p.setToZero(); }
void BasisFunc::conjugate() { kx_ = -kx_; kz_ = -kz_; for (auto& p : u_) p.conjugate(); }
void BasisFunc::fill(const BasisFunc& f) { for (int n = 0; n < Nd_ && n < f.Nd_; ++n) u_[n].fill(f.u_[n]); }
void BasisFunc::chebyfft(const ChebyTransform& t) { for (auto& p : u_) p.chebyfft(t); state_ = Spectral; }
void BasisFunc::ichebyfft(const ChebyTransform& t) { for (auto& p : u_) p.ichebyfft(t); state_ = Physical; }
void BasisFunc::makeSpectral(const ChebyTransform& t) { if (state_ == Physical) chebyfft(t); }
void BasisFunc::makePhysical(const ChebyTransform& t) { if (state_ == Spectral) ichebyfft(t); }
void BasisFunc::makeState(fieldstate s, const ChebyTransform& t) { if (s == Physical) makePhysical(t); else makeSpectral(t); }
void BasisFunc::chebyfft() { ChebyTransform t(Ny_); chebyfft(t); }
void BasisFunc::ichebyfft() { ChebyTransform t(Ny_); ichebyfft(t); }
void BasisFunc::makeSpectral() { ChebyTransform t(Ny_); makeSpectral(t); }
void BasisFunc::makePhysical() { ChebyTransform t(Ny_); makePhysical(t); }
void BasisFunc::makeState(fieldstate s) { ChebyTransform t(Ny_); makeState(s, t); }
bool BasisFunc::geomCongruent(const BasisFunc& f) const { return f.Ny_ == Ny_ && f.Lx_ == Lx_ && f.Lz_ == Lz_ && f.a_ == a_ && f.b_ == b_; }
bool BasisFunc::congruent(const BasisFunc& f) const { return geomCongruent(f) && f.kx_ == kx_ && f.kz_ == kz_ && f.state_ == state_; }
bool BasisFunc::interoperable(const BasisFunc& f) const { return geomCongruent(f) && f.state_ == state_ && f.Nd_ == Nd_; }
BasisFunc& BasisFunc::operator*=(Real c) { for (auto& p : u_) p *= c; return *this; }
BasisFunc& BasisFunc::operator*=(Complex c) { for (auto& p : u_) p *= c; return *this; }
BasisFunc& BasisFunc::operator+=(const BasisFunc& g) { for (int n = 0; n < Nd_; ++n) u_[n] += g.u_[n]; return *this; }
BasisFunc& BasisFunc::operator-=(const BasisFunc& g) { for (int n = 0; n < Nd_; ++n) u_[n] -= g.u_[n]; return *this; }

BasisFunc conjugate(const BasisFunc& f) { BasisFunc g(f); g.conjugate(); return g; }
Real L2Norm2(const BasisFunc& f, bool normalize) {
    Real s = 0.0;
    for (int n = 0; n < f.Nd(); ++n) s += L2Norm2(f[n], normalize);
    if (!normalize) s *= f.Lx() * f.Lz();
    return s;
}
Real L2Norm(const BasisFunc& f, bool normalize) { return sqrt(L2Norm2(f, normalize)); }
Real L2Dist2(const BasisFunc& f, const BasisFunc& g, bool normalize) {
    if (f.kx() != g.kx() || f.kz() != g.kz()) return L2Norm2(f, normalize) + L2Norm2(g, normalize);
    Real s = 0.0;
    for (int n = 0; n < f.Nd(); ++n) s += L2Dist2(f[n], g[n], normalize);
    if (!normalize) s *= f.Lx() * f.Lz();
    return s;
}
Real L2Dist(const BasisFunc& f, const BasisFunc& g, bool normalize) { return sqrt(L2Dist2(f, g, normalize)); }
Complex L2InnerProduct(const BasisFunc& f, const BasisFunc& g, bool normalize) {
    if (f.kx() != g.kx() || f.kz() != g.kz()) return Complex(0.0, 0.0);
    Complex s(0.0, 0.0);
    for (int n = 0; n < f.Nd(); ++n) s += L2InnerProduct(f[n], g[n], normalize);
    if (!normalize) s *= f.Lx() * f.Lz();
    return s;
}
Real divNorm2(const BasisFunc& f, bool normalize) {
    ComplexChebyCoeff d = f.u();
    d *= Complex(0.0, 2 * pi * f.kx() / f.Lx());
    ComplexChebyCoeff t = f.w();
    t *= Complex(0.0, 2 * pi * f.kz() / f.Lz());
    d += t;
    diff(f.v(), t);
    d += t;
    return L2Norm2(d, normalize);
}
Real divNorm(const BasisFunc& f, bool normalize) { return sqrt(divNorm2(f, normalize)); }
Real bcNorm2(const BasisFunc& f, bool normalize) {
    Real s = 0.0;
    for (int n = 0; n < f.Nd(); ++n) s += abs2(f[n].eval_a()) + abs2(f[n].eval_b());
    if (!normalize) s *= f.Lx() * f.Lz();
    return s;
}
Real bcNorm(const BasisFunc& f, bool normalize) { return sqrt(bcNorm2(f, normalize)); }

}  // namespace chflow
