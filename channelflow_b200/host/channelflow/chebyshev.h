// 1-D Chebyshev expansions on the host (small O(N) / O(N^2) work: base-flow profiles, diagnostics).
// Same public names as the reference's channelflow/chebyshev.h:43-197 (ChebyCoeff, ComplexChebyCoeff,
// ChebyTransform, diff, diff2, integrate); the y-transform here is a direct O(N^2) DCT-I, the bulk transforms of
// FlowFields run on the GPU (cfgpu_field_make_*).
#ifndef CFB200_CHEBYSHEV_H
#define CFB200_CHEBYSHEV_H
#include <vector>

#include "cfbasics/mathdefs.h"

namespace chflow {

class ChebyTransform {
   public:
    explicit ChebyTransform(int N = 0) : N_(N) {}
    int N() const { return N_; }
    int length() const { return N_; }

   private:
    int N_;
};

class ChebyCoeff {
   public:
    ChebyCoeff() : a_(0), b_(0), state_(Spectral) {}
    ChebyCoeff(int N, Real a = -1, Real b = 1, fieldstate s = Spectral) : data_(N, 0.0), a_(a), b_(b), state_(s) {}

    Real& operator[](int n) { return data_[n]; }
    const Real& operator[](int n) const { return data_[n]; }
    Real& operator()(int n) { return data_[n]; }
    const Real& operator()(int n) const { return data_[n]; }

    int length() const { return (int)data_.size(); }
    int numModes() const { return (int)data_.size(); }
    int N() const { return (int)data_.size(); }
    void resize(int N) { data_.resize(N, 0.0); }
    Real a() const { return a_; }
    Real b() const { return b_; }
    Real L() const { return b_ - a_; }
    void setBounds(Real a, Real b) { a_ = a; b_ = b; }
    fieldstate state() const { return state_; }
    void setState(fieldstate s) { state_ = s; }
    void setToZero() { for (auto& x : data_) x = 0.0; }

    Real eval_a() const;  // u(a)
    Real eval_b() const;  // u(b)
    Real mean() const;

    void makePhysical();
    void makeSpectral();
    void makePhysical(const ChebyTransform&) { makePhysical(); }
    void makeSpectral(const ChebyTransform&) { makeSpectral(); }
    void makeState(fieldstate s) { if (s == Physical) makePhysical(); else makeSpectral(); }

    ChebyCoeff& operator*=(Real c) { for (auto& x : data_) x *= c; return *this; }
    ChebyCoeff& operator+=(const ChebyCoeff& o) { for (int i = 0; i < length(); ++i) data_[i] += o.data_[i]; return *this; }
    ChebyCoeff& operator-=(const ChebyCoeff& o) { for (int i = 0; i < length(); ++i) data_[i] -= o.data_[i]; return *this; }

    const std::vector<Real>& data() const { return data_; }

   private:
    std::vector<Real> data_;
    Real a_, b_;
    fieldstate state_;
};

class ComplexChebyCoeff {
   public:
    ComplexChebyCoeff() {}
    ComplexChebyCoeff(int N, Real a = -1, Real b = 1, fieldstate s = Spectral) : re(N, a, b, s), im(N, a, b, s) {}
    Complex operator[](int n) const { return Complex(re[n], im[n]); }
    void set(int n, Complex c) { re[n] = c.real(); im[n] = c.imag(); }
    int length() const { return re.length(); }
    ChebyCoeff re, im;
};

inline ChebyCoeff Re(const ComplexChebyCoeff& c) { return c.re; }
inline ChebyCoeff Im(const ComplexChebyCoeff& c) { return c.im; }

void diff(const ChebyCoeff& u, ChebyCoeff& dudy);
ChebyCoeff diff(const ChebyCoeff& u);
void diff2(const ChebyCoeff& u, ChebyCoeff& d2udy2);
ChebyCoeff diff2(const ChebyCoeff& u);
void integrate(const ChebyCoeff& dudy, ChebyCoeff& u);
ChebyCoeff integrate(const ChebyCoeff& dudy);
std::vector<Real> chebypoints(int N, Real a, Real b);

Real L2Norm2(const ChebyCoeff& u, bool normalize = true);
Real L2Norm(const ChebyCoeff& u, bool normalize = true);
Real L2InnerProduct(const ChebyCoeff& u, const ChebyCoeff& v, bool normalize = true);

}  // namespace chflow
#endif
