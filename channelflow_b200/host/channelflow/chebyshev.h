// Real- and complex-valued 1-d Chebyshev expansions on the host (base-flow profiles, single-mode profiles, diagnostics).
// Same public surface as the reference's channelflow/chebyshev.h:43-309.  These are O(Ny) objects: the transform is a
// direct O(Ny^2) DCT-I with extended-precision cosines (the reference calls FFTW REDFT00); the bulk y-transforms of
// FlowFields run on the GPU as a DMMA contraction (cfgpu_field_make_*).
#ifndef CFB200_CHEBYSHEV_H
#define CFB200_CHEBYSHEV_H
#include <memory>
#include <type_traits>

#include "cfbasics/cfvector.h"
#include "cfbasics/mathdefs.h"
#include "channelflow/cfmpi.h"

#ifndef FFTW_ESTIMATE
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_MEASURE (0U)
#define FFTW_PATIENT (1U << 5)
#define FFTW_EXHAUSTIVE (1U << 3)
#endif

namespace chflow {

enum BC { Free, Diri };
enum NormType { Uniform, Chebyshev };  // uniform weight in y, or 1/sqrt(1-y^2)

// FFTW planner state does not exist here; accepted and ignored
inline void fftw_loadwisdom(const char* = nullptr) {}
inline void fftw_savewisdom(const char* = nullptr) {}

Vector chebypoints(int N, Real a, Real b);
Real chebyIP(int m, int n);  // integral over [-1,1] of T_m T_n
inline int cheby_c(int n) { return n > 0 ? 1 : (n == 0 ? 2 : 0); }
Real legendre(int n, Real x);
Real chebyshev(int n, Real x);
void gaussLegendreQuadrature(int N, Real a, Real b, Vector& x, Vector& w);

class ChebyTransform;

class ChebyCoeff : public Vector {
   public:
    ChebyCoeff();
    ChebyCoeff(int N, Real a = -1, Real b = 1, fieldstate s = Spectral);
    ChebyCoeff(const Vector& v, Real a, Real b, fieldstate s = Spectral);
    ChebyCoeff(int N, const ChebyCoeff& g);   // first N coefficients of g
    ChebyCoeff(const std::string& filebase);  // ascii file written by save()
    ~ChebyCoeff();

    void save(const std::string& filebase, fieldstate s = Physical) const;
    void binaryDump(std::ostream& os) const;
    void binaryLoad(std::istream& is);
    void reconfig(const ChebyCoeff& f);
    void randomize(Real magn, Real smoothness, BC aBC, BC bBC);
    void setBounds(Real a, Real b);
    void setState(fieldstate s);
    void setToZero();
    void fill(const ChebyCoeff& g);
    void interpolate(const ChebyCoeff& g);
    void reflect(const ChebyCoeff& g, parity p);

    Real eval_a() const;
    Real eval_b() const;
    Real eval(Real x) const;
    ChebyCoeff eval(const Vector& x) const;
    void eval(const Vector& x, ChebyCoeff& g) const;
    Real slope_a() const;
    Real slope_b() const;
    Real a() const { return a_; }
    Real b() const { return b_; }
    Real L() const { return b_ - a_; }
    int N() const { return (int)data_.size(); }
    int numModes() const { return (int)data_.size(); }
    fieldstate state() const { return state_; }
    Real mean() const;

    ChebyCoeff& operator*=(Real c);
    ChebyCoeff& operator+=(const ChebyCoeff& g);
    ChebyCoeff& operator-=(const ChebyCoeff& g);
    ChebyCoeff& operator*=(const ChebyCoeff& g);  // pointwise, Physical only

    void chebyfft();
    void ichebyfft();
    void makeSpectral();
    void makePhysical();
    void makeState(fieldstate s);
    void chebyfft(const ChebyTransform& t);
    void ichebyfft(const ChebyTransform& t);
    void makeSpectral(const ChebyTransform& t);
    void makePhysical(const ChebyTransform& t);
    void makeState(fieldstate s, const ChebyTransform& t);

    bool congruent(const ChebyCoeff& g) const;
    friend void swap(ChebyCoeff& f, ChebyCoeff& g);
    const std::vector<Real>& data() const { return data_; }

   private:
    Real a_, b_;
    fieldstate state_;
    friend class ChebyTransform;
};

class ComplexChebyCoeff {
   public:
    ComplexChebyCoeff();
    ComplexChebyCoeff(int N, Real a = -1, Real b = 1, fieldstate s = Spectral);
    ComplexChebyCoeff(int N, const ComplexChebyCoeff& f);
    ComplexChebyCoeff(const ChebyCoeff& re, const ChebyCoeff& im);
    ComplexChebyCoeff(const std::string& filebase);

    void reconfig(const ComplexChebyCoeff& f);
    void resize(int N);
    void randomize(Real magn, Real smoothness, BC aBC, BC bBC);
    void setToZero();
    void setBounds(Real a, Real b);
    void setState(fieldstate s);
    void fill(const ComplexChebyCoeff& g);
    void interpolate(const ComplexChebyCoeff& g);
    void reflect(const ComplexChebyCoeff& g, parity p);

    Complex eval_a() const;
    Complex eval_b() const;
    Complex eval(Real x) const;
    Complex slope_a() const;
    Complex slope_b() const;
    Complex mean() const;
    Real a() const { return re.a(); }
    Real b() const { return re.b(); }
    Real L() const { return re.L(); }
    int N() const { return re.N(); }
    int length() const { return re.length(); }
    int numModes() const { return re.numModes(); }
    fieldstate state() const { return re.state(); }
    Complex operator[](int n) const { return Complex(re[n], im[n]); }
    void set(int n, Complex c) { re[n] = c.real(); im[n] = c.imag(); }
    void add(int n, Complex c) { re[n] += c.real(); im[n] += c.imag(); }
    void sub(int n, Complex c) { re[n] -= c.real(); im[n] -= c.imag(); }

    ComplexChebyCoeff& operator+=(const ComplexChebyCoeff& f);
    ComplexChebyCoeff& operator-=(const ComplexChebyCoeff& f);
    ComplexChebyCoeff& operator*=(Real c);
    ComplexChebyCoeff& operator*=(Complex c);
    ComplexChebyCoeff& operator*=(const ComplexChebyCoeff& c);  // pointwise
    void conjugate();
    void save(const std::string& filebase, fieldstate s = Physical) const;
    void binaryDump(std::ostream& os) const;
    void binaryLoad(std::istream& is);
    bool congruent(const ComplexChebyCoeff& g) const;

    void chebyfft();
    void ichebyfft();
    void makeSpectral();
    void makePhysical();
    void makeState(fieldstate s);
    void chebyfft(const ChebyTransform& t);
    void ichebyfft(const ChebyTransform& t);
    void makeSpectral(const ChebyTransform& t);
    void makePhysical(const ChebyTransform& t);
    void makeState(fieldstate s, const ChebyTransform& t);
    friend void swap(ComplexChebyCoeff& f, ComplexChebyCoeff& g);

    ChebyCoeff re;
    ChebyCoeff im;
};

// Cosine table of one transform length (the reference holds an FFTW plan)
class ChebyTransform {
   public:
    ChebyTransform(int N = 0, uint fftw_flags = FFTW_ESTIMATE);
    int N() const { return N_; }
    int length() const { return N_; }
    // out[j] = sum_n w(n) in[n] cos(pi j n/(N-1)) with the end weights of REDFT00 (chebyshev.cpp:262-302)
    void inverse(std::vector<Real>& x) const;  // coefficients -> Gauss-Lobatto values
    void forward(std::vector<Real>& x) const;  // values -> coefficients

   private:
    int N_;
    uint flags_;
    std::shared_ptr<std::vector<Real>> cos_;  // cos(pi k/(N-1)), k = 0 .. 2(N-1)-1
    friend class ChebyCoeff;
};

ChebyCoeff operator*(Real c, const ChebyCoeff& g);
ChebyCoeff operator+(const ChebyCoeff& f, const ChebyCoeff& g);
ChebyCoeff operator-(const ChebyCoeff& f, const ChebyCoeff& g);
bool operator==(const ChebyCoeff& f, const ChebyCoeff& g);
bool operator!=(const ChebyCoeff& f, const ChebyCoeff& g);

void diff(const ChebyCoeff& f, ChebyCoeff& df);
void diff2(const ChebyCoeff& f, ChebyCoeff& d2f);
void diff2(const ChebyCoeff& f, ChebyCoeff& d2f, ChebyCoeff& tmp);
void diff(const ChebyCoeff& f, ChebyCoeff& df, int n);
ChebyCoeff diff(const ChebyCoeff& f);
ChebyCoeff diff2(const ChebyCoeff& f);
ChebyCoeff diff(const ChebyCoeff& f, int n);
void integrate(const ChebyCoeff& df, ChebyCoeff& f);
ChebyCoeff integrate(const ChebyCoeff& df);
void legendre(int n, ChebyCoeff& u, ChebyTransform& trans, bool normalize = false);

Real L2Norm2(const ChebyCoeff& f, bool normalize = true);
Real L2Dist2(const ChebyCoeff& f, const ChebyCoeff& g, bool normalize = true);
Real L2Norm(const ChebyCoeff& f, bool normalize = true);
Real L2Dist(const ChebyCoeff& f, const ChebyCoeff& g, bool normalize = true);
Real L2InnerProduct(const ChebyCoeff& f, const ChebyCoeff& g, bool normalize = true);
Real chebyNorm2(const ChebyCoeff& f, bool normalize = true);
Real chebyDist2(const ChebyCoeff& f, const ChebyCoeff& g, bool normalize = true);
Real chebyNorm(const ChebyCoeff& f, bool normalize = true);
Real chebyDist(const ChebyCoeff& f, const ChebyCoeff& g, bool normalize = true);
Real chebyInnerProduct(const ChebyCoeff& f, const ChebyCoeff& g, bool normalize = true);
Real norm2(const ChebyCoeff& f, NormType n, bool normalize = true);
Real norm(const ChebyCoeff& f, NormType n, bool normalize = true);
Real dist2(const ChebyCoeff& f, const ChebyCoeff& g, NormType n, bool normalize = true);
Real dist(const ChebyCoeff& f, const ChebyCoeff& g, NormType n, bool normalize = true);
Real innerProduct(const ChebyCoeff& f, const ChebyCoeff& g, NormType n, bool normalize = true);
Real L1Norm(const ChebyCoeff& f, bool normalize = true);
Real L1Dist(const ChebyCoeff& f, const ChebyCoeff& g, bool normalize = true);
Real LinfNorm(const ChebyCoeff& f);
Real LinfDist(const ChebyCoeff& f, const ChebyCoeff& g);

ComplexChebyCoeff operator*(Real c, const ComplexChebyCoeff& g);
ComplexChebyCoeff operator+(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g);
ComplexChebyCoeff operator-(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g);
bool operator==(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g);
bool operator!=(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g);
void diff(const ComplexChebyCoeff& f, ComplexChebyCoeff& df);
void diff2(const ComplexChebyCoeff& f, ComplexChebyCoeff& d2f);
void diff2(const ComplexChebyCoeff& f, ComplexChebyCoeff& d2f, ComplexChebyCoeff& tmp);
void diff(const ComplexChebyCoeff& f, ComplexChebyCoeff& df, int n);
ComplexChebyCoeff diff(const ComplexChebyCoeff& f);
ComplexChebyCoeff diff2(const ComplexChebyCoeff& f);
ComplexChebyCoeff diff(const ComplexChebyCoeff& f, int n);
void integrate(const ComplexChebyCoeff& df, ComplexChebyCoeff& f);
ComplexChebyCoeff integrate(const ComplexChebyCoeff& df);
std::ostream& operator<<(std::ostream& os, const ComplexChebyCoeff& f);

inline ChebyCoeff& Re(ComplexChebyCoeff& f) { return f.re; }
inline ChebyCoeff& Im(ComplexChebyCoeff& f) { return f.im; }
inline const ChebyCoeff& Re(const ComplexChebyCoeff& f) { return f.re; }
inline const ChebyCoeff& Im(const ComplexChebyCoeff& f) { return f.im; }

Real L2Norm2(const ComplexChebyCoeff& f, bool normalize = true);
Real L2Dist2(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, bool normalize = true);
Real L2Norm(const ComplexChebyCoeff& f, bool normalize = true);
Real L2Dist(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, bool normalize = true);
Complex L2InnerProduct(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, bool normalize = true);
Real chebyNorm2(const ComplexChebyCoeff& f, bool normalize = true);
Real chebyDist2(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, bool normalize = true);
Real chebyNorm(const ComplexChebyCoeff& f, bool normalize = true);
Real chebyDist(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, bool normalize = true);
Complex chebyInnerProduct(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, bool normalize = true);
Real norm2(const ComplexChebyCoeff& f, NormType n, bool normalize = true);
Real norm(const ComplexChebyCoeff& f, NormType n, bool normalize = true);
Real dist2(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, NormType n, bool normalize = true);
Real dist(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, NormType n, bool normalize = true);
Complex innerProduct(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, NormType n, bool normalize = true);
Real L1Norm(const ComplexChebyCoeff& f, bool normalize = true);
Real L1Dist(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g, bool normalize = true);
Real LinfNorm(const ComplexChebyCoeff& f);
Real LinfDist(const ComplexChebyCoeff& f, const ComplexChebyCoeff& g);

}  // namespace chflow
#endif
