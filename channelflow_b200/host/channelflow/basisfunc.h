// BasisFunc: the y-profiles of one Fourier mode (kx,kz) of an Nd-component field -- a tuple of ComplexChebyCoeff with its
// wavenumbers and box (reference channelflow/basisfunc.h:25-215).  Host object (O(Ny)); used to move single modes in and
// out of FlowFields (eigenfunctions, test fields).  The Galerkin basis construction of the reference header is outside
// this package's scope.
#ifndef CFB200_BASISFUNC_H
#define CFB200_BASISFUNC_H
#include <vector>

#include "cfbasics/cfarray.h"
#include "cfbasics/mathdefs.h"
#include "channelflow/chebyshev.h"

namespace chflow {

inline int i3j(int i, int j) { return 3 * i + j; }

class BasisFunc {
   public:
    BasisFunc();
    BasisFunc(int Nd, int Ny, int kx, int kz, Real Lx, Real Lz, Real a, Real b, fieldstate s = Spectral);
    BasisFunc(int Ny, int kx, int kz, Real Lx, Real Lz, Real a, Real b, fieldstate s = Spectral);  // Nd = 3
    BasisFunc(const std::string& filebase);  // filebase.asc: "% Nd Ny kx kz Lx Lz a b state", then Ny rows of (re im) x Nd
    BasisFunc(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, const ComplexChebyCoeff& w, int kx, int kz, Real Lx, Real Lz);

    void save(const std::string& filebase, fieldstate s = Physical) const;
    void binaryDump(std::ostream& os) const;
    void binaryLoad(std::istream& is);

    int Nd() const { return Nd_; }
    int Ny() const { return Ny_; }
    int kx() const { return kx_; }
    int kz() const { return kz_; }
    Real Lx() const { return Lx_; }
    Real Lz() const { return Lz_; }
    Real a() const { return a_; }
    Real b() const { return b_; }
    fieldstate state() const { return state_; }

    void reconfig(const BasisFunc& f);
    void resize(int Ny);
    void setBounds(Real Lx, Real Lz, Real a, Real b);
    void setkxkz(int kx, int kz) { kx_ = kx; kz_ = kz; }
    void setState(fieldstate s);
    void setToZero();
    void conjugate();  // kx, kz -> -kx, -kz and complex-conjugate profiles
    void fill(const BasisFunc& f);

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

    const ComplexChebyCoeff& u() const { return u_[0]; }
    const ComplexChebyCoeff& v() const { return u_[1]; }
    const ComplexChebyCoeff& w() const { return u_[2]; }
    ComplexChebyCoeff& u() { return u_[0]; }
    ComplexChebyCoeff& v() { return u_[1]; }
    ComplexChebyCoeff& w() { return u_[2]; }
    const ComplexChebyCoeff& operator[](int i) const { return u_[i]; }
    ComplexChebyCoeff& operator[](int i) { return u_[i]; }

    bool geomCongruent(const BasisFunc& f) const;
    bool congruent(const BasisFunc& f) const;
    bool interoperable(const BasisFunc& f) const;

    BasisFunc& operator*=(Real c);
    BasisFunc& operator*=(Complex c);
    BasisFunc& operator+=(const BasisFunc& g);
    BasisFunc& operator-=(const BasisFunc& g);

   private:
    int Nd_, Ny_, kx_, kz_;
    Real Lx_, Lz_, a_, b_;
    fieldstate state_;
    std::vector<ComplexChebyCoeff> u_;
};

BasisFunc conjugate(const BasisFunc& f);
Real L2Norm(const BasisFunc& f, bool normalize = true);
Real L2Norm2(const BasisFunc& f, bool normalize = true);
Real L2Dist(const BasisFunc& f, const BasisFunc& g, bool normalize = true);
Real L2Dist2(const BasisFunc& f, const BasisFunc& g, bool normalize = true);
Complex L2InnerProduct(const BasisFunc& f, const BasisFunc& g, bool normalize = true);
Real divNorm(const BasisFunc& f, bool normalize = true);
Real divNorm2(const BasisFunc& f, bool normalize = true);
Real bcNorm(const BasisFunc& f, bool normalize = true);
Real bcNorm2(const BasisFunc& f, bool normalize = true);

}  // namespace chflow
#endif
