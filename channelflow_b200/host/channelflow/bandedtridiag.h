// chflow::BandedTridiag -- the bordered tridiagonal matrix of the Chebyshev tau Helmholtz problem (reference
// channelflow/bandedtridiag.h:29-117): dense first row + tridiagonal rest.  The element accessors work on a host copy of
// the reference's storage scheme (a_[4M-2]: the first row stored reversed in front of the (up, diag, lo) triplets, so that
// band(j), diag(i), updiag(i), lodiag(i) are the same addresses as in the reference); ULdecomp / ULsolve* / multiply*
// run on the device (cfgpu_tridiag, csrc/tau.cu:tridiag_kernel).
#ifndef CHANNELFLOW_BANDEDTRIDIAG_H
#define CHANNELFLOW_BANDEDTRIDIAG_H

#include <cassert>
#include <string>
#include <vector>

#include "cfbasics/cfvector.h"
#include "cfbasics/mathdefs.h"

namespace chflow {

typedef double Real;

class BandedTridiag {
   public:
    BandedTridiag() = default;
    explicit BandedTridiag(int M);
    BandedTridiag(const BandedTridiag& A);  // like the reference: the copy is un-factorised
    explicit BandedTridiag(const std::string& filebase);
    BandedTridiag& operator=(const BandedTridiag& A) = default;
    bool operator==(const BandedTridiag& A) const;
    int numrows() const { return M_; }

    Real& band(int j) { assert(j >= 0 && j < M_); return a_[M_ - 1 - j]; }           // A[0,j]
    Real& diag(int i) { assert(i >= 0 && i < M_); return a_[M_ - 1 + 3 * i]; }       // A[i,i]
    Real& updiag(int i) { assert(i >= 0 && i < M_); return a_[M_ - 2 + 3 * i]; }     // A[i,i+1]
    Real& lodiag(int i) { assert(i >= 0 && i < M_); return a_[M_ + 3 * i]; }         // A[i,i-1]
    const Real& band(int j) const { assert(j >= 0 && j < M_); return a_[M_ - 1 - j]; }
    const Real& diag(int i) const { assert(i >= 0 && i < M_); return a_[M_ - 1 + 3 * i]; }
    const Real& updiag(int i) const { assert(i >= 0 && i < M_); return a_[M_ - 2 + 3 * i]; }
    const Real& lodiag(int i) const { assert(i >= 0 && i < M_); return a_[M_ + 3 * i]; }
    Real& elem(int i, int j) { return a_[index(i, j)]; }
    const Real& elem(int i, int j) const { return a_[index(i, j)]; }

    void ULdecomp();  // no pivoting
    void ULsolve(Vector& b) const { ULsolveStrided(b, 0, 1); }
    void multiply(const Vector& x, Vector& b) const { multiplyStrided(x, b, 0, 1); }
    void ULsolveStrided(Vector& b, int offset, int stride) const;
    void multiplyStrided(const Vector& x, Vector& b, int offset, int stride) const;

    void print() const;
    void ULprint() const;
    void test() const;
    void save(const std::string& filebase) const;  // rows "i j Aij"

   private:
    int M_ = 0;
    std::vector<Real> a_, invdiag_;
    bool UL_ = false;
    int index(int i, int j) const {
        assert(i == 0 || (i >= 0 && i < M_ && j >= 0 && j < M_ && (i - j <= 1 && j - i <= 1)));
        return i == 0 ? M_ - 1 - j : M_ - 1 + 3 * i + (i - j);  // j = i+1 -> -1 (up), j = i-1 -> +1 (lo)
    }
};

}  // namespace chflow
#endif
