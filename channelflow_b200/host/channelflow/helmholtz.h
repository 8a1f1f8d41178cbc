// chflow::HelmholtzSolver -- nu u'' - lambda u = f on [a,b] with Dirichlet data, Chebyshev tau method (reference
// channelflow/helmholtz.h:26-61, Canuto & Hussaini 5.1.2).  The object only records the operator; every solve is one
// launch of the batched device solver (cfgpu_helmholtz_solve, csrc/tau.cu:helmholtz_batch_kernel), which builds the UL
// factors of the even/odd bordered tridiagonal systems in shared memory and solves one column per warp -- the same code
// path the time stepper's tau solve uses.
#ifndef CHANNELFLOW_HELMHOLTZ_H
#define CHANNELFLOW_HELMHOLTZ_H

#include "cfbasics/mathdefs.h"
#include "channelflow/bandedtridiag.h"
#include "channelflow/chebyshev.h"

namespace chflow {

class HelmholtzSolver {
   public:
    HelmholtzSolver() = default;
    HelmholtzSolver(int numberModes, Real a, Real b, Real lambda, Real nu = 1.0);

    // Dirichlet data ua = u(a), ub = u(b)
    void solve(ChebyCoeff& u, const ChebyCoeff& f, Real ua, Real ub) const;
    void verify(const ChebyCoeff& u, const ChebyCoeff& f, Real ua, Real ub, bool verbose = false) const;
    Real residual(const ChebyCoeff& u, const ChebyCoeff& f, Real ua, Real ub) const;

    // nu u'' - lambda u - mu = f with mean(u) = umean: solves for u and the constant mu
    void solve(ChebyCoeff& u, Real& mu, const ChebyCoeff& f, Real umean, Real ua, Real ub) const;
    void verify(ChebyCoeff& u, Real& mu, const ChebyCoeff& f, Real umean, Real ua, Real ub) const;
    Real residual(const ChebyCoeff& u, Real mu, const ChebyCoeff& f, Real umean, Real ua, Real ub) const;

    Real lambda() const { return lambda_; }

   private:
    int nModes_ = 0;
    Real a_ = 0, b_ = 0, lambda_ = 0, nu_ = 0;
    // tau residual sums of nu u'' - lambda u - f: over the first N-2 modes, and over all of them
    void tau_residuals(const ChebyCoeff& u, const ChebyCoeff& f, Real& tau, Real& all) const;
};

}  // namespace chflow
#endif
