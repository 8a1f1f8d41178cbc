// chflow::TauSolver -- the Kleiser-Schumann influence-matrix solver of one Fourier mode (reference
// channelflow/tausolver.h:30-115; Canuto & Hussaini 7.3.18-20):
//      nu u'' - lambda u - grad P = -R,   div u = 0,   u(+-1) = 0
// The object records the mode and the operator; solve() runs the batched device tau solver on that one mode
// (cfgpu_tausolve_mode: the same tau_setup / tau_solve kernels the DNS uses for all modes at once).  verify() measures the
// residuals of a given solution with the host Chebyshev calculus, as a check that is independent of the solver.
#ifndef CHANNELFLOW_TAUSOLVER_H
#define CHANNELFLOW_TAUSOLVER_H

#include <string>

#include "cfbasics/cfvector.h"
#include "cfbasics/mathdefs.h"
#include "channelflow/chebyshev.h"
#include "channelflow/helmholtz.h"

namespace chflow {

Real divcheck(std::string& label, int kx, int kz, Real Lx, Real Lz, const ComplexChebyCoeff& u, const ComplexChebyCoeff& v,
              const ComplexChebyCoeff& w, bool verbose = false);

class TauSolver {
   public:
    TauSolver() = default;
    TauSolver(int kx, int kz, Real Lx, Real Lz, Real a, Real b, Real lambda, Real nu, int Ny, bool tauCorrection = true);

    void solve(ComplexChebyCoeff& u, ComplexChebyCoeff& v, ComplexChebyCoeff& w, ComplexChebyCoeff& P,
               const ComplexChebyCoeff& Rx, const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz) const;
    Real verify(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, const ComplexChebyCoeff& w, const ComplexChebyCoeff& P,
                const ComplexChebyCoeff& Rx, const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz, bool verbose = false) const;

    // the (0,0) mode with unknown mean pressure gradients and prescribed bulk velocities
    void solve(ComplexChebyCoeff& u, ComplexChebyCoeff& v, ComplexChebyCoeff& w, ComplexChebyCoeff& P, Real& dPdx, Real& dPdz,
               const ComplexChebyCoeff& Rx, const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz, Real umean, Real wmean) const;
    Real verify(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, const ComplexChebyCoeff& w, const ComplexChebyCoeff& P,
                Real dPdx, Real dPdz, const ComplexChebyCoeff& Rx, const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz,
                Real umean, Real wmean, bool verbose = false) const;

    // Stages of the reference's host algorithm ("not meant to be used alone", tausolver.h:59-66).  The device solver fuses
    // them into one kernel and exposes no per-stage state: these report an error.
    void influenceCorrection(ChebyCoeff& P, ChebyCoeff& v) const;
    void solve_P_and_v(ChebyCoeff& P, ChebyCoeff& v, const ChebyCoeff& r, const ChebyCoeff& Ry, Real& sigmaNb1, Real& sigmaNb) const;
    Real verify_P_and_v(const ChebyCoeff& P, const ChebyCoeff& v, const ChebyCoeff& r, const ChebyCoeff& Ry, Real sigmaNb1,
                        Real sigmaNb, bool verbose = false) const;

    int kx() const { return kx_; }
    int kz() const { return kz_; }
    Real lambda() const { return lambda_; }
    Real nu() const { return nu_; }

   private:
    int N_ = 0, kx_ = 0, kz_ = 0;
    Real Lx_ = 0, Lz_ = 0, a_ = 0, b_ = 0, lambda_ = 0, nu_ = 0;
    bool tauCorrection_ = true;
    void device_solve(ComplexChebyCoeff& u, ComplexChebyCoeff& v, ComplexChebyCoeff& w, ComplexChebyCoeff& P, const ComplexChebyCoeff& Rx,
                      const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz, int constraint, Real umean, Real wmean, Real* dPd) const;
};

}  // namespace chflow
#endif
