// chflow::PoissonSolver / PressureSolver (reference channelflow/poissonsolver.h:21-104): lapl u = f with Dirichlet data, and
// the pressure of a velocity field from lapl p = -div N(u) with dp/dy = nu v_yy at the walls.  The solver objects only
// record the geometry: the per-mode Helmholtz operators are factorised inside the device kernel that uses them
// (cfgpu_poisson_solve, csrc/tau.cu:poisson_kernel -- one warp per Fourier mode), nothing is stored per mode.
#ifndef CHANNELFLOW_POISSONSOLVER_H
#define CHANNELFLOW_POISSONSOLVER_H

#include "cfbasics/cfvector.h"
#include "cfbasics/mathdefs.h"
#include "channelflow/chebyshev.h"
#include "channelflow/diffops.h"
#include "channelflow/dns.h"
#include "channelflow/flowfield.h"
#include "channelflow/helmholtz.h"

namespace chflow {

class PoissonSolver {
   public:
    PoissonSolver() = default;
    PoissonSolver(const FlowField& u);
    PoissonSolver(int Nx, int Ny, int Nz, int Nd, Real Lx, Real Lz, Real a, Real b, CfMPI* cfmpi = nullptr);

    void solve(FlowField& u, const FlowField& f) const;                       // u = 0 on the walls
    void solve(FlowField& u, const FlowField& f, const FlowField& bc) const;  // u = bc on the walls
    Real verify(const FlowField& u, const FlowField& f) const;
    Real verify(const FlowField& u, const FlowField& f, const FlowField& bc) const;

    bool geomCongruent(const FlowField& u) const;
    bool congruent(const FlowField& u) const;

   protected:
    int Mx_ = 0, My_ = 0, Mz_ = 0, Nz_ = 0, Nd_ = 0;
    Real Lx_ = 0, Lz_ = 0, a_ = 0, b_ = 0;
    void prepare(FlowField& u, const FlowField& f) const;
};

class PressureSolver : public PoissonSolver {
   public:
    PressureSolver() = default;
    PressureSolver(int Nx, int Ny, int Nz, Real Lx, Real Lz, Real a, Real b, const ChebyCoeff& U, const ChebyCoeff& W, Real nu,
                   Real Vsuck, NonlinearMethod nonl_method, CfMPI* cfmpi = nullptr);
    PressureSolver(const FlowField& u, Real nu, Real Vsuck, NonlinearMethod nonl_method);
    PressureSolver(const FlowField& u, const ChebyCoeff& U, const ChebyCoeff& W, Real nu, Real Vsuck, NonlinearMethod nonl_method);

    FlowField solve(const FlowField& u);
    void solve(FlowField& p, FlowField u);
    Real verify(const FlowField& p, const FlowField& u);

   private:
    ChebyCoeff U_, W_;
    FlowField nonl_, tmp_, div_nonl_;
    Real nu_ = 0, Vsuck_ = 0;
    NonlinearMethod nonl_method_ = Rotational;
    void minus_div_nonlinear(const FlowField& u);  // div_nonl_ = -div N(u)
};

}  // namespace chflow
#endif
