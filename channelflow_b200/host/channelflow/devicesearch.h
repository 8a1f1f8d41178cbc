// Newton-Krylov-hookstep search for invariant solutions with every vector in HBM -- the device-resident counterpart of the
// reference's findsoln stack for the fixed-(T, sigma) case:
//   cfDSI::eval      (channelflow/cfdsi.cpp:52-61, 686-775)   G(x) = (sigma f^T(u(x)) - u(x)) [/T]   -> DeviceDSI::eval
//   GMRES            (nsolver/gmres.cpp:37-102)                Arnoldi + modified Gram-Schmidt        -> cfgpu_vec_dot / axpy
//   Newton-hookstep  (nsolver/newtonalgorithm.cpp:334-790)     trust-region step in the Krylov space  -> hookstepSearch
// The state vector x is the field2vector packing of the velocity field (flowfield.cpp:4448-4760, device kernels in
// csrc/vecpack.cu); one evaluation of G is one DNS integration over T (CUDA-graph replay on the small grids these searches
// run on).  Only the small Hessenberg problem (at most Ngmres+1 by Ngmres) lives on the host.  The search flags keep the
// reference's names and defaults (nsolver/newtonalgorithm.h, programs/findsoln.cpp).  The phase shifts ax, az of sigma can be
// unknowns of the search (-xrel, -zrel of findsoln: cfdsi.cpp:531-572 appends them to the state vector and the Newton step is
// kept orthogonal to the translation directions du/dx, du/dz); the integration time T stays fixed (no -orb search).
#ifndef CHANNELFLOW_DEVICESEARCH_H
#define CHANNELFLOW_DEVICESEARCH_H

#include <iostream>
#include <vector>

#include "channelflow/dns.h"
#include "channelflow/flowfield.h"
#include "channelflow/symmetry.h"

namespace chflow {

struct DeviceSearchFlags {
    Real epsSearch = 1e-13;  // stop when L2Norm(G(u)) < epsSearch
    Real epsKrylov = 1e-14;  // stop the Arnoldi iteration when the new Krylov vector is this small
    Real epsDx = 1e-7;       // relative size of the finite-difference step of DG
    Real epsGMRES = 1e-3;    // GMRES target for |DG dx + G| / |G|
    Real epsGMRESf = 0.05;   // accept the last iterate below this
    bool centdiff = false;   // centred differences for DG
    int Nnewton = 20, Ngmres = 120, Nhook = 20;
    Real delta = 0.01, deltaMin = 1e-12, deltaMax = 0.1, deltaFuzz = 1e-6;  // trust-region radius (2-norm of the step vector)
    Real lambdaMin = 0.2, lambdaMax = 1.5, lambdaRequiredReduction = 0.5;
    Real improvReq = 1e-3, improvOk = 0.10, improvGood = 0.75, improvAcc = 0.10;
    std::ostream* logstream = &std::cout;
};

// G(x) for a (relative) equilibrium / periodic orbit with fixed T and sigma
class DeviceDSI {
   public:
    DeviceDSI(const FlowField& u, const DNSFlags& flags, const TimeStep& dt, const FieldSymmetry& sigma, Real T, bool Tnormalize,
              bool xrelative = false, bool zrelative = false);
    long size() const { return size_; }
    bool xrelative() const { return xrel_; }
    bool zrelative() const { return zrel_; }
    const FieldSymmetry& sigma() const { return sigma_; }
    void setShifts(Real ax, Real az) { sigma_ = FieldSymmetry(sigma_.sx(), sigma_.sy(), sigma_.sz(), ax, az, sigma_.s()); }
    void tangent(const DeviceVector& x, int dir, DeviceVector& t) const;  // packing of du/dx (dir 0) or du/dz (dir 1)
    void makeVector(const FlowField& u, DeviceVector& x) const { field2vector(u, x); }
    void extractVector(const DeviceVector& x, FlowField& u) const;
    void eval(const DeviceVector& x, DeviceVector& Gx);   // one DNS integration
    void f(const FlowField& u, FlowField& fu);            // f^T(u)
    void G(const FlowField& u, FlowField& Gu);            // sigma f^T(u) - u [/T]
    Real residual(const DeviceVector& Gx) const;          // L2Norm of the field G packs
    int evaluations() const { return fcount_; }
    Real CFL() const { return CFL_; }
    Real steps_per_eval() const { return steps_; }

   private:
    FlowField proto_;
    DNSFlags flags_;
    TimeStep dt_;
    FieldSymmetry sigma_;
    Real T_;
    bool Tnormalize_;
    bool xrel_ = false, zrel_ = false;
    long size_;
    int fcount_ = 0;
    Real CFL_ = 0, steps_ = 0;
};

struct DeviceSearchResult {
    bool converged = false;
    int newtonSteps = 0, fevals = 0, gmresIterations = 0;
    Real residual = 0;                 // final L2Norm(G)
    std::vector<Real> history;         // L2Norm(G) after each accepted Newton-hookstep step (history[0]: initial guess)
};

// Newton-hookstep iteration on G(x) = 0 from the initial guess in x (overwritten by the best solution found); with
// dsi.xrelative()/zrelative() the shifts of dsi.sigma() are unknowns too and hold the solution's values on return
DeviceSearchResult hookstepSearch(DeviceDSI& dsi, DeviceVector& x, const DeviceSearchFlags& flags);

}  // namespace chflow
#endif
