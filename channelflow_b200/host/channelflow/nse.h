// chflow::NSE -- the Navier-Stokes operator of the DNS (nonlinear term, linear term, implicit tau solve).
// Public surface of the reference's channelflow/nse.h:20-169; the work is done by the cfgpu_nse_* entry points
// (include/cfgpu.h).  One addition used by the time steppers: solve_lincomb(), which fuses the right-hand-side
// accumulation (FlowField::add passes in dnsalgo.cpp:217-224) into the solve kernel.
#ifndef CFB200_NSE_H
#define CFB200_NSE_H
#include <memory>
#include <vector>

#include "channelflow/diffops.h"
#include "channelflow/dnsflags.h"
#include "channelflow/flowfield.h"

namespace chflow {

void navierstokesNL(const FlowField& u, ChebyCoeff Ubase, ChebyCoeff Wbase, FlowField& f, FlowField& tmp, DNSFlags& flags);

class NSE {
   public:
    NSE();
    NSE(const NSE& nse);
    NSE(const std::vector<FlowField>& fields, const DNSFlags& flags);
    NSE(const std::vector<FlowField>& fields, const std::vector<ChebyCoeff>& base, const DNSFlags& flags);
    virtual ~NSE();

    virtual void nonlinear(const std::vector<FlowField>& infields, std::vector<FlowField>& outfields);
    virtual void linear(const std::vector<FlowField>& infields, std::vector<FlowField>& outfields);
    virtual void solve(std::vector<FlowField>& outfields, const std::vector<FlowField>& infields, const int i = 0);
    // outfields = solution of the implicit problem with rhs = sum_j coef[j] * terms[j]   (extension, see above)
    virtual void solve_lincomb(std::vector<FlowField>& outfields, const std::vector<Real>& coef,
                               const std::vector<const FlowField*>& terms, const int i = 0);

    virtual void reset_lambda(const std::vector<Real> lambda_t);
    virtual std::vector<FlowField> createRHS(const std::vector<FlowField>& fields) const;
    virtual std::vector<cfarray<FieldSymmetry>> createSymmVec() const;  // symmetries confining (u, q) to a subspace

    int taskid() const { return 0; }
    void reset_gradp(Real dPdx, Real dPdz);
    void reset_bulkv(Real Ubulk, Real Wbulk);

    int Nx() const { return Nx_; }
    int Ny() const { return My_; }
    int Nz() const { return Nz_; }
    Real Lx() const { return Lx_; }
    Real Lz() const { return Lz_; }
    Real a() const { return a_; }
    Real b() const { return b_; }
    Real nu() const { return flags_.nu; }
    Real dPdx() const;
    Real dPdz() const;
    Real Ubulk() const { return UbulkAct_; }
    Real Wbulk() const { return WbulkAct_; }
    Real dPdxRef() const { return dPdxRef_; }
    Real dPdzRef() const { return dPdzRef_; }
    Real UbulkRef() const { return UbulkRef_; }
    Real WbulkRef() const { return WbulkRef_; }
    virtual const ChebyCoeff& Ubase() const { return Ubase_; }
    virtual const ChebyCoeff& Wbase() const { return Wbase_; }
    const DNSFlags& flags() const { return flags_; }

    Real CFLfactor(const FlowField& u) const;  // max (u_i+U_i)/dx_i on the device
    static Real cflfactor_of(const FlowField& u, const ChebyCoeff& U, const ChebyCoeff& W, const DNSFlags& flags);

   protected:
    std::vector<Real> lambda_t_;
    DNSFlags flags_;
    int Nx_ = 0, My_ = 0, Nz_ = 0;
    Real Lx_ = 0, Lz_ = 0, a_ = 0, b_ = 0;
    Real dPdxRef_ = 0, dPdzRef_ = 0, UbulkRef_ = 0, UbulkAct_ = 0, UbulkBase_ = 0, WbulkRef_ = 0, WbulkAct_ = 0, WbulkBase_ = 0;
    mutable Real dPdxAct_ = 0, dPdzAct_ = 0;
    mutable bool dPd_on_device_ = false;
    ChebyCoeff Ubase_, Wbase_;
    cfgpu_nse dev_ = nullptr;

    void createCFBaseFlow();
    void initCFConstraint(const FlowField& u);
    void create_device();
    void push_constraint();
};

Real viscosity(Real Reynolds, VelocityScale vscale, MeanConstraint constraint, Real dPdx, Real Ubulk, Real Uwall, Real h);
ChebyCoeff laminarProfile(Real nu, MeanConstraint constraint, Real dPdx, Real Ubulk, Real Vsuck, Real a, Real b, Real ua,
                          Real ub, int Ny);
ChebyCoeff laminarProfile(const DNSFlags& flags, Real a, Real b, int Ny);

}  // namespace chflow
#endif
