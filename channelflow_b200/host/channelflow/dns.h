// chflow::DNS -- facade that picks the time-stepping algorithm, initialises multistep schemes with a one-step
// scheme and advances the fields.  Same public surface as the reference's channelflow/dns.h:30-97.
#ifndef CFB200_DNS_H
#define CFB200_DNS_H
#include <memory>
#include <vector>

#include "channelflow/dnsalgo.h"

namespace chflow {

class DNS {
   public:
    DNS();
    DNS(const DNS& dns);
    DNS(const std::vector<FlowField>& fields, const DNSFlags& flags);
    DNS(const std::vector<FlowField>& fields, const std::vector<ChebyCoeff>& base, const DNSFlags& flags);
    virtual ~DNS();
    DNS& operator=(const DNS& dns);

    void advance(std::vector<FlowField>& fields, int nSteps = 1);
    void project();
    // true pressure p <-> the modified pressure q = p + 1/2 |u + Ubase|^2 of the rotational form (dns.cpp:372-444)
    void uq2p(FlowField u, FlowField q, FlowField& p) const;
    void up2q(FlowField u, FlowField p, FlowField& q) const;

    virtual void reset_dt(Real dt);
    virtual void reset_time(Real t);
    virtual void reset_gradp(Real dPdx, Real dPdz);
    virtual void reset_bulkv(Real Ubulk, Real Wbulk);

    bool push(const std::vector<FlowField>& fields);
    virtual bool full() const;
    virtual int order() const;
    virtual int Ninitsteps() const;

    Real nu() const;
    virtual Real dt() const;
    virtual Real CFL(FlowField& u) const;
    virtual Real time() const;
    virtual Real dPdx() const;
    Real dPdz() const;
    virtual Real Ubulk() const;
    Real Wbulk() const;
    virtual Real dPdxRef() const;
    Real dPdzRef() const;
    virtual Real UbulkRef() const;
    Real WbulkRef() const;
    virtual const ChebyCoeff& Ubase() const;
    virtual const ChebyCoeff& Wbase() const;
    const DNSFlags& flags() const;
    virtual TimeStepMethod timestepping() const;
    // apply sigma[m] to every copy of field m the algorithms hold (dns.cpp:175-180, dnsalgo.cpp:264-272)
    void operator*=(const std::vector<FieldSymmetry>& sigma);

   protected:
    std::shared_ptr<NSE> main_nse_, init_nse_;
    DNSAlgorithm* main_algorithm_ = nullptr;
    DNSAlgorithm* init_algorithm_ = nullptr;

    DNSAlgorithm* newAlgorithm(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags);
    void build(const std::vector<FlowField>& fields, const std::vector<ChebyCoeff>* base, DNSFlags flags);
};

// ---------------------------------------------------------------------------------------------------------------------
// Poincare sections (dns.h:157-216, dns.cpp:450-704).  A PoincareCondition is a scalar function h of the velocity field, the
// section is h(u) == 0.  DNSPoincare::advanceToSection advances (u, q) by nSteps time steps; when h changes sign over the
// stride (in the requested direction, after Tmin) the stride is integrated again step by step from the saved start and
// the crossing is located on the quadratic interpolant in time through the three steps that bracket it, by a secant-
// started Newton iteration on h.  Every field operation (h itself, the interpolant, the symmetry map back into the
// fundamental domain) runs on the device.
//
// Note on the reference: in its current form dns.cpp:526-528 advances a *copy* of (u, p) (the std::vector built from them),
// so its u never moves and no crossing can be detected.  This class implements the documented behaviour -- the caller's
// fields are advanced -- which is also what cfdsi.cpp:779-830 relies on.
class PoincareCondition {
   public:
    virtual ~PoincareCondition() {}
    virtual Real operator()(const FlowField& u) = 0;
};

// h(u) = (u, estar) - (ustar, estar)
class PlaneIntersection : public PoincareCondition {
   public:
    PlaneIntersection() {}
    PlaneIntersection(const FlowField& ustar, const FlowField& estar);
    Real operator()(const FlowField& u) override;

   private:
    FlowField estar_;
    Real cstar_ = 0;
};

// h(u) = I - D = wallshear(u) - dissipation(u)
class DragDissipation : public PoincareCondition {
   public:
    Real operator()(const FlowField& u) override;
};

class DNSPoincare : public DNS {
   public:
    DNSPoincare();
    DNSPoincare(FlowField& u, PoincareCondition* h, const DNSFlags& flags);
    // e[n], sigma[n]: the fundamental domain is (u, e[n]) >= 0 for all n, sigma[n] maps a field with (u, e[n]) < 0 back
    DNSPoincare(FlowField& u, const cfarray<FlowField>& e, const cfarray<FieldSymmetry>& sigma, PoincareCondition* h,
                const DNSFlags& flags);

    bool advanceToSection(FlowField& u, FlowField& q, int nSteps, int crosssign = 0, Real Tmin = 0, Real epsilon = 1e-13);

    const FlowField& ucrossing() const { return ucrossing_; }
    const FlowField& pcrossing() const { return pcrossing_; }
    Real hcrossing() const { return hcrossing_; }  // h at the crossing found
    Real tcrossing() const { return tcrossing_; }
    int scrossing() const { return scrossing_; }   // sign of dh/dt at the crossing
    Real hcurrent() const { return hcurrent_; }    // h(u) after the last stride

   private:
    cfarray<FlowField> e_;
    cfarray<FieldSymmetry> sigma_;
    PoincareCondition* h_ = nullptr;
    FlowField ucrossing_, pcrossing_;
    Real tcrossing_ = 0, hcrossing_ = 0, hcurrent_ = 0, t0_ = 0;
    int scrossing_ = 0;
};

}  // namespace chflow
#endif
