// Time-stepping algorithms of the DNS -- same class family and semantics as the reference's
// channelflow/dnsalgo.h:31-173 (DNSAlgorithm, MultistepDNS [CNFE1, SBDF1-4], RungeKuttaDNS [CNRK2],
// CNABstyleDNS [CNAB2, SMRK2]).  Each stage is: NSE::nonlinear (device pipeline), optional NSE::linear, and
// NSE::solve_lincomb, which accumulates the stage's right-hand side inside the batched tau-solve kernel.
#ifndef CFB200_DNSALGO_H
#define CFB200_DNSALGO_H
#include <memory>
#include <vector>

#include "channelflow/nse.h"

namespace chflow {

class DNSAlgorithm {
   public:
    DNSAlgorithm();
    DNSAlgorithm(const DNSAlgorithm& dns);
    DNSAlgorithm(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags);
    virtual ~DNSAlgorithm();

    virtual void advance(std::vector<FlowField>& fields, int nSteps = 1) = 0;
    virtual void project() {}  // project the state the algorithm holds onto the symmetric subspace of the flags
    virtual void operator*=(const std::vector<FieldSymmetry>&) {}  // map the stored history (dnsalgo.cpp:50, 264-272)
    cfarray<FieldSymmetry> symmetries(int ifield) const { return symmetries_[ifield]; }
    virtual void reset_dt(Real dt) = 0;
    virtual bool push(const std::vector<FlowField>& fields);
    virtual bool full() const;
    virtual DNSAlgorithm* clone(const std::shared_ptr<NSE>& nse) const = 0;

    void reset_time(Real t) { t_ = t; }
    void reset_nse(std::shared_ptr<NSE> nse) { nse_ = nse; }
    int order() const { return order_; }
    int Ninitsteps() const { return Ninitsteps_; }
    Real dt() const { return flags_.dt; }
    Real CFL(FlowField& u) const;
    Real time() const { return t_; }
    const DNSFlags& flags() const { return flags_; }
    TimeStepMethod timestepping() const { return flags_.timestepping; }

   protected:
    DNSFlags flags_;
    int order_ = 0, numfields_ = 0, Ninitsteps_ = 0;
    Real t_ = 0;
    std::vector<Real> lambda_t_;
    std::shared_ptr<NSE> nse_;
    std::vector<cfarray<FieldSymmetry>> symmetries_;  // per field (velocity, pressure)
    void tick() const;
    void endline() const;
};

class MultistepDNS : public DNSAlgorithm {
   public:
    MultistepDNS();
    MultistepDNS(const MultistepDNS& dns);
    MultistepDNS(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags);
    void advance(std::vector<FlowField>& fields, int nSteps = 1) override;
    void project() override;
    void operator*=(const std::vector<FieldSymmetry>& sigma) override;
    void reset_dt(Real dt) override;
    bool push(const std::vector<FlowField>& fields) override;
    bool full() const override { return countdown_ == 0; }
    DNSAlgorithm* clone(const std::shared_ptr<NSE>& nse) const override;

   protected:
    Real eta_ = 0;
    std::vector<Real> alpha_, beta_;
    std::vector<std::vector<FlowField>> fields_, nonlf_;  // history of (u,q) and of f = NL(u)
    int countdown_ = 0;
};

class RungeKuttaDNS : public DNSAlgorithm {
   public:
    RungeKuttaDNS();
    RungeKuttaDNS(const RungeKuttaDNS& dns);
    RungeKuttaDNS(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags);
    void advance(std::vector<FlowField>& fields, int nSteps = 1) override;
    void reset_dt(Real dt) override;
    DNSAlgorithm* clone(const std::shared_ptr<NSE>& nse) const override;

   protected:
    int Nsubsteps_ = 0;
    std::vector<FlowField> Qj1_, Qj_;
    std::vector<FlowField> lt_;  // linear term work field
    std::vector<Real> A_, B_, C_;
};

class CNABstyleDNS : public DNSAlgorithm {
   public:
    CNABstyleDNS();
    CNABstyleDNS(const CNABstyleDNS& dns);
    CNABstyleDNS(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags);
    void advance(std::vector<FlowField>& fields, int nSteps = 1) override;
    void reset_dt(Real dt) override;
    bool push(const std::vector<FlowField>& fields) override;
    bool full() const override { return full_; }
    DNSAlgorithm* clone(const std::shared_ptr<NSE>& nse) const override;

   protected:
    int Nsubsteps_ = 0;
    bool full_ = false;
    std::vector<FlowField> fj1_, fj_;
    std::vector<FlowField> lt_;  // linear term work field
    std::vector<Real> alpha_, beta_, gamma_, zeta_;
};

}  // namespace chflow
#endif
