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

   protected:
    std::shared_ptr<NSE> main_nse_, init_nse_;
    DNSAlgorithm* main_algorithm_ = nullptr;
    DNSAlgorithm* init_algorithm_ = nullptr;

    DNSAlgorithm* newAlgorithm(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags);
    void build(const std::vector<FlowField>& fields, const std::vector<ChebyCoeff>* base, DNSFlags flags);
};

}  // namespace chflow
#endif
