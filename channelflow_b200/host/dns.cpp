// chflow::DNS facade (reference dns.cpp:22-163, 192-330).
#include "channelflow/dns.h"

namespace chflow {

DNS::DNS() {}

DNS::DNS(const DNS& d)
    : main_nse_(d.main_nse_ ? new NSE(*d.main_nse_) : nullptr), init_nse_(d.init_nse_ ? new NSE(*d.init_nse_) : nullptr),
      main_algorithm_(d.main_algorithm_ ? d.main_algorithm_->clone(main_nse_) : nullptr),
      init_algorithm_(d.init_algorithm_ ? d.init_algorithm_->clone(init_nse_) : nullptr) {}

void DNS::build(const std::vector<FlowField>& fields, const std::vector<ChebyCoeff>* base, DNSFlags flags) {
    auto make_nse = [&](const DNSFlags& f) {
        return std::shared_ptr<NSE>(base ? new NSE(fields, *base, f) : new NSE(fields, f));
    };
    main_nse_ = make_nse(flags);
    main_algorithm_ = newAlgorithm(fields, main_nse_, flags);
    if (!main_algorithm_->full() && flags.initstepping != flags.timestepping) {
        DNSFlags initflags = flags;
        initflags.timestepping = flags.initstepping;
        init_nse_ = make_nse(initflags);
        init_algorithm_ = newAlgorithm(fields, init_nse_, initflags);
        if (init_algorithm_->Ninitsteps() != 0)
            std::cerr << "DNS::DNS(fields, flags) :\n" << flags.initstepping << " can't initialize " << flags.timestepping
                      << " since it needs initialization itself.\n";
    }
}

DNS::DNS(const std::vector<FlowField>& fields, const DNSFlags& flags) { build(fields, nullptr, flags); }

DNS::DNS(const std::vector<FlowField>& fields, const std::vector<ChebyCoeff>& base, const DNSFlags& flags_) {
    DNSFlags flags = flags_;
    flags.baseflow = ArbitraryBase;
    build(fields, &base, flags);
}

DNS::~DNS() {
    delete main_algorithm_;
    delete init_algorithm_;
}

DNS& DNS::operator=(const DNS& d) {
    if (this == &d) return *this;
    delete main_algorithm_;
    delete init_algorithm_;
    main_nse_ = d.main_nse_ ? std::shared_ptr<NSE>(new NSE(*d.main_nse_)) : nullptr;
    init_nse_ = d.init_nse_ ? std::shared_ptr<NSE>(new NSE(*d.init_nse_)) : nullptr;
    main_algorithm_ = d.main_algorithm_ ? d.main_algorithm_->clone(main_nse_) : nullptr;
    init_algorithm_ = d.init_algorithm_ ? d.init_algorithm_->clone(init_nse_) : nullptr;
    return *this;
}

DNSAlgorithm* DNS::newAlgorithm(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags) {
    switch (flags.timestepping) {
        case CNFE1: case SBDF1: case SBDF2: case SBDF3: case SBDF4:
            return new MultistepDNS(fields, nse, flags);
        case CNRK2:
            return new RungeKuttaDNS(fields, nse, flags);
        case SMRK2: case CNAB2:
            return new CNABstyleDNS(fields, nse, flags);
        default:
            std::cerr << "DNS::newAlgorithm : algorithm " << flags.timestepping << " is unimplemented" << std::endl;
    }
    return nullptr;
}

void DNS::advance(std::vector<FlowField>& fields, int Nsteps) {
    assert(main_algorithm_);
    if (!main_algorithm_->full() && !init_algorithm_)
        cferror("DNS::advance(u,q,Nsteps) : the main algorithm is uninitialized and the initialization algorithm is not set.");
    for (size_t j = 1; j < fields.size(); ++j)
        if (!fields[j].geomCongruent(fields[0]))
            fields[j].resize(fields[0].Nx(), fields[0].Ny(), fields[0].Nz(), fields[j].Nd(), fields[0].Lx(), fields[0].Lz(),
                             fields[0].a(), fields[0].b(), fields[0].cfmpi());
    int n = 0;
    // every symmetryprojectioninterval time units the state is projected back onto the symmetric subspace (dns.cpp:152-156)
    if (flags().symmetries.length() > 0 && (int)(main_algorithm_->time()) % (int)(flags().symmetryprojectioninterval) == 0) {
        main_algorithm_->project();
        for (size_t j = 0; j < fields.size(); ++j) fields[j].project(main_algorithm_->symmetries((int)j));
    }
    while (!main_algorithm_->full() && n < Nsteps) {
        main_algorithm_->push(fields);
        init_algorithm_->advance(fields, 1);
        ++n;
    }
    main_algorithm_->advance(fields, Nsteps - n);
}

void DNS::project() {}

// q = p + s/2 (u + U e_x + W e_z)^2 pointwise; everything on the device: base profiles are added to the (0,0) mode, the
// kinetic energy is one pointwise kernel
static void modified_pressure(const DNS& dns, FlowField& u, const FlowField& in, FlowField& out, Real sign) {
    u.makeSpectral();
    std::vector<ChebyCoeff> UW = {dns.Ubase(), dns.Wbase()};
    if (UW[0].length() == 0) UW[0] = ChebyCoeff(u.Ny(), u.a(), u.b(), Spectral);
    if (UW[1].length() == 0) UW[1] = ChebyCoeff(u.Ny(), u.a(), u.b(), Spectral);
    UW[0].makeSpectral();
    UW[1].makeSpectral();
    u += UW;
    FlowField e;
    energy(u, e);  // 1/2 |u_tot|^2, spectral on return
    FlowField tmp(in);
    tmp.makeSpectral();
    tmp.add(sign, e);
    tmp.makeState(in.xzstate(), in.ystate());
    out = tmp;
}
void DNS::uq2p(FlowField u, FlowField q, FlowField& p) const {
    if (flags().nonlinearity != Rotational) { p = q; return; }
    modified_pressure(*this, u, q, p, -1.0);
}
void DNS::up2q(FlowField u, FlowField p, FlowField& q) const {
    if (flags().nonlinearity != Rotational) { q = p; return; }
    modified_pressure(*this, u, p, q, 1.0);
}

void DNS::reset_dt(Real dt) {
    main_algorithm_->reset_dt(dt);
    if (init_algorithm_) init_algorithm_->reset_dt(dt);
}
void DNS::reset_time(Real t) {
    if (init_algorithm_) init_algorithm_->reset_time(t);
    if (main_algorithm_) main_algorithm_->reset_time(t);
}
void DNS::reset_gradp(Real dPdx, Real dPdz) {
    if (init_nse_) init_nse_->reset_gradp(dPdx, dPdz);
    if (main_nse_) main_nse_->reset_gradp(dPdx, dPdz);
}
void DNS::reset_bulkv(Real Ubulk, Real Wbulk) {
    if (init_nse_) init_nse_->reset_bulkv(Ubulk, Wbulk);
    if (main_nse_) main_nse_->reset_bulkv(Ubulk, Wbulk);
}
bool DNS::push(const std::vector<FlowField>& fields) { return main_algorithm_ ? main_algorithm_->push(fields) : false; }
bool DNS::full() const { return main_algorithm_ ? main_algorithm_->full() : false; }
int DNS::order() const { return main_algorithm_ ? main_algorithm_->order() : (init_algorithm_ ? init_algorithm_->order() : 0); }
int DNS::Ninitsteps() const { return main_algorithm_ ? main_algorithm_->Ninitsteps() : 0; }
Real DNS::nu() const { return main_nse_ ? main_nse_->nu() : (init_nse_ ? init_nse_->nu() : 0.0); }
Real DNS::dt() const { return main_algorithm_ ? main_algorithm_->dt() : (init_algorithm_ ? init_algorithm_->dt() : 0.0); }
Real DNS::CFL(FlowField& u) const {
    return main_algorithm_ ? main_algorithm_->CFL(u) : (init_algorithm_ ? init_algorithm_->CFL(u) : 0.0);
}
Real DNS::time() const { return main_algorithm_ ? main_algorithm_->time() : (init_algorithm_ ? init_algorithm_->time() : 0.0); }
#define NSE_GETTER(name) \
    Real DNS::name() const { return main_nse_ ? main_nse_->name() : (init_nse_ ? init_nse_->name() : 0.0); }
NSE_GETTER(dPdx) NSE_GETTER(dPdz) NSE_GETTER(Ubulk) NSE_GETTER(Wbulk) NSE_GETTER(dPdxRef) NSE_GETTER(dPdzRef)
NSE_GETTER(UbulkRef) NSE_GETTER(WbulkRef)
const ChebyCoeff& DNS::Ubase() const { return main_nse_ ? main_nse_->Ubase() : init_nse_->Ubase(); }
const ChebyCoeff& DNS::Wbase() const { return main_nse_ ? main_nse_->Wbase() : init_nse_->Wbase(); }
const DNSFlags& DNS::flags() const { return main_algorithm_ ? main_algorithm_->flags() : init_algorithm_->flags(); }
TimeStepMethod DNS::timestepping() const {
    return main_algorithm_ ? main_algorithm_->timestepping() : init_algorithm_->timestepping();
}

}  // namespace chflow
