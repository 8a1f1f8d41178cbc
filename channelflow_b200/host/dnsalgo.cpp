// Time steppers; scheme constants and stage formulas as in the reference's dnsalgo.cpp
// (SBDF tables :106-150, CNRK2 :358-374, CNAB2/SMRK2 :515-555, stage updates :202-253, :419-470, :641-697).
#include "channelflow/dnsalgo.h"

namespace chflow {

DNSAlgorithm::DNSAlgorithm() {}
DNSAlgorithm::DNSAlgorithm(const DNSAlgorithm& d)
    : flags_(d.flags_), order_(d.order_), numfields_(d.numfields_), Ninitsteps_(d.Ninitsteps_), t_(d.t_),
      lambda_t_(d.lambda_t_), nse_(d.nse_), symmetries_(d.symmetries_) {}
DNSAlgorithm::DNSAlgorithm(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags)
    : flags_(flags), numfields_((int)fields.size()), t_(flags.t0), nse_(nse), symmetries_(nse->createSymmVec()) {}
DNSAlgorithm::~DNSAlgorithm() {}
bool DNSAlgorithm::push(const std::vector<FlowField>&) { return true; }
bool DNSAlgorithm::full() const { return true; }

Real DNSAlgorithm::CFL(FlowField& u) const {
    Real cfl = nse_->CFLfactor(u);
    cfl *= flags_.dealias_xz() ? 2.0 * pi / 3.0 * flags_.dt : pi * flags_.dt;
    return cfl;
}
void DNSAlgorithm::tick() const {
    if (flags_.verbosity == PrintTime || flags_.verbosity == PrintAll) *flags_.logstream << t_ << ' ' << std::flush;
    else if (flags_.verbosity == PrintTicks) *flags_.logstream << '.' << std::flush;
}
void DNSAlgorithm::endline() const {
    if (flags_.verbosity == PrintTime || flags_.verbosity == PrintAll || flags_.verbosity == PrintTicks)
        *flags_.logstream << std::endl;
}

// zero fields of the same shape and state (no device memory is touched until they are first written)
static std::vector<FlowField> zeros_like(const std::vector<FlowField>& fields) {
    std::vector<FlowField> z;
    z.reserve(fields.size());
    for (const auto& f : fields) {
        z.emplace_back(f.Nx(), f.Ny(), f.Nz(), f.Nd(), f.Lx(), f.Lz(), f.a(), f.b(), f.cfmpi(), f.xzstate(), f.ystate());
        z.back().setPadded(f.padded());
    }
    return z;
}

// =============================================================================================== multistep
MultistepDNS::MultistepDNS() {}
MultistepDNS::MultistepDNS(const MultistepDNS& d)
    : DNSAlgorithm(d), eta_(d.eta_), alpha_(d.alpha_), beta_(d.beta_), fields_(d.fields_), nonlf_(d.nonlf_),
      countdown_(d.countdown_) {}

MultistepDNS::MultistepDNS(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags)
    : DNSAlgorithm(fields, nse, flags) {
    switch (flags.timestepping) {
        case CNFE1:
        case SBDF1:
            order_ = 1; eta_ = 1.0; alpha_ = {-1.0}; beta_ = {1.0};
            break;
        case SBDF2:
            order_ = 2; eta_ = 1.5; alpha_ = {-2.0, 0.5}; beta_ = {2.0, -1.0};
            break;
        case SBDF3:
            order_ = 3; eta_ = 11.0 / 6.0; alpha_ = {-3.0, 1.5, -1.0 / 3.0}; beta_ = {3.0, -3.0, 1.0};
            break;
        case SBDF4:
            order_ = 4; eta_ = 25.0 / 12.0; alpha_ = {-4.0, 3.0, -4.0 / 3.0, 0.25}; beta_ = {4.0, -6.0, 4.0, -1.0};
            break;
        default:
            cferror("MultistepDNS: flags.timestepping is a non-multistepping algorithm");
    }
    lambda_t_ = {eta_ / flags_.dt};
    nse_->reset_lambda(lambda_t_);
    std::vector<FlowField> z = zeros_like(fields);
    fields_.assign(order_, z);
    nonlf_.assign(order_, z);
    Ninitsteps_ = order_ - 1;
    countdown_ = Ninitsteps_;
}

DNSAlgorithm* MultistepDNS::clone(const std::shared_ptr<NSE>& nse) const {
    DNSAlgorithm* c = new MultistepDNS(*this);
    c->reset_nse(nse);
    return c;
}

void MultistepDNS::reset_dt(Real dt) {
    flags_.dt = dt;
    lambda_t_ = {eta_ / flags_.dt};
    nse_->reset_lambda(lambda_t_);
    countdown_ = Ninitsteps_;
}

// CFGPU_GRAPH=0 never, 1 always (single GPU), unset: grids of at most 2^20 points, where a step is launch-latency bound
static bool graph_replay_wanted(const FlowField& u, int Nsteps) {
    const char* env = getenv("CFGPU_GRAPH");
    if (env && atoi(env) == 0) return false;
    int nranks = 1;
    cfgpu_comm_rank(cfgpu_context(), nullptr, &nranks);
    if (nranks != 1) return false;
    const bool small = (long)u.Nx() * u.Ny() * u.Nz() <= (1L << 20);
    return (env ? atoi(env) == 1 : small);
}

void MultistepDNS::advance(std::vector<FlowField>& fieldsn, int Nsteps) {
    const int J = order_ - 1;
    // The caller's fields become history slot 0 and come back at the end: O(1) handle exchanges instead of the
    // reference's two deep copies per call (dnsalgo.cpp:207,247); slot 0 is overwritten on entry of the next call anyway.
    for (int l = 0; l < numfields_; ++l) {
        if (!fields_[0][l].congruent(fieldsn[l])) fields_[0][l] = fieldsn[l];
        else {
            swap(fields_[0][l], fieldsn[l]);
            fields_[0][l].setPadded(fieldsn[l].padded());  // swap exchanges the data only (as flowfield.cpp:4076-4090)
        }
    }
    std::vector<Real> coef(2 * order_);
    std::vector<const FlowField*> terms(2 * order_);
    auto one_step = [&](bool recorded_only = false) {
        nse_->nonlinear(fields_[0], nonlf_[0]);
        // rhs = sum_j (-alpha_j/dt) u_j + (-beta_j) f_j, accumulated inside the solve kernel
        for (int j = 0; j < order_; ++j) {
            coef[2 * j] = -alpha_[j] / flags_.dt;
            terms[2 * j] = &fields_[j][0];
            coef[2 * j + 1] = -beta_[j];
            terms[2 * j + 1] = &nonlf_[j][0];
        }
        nse_->solve_lincomb(fields_[J], coef, terms, 0);
        for (int j = order_ - 1; j > 0; --j)
            for (int l = 0; l < numfields_; ++l) {
                swap(nonlf_[j][l], nonlf_[j - 1][l]);
                swap(fields_[j][l], fields_[j - 1][l]);
            }
        if (recorded_only) return;
        t_ += flags_.dt;
        tick();
    };
    int step = 0;
    // Launch-bound grids: `order_` consecutive steps rotate every history buffer back to its role, so that sequence is
    // captured once per call as a CUDA graph and replayed (one graph launch instead of 6 order_ kernel launches).  The
    // sequence runs eagerly first, which also brings every work space to its final size.  Capture + instantiation cost about
    // 1 ms per call (measured: break-even near 60 steps at 32x33x32), hence the minimum call length.
    if (Nsteps >= 16 * order_ && graph_replay_wanted(fields_[0][0], Nsteps)) {
        cfgpu_ctx ctx = cfgpu_context();
        for (int k = 0; k < order_; ++k, ++step) one_step();
        if (cfgpu_graph_begin(ctx) == 0) {
            for (int k = 0; k < order_; ++k) one_step(true);  // recorded, not executed
            int id = -1;
            if (cfgpu_graph_end(ctx, &id) != 0) {
                cfgpu_graph_abort(ctx);
                cferror(std::string("MultistepDNS::advance: CUDA graph capture of the step sequence failed: ") + cfgpu_last_error());
            }
            while (Nsteps - step >= order_) {
                if (cfgpu_graph_launch(ctx, id) != 0) cferror(std::string("MultistepDNS::advance: ") + cfgpu_last_error());
                for (int k = 0; k < order_; ++k, ++step) {
                    t_ += flags_.dt;
                    tick();
                }
            }
            cfgpu_graph_destroy(ctx, id);
        }
    }
    for (; step < Nsteps; ++step) one_step();
    for (int l = 0; l < numfields_; ++l) {
        if (!fields_[0][l].congruent(fieldsn[l])) fieldsn[l] = fields_[0][l];
        else {
            fieldsn[l].setPadded(fields_[0][l].padded());
            swap(fields_[0][l], fieldsn[l]);
        }
    }
    endline();
}

// dnsalgo.cpp:255-262: the whole history is projected
void MultistepDNS::project() {
    for (int n = 0; n < order_; ++n)
        for (int m = 0; m < numfields_; ++m) {
            fields_[n][m].project(symmetries_[m]);
            nonlf_[n][m].project(symmetries_[m]);
        }
}

void MultistepDNS::operator*=(const std::vector<FieldSymmetry>& sigma) {
    assert((int)sigma.size() == numfields_);
    for (int n = 0; n < order_; ++n)
        for (int m = 0; m < numfields_; ++m) {
            fields_[n][m] *= sigma[m];
            nonlf_[n][m] *= sigma[m];
        }
}

bool MultistepDNS::push(const std::vector<FlowField>& fields) {
    for (int j = order_ - 1; j > 0; --j)
        for (int l = 0; l < numfields_; ++l) {
            swap(nonlf_[j][l], nonlf_[j - 1][l]);
            swap(fields_[j][l], fields_[j - 1][l]);
        }
    if (order_ > 1) {
        fields_[1] = fields;
        nse_->nonlinear(fields_[1], nonlf_[1]);
    }
    t_ += flags_.dt;
    --countdown_;
    return full();
}

// =============================================================================================== Runge-Kutta
RungeKuttaDNS::RungeKuttaDNS() {}
RungeKuttaDNS::RungeKuttaDNS(const RungeKuttaDNS& d)
    : DNSAlgorithm(d), Nsubsteps_(d.Nsubsteps_), Qj1_(d.Qj1_), Qj_(d.Qj_), lt_(d.lt_), A_(d.A_), B_(d.B_), C_(d.C_) {}

RungeKuttaDNS::RungeKuttaDNS(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags)
    : DNSAlgorithm(fields, nse, flags), Qj1_(zeros_like(fields)), Qj_(zeros_like(fields)), lt_(zeros_like({fields[0]})) {
    if (flags_.timestepping != CNRK2) cferror("RungeKuttaDNS: flags.timestepping is a non-runge-kutta algorithm");
    order_ = 2; Nsubsteps_ = 3; Ninitsteps_ = 0;
    A_ = {0.0, -5.0 / 9.0, -153.0 / 128.0};
    B_ = {1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0};
    C_ = {1.0 / 6.0, 5.0 / 24.0, 1.0 / 8.0};
    reset_dt(flags_.dt);
}
DNSAlgorithm* RungeKuttaDNS::clone(const std::shared_ptr<NSE>& nse) const {
    DNSAlgorithm* c = new RungeKuttaDNS(*this);
    c->reset_nse(nse);
    return c;
}
void RungeKuttaDNS::reset_dt(Real dt) {
    flags_.dt = dt;
    lambda_t_.resize(Nsubsteps_);
    for (int j = 0; j < Nsubsteps_; ++j) lambda_t_[j] = 1.0 / (C_[j] * flags_.dt);
    nse_->reset_lambda(lambda_t_);
}
void RungeKuttaDNS::advance(std::vector<FlowField>& fields, int Nsteps) {
    std::vector<FlowField>& lt = lt_;  // work field of the linear term, allocated once (the reference re-creates it per call)
    for (int n = 0; n < Nsteps; ++n) {
        for (int j = 0; j < Nsubsteps_; ++j) {
            Qj_[0] *= A_[j];
            nse_->nonlinear(fields, Qj1_);
            Qj_[0] -= Qj1_[0];
            nse_->linear(fields, lt);
            nse_->solve_lincomb(fields, {1.0, lambda_t_[j], B_[j] / C_[j]}, {&lt[0], &fields[0], &Qj_[0]}, j);
        }
        t_ += flags_.dt;
        tick();
    }
    if (flags_.verbosity == PrintTime || flags_.verbosity == PrintAll) *flags_.logstream << std::endl;
}

// =============================================================================================== CNAB style
CNABstyleDNS::CNABstyleDNS() {}
CNABstyleDNS::CNABstyleDNS(const CNABstyleDNS& d)
    : DNSAlgorithm(d), Nsubsteps_(d.Nsubsteps_), full_(d.full_), fj1_(d.fj1_), fj_(d.fj_), lt_(d.lt_), alpha_(d.alpha_), beta_(d.beta_),
      gamma_(d.gamma_), zeta_(d.zeta_) {}

CNABstyleDNS::CNABstyleDNS(const std::vector<FlowField>& fields, const std::shared_ptr<NSE>& nse, const DNSFlags& flags)
    : DNSAlgorithm(fields, nse, flags), fj1_(zeros_like(fields)), fj_(zeros_like(fields)), lt_(zeros_like({fields[0]})) {
    switch (flags_.timestepping) {
        case CNAB2:
            order_ = 2; Nsubsteps_ = 1; Ninitsteps_ = 1; full_ = false;
            alpha_ = {0.5}; beta_ = {0.5}; gamma_ = {1.5}; zeta_ = {-0.5};
            break;
        case SMRK2:
            order_ = 2; Nsubsteps_ = 3; Ninitsteps_ = 0; full_ = true;
            alpha_ = {29.0 / 96.0, -3.0 / 40.0, 1.0 / 6.0};
            beta_ = {37.0 / 160.0, 5.0 / 24.0, 1.0 / 6.0};
            gamma_ = {8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0};
            zeta_ = {0.0, -17.0 / 60.0, -5.0 / 12.0};
            break;
        default:
            cferror("CNABstyleDNS: flags.timestepping is not a CNAB-style algorithm");
    }
    lambda_t_.resize(Nsubsteps_);
    for (int j = 0; j < Nsubsteps_; ++j) lambda_t_[j] = 1.0 / (beta_[j] * flags_.dt);
    nse_->reset_lambda(lambda_t_);
}
DNSAlgorithm* CNABstyleDNS::clone(const std::shared_ptr<NSE>& nse) const {
    DNSAlgorithm* c = new CNABstyleDNS(*this);
    c->reset_nse(nse);
    return c;
}
void CNABstyleDNS::reset_dt(Real dt) {
    flags_.dt = dt;
    for (int j = 0; j < Nsubsteps_; ++j) lambda_t_[j] = 1.0 / (beta_[j] * flags_.dt);
    nse_->reset_lambda(lambda_t_);
    if (flags_.timestepping == CNAB2) full_ = false;
}
bool CNABstyleDNS::push(const std::vector<FlowField>& fields) {
    for (int l = 0; l < numfields_; ++l) swap(fj_[l], fj1_[l]);
    nse_->nonlinear(fields, fj_);
    t_ += flags_.dt;
    full_ = true;
    return full_;
}
void CNABstyleDNS::advance(std::vector<FlowField>& fields, int Nsteps) {
    std::vector<FlowField>& lt = lt_;
    for (int n = 0; n < Nsteps; ++n) {
        for (int j = 0; j < Nsubsteps_; ++j) {
            for (int l = 0; l < numfields_; ++l) swap(fj_[l], fj1_[l]);
            nse_->nonlinear(fields, fj_);
            const Real a_b = alpha_[j] / beta_[j], g_b = gamma_[j] / beta_[j], z_b = zeta_[j] / beta_[j];
            nse_->linear(fields, lt);
            // rhs = lambda_j u + (a/b) L(u,q) - (g/b) f_j - (z/b) f_{j-1}
            nse_->solve_lincomb(fields, {lambda_t_[j], a_b, -g_b, -z_b}, {&fields[0], &lt[0], &fj_[0], &fj1_[0]}, j);
        }
        t_ += flags_.dt;
        tick();
    }
    if (flags_.verbosity == PrintTime || flags_.verbosity == PrintAll) *flags_.logstream << std::endl;
}

}  // namespace chflow
