// chflow::NSE over the cfgpu C-ABI.  Host part: base flow, mean constraint bookkeeping (reference nse.cpp:212-289,
// 589-671, 724-745, 774-951); device part: cfgpu_nse_* (nonlinear, linear, solve, reset_lambda).
#include "channelflow/nse.h"

#include "channelflow/diffops.h"

namespace chflow {

#define CK(call) cfgpu_check((call), #call)

NSE::NSE() {}

NSE::NSE(const NSE& o)
    : lambda_t_(o.lambda_t_), flags_(o.flags_), Nx_(o.Nx_), My_(o.My_), Nz_(o.Nz_), Lx_(o.Lx_), Lz_(o.Lz_), a_(o.a_), b_(o.b_),
      dPdxRef_(o.dPdxRef_), dPdzRef_(o.dPdzRef_), UbulkRef_(o.UbulkRef_), UbulkAct_(o.UbulkAct_), UbulkBase_(o.UbulkBase_),
      WbulkRef_(o.WbulkRef_), WbulkAct_(o.WbulkAct_), WbulkBase_(o.WbulkBase_), dPdxAct_(o.dPdx()), dPdzAct_(o.dPdz()),
      Ubase_(o.Ubase_), Wbase_(o.Wbase_) {
    if (o.dev_) {
        create_device();
        if (!lambda_t_.empty()) CK(cfgpu_nse_reset_lambda(dev_, lambda_t_.data(), (int)lambda_t_.size()));
    }
}

static void geometry_from(const FlowField& u, int& Nx, int& My, int& Nz, Real& Lx, Real& Lz, Real& a, Real& b) {
    Nx = u.Nx(); My = u.Ny(); Nz = u.Nz(); Lx = u.Lx(); Lz = u.Lz(); a = u.a(); b = u.b();
}

NSE::NSE(const std::vector<FlowField>& fields, const DNSFlags& flags) : flags_(flags) {
    assert(fields[0].vectorDim() == 3);
    geometry_from(fields[0], Nx_, My_, Nz_, Lx_, Lz_, a_, b_);
    createCFBaseFlow();
    initCFConstraint(fields[0]);
    create_device();
}

NSE::NSE(const std::vector<FlowField>& fields, const std::vector<ChebyCoeff>& base, const DNSFlags& flags)
    : flags_(flags), Ubase_(base[0]), Wbase_(base[1]) {
    assert(fields[0].vectorDim() == 3);
    geometry_from(fields[0], Nx_, My_, Nz_, Lx_, Lz_, a_, b_);
    Ubase_.makeSpectral();
    Wbase_.makeSpectral();
    initCFConstraint(fields[0]);
    create_device();
}

NSE::~NSE() {
    if (dev_) cfgpu_nse_destroy(dev_);
}

void NSE::create_device() {
    cfgpu_nse_config c;
    c.nu = flags_.nu;
    c.Vsuck = flags_.Vsuck;
    c.rotation = flags_.rotation;
    c.nonlinearity = (int)flags_.nonlinearity;
    c.dealias_xz = flags_.dealias_xz() ? 1 : 0;
    c.dealias_y = flags_.dealias_y() ? 1 : 0;
    c.taucorrection = flags_.taucorrection ? 1 : 0;
    c.constraint = (int)flags_.constraint;
    c.dPdxRef = dPdxRef_;
    c.dPdzRef = dPdzRef_;
    c.UbulkRef_minus_base = UbulkRef_ - UbulkBase_;
    c.WbulkRef_minus_base = WbulkRef_ - WbulkBase_;
    std::vector<Real> U(My_, 0.0), W(My_, 0.0);
    for (int n = 0; n < My_ && n < Ubase_.length(); ++n) U[n] = Ubase_[n];
    for (int n = 0; n < My_ && n < Wbase_.length(); ++n) W[n] = Wbase_[n];
    CK(cfgpu_nse_create(cfgpu_context(), Nx_, My_, Nz_, Lx_, Lz_, a_, b_, &c, U.data(), W.data(), &dev_));
}

void NSE::push_constraint() {
    if (dev_)
        CK(cfgpu_nse_set_constraint(dev_, (int)flags_.constraint, dPdxRef_, dPdzRef_, UbulkRef_ - UbulkBase_,
                                    WbulkRef_ - WbulkBase_));
}

void NSE::createCFBaseFlow() {
    Ubase_ = ChebyCoeff(My_, a_, b_, Spectral);
    Wbase_ = ChebyCoeff(My_, a_, b_, Spectral);
    switch (flags_.baseflow) {
        case ZeroBase:
            break;
        case LinearBase:
            Ubase_[1] = 1;
            break;
        case ParabolicBase:
            Ubase_[0] = 0.5;
            Ubase_[2] = -0.5;
            break;
        case SuctionBase:
            Ubase_ = laminarProfile(flags_.nu, PressureGradient, 0, flags_.Ubulk, flags_.Vsuck, a_, b_, -0.5, 0.5, My_);
            break;
        case LaminarBase:
            Ubase_ = laminarProfile(flags_.nu, flags_.constraint, flags_.dPdx, flags_.Ubulk, flags_.Vsuck, a_, b_,
                                    flags_.ulowerwall, flags_.uupperwall, My_);
            Wbase_ = laminarProfile(flags_.nu, flags_.constraint, flags_.dPdz, flags_.Wbulk, flags_.Vsuck, a_, b_,
                                    flags_.wlowerwall, flags_.wupperwall, My_);
            break;
        default:
            cferror("error in NSE::createBaseFlow : flags.baseflow should be ZeroBase, LinearBase, ParabolicBase, "
                    "LaminarBase or SuctionBase; other cases require the DNS(fields, base, flags) constructor.");
    }
}

void NSE::initCFConstraint(const FlowField& u) {
    UbulkBase_ = Ubase_.mean();
    WbulkBase_ = Wbase_.mean();
    ComplexChebyCoeff u00 = u.profile(0, 0, 0), w00 = u.profile(0, 0, 2);
    Real ub = u00.re.mean(), wb = w00.re.mean();
    if (std::abs(ub) < 1e-15) ub = 0.0;
    if (std::abs(wb) < 1e-15) wb = 0.0;
    UbulkAct_ = UbulkBase_ + ub;
    WbulkAct_ = WbulkBase_ + wb;
    const Real Ly = b_ - a_;
    {   // wall shear of the total mean profile (reference getdPdx(u + Ubase), nse.cpp:647-660)
        ChebyCoeff t = u00.re;
        t += Ubase_;
        ChebyCoeff d = diff(t);
        dPdxAct_ = flags_.nu * (d.eval_b() - d.eval_a()) / Ly;
        t = w00.re;
        t += Wbase_;
        d = diff(t);
        dPdzAct_ = flags_.nu * (d.eval_b() - d.eval_a()) / Ly;
    }
    if (flags_.constraint == BulkVelocity) {
        UbulkRef_ = flags_.Ubulk;
        WbulkRef_ = flags_.Wbulk;
    } else {
        dPdxAct_ = dPdxRef_ = flags_.dPdx;
        dPdzAct_ = dPdzRef_ = flags_.dPdz;
    }
}

void NSE::reset_lambda(const std::vector<Real> lambda_t) {
    lambda_t_ = lambda_t;
    CK(cfgpu_nse_reset_lambda(dev_, lambda_t_.data(), (int)lambda_t_.size()));
}

void NSE::nonlinear(const std::vector<FlowField>& infields, std::vector<FlowField>& outfields) {
    FlowField& f = outfields[0];
    const FlowField& u = infields[0];
    if (!u.geomCongruent(f) || f.Nd() != 3) f.resize(u.Nx(), u.Ny(), u.Nz(), 3, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi());
    CK(cfgpu_nse_nonlinear(dev_, u.device(), f.device_mut()));
    f.setState(Spectral, Spectral);
    if (flags_.dealias_xz()) f.setPadded(true);
}

void NSE::linear(const std::vector<FlowField>& infields, std::vector<FlowField>& outfields) {
    assert(infields.size() == outfields.size() + 1);
    CK(cfgpu_nse_linear(dev_, infields[0].device(), infields[1].device(), outfields[0].device_mut()));
    outfields[0].setState(Spectral, Spectral);
}

void NSE::solve(std::vector<FlowField>& outfields, const std::vector<FlowField>& rhs, const int s) {
    assert(outfields.size() == rhs.size() + 1);
    solve_lincomb(outfields, {1.0}, {&rhs[0]}, s);
}

void NSE::solve_lincomb(std::vector<FlowField>& outfields, const std::vector<Real>& coef,
                        const std::vector<const FlowField*>& terms, const int s) {
    std::vector<cfgpu_field> t(terms.size());
    for (size_t j = 0; j < terms.size(); ++j) t[j] = terms[j]->device();
    CK(cfgpu_nse_solve(dev_, s, (int)terms.size(), coef.data(), t.data(), outfields[0].device_mut(), outfields[1].device_mut()));
    outfields[0].setState(Spectral, Spectral);
    outfields[1].setState(Spectral, Spectral);
    if (flags_.constraint == BulkVelocity) dPd_on_device_ = true;
}

std::vector<FlowField> NSE::createRHS(const std::vector<FlowField>& fields) const { return {fields[0]}; }

// velocity: the flags' symmetries; pressure: the identity (nse.cpp:578-587)
std::vector<cfarray<FieldSymmetry>> NSE::createSymmVec() const {
    cfarray<FieldSymmetry> usym = SymmetryList(1), psym = SymmetryList(1);
    if (flags_.symmetries.length() > 0) usym = flags_.symmetries;
    return {usym, psym};
}

Real NSE::dPdx() const {
    if (dPd_on_device_) {
        CK(cfgpu_nse_get_dPd(dev_, &dPdxAct_, &dPdzAct_));
        dPd_on_device_ = false;
    }
    return dPdxAct_;
}
Real NSE::dPdz() const {
    dPdx();
    return dPdzAct_;
}

void NSE::reset_gradp(Real dPdx, Real dPdz) {
    flags_.constraint = PressureGradient;
    flags_.dPdx = dPdx; flags_.dPdz = dPdz; flags_.Ubulk = 0.0; flags_.Wbulk = 0.0;
    dPdxRef_ = dPdx; dPdzRef_ = dPdz; UbulkRef_ = 0.0; WbulkRef_ = 0.0;
    push_constraint();
}
void NSE::reset_bulkv(Real Ubulk, Real Wbulk) {
    flags_.constraint = BulkVelocity;
    flags_.Ubulk = Ubulk; flags_.Wbulk = Wbulk; flags_.dPdx = 0.0; flags_.dPdz = 0.0;
    UbulkRef_ = Ubulk; WbulkRef_ = Wbulk; dPdxRef_ = 0.0; dPdzRef_ = 0.0;
    push_constraint();
}

Real NSE::CFLfactor(const FlowField& u) const {
    Real r = 0;
    CK(cfgpu_nse_cflfactor(dev_, u.device(), &r));
    return r;
}

Real NSE::cflfactor_of(const FlowField& u, const ChebyCoeff& U, const ChebyCoeff& W, const DNSFlags& flags) {
    FlowField v(u);
    v.makeSpectral();
    cfgpu_nse_config c;
    c.nu = 1.0; c.Vsuck = 0; c.rotation = 0; c.nonlinearity = 0;
    c.dealias_xz = flags.dealias_xz() ? 1 : 0;
    c.dealias_y = 0; c.taucorrection = 1; c.constraint = 0;
    c.dPdxRef = c.dPdzRef = c.UbulkRef_minus_base = c.WbulkRef_minus_base = 0;
    std::vector<Real> Uc(u.Ny(), 0.0), Wc(u.Ny(), 0.0);
    for (int n = 0; n < u.Ny() && n < U.length(); ++n) Uc[n] = U[n];
    for (int n = 0; n < u.Ny() && n < W.length(); ++n) Wc[n] = W[n];
    cfgpu_nse h = nullptr;
    CK(cfgpu_nse_create(cfgpu_context(), u.Nx(), u.Ny(), u.Nz(), u.Lx(), u.Lz(), u.a(), u.b(), &c, Uc.data(), Wc.data(), &h));
    Real r = 0;
    CK(cfgpu_nse_cflfactor(h, v.device(), &r));
    cfgpu_nse_destroy(h);
    return r;
}

// The nonlinear term as a free function (nse.cpp:12-91): f = N(u + Ubase e_x + Wbase e_z - Vsuck e_y) in the form
// flags.nonlinearity names, NOT de-aliased (NSE::nonlinear does that afterwards, nse.cpp:383-391); the Alternating forms
// toggle flags.nonlinearity as in the reference.  tmp is scratch of the reference's host algorithm: unused, the device
// pipeline owns its pencils.
void navierstokesNL(const FlowField& u, ChebyCoeff Ubase, ChebyCoeff Wbase, FlowField& f, FlowField&, DNSFlags& flags) {
    assert(u.xzstate() == Spectral && u.ystate() == Spectral && Ubase.state() == Spectral && Wbase.state() == Spectral);
    cfgpu_nse_config c;
    c.nu = flags.nu; c.Vsuck = flags.Vsuck; c.rotation = flags.rotation;
    c.nonlinearity = (int)flags.nonlinearity;
    c.dealias_xz = 0; c.dealias_y = 0; c.taucorrection = 1; c.constraint = 0;
    c.dPdxRef = c.dPdzRef = c.UbulkRef_minus_base = c.WbulkRef_minus_base = 0;
    std::vector<Real> Uc(u.Ny(), 0.0), Wc(u.Ny(), 0.0);
    for (int n = 0; n < u.Ny() && n < Ubase.length(); ++n) Uc[n] = Ubase[n];
    for (int n = 0; n < u.Ny() && n < Wbase.length(); ++n) Wc[n] = Wbase[n];
    cfgpu_nse h = nullptr;
    CK(cfgpu_nse_create(cfgpu_context(), u.Nx(), u.Ny(), u.Nz(), u.Lx(), u.Lz(), u.a(), u.b(), &c, Uc.data(), Wc.data(), &h));
    if (!u.geomCongruent(f) || f.Nd() != 3) f.resize(u.Nx(), u.Ny(), u.Nz(), 3, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi());
    const int rc = cfgpu_nse_nonlinear(h, u.device(), f.device_overwrite());
    cfgpu_nse_destroy(h);
    CK(rc);
    f.setState(Spectral, Spectral);
    f.setPadded(false);
    if (flags.nonlinearity == Alternating) flags.nonlinearity = Alternating_;
    else if (flags.nonlinearity == Alternating_) flags.nonlinearity = Alternating;
}

// ---- the forms of the nonlinear term as free functions (diffops.h:226-231)
static void nl_form(NonlinearMethod m, const FlowField& u, const ChebyCoeff& U, const ChebyCoeff& W, FlowField& f, fieldstate finalstate) {
    DNSFlags flags;
    flags.nonlinearity = m;
    flags.Vsuck = 0.0;
    FlowField tmp;
    navierstokesNL(u, U, W, f, tmp, flags);
    if (finalstate == Physical) f.makePhysical();
}
static ChebyCoeff zero_profile(const FlowField& u) { return ChebyCoeff(u.Ny(), u.a(), u.b(), Spectral); }
void rotationalNL(const FlowField& u, FlowField& f, FlowField&, const fieldstate fs) { nl_form(Rotational, u, zero_profile(u), zero_profile(u), f, fs); }
void convectionNL(const FlowField& u, FlowField& f, FlowField&, const fieldstate fs) { nl_form(Convection, u, zero_profile(u), zero_profile(u), f, fs); }
void divergenceNL(const FlowField& u, FlowField& f, FlowField&, const fieldstate fs) { nl_form(Divergence, u, zero_profile(u), zero_profile(u), f, fs); }
void skewsymmetricNL(const FlowField& u, FlowField& f, FlowField&, const fieldstate fs) { nl_form(SkewSymmetric, u, zero_profile(u), zero_profile(u), f, fs); }
void linearizedNL(const FlowField& u, const ChebyCoeff& U, const ChebyCoeff& W, FlowField& f, const fieldstate fs) {
    ChebyCoeff Us(U), Ws(W);
    Us.makeSpectral();
    Ws.makeSpectral();
    nl_form(LinearAboutProfile, u, Us, Ws, f, fs);
}

// ---------------------------------------------------------------------------------------------- free functions
Real viscosity(Real Reynolds, VelocityScale vscale, MeanConstraint constraint, Real dPdx, Real Ubulk, Real Uwall, Real h) {
    if (vscale == WallScale) return fabs(Uwall) * h / Reynolds;
    if (constraint == PressureGradient) return sqrt(pow(h, 3) * fabs(dPdx) / (2 * Reynolds));
    return 1.5 * Ubulk * h / Reynolds;
}

// Laminar solution of nu U'' + Vsuck U' = dPdx with U(a)=ua, U(b)=ub (reference nse.cpp:823-943):
// quadratic for Vsuck == 0, exponential (ASBL) otherwise, Taylor-expanded in mu = Vsuck H / nu for |mu| <= 0.1.
ChebyCoeff laminarProfile(Real nu, MeanConstraint constraint, Real dPdx, Real Ubulk, Real Vsuck, Real a, Real b, Real ua,
                          Real ub, int Ny) {
    ChebyCoeff u(Ny, a, b, Spectral);
    const Real H = b - a;
    const Real mu = Vsuck * H / nu;
    if (std::abs(mu) == 0.0) {
        if (constraint == BulkVelocity) {
            u[0] = 0.125 * (ub + ua) + 0.75 * Ubulk;
            u[1] = 0.5 * (ub - ua);
            u[2] = 0.375 * (ub + ua) - 0.75 * Ubulk;
        } else {
            dPdx *= square((b - a) / 2);
            u[0] = 0.5 * (ub + ua) - 0.25 * dPdx / nu;
            u[1] = 0.5 * (ub - ua);
            u[2] = 0.25 * dPdx / nu;
        }
        return u;
    }
    u.setState(Physical);
    const Real dU = ub - ua;
    const Real em = expm1(-mu);
    Real G = dPdx * square(H) / nu;  // dimensionless pressure gradient
    const Vector y = chebypoints(Ny, a, b);
    if (std::abs(mu) > 1e-01) {
        if (constraint == BulkVelocity) {
            const Real k = -1.0 / em - 1 / mu;
            G = mu * (Ubulk - ua - dU * k) / (0.5 - k);
        }
        for (int i = 0; i < Ny; ++i) {
            const Real s = (y[i] - a) / H;
            const Real es = expm1(-mu * s);
            u[i] = ua + dU * es / em + G * (s - es / em) / mu;
        }
    } else {
        if (constraint == BulkVelocity) {
            const Real m2 = mu * mu, A = ub - ua, B = ua + ub - 2 * Ubulk;
            G = 6 * B + mu * (A + mu * (B / 10 + m2 * (-B / 1400 + m2 * B / 126000)));
        }
        for (int i = 0; i < Ny; ++i) {
            const Real s = (y[i] - a) / H, s2 = s * s, s3 = s2 * s, s4 = s3 * s, t = s - 1;
            const Real p = s * t, p2 = square(p), q = 2 * s - 1;
            const Real c3 = 3 * s2 - 3 * s - 1, c4 = 2 * s2 - 2 * s - 1, c5 = 3 * s4 - 6 * s3 + 3 * s + 1,
                       c6 = 3 * s4 - 6 * s3 + 4 * s + 2;
            // series of expm1(-mu s)/expm1(-mu) and of (s - that)/mu
            const Real t1 = s + mu * (-p / 2 + mu * (p * q / 12 + mu * (-p2 / 24 + mu * (p * q * c3 / 720 +
                            mu * (-p2 * c4 / 1440 + mu * (p * q * c5 / 30240))))));
            const Real t2 = p / 2 + mu * (-p * q / 12 + mu * (p2 / 24 + mu * (-p * q * c3 / 720 + mu * (p2 * c4 / 1440 +
                            mu * (-p * q * c5 / 30240 + mu * (p2 * c6 / 120960))))));
            u[i] = ua + dU * t1 + G * t2;
        }
    }
    u.makeSpectral();
    return u;
}

ChebyCoeff laminarProfile(const DNSFlags& flags, Real a, Real b, int Ny) {
    if (flags.baseflow == SuctionBase)
        return laminarProfile(flags.nu, PressureGradient, 0, flags.Ubulk, flags.Vsuck, a, b, -0.5, 0.5, Ny);
    return laminarProfile(flags.nu, flags.constraint, flags.dPdx, flags.Ubulk, flags.Vsuck, a, b, flags.ulowerwall,
                          flags.uupperwall, Ny);
}

}  // namespace chflow
