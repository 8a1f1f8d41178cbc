// DNSFlags / TimeStep host logic; behaviour follows the reference's dnsflags.cpp (enum <-> string tables :276-597,
// TimeStep :749-975).
#include "channelflow/dnsflags.h"

#include <climits>

namespace chflow {

DNSFlags::DNSFlags(Real nu_, Real dPdx_, Real dPdz_, Real Ubulk_, Real Wbulk_, Real Uwall_, Real ulowerwall_,
                   Real uupperwall_, Real wlowerwall_, Real wupperwall_, Real theta_, Real Vsuck_, Real rotation_, Real t0_,
                   Real T_, Real dT_, Real dt_, bool variabledt_, Real dtmin_, Real dtmax_, Real CFLmin_, Real CFLmax_,
                   Real symmetryprojectioninterval_, BaseFlow baseflow_, MeanConstraint constraint_,
                   TimeStepMethod timestepping_, TimeStepMethod initstepping_, NonlinearMethod nonlinearity_,
                   Dealiasing dealiasing_, BodyForce* bodyforce_, bool taucorrection_, Verbosity verbosity_,
                   std::ostream* logstream_)
    : baseflow(baseflow_), constraint(constraint_), timestepping(timestepping_), initstepping(initstepping_),
      nonlinearity(nonlinearity_), dealiasing(dealiasing_), bodyforce(bodyforce_), taucorrection(taucorrection_), nu(nu_),
      Vsuck(Vsuck_), rotation(rotation_), theta(theta_), dPdx(dPdx_), dPdz(dPdz_), Ubulk(Ubulk_), Wbulk(Wbulk_),
      Uwall(Uwall_), ulowerwall(ulowerwall_), uupperwall(uupperwall_), wlowerwall(wlowerwall_), wupperwall(wupperwall_),
      t0(t0_), T(T_), dT(dT_), dt(dt_), variabledt(variabledt_), dtmin(dtmin_), dtmax(dtmax_), CFLmin(CFLmin_),
      CFLmax(CFLmax_), symmetryprojectioninterval((int)symmetryprojectioninterval_), verbosity(verbosity_),
      logstream(logstream_) {
    if (dealias_y() && nonlinearity != Rotational) {
        std::cerr << "DNSFlags::DNSFlags: DealiasY and DealiasXYZ work only with Rotational nonlinearity.\n"
                  << "Setting nonlinearity to Rotational." << std::endl;
        nonlinearity = Rotational;
    }
}

namespace {
template <class E>
E lookup(const std::string& s, const std::vector<std::pair<const char*, E>>& tab, const char* what) {
    std::string t;
    for (char c : s) t += (char)tolower(c);
    for (auto& kv : tab)
        if (t.find(kv.first) != std::string::npos) return kv.second;
    cferror(std::string("unrecognized ") + what + " : " + s);
}
}  // namespace

BaseFlow s2baseflow(const std::string& s) {
    return lookup<BaseFlow>(s, {{"zero", ZeroBase}, {"linear", LinearBase}, {"parabolic", ParabolicBase}, {"laminar", LaminarBase},
                                {"suction", SuctionBase}, {"arbitrary", ArbitraryBase}}, "base flow");
}
MeanConstraint s2constraint(const std::string& s) {
    return lookup<MeanConstraint>(s, {{"gradp", PressureGradient}, {"pressure", PressureGradient}, {"bulkv", BulkVelocity},
                                      {"velocity", BulkVelocity}}, "mean constraint");
}
TimeStepMethod s2stepmethod(const std::string& s) {
    return lookup<TimeStepMethod>(s, {{"cnfe1", CNFE1}, {"cnab2", CNAB2}, {"cnrk2", CNRK2}, {"smrk2", SMRK2}, {"sbdf1", SBDF1},
                                      {"sbdf2", SBDF2}, {"sbdf3", SBDF3}, {"sbdf4", SBDF4}}, "time-stepping method");
}
NonlinearMethod s2nonlmethod(const std::string& s) {
    return lookup<NonlinearMethod>(s, {{"rot", Rotational}, {"conv", Convection}, {"skew", SkewSymmetric}, {"alt", Alternating},
                                       {"div", Divergence}, {"lin", LinearAboutProfile}}, "nonlinear method");
}
Dealiasing s2dealiasing(const std::string& s) {
    return lookup<Dealiasing>(s, {{"nodealiasing", NoDealiasing}, {"dealiasxyz", DealiasXYZ}, {"dealiasxz", DealiasXZ},
                                  {"dealiasy", DealiasY}}, "dealiasing");
}
Verbosity s2verbosity(const std::string& s) {
    return lookup<Verbosity>(s, {{"silent", Silent}, {"printtime", PrintTime}, {"printticks", PrintTicks},
                                 {"verifytausolve", VerifyTauSolve}, {"printall", PrintAll}}, "verbosity");
}
VelocityScale s2velocityscale(const std::string& s) {
    return lookup<VelocityScale>(s, {{"wall", WallScale}, {"parab", ParabolicScale}}, "velocity scale");
}
std::string baseflow2string(BaseFlow b) {
    static const char* n[] = {"ZeroBase", "LinearBase", "ParabolicBase", "LaminarBase", "SuctionBase", "ArbitraryBase"};
    return n[(int)b];
}
std::string constraint2string(MeanConstraint m) { return m == PressureGradient ? "PressureGradient" : "BulkVelocity"; }
std::string stepmethod2string(TimeStepMethod t) {
    static const char* n[] = {"CNFE1", "CNAB2", "CNRK2", "SMRK2", "SBDF1", "SBDF2", "SBDF3", "SBDF4"};
    return n[(int)t];
}
std::string nonlmethod2string(NonlinearMethod m) {
    static const char* n[] = {"Rotational", "Convection", "Divergence", "SkewSymmetric", "Alternating", "Alternating_",
                              "LinearAboutProfile"};
    return n[(int)m];
}
std::string dealiasing2string(Dealiasing d) {
    static const char* n[] = {"NoDealiasing", "DealiasXZ", "DealiasY", "DealiasXYZ"};
    return n[(int)d];
}
std::ostream& operator<<(std::ostream& os, BaseFlow b) { return os << baseflow2string(b); }
std::ostream& operator<<(std::ostream& os, MeanConstraint m) { return os << constraint2string(m); }
std::ostream& operator<<(std::ostream& os, TimeStepMethod t) { return os << stepmethod2string(t); }
std::ostream& operator<<(std::ostream& os, NonlinearMethod n) { return os << nonlmethod2string(n); }
std::ostream& operator<<(std::ostream& os, Dealiasing d) { return os << dealiasing2string(d); }

std::ostream& operator<<(std::ostream& os, const DNSFlags& f) {
    const char* s = ", ";
    const auto p = os.precision();
    os.precision(16);
    os << "nu==" << f.nu << s << "Vsuck==" << f.Vsuck << s << "rotation==" << f.rotation << s << "dPdx==" << f.dPdx << s
       << "Ubulk==" << f.Ubulk << s << "dt==" << f.dt << s << f.baseflow << s << f.constraint << s << f.timestepping << s
       << f.initstepping << s << f.nonlinearity << s << f.dealiasing << s
       << (f.taucorrection ? "TauCorrection" : "NoTauCorrection");
    os.precision(p);
    return os;
}

// ---------------------------------------------------------------------------------------------- TimeStep
TimeStep::TimeStep() : n_(0), N_(0), dt_(0), dtmin_(0), dtmax_(0), dT_(0), T_(0), CFLmin_(0), CFL_(0), CFLmax_(0), variable_(false) {}

TimeStep::TimeStep(Real dt, Real dtmin, Real dtmax, Real dT, Real CFLmin, Real CFLmax, bool variable)
    : n_(0), N_(0), dt_(dt), dtmin_(dtmin), dtmax_(dtmax), dT_(dT), T_(0), CFLmin_(CFLmin), CFL_((CFLmax + CFLmin) / 2),
      CFLmax_(CFLmax), variable_(variable) {
    if (dtmin < 0 || dt < dtmin || dtmax < dt) cferror("TimeStep: condition 0 <= dtmin <= dt <= dtmax does not hold");
    if (CFLmin < 0 || CFLmax < CFLmin) cferror("TimeStep: condition 0 <= CFLmin <= CFLmax does not hold");
    if (dT < dtmin) cferror("TimeStep: dT < dtmin");
    n_ = Greater(iround(dT / dt), 1);
    dt_ = dT_ / n_;
    while (dt_ < dtmin_ && n_ >= 2 && dT_ != 0) dt_ = dT_ / --n_;
    while (dt_ > dtmax_ && n_ <= INT_MAX && dT_ != 0) dt_ = dT_ / ++n_;
}
TimeStep::TimeStep(DNSFlags& f) : TimeStep(f.dt, f.dtmin, f.dtmax, f.dT, f.CFLmin, f.CFLmax, f.variabledt) {}

bool TimeStep::adjust(Real CFL, bool verbose, std::ostream& os) {
    CFL_ = CFL;
    if (variable_ && (CFL <= CFLmin_ || CFL >= CFLmax_)) return adjustToMiddle(CFL, verbose, os);
    return false;
}
bool TimeStep::adjustToMiddle(Real CFL, bool verbose, std::ostream& os) {
    if (dtmin_ == dtmax_ || dT_ == 0.0) return false;
    int n = Greater(iround(2 * n_ * CFL / (CFLmax_ + CFLmin_)), 1);
    Real dt = dT_ / n;
    while (dt < dtmin_ && dt < dT_) dt = dT_ / --n;
    while (dt > dtmax_ && n <= INT_MAX) dt = dT_ / ++n;
    CFL *= dt / dt_;
    if (verbose && (CFL > CFLmax_ || CFL < CFLmin_))
        os << "TimeStep::adjust(CFL) : dt " << (CFL > CFLmax_ ? "bottomed" : "topped") << " out at\n dt  == " << dt
           << "\n CFL == " << CFL << "\n n   == " << n << std::endl;
    const bool adjustment = n != n_;
    if (adjustment) {
        if (verbose)
            os << "TimeStep::adjust(CFL) { \n   n : " << n_ << " -> " << n << "\n  dt : " << dt_ << " -> " << dt
               << "\n CFL : " << CFL_ << " -> " << CFL << "\n}" << std::endl;
        n_ = n;
        dt_ = dt;
        CFL_ = CFL;
    }
    return adjustment;
}
bool TimeStep::adjust(Real a, Real a_max, bool verbose, std::ostream& os) {
    if (variable_ && a >= a_max) return adjustToDesired(a, a_max, verbose, os);
    return false;
}
bool TimeStep::adjustToDesired(Real a, Real a_des, bool verbose, std::ostream& os) {
    if (dtmin_ == dtmax_ || dT_ == 0.0) return false;
    int n = Greater(iround(n_ * a / a_des), 1);
    Real dt = dT_ / n;
    while (dt < dtmin_ && dt < dT_) dt = dT_ / --n;
    while (dt > dtmax_ && n <= INT_MAX) dt = dT_ / ++n;
    a *= dt / dt_;
    if (verbose && a > a_des) os << "TimeStep::adjust(a) : dt bottomed out at dt == " << dt << std::endl;
    const bool adjustment = n != n_;
    if (adjustment) {
        n_ = n;
        dt_ = dt;
    }
    return adjustment;
}
bool TimeStep::adjust_for_T(Real T, bool verbose, std::ostream& os) {
    T_ = T;
    if (T < 0) cferror("TimeStep::adjust_for_T : can't integrate backwards in time.");
    if (T == 0) {
        const bool adjustment = dt_ != 0;
        dt_ = 0; n_ = 0; dT_ = 0; T_ = 0;
        return adjustment;
    }
    const int N = Greater(iround(T / dT_), 1);
    const Real dT = T / N;
    int n = Greater(iround(dT / dt_), 1);
    Real dt = dT / n;
    while (dt < dtmin_ && n > 2 && dT != 0) dt = dT / --n;
    while (dt > dtmax_ && n <= INT_MAX && dT != 0) dt = dT / ++n;
    const Real CFL = dt * CFL_ / dt_;
    const bool adjustment = dt != dt_;
    if (adjustment && verbose)
        os << "TimeStep::adjust_for_T(Real T) { dT : " << dT_ << " -> " << dT << ", dt : " << dt_ << " -> " << dt << " }" << std::endl;
    n_ = n; N_ = N; dt_ = dt; dT_ = dT; CFL_ = CFL;
    return adjustment;
}
std::ostream& operator<<(std::ostream& os, const TimeStep& dt) {
    os << "{dt=" << dt.dt() << ", n=" << dt.n() << ", dT=" << dt.dT() << ", N=" << dt.N() << ", dtmin=" << dt.dtmin()
       << ", dtmax=" << dt.dtmax() << ", CFLmin=" << dt.CFLmin() << ", CFL=" << dt.CFL() << ", CFLmax=" << dt.CFLmax()
       << ", variable=" << dt.variable() << "}";
    return os;
}

}  // namespace chflow
