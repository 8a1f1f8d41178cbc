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

std::ostream& operator<<(std::ostream& os, VelocityScale v) { return os << (v == WallScale ? "WallScale" : "ParabolicScale"); }
std::ostream& operator<<(std::ostream& os, Verbosity v) {
    static const char* n[] = {"Silent", "PrintTime", "PrintTicks", "VerifyTauSolve", "PrintAll"};
    return os << n[(int)v];
}

// ---- BodyForce (never evaluated by the time steppers, see the header)
Vector BodyForce::operator()(Real x, Real y, Real z, Real t) {
    Vector f(3);
    eval(x, y, z, t, f[0], f[1], f[2]);
    return f;
}
void BodyForce::eval(Real, Real, Real, Real, Real& fx, Real& fy, Real& fz) { fx = fy = fz = 0.0; }
bool BodyForce::isOn(Real) { return false; }
void BodyForce::eval(Real t, FlowField& f) {
    f.makePhysical();
    for (int ny = 0; ny < f.Ny(); ++ny)
        for (int nx = 0; nx < f.Nx(); ++nx)
            for (int nz = 0; nz < f.Nz(); ++nz) eval(f.x(nx), f.y(ny), f.z(nz), t, f(nx, ny, nz, 0), f(nx, ny, nz, 1), f(nx, ny, nz, 2));
    f.makeSpectral();
}

// ---- command-line construction: option names, defaults and help texts are the user interface of the reference's
// programs (dnsflags.cpp:160-264) and are kept verbatim
DNSFlags::DNSFlags(ArgList& args, const bool laurette) : DNSFlags() {
    args.section("System parameters");
    const Real Reynolds = args.getreal("-R", "--Reynolds", 400, "pseudo-Reynolds number == 1/nu");
    const Real nuarg = args.getreal("-nu", "--nu", 0, "kinematic viscosity (takes precedence over Reynolds, if nonzero)");
    nu = nuarg != 0 ? nuarg : 1.0 / Reynolds;
    args2BC(args);
    args2numerics(args, laurette);
}
void DNSFlags::args2BC(ArgList& args) {
    args.section("Boundary conditions");
    baseflow = s2baseflow(args.getstr("-bf", "--baseflow", "laminar", "set base flow to one of [zero|laminar|linear|parabolic|suction|arbitrary]"));
    constraint = s2constraint(args.getstr("-mc", "--meanconstraint", "gradp", "fix one of two flow constraints [gradp|bulkv]"));
    const Real dPds = args.getreal("-dPds", "--dPds", 0.0, "magnitude of imposed pressure gradient along streamwise s");
    const Real Ub = args.getreal("-Ubulk", "--Ubulk", 0.0, "magnitude of imposed bulk velocity");
    Uwall = args.getreal("-Uwall", "--Uwall", 1.0, "magnitude of imposed wall velocity, +/-Uwall at y = +/-h");
    theta = args.getreal("-theta", "--theta", 0.0, "angle of base flow relative to x-axis");
    Vsuck = args.getreal("-Vs", "--Vsuck", 0.0, "wall-normal suction velocity");
    rotation = args.getreal("-rot", "--rotation", 0.0, "rotation around the z-axis");
    const Real c = cos(theta), sn = sin(theta);
    ulowerwall = -Uwall * c; uupperwall = Uwall * c;
    wlowerwall = -Uwall * sn; wupperwall = Uwall * sn;
    dPdx = dPds * c; dPdz = dPds * sn;
    Ubulk = Ub * c; Wbulk = Ub * sn;
}
void DNSFlags::args2numerics(ArgList& args, const bool laurette) {
    args.section("Numerical setup");
    dealiasing = s2dealiasing(args.getstr("-da", "--dealiasing", "DealiasXZ", "define dealiasing behavior, one of [NoDealiasing|DealiasXZ|DealiasY|DealiasXYZ]"));
    t0 = args.getreal("-T0", "--T0", 0.0, "start time of DNS or period of map f^T(u)");
    T = args.getreal("-T", "--T1", 20.0, "final time of DNS or period of map f^T(u)");
    dT = args.getreal("-dT", "--dT", 1.0, "save interval");
    dt = args.getreal("-dt", "--dt", 0.03125, "timestep");
    variabledt = args.getbool("-vdt", "--variabledt", true, "adjust dt to keep CFLmin<=CFL<CFLmax");
    dtmin = args.getreal("-dtmin", "--dtmin", 0.001, "minimum time step");
    dtmax = args.getreal("-dtmax", "--dtmax", 0.2, "maximum time step");
    CFLmin = args.getreal("-CFLmin", "--CFLmin", 0.40, "minimum CFL number");
    CFLmax = args.getreal("-CFLmax", "--CFLmax", 0.60, "maximum CFL number");
    timestepping = s2stepmethod(args.getstr("-ts", "--timestepping", "sbdf3", "timestepping algorithm,  one of [cnfe1|cnab2|cnrk2|smrk2|sbdf1|sbdf2|sbdf3|sbdf4]"));
    initstepping = s2stepmethod(args.getstr("-is", "--initstepping", "smrk2",
                                            "timestepping algorithm for initializing multistep algorithms,  one of [cnfe1|cnrk2|smrk2|sbdf1]"));
    nonlinearity = s2nonlmethod(args.getstr("-nl", "--nonlinearity", "rot", "method of calculating nonlinearity, one of [rot|conv|div|skew|alt|linear]"));
    symmetryprojectioninterval = args.getint("-symmpi", "--symmetryprojection", 100, "project onto symmetries at this time interval");
    const std::string symmstr = args.getstr("-symms", "--symmetries", "",
                                            "constrain u(t) to invariant symmetric subspace, argument is the filename for a file listing the "
                                            "generators of the isotropy group");
    verbosity = Silent;
    if (!symmstr.empty()) {
        symmetries_file = symmstr;
        symmetries = SymmetryList(symmstr);
    }
    if (laurette) {
        dT = T; dt = T; dtmax = T;
        variabledt = false;
        initstepping = timestepping = SBDF1;
    }
}
// one "value  %name" line per flag, the form DNSFlags::load reads back (dnsflags.cpp:610-740)
void DNSFlags::save(const std::string& outdir) const {
    if (mpirank() != 0) return;
    std::ofstream os((pathfix(outdir) + "dnsflags.txt").c_str());
    if (!os.good()) cferror("DNSFlags::save(outdir) :  can't open file " + outdir + "dnsflags.txt");
    os << std::setprecision(REAL_DIGITS);
    auto line = [&](auto v, const char* name) { os << std::left << std::setw(REAL_IOWIDTH) << v << "  %" << name << "\n"; };
    line(nu, "nu"); line(Vsuck, "Vsuck"); line(rotation, "rotation"); line(theta, "theta");
    line(dPdx, "dPdx"); line(dPdz, "dPdz"); line(Ubulk, "Ubulk"); line(Wbulk, "Wbulk"); line(Uwall, "Uwall");
    line(ulowerwall, "ulowerwall"); line(uupperwall, "uupperwall"); line(wlowerwall, "wlowerwall"); line(wupperwall, "wupperwall");
    line(t0, "t0"); line(T, "T"); line(dT, "dT"); line(dt, "dt"); line(variabledt, "variabledt");
    line(dtmin, "dtmin"); line(dtmax, "dtmax"); line(CFLmin, "CFLmin"); line(CFLmax, "CFLmax");
    line(symmetryprojectioninterval, "symmetryprojectioninterval");
    line(baseflow, "baseflow"); line(constraint, "constraint"); line(timestepping, "timestepping"); line(initstepping, "initstepping");
    line(nonlinearity, "nonlinearity"); line(dealiasing, "dealiasing");
    line(bodyforce ? "nonzero_bodyforce" : "zero_bodyforce", "bodyforce");
    line(taucorrection, "taucorrection"); line(verbosity, "verbosity");
}
void DNSFlags::load(int, const std::string indir) {
    std::ifstream is((pathfix(indir) + "dnsflags.txt").c_str());
    if (!is.good()) cferror("DNSFlags::load(taskid, indir): can't open file " + indir + "dnsflags.txt");
    std::string value, name;
    while (is >> value >> name) {
        const Real r = std::atof(value.c_str());
        if (name == "%nu") nu = r; else if (name == "%Vsuck") Vsuck = r; else if (name == "%rotation") rotation = r;
        else if (name == "%theta") theta = r; else if (name == "%dPdx") dPdx = r; else if (name == "%dPdz") dPdz = r;
        else if (name == "%Ubulk") Ubulk = r; else if (name == "%Wbulk") Wbulk = r; else if (name == "%Uwall") Uwall = r;
        else if (name == "%ulowerwall") ulowerwall = r; else if (name == "%uupperwall") uupperwall = r;
        else if (name == "%wlowerwall") wlowerwall = r; else if (name == "%wupperwall") wupperwall = r;
        else if (name == "%t0") t0 = r; else if (name == "%T") T = r; else if (name == "%dT") dT = r; else if (name == "%dt") dt = r;
        else if (name == "%variabledt") variabledt = value != "0"; else if (name == "%dtmin") dtmin = r; else if (name == "%dtmax") dtmax = r;
        else if (name == "%CFLmin") CFLmin = r; else if (name == "%CFLmax") CFLmax = r;
        else if (name == "%symmetryprojectioninterval") symmetryprojectioninterval = (int)r;
        else if (name == "%baseflow") baseflow = s2baseflow(value); else if (name == "%constraint") constraint = s2constraint(value);
        else if (name == "%timestepping") timestepping = s2stepmethod(value); else if (name == "%initstepping") initstepping = s2stepmethod(value);
        else if (name == "%nonlinearity") nonlinearity = s2nonlmethod(value); else if (name == "%dealiasing") dealiasing = s2dealiasing(value);
        else if (name == "%bodyforce") { if (value != "zero_bodyforce") cferror("bodyforce not zero, this is not implemented for a restart."); }
        else if (name == "%taucorrection") taucorrection = value != "0"; else if (name == "%verbosity") verbosity = s2verbosity(value);
    }
}

std::ostream& operator<<(std::ostream& os, const DNSFlags& f) {
    const char* s = ", ";
    const auto p = os.precision();
    os.precision(16);
    // one record per run in the reference's field order (dnsflags.cpp:976-993): logs of the two builds can be diffed
    const std::pair<const char*, Real> nums[] = {{"nu", f.nu}, {"Vsuck", f.Vsuck}, {"rotation", f.rotation}, {"theta", f.theta}, {"dPdx", f.dPdx},
        {"dPdz", f.dPdz}, {"Ubulk", f.Ubulk}, {"Wbulk", f.Wbulk}, {"uwall", f.Uwall}, {"uupper", f.uupperwall}, {"ulower", f.ulowerwall},
        {"wupper", f.wupperwall}, {"wlower", f.wlowerwall}, {"t0", f.t0}, {"dT", f.dT}, {"dt", f.dt}};
    for (const auto& kv : nums) os << kv.first << "==" << kv.second << s;
    os << "variabledt==" << f.variabledt << s << "dtmin==" << f.dtmin << s << "dtmax==" << f.dtmax << s << "CFLmin==" << f.CFLmin << s
       << "CFLmax==" << f.CFLmax << s << f.baseflow << s << f.constraint << s << f.timestepping << s << f.initstepping << s << f.nonlinearity
       << s << f.dealiasing << s << (f.bodyforce ? "nonzero_bodyforce" : "zero_bodyforce") << s
       << (f.taucorrection ? "TauCorrection" : "NoTauCorrection") << s << f.verbosity;
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
