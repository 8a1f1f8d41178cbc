// See channelflow/utilfuncs.h (reference utilfuncs.cpp:23-63, 712-906).
#include "channelflow/utilfuncs.h"

#include <ctime>

namespace chflow {

void WriteProcessInfo(int argc, char* argv[], std::string filename, std::ios::openmode mode) {
    if (mpirank() != 0) return;
    std::ofstream os(filename.c_str(), mode);
    os << "Command:  ";
    for (int n = 0; n < argc; ++n) os << argv[n] << ' ';
    os << "\nPWD:      " << pwd() << '\n';
    char host[1024] = {0};
    gethostname(host, sizeof host - 1);
    os << "Host:     " << host << ", PID: " << getpid() << '\n';
    char stamp[80];
    const time_t now = time(nullptr);
    strftime(stamp, sizeof stamp, "%Y-%m-%d %I:%M:%S", localtime(&now));
    os << "Time:     " << stamp << '\n';
    os << "Version:  " << CHANNELFLOW_VERSION << "\nDevice:   B200 (sm_100a) through libcfgpu: " << cfgpu_version() << "\n\n\n";
}

// f -= s0 T0 + .. + s3 T3 so that f = f' = 0 at both walls (utilfuncs.cpp:730-756)
void fixDiriNeum(ChebyCoeff& f) {
    const Real ya = f.a(), yb = f.b();
    f.setBounds(-1, 1);
    const Real a = f.eval_a(), b = f.eval_b(), c = f.slope_a(), d = f.slope_b();
    f[0] -= 0.5 * (a + b) + 0.125 * (c - d);
    f[1] -= 0.5625 * (b - a) - 0.0625 * (c + d);
    f[2] -= 0.125 * (d - c);
    f[3] -= 0.0625 * (a - b + c + d);
    f.setBounds(ya, yb);
}
void fixDiriNeum(ComplexChebyCoeff& f) { fixDiriNeum(f.re); fixDiriNeum(f.im); }

FieldSeries::FieldSeries() : emptiness_(0) {}
FieldSeries::FieldSeries(int N) : t_(N), f_(N), emptiness_(N) {}
void FieldSeries::push(const FlowField& f, Real t) {
    for (int n = f_.N() - 1; n > 0; --n) {
        if (f_[n].congruent(f_[n - 1])) swap(f_[n], f_[n - 1]);
        else f_[n] = f_[n - 1];
        t_[n] = t_[n - 1];
    }
    if (f_.N() > 0) { f_[0] = f; t_[0] = t; }
    if (emptiness_ > 0) --emptiness_;
}
bool FieldSeries::full() const { return emptiness_ == 0; }
void FieldSeries::interpolate(FlowField& f, Real t) const {  // Lagrange weights in time, fields combined on the device
    if (!full()) cferror("FieldSeries::interpolate(Real t, FlowField& f) : FieldSeries is not completely initialized.");
    const int N = f_.N();
    f = f_[0];
    f.setToZero();
    for (int i = 0; i < N; ++i) {
        Real w = 1.0;
        for (int j = 0; j < N; ++j)
            if (j != i) w *= (t - t_[j]) / (t_[i] - t_[j]);
        f.add(w, f_[i]);
    }
}

Real tFromFilename(const std::string filename) {
    size_t b = filename.find_last_of('/');
    b = b == std::string::npos ? 0 : b + 1;
    while (b < filename.size() && !(std::isdigit((unsigned char)filename[b]) || filename[b] == '-')) ++b;
    size_t e = b;
    while (e < filename.size() && (std::isdigit((unsigned char)filename[e]) || filename[e] == '.' || filename[e] == '-')) ++e;
    std::string num = filename.substr(b, e - b);
    while (!num.empty() && num.back() == '.') num.pop_back();  // the dot of the extension
    return std::atof(num.c_str());
}
bool comparetimes(const std::string& s0, const std::string& s1) { return tFromFilename(s0) < tFromFilename(s1); }
void channelflowVersion(int& major, int& minor, int& update) { major = 2; minor = 0; update = 0; }

// option names and help texts: the command-line interface of the reference's programs (utilfuncs.cpp:823-862)
DNSFlags setBaseFlowFlags(ArgList& args, std::string& Uname, std::string& Wname) {
    args.section("Base flow options");
    const std::string bf = args.getstr("-bf", "--baseflow", "laminar", "set base flow to one of [zero|laminar|linear|parabolic|suction]");
    Uname = args.getstr("-ub", "--Ubase", "", "input baseflow file of arbitrary U-baseflow (takes precedence over -bf option)");
    Wname = args.getstr("-wb", "--Wbase", "", "input baseflow file of arbitrary W-baseflow (takes precedence over -bf option)");
    const Real Reynolds = args.getreal("-R", "--Reynolds", 400, "pseudo-Reynolds number == 1/nu");
    const Real nuarg = args.getreal("-nu", "--nu", 0, "kinematic viscosity (takes precedence over Reynolds, if nonzero)");
    const std::string mean = args.getstr("-mc", "--meanconstraint", "gradp", "fix one of two flow constraints [gradp|bulkv]");
    const Real dPds = args.getreal("-dPds", "--dPds", 0.0, "magnitude of imposed pressure gradient along streamwise s");
    const Real Ub = args.getreal("-Ubulk", "--Ubulk", 0.0, "magnitude of imposed bulk velocity");
    const Real Uw = args.getreal("-Uwall", "--Uwall", 1.0, "magnitude of imposed wall velocity, +/-Uwall at y = +/-h");
    const Real th = args.getreal("-theta", "--theta", 0.0, "angle of base flow relative to x-axis");
    const Real Vs = args.getreal("-Vs", "--Vsuck", 0.0, "wall-normal suction velocity");
    DNSFlags flags;
    flags.baseflow = s2baseflow(bf);
    flags.nu = nuarg != 0 ? nuarg : 1.0 / Reynolds;
    flags.constraint = s2constraint(mean);
    flags.theta = th; flags.Uwall = Uw; flags.Vsuck = Vs;
    flags.ulowerwall = -Uw * cos(th); flags.uupperwall = Uw * cos(th);
    flags.wlowerwall = -Uw * sin(th); flags.wupperwall = Uw * sin(th);
    flags.dPdx = dPds * cos(th); flags.dPdz = dPds * sin(th);
    flags.Ubulk = Ub * cos(th); flags.Wbulk = Ub * sin(th);
    return flags;
}
std::vector<ChebyCoeff> baseFlow(int Ny, Real a, Real b, DNSFlags& flags, std::string Uname, std::string Wname) {
    ChebyCoeff U(Ny, a, b, Spectral), W(Ny, a, b, Spectral);
    if (!Uname.empty() || !Wname.empty()) flags.baseflow = ArbitraryBase;
    switch (flags.baseflow) {
        case ZeroBase: std::cout << "Baseflow: zero" << std::endl; break;
        case LinearBase: std::cout << "Baseflow: linear" << std::endl; U[1] = 1; break;
        case ParabolicBase: std::cout << "Baseflow: parabolic" << std::endl; U[0] = 0.5; U[2] = -0.5; break;
        case SuctionBase:
            std::cout << "Baseflow: suction" << std::endl;
            U = laminarProfile(flags.nu, PressureGradient, 0, flags.Ubulk, flags.Vsuck, a, b, -0.5, 0.5, Ny);
            break;
        case LaminarBase:
            std::cout << "Baseflow: laminar" << std::endl;
            U = laminarProfile(flags.nu, flags.constraint, flags.dPdx, flags.Ubulk, flags.Vsuck, a, b, flags.ulowerwall, flags.uupperwall, Ny);
            W = laminarProfile(flags.nu, flags.constraint, flags.dPdz, flags.Wbulk, flags.Vsuck, a, b, flags.wlowerwall, flags.wupperwall, Ny);
            break;
        case ArbitraryBase:
            std::cout << "Baseflow: reading from file" << std::endl;
            if (!Uname.empty()) U = ChebyCoeff(Uname);
            if (!Wname.empty()) W = ChebyCoeff(Wname);
            break;
        default: cferror("Unknown base flow !!!");
    }
    return {U, W};
}

}  // namespace chflow
