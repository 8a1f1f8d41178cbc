// Small utilities of the Channelflow API (reference cfbasics/cfbasics.h): string helpers, file-name helpers, the
// big-endian binary primitives of the .ff / .bf file formats, directory helpers, timers, 1-d interpolation.  The Eigen
// based parts of the reference header (Newton-hookstep linear algebra) are not here: the solver in this package keeps its
// Krylov vectors on the device (channelflow/flowfield.h: DeviceVector; host/nsolver.*).
#ifndef CFB200_CFBASICS_H
#define CFB200_CFBASICS_H
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cfbasics/cfarray.h"
#include "cfbasics/mathdefs.h"

namespace chflow {

enum HookstepPhase { ConstantDelta, ReducingDelta, IncreasingDelta, Finished };
enum ResidualImprovement { Unacceptable, Poor, Ok, Good, Accurate, NegaCurve };
enum SolverMethod { SolverEigen, SolverGMRES, SolverFGMRES, SolverBiCGStab };
enum OptimizationMethod { None, Linear, Hookstep };
enum SolutionType { Equilibrium, PeriodicOrbit };
enum fEvalType { fEval, DfEval, HookstepEval };

inline int mpirank() {  // rank this process was launched with (0 in single-process runs)
    for (const char* n : {"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"})
        if (const char* v = std::getenv(n)) return std::atoi(v);
    return 0;
}
inline std::string r2s(Real r) {
    char b[32];
    std::snprintf(b, sizeof b, "%g", r);
    return b;
}
inline std::string i2s(int n, int length = 0, char pad = '0') {
    std::string s = std::to_string(n);
    if ((int)s.size() < length) s.insert(0, length - s.size(), pad);
    return s;
}
inline void cfpause() {
    if (mpirank() == 0) {
        std::cout << "cfpause..." << std::flush;
        std::string s;
        std::getline(std::cin, s);
    }
}
inline bool fileExists(const std::string& filename) {
    struct stat st;
    return ::stat(filename.c_str(), &st) == 0;
}
inline std::string FillZeros(int i, int n) { return i2s(i, n, '0'); }
inline std::string FillZeros(Real t, int n) {
    std::string s = r2s(t);
    if ((int)s.size() < n) s.insert(0, n - s.size(), '0');
    return s;
}
inline std::string pwd() {
    char b[4096];
    return ::getcwd(b, sizeof b) ? std::string(b) : std::string();
}
inline std::string appendSuffix(const std::string& filebase, const std::string& ext) {
    const size_t L = filebase.size(), E = ext.size();
    return (L < E || filebase.compare(L - E, E, ext) != 0) ? filebase + ext : filebase;
}
inline std::string removeSuffix(const std::string& filename, const std::string& ext) {
    const size_t p = filename.find(ext);
    return p == std::string::npos ? filename : filename.substr(0, p);
}
inline bool hasSuffix(const std::string& filename, const std::string& ext) { return filename.find(ext) != std::string::npos; }
inline bool isReadable(const std::string& filename) {
    std::ifstream is(filename.c_str(), std::ios::in | std::ios::binary);
    return (bool)is;
}
// open <filebase>, else <filebase><ext>; returns the name tried last
inline std::string ifstreamOpen(std::ifstream& is, const std::string& filebase, const std::string& ext,
                                std::ios_base::openmode mode = std::ios::in) {
    mode |= std::ios::in;
    is.open(filebase.c_str(), mode);
    if (is) return filebase;
    is.close();
    is.clear();
    const std::string name = filebase + ext;
    is.open(name.c_str(), mode);
    if (!is) is.close();
    return name;
}

// ---- binary primitives: big-endian on disk, 32-bit int, IEEE double (the .ff / .bf formats)
inline void write(std::ostream& os, int n) {
    const uint32_t v = htonl((uint32_t)n);
    os.write(reinterpret_cast<const char*>(&v), 4);
}
inline void read(std::istream& is, int& n) {
    uint32_t v = 0;
    is.read(reinterpret_cast<char*>(&v), 4);
    n = (int)ntohl(v);
}
inline void write(std::ostream& os, Real x) {
    unsigned char b[8];
    std::memcpy(b, &x, 8);
    if (htonl(1) != 1) std::reverse(b, b + 8);
    os.write(reinterpret_cast<const char*>(b), 8);
}
inline void read(std::istream& is, Real& x) {
    unsigned char b[8];
    is.read(reinterpret_cast<char*>(b), 8);
    if (htonl(1) != 1) std::reverse(b, b + 8);
    std::memcpy(&x, b, 8);
}
inline void write(std::ostream& os, bool b) { const char c = b ? '1' : '0'; os.write(&c, 1); }
inline void read(std::istream& is, bool& b) { char c = '0'; is.read(&c, 1); b = c != '0'; }
inline void write(std::ostream& os, fieldstate s) { const char c = s == Spectral ? 'S' : 'P'; os.write(&c, 1); }
inline void read(std::istream& is, fieldstate& s) { char c = 'S'; is.read(&c, 1); s = c == 'S' ? Spectral : Physical; }
inline void write(std::ostream& os, Complex z) { write(os, z.real()); write(os, z.imag()); }
inline void read(std::istream& is, Complex& z) { Real a = 0, b = 0; read(is, a); read(is, b); z = Complex(a, b); }

// ---- scalars in ascii files
inline void save(Real c, const std::string& filebase) {
    if (mpirank() != 0) return;
    std::ofstream os(appendSuffix(filebase, ".asc").c_str());
    if (!os.good()) cferror("save(Real, filebase) :  can't open file " + filebase);
    os << std::setprecision(REAL_DIGITS) << c << '\n';
}
inline void load(Real& c, const std::string& filebase) {
    std::ifstream is(appendSuffix(filebase, ".asc").c_str());
    if (!is.good()) cferror("load(Real, filebase) :  can't open file " + filebase + ".asc");
    is >> c;
}
inline void save(Complex c, const std::string& filebase) {
    if (mpirank() != 0) return;
    std::ofstream os(appendSuffix(filebase, ".asc").c_str());
    if (!os.good()) cferror("save(Complex, filebase) :  can't open file " + filebase);
    os << std::setprecision(REAL_DIGITS) << c.real() << ' ' << c.imag() << '\n';
}
inline void load(Complex& c, const std::string& filebase) {
    std::ifstream is(appendSuffix(filebase, ".asc").c_str());
    if (!is.good()) cferror("load(Complex, filebase) :  can't open file " + filebase + ".asc");
    Real a = 0, b = 0;
    is >> a >> b;
    c = Complex(a, b);
}

inline void mkdir(const std::string& dirname) {
    if (mpirank() == 0) ::mkdir(dirname.c_str(), 0755);
}
inline void rename(const std::string& oldname, const std::string& newname) {
    if (mpirank() == 0) ::rename(oldname.c_str(), newname.c_str());
}
inline std::string pathfix(const std::string& path) { return (!path.empty() && path.back() != '/') ? path + "/" : path; }
inline std::string t2s(Real t, bool inttime) {  // the reference declares the flag as `decimals` and defines it as `inttime`
    char b[32];
    if (inttime) std::snprintf(b, sizeof b, "%d", iround(t));
    else std::snprintf(b, sizeof b, "%.3f", t);
    return b;
}
inline std::string clip(const std::string& filename, const std::string& ext) { return removeSuffix(filename, ext); }
inline std::string stub(const std::string& filename, const std::string& ext) {
    const std::string f = removeSuffix(filename, ext);
    const size_t s = f.find_last_of('/');
    return s == std::string::npos ? f : f.substr(s + 1);
}
inline Real executionTime() {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    static const double t0 = tv.tv_sec + 1e-6 * tv.tv_usec;
    return tv.tv_sec + 1e-6 * tv.tv_usec - t0;
}
inline Real linearInterpolate(Real x0, Real y0, Real x1, Real y1, Real x) { return y0 + (y1 - y0) * (x - x0) / (x1 - x0); }
// Lagrange interpolation through all (xn, fn)
inline Real polynomialInterpolate(const cfarray<Real>& fn, const cfarray<Real>& xn, Real x) {
    Real s = 0.0;
    for (int i = 0; i < fn.length(); ++i) {
        Real w = fn[i];
        for (int j = 0; j < fn.length(); ++j)
            if (j != i) w *= (x - xn[j]) / (xn[i] - xn[j]);
        s += w;
    }
    return s;
}
inline Real quadraticInterpolate(const cfarray<Real>& fn, const cfarray<Real>& xn, Real x) { return polynomialInterpolate(fn, xn, x); }
inline bool isconst(cfarray<Real> f, Real eps = 1e-13) {
    for (int i = 1; i < f.length(); ++i)
        if (std::fabs(f[i] - f[0]) > eps) return false;
    return true;
}
template <class T>
inline void push(const T& t, cfarray<T>& a) {
    for (int i = a.length() - 1; i > 0; --i) a[i] = a[i - 1];
    if (a.length() > 0) a[0] = t;
}
inline void openfile(std::ofstream& f, std::string filename, std::ios::openmode openflag = std::ios::out) {
    if (mpirank() == 0) f.open(filename.c_str(), openflag);
}
inline void printout(const std::string& message, int taskid, bool newline = true, std::ostream& os = std::cout) {
    if (taskid != 0) return;
    os << message;
    if (newline) os << std::endl;
    os << std::flush;
}
inline void printout(const std::string& message, bool newline = true, std::ostream& os = std::cout) {
    printout(message, mpirank(), newline, os);
}
inline void printout(const std::string& message, std::ostream& os) { printout(message, true, os); }
inline void printout(const std::stringstream& sstr, std::ostream& os = std::cout) { printout(sstr.str(), false, os); }
inline std::string FillSpaces(int i, int n) { return i2s(i, n, ' '); }   // FillSpaces(12, 5) = "   12"
inline std::string FillSpaces(Real t, int n) { return FillSpaces((int)t, n); }
inline std::string fuzzyless(Real x, Real eps) { return x < eps ? "TRUE  " : (x < std::sqrt(eps) ? "APPROX" : "false "); }
inline std::ostream& operator<<(std::ostream& os, HookstepPhase p) {
    static const char* n[] = {"ConstantDelta", "ReducingDelta", "IncreasingDelta", "Finished"};
    return os << n[(int)p];
}
inline std::ostream& operator<<(std::ostream& os, ResidualImprovement i) {
    static const char* n[] = {"Unacceptable", "Poor", "Ok", "Good", "Accurate", "NegativeCurvature"};
    return os << n[(int)i];
}
inline std::ostream& operator<<(std::ostream& os, SolutionType s) { return os << (s == Equilibrium ? "Equilibrium" : "PeriodicOrbit"); }

}  // namespace chflow

// ---- vector helpers of the reference header for Eigen::VectorXd (cfbasics.h:97-115, 711-780), available when an
// <Eigen/Dense> is on the include path (the real one, or channelflow_b200/host/compat)
#if defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
namespace chflow {
inline void setToZero(Eigen::VectorXd& x) { x.setZero(); }
inline Real L2Norm2(const Eigen::VectorXd& x, int cutoff = 0) {
    Real s = 0;
    for (long i = 0; i < (long)x.size() - cutoff; ++i) s += x(i) * x(i);
    return s;
}
inline Real L2Norm(const Eigen::VectorXd& x, int cutoff = 0) { return std::sqrt(L2Norm2(x, cutoff)); }
inline Real L2Dist2(const Eigen::VectorXd& x, const Eigen::VectorXd& y, int cutoff = 0) {
    Real s = 0;
    for (long i = 0; i < (long)x.size() - cutoff; ++i) s += (x(i) - y(i)) * (x(i) - y(i));
    return s;
}
inline Real L2Dist(const Eigen::VectorXd& x, const Eigen::VectorXd& y, int cutoff = 0) { return std::sqrt(L2Dist2(x, y, cutoff)); }
inline Real L2IP(const Eigen::VectorXd& x, const Eigen::VectorXd& y) { Real s = 0; for (long i = 0; i < (long)x.size(); ++i) s += x(i) * y(i); return s; }
inline void print(const Eigen::VectorXd& x) { for (long i = 0; i < (long)x.size(); ++i) std::cout << x(i) << '\n'; }
inline void save(const Eigen::VectorXd& x, const std::string& filebase) {
    std::ofstream os(appendSuffix(filebase, ".asc").c_str());
    os << std::setprecision(17) << x.size() << '\n';
    for (long i = 0; i < (long)x.size(); ++i) os << x(i) << '\n';
}
inline void load(Eigen::VectorXd& x, const std::string& filebase) {
    std::ifstream is(appendSuffix(filebase, ".asc").c_str());
    long n = 0;
    is >> n;
    x.resize(n);
    for (long i = 0; i < n; ++i) is >> x(i);
}
}  // namespace chflow
#endif
#endif

#endif
