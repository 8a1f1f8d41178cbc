// Basic numeric typedefs and helpers shared by the host classes.
// Same names as the reference's cfbasics/mathdefs.h (Real, Complex, lint, fieldstate, parity, pi, I, small inline
// helpers, randomReal / randomComplex over the libc drand48 stream, stream operators) so that code written against
// Channelflow compiles unchanged; there is no MPI here (one process per GPU), so lint is ptrdiff_t and MPI_Comm a stub.
#ifndef CFB200_MATHDEFS_H
#define CFB200_MATHDEFS_H
#include <arpa/inet.h>

#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>

#include "channelflow/config.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef DEBUG
#ifdef NDEBUG
#define DEBUG 0
#else
#define DEBUG 1
#endif
#endif

namespace chflow {

typedef double Real;
typedef std::complex<double> Complex;
typedef std::ptrdiff_t lint;
typedef unsigned int uint;
typedef int MPI_Comm;
static const int MPI_COMM_WORLD = 0;

enum fieldstate { Physical, Spectral };
enum parity { Even, Odd };

const int REAL_DIGITS = 17;
const int REAL_IOWIDTH = 24;
const Real pi = 3.14159265358979323846264338327950288;
const Complex I(0.0, 1.0);

// print the message and terminate (reference: cfbasics/cfbasics.h:272-278)
[[noreturn]] inline void cferror(const std::string& message) {
    std::cerr << message << std::endl;
    std::exit(1);
}

inline int kronecker(int m, int n) { return m == n ? 1 : 0; }
inline int square(int x) { return x * x; }
inline int cube(int x) { return x * x * x; }
inline Real square(Real x) { return x * x; }
inline Real cube(Real x) { return x * x * x; }
inline Real nr_sign(Real a, Real b) { return b >= 0.0 ? std::fabs(a) : -std::fabs(a); }
inline void swap(int& a, int& b) { const int t = a; a = b; b = t; }
inline void swap(Real& a, Real& b) { const Real t = a; a = b; b = t; }
inline void swap(Complex& a, Complex& b) { const Complex t = a; a = b; b = t; }
inline int Greater(int a, int b) { return a > b ? a : b; }
inline int lesser(int a, int b) { return a < b ? a : b; }
inline Real Greater(Real a, Real b) { return a > b ? a : b; }
inline Real lesser(Real a, Real b) { return a < b ? a : b; }
inline Real Re(const Complex& z) { return z.real(); }
inline Real Im(const Complex& z) { return z.imag(); }
inline Real abs2(const Complex& z) { return std::norm(z); }
inline int iround(Real x) { return int(x > 0.0 ? x + 0.5 : x - 0.5); }
inline int intpow(int x, int n) {
    if (n < 0) cferror("int pow(int, int) : can't do negative exponents, use Real pow(Real,int)");
    int r = 1;
    for (; n > 0; n >>= 1, x *= x)
        if (n & 1) r *= x;
    return r;
}
// sqrt(a^2 + b^2) without overflow
inline Real pythag(Real a, Real b) {
    const Real p = std::fabs(a), q = std::fabs(b);
    if (p > q) return p * std::sqrt(1.0 + square(q / p));
    if (q > p) return q * std::sqrt(1.0 + square(p / q));
    return 0.0;  // (sic: equal magnitudes give 0 in the reference too)
}
// signed distance of (Lx,Lz) from a target along the direction phi
inline Real spythag(Real Lx, Real Lxtarg, Real Lz, Real Lztarg, Real phi) {
    const Real dLx = Lxtarg - Lx, dLz = Lztarg - Lz;
    const Real proj = std::cos(phi) * dLx + std::sin(phi) * dLz;
    return (proj > 0 ? 1 : (proj < 0 ? -1 : 0)) * pythag(dLx, dLz);
}
inline bool isPowerOfTwo(int n) { return n >= 0 && (n & (n - 1)) == 0; }

// the reference draws every random number from the libc drand48 stream (serial build): same stream, same fields
inline Real randomReal(Real a = 0, Real b = 1) { return a + (b - a) * drand48(); }
inline Complex randomComplex() {  // gaussian about zero (polar Box-Muller on the drand48 stream)
    Real a, b, r2;
    do {
        a = randomReal(-1, 1);
        b = randomReal(-1, 1);
        r2 = a * a + b * b;
    } while (r2 >= 1.0 || r2 == 0.0);
    return std::sqrt(-std::log(r2) / r2) * Complex(a, b);
}

inline std::ostream& operator<<(std::ostream& os, Complex z) { return os << '(' << z.real() << ", " << z.imag() << ')'; }
inline std::ostream& operator<<(std::ostream& os, fieldstate s) { return os << (s == Spectral ? 'S' : 'P'); }
inline std::istream& operator>>(std::istream& is, fieldstate& s) {
    char c = ' ';
    while (c == ' ' && is.good()) is >> c;
    if (c == 'P') s = Physical;
    else if (c == 'S') s = Spectral;
    else {
        std::cerr << "read fieldstate error: unknown fieldstate " << c << std::endl;
        s = Spectral;
    }
    return is;
}

}  // namespace chflow
#endif
