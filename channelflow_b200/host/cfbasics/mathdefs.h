// Basic numeric typedefs and helpers shared by the host classes.
// Mirrors the names of the reference's cfbasics/mathdefs.h:61-77,150-208 (Real, Complex, fieldstate, pi, ...).
#ifndef CFB200_MATHDEFS_H
#define CFB200_MATHDEFS_H
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <string>

namespace chflow {

typedef double Real;
typedef std::complex<double> Complex;
typedef std::ptrdiff_t lint;
typedef unsigned int uint;

enum fieldstate { Physical, Spectral };

const Real pi = 3.14159265358979323846264338327950288;
const Complex I(0.0, 1.0);

inline int Greater(int a, int b) { return a > b ? a : b; }
inline int lesser(int a, int b) { return a < b ? a : b; }
inline Real Greater(Real a, Real b) { return a > b ? a : b; }
inline Real lesser(Real a, Real b) { return a < b ? a : b; }
inline Real Re(const Complex& z) { return z.real(); }
inline Real Im(const Complex& z) { return z.imag(); }
inline Real square(Real x) { return x * x; }
inline int iround(Real x) { return int(x > 0.0 ? x + 0.5 : x - 0.5); }

// reference: cfbasics/cfbasics.h:272-278 (print, finalize, exit(1))
[[noreturn]] inline void cferror(const std::string& message) {
    std::cerr << message << std::endl;
    std::exit(1);
}

inline std::ostream& operator<<(std::ostream& os, fieldstate s) { return os << (s == Spectral ? 'S' : 'P'); }

}  // namespace chflow
#endif
