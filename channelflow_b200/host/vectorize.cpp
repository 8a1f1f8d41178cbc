// field2vector / vector2field / fixDiri* (reference flowfield.cpp:4448-4760, utilfuncs.cpp:712-765): the map between a
// divergence-free, no-slip velocity field and the vector of its linearly independent real coefficients
// (Gibson, Halcrow & Cvitanovic 2008, table 1) that nsolver's Newton-Krylov / Arnoldi iterations work on.
// Host-side over the FlowField's host mirror: a de-aliased spectral field crosses PCIe as its retained box only
// (cfgpu_field_download_padded), so one call moves 44 % of the array once; a device-side pack kernel is a "next" row.
#include <cmath>

#include "channelflow/flowfield.h"

namespace chflow {

void fixDiri(ChebyCoeff& f) {
    const Real fa = f.eval_a(), fb = f.eval_b();
    const Real mean = 0.5 * (fb + fa), slop = 0.5 * (fb - fa);
    f[0] -= mean;
    f[1] -= slop;
}
void fixDiriMean(ChebyCoeff& f) {
    const Real fa = f.eval_a(), fb = f.eval_b(), fm = f.mean();
    f[0] -= 0.125 * (fa + fb) + 0.75 * fm;
    f[1] -= 0.5 * (fb - fa);
    f[2] -= 0.375 * (fa + fb) - 0.75 * fm;
}
void fixDiri(ComplexChebyCoeff& f) { fixDiri(f.re); fixDiri(f.im); }
void fixDiriMean(ComplexChebyCoeff& f) { fixDiriMean(f.re); fixDiriMean(f.im); }

int field2vector_size(const FlowField& u) {
    const int Kx = u.kxmaxDealiased(), Kz = u.kzmaxDealiased(), Ny = u.Ny();
    int N = 2 * (Ny - 2);
    N += Kx * (2 * (Ny - 2) + 2 * (Ny - 4));
    N += Kz * (2 * (Ny - 2) + 2 * (Ny - 4));
    N += 2 * Kx * Kz * (2 * (Ny - 2) + 2 * (Ny - 4));
    return N;
}

void field2vector(const FlowField& u, Real* a) {
    assert(u.xzstate() == Spectral && u.ystate() == Spectral);
    assert(u.Nd() == 3);
    const int Kx = u.kxmaxDealiased(), Kz = u.kzmaxDealiased(), Ny = u.Ny();
    int n = 0;
    for (int ny = 2; ny < Ny; ++ny) a[n++] = u.cmplx(0, ny, 0, 0).real();
    for (int ny = 2; ny < Ny; ++ny) a[n++] = u.cmplx(0, ny, 0, 2).real();
    for (int kx = 1; kx <= Kx; ++kx) {
        const int mx = u.mx(kx);
        for (int ny = 2; ny < Ny; ++ny) {
            a[n++] = u.cmplx(mx, ny, 0, 2).real();
            a[n++] = u.cmplx(mx, ny, 0, 2).imag();
        }
        for (int ny = 3; ny < Ny - 1; ++ny) {
            a[n++] = u.cmplx(mx, ny, 0, 0).real();
            a[n++] = u.cmplx(mx, ny, 0, 0).imag();
        }
    }
    for (int kz = 1; kz <= Kz; ++kz) {
        const int mz = u.mz(kz);
        for (int ny = 2; ny < Ny; ++ny) {
            a[n++] = u.cmplx(0, ny, mz, 0).real();
            a[n++] = u.cmplx(0, ny, mz, 0).imag();
        }
        for (int ny = 3; ny < Ny - 1; ++ny) {
            a[n++] = u.cmplx(0, ny, mz, 2).real();
            a[n++] = u.cmplx(0, ny, mz, 2).imag();
        }
    }
    for (int kx = -Kx; kx <= Kx; ++kx) {
        if (kx == 0) continue;
        const int mx = u.mx(kx);
        for (int kz = 1; kz <= Kz; ++kz) {
            const int mz = u.mz(kz);
            for (int ny = 2; ny < Ny; ++ny) {
                a[n++] = u.cmplx(mx, ny, mz, 0).real();
                a[n++] = u.cmplx(mx, ny, mz, 0).imag();
            }
            for (int ny = 3; ny < Ny - 1; ++ny) {
                a[n++] = u.cmplx(mx, ny, mz, 2).real();
                a[n++] = u.cmplx(mx, ny, mz, 2).imag();
            }
        }
    }
}

namespace {
Complex eval_a(const ComplexChebyCoeff& f) { return Complex(f.re.eval_a(), f.im.eval_a()); }
Complex eval_b(const ComplexChebyCoeff& f) { return Complex(f.re.eval_b(), f.im.eval_b()); }
Complex mean(const ComplexChebyCoeff& f) { return Complex(f.re.mean(), f.im.mean()); }
void sub(ComplexChebyCoeff& f, int n, Complex c) { f.re[n] -= c.real(); f.im[n] -= c.imag(); }
void mul(ComplexChebyCoeff& f, Complex c) {
    for (int n = 0; n < f.length(); ++n) f.set(n, f[n] * c);
}
void integrate(const ComplexChebyCoeff& d, ComplexChebyCoeff& u) {
    chflow::integrate(d.re, u.re);
    chflow::integrate(d.im, u.im);
}
void store(FlowField& u, int mx, int mz, int i, const ComplexChebyCoeff& f) {
    for (int ny = 0; ny < u.Ny(); ++ny) u.cmplx(mx, ny, mz, i) = f[ny];
}
}  // namespace

void vector2field(const Real* a, FlowField& u) {
    assert(u.Nd() == 3);
    u.setToZero();
    u.setState(Spectral, Spectral);
    const int Kx = u.kxmaxDealiased(), Kz = u.kzmaxDealiased(), Ny = u.Ny();
    const Real ya = u.a(), yb = u.b(), Lx = u.Lx(), Lz = u.Lz();
    int n = 0;
    {   // (0,0) Fourier mode
        ComplexChebyCoeff f0(Ny, ya, yb, Spectral), f2(Ny, ya, yb, Spectral);
        for (int ny = 2; ny < Ny; ++ny) f0.re[ny] = a[n++];
        fixDiri(f0.re);
        store(u, 0, 0, 0, f0);
        for (int ny = 2; ny < Ny; ++ny) f2.re[ny] = a[n++];
        fixDiri(f2.re);
        store(u, 0, 0, 2, f2);
    }
    for (int kx = 1; kx <= Kx; ++kx) {  // (kx,0), kx > 0, and their conjugates at (-kx,0)
        const int mx = u.mx(kx);
        ComplexChebyCoeff f0(Ny, ya, yb, Spectral), f1(Ny, ya, yb, Spectral), f2(Ny, ya, yb, Spectral);
        for (int ny = 2; ny < Ny; ++ny) {
            f2.re[ny] = a[n++];
            f2.im[ny] = a[n++];
        }
        fixDiri(f2);
        for (int ny = 3; ny < Ny - 1; ++ny) {
            f0.re[ny] = a[n++];
            f0.im[ny] = a[n++];
        }
        f0.re[Ny - 1] = 0.0;
        f0.im[Ny - 1] = 0.0;
        fixDiriMean(f0);
        integrate(f0, f1);
        sub(f1, 0, 0.5 * (eval_a(f1) + eval_b(f1)));
        mul(f1, Complex(0.0, -(2 * pi * kx) / Lx));
        store(u, mx, 0, 0, f0);
        store(u, mx, 0, 1, f1);
        store(u, mx, 0, 2, f2);
        const int mxm = u.mx(-kx);
        for (int ny = 0; ny < Ny; ++ny) {
            u.cmplx(mxm, ny, 0, 0) = std::conj(f0[ny]);
            u.cmplx(mxm, ny, 0, 1) = std::conj(f1[ny]);
            u.cmplx(mxm, ny, 0, 2) = std::conj(f2[ny]);
        }
    }
    for (int kz = 1; kz <= Kz; ++kz) {  // (0,kz), kz > 0
        const int mz = u.mz(kz);
        ComplexChebyCoeff f0(Ny, ya, yb, Spectral), f1(Ny, ya, yb, Spectral), f2(Ny, ya, yb, Spectral);
        for (int ny = 2; ny < Ny; ++ny) {
            f0.re[ny] = a[n++];
            f0.im[ny] = a[n++];
        }
        fixDiri(f0);
        for (int ny = 3; ny < Ny - 1; ++ny) {
            f2.re[ny] = a[n++];
            f2.im[ny] = a[n++];
        }
        f2.re[Ny - 1] = 0.0;
        f2.im[Ny - 1] = 0.0;
        fixDiriMean(f2);
        integrate(f2, f1);
        sub(f1, 0, 0.5 * (eval_a(f1) + eval_b(f1)));
        mul(f1, Complex(0.0, -(2 * pi * kz) / Lz));
        store(u, 0, mz, 0, f0);
        store(u, 0, mz, 1, f1);
        store(u, 0, mz, 2, f2);
    }
    for (int kx = -Kx; kx <= Kx; ++kx) {
        if (kx == 0) continue;
        const int mx = u.mx(kx);
        for (int kz = 1; kz <= Kz; ++kz) {
            const int mz = u.mz(kz);
            ComplexChebyCoeff f0(Ny, ya, yb, Spectral), f1(Ny, ya, yb, Spectral), f2(Ny, ya, yb, Spectral);
            for (int ny = 2; ny < Ny; ++ny) {
                f0.re[ny] = a[n++];
                f0.im[ny] = a[n++];
            }
            fixDiri(f0);
            store(u, mx, mz, 0, f0);
            for (int ny = 3; ny < Ny - 1; ++ny) {
                f2.re[ny] = a[n++];
                f2.im[ny] = a[n++];
            }
            f2.re[Ny - 1] = -f0.re[Ny - 1] * (kx * Lz) / (kz * Lx);
            f2.im[Ny - 1] = -f0.im[Ny - 1] * (kx * Lz) / (kz * Lx);
            // adjust coefficients 0,1,2 of f2 so that f2(+-1) = 0 and kz/Lz mean(f2) + kx/Lx mean(f0) = 0
            const Complex f2a = eval_a(f2), f2b = eval_b(f2), f0m = mean(f0);
            const Complex f2m = mean(f2) + (kx * Lz) / (kz * Lx) * f0m;
            sub(f2, 0, 0.125 * (f2a + f2b) + 0.75 * f2m);
            sub(f2, 1, 0.5 * (f2b - f2a));
            sub(f2, 2, 0.375 * (f2a + f2b) - 0.75 * f2m);
            store(u, mx, mz, 2, f2);
            mul(f0, Complex(0, -2 * pi * kx / Lx));
            mul(f2, Complex(0, -2 * pi * kz / Lz));
            f0.re += f2.re;
            f0.im += f2.im;
            integrate(f0, f1);
            sub(f1, 0, 0.5 * (eval_a(f1) + eval_b(f1)));
            store(u, mx, mz, 1, f1);
        }
    }
    u.setPadded(true);
}

void fixdivnoslip(FlowField& u) {
    std::vector<Real> v(field2vector_size(u));
    field2vector(u, v.data());
    vector2field(v.data(), u);
}

}  // namespace chflow
