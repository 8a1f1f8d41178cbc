// Host classes of the 1-d wall-normal solvers -- BandedTridiag, HelmholtzSolver, TauSolver -- over the device back ends
// cfgpu_tridiag / cfgpu_helmholtz_solve / cfgpu_tausolve_mode (reference channelflow/bandedtridiag.cpp, helmholtz.cpp,
// tausolver.cpp).  In the time stepper these solvers exist only as batched kernels over all Fourier modes (csrc/tau.cu);
// the classes here are the single-system view of the same kernels that the reference's unit tests and tools program
// against.  Residual checks (verify / residual) are host-side Chebyshev calculus on purpose: they must not share code
// with the solver they check.
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>

#include "cfgpu.h"
#include "channelflow/flowfield.h"
#include "channelflow/tausolver.h"

using namespace std;

namespace chflow {

namespace {
void dev(int rc, const char* what) {
    if (rc != 0) cferror(string(what) + ": " + cfgpu_last_error());
}
}  // namespace

// =========================================================================================================== BandedTridiag
BandedTridiag::BandedTridiag(int M) : M_(M), a_(M > 0 ? 4 * M - 2 : 0, 0.0), invdiag_(M, 0.0) { assert(M >= 0); }

BandedTridiag::BandedTridiag(const BandedTridiag& A) : M_(A.M_), a_(A.a_), invdiag_(A.M_, 0.0), UL_(false) {}

bool BandedTridiag::operator==(const BandedTridiag& A) const {
    if (M_ != A.M_ || UL_ != A.UL_) return false;
    for (size_t n = 0; n < a_.size(); ++n)
        if (fabs(a_[n] - A.a_[n]) > 1e-13) {
            cout << "BandedTridiag::operator== failed on a[" << n << "]\n" << setprecision(REAL_DIGITS) << a_[n] << ' ' << A.a_[n] << endl;
            return false;
        }
    return true;
}

void BandedTridiag::ULdecomp() {
    if (M_ >= 2)
        dev(cfgpu_tridiag(cfgpu_context(), 0, M_, a_.data(), invdiag_.data(), nullptr, nullptr, 0, 0, 1), "BandedTridiag::ULdecomp");
    else if (M_ == 1)
        invdiag_[0] = 1.0 / a_[0];
    UL_ = true;
}

void BandedTridiag::ULsolveStrided(Vector& b, int offset, int stride) const {
    assert(UL_);
    if (!((offset == 0 || offset == 1) && (stride == 1 || stride == 2)))
        cferror("BandedTridiag::ULsolveStrided(Vector& b, int offset, int stride) : offset must be 0 or 1, stride 1 or 2");
    if (M_ == 1) { b[offset] /= a_[0]; return; }
    BandedTridiag* me = const_cast<BandedTridiag*>(this);  // the C-ABI takes the factors as in/out host arrays
    dev(cfgpu_tridiag(cfgpu_context(), 1, M_, me->a_.data(), me->invdiag_.data(), b.pointer(), nullptr, b.length(), offset, stride),
        "BandedTridiag::ULsolveStrided");
}

void BandedTridiag::multiplyStrided(const Vector& x, Vector& b, int offset, int stride) const {
    assert((offset == 0 || offset == 1) && (stride == 1 || stride == 2));
    if (M_ == 1) { b[offset] = a_[0] * x[offset]; return; }
    assert(b.length() >= x.length());
    BandedTridiag* me = const_cast<BandedTridiag*>(this);
    dev(cfgpu_tridiag(cfgpu_context(), 2, M_, me->a_.data(), me->invdiag_.data(), const_cast<Real*>(x.pointer()), b.pointer(), x.length(),
                      offset, stride),
        "BandedTridiag::multiplyStrided");
}

// text form: header "% M ul", then "i j Aij" for the dense row and for every stored element of rows 1..M-1
void BandedTridiag::save(const string& filebase) const {
    ofstream os((filebase + ".asc").c_str());
    os << setprecision(REAL_DIGITS) << "% " << M_ << ' ' << (UL_ ? 1 : 0) << endl;
    for (int j = 0; j < M_; ++j) os << "0 " << j << ' ' << band(j) << '\n';
    for (int i = 1; i < M_; ++i)
        for (int j = i - 1; j <= i + 1 && j < M_; ++j) os << i << ' ' << j << ' ' << elem(i, j) << '\n';
}

BandedTridiag::BandedTridiag(const string& filebase) {
    ifstream is;
    const string filename = ifstreamOpen(is, filebase, ".asc");
    char c = 0;
    is >> c;
    if (c != '%') cferror("BandedTridiag(filebase): bad header in file " + filename);
    int ul = 0;
    is >> M_ >> ul;
    UL_ = ul != 0;
    a_.assign(M_ > 0 ? 4 * M_ - 2 : 0, 0.0);
    invdiag_.assign(M_, 0.0);
    const int nnz = M_ <= 1 ? M_ : 4 * (M_ - 1);
    for (int n = 0; n < nnz; ++n) {
        int i, j;
        Real x;
        is >> i >> j >> x;
        elem(i, j) = x;
    }
    if (UL_)
        for (int i = 0; i < M_; ++i) invdiag_[i] = 1.0 / diag(i);
}

void BandedTridiag::print() const {
    cout << "[\n";
    for (int i = 0; i < M_; ++i) {
        for (int j = 0; j < M_; ++j) cout << ((i == 0 || (i - j <= 1 && j - i <= 1)) ? elem(i, j) : 0.0) << ' ';
        cout << ";\n";
    }
    cout << "]\n";
}

void BandedTridiag::ULprint() const {
    cout << "U = [\n";
    for (int i = 0; i < M_; ++i) {
        for (int j = 0; j < M_; ++j) cout << (i == 0 ? band(j) : i == j ? 1.0 : i == j - 1 ? updiag(i) : 0.0) << ' ';
        cout << ";\n";
    }
    cout << "]\nL = [\n";
    for (int i = 0; i < M_; ++i) {
        for (int j = 0; j < M_; ++j) cout << (i == j + 1 ? lodiag(i) : i == j ? diag(i) : 0.0) << ' ';
        cout << ";\n";
    }
    cout << "]\n";
}

// self check: A x = b solved back to x for a well-conditioned random matrix
void BandedTridiag::test() const {
    const int M = M_ > 1 ? M_ : 8;
    BandedTridiag A(M);
    Vector x(M), b(M);
    for (int i = 0; i < M; ++i) {
        x[i] = drand48();
        A.band(i) = 1.0 + 0.1 * drand48();
        A.diag(i) = 1.0 + 0.1 * drand48();
        if (i > 0) A.lodiag(i) = 0.1 * drand48();
        if (i > 0 && i < M - 1) A.updiag(i) = 0.1 * drand48();
    }
    A.multiply(x, b);
    A.ULdecomp();
    A.ULsolve(b);
    cout << "BandedTridiag::test: L1 error of UL solve == " << L1Norm(x - b) << endl;
}

// ========================================================================================================= HelmholtzSolver
HelmholtzSolver::HelmholtzSolver(int numberModes, Real a, Real b, Real lambda, Real nu)
    : nModes_(numberModes), a_(a), b_(b), lambda_(lambda), nu_(nu) {
    assert(nModes_ % 2 == 1 && nModes_ > 2);
}

void HelmholtzSolver::solve(ChebyCoeff& u, const ChebyCoeff& f, Real ua, Real ub) const {
    assert(f.state() == Spectral && f.length() == nModes_);
    if (u.length() != nModes_) u = ChebyCoeff(nModes_, a_, b_, Spectral);
    dev(cfgpu_helmholtz_solve(cfgpu_context(), nModes_, a_, b_, lambda_, nu_, 1, f.pointer(), &ua, &ub, u.pointer()), "HelmholtzSolver::solve");
    u.setState(Spectral);
}

// u = u1 + (mu/nu) u3 with  nu u1'' - lambda u1 = f (Dirichlet data)  and  nu u3'' - lambda u3 = nu (homogeneous):
// both solved in one launch, mu from the mean constraint, then the full problem once more with f + mu so that the result
// is the solver's own answer to that right-hand side (helmholtz.cpp:158-213)
void HelmholtzSolver::solve(ChebyCoeff& u, Real& mu, const ChebyCoeff& f, Real umean, Real ua, Real ub) const {
    assert(f.state() == Spectral && f.length() == nModes_);
    const int N = nModes_;
    vector<Real> rhs(2 * N, 0.0), sol(2 * N);
    for (int n = 0; n < N; ++n) rhs[n] = f[n];
    rhs[N] = nu_;
    const Real uas[2] = {ua, 0.0}, ubs[2] = {ub, 0.0};
    dev(cfgpu_helmholtz_solve(cfgpu_context(), N, a_, b_, lambda_, nu_, 2, rhs.data(), uas, ubs, sol.data()), "HelmholtzSolver::solve");
    ChebyCoeff part(N, a_, b_, Spectral);
    Real means[2];
    for (int c = 0; c < 2; ++c) {
        for (int n = 0; n < N; ++n) part[n] = sol[c * N + n];
        means[c] = part.mean();
    }
    mu = nu_ * (umean - means[0]) / means[1];
    ChebyCoeff g(f);
    g[0] += mu;
    solve(u, g, ua, ub);
}

void HelmholtzSolver::tau_residuals(const ChebyCoeff& u, const ChebyCoeff& f, Real& tau, Real& all) const {
    assert(u.length() == nModes_);
    const ChebyCoeff uyy = diff2(u);
    tau = all = 0.0;
    for (int n = 0; n < nModes_; ++n) {
        const Real r = fabs(nu_ * uyy[n] - lambda_ * u[n] - f[n]);
        all += r;
        if (n < nModes_ - 2) tau += r;
    }
}

Real HelmholtzSolver::residual(const ChebyCoeff& u, const ChebyCoeff& f, Real, Real) const {
    Real tau, all;
    tau_residuals(u, f, tau, all);
    return tau;
}

void HelmholtzSolver::verify(const ChebyCoeff& u, const ChebyCoeff& f, Real ua, Real ub, bool verbose) const {
    Real tau, all;
    tau_residuals(u, f, tau, all);
    Real norm = Greater(L1Norm(u), L1Norm(f));
    if (norm <= 1.0) norm = 1.0;
    const Real ea = fabs(u.eval_a() - ua), eb = fabs(u.eval_b() - ub);
    if (verbose)
        cerr << "Helmholtz::verify() { \nN nu lambda == " << nModes_ - 1 << ' ' << nu_ << ' ' << lambda_
             << "\ntauNorm(nu*uyy - lambda*u - f) == " << tau << "\n L1Norm(nu*uyy - lambda*u - f) == " << all
             << "\nfabs(uas - ua)      == " << ea << "\nfabs(ubs - ub)      == " << eb << "\n} Helmholtz::verify()" << endl;
    assert(tau / norm < 1.0 && ea / norm < 1.0 && eb / norm < 1.0);
    (void)norm;
}

void HelmholtzSolver::verify(ChebyCoeff& u, Real& mu, const ChebyCoeff& f, Real umean, Real ua, Real ub) const {
    cerr << "Helmholtz::verify(u,f,a,b,mu,umean) {" << endl;
    verify(u, f, ua, ub);
    cerr << "mu == " << mu << "\numean - mean(u) === " << umean - u.mean() << "\n} Helmholtz::verify(u,f,ua,ub,mu,umean)" << endl;
}

Real HelmholtzSolver::residual(const ChebyCoeff& u, Real mu, const ChebyCoeff& f, Real umean, Real ua, Real ub) const {
    ChebyCoeff g(f);
    g[0] += mu;
    return residual(u, g, ua, ub) + std::abs(umean - u.mean());
}

// =============================================================================================================== TauSolver
Real divcheck(string& label, int kx, int kz, Real kxLx, Real kzLz, const ComplexChebyCoeff& u, const ComplexChebyCoeff& v,
              const ComplexChebyCoeff& w, bool verbose) {
    const int N = u.length();
    if (v.length() != N || w.length() != N) {
        cout << "divcheck length problem!" << endl;
        exit(1);
    }
    const ComplexChebyCoeff vy = diff(v);
    Real s = 0.0;
    for (int n = N - 1; n >= 0; --n) s += abs2(vy[n] + (pi * 2 * (kxLx * u[n] + kzLz * w[n])) * I);
    s = sqrt(s);
    if (s > 1e-13 || verbose)
        cout << label << "\nkx, kz == " << kx << ", " << kz << "\nkxLx, kzLz == " << kxLx << ", " << kzLz << "\ndivergence == " << s << endl;
    return s;
}

TauSolver::TauSolver(int kx, int kz, Real Lx, Real Lz, Real a, Real b, Real lambda, Real nu, int Ny, bool tauCorrection)
    : N_(Ny), kx_(kx), kz_(kz), Lx_(Lx), Lz_(Lz), a_(a), b_(b), lambda_(lambda), nu_(nu), tauCorrection_(tauCorrection) {}

// One mode through the batched kernels.  A real field stores kz >= 0 only: the mode (kx, kz < 0) is solved as its
// conjugate partner (-kx, -kz) with conjugated data (the operator is real)
void TauSolver::device_solve(ComplexChebyCoeff& u, ComplexChebyCoeff& v, ComplexChebyCoeff& w, ComplexChebyCoeff& P, const ComplexChebyCoeff& Rx,
                             const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz, int constraint, Real umean, Real wmean, Real* dPd) const {
    const int N = N_;
    assert(Rx.length() == N && Ry.length() == N && Rz.length() == N);
    const bool conj = kz_ < 0;
    const Real s = conj ? -1.0 : 1.0;
    vector<Real> R(6 * N), out(8 * N);
    const ComplexChebyCoeff* Rs[3] = {&Rx, &Ry, &Rz};
    for (int i = 0; i < 3; ++i)
        for (int n = 0; n < N; ++n) {
            R[(i * N + n) * 2] = Rs[i]->re[n];
            R[(i * N + n) * 2 + 1] = s * Rs[i]->im[n];
        }
    dev(cfgpu_tausolve_mode(cfgpu_context(), N, conj ? -kx_ : kx_, conj ? -kz_ : kz_, Lx_, Lz_, a_, b_, lambda_, nu_, tauCorrection_ ? 1 : 0,
                            constraint, umean, wmean, R.data(), out.data(), dPd),
        "TauSolver::solve");
    ComplexChebyCoeff* outs[4] = {&u, &v, &w, &P};
    for (int i = 0; i < 4; ++i) {
        if (outs[i]->length() != N) *outs[i] = ComplexChebyCoeff(N, a_, b_, Spectral);
        for (int n = 0; n < N; ++n) outs[i]->set(n, Complex(out[(i * N + n) * 2], s * out[(i * N + n) * 2 + 1]));
        outs[i]->setState(Spectral);
    }
}

void TauSolver::solve(ComplexChebyCoeff& u, ComplexChebyCoeff& v, ComplexChebyCoeff& w, ComplexChebyCoeff& P, const ComplexChebyCoeff& Rx,
                      const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz) const {
    device_solve(u, v, w, P, Rx, Ry, Rz, 0, 0.0, 0.0, nullptr);
}

void TauSolver::solve(ComplexChebyCoeff& u, ComplexChebyCoeff& v, ComplexChebyCoeff& w, ComplexChebyCoeff& P, Real& dPdx, Real& dPdz,
                      const ComplexChebyCoeff& Rx, const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz, Real umean, Real wmean) const {
    assert(kx_ == 0 && kz_ == 0);
    Real dPd[2] = {0.0, 0.0};
    device_solve(u, v, w, P, Rx, Ry, Rz, 1, umean, wmean, dPd);
    dPdx = dPd[0];
    dPdz = dPd[1];
}

Real TauSolver::verify(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, const ComplexChebyCoeff& w, const ComplexChebyCoeff& P,
                       const ComplexChebyCoeff& Rx, const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz, bool verbose) const {
    return verify(u, v, w, P, 0.0, 0.0, Rx, Ry, Rz, Re(u.mean()), Re(w.mean()), verbose);
}

// Sum of the L2 residuals of the three momentum equations, the pressure Poisson equation and continuity, the boundary
// values of u, v, v' and the mean-velocity mismatches (the error measure of tausolver.cpp:470-625, which tausolverTest
// compares with its tolerance).
Real TauSolver::verify(const ComplexChebyCoeff& u, const ComplexChebyCoeff& v, const ComplexChebyCoeff& w, const ComplexChebyCoeff& P, Real dPdx,
                       Real dPdz, const ComplexChebyCoeff& Rx, const ComplexChebyCoeff& Ry, const ComplexChebyCoeff& Rz, Real umean, Real wmean,
                       bool verbose) const {
    const int N = N_;
    const Real ax = 2 * pi * kx_ / Lx_, az = 2 * pi * kz_ / Lz_;
    const Real kappa2 = 4 * square(pi) * (square(kx_ / Lx_) + square(kz_ / Lz_));
    auto trunc_dist = [N](const ComplexChebyCoeff& f, const ComplexChebyCoeff& g) {  // distance over the first N-2 modes
        return L2Dist(ComplexChebyCoeff(N - 2, f), ComplexChebyCoeff(N - 2, g));
    };
    Real error = 0.0;
    auto report = [&](const char* what, const ComplexChebyCoeff& lhs, const ComplexChebyCoeff& rhs) {
        const Real l2 = L2Dist(lhs, rhs);
        error += l2;
        if (verbose) cout << "L2Norm(rhs) == " << L2Norm(rhs) << "\ntauDist(" << what << ") == " << trunc_dist(lhs, rhs) << "\n L2Dist(" << what << ") == " << l2 << endl;
    };
    if (verbose) cout << "TauSolver::verify(u,v,w,P,dPdx,dPdz,Rx,Ry,Rz,umean,wmean,verbose) {\n kx kz == " << kx_ << ' ' << kz_ << endl;

    const ComplexChebyCoeff Py = diff(P);
    // momentum: lambda q - nu q'' + (grad P)_q (+ mean pressure gradient) = R_q
    auto momentum = [&](const ComplexChebyCoeff& q, int dir) {
        ComplexChebyCoeff lhs = diff2(q);
        lhs *= -nu_;
        for (int n = 0; n < N; ++n) {
            const Complex gradP = dir == 1 ? Py[n] : Complex(0.0, dir == 0 ? ax : az) * P[n];
            lhs.add(n, lambda_ * q[n] + gradP);
        }
        if (dir == 0) lhs.re[0] += dPdx;
        if (dir == 2) lhs.re[0] += dPdz;
        return lhs;
    };
    report("nu u'' - lambda u - dP/dx, -Rx", momentum(u, 0), Rx);
    report("nu v'' - lambda v - dP/dy, -Ry", momentum(v, 1), Ry);
    report("nu w'' - lambda w - dP/dz, -Rz", momentum(w, 2), Rz);

    // pressure: P'' - kappa^2 P = div R
    ComplexChebyCoeff lapP = diff(Py), divR = diff(Ry);
    for (int n = 0; n < N; ++n) {
        lapP.add(n, -kappa2 * P[n]);
        divR.add(n, I * (ax * Rx[n] + az * Rz[n]));
    }
    report("P'' - k^2 P, div R", lapP, divR);

    // continuity
    ComplexChebyCoeff divu = diff(v);
    for (int n = 0; n < N; ++n) divu.add(n, I * (ax * u[n] + az * w[n]));
    const Real l2div = L2Norm(divu);
    error += l2div;
    if (verbose) cout << " L2Norm(div) == " << l2div << endl;

    // boundary values (the reference's w check re-uses u's values: kept, it is part of the test's error measure)
    const ComplexChebyCoeff vy = diff(v);
    const Complex bvals[8] = {u.eval_a(), u.eval_b(), v.eval_a(), v.eval_b(), vy.eval_a(), vy.eval_b(), u.eval_a(), u.eval_b()};
    for (const Complex& c : bvals) error += abs(c);
    if (verbose)
        cout << "u(a),u(b) == " << bvals[0] << ' ' << bvals[1] << "\nv(a),v(b) == " << bvals[2] << ' ' << bvals[3] << "\nv' at a,b == " << bvals[4]
             << ' ' << bvals[5] << endl;

    const Real eu = abs2(Re(u.mean()) - umean), ew = abs2(Re(w.mean()) - wmean);
    error += eu + ew;
    if (verbose)
        cout << "abs2(u.mean() - umean) == " << eu << "\nabs2(w.mean() - wmean) == " << ew << "\ntotal verification error == " << error
             << "\n} TauSolver::verify(...)" << endl;
    return error;
}

void TauSolver::influenceCorrection(ChebyCoeff&, ChebyCoeff&) const {
    cferror("TauSolver::influenceCorrection: the device tau solver fuses the influence-matrix correction into solve(); no per-stage state exists");
}
void TauSolver::solve_P_and_v(ChebyCoeff&, ChebyCoeff&, const ChebyCoeff&, const ChebyCoeff&, Real&, Real&) const {
    cferror("TauSolver::solve_P_and_v: the device tau solver fuses the P/v stage into solve(); use solve()");
}
Real TauSolver::verify_P_and_v(const ChebyCoeff&, const ChebyCoeff&, const ChebyCoeff&, const ChebyCoeff&, Real, Real, bool) const {
    cferror("TauSolver::verify_P_and_v: the device tau solver fuses the P/v stage into solve(); use verify()");
    return 0.0;
}

}  // namespace chflow
