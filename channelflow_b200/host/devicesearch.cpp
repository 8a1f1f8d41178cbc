// Device-resident Newton-Krylov-hookstep search (see channelflow/devicesearch.h for what it replaces).
#include "channelflow/devicesearch.h"

#include <algorithm>
#include <cmath>
#include <iomanip>

#include "channelflow/diffops.h"

using namespace std;

namespace chflow {

// ===================================================================================================== G(x)
DeviceDSI::DeviceDSI(const FlowField& u, const DNSFlags& flags, const TimeStep& dt, const FieldSymmetry& sigma, Real T, bool Tnormalize,
                     bool xrelative, bool zrelative)
    : proto_(u.Nx(), u.Ny(), u.Nz(), u.Nd(), u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi()), flags_(flags), dt_(dt), sigma_(sigma), T_(T),
      Tnormalize_(Tnormalize), xrel_(xrelative), zrel_(zrelative), size_(field2vector_size(u)) {
    flags_.verbosity = Silent;
}

void DeviceDSI::tangent(const DeviceVector& x, int dir, DeviceVector& t) const {
    FlowField u(proto_), du;
    vector2field(x, u);
    if (dir == 0) xdiff(u, du);
    else zdiff(u, du);
    du.setPadded(true);
    if (t.size() != size_) t.resize(size_);
    field2vector(du, t);
}

void DeviceDSI::extractVector(const DeviceVector& x, FlowField& u) const {
    if (!u.geomCongruent(proto_) || u.Nd() != proto_.Nd()) u = proto_;
    vector2field(x, u);
}

// f^T(u): a fresh DNS per evaluation, as cfdsi.cpp:705-775 (the time stepper's tau factors are rebuilt by two setup kernels)
void DeviceDSI::f(const FlowField& u, FlowField& fu) {
    ++fcount_;
    if (T_ < 0) cferror("DeviceDSI::f: negative integration time");
    if (T_ == 0) {
        fu = u;
        return;
    }
    DNSFlags flags(flags_);
    TimeStep dt(dt_);
    dt.adjust_for_T(T_, false);
    flags.dt = dt;
    vector<FlowField> fields = {u, FlowField(u.Nx(), u.Ny(), u.Nz(), 1, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi())};
    DNS dns(fields, flags);
    if (dt.variable()) {  // one trial step fixes dt from the CFL number, then restart from u
        dns.advance(fields, 1);
        dt.adjust(dns.CFL(fields[0]), false);
        dns.reset_dt(dt);
        fields[0] = u;
    }
    steps_ = 0;
    if (!dt.variable()) {
        // fixed dt: the whole integration is one advance() call, which replays the step sequence as a CUDA graph
        CFL_ = dns.CFL(fields[0]);
        dns.advance(fields, dt.N() * dt.n());
        steps_ = (Real)dt.N() * dt.n();
    } else {
        for (int s = 1; s <= dt.N(); ++s) {
            CFL_ = dns.CFL(fields[0]);
            dns.advance(fields, dt.n());
            steps_ += dt.n();
            if (dt.adjust(CFL_, false)) dns.reset_dt(dt);
        }
    }
    if (fu.congruent(fields[0])) swap(fu, fields[0]);
    else fu = fields[0];
    const Real nrm = L2Norm(fu);
    if (!std::isfinite(nrm)) cferror("DeviceDSI::f: f^T(u) is not finite");
}

void DeviceDSI::G(const FlowField& u, FlowField& Gu) {
    f(u, Gu);
    Gu *= sigma_;
    Gu -= u;
    if (Tnormalize_) Gu *= 1.0 / T_;
}

void DeviceDSI::eval(const DeviceVector& x, DeviceVector& Gx) {
    FlowField u(proto_), Gu;
    vector2field(x, u);
    G(u, Gu);
    if (Gx.size() != size_) Gx.resize(size_);
    field2vector(Gu, Gx);
}

Real DeviceDSI::residual(const DeviceVector& Gx) const {
    FlowField Gu(proto_);
    vector2field(Gx, Gu);
    return L2Norm(Gu);
}

// ===================================================================================================== small dense algebra
namespace {

// column-major dense matrix for the Hessenberg problem
struct Dense {
    int m = 0, n = 0;
    vector<Real> a;
    Dense() {}
    Dense(int m_, int n_) : m(m_), n(n_), a((size_t)m_ * n_, 0.0) {}
    Real& operator()(int i, int j) { return a[(size_t)j * m + i]; }
    Real operator()(int i, int j) const { return a[(size_t)j * m + i]; }
};

// One-sided Jacobi SVD  A = U diag(d) V^T  (A m x n, m >= n): rotations orthogonalise the columns of A in place
void jacobi_svd(const Dense& A, Dense& U, vector<Real>& d, Dense& V) {
    const int m = A.m, n = A.n;
    U = A;
    V = Dense(n, n);
    for (int j = 0; j < n; ++j) V(j, j) = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        Real off = 0.0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                Real app = 0, aqq = 0, apq = 0;
                for (int i = 0; i < m; ++i) {
                    app += U(i, p) * U(i, p);
                    aqq += U(i, q) * U(i, q);
                    apq += U(i, p) * U(i, q);
                }
                if (fabs(apq) <= 1e-300 || fabs(apq) <= 1e-16 * sqrt(app * aqq)) continue;
                off = max(off, fabs(apq) / sqrt(app * aqq));
                const Real zeta = (aqq - app) / (2.0 * apq);
                const Real t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const Real c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < m; ++i) {
                    const Real up = U(i, p), uq = U(i, q);
                    U(i, p) = c * up - s * uq;
                    U(i, q) = s * up + c * uq;
                }
                for (int i = 0; i < n; ++i) {
                    const Real vp = V(i, p), vq = V(i, q);
                    V(i, p) = c * vp - s * vq;
                    V(i, q) = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    d.assign(n, 0.0);
    for (int j = 0; j < n; ++j) {
        Real s = 0;
        for (int i = 0; i < m; ++i) s += U(i, j) * U(i, j);
        d[j] = sqrt(s);
        if (d[j] > 0)
            for (int i = 0; i < m; ++i) U(i, j) /= d[j];
    }
}

// The Krylov-space model of one Newton step:  min_y |H y - beta e1|  (H (n+1) x n from Arnoldi), optionally with |y| <= delta
struct KrylovModel {
    int n = 0;
    Dense U, V;
    vector<Real> d, bh;  // singular values, U^T (beta e1)
    Real beta = 0;
    void factor(const Dense& H, int n_, Real beta_) {
        n = n_;
        beta = beta_;
        Dense Hn(n + 1, n);
        for (int j = 0; j < n; ++j)
            for (int i = 0; i <= n; ++i) Hn(i, j) = H(i, j);
        jacobi_svd(Hn, U, d, V);
        bh.assign(n, 0.0);
        for (int j = 0; j < n; ++j) bh[j] = beta * U(0, j);
    }
    // y(mu) = V diag(d / (d^2 + mu)) bh: mu = 0 is the least-squares (Newton) step, mu > 0 shortens it (hookstep)
    vector<Real> step(Real mu) const {
        vector<Real> y(n, 0.0);
        const Real dmax = n ? *max_element(d.begin(), d.end()) : 0.0;
        for (int j = 0; j < n; ++j) {
            if (d[j] <= 1e-14 * dmax) continue;
            const Real c = d[j] * bh[j] / (d[j] * d[j] + mu);
            for (int i = 0; i < n; ++i) y[i] += V(i, j) * c;
        }
        return y;
    }
    static Real norm(const vector<Real>& y) {
        Real s = 0;
        for (Real v : y) s += v * v;
        return sqrt(s);
    }
    // |H y - beta e1| for y = step(mu): the part of beta e1 outside range(U) plus the damped components
    Real model_residual(Real mu) const {
        Real in_range = 0, res = 0;
        const Real dmax = n ? *max_element(d.begin(), d.end()) : 0.0;
        for (int j = 0; j < n; ++j) {
            in_range += bh[j] * bh[j];
            const Real keep = d[j] <= 1e-14 * dmax ? 1.0 : mu / (d[j] * d[j] + mu);
            res += keep * keep * bh[j] * bh[j];
        }
        return sqrt(max(0.0, beta * beta - in_range) + res);
    }
    // the step of length delta (or the Newton step if that is shorter); returns mu
    Real constrained(Real delta, vector<Real>& y) const {
        y = step(0.0);
        if (norm(y) <= delta) return 0.0;
        Real lo = 0.0, hi = 1.0;
        while (norm(step(hi)) > delta) hi *= 4.0;
        for (int it = 0; it < 200; ++it) {
            const Real mid = 0.5 * (lo + hi);
            (norm(step(mid)) > delta ? lo : hi) = mid;
            if (hi - lo <= 1e-14 * hi) break;
        }
        y = step(hi);
        return hi;
    }
};

}  // namespace

// ===================================================================================================== the search
namespace {
// state vector of the search: the packed field plus the phase-shift unknowns (ax, az; zero and inert when not searched)
struct XVec {
    DeviceVector v;
    Real e[2] = {0.0, 0.0};
    explicit XVec(long n = 0) : v(n) {}
    Real dot(const XVec& o) const { return v.dot(o.v) + e[0] * o.e[0] + e[1] * o.e[1]; }
    Real norm() const { return sqrt(dot(*this)); }
    void axpy(Real a, const XVec& x) { v.axpy(a, x.v); e[0] += a * x.e[0]; e[1] += a * x.e[1]; }
    void scale(Real a) { v.scale(a); e[0] *= a; e[1] *= a; }
    void zero() { v.setToZero(); e[0] = e[1] = 0.0; }
};
}  // namespace

DeviceSearchResult hookstepSearch(DeviceDSI& dsi, DeviceVector& x0, const DeviceSearchFlags& fl) {
    ostream& os = *fl.logstream;
    DeviceSearchResult out;
    const long N = dsi.size();
    const bool rel[2] = {dsi.xrelative(), dsi.zrelative()};
    XVec x(N), Gx(N), xt(N), Gt(N), dx(N), w(N);
    x.v = x0;
    x.e[0] = rel[0] ? dsi.sigma().ax() : 0.0;
    x.e[1] = rel[1] ? dsi.sigma().az() : 0.0;
    const Real shift0[2] = {dsi.sigma().ax(), dsi.sigma().az()};
    // G at an extended state: the shifts of sigma are taken from the state where they are unknowns; the extra rows of G are 0
    auto evalG = [&](const XVec& X, XVec& GX) {
        dsi.setShifts(rel[0] ? X.e[0] : shift0[0], rel[1] ? X.e[1] : shift0[1]);
        dsi.eval(X.v, GX.v);
        GX.e[0] = GX.e[1] = 0.0;
    };
    evalG(x, Gx);
    Real gnorm = Gx.norm();              // 2-norm of the packed residual: what GMRES and the trust region see
    Real resid = dsi.residual(Gx.v);     // L2Norm(G): the convergence measure
    out.history.push_back(resid);
    Real delta = fl.delta;
    os << setprecision(8) << "hookstepSearch: N == " << N << " + " << int(rel[0]) + int(rel[1]) << " unknowns, L2Norm(G) == " << resid << endl;

    for (int newton = 0; newton < fl.Nnewton && resid >= fl.epsSearch; ++newton) {
        // translation directions at the current state (unit vectors): the Newton step is kept orthogonal to them, which
        // closes the system for the shift unknowns (the constraint rows of nsolver/newtonalgorithm.cpp)
        DeviceVector tang[2];
        for (int d = 0; d < 2; ++d)
            if (rel[d]) {
                dsi.tangent(x.v, d, tang[d]);
                const Real tn = tang[d].norm();
                if (tn > 0) tang[d].scale(1.0 / tn);
            }
        // ---- GMRES on A dX = -G,  A dX = [ (G(X + eps dX) - G(X)) / eps ; <dx, du/dx> ; <dx, du/dz> ]
        vector<XVec> Q;
        Q.reserve(fl.Ngmres + 1);
        Q.emplace_back(Gx);
        Q[0].scale(-1.0 / gnorm);
        Dense H(fl.Ngmres + 2, fl.Ngmres + 1);
        KrylovModel km;
        const Real xnorm = max(x.norm(), 1e-300);
        int n = 0;
        Real gres = 1.0;
        while (n < fl.Ngmres) {
            const Real eps = fl.epsDx * xnorm;  // |q| = 1
            xt = x;
            xt.axpy(eps, Q[n]);
            evalG(xt, w);
            if (fl.centdiff) {
                xt = x;
                xt.axpy(-eps, Q[n]);
                evalG(xt, Gt);
                w.axpy(-1.0, Gt);
                w.scale(0.5 / eps);
            } else {
                w.axpy(-1.0, Gx);
                w.scale(1.0 / eps);
            }
            for (int d = 0; d < 2; ++d) w.e[d] = rel[d] ? Q[n].v.dot(tang[d]) : 0.0;
            for (int j = 0; j <= n; ++j) {  // modified Gram-Schmidt on the device
                const Real h = w.dot(Q[j]);
                H(j, n) = h;
                w.axpy(-h, Q[j]);
            }
            const Real hn = w.norm();
            H(n + 1, n) = hn;
            ++n;
            ++out.gmresIterations;
            km.factor(H, n, gnorm);
            gres = km.model_residual(0.0) / gnorm;
            os << "  gmres " << setw(3) << n << "  residual " << gres << endl;
            if (gres < fl.epsGMRES || hn <= fl.epsKrylov) break;
            Q.emplace_back(w);
            Q[n].scale(1.0 / hn);
        }
        if (gres > fl.epsGMRESf) os << "  GMRES stopped at residual " << gres << " > epsGMRESfinal: the step is taken from the Krylov space built so far" << endl;

        // ---- hookstep: shrink or grow the trust region until the step improves the residual as the linear model predicts
        bool accepted = false;
        Real best_g = gnorm;
        for (int hook = 0; hook < fl.Nhook; ++hook) {
            vector<Real> y;
            const Real mu = km.constrained(delta, y);
            const Real ynorm = KrylovModel::norm(y);
            dx.zero();
            for (int j = 0; j < n; ++j) dx.axpy(y[j], Q[j]);
            xt = x;
            xt.axpy(1.0, dx);
            evalG(xt, Gt);
            const Real gt = Gt.norm();
            const Real predicted = km.model_residual(mu);     // |G + A dX| of the linear model
            const Real gain_model = gnorm - predicted, gain = gnorm - gt;
            os << "  hookstep " << hook << ": delta " << delta << " |dx| " << ynorm << " mu " << mu << "  |G| " << gnorm << " -> " << gt
               << " (linear model " << predicted << ")" << endl;
            const bool improved = gt < gnorm && gain >= fl.improvReq * gain_model;
            if (!improved) {
                // reduce the radius by at least lambdaRequiredReduction, at most to lambdaMin of the current one
                Real lambda = fl.lambdaRequiredReduction;
                if (gt > gnorm && gain_model > 0) lambda = max(fl.lambdaMin, min(lambda, gain_model / (2.0 * (gt - predicted))));
                delta = min(delta, ynorm) * lambda;
                if (delta < fl.deltaMin) {
                    os << "  trust region radius below deltaMin: stopping" << endl;
                    break;
                }
                continue;
            }
            // accept; adjust the radius for the next Newton step from the quality of the model
            const Real quality = gain_model > 0 ? gain / gain_model : 1.0;
            x = xt;
            Gx = Gt;
            best_g = gt;
            accepted = true;
            if (quality > fl.improvGood && mu > 0.0) delta = min(fl.deltaMax, fl.lambdaMax * delta);
            else if (quality < fl.improvOk) delta = max(fl.deltaMin, fl.lambdaRequiredReduction * delta);
            if (mu == 0.0) delta = max(delta, min(fl.deltaMax, ynorm));  // a full Newton step fitted: no reason to stay smaller
            break;
        }
        if (!accepted) break;
        gnorm = best_g;
        resid = dsi.residual(Gx.v);
        ++out.newtonSteps;
        out.history.push_back(resid);
        os << "newton step " << out.newtonSteps << ": L2Norm(G) == " << resid << "  ax == " << setprecision(17) << (rel[0] ? x.e[0] : shift0[0])
           << " az == " << (rel[1] ? x.e[1] : shift0[1]) << setprecision(8) << "  (DNS integrations so far: " << dsi.evaluations() << ")" << endl;
    }
    dsi.setShifts(rel[0] ? x.e[0] : shift0[0], rel[1] ? x.e[1] : shift0[1]);
    x0 = x.v;
    out.residual = resid;
    out.converged = resid < fl.epsSearch;
    out.fevals = dsi.evaluations();
    return out;
}

}  // namespace chflow
