// Poincare sections on top of the device DNS: PoincareCondition family, DNSPoincare::advanceToSection, and the FlowField
// interpolants it uses.  Behaviour follows channelflow/dns.cpp:450-704 and flowfield.cpp:4136-4287 (see the note in dns.h
// on where the reference's current code does not do what its own comments and callers say).
#include <cmath>
#include <cstdlib>

#include "cfbasics/cfbasics.h"
#include "channelflow/diffops.h"
#include "channelflow/dns.h"

namespace chflow {

namespace {
// Lagrange weights of the nodes xn at x
std::vector<Real> lagrange_weights(const cfarray<Real>& xn, Real x) {
    const int N = xn.length();
    std::vector<Real> w(N, 1.0);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            if (j != i) w[i] *= (x - xn[j]) / (xn[i] - xn[j]);
    return w;
}

FlowField lagrange_field(cfarray<FlowField>& un, const cfarray<Real>& mun, Real mu, const char* who) {
    const int N = un.length();
    if (N < 1 || mun.length() != N) cferror(std::string("error in ") + who + ": un and mun must have the same, nonzero length");
    for (int n = 1; n < N; ++n)
        if (un[n].Nx() != un[0].Nx() || un[n].Ny() != un[0].Ny() || un[n].Nz() != un[0].Nz() || un[n].Nd() != un[0].Nd())
            cferror(std::string("error in ") + who + ": incompatible grids in un");
    const std::vector<Real> w = lagrange_weights(mun, mu);
    cfarray<Real> g(N);
    auto geom = [&](Real (FlowField::*get)() const) {
        for (int n = 0; n < N; ++n) g[n] = (un[n].*get)();
        return isconst(g) ? g[0] : polynomialInterpolate(g, mun, mu);
    };
    const Real Lx = geom(&FlowField::Lx), Lz = geom(&FlowField::Lz), a = geom(&FlowField::a), b = geom(&FlowField::b);
    // the data are combined in spectral state (the transforms are linear); the inputs keep their states
    bool padded = true;
    FlowField u(un[0].Nx(), un[0].Ny(), un[0].Nz(), un[0].Nd(), Lx, Lz, a, b, un[0].cfmpi(), Spectral, Spectral);
    for (int n = 0; n < N; ++n) {
        padded = padded && un[n].padded();
        if (un[n].xzstate() == Spectral && un[n].ystate() == Spectral && un[n].geomCongruent(u)) {
            u.add(w[n], un[n]);
        } else {
            FlowField t(un[n]);
            t.makeSpectral();
            if (!t.geomCongruent(u)) {   // box lengths or walls vary with mu: the same coefficients on the interpolated box
                std::vector<Real> raw(t.Nloc());
                t.raw_download(raw.data());
                t = FlowField(u.Nx(), u.Ny(), u.Nz(), u.Nd(), Lx, Lz, a, b, u.cfmpi(), Spectral, Spectral);
                t.raw_upload(raw.data());
            }
            u.add(w[n], t);
        }
    }
    if (padded) u.zeroPaddedModes();
    return u;
}
}  // namespace

FlowField quadraticInterpolate(cfarray<FlowField>& un, const cfarray<Real>& mun, Real mu, Real) {
    if (un.length() != 3 || mun.length() != 3) cferror("error in quadraticInterpolate(cfarray<FlowField>&, ...): three fields and three parameters needed");
    return lagrange_field(un, mun, mu, "quadraticInterpolate(cfarray<FlowField>&, cfarray<Real>&, Real, Real)");
}
FlowField polynomialInterpolate(cfarray<FlowField>& un, cfarray<Real>& mun, Real mu) {
    return lagrange_field(un, mun, mu, "polynomialInterpolate(cfarray<FlowField>&, cfarray<Real>&, Real)");
}

// ------------------------------------------------------------------------------------------------- conditions
PlaneIntersection::PlaneIntersection(const FlowField& ustar, const FlowField& estar) : estar_(estar), cstar_(L2IP(ustar, estar)) {}
Real PlaneIntersection::operator()(const FlowField& u) { return L2IP(u, estar_) - cstar_; }
Real DragDissipation::operator()(const FlowField& u) { return wallshear(u) - dissipation(u); }

// ------------------------------------------------------------------------------------------------- DNSPoincare
void DNS::operator*=(const std::vector<FieldSymmetry>& sigma) {
    if (init_algorithm_) *init_algorithm_ *= sigma;
    if (main_algorithm_) *main_algorithm_ *= sigma;
}

namespace {
// (u, q) with a zero pressure field on u's grid: the algorithms keep both (the reference constructs from {u} alone)
std::vector<FlowField> with_pressure(const FlowField& u) {
    return {u, FlowField(u.Nx(), u.Ny(), u.Nz(), 1, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi(), Spectral, Spectral)};
}
}  // namespace

DNSPoincare::DNSPoincare() : DNS() {}
DNSPoincare::DNSPoincare(FlowField& u, PoincareCondition* h, const DNSFlags& flags)
    : DNS(with_pressure(u), flags), h_(h), hcurrent_((*h)(u)), t0_(flags.t0) {}
DNSPoincare::DNSPoincare(FlowField& u, const cfarray<FlowField>& e, const cfarray<FieldSymmetry>& sigma, PoincareCondition* h,
                         const DNSFlags& flags)
    : DNS(with_pressure(u), flags), e_(e), sigma_(sigma), h_(h), hcurrent_((*h)(u)), t0_(flags.t0) {}

bool DNSPoincare::advanceToSection(FlowField& u, FlowField& q, int nSteps, int crosssign, Real Tmin, Real epsilon) {
    std::ostream& os = *flags().logstream;
    FlowField ustart(u), qstart(q);

    {   // the coarse stride
        std::vector<FlowField> state = {u, q};
        advance(state, nSteps);
        u = state[0];
        q = state[1];
    }

    // map back into the fundamental domain, and the start of the stride and the stepper's history with it
    if (e_.length() > 0) {
        FieldSymmetry s, identity;
        for (int n = 0; n < e_.length(); ++n)
            if (L2IP(u, e_[n]) < 0) s *= sigma_[n];
        if (s != identity) {
            u *= s;
            q *= s;
            ustart *= s;
            qstart *= s;
            *this *= std::vector<FieldSymmetry>{s, FieldSymmetry()};
        }
    }

    const Real tend = DNS::time();
    if (!(tend - t0_ > Tmin)) return false;

    const Real h0 = (*h_)(ustart), h1 = (*h_)(u);
    hcurrent_ = h1;
    const bool up = h0 < 0 && 0 <= h1, down = h0 > 0 && 0 >= h1;
    if (!((crosssign > 0 && up) || (crosssign < 0 && down) || (crosssign == 0 && (up || down)))) return false;
    os << (up ? '+' : '-') << std::flush;
    scrossing_ = up ? 1 : -1;

    // second pass over the stride, one step at a time, keeping the last three states: index 0 is the newest
    const Real dt = DNS::dt(), tstart = tend - dt * nSteps;
    cfarray<Real> ts(3), hs(3);
    cfarray<FlowField> us(3), qs(3);
    ts[0] = tstart; hs[0] = h0; us[0] = ustart; qs[0] = qstart;
    ts[1] = ts[2] = 0.0; hs[1] = hs[2] = 0.0;

    DNSFlags fine = DNS::flags();
    fine.verbosity = Silent;
    fine.t0 = tstart;
    DNS dns({us[0], qs[0]}, fine);

    int have = 1;
    for (Real t = tstart; t <= tend + dt; t += dt) {
        for (int n = 2; n > 0; --n) {
            ts[n] = ts[n - 1]; hs[n] = hs[n - 1];
            us[n] = us[n - 1]; qs[n] = qs[n - 1];
        }
        std::vector<FlowField> state = {us[0], qs[0]};
        dns.advance(state, 1);
        us[0] = state[0];
        qs[0] = state[1];
        hs[0] = (*h_)(us[0]);
        ts[0] = t + dt;
        os << ':' << std::flush;

        if (++have < 3 || !((hs[2] < 0 && 0 <= hs[0]) || (hs[2] > 0 && 0 >= hs[0]))) continue;

        // Newton iteration on g(s) = h(v(s)), v the quadratic interpolant of the three states; start from the inverse
        // interpolation s(h) at h = 0, derivative by a forward difference of relative size 1e-9 (dns.cpp:641-684)
        Real s = polynomialInterpolate(ts, hs, 0.0);
        const Real rel = 1e-9;
        const int maxit = 6;
        Real g = 0;
        for (int it = 0; it < maxit; ++it) {
            FlowField v = polynomialInterpolate(us, ts, s);
            g = (*h_)(v);
            const bool good = std::abs(g) < epsilon / 2;
            if (good || it == maxit - 1) {
                os << (good ? "|" : "~|") << std::flush;
                tcrossing_ = s;
                ucrossing_ = v;
                pcrossing_ = polynomialInterpolate(qs, ts, s);
                break;
            }
            FlowField vd = polynomialInterpolate(us, ts, s + rel * s);
            const Real dgds = ((*h_)(vd) - g) / (rel * s);
            s -= g / dgds;
        }
        hcrossing_ = g;
        return true;
    }
    os << "DNSPoincare::advanceToSection: the stride crossed the section but the step-by-step pass over it did not. Exiting." << std::endl;
    std::exit(1);
}

}  // namespace chflow
