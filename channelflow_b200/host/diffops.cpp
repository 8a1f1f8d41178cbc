// FlowField norms, differential operators and pointwise products over the device layer (see channelflow/diffops.h).
// What each function computes follows the reference's diffops.cpp (line ranges beside the functions); how it is computed
// does not: one generic spectral-operator kernel, one pointwise kernel, batched reductions.
#include "channelflow/diffops.h"

#include <sstream>

namespace chflow {

namespace {
struct Term { int out, in, nx, ny, nz; Real coef; };
// out = sum of terms applied to f (f is made spectral for the call and restored)
void apply(const FlowField& f_, FlowField& out, int Ndout, const std::vector<Term>& terms) {
    FlowField& f = const_cast<FlowField&>(f_);
    const fieldstate sxz = f.xzstate(), sy = f.ystate();
    f.makeSpectral();
    if (!f.geomCongruent(out) || out.Nd() != Ndout) out.resize(f.Nx(), f.Ny(), f.Nz(), Ndout, f.Lx(), f.Lz(), f.a(), f.b(), f.cfmpi());
    std::vector<int> o, i, nx, ny, nz;
    std::vector<Real> c;
    for (const Term& t : terms) { o.push_back(t.out); i.push_back(t.in); nx.push_back(t.nx); ny.push_back(t.ny); nz.push_back(t.nz); c.push_back(t.coef); }
    cfgpu_check(cfgpu_field_diffop(out.device_overwrite(), f.device(), (int)terms.size(), o.data(), i.data(), nx.data(), ny.data(), nz.data(), c.data()),
                "cfgpu_field_diffop");
    out.setState(Spectral, Spectral);
    out.setPadded(f.padded());
    f.makeState(sxz, sy);
}
// out = op(f, g) pointwise in the physical state (inputs are restored to their states)
void pointwise(int op, const FlowField& f_, const FlowField* g_, FlowField& out, int Ndout) {
    FlowField& f = const_cast<FlowField&>(f_);
    FlowField* g = const_cast<FlowField*>(g_);
    const fieldstate fxz = f.xzstate(), fy = f.ystate();
    const fieldstate gxz = g ? g->xzstate() : Physical, gy = g ? g->ystate() : Physical;
    f.makePhysical();
    if (g) g->makePhysical();
    if (!f.geomCongruent(out) || out.Nd() != Ndout) out.resize(f.Nx(), f.Ny(), f.Nz(), Ndout, f.Lx(), f.Lz(), f.a(), f.b(), f.cfmpi());
    out.setState(Physical, Physical);
    out.setPadded(false);
    cfgpu_check(cfgpu_field_pointwise(op, out.device_overwrite(), f.device(), g ? g->device() : nullptr), "cfgpu_field_pointwise");
    f.makeState(fxz, fy);
    if (g && g != &f) g->makeState(gxz, gy);
}
}  // namespace

// ------------------------------------------------------------------------------------------------ norms (diffops.cpp:18-800)
Real L2Norm2(const FlowField& u, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_l2norm2(u.device(), normalize ? 1 : 0, &r), "cfgpu_l2norm2");
    return r;
}
Real L2Norm(const FlowField& u, bool normalize) { return sqrt(L2Norm2(u, normalize)); }
Real L2Norm2_3d(const FlowField& u, bool normalize) {
    double r = 0;
    cfgpu_check(cfgpu_l2norm2_3d(u.device(), normalize ? 1 : 0, &r), "cfgpu_l2norm2_3d");
    return r;
}
Real L2Norm3d(const FlowField& u, bool normalize) { return sqrt(L2Norm2_3d(u, normalize)); }
Real L2Dist2(const FlowField& u, const FlowField& v, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_l2dist2(u.device(), v.device(), normalize ? 1 : 0, &r), "cfgpu_l2dist2");
    return r;
}
Real L2Dist(const FlowField& u, const FlowField& v, bool normalize) { return sqrt(L2Dist2(u, v, normalize)); }
Real L2InnerProduct(const FlowField& u, const FlowField& v, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_l2ip(u.device(), v.device(), normalize ? 1 : 0, &r), "cfgpu_l2ip");
    return r;
}
// Chebyshev-weighted norms (diffops.cpp:259-350): the y inner product with weight 1/sqrt(1-y^2), diagonal in the coefficients
Real chebyNorm2(const FlowField& u, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_chebyform(u.device(), nullptr, 0, normalize ? 1 : 0, &r), "cfgpu_chebyform");
    return r;
}
Real chebyNorm(const FlowField& u, bool normalize) { return sqrt(chebyNorm2(u, normalize)); }
Real chebyDist2(const FlowField& u, const FlowField& v, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_chebyform(u.device(), v.device(), 1, normalize ? 1 : 0, &r), "cfgpu_chebyform");
    return r;
}
Real chebyDist(const FlowField& u, const FlowField& v, bool normalize) { return sqrt(chebyDist2(u, v, normalize)); }
static void default_box(const FlowField& f, int& kxmax, int& kzmax, bool zero_means_default) {
    if (kxmax < 0 || kxmax > f.kxmax() || (zero_means_default && kxmax == 0)) kxmax = f.padded() ? f.kxmaxDealiased() : f.kxmax();
    if (kzmax < 0 || kzmax > f.kzmax() || (zero_means_default && kzmax == 0)) kzmax = f.padded() ? f.kzmaxDealiased() : f.kzmax();
}
Real L2Norm2(const FlowField& f, int kxmax, int kzmax, bool normalize) {
    default_box(f, kxmax, kzmax, false);
    Real r = 0;
    cfgpu_check(cfgpu_l2form_box(f.device(), nullptr, 0, kxmax, kzmax, 1, normalize ? 1 : 0, &r), "cfgpu_l2form_box");
    return r;
}
Real L2Norm(const FlowField& f, int kxmax, int kzmax, bool normalize) { return sqrt(L2Norm2(f, kxmax, kzmax, normalize)); }
Real L2Dist2(const FlowField& f, const FlowField& g, int kxmax, int kzmax, bool normalize) {
    default_box(f, kxmax, kzmax, true);  // (the reference treats 0 as "default" here, diffops.cpp:655-661)
    Real r = 0;
    cfgpu_check(cfgpu_l2form_box(f.device(), g.device(), 1, kxmax, kzmax, 1, normalize ? 1 : 0, &r), "cfgpu_l2form_box");
    return r;
}
Real L2Dist(const FlowField& f, const FlowField& g, int kxmax, int kzmax, bool normalize) { return sqrt(L2Dist2(f, g, kxmax, kzmax, normalize)); }
Real L2InnerProduct(const FlowField& f, const FlowField& g, int kxmax, int kzmax, bool normalize) {
    if (kxmax < 0 || kxmax > f.kxmax() || kxmax > g.kxmax())
        kxmax = lesser(f.padded() ? f.kxmaxDealiased() : f.kxmax(), g.padded() ? g.kxmaxDealiased() : g.kxmax());
    if (kzmax < 0 || kzmax > f.kzmax() || kzmax > g.kzmax())
        kzmax = lesser(f.padded() ? f.kzmaxDealiased() : f.kzmax(), g.padded() ? g.kzmaxDealiased() : g.kzmax());
    Real r = 0;
    cfgpu_check(cfgpu_l2form_box(f.device(), g.device(), 2, kxmax, kzmax, 1, normalize ? 1 : 0, &r), "cfgpu_l2form_box");
    return r;
}
Real bcNorm2(const FlowField& f, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_bcnorm2(f.device(), nullptr, normalize ? 1 : 0, &r), "cfgpu_bcnorm2");
    return r;
}
Real bcDist2(const FlowField& f, const FlowField& g, bool normalize) {
    Real r = 0;
    cfgpu_check(cfgpu_bcnorm2(f.device(), g.device(), normalize ? 1 : 0, &r), "cfgpu_bcnorm2");
    return r;
}
Real bcNorm(const FlowField& f, bool normalize) { return sqrt(bcNorm2(f, normalize)); }
Real bcDist(const FlowField& f, const FlowField& g, bool normalize) { return sqrt(bcDist2(f, g, normalize)); }
// divergence with plain i k factors on every stored mode, summed without the kz ghost weights (diffops.cpp:91-116 through
// basisfunc.cpp:520-531)
static Real divnorm2_of(const FlowField& f, bool normalize) {
    FlowField d;
    // (no Nyquist rule here: basisfunc.cpp:524-526 multiplies by 2 pi k/L for every mode; the generic operator drops the
    // odd-order Nyquist term, which only matters for un-dealiased fields with energy in the Nyquist modes)
    apply(f, d, 1, {{0, 0, 1, 0, 0, 1.0}, {0, 1, 0, 1, 0, 1.0}, {0, 2, 0, 0, 1, 1.0}});
    Real r = 0;
    cfgpu_check(cfgpu_l2form_box(d.device(), nullptr, 0, f.Nx(), f.Nz(), 0, normalize ? 1 : 0, &r), "cfgpu_l2form_box");
    if (!normalize) r /= f.Lx() * f.Lz();  // basisfunc L2Norm2(.., false) multiplies by Ly only
    return r;
}
Real divNorm2(const FlowField& f, bool normalize) {
    assert(f.Nd() == 3);
    return divnorm2_of(f, normalize);
}
Real divDist2(const FlowField& f, const FlowField& g, bool normalize) {
    FlowField d(f);
    d -= g;
    return divnorm2_of(d, normalize);
}
Real divNorm(const FlowField& f, bool normalize) { return sqrt(divNorm2(f, normalize)); }
Real divDist(const FlowField& f, const FlowField& g, bool normalize) { return sqrt(divDist2(f, g, normalize)); }

Real dissipation(const FlowField& u, bool normalize) {  // diffops.cpp:767-777: sum_i || grad u_i ||^2
    FlowField g;
    grad(u, g);
    return L2Norm2(g, normalize);
}
Real wallshearLower(const FlowField& f, bool normalize) {
    Real I = 0.5 * sqrt(square(f.dudy_a()) + square(f.dwdy_a()));
    if (!normalize) I *= 2 * f.Lx() * f.Lz();
    return I;
}
Real wallshearUpper(const FlowField& f, bool normalize) {
    Real I = 0.5 * sqrt(square(f.dudy_b()) + square(f.dwdy_b()));
    if (!normalize) I *= 2 * f.Lx() * f.Lz();
    return I;
}
Real wallshear(const FlowField& f, bool normalize) { return wallshearLower(f, normalize) + wallshearUpper(f, normalize); }
Real L2Norm_uvw(const FlowField& u, const bool ux, const bool uy, const bool uz) {  // diffops.cpp:3885-3929
    const bool use[3] = {ux, uy, uz};
    Real sum = 0.0;
    for (int i = 0; i < 3 && i < u.Nd(); ++i)
        if (use[i]) sum += L2Norm2(u[i], true) * (u.b() - u.a());
    return sqrt(sum);
}
Real Ecf(const FlowField& u) { return pow(L2Norm_uvw(u, false, true, true), 2); }

// ------------------------------------------------------------------------------------------------ differential operators
void diff(const FlowField& f, FlowField& df, int nx, int ny, int nz) {  // diffops.cpp:1650-1782
    if (ny > 2) {  // higher y orders: repeated second / first derivatives
        FlowField t;
        diff(f, t, nx, 2, nz);
        diff(t, df, 0, ny - 2, 0);
        return;
    }
    std::vector<Term> t;
    for (int i = 0; i < f.Nd(); ++i) t.push_back({i, i, nx, ny, nz, 1.0});
    if (f.Nd() * 1 > 27) cferror("diff: too many components");
    apply(f, df, f.Nd(), t);
}
void xdiff(const FlowField& f, FlowField& d, int n) { diff(f, d, n, 0, 0); }
void ydiff(const FlowField& f, FlowField& d, int n) { diff(f, d, 0, n, 0); }
void zdiff(const FlowField& f, FlowField& d, int n) { diff(f, d, 0, 0, n); }
void diff(const FlowField& f, FlowField& df, int i, int n) { diff(f, df, i == 0 ? n : 0, i == 1 ? n : 0, i == 2 ? n : 0); }
void grad(const FlowField& f, FlowField& g) {  // diffops.cpp:1784-1941
    if (f.Nd() != 1 && f.Nd() != 3) cferror("grad(f, gradf): f must be 1d or 3d");
    std::vector<Term> t;
    for (int i = 0; i < f.Nd(); ++i)
        for (int j = 0; j < 3; ++j) t.push_back({f.Nd() == 1 ? j : 3 * i + j, i, j == 0, j == 1, j == 2, 1.0});
    apply(f, g, 3 * f.Nd(), t);
}
void lapl(const FlowField& f, FlowField& l) {  // diffops.cpp:2042-2130
    std::vector<Term> t;
    for (int i = 0; i < f.Nd(); ++i) {
        t.push_back({i, i, 2, 0, 0, 1.0});
        t.push_back({i, i, 0, 2, 0, 1.0});
        t.push_back({i, i, 0, 0, 2, 1.0});
    }
    apply(f, l, f.Nd(), t);
}
void curl(const FlowField& f, FlowField& c) {  // diffops.cpp:2229-2334
    assert(f.Nd() == 3);
    apply(f, c, 3, {{0, 2, 0, 1, 0, 1.0}, {0, 1, 0, 0, 1, -1.0},    // w_y - v_z
                    {1, 0, 0, 0, 1, 1.0}, {1, 2, 1, 0, 0, -1.0},    // u_z - w_x
                    {2, 1, 1, 0, 0, 1.0}, {2, 0, 0, 1, 0, -1.0}});  // v_x - u_y
}
void div(const FlowField& f, FlowField& d, const fieldstate finalstate) {  // diffops.cpp:2470-2558
    if (f.Nd() == 3) apply(f, d, 1, {{0, 0, 1, 0, 0, 1.0}, {0, 1, 0, 1, 0, 1.0}, {0, 2, 0, 0, 1, 1.0}});
    else if (f.Nd() == 9) {
        std::vector<Term> t;
        for (int i = 0; i < 3; ++i)  // d/dx_j f_ij with f_ij at component 3i+j
            for (int j = 0; j < 3; ++j) t.push_back({i, 3 * i + j, j == 0, j == 1, j == 2, 1.0});
        apply(f, d, 3, t);
    } else cferror("div(f, divf): f must be 3d or 9d");
    if (finalstate == Physical) d.makePhysical();
}
void norm2(const FlowField& f, FlowField& n2) { pointwise(3, f, nullptr, n2, 1); n2.makeSpectral(); }  // diffops.cpp:2132-2170
void norm(const FlowField& f, FlowField& n) { pointwise(4, f, nullptr, n, 1); n.makeSpectral(); }
void cross(const FlowField& f, const FlowField& g, FlowField& fxg, const fieldstate finalstate) {  // diffops.cpp:2560-2611
    pointwise(0, f, &g, fxg, 3);
    if (finalstate == Spectral) fxg.makeSpectral();
}
void outer(const FlowField& f, const FlowField& g, FlowField& fg) { pointwise(1, f, &g, fg, f.Nd() * g.Nd()); fg.makeSpectral(); }  // :2336-2388
void dot(const FlowField& f, const FlowField& g, FlowField& fdg) { pointwise(2, f, &g, fdg, 1); fdg.makeSpectral(); }             // :2390-2468
void energy(const FlowField& u, FlowField& e) { pointwise(5, u, nullptr, e, 1); e.makeSpectral(); }                               // :2613-2643
void energy(const FlowField& u, const ChebyCoeff& U, FlowField& e) {  // energy of u + U e_x (:2645-2700)
    FlowField t(u);
    t.makeSpectral();
    ChebyCoeff Us(U);
    Us.makeSpectral();
    t += Us;
    energy(t, e);
}
// u . grad v (diffops.cpp:3586-3643); tmp receives grad v (physical on return, as in the reference)
void dotgrad(const FlowField& u, const FlowField& v, FlowField& udgv, FlowField& tmp) {
    grad(v, tmp);
    FlowField& uu = const_cast<FlowField&>(u);
    const fieldstate sxz = uu.xzstate(), sy = uu.ystate();
    uu.makePhysical();
    tmp.makePhysical();
    const int vd = v.Nd();
    if (!u.geomCongruent(udgv) || udgv.Nd() != vd) udgv.resize(u.Nx(), u.Ny(), u.Nz(), vd, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi());
    udgv.setState(Physical, Physical);
    udgv.setToZero();
    // sum_j u_j d_j v_i: one pointwise product per (i): gather the three gradient components of v_i and dot with u
    for (int i = 0; i < vd; ++i) {
        FlowField gi(u.Nx(), u.Ny(), u.Nz(), 3, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi(), Physical, Physical);
        for (int j = 0; j < 3; ++j) gi.setComponent(j, tmp, vd == 1 ? j : 3 * i + j);
        FlowField d;
        pointwise(2, uu, &gi, d, 1);
        udgv.setComponent(i, d, 0);
    }
    udgv.makeSpectral();
    uu.makeState(sxz, sy);
}
FlowField dotgrad(const FlowField& u, const FlowField& v, FlowField& tmp) { FlowField r; dotgrad(u, v, r, tmp); return r; }

#define RETURNING(name, ...) { FlowField r; name(__VA_ARGS__, r); return r; }
FlowField xdiff(const FlowField& f, int n) { FlowField r; xdiff(f, r, n); return r; }
FlowField ydiff(const FlowField& f, int n) { FlowField r; ydiff(f, r, n); return r; }
FlowField zdiff(const FlowField& f, int n) { FlowField r; zdiff(f, r, n); return r; }
FlowField diff(const FlowField& f, int i, int n) { FlowField r; diff(f, r, i, n); return r; }
FlowField diff(const FlowField& f, int nx, int ny, int nz) { FlowField r; diff(f, r, nx, ny, nz); return r; }
FlowField grad(const FlowField& f) RETURNING(grad, f)
FlowField lapl(const FlowField& f) RETURNING(lapl, f)
FlowField curl(const FlowField& f) RETURNING(curl, f)
FlowField norm(const FlowField& f) RETURNING(norm, f)
FlowField norm2(const FlowField& f) RETURNING(norm2, f)
FlowField div(const FlowField& f) { FlowField r; div(f, r); return r; }
FlowField cross(const FlowField& f, const FlowField& g) { FlowField r; cross(f, g, r); return r; }
FlowField outer(const FlowField& f, const FlowField& g) RETURNING(outer, f, g)
FlowField dot(const FlowField& f, const FlowField& g) RETURNING(dot, f, g)
FlowField energy(const FlowField& u) RETURNING(energy, u)
FlowField energy(const FlowField& u, ChebyCoeff& U) RETURNING(energy, u, U)
#undef RETURNING

// ------------------------------------------------------------------------------------------------ mean-flow diagnostics
Real getdPdx(const FlowField& u, Real nu) { return nu * (u.dudy_b() - u.dudy_a()) / (u.b() - u.a()); }
Real getdPdz(const FlowField& u, Real nu) { return nu * (u.dwdy_b() - u.dwdy_a()) / (u.b() - u.a()); }
Real getUbulk(const FlowField& u) {
    Real ubulk = u.profile(0, 0, 0).re.mean();
    if (std::abs(ubulk) < 1e-15) ubulk = 0.0;
    return ubulk;
}
Real getWbulk(const FlowField& u) {
    Real wbulk = u.profile(0, 0, 2).re.mean();
    if (std::abs(wbulk) < 1e-15) wbulk = 0.0;
    return wbulk;
}

// one line of run-time diagnostics (diffops.cpp:3933-3968): same columns, same widths
std::string fieldstatsheader() {
    std::stringstream h;
    for (const char* c : {"L2", "u2", "v2", "w2", "e3d", "ecf", "ubulk", "wbulk", "wallshear", "wallshear_a", "wallshear_b", "dissipation"})
        h << std::setw(14) << c;
    return h.str();
}
std::string fieldstatsheader_t(const std::string tname) {
    std::stringstream h;
    h << std::setw(8) << "#(" << tname << ")" << fieldstatsheader();
    return h.str();
}
std::string fieldstats(const FlowField& u) {
    std::stringstream s;
    const Real l2n = L2Norm(u);
    if (std::isnan(l2n)) cferror("L2Norm(u) is nan");
    s << std::setw(14) << l2n << std::setw(14) << L2Norm_uvw(u, true, false, false) << std::setw(14) << L2Norm_uvw(u, false, true, false)
      << std::setw(14) << L2Norm_uvw(u, false, false, true) << std::setw(14) << L2Norm3d(u) << std::setw(14) << Ecf(u) << std::setw(14)
      << getUbulk(u) << std::setw(14) << getWbulk(u) << std::setw(14) << wallshear(u) << std::setw(14) << wallshearLower(u)
      << std::setw(14) << -1 * wallshearUpper(u) << std::setw(14) << dissipation(u);
    return s.str();
}
std::string fieldstats_t(const FlowField& u, Real t) {
    std::stringstream s;
    s << std::setw(8) << t << fieldstats(u);
    return s.str();
}

}  // namespace chflow
