// PoissonSolver / PressureSolver over the device kernels (reference channelflow/poissonsolver.cpp).
#include "channelflow/poissonsolver.h"

#include <iostream>

#include "cfgpu.h"
#include "channelflow/nse.h"

using namespace std;

namespace chflow {

namespace {
void dev(int rc, const char* what) {
    if (rc != 0) cferror(string(what) + ": " + cfgpu_last_error());
}
}  // namespace

PoissonSolver::PoissonSolver(const FlowField& u)
    : Mx_(u.Mx()), My_(u.My()), Mz_(u.Mz()), Nz_(u.Nz()), Nd_(u.Nd()), Lx_(u.Lx()), Lz_(u.Lz()), a_(u.a()), b_(u.b()) {}

PoissonSolver::PoissonSolver(int Nx, int Ny, int Nz, int Nd, Real Lx, Real Lz, Real a, Real b, CfMPI*)
    : Mx_(Nx), My_(Ny), Mz_(Nz / 2 + 1), Nz_(Nz), Nd_(Nd), Lx_(Lx), Lz_(Lz), a_(a), b_(b) {}

bool PoissonSolver::geomCongruent(const FlowField& u) const {
    return u.Mx() == Mx_ && u.My() == My_ && u.Mz() == Mz_ && u.Lx() == Lx_ && u.Lz() == Lz_ && u.a() == a_ && u.b() == b_;
}
bool PoissonSolver::congruent(const FlowField& u) const { return geomCongruent(u) && u.Nd() == Nd_; }

void PoissonSolver::prepare(FlowField& u, const FlowField& f) const {
    assert(congruent(f));
    f.assertState(Spectral, Spectral);
    if (!congruent(u)) u = FlowField(f.Nx(), f.Ny(), f.Nz(), f.Nd(), f.Lx(), f.Lz(), f.a(), f.b(), f.cfmpi());
}

void PoissonSolver::solve(FlowField& u, const FlowField& f) const {
    prepare(u, f);
    dev(cfgpu_poisson_solve(u.device_overwrite(), f.device(), nullptr), "PoissonSolver::solve");
    u.setState(Spectral, Spectral);
}

void PoissonSolver::solve(FlowField& u, const FlowField& f, const FlowField& bc) const {
    assert(congruent(bc) && bc.xzstate() == Spectral);
    prepare(u, f);
    if (bc.ystate() == Spectral)
        dev(cfgpu_poisson_solve(u.device_overwrite(), f.device(), bc.device()), "PoissonSolver::solve");
    else {
        FlowField bcs(bc);
        bcs.makeSpectral_y();
        dev(cfgpu_poisson_solve(u.device_overwrite(), f.device(), bcs.device()), "PoissonSolver::solve");
    }
    u.setState(Spectral, Spectral);
}

// || lapl u - f || + wall mismatch
Real PoissonSolver::verify(const FlowField& u, const FlowField& f) const {
    assert(congruent(u) && congruent(f));
    FlowField lapl_u;
    lapl(u, lapl_u);
    const Real l2err = L2Dist(lapl_u, f), bcerr = bcNorm(u);
    cout << "PoissonSolver::verify(u, f) {\n  L2Norm(u)         == " << L2Norm(u) << "\n  L2Norm(f)         == " << L2Norm(f)
         << "\n  L2Norm(lapl u)    == " << L2Norm(lapl_u) << "\n  L2Dist(lapl u, f) == " << l2err << "\n  bcNorm(u)         == " << bcerr
         << "\n} // PoissonSolver::verify(u, f)\n";
    return l2err + bcerr;
}

Real PoissonSolver::verify(const FlowField& u, const FlowField& f, const FlowField& bc) const {
    assert(congruent(u) && congruent(f));
    FlowField lapl_u;
    lapl(u, lapl_u);
    const Real l2err = L2Dist(lapl_u, f), bcerr = bcDist(u, bc);
    cout << "PoissonSolver::verify(u, f) {\n  L2Norm(u)         == " << L2Norm(u) << "\n  L2Norm(f)         == " << L2Norm(f)
         << "\n  L2Norm(lapl u)    == " << L2Norm(lapl_u) << "\n  L2Dist(lapl u, f) == " << l2err << "\n  bcNorm(u)         == " << bcNorm(u)
         << "\n  bcNorm(bc)        == " << bcNorm(bc) << "\n  bcDist(u,bc)      == " << bcerr << "\n} // PoissonSolver::verify(u, f)\n";
    return l2err + bcerr;
}

// ------------------------------------------------------------------------------------------------------- PressureSolver
PressureSolver::PressureSolver(int Nx, int Ny, int Nz, Real Lx, Real Lz, Real a, Real b, const ChebyCoeff& U, const ChebyCoeff& W, Real nu,
                               Real Vsuck, NonlinearMethod nonl, CfMPI* cfmpi)
    : PoissonSolver(Nx, Ny, Nz, 1, Lx, Lz, a, b, cfmpi), U_(U), W_(W), nu_(nu), Vsuck_(Vsuck), nonl_method_(nonl) {
    assert(U_.N() == Ny && W_.N() == Ny);
    U_.makeSpectral();
    W_.makeSpectral();
}

PressureSolver::PressureSolver(const FlowField& u, Real nu, Real Vsuck, NonlinearMethod nl)
    : PoissonSolver(u.Nx(), u.Ny(), u.Nz(), 1, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi()),
      U_(u.Ny(), u.a(), u.b(), Spectral), W_(u.Ny(), u.a(), u.b(), Spectral), nu_(nu), Vsuck_(Vsuck), nonl_method_(nl) {}

PressureSolver::PressureSolver(const FlowField& u, const ChebyCoeff& U, const ChebyCoeff& W, Real nu, Real Vsuck, NonlinearMethod nl)
    : PoissonSolver(u.Nx(), u.Ny(), u.Nz(), 1, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi()), U_(U), W_(W), nu_(nu), Vsuck_(Vsuck), nonl_method_(nl) {
    assert(U_.N() == u.Ny() && W_.N() == u.Ny());
    U_.makeSpectral();
    W_.makeSpectral();
}

void PressureSolver::minus_div_nonlinear(const FlowField& u) {
    DNSFlags flags;
    flags.nonlinearity = nonl_method_;
    flags.Vsuck = Vsuck_;
    flags.nu = nu_;
    navierstokesNL(u, U_, W_, nonl_, tmp_, flags);
    div(nonl_, div_nonl_);
    div_nonl_ *= -1.0;
}

FlowField PressureSolver::solve(const FlowField& u) {
    FlowField p;
    solve(p, u);
    return p;
}

// I. Dirichlet solution of lapl p = -div N(u);  II. add the homogeneous solution that sets dp/dy = nu v_yy at the walls
// (every mode but the mean one), built on the device at the Gauss-Lobatto points and transformed by the y-GEMM
void PressureSolver::solve(FlowField& p, FlowField u) {
    assert(u.xzstate() == Spectral && u.ystate() == Spectral && geomCongruent(u));
    minus_div_nonlinear(u);
    PoissonSolver::solve(p, div_nonl_);
    FlowField g(u.Nx(), u.Ny(), u.Nz(), 1, u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi());
    dev(cfgpu_pressure_neumann(g.device_overwrite(), p.device(), u.device(), nu_), "PressureSolver::solve");
    g.setState(Spectral, Physical);
    g.makeSpectral_y();
    p += g;
}

Real PressureSolver::verify(const FlowField& p, const FlowField& u) {
    assert(u.xzstate() == Spectral && u.ystate() == Spectral && congruent(p) && geomCongruent(u));
    {   // the reference's check runs with default flags apart from the nonlinearity (poissonsolver.cpp:440-446)
        DNSFlags flags;
        flags.nonlinearity = nonl_method_;
        navierstokesNL(u, U_, W_, nonl_, tmp_, flags);
        div(nonl_, div_nonl_);
        div_nonl_ *= -1.0;
    }
    PoissonSolver::verify(p, div_nonl_);
    FlowField lapl_p;
    lapl(p, lapl_p);
    const Real l2err = L2Dist(lapl_p, div_nonl_);
    FlowField dpdy, nu_vyy;
    ydiff(p, dpdy);
    ydiff(u[1], nu_vyy, 2);
    nu_vyy *= nu_;
    const Real bcerr = bcDist(dpdy, nu_vyy);
    cout << "PressureSolver::verify(p,u) {\n  L2Norm(u)           == " << L2Norm(u) << "\n  L2Norm(div(nonl(u)) == " << L2Norm(div_nonl_)
         << "\n  L2Norm(lapl p)      == " << L2Norm(lapl_p) << "\n  L2Dist(lapl p, div(nonl(u))) == " << l2err
         << "\n  bcNorm(dpdy)         == " << bcNorm(dpdy) << "\n  bcNorm(nu_vyy)       == " << bcNorm(nu_vyy)
         << "\n  bcDist(dpdy, nu_vyy) == " << bcerr << "\n} // PressureSolver::verify(p,u)\n";
    return l2err + bcerr;
}

}  // namespace chflow
