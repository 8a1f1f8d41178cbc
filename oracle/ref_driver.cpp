/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * C-ABI driver around the *unmodified* Channelflow reference classes (compiled from /root/reference by
 * oracle/Makefile into oracle/_ref/libchflow_ref.so).  Lets the Python tests and bench.py's CPU arms call
 * the reference's own FlowField / DNS / NSE / TauSolver / HelmholtzSolver code through ctypes.
 * Nothing under channelflow_b200/ links or loads this library.
 */
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

#include "channelflow/chebyshev.h"
#include "channelflow/diffops.h"
#include "channelflow/dns.h"
#include "channelflow/dnsflags.h"
#include "channelflow/flowfield.h"
#include "channelflow/symmetry.h"
#include "channelflow/helmholtz.h"
#include "channelflow/nse.h"
#include "channelflow/tausolver.h"
#include "channelflow/utilfuncs.h"

using namespace chflow;

// stands in for the cmake-generated GitSHA1.cpp (template GitSHA1.cpp.in)
const char* g_GIT_SHA1 = "oracle-build";

extern "C" {

struct RefFlags {
    double nu, dPdx, dPdz, Ubulk, Wbulk, ulowerwall, uupperwall, wlowerwall, wupperwall, Vsuck, rotation, t0, dt;
    int baseflow, constraint, timestepping, initstepping, nonlinearity, dealiasing, taucorrection;
};

static std::ostringstream g_sink;

static DNSFlags to_flags(const RefFlags* f) {
    DNSFlags fl;
    fl.nu = f->nu;
    fl.dPdx = f->dPdx;
    fl.dPdz = f->dPdz;
    fl.Ubulk = f->Ubulk;
    fl.Wbulk = f->Wbulk;
    fl.ulowerwall = f->ulowerwall;
    fl.uupperwall = f->uupperwall;
    fl.wlowerwall = f->wlowerwall;
    fl.wupperwall = f->wupperwall;
    fl.Vsuck = f->Vsuck;
    fl.rotation = f->rotation;
    fl.t0 = f->t0;
    fl.dt = f->dt;
    fl.baseflow = (BaseFlow)f->baseflow;
    fl.constraint = (MeanConstraint)f->constraint;
    fl.timestepping = (TimeStepMethod)f->timestepping;
    fl.initstepping = (TimeStepMethod)f->initstepping;
    fl.nonlinearity = (NonlinearMethod)f->nonlinearity;
    fl.dealiasing = (Dealiasing)f->dealiasing;
    fl.taucorrection = f->taucorrection != 0;
    fl.verbosity = Silent;
    fl.logstream = &g_sink;
    return fl;
}

// ---------------------------------------------------------------- FlowField
void* ref_field_create(int Nx, int Ny, int Nz, int Nd, double Lx, double Lz, double a, double b) {
    return new FlowField(Nx, Ny, Nz, Nd, Lx, Lz, a, b);
}
void ref_field_free(void* h) { delete (FlowField*)h; }
// pointer to the raw storage (reference serial layout nz + Nzpad*(nx + Nx*(ny + Ny*i)))
double* ref_field_data(void* h) {
    FlowField* u = (FlowField*)h;
    return reinterpret_cast<double*>(&u->cmplx(0, 0, 0, 0));
}
long ref_field_nloc(void* h) {
    FlowField* u = (FlowField*)h;
    return (long)u->Nx() * u->Ny() * (2 * (u->Nz() / 2 + 1)) * u->Nd();
}
void ref_field_set_state(void* h, int xz, int y) { ((FlowField*)h)->setState((fieldstate)xz, (fieldstate)y); }
void ref_field_get_state(void* h, int* xz, int* y) {
    *xz = (int)((FlowField*)h)->xzstate();
    *y = (int)((FlowField*)h)->ystate();
}
void ref_field_set_padded(void* h, int p) { ((FlowField*)h)->setPadded(p != 0); }
int ref_field_padded(void* h) { return ((FlowField*)h)->padded() ? 1 : 0; }
void ref_field_copy(void* dst, void* src) { *(FlowField*)dst = *(FlowField*)src; }
void ref_field_zero(void* h) { ((FlowField*)h)->setToZero(); }
void ref_make_physical(void* h) { ((FlowField*)h)->makePhysical(); }
void ref_make_spectral(void* h) { ((FlowField*)h)->makeSpectral(); }
void ref_make_physical_y(void* h) { ((FlowField*)h)->makePhysical_y(); }
void ref_make_spectral_y(void* h) { ((FlowField*)h)->makeSpectral_y(); }
void ref_make_physical_xz(void* h) { ((FlowField*)h)->makePhysical_xz(); }
void ref_make_spectral_xz(void* h) { ((FlowField*)h)->makeSpectral_xz(); }
void ref_zero_padded_modes(void* h) { ((FlowField*)h)->zeroPaddedModes(); }
// flowfield.cpp:1274-1433
void ref_field_symmetry(void* h, int s, int sx, int sy, int sz, double ax, double az) { *((FlowField*)h) *= FieldSymmetry(sx, sy, sz, ax, az, s); }

// tools/randomfield.cpp:50-67
void ref_randomfield(void* h, int seed, double magn, double smooth, int meanflow) {
    FlowField& u = *(FlowField*)h;
    srand48(seed);
    u.setToZero();
    u.setState(Spectral, Spectral);
    u.addPerturbations(u.kxmaxDealiased(), u.kzmaxDealiased(), 1.0, 1 - smooth, meanflow != 0);
    u *= magn / L2Norm(u);
    u.setPadded(true);
}

// FlowField::readNetCDF semantics (flowfield.cpp:3605-3826) for the "IO without padded modes" branch:
// var[c] has dims (Nz_io, Ny, Nx_io); the field must already have the full-grid geometry.
void ref_load_padded_physical(void* h, const double* var, int Nx_io, int Nz_io) {
    FlowField& u = *(FlowField*)h;
    const int Ny = u.Ny(), Nd = u.Nd();
    const int Mz_io = Nz_io / 2 + 1, Nzpad_io = 2 * Mz_io;
    const long Nloc_io = 2L * Nx_io * Mz_io * Ny * Nd;
    double* rio = (double*)fftw_malloc(Nloc_io * sizeof(double));
    memset(rio, 0, Nloc_io * sizeof(double));
    for (int i = 0; i < Nd; ++i)
        for (int nx = 0; nx < Nx_io; ++nx)
            for (int ny = 0; ny < Ny; ++ny)
                for (int nz = 0; nz < Nz_io; ++nz)
                    rio[nz + Nzpad_io * (nx + (long)Nx_io * (ny + Ny * i))] =
                        var[(long)i * Nz_io * Ny * Nx_io + nx + (long)Nx_io * (ny + (long)Ny * nz)];
    u.setState(Physical, Physical);
    u.setPadded(true);
    u.addPaddedModes(rio, Nx_io, 0, Mz_io, 0);
    u.makeSpectral();
    fftw_free(rio);
}

double ref_l2norm(void* h) { return L2Norm(*(FlowField*)h); }
double ref_l2norm2(void* h, int normalize) { return L2Norm2(*(FlowField*)h, normalize != 0); }
double ref_l2norm3d(void* h) { return L2Norm3d(*(FlowField*)h); }
double ref_l2dist(void* a, void* b) { return L2Dist(*(FlowField*)a, *(FlowField*)b); }
double ref_l2ip(void* a, void* b) { return L2InnerProduct(*(FlowField*)a, *(FlowField*)b); }
double ref_divnorm(void* h) { return divNorm(*(FlowField*)h); }
double ref_wallshear(void* h) { return wallshear(*(FlowField*)h); }
double ref_dissipation(void* h) { return dissipation(*(FlowField*)h); }
double ref_bcnorm(void* h) { return bcNorm(*(FlowField*)h); }

int ref_field2vector_size(void* h) { return field2vector_size(*(FlowField*)h); }
void ref_field2vector(void* h, double* x) {
    Eigen::VectorXd v;
    field2vector(*(FlowField*)h, v);
    for (long i = 0; i < (long)v.size(); ++i) x[i] = v(i);
}
void ref_vector2field(const double* x, long n, void* h) {
    Eigen::VectorXd v(n);
    for (long i = 0; i < n; ++i) v(i) = x[i];
    vector2field(v, *(FlowField*)h);
}

// ---------------------------------------------------------------- NSE / nonlinear
// f = NSE::nonlinear(u) (navierstokesNL + zeroPaddedModes), nse.cpp:383-391
void ref_nonlinear(void* uh, void* fh, const RefFlags* rf) {
    DNSFlags flags = to_flags(rf);
    FlowField& u = *(FlowField*)uh;
    FlowField q(u.Nx(), u.Ny(), u.Nz(), 1, u.Lx(), u.Lz(), u.a(), u.b());
    std::vector<FlowField> fields = {u, q};
    NSE nse(fields, flags);
    std::vector<FlowField> out = {*(FlowField*)fh};
    nse.nonlinear(fields, out);
    *(FlowField*)fh = out[0];
}

void ref_base_profiles(void* uh, const RefFlags* rf, double* Ubase, double* Wbase) {
    DNSFlags flags = to_flags(rf);
    FlowField& u = *(FlowField*)uh;
    FlowField q(u.Nx(), u.Ny(), u.Nz(), 1, u.Lx(), u.Lz(), u.a(), u.b());
    std::vector<FlowField> fields = {u, q};
    NSE nse(fields, flags);
    for (int n = 0; n < u.Ny(); ++n) {
        Ubase[n] = nse.Ubase()[n];
        Wbase[n] = nse.Wbase()[n];
    }
}

// ---------------------------------------------------------------- DNS
struct RefDNS {
    std::vector<FlowField> fields;
    DNS* dns;
};

void* ref_dns_create(void* uh, void* qh, const RefFlags* rf) {
    RefDNS* d = new RefDNS();
    d->fields = {*(FlowField*)uh, *(FlowField*)qh};
    d->dns = new DNS(d->fields, to_flags(rf));
    return d;
}
void ref_dns_free(void* h) {
    RefDNS* d = (RefDNS*)h;
    delete d->dns;
    delete d;
}
void ref_dns_advance(void* h, int n) {
    RefDNS* d = (RefDNS*)h;
    d->dns->advance(d->fields, n);
}
void ref_dns_get(void* h, void* uh, void* qh) {
    RefDNS* d = (RefDNS*)h;
    if (uh) *(FlowField*)uh = d->fields[0];
    if (qh) *(FlowField*)qh = d->fields[1];
}
void ref_dns_set(void* h, void* uh, void* qh) {
    RefDNS* d = (RefDNS*)h;
    if (uh) d->fields[0] = *(FlowField*)uh;
    if (qh) d->fields[1] = *(FlowField*)qh;
}
double ref_dns_cfl(void* h) {
    RefDNS* d = (RefDNS*)h;
    return d->dns->CFL(d->fields[0]);
}
void ref_dns_reset_dt(void* h, double dt) { ((RefDNS*)h)->dns->reset_dt(dt); }
double ref_dns_time(void* h) { return ((RefDNS*)h)->dns->time(); }
double ref_dns_dPdx(void* h) { return ((RefDNS*)h)->dns->dPdx(); }
double ref_dns_Ubulk(void* h) { return ((RefDNS*)h)->dns->Ubulk(); }

// ---------------------------------------------------------------- 1-D solvers
// TauSolver::solve (tausolver.cpp:347-402). Profiles are length-N arrays (re, im separately).
void ref_tausolve(int kx, int kz, double Lx, double Lz, double a, double b, double lambda, double nu, int N,
                  int taucorr, const double* Rx_re, const double* Rx_im, const double* Ry_re, const double* Ry_im,
                  const double* Rz_re, const double* Rz_im, double* out /* u,v,w,P each re[N],im[N] */) {
    TauSolver ts(kx, kz, Lx, Lz, a, b, lambda, nu, N, taucorr != 0);
    ComplexChebyCoeff u(N, a, b, Spectral), v(N, a, b, Spectral), w(N, a, b, Spectral), P(N, a, b, Spectral);
    ComplexChebyCoeff Rx(N, a, b, Spectral), Ry(N, a, b, Spectral), Rz(N, a, b, Spectral);
    for (int n = 0; n < N; ++n) {
        Rx.re[n] = Rx_re[n]; Rx.im[n] = Rx_im[n];
        Ry.re[n] = Ry_re[n]; Ry.im[n] = Ry_im[n];
        Rz.re[n] = Rz_re[n]; Rz.im[n] = Rz_im[n];
    }
    ts.solve(u, v, w, P, Rx, Ry, Rz);
    ComplexChebyCoeff* o[4] = {&u, &v, &w, &P};
    for (int c = 0; c < 4; ++c)
        for (int n = 0; n < N; ++n) {
            out[(2 * c) * N + n] = o[c]->re[n];
            out[(2 * c + 1) * N + n] = o[c]->im[n];
        }
}

// bulk-velocity-constrained variant for kx=kz=0 (tausolver.cpp:404-450)
void ref_tausolve_bulk(double Lx, double Lz, double a, double b, double lambda, double nu, int N, int taucorr,
                       const double* Rx_re, const double* Ry_re, const double* Rz_re, double umean, double wmean,
                       double* out, double* dPdx, double* dPdz) {
    TauSolver ts(0, 0, Lx, Lz, a, b, lambda, nu, N, taucorr != 0);
    ComplexChebyCoeff u(N, a, b, Spectral), v(N, a, b, Spectral), w(N, a, b, Spectral), P(N, a, b, Spectral);
    ComplexChebyCoeff Rx(N, a, b, Spectral), Ry(N, a, b, Spectral), Rz(N, a, b, Spectral);
    for (int n = 0; n < N; ++n) {
        Rx.re[n] = Rx_re[n];
        Ry.re[n] = Ry_re[n];
        Rz.re[n] = Rz_re[n];
    }
    ts.solve(u, v, w, P, *dPdx, *dPdz, Rx, Ry, Rz, umean, wmean);
    ComplexChebyCoeff* o[4] = {&u, &v, &w, &P};
    for (int c = 0; c < 4; ++c)
        for (int n = 0; n < N; ++n) {
            out[(2 * c) * N + n] = o[c]->re[n];
            out[(2 * c + 1) * N + n] = o[c]->im[n];
        }
}

// HelmholtzSolver::solve (helmholtz.cpp:79-95)
void ref_helmholtz(int N, double a, double b, double lambda, double nu, const double* f, double ua, double ub,
                   double* u) {
    HelmholtzSolver h(N, a, b, lambda, nu);
    ChebyCoeff uu(N, a, b, Spectral), ff(N, a, b, Spectral);
    for (int n = 0; n < N; ++n) ff[n] = f[n];
    h.solve(uu, ff, ua, ub);
    for (int n = 0; n < N; ++n) u[n] = uu[n];
}

// ChebyTransform round trip helpers (chebyshev.cpp:262-302)
void ref_cheby_make_physical(int N, double* c) {
    ChebyCoeff u(N, -1, 1, Spectral);
    for (int n = 0; n < N; ++n) u[n] = c[n];
    ChebyTransform t(N);
    u.makePhysical(t);
    for (int n = 0; n < N; ++n) c[n] = u[n];
}
void ref_cheby_make_spectral(int N, double* c) {
    ChebyCoeff u(N, -1, 1, Physical);
    for (int n = 0; n < N; ++n) u[n] = c[n];
    ChebyTransform t(N);
    u.makeSpectral(t);
    for (int n = 0; n < N; ++n) c[n] = u[n];
}

void ref_laminar_profile(const RefFlags* rf, double a, double b, int Ny, double* U) {
    DNSFlags flags = to_flags(rf);
    ChebyCoeff u = laminarProfile(flags, a, b, Ny);
    for (int n = 0; n < Ny; ++n) U[n] = u[n];
}

}  // extern "C"
