// Flat C driver over the host classes (chflow::FlowField / NSE / DNS): the same entry points, with the same
// argument meaning, as oracle/ref_driver.cpp offers over the reference's classes, so the parity tests can run
// one script against both libraries.  Prefix cf_ (reference driver: ref_).
#include <cstring>
#include <sstream>

#include "channelflow/diffops.h"
#include "channelflow/dns.h"
#include "channelflow/devicesearch.h"
#include "channelflow/symmetry.h"

using namespace chflow;

extern "C" {

struct CfFlags {
    double nu, dPdx, dPdz, Ubulk, Wbulk, ulowerwall, uupperwall, wlowerwall, wupperwall, Vsuck, rotation, t0, dt;
    int baseflow, constraint, timestepping, initstepping, nonlinearity, dealiasing, taucorrection;
};

static std::ostringstream g_sink;

static DNSFlags to_flags(const CfFlags* f) {
    DNSFlags fl;
    fl.nu = f->nu; fl.dPdx = f->dPdx; fl.dPdz = f->dPdz; fl.Ubulk = f->Ubulk; fl.Wbulk = f->Wbulk;
    fl.ulowerwall = f->ulowerwall; fl.uupperwall = f->uupperwall; fl.wlowerwall = f->wlowerwall; fl.wupperwall = f->wupperwall;
    fl.Vsuck = f->Vsuck; fl.rotation = f->rotation; fl.t0 = f->t0; fl.dt = f->dt;
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

void* cf_field_create(int Nx, int Ny, int Nz, int Nd, double Lx, double Lz, double a, double b) {
    return new FlowField(Nx, Ny, Nz, Nd, Lx, Lz, a, b);
}
void cf_field_free(void* h) { delete (FlowField*)h; }
long cf_field_nloc(void* h) { return (long)((FlowField*)h)->Nloc(); }
void cf_field_upload(void* h, const double* data) { ((FlowField*)h)->raw_upload(data); }
void cf_field_download(void* h, double* data) { ((FlowField*)h)->raw_download(data); }
void cf_field_set_state(void* h, int xz, int y) { ((FlowField*)h)->setState((fieldstate)xz, (fieldstate)y); }
void cf_field_get_state(void* h, int* xz, int* y) {
    *xz = (int)((FlowField*)h)->xzstate();
    *y = (int)((FlowField*)h)->ystate();
}
void cf_field_set_padded(void* h, int p) { ((FlowField*)h)->setPadded(p != 0); }
int cf_field_padded(void* h) { return ((FlowField*)h)->padded() ? 1 : 0; }
void cf_field_copy(void* dst, void* src) { *(FlowField*)dst = *(FlowField*)src; }
void cf_field_zero(void* h) { ((FlowField*)h)->setToZero(); }
void cf_make_physical(void* h) { ((FlowField*)h)->makePhysical(); }
void cf_make_spectral(void* h) { ((FlowField*)h)->makeSpectral(); }
void cf_make_physical_y(void* h) { ((FlowField*)h)->makePhysical_y(); }
void cf_make_spectral_y(void* h) { ((FlowField*)h)->makeSpectral_y(); }
void cf_make_physical_xz(void* h) { ((FlowField*)h)->makePhysical_xz(); }
void cf_make_spectral_xz(void* h) { ((FlowField*)h)->makeSpectral_xz(); }
void cf_zero_padded_modes(void* h) { ((FlowField*)h)->zeroPaddedModes(); }
void cf_field_symmetry(void* h, int s, int sx, int sy, int sz, double ax, double az) { *((FlowField*)h) *= FieldSymmetry(sx, sy, sz, ax, az, s); }
double cf_cmplx_get(void* h, int mx, int my, int mz, int i, int part) {
    const FlowField& u = *(FlowField*)h;
    const Complex c = u.cmplx(mx, my, mz, i);
    return part ? c.imag() : c.real();
}
void cf_cmplx_set(void* h, int mx, int my, int mz, int i, double re, double im) {
    ((FlowField*)h)->cmplx(mx, my, mz, i) = Complex(re, im);
}
void cf_field_save(void* h, const char* filebase) { ((FlowField*)h)->save(filebase); }  // suffix picks .ff (default) / .nc / .asc
void* cf_field_load(const char* filebase) { return new FlowField(std::string(filebase)); }

double cf_l2norm(void* h) { return L2Norm(*(FlowField*)h); }
double cf_l2norm2(void* h, int normalize) { return L2Norm2(*(FlowField*)h, normalize != 0); }
double cf_l2dist(void* a, void* b) { return L2Dist(*(FlowField*)a, *(FlowField*)b); }
double cf_l2ip(void* a, void* b) { return L2InnerProduct(*(FlowField*)a, *(FlowField*)b); }
void cf_field_axpby(void* y, double a, void* x, double b, void* z) {
    if (z) ((FlowField*)y)->add(a, *(FlowField*)x, b, *(FlowField*)z);
    else ((FlowField*)y)->add(a, *(FlowField*)x);
}
void cf_field_scale(void* y, double s) { *(FlowField*)y *= s; }

// ---- device state vectors (DeviceVector over cfgpu_vec)
void* cf_vec_create(long n) { return new DeviceVector(n); }
void cf_vec_free(void* v) { delete (DeviceVector*)v; }
long cf_vec_size(void* v) { return ((DeviceVector*)v)->size(); }
void cf_vec_upload(void* v, const double* x) { ((DeviceVector*)v)->upload(x); }
void cf_vec_download(void* v, double* x) { ((DeviceVector*)v)->download(x); }
double cf_vec_dot(void* a, void* b) { return ((DeviceVector*)a)->dot(*(DeviceVector*)b); }
double cf_vec_norm(void* a) { return ((DeviceVector*)a)->norm(); }
void cf_vec_axpy(void* y, double a, void* x) { ((DeviceVector*)y)->axpy(a, *(DeviceVector*)x); }
void cf_vec_axpby(void* y, double a, void* x, double b) { ((DeviceVector*)y)->axpby(a, *(DeviceVector*)x, b); }
void cf_vec_scale(void* y, double s) { ((DeviceVector*)y)->scale(s); }
void cf_field2vector_dev(void* u, void* v) { field2vector(*(FlowField*)u, *(DeviceVector*)v); }
void cf_vector2field_dev(void* v, void* u) { vector2field(*(DeviceVector*)v, *(FlowField*)u); }

void cf_nonlinear(void* uh, void* fh, const CfFlags* rf) {
    DNSFlags flags = to_flags(rf);
    FlowField& u = *(FlowField*)uh;
    FlowField q(u.Nx(), u.Ny(), u.Nz(), 1, u.Lx(), u.Lz(), u.a(), u.b());
    std::vector<FlowField> fields = {u, q};
    NSE nse(fields, flags);
    std::vector<FlowField> out = {*(FlowField*)fh};
    nse.nonlinear(fields, out);
    *(FlowField*)fh = out[0];
}
void cf_base_profiles(void* uh, const CfFlags* rf, double* Ubase, double* Wbase) {
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

struct CfDNS {
    std::vector<FlowField> fields;
    DNS* dns;
};
void* cf_dns_create(void* uh, void* qh, const CfFlags* rf) {
    CfDNS* d = new CfDNS();
    d->fields = {*(FlowField*)uh, *(FlowField*)qh};
    d->dns = new DNS(d->fields, to_flags(rf));
    return d;
}
void cf_dns_free(void* h) {
    CfDNS* d = (CfDNS*)h;
    delete d->dns;
    delete d;
}
void cf_dns_advance(void* h, int n) {
    CfDNS* d = (CfDNS*)h;
    d->dns->advance(d->fields, n);
}
void cf_dns_get(void* h, void* uh, void* qh) {
    CfDNS* d = (CfDNS*)h;
    if (uh) *(FlowField*)uh = d->fields[0];
    if (qh) *(FlowField*)qh = d->fields[1];
}
void cf_dns_set(void* h, void* uh, void* qh) {
    CfDNS* d = (CfDNS*)h;
    if (uh) d->fields[0] = *(FlowField*)uh;
    if (qh) d->fields[1] = *(FlowField*)qh;
}
// u <- sigma u on the DNS' own velocity field and on the time-stepping history (DNS::operator*=, dns.cpp:175-180)
void cf_dns_symmetry(void* h, int s, int sx, int sy, int sz, double ax, double az) {
    CfDNS* d = (CfDNS*)h;
    const FieldSymmetry sigma(sx, sy, sz, ax, az, s);
    d->fields[0] *= sigma;
    *d->dns *= std::vector<FieldSymmetry>{sigma, FieldSymmetry()};
}
double cf_dns_cfl(void* h) {
    CfDNS* d = (CfDNS*)h;
    return d->dns->CFL(d->fields[0]);
}
void cf_dns_reset_dt(void* h, double dt) { ((CfDNS*)h)->dns->reset_dt(dt); }
double cf_dns_time(void* h) { return ((CfDNS*)h)->dns->time(); }
double cf_dns_dPdx(void* h) { return ((CfDNS*)h)->dns->dPdx(); }
double cf_dns_Ubulk(void* h) { return ((CfDNS*)h)->dns->Ubulk(); }
double cf_l2norm3d(void* uh) { return L2Norm3d(*(FlowField*)uh); }
int cf_field2vector_size(void* uh) { return field2vector_size(*(FlowField*)uh); }
void cf_field2vector(void* uh, double* x) { field2vector(*(FlowField*)uh, x); }
void cf_vector2field(const double* x, void* uh) { vector2field(x, *(FlowField*)uh); }
void cf_sync() { cfgpu_sync(cfgpu_context()); }
// ---- multi-GPU plumbing (one process per GPU; the launcher distributes the NCCL id)
int cf_comm_unique_id(void* id128) { return cfgpu_comm_unique_id(id128); }
void cf_comm_init_nccl(int rank, int nranks, const void* id128) {
    cfgpu_check(cfgpu_comm_init_nccl(cfgpu_context(), rank, nranks, id128), "cfgpu_comm_init_nccl");
}
void cf_comm_init_external(int rank, int nranks, cfgpu_exchange_fn ex, cfgpu_allreduce_fn ar) {
    cfgpu_check(cfgpu_comm_init_external(cfgpu_context(), rank, nranks, ex, ar, nullptr), "cfgpu_comm_init_external");
}
void cf_comm_ranges(int nmx, int Ny, int rank, int* r4) {
    cfgpu_check(cfgpu_comm_ranges(cfgpu_context(), nmx, Ny, rank, r4, r4 + 1, r4 + 2, r4 + 3), "cfgpu_comm_ranges");
}
void cf_field_allgather(void* uh) {
    FlowField& u = *(FlowField*)uh;
    cfgpu_check(cfgpu_field_allgather(u.device_mut()), "cfgpu_field_allgather");
}
long long cf_launch_count() {
    long long n = 0;
    cfgpu_launch_count(cfgpu_context(), &n);
    return n;
}
void cf_timer_start() { cfgpu_timer_start(cfgpu_context()); }
double cf_timer_stop() {
    double ms = 0;
    cfgpu_timer_stop(cfgpu_context(), &ms);
    return ms;
}

void cf_profile_enable(int on) { cfgpu_profile_enable(cfgpu_context(), on); }
void cf_profile_read(double* ms, long long* calls, int reset) { cfgpu_profile_read(cfgpu_context(), ms, calls, reset); }

// Poincare section search (channelflow/dns.h: DNSPoincare): strides of nSteps steps from (u, q) until the section is crossed
// (at most maxstrides).  kind 0: h = wallshear - dissipation (DragDissipation); kind 1: h = (u, estar) - (ustar, estar)
// (PlaneIntersection).  u, q end at the last stride's state; ucross / pcross (may be null) receive the crossing.
// out = {found, tcrossing, hcrossing, scrossing, strides taken, hcurrent}
void cf_poincare_search(void* uh, void* qh, const CfFlags* rf, int kind, void* ustar, void* estar, int nSteps, int maxstrides,
                        int crosssign, double Tmin, double eps, double* out, void* ucross, void* pcross) {
    FlowField& u = *(FlowField*)uh;
    FlowField& q = *(FlowField*)qh;
    std::unique_ptr<PoincareCondition> h;
    if (kind == 0) h.reset(new DragDissipation());
    else h.reset(new PlaneIntersection(*(FlowField*)ustar, *(FlowField*)estar));
    DNSPoincare dns(u, h.get(), to_flags(rf));
    bool found = false;
    int n = 0;
    while (n < maxstrides && !found) {
        found = dns.advanceToSection(u, q, nSteps, crosssign, Tmin, eps);
        ++n;
    }
    out[0] = found ? 1.0 : 0.0;
    out[1] = found ? dns.tcrossing() : 0.0;
    out[2] = found ? dns.hcrossing() : 0.0;
    out[3] = found ? dns.scrossing() : 0.0;
    out[4] = n;
    out[5] = dns.hcurrent();
    if (found && ucross) *(FlowField*)ucross = dns.ucrossing();
    if (found && pcross) *(FlowField*)pcross = dns.pcrossing();
}

// Newton-Krylov-hookstep search on the device (channelflow/devicesearch.h): u is the initial guess and receives the solution.
// sigma = {s, sx, sy, sz, ax, az} (ax, az updated on return); par = {epsSearch, epsGMRES, epsDx, delta, Nnewton, Ngmres, Nhook,
// Tnormalize, verbose, xrelative, zrelative};
// out = {converged, newtonSteps, fevals, gmresIterations, residual, steps_per_eval, history[0..]} (at most nout entries)
void cf_hookstep_search(void* uh, const CfFlags* rf, double T, double dt, double* sigma, const double* par, double* out, int nout) {
    FlowField& u = *(FlowField*)uh;
    DNSFlags flags = to_flags(rf);
    flags.dt = dt;
    TimeStep ts(dt, dt, dt, 1.0, 0.0, 1e9, false);
    FieldSymmetry sg((int)sigma[1], (int)sigma[2], (int)sigma[3], sigma[4], sigma[5], (int)sigma[0]);
    DeviceSearchFlags sf;
    sf.epsSearch = par[0]; sf.epsGMRES = par[1]; sf.epsDx = par[2]; sf.delta = par[3];
    sf.Nnewton = (int)par[4]; sf.Ngmres = (int)par[5]; sf.Nhook = (int)par[6];
    std::ostringstream sink;
    if (par[8] == 0) sf.logstream = &sink;
    u.makeSpectral();
    DeviceDSI dsi(u, flags, ts, sg, T, par[7] != 0, par[9] != 0, par[10] != 0);
    DeviceVector x;
    dsi.makeVector(u, x);
    const DeviceSearchResult r = hookstepSearch(dsi, x, sf);
    dsi.extractVector(x, u);
    sigma[4] = dsi.sigma().ax();
    sigma[5] = dsi.sigma().az();
    const double head[6] = {r.converged ? 1.0 : 0.0, (double)r.newtonSteps, (double)r.fevals, (double)r.gmresIterations, r.residual, dsi.steps_per_eval()};
    for (int i = 0; i < nout; ++i) out[i] = i < 6 ? head[i] : (i - 6 < (int)r.history.size() ? r.history[i - 6] : -1.0);
}
// tools/randomfield.cpp:49-64: the reference's random initial condition (serial drand48 stream, seed, Gaussian coefficients with
// spectral decay 1 - smooth over the whole retained box, divergence-free, no-slip, rescaled to L2Norm magn)
void cf_randomfield(void* h, int seed, double magn, double smooth, int meanflow) {
    FlowField& u = *(FlowField*)h;
    srand48(seed);
    u.setToZero();
    u.setState(Spectral, Spectral);
    u.addPerturbations(u.kxmaxDealiased(), u.kzmaxDealiased(), 1.0, 1.0 - smooth, meanflow != 0);
    u *= magn / L2Norm(u);
    u.setPadded(true);
}
void cf_laminar_profile(const CfFlags* rf, double a, double b, int Ny, double* U) {
    DNSFlags flags = to_flags(rf);
    ChebyCoeff u = laminarProfile(flags, a, b, Ny);
    for (int n = 0; n < Ny; ++n) U[n] = u[n];
}

// TimeStep (dnsflags.cpp:749-975) for the host-logic tests
void* cf_timestep_create(double dt, double dtmin, double dtmax, double dT, double CFLmin, double CFLmax, int variable) {
    return new TimeStep(dt, dtmin, dtmax, dT, CFLmin, CFLmax, variable != 0);
}
void cf_timestep_free(void* h) { delete (TimeStep*)h; }
int cf_timestep_adjust(void* h, double cfl) { return ((TimeStep*)h)->adjust(cfl, false) ? 1 : 0; }
int cf_timestep_adjust_for_T(void* h, double T) { return ((TimeStep*)h)->adjust_for_T(T, false) ? 1 : 0; }
int cf_timestep_n(void* h) { return ((TimeStep*)h)->n(); }
int cf_timestep_N(void* h) { return ((TimeStep*)h)->N(); }
double cf_timestep_dt(void* h) { return ((TimeStep*)h)->dt(); }
double cf_timestep_dT(void* h) { return ((TimeStep*)h)->dT(); }
double cf_timestep_CFL(void* h) { return ((TimeStep*)h)->CFL(); }

}  // extern "C"
