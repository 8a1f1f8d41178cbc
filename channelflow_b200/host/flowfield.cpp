// chflow::FlowField over the cfgpu C-ABI.  See channelflow/flowfield.h.
#include "channelflow/flowfield.h"

#include <algorithm>

#include <arpa/inet.h>

#include <cstring>
#include <fstream>

#include "channelflow/nse.h"

namespace chflow {

// ------------------------------------------------------------------------------------------- context
cfgpu_ctx cfgpu_context() {
    static cfgpu_ctx ctx = nullptr;
    if (!ctx) {
        int dev = 0;
        if (const char* e = std::getenv("CFGPU_DEVICE")) dev = std::atoi(e);
        else if (const char* l = std::getenv("LOCAL_RANK")) dev = std::atoi(l);
        if (cfgpu_init(dev, &ctx) != 0) cferror(std::string("cfgpu_init failed: ") + cfgpu_last_error());
    }
    return ctx;
}
void cfgpu_check(int status, const char* where) {
    if (status != 0) cferror(std::string(where) + ": " + cfgpu_last_error());
}
#define CK(call) cfgpu_check((call), #call)

// ------------------------------------------------------------------------------------------- ctors
FlowField::FlowField() {}

FlowField::FlowField(int Nx, int Ny, int Nz, int Nd, Real Lx, Real Lz, Real a, Real b, CfMPI* cfmpi, fieldstate xzstate,
                     fieldstate ystate, uint) {
    resize(Nx, Ny, Nz, Nd, Lx, Lz, a, b, cfmpi);
    xzstate_ = xzstate;
    ystate_ = ystate;
}

FlowField::FlowField(const FlowField& u) { *this = u; }

FlowField::~FlowField() {
    if (dev_) cfgpu_field_destroy(dev_);
}

FlowField& FlowField::operator=(const FlowField& u) {
    if (this == &u) return *this;
    if (!u.dev_) {
        if (dev_) cfgpu_field_destroy(dev_);
        dev_ = nullptr;
        Nx_ = Ny_ = Nz_ = Nd_ = 0;
        return *this;
    }
    resize(u.Nx_, u.Ny_, u.Nz_, u.Nd_, u.Lx_, u.Lz_, u.a_, u.b_, u.cfmpi_);
    xzstate_ = u.xzstate_;
    ystate_ = u.ystate_;
    padded_ = u.padded_;
    CK(cfgpu_field_copy(dev_, u.device()));
    dev_valid_ = true;
    host_valid_ = false;
    return *this;
}

void FlowField::resize(int Nx, int Ny, int Nz, int Nd, Real Lx, Real Lz, Real a, Real b, CfMPI* cfmpi, uint) {
    if (dev_ && Nx == Nx_ && Ny == Ny_ && Nz == Nz_ && Nd == Nd_ && Lx == Lx_ && Lz == Lz_ && a == a_ && b == b_) return;
    if (dev_) cfgpu_field_destroy(dev_);
    dev_ = nullptr;
    Nx_ = Nx; Ny_ = Ny; Nz_ = Nz; Nd_ = Nd; Lx_ = Lx; Lz_ = Lz; a_ = a; b_ = b;
    cfmpi_ = cfmpi;
    host_.clear();
    host_valid_ = false;
    dev_valid_ = true;
    if (Nx > 0 && Ny > 0 && Nz > 0 && Nd > 0) CK(cfgpu_field_create(cfgpu_context(), Nx, Ny, Nz, Nd, Lx, Lz, a, b, &dev_));
}

void FlowField::reconfig(const FlowField& u, uint) {
    resize(u.Nx(), u.Ny(), u.Nz(), u.Nd(), u.Lx(), u.Lz(), u.a(), u.b(), u.cfmpi());
    setToZero();
}

// ------------------------------------------------------------------------------------------- mirror
void FlowField::push_state() const {
    CK(cfgpu_field_set_state(dev_, xzstate_ == Spectral ? CFGPU_SPECTRAL : CFGPU_PHYSICAL,
                             ystate_ == Spectral ? CFGPU_SPECTRAL : CFGPU_PHYSICAL));
    CK(cfgpu_field_set_padded(dev_, padded_ ? 1 : 0));
}
void FlowField::host_sync() const {
    if (host_valid_) return;
    const bool fresh = host_.empty();
    host_.resize((size_t)Nloc());  // value-initialised: zeros
    if (padded_ && xzstate_ == Spectral && box_ok()) {
        // only the retained box travels; whatever an earlier (physical / un-padded) state left elsewhere must not survive
        if (!fresh && !host_box_clean_) std::fill(host_.begin(), host_.end(), 0.0);
        CK(cfgpu_field_download_padded(dev_, host_.data()));
        host_box_clean_ = true;
    } else {
        CK(cfgpu_field_download(dev_, host_.data()));
        host_box_clean_ = false;
    }
    host_valid_ = true;
}
void FlowField::host_dirty() {
    host_sync();
    host_box_clean_ = false;  // the caller may write anywhere
    dev_valid_ = false;
}
cfgpu_field FlowField::device() const {
    if (!dev_) cferror("FlowField: operation on an empty field");
    if (!dev_valid_) {
        upload_from(host_.data());
        dev_valid_ = true;
    }
    push_state();
    return dev_;
}
cfgpu_field FlowField::device_overwrite() {
    if (!dev_) cferror("FlowField: operation on an empty field");
    dev_valid_ = true;
    host_valid_ = false;
    push_state();
    return dev_;
}
cfgpu_field FlowField::device_mut() {
    cfgpu_field d = device();
    host_valid_ = false;
    return d;
}
// de-aliased spectral fields travel as their retained box only (include/cfgpu.h: cfgpu_field_upload_padded)
bool FlowField::box_ok() const { return Nx_ / 3 - 1 >= 0 && Nz_ / 3 - 1 >= 0; }
void FlowField::upload_from(const Real* data) const {
    if (padded_ && xzstate_ == Spectral && box_ok()) CK(cfgpu_field_upload_padded(dev_, data, ystate_ == Spectral));
    else CK(cfgpu_field_upload(dev_, data, xzstate_ == Spectral, ystate_ == Spectral));
}
void FlowField::raw_upload(const Real* data) {
    upload_from(data);
    dev_valid_ = true;
    host_valid_ = false;
}
void FlowField::raw_download(Real* data) const {
    cfgpu_field d = device();
    if (padded_ && xzstate_ == Spectral && box_ok()) CK(cfgpu_field_download_padded(d, data));
    else CK(cfgpu_field_download(d, data));
}

Real& FlowField::operator()(int nx, int ny, int nz, int i) {
    assert(xzstate_ == Physical);
    host_dirty();
    return host_[flatten(nx, ny, nz, i)];
}
const Real& FlowField::operator()(int nx, int ny, int nz, int i) const {
    assert(xzstate_ == Physical);
    host_sync();
    return host_[flatten(nx, ny, nz, i)];
}
Complex& FlowField::cmplx(int mx, int my, int mz, int i) {
    assert(xzstate_ == Spectral);
    host_dirty();
    return reinterpret_cast<Complex*>(host_.data())[complex_flatten(mx, my, mz, i)];
}
const Complex& FlowField::cmplx(int mx, int my, int mz, int i) const {
    assert(xzstate_ == Spectral);
    host_sync();
    return reinterpret_cast<const Complex*>(host_.data())[complex_flatten(mx, my, mz, i)];
}

ComplexChebyCoeff FlowField::profile(int mx, int mz, int i) const {
    ComplexChebyCoeff p(Ny_, a_, b_, ystate_);
    std::vector<Real> buf(2 * (size_t)Ny_);
    CK(cfgpu_field_get_profile(device(), mx, mz, i, buf.data()));
    for (int n = 0; n < Ny_; ++n) p.set(n, Complex(buf[2 * n], buf[2 * n + 1]));
    return p;
}

// ------------------------------------------------------------------------------------------- transforms
void FlowField::makeSpectral_xz() {
    if (xzstate_ == Spectral) return;
    CK(cfgpu_field_make_spectral_xz(device_mut()));
    xzstate_ = Spectral;
}
void FlowField::makePhysical_xz() {
    if (xzstate_ == Physical) return;
    CK(cfgpu_field_make_physical_xz(device_mut()));
    xzstate_ = Physical;
}
void FlowField::makeSpectral_y() {
    if (ystate_ == Spectral) return;
    CK(cfgpu_field_make_spectral_y(device_mut()));
    ystate_ = Spectral;
}
void FlowField::makePhysical_y() {
    if (ystate_ == Physical) return;
    CK(cfgpu_field_make_physical_y(device_mut()));
    ystate_ = Physical;
}
void FlowField::makeSpectral() {
    makeSpectral_xz();
    makeSpectral_y();
}
void FlowField::makePhysical() {
    makePhysical_y();
    makePhysical_xz();
}
void FlowField::makeState(fieldstate xz, fieldstate y) {
    if (y == Physical && xz == Physical) makePhysical();
    else if (y == Spectral && xz == Spectral) makeSpectral();
    else if (y == Physical && xz == Spectral) { makeSpectral_xz(); makePhysical_y(); }
    else { makeSpectral_y(); makePhysical_xz(); }
}

void FlowField::setToZero() {
    if (!dev_) return;
    CK(cfgpu_field_zero(dev_));
    dev_valid_ = true;
    host_valid_ = false;
}

Complex FlowField::Dx(int mx) const {
    const int k = kx(mx);
    return Complex(0.0, 2 * pi * k / Lx_ * ((k == kxmax()) ? 0 : 1));
}
Complex FlowField::Dz(int mz) const {
    const int k = kz(mz);
    return Complex(0.0, 2 * pi * k / Lz_ * ((k == kzmax()) ? 0 : 1));
}

// ------------------------------------------------------------------------------------------- arithmetic
FlowField& FlowField::operator*=(Real x) {
    CK(cfgpu_field_scale(device_mut(), x));
    return *this;
}
static void add_profile(FlowField& u, cfgpu_field d, int i, const ChebyCoeff& U, Real s) {
    std::vector<Real> buf(2 * (size_t)u.Ny(), 0.0);
    for (int n = 0; n < u.Ny() && n < U.length(); ++n) buf[2 * n] = U[n];
    cfgpu_check(cfgpu_field_add_profile(d, 0, 0, i, buf.data(), s), "cfgpu_field_add_profile");
}
FlowField& FlowField::operator+=(const ChebyCoeff& U) {
    assert(xzstate_ == Spectral && ystate_ == U.state());
    add_profile(*this, device_mut(), 0, U, 1.0);
    return *this;
}
FlowField& FlowField::operator-=(const ChebyCoeff& U) {
    add_profile(*this, device_mut(), 0, U, -1.0);
    return *this;
}
FlowField& FlowField::operator+=(const std::vector<ChebyCoeff>& UW) {
    cfgpu_field d = device_mut();
    add_profile(*this, d, 0, UW[0], 1.0);
    add_profile(*this, d, 2, UW[1], 1.0);
    return *this;
}
FlowField& FlowField::operator-=(const std::vector<ChebyCoeff>& UW) {
    cfgpu_field d = device_mut();
    add_profile(*this, d, 0, UW[0], -1.0);
    add_profile(*this, d, 2, UW[1], -1.0);
    return *this;
}
FlowField& FlowField::operator+=(const FlowField& u) {
    add(1.0, u);
    return *this;
}
FlowField& FlowField::operator-=(const FlowField& u) {
    add(-1.0, u);
    return *this;
}
void FlowField::add(const Real a, const FlowField& u) {
    assert(congruent(u));
    CK(cfgpu_field_axpby(device_mut(), a, u.device(), 0.0, nullptr));
}
void FlowField::add(const Real a, const FlowField& u, const Real b, const FlowField& v) {
    assert(congruent(u) && congruent(v));
    CK(cfgpu_field_axpby(device_mut(), a, u.device(), b, v.device()));
}

bool FlowField::geomCongruent(const FlowField& v, Real eps) const {
    return (v.Nx_ == Nx_ && v.Ny_ == Ny_ && v.Nz_ == Nz_ && std::abs(v.Lx_ - Lx_) / Greater(Lx_, 1.0) < eps &&
            std::abs(v.Lz_ - Lz_) / Greater(Lz_, 1.0) < eps && std::abs(v.a_ - a_) / Greater(std::abs(a_), 1.0) < eps &&
            std::abs(v.b_ - b_) / Greater(std::abs(b_), 1.0) < eps);
}
bool FlowField::congruent(const FlowField& v, Real eps) const {
    return geomCongruent(v, eps) && v.Nd_ == Nd_ && v.xzstate_ == xzstate_ && v.ystate_ == ystate_;
}

void swap(FlowField& f, FlowField& g) {
    assert(f.congruent(g));
    // O(1): exchange device handles, mirrors and flags (reference: flowfield.cpp:4076-4090)
    std::swap(f.dev_, g.dev_);
    std::swap(f.host_, g.host_);
    std::swap(f.host_valid_, g.host_valid_);
    std::swap(f.host_box_clean_, g.host_box_clean_);
    std::swap(f.dev_valid_, g.dev_valid_);
}

void FlowField::setState(fieldstate xz, fieldstate y) {
    xzstate_ = xz;
    ystate_ = y;
}
void FlowField::assertState(fieldstate xz, fieldstate y) const { assert(xzstate_ == xz && ystate_ == y); (void)xz; (void)y; }

void FlowField::zeroPaddedModes() {
    fieldstate xzs = xzstate_;
    makeSpectral_xz();
    CK(cfgpu_field_zero_padded_modes(device_mut()));
    padded_ = true;
    makeState(xzs, ystate_);
}
void FlowField::setPadded(bool b) { padded_ = b; }

// ------------------------------------------------------------------------------------------- diagnostics
static Real wall_slope(const FlowField& u, int i, bool upper) {
    ComplexChebyCoeff p = u.profile(0, 0, i);
    ChebyCoeff d = diff(p.re);
    return upper ? d.eval_b() : d.eval_a();
}
Real FlowField::dudy_a() const { return wall_slope(*this, 0, false); }
Real FlowField::dudy_b() const { return wall_slope(*this, 0, true); }
Real FlowField::dwdy_a() const { return wall_slope(*this, 2, false); }
Real FlowField::dwdy_b() const { return wall_slope(*this, 2, true); }

Real FlowField::CFLfactor(ChebyCoeff Ubase, ChebyCoeff Wbase) const {
    // max over the grid of (u_i + U_i)/dx_i, signed (reference: flowfield.cpp:4035-4068); computed on the device
    // through a throw-away NSE operator (base flow only matters through U, W)
    DNSFlags flags;
    flags.baseflow = ArbitraryBase;
    flags.dealiasing = padded_ ? DealiasXZ : NoDealiasing;
    Ubase.makeSpectral();
    Wbase.makeSpectral();
    return NSE::cflfactor_of(*this, Ubase, Wbase, flags);
}
Real FlowField::CFLfactor() const {
    ChebyCoeff Z(Ny_, a_, b_, Spectral);
    return CFLfactor(Z, Z);
}

// ------------------------------------------------------------------------------------------- .ff files
// Channelflow's native binary format (reference writer flowfield.cpp:2642-2725, reader :359-428): big-endian.
namespace {
void wr_int(std::ostream& os, int n) { uint32_t v = htonl((uint32_t)n); os.write((char*)&v, 4); }
void wr_real(std::ostream& os, Real x) {
    unsigned char b[8];
    std::memcpy(b, &x, 8);
    for (int i = 0; i < 4; ++i) std::swap(b[i], b[7 - i]);
    os.write((char*)b, 8);
}
void wr_char(std::ostream& os, char c) { os.write(&c, 1); }
int rd_int(std::istream& is) { uint32_t v; is.read((char*)&v, 4); return (int)ntohl(v); }
Real rd_real(std::istream& is) {
    unsigned char b[8];
    is.read((char*)b, 8);
    for (int i = 0; i < 4; ++i) std::swap(b[i], b[7 - i]);
    Real x;
    std::memcpy(&x, b, 8);
    return x;
}
char rd_char(std::istream& is) { char c; is.read(&c, 1); return c; }
std::string with_ff(const std::string& f) { return (f.size() > 3 && f.substr(f.size() - 3) == ".ff") ? f : f + ".ff"; }
}  // namespace

void FlowField::binarySave(const std::string& filebase) const {
    std::ofstream os(with_ff(filebase).c_str(), std::ios::out | std::ios::binary);
    if (!os.good()) cferror("FlowField::binarySave(filebase) : can't open file " + with_ff(filebase));
    wr_int(os, 2); wr_int(os, 0); wr_int(os, 9);
    wr_int(os, Nx_); wr_int(os, Ny_); wr_int(os, Nz_); wr_int(os, Nd_);
    wr_char(os, xzstate_ == Spectral ? 'S' : 'P'); wr_char(os, ystate_ == Spectral ? 'S' : 'P');
    wr_real(os, Lx_); wr_real(os, Lz_); wr_real(os, a_); wr_real(os, b_);
    wr_char(os, padded_ ? '1' : '0');
    host_sync();
    const Real* r = host_.data();
    if (padded_ && xzstate_ == Spectral) {
        const int Nxd = 2 * (Nx_ / 6), Nzd = 2 * (Nz_ / 3) + 1;
        for (int i = 0; i < Nd_; ++i)
            for (int ny = 0; ny < Ny_; ++ny) {
                for (int nx = 0; nx <= Nxd; ++nx)
                    for (int nz = 0; nz <= Nzd; ++nz) wr_real(os, r[flatten(nx, ny, nz, i)]);
                for (int nx = Nx_ - Nxd; nx < Nx_; ++nx)
                    for (int nz = 0; nz <= Nzd; ++nz) wr_real(os, r[flatten(nx, ny, nz, i)]);
            }
    } else {
        const size_t N = (size_t)Nloc();
        for (size_t k = 0; k < N; ++k) wr_real(os, r[k]);
    }
}

bool read_netcdf_field(const std::string& filename, FlowField& u, CfMPI* cfmpi);  // ncfile.cpp

// filebase, filebase.ff (the reference's binary format) or filebase.nc (NetCDF-4 as written by stock Channelflow); with no
// suffix given .nc is tried first, as in flowfield.cpp:130-190
FlowField::FlowField(const std::string& filebase, CfMPI* cfmpi) {
    const bool want_nc = hasSuffix(filebase, ".nc"), want_ff = hasSuffix(filebase, ".ff");
    if (hasSuffix(filebase, ".h5")) cferror("FlowField(filebase) : HDF5 field files are not supported, convert with fieldconvert: " + filebase);
    if (!want_ff && read_netcdf_field(want_nc ? filebase : filebase + ".nc", *this, cfmpi)) return;
    if (want_nc) cferror("FlowField(filebase) : can't open " + filebase);
    std::ifstream is(with_ff(filebase).c_str(), std::ios::in | std::ios::binary);
    if (!is.good()) cferror("FlowField(filebase) : can't open " + with_ff(filebase) + " or " + filebase + ".nc");
    rd_int(is); rd_int(is); rd_int(is);
    const int Nx = rd_int(is), Ny = rd_int(is), Nz = rd_int(is), Nd = rd_int(is);
    const fieldstate xz = rd_char(is) == 'S' ? Spectral : Physical;
    const fieldstate ys = rd_char(is) == 'S' ? Spectral : Physical;
    const Real Lx = rd_real(is), Lz = rd_real(is), a = rd_real(is), b = rd_real(is);
    const bool padded = rd_char(is) == '1';
    resize(Nx, Ny, Nz, Nd, Lx, Lz, a, b, cfmpi);
    xzstate_ = xz; ystate_ = ys; padded_ = padded;
    host_.assign((size_t)Nloc(), 0.0);
    host_box_clean_ = false;
    if (padded && xz == Spectral) {
        const int Nxd = 2 * (Nx_ / 6), Nzd = 2 * (Nz_ / 3) + 1;
        for (int i = 0; i < Nd_; ++i)
            for (int ny = 0; ny < Ny_; ++ny) {
                for (int nx = 0; nx <= Nxd; ++nx)
                    for (int nz = 0; nz <= Nzd; ++nz) host_[flatten(nx, ny, nz, i)] = rd_real(is);
                for (int nx = Nx_ - Nxd; nx < Nx_; ++nx)
                    for (int nz = 0; nz <= Nzd; ++nz) host_[flatten(nx, ny, nz, i)] = rd_real(is);
            }
    } else {
        for (size_t k = 0; k < host_.size(); ++k) host_[k] = rd_real(is);
    }
    host_valid_ = true;
    dev_valid_ = false;
    makeSpectral();
}

// ------------------------------------------------------------------------------------------- free operators
FlowField operator*(const Real a, const FlowField& w) {
    FlowField v(w);
    v *= a;
    return v;
}
FlowField operator+(const FlowField& v, const FlowField& w) {
    FlowField r(v);
    r += w;
    return r;
}
FlowField operator-(const FlowField& v, const FlowField& w) {
    FlowField r(v);
    r -= w;
    return r;
}

}  // namespace chflow
