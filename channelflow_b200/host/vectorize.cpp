// field2vector / vector2field / fixDiri* (reference flowfield.cpp:4448-4760, utilfuncs.cpp:712-765): the map between a
// divergence-free, no-slip velocity field and the vector of its linearly independent real coefficients
// (Gibson, Halcrow & Cvitanovic 2008, table 1) that nsolver's Newton-Krylov / Arnoldi iterations work on.
// The maps run on the device (csrc/vecpack.cu, one warp per Fourier mode) on device-resident vectors (DeviceVector =
// cfgpu_vec); the raw-pointer forms of the reference API stage the vector through PCIe.
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

// ---- device state vectors (cfgpu_vec): the pack / unpack kernels run on the device (csrc/vecpack.cu); the raw-pointer
// forms below stage through a cached device vector of the right length
DeviceVector::DeviceVector(long n) { resize(n); }
DeviceVector::DeviceVector(const DeviceVector& o) {
    resize(o.n_);
    if (n_) cfgpu_check(cfgpu_vec_copy(v_, o.v_), "cfgpu_vec_copy");
}
DeviceVector& DeviceVector::operator=(const DeviceVector& o) {
    if (this == &o) return *this;
    resize(o.n_);
    if (n_) cfgpu_check(cfgpu_vec_copy(v_, o.v_), "cfgpu_vec_copy");
    return *this;
}
DeviceVector::~DeviceVector() {
    if (v_) cfgpu_vec_destroy(v_);
}
void DeviceVector::resize(long n) {
    if (n == n_ && v_) return;
    if (v_) cfgpu_vec_destroy(v_);
    v_ = nullptr;
    n_ = n;
    cfgpu_check(cfgpu_vec_create(cfgpu_context(), n, &v_), "cfgpu_vec_create");
}
void DeviceVector::setToZero() { cfgpu_check(cfgpu_vec_zero(v_), "cfgpu_vec_zero"); }
void DeviceVector::upload(const Real* x) { cfgpu_check(cfgpu_vec_upload(v_, x), "cfgpu_vec_upload"); }
void DeviceVector::download(Real* x) const { cfgpu_check(cfgpu_vec_download(v_, x), "cfgpu_vec_download"); }
Real DeviceVector::dot(const DeviceVector& o) const {
    Real r = 0;
    cfgpu_check(cfgpu_vec_dot(v_, o.v_, &r), "cfgpu_vec_dot");
    return r;
}
Real DeviceVector::norm() const {
    Real r = 0;
    cfgpu_check(cfgpu_vec_nrm2(v_, &r), "cfgpu_vec_nrm2");
    return r;
}
void DeviceVector::axpy(Real a, const DeviceVector& x) { cfgpu_check(cfgpu_vec_axpy(v_, a, x.v_), "cfgpu_vec_axpy"); }
void DeviceVector::axpby(Real a, const DeviceVector& x, Real b) { cfgpu_check(cfgpu_vec_axpby(v_, a, x.v_, b), "cfgpu_vec_axpby"); }
void DeviceVector::scale(Real s) { cfgpu_check(cfgpu_vec_scal(v_, s), "cfgpu_vec_scal"); }

void field2vector(const FlowField& u, DeviceVector& x) {
    assert(u.xzstate() == Spectral && u.ystate() == Spectral);
    const long N = field2vector_size(u);
    if (x.size() < N) x.resize(N);
    cfgpu_check(cfgpu_field2vector(u.device(), x.handle()), "cfgpu_field2vector");
}
void vector2field(const DeviceVector& x, FlowField& u) {
    u.setState(Spectral, Spectral);
    u.setPadded(true);
    cfgpu_check(cfgpu_vector2field(x.handle(), u.device_overwrite()), "cfgpu_vector2field");
}

static DeviceVector& staging(long n) {
    static DeviceVector v;
    if (v.size() != n) v.resize(n);
    return v;
}
void field2vector(const FlowField& u, Real* a) {
    DeviceVector& x = staging(field2vector_size(u));
    field2vector(u, x);
    x.download(a);
}
void vector2field(const Real* a, FlowField& u) {
    DeviceVector& x = staging(field2vector_size(u));
    x.upload(a);
    vector2field(x, u);
}

void fixdivnoslip(FlowField& u) {
    std::vector<Real> v(field2vector_size(u));
    field2vector(u, v.data());
    vector2field(v.data(), u);
}

}  // namespace chflow
