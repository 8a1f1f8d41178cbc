// chflow::FlowField -- Fourier x Chebyshev x Fourier field whose storage lives in B200 HBM.
//
// Same public surface as the hot-path part of the reference's channelflow/flowfield.h:52-345 (constructors,
// element access, transforms, state protocol, arithmetic, swap, dealiasing, geometry queries, .ff I/O), so that
// code written against Channelflow's FlowField keeps compiling.  Differences in mechanism, not in behaviour:
//   * data are an opaque cfgpu_field (include/cfgpu.h) in the reference's serial layout;
//   * operator()/cmplx() return references into a lazily synchronised host mirror (downloaded on first host
//     access after a device operation, uploaded before the next device operation);
//   * transforms / arithmetic / norms run as CUDA kernels on the context's stream.
#ifndef CFB200_FLOWFIELD_H
#define CFB200_FLOWFIELD_H
#include <cassert>
#include <string>
#include <vector>

#include "cfbasics/cfarray.h"
#include "cfbasics/cfvector.h"
#include "cfbasics/mathdefs.h"
#include "cfgpu.h"
#include "channelflow/basisfunc.h"
#include "channelflow/cfmpi.h"
#include "channelflow/chebyshev.h"

namespace chflow {

cfgpu_ctx cfgpu_context();  // process-wide device context (device = $CFGPU_DEVICE, else $LOCAL_RANK, else 0)
void cfgpu_check(int status, const char* where);

// std allocator over cfgpu_host_alloc: the host mirror lives in page-locked memory
template <class T>
struct PinnedAllocator {
    typedef T value_type;
    PinnedAllocator() {}
    template <class U> PinnedAllocator(const PinnedAllocator<U>&) {}
    T* allocate(std::size_t n) {
        void* p = nullptr;
        cfgpu_check(cfgpu_host_alloc(&p, (unsigned long long)(n * sizeof(T))), "cfgpu_host_alloc");
        return static_cast<T*>(p);
    }
    void deallocate(T* p, std::size_t) { cfgpu_host_free(p); }
    template <class U> bool operator==(const PinnedAllocator<U>&) const { return true; }
    template <class U> bool operator!=(const PinnedAllocator<U>&) const { return false; }
};

class FieldSymmetry;

class FlowField {
   public:
    FlowField();
    FlowField(int Nx, int Ny, int Nz, int Nd, Real Lx, Real Lz, Real a, Real b, CfMPI* cfmpi = nullptr,
              fieldstate xzstate = Spectral, fieldstate ystate = Spectral, uint fftw_flags = FFTW_ESTIMATE);
    FlowField(const FlowField& u);
    explicit FlowField(const std::string& filebase, CfMPI* cfmpi = nullptr);  // reads filebase.ff
    ~FlowField();
    FlowField& operator=(const FlowField& u);

    void resize(int Nx, int Ny, int Nz, int Nd, Real Lx, Real Lz, Real a, Real b, CfMPI* cfmpi = nullptr,
                uint fftw_flags = FFTW_ESTIMATE);
    void reconfig(const FlowField& u, uint fftw_flags = FFTW_ESTIMATE);
    void optimizeFFTW(uint = 0) {}

    // element access (host mirror)
    Real& operator()(int nx, int ny, int nz, int i);
    const Real& operator()(int nx, int ny, int nz, int i) const;
    Complex& cmplx(int mx, int my, int mz, int i);
    const Complex& cmplx(int mx, int my, int mz, int i) const;
    Real& operator()(int nx, int ny, int nz, int i, int j) { return (*this)(nx, ny, nz, i + Nd_ * j); }
    Complex& cmplx(int mx, int my, int mz, int i, int j) { return cmplx(mx, my, mz, i + Nd_ * j); }

    ComplexChebyCoeff profile(int mx, int mz, int i) const;
    BasisFunc profile(int mx, int mz) const;
    FlowField operator[](int i) const;                                 // the i-th component as a 1-component field
    void setComponent(int i, const FlowField& src, int j);             // this[i] = src[j] (same grid and state)

    // random divergence-free, no-slip perturbations on the libc drand48 stream (flowfield.cpp:2060-2177): the rule of
    // tools/randomfield.cpp; profiles are built on the host, the closing transform round trip runs on the device
    void perturb(Real magnitude, Real spectralDecay, bool meanflow = true);
    void addPerturbation(int kx, int kz, Real mag, Real spectralDecay);
    void addPerturbation1D(int kx, int kz, Real mag, Real spectralDecay);
    void addPerturbations(int kxmax, int kzmax, Real mag, Real spectralDecay, bool meanflow = true);
    void addPerturbations(Real magnitude, Real spectralDecay, bool meanflow = true);

    void makeSpectral_xz();
    void makePhysical_xz();
    void makeSpectral_y();
    void makePhysical_y();
    void makeSpectral();
    void makePhysical();
    void makeState(fieldstate xzstate, fieldstate ystate);

    void setToZero();
    void interpolate(FlowField f);   // spectral interpolation of f onto this grid (same box), flowfield.cpp:691-792

    int numXmodes() const { return Nx_; }
    int numYmodes() const { return Ny_; }
    int numZmodes() const { return Nz_ / 2 + 1; }
    int numXgridpts() const { return Nx_; }
    int numYgridpts() const { return Ny_; }
    int numZgridpts() const { return Nz_; }
    int vectorDim() const { return Nd_; }
    int Nx() const { return Nx_; }
    int Ny() const { return Ny_; }
    int Nz() const { return Nz_; }
    int Nd() const { return Nd_; }
    int Mx() const { return Nx_; }
    int My() const { return Ny_; }
    int Mz() const { return Nz_ / 2 + 1; }
    lint Nloc() const { return (lint)Nx_ * Ny_ * Nzpad() * Nd_; }
    lint Nxloc() const { return Nx_; }
    lint nxlocmin() const { return 0; }
    lint Mxloc() const { return Nx_; }
    lint mxlocmin() const { return 0; }
    lint Nyloc() const { return Ny_; }
    lint nylocmin() const { return 0; }
    lint nylocmax() const { return Ny_; }
    lint Mzloc() const { return Mz(); }
    lint mzlocmin() const { return 0; }

    int mx(int kx) const { return kx >= 0 ? kx : kx + Nx_; }
    int mz(int kz) const { return kz; }
    int kx(int mx) const { return mx <= Nx_ / 2 ? mx : mx - Nx_; }
    int kz(int mz) const { return mz; }
    int kxmax() const { return Nx_ / 2; }
    int kzmax() const { return Nz_ / 2; }
    int kxmin() const { return Nx_ / 2 + 1 - Nx_; }
    int kzmin() const { return 0; }
    int kxmaxDealiased() const { return Nx_ / 3 - 1; }
    int kzmaxDealiased() const { return Nz_ / 3 - 1; }
    int kxminDealiased() const { return -(Nx_ / 3 - 1); }
    int kzminDealiased() const { return 0; }
    bool isAliased(int kx, int kz) const { return std::abs(kx) > kxmaxDealiased() || std::abs(kz) > kzmaxDealiased(); }

    Real Lx() const { return Lx_; }
    Real Ly() const { return b_ - a_; }
    Real Lz() const { return Lz_; }
    Real a() const { return a_; }
    Real b() const { return b_; }
    Real x(int nx) const { return nx * Lx_ / Nx_; }
    Real y(int ny) const { return 0.5 * ((b_ + a_) + (b_ - a_) * cos(pi * ny / (Ny_ - 1))); }
    Real z(int nz) const { return nz * Lz_ / Nz_; }

    int nproc0() const { return 1; }
    int nproc1() const { return 1; }
    int taskid() const { return 0; }
    int numtasks() const { return 1; }
    int task_coeff(int, int) const { return 0; }
    CfMPI* cfmpi() const { return cfmpi_; }

    Complex Dx(int mx) const;
    Complex Dz(int mz) const;
    Complex Dx(int mx, int n) const;
    Complex Dz(int mz, int n) const;
    Vector xgridpts() const;
    Vector ygridpts() const;
    Vector zgridpts() const;
    lint Nxlocmax() const { return Nx_; }
    lint Nylocpad() const { return Ny_; }
    lint Nypad() const { return Ny_; }
    int key0() const { return 0; }
    int color0() const { return 0; }
    int taskid_world() const { return 0; }
    int task_coeffp(int, int) const { return 0; }
    int task_coeff(int, int, int, int) const { return 0; }
    MPI_Comm* comm_world() const { static MPI_Comm c = 0; return &c; }

    FlowField& operator*=(Real x);
    FlowField& operator*=(const FieldSymmetry& s);  // u <- s(u), on the device (symmetry.cpp)
    FlowField& project(const FieldSymmetry& s);      // u <- (u + s u)/2
    FlowField& project(const cfarray<FieldSymmetry>& s);
    FlowField& operator+=(const Real& a);        // u(0,0,0,0) += a
    FlowField& operator-=(const Real& a);
    FlowField& operator+=(const ComplexChebyCoeff& U);
    FlowField& operator-=(const ComplexChebyCoeff& U);
    FlowField& operator+=(const BasisFunc& U);   // adds U to mode (U.kx, U.kz)
    FlowField& operator-=(const BasisFunc& U);
    bool congruent(const BasisFunc& phi) const;
    FlowField& operator+=(const ChebyCoeff& U);  // u(0,*,0,0) += U
    FlowField& operator-=(const ChebyCoeff& U);
    FlowField& operator+=(const std::vector<ChebyCoeff>& UW);
    FlowField& operator-=(const std::vector<ChebyCoeff>& UW);
    FlowField& operator+=(const FlowField& u);
    FlowField& operator-=(const FlowField& u);
    void add(const Real a, const FlowField& u);
    void add(const Real a, const FlowField& u, const Real b, const FlowField& v);

    bool geomCongruent(const FlowField& f, Real eps = 1e-13) const;
    bool congruent(const FlowField& f, Real eps = 1e-13) const;
    friend void swap(FlowField& f, FlowField& g);
// Lagrange interpolant sum_n w_n(mu) un[n] through (mun[n], un[n]), box lengths and wall positions interpolated alike
// (flowfield.cpp:4136-4287; there point by point on the physical grid -- the interpolant is linear in the data, so here it
// is length(un) axpy's on the device in whatever state the fields are in).  The result is spectral, padded modes zeroed
// when the inputs' are.
FlowField quadraticInterpolate(cfarray<FlowField>& un, const cfarray<Real>& mun, Real mu, Real eps = 1e-13);
FlowField polynomialInterpolate(cfarray<FlowField>& un, cfarray<Real>& mun, Real mu);

    void binarySave(const std::string& filebase) const;
    void asciiSave(const std::string& filebase) const;
    void save(const std::string& filebase, std::vector<std::string> component_names = std::vector<std::string>()) const;
    // NetCDF with the reference's dimensions / variables / attributes (flowfield.cpp:3225-3597), classic CDF-2 container (ncfile.cpp)
    void writeNetCDF(const std::string& filebase, std::vector<std::string> component_names = std::vector<std::string>()) const;
    void saveProfile(int mx, int mz, const std::string& filebase) const;
    void saveProfile(int mx, int mz, const std::string& filebase, const ChebyTransform& t) const;
    void saveSpectrum(const std::string& filebase, int i, int ny = -1, bool kxorder = true, bool showpadding = false) const;
    void saveSpectrum(const std::string& filebase, bool kxorder = true, bool showpadding = false) const;
    Real energy(bool normalize = true) const;
    Real energy(int mx, int mz, bool normalize = true) const;
    void rescale(Real Lx, Real Lz);

    Real dudy_a() const;
    Real dudy_b() const;
    Real dwdy_a() const;
    Real dwdy_b() const;
    Real CFLfactor() const;
    Real CFLfactor(ChebyCoeff Ubase, ChebyCoeff Wbase) const;

    void setState(fieldstate xz, fieldstate y);
    void assertState(fieldstate xz, fieldstate y) const;
    fieldstate xzstate() const { return xzstate_; }
    fieldstate ystate() const { return ystate_; }

    void zeroPaddedModes();
    void setPadded(bool b);
    bool padded() const { return padded_; }

    // ---- device side (used by NSE / diffops; not part of the reference API)
    bool box_ok() const;
    void upload_from(const Real* data) const;
    cfgpu_field device() const;          // device copy is current on return; host mirror stays valid
    cfgpu_field device_mut();            // as above, and the host mirror is invalidated
    cfgpu_field device_overwrite();      // the caller overwrites the whole field on the device: nothing is uploaded first
    void raw_upload(const Real* data);   // whole array, reference layout
    void raw_download(Real* data) const;

   private:
    int Nx_ = 0, Ny_ = 0, Nz_ = 0, Nd_ = 0;
    Real Lx_ = 0, Lz_ = 0, a_ = 0, b_ = 0;
    bool padded_ = false;
    fieldstate xzstate_ = Spectral, ystate_ = Spectral;
    CfMPI* cfmpi_ = nullptr;

    mutable cfgpu_field dev_ = nullptr;
    mutable std::vector<Real, PinnedAllocator<Real>> host_;
    mutable bool host_valid_ = false;  // host mirror holds the current data
    mutable bool host_box_clean_ = false;  // host mirror is known to be zero outside the retained (de-aliased) box
    mutable bool dev_valid_ = true;    // device copy holds the current data

    int Nzpad() const { return 2 * (Nz_ / 2 + 1); }
    size_t flatten(int nx, int ny, int nz, int i) const { return nz + (size_t)Nzpad() * (nx + (size_t)Nx_ * (ny + (size_t)Ny_ * i)); }
    size_t complex_flatten(int mx, int my, int mz, int i) const {
        return mz + (size_t)(Nz_ / 2 + 1) * (mx + (size_t)Nx_ * (my + (size_t)Ny_ * i));
    }
    void host_sync() const;   // make the host mirror current
    void host_dirty();        // host mirror current and modified => device stale
    void push_state() const;  // copy state flags to the device object
};

FlowField operator*(const Real a, const FlowField& w);
FlowField operator+(const FlowField& v, const FlowField& w);
FlowField operator-(const FlowField& v, const FlowField& w);
void swap(FlowField& f, FlowField& g);

// Device-resident vector of reals (cfgpu_vec): the state vectors of the Newton-Krylov / Arnoldi iterations stay in HBM;
// dot / norm / axpy replace the Eigen VectorXd algebra of nsolver (cfbasics.h:711-780, gmres.cpp:37-102).
class DeviceVector {
   public:
    DeviceVector() {}
    explicit DeviceVector(long n);
    DeviceVector(const DeviceVector& o);
    DeviceVector& operator=(const DeviceVector& o);
    ~DeviceVector();
    void resize(long n);  // contents are zeroed
    long size() const { return n_; }
    void setToZero();
    void upload(const Real* x);
    void download(Real* x) const;
    Real dot(const DeviceVector& o) const;
    Real norm() const;
    void axpy(Real a, const DeviceVector& x);            // this += a x
    void axpby(Real a, const DeviceVector& x, Real b);   // this = a x + b this
    void scale(Real s);
    cfgpu_vec handle() const { return v_; }

   private:
    cfgpu_vec v_ = nullptr;
    long n_ = 0;
};
void field2vector(const FlowField& u, DeviceVector& x);
void vector2field(const DeviceVector& x, FlowField& u);

// The field2vector and vector2field functions assume zero divergence and no-slip BCs (reference flowfield.h:617-621).
// Raw-pointer forms plus adaptors for any vector type with size()/resize()/data() (Eigen::VectorXd, std::vector).
int field2vector_size(const FlowField& u);
void field2vector(const FlowField& u, Real* a);
void vector2field(const Real* a, FlowField& u);
void fixdivnoslip(FlowField& u);
template <class Vec>
inline auto field2vector(const FlowField& u, Vec& v) -> decltype(v.resize(1), void()) {
    const int N = field2vector_size(u);
    if ((int)v.size() < N) v.resize(N);
    field2vector(u, v.data());
}
template <class Vec>
inline auto vector2field(const Vec& v, FlowField& u) -> decltype(v.size(), void()) {
    vector2field(v.data(), u);
}
// utilfuncs.h:76-81
void fixDiri(ChebyCoeff& f);
void fixDiriMean(ChebyCoeff& f);
void fixDiri(ComplexChebyCoeff& f);
void fixDiriMean(ComplexChebyCoeff& f);

}  // namespace chflow
#endif
