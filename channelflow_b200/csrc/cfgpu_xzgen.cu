// Generic in-place FlowField::makePhysical_xz / makeSpectral_xz (flowfield.cpp:1850-1886) on the full, padded
// reference layout (all modes incl. Nyquist) -- used by the FlowField API outside the fused DNS pipeline.
// x: batched c2c over mx for TZ kz-columns per CTA; z: two real lines share one complex FFT (c2r / r2c).
// The forward transform is scaled by 1/(Nx*Nz) like the reference.
#include "cfgpu_internal.h"

namespace cfgpu {
namespace {
constexpr int XG_THREADS = 256;

template <int DIR>
__global__ void __launch_bounds__(XG_THREADS) xgen_kernel(double2* __restrict__ c, int Nx, int Mz, int TZ, FftPlanDev plan) {
    double2* a = dyn_smem<double2>();
    double2* b = a + (size_t)Nx * TZ;
    const int tid = threadIdx.x;
    const int mz0 = blockIdx.x * TZ;
    double2* base = c + (size_t)blockIdx.y * Nx * Mz;  // (i, ny) plane
    for (int idx = tid; idx < Nx * TZ; idx += XG_THREADS) {
        const int mx = idx / TZ, cc = idx - mx * TZ;
        a[idx] = (mz0 + cc < Mz) ? base[(size_t)mx * Mz + mz0 + cc] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const double2* res = fft_smem<DIR>(a, b, plan, plan.tw, TZ, tid, XG_THREADS);
    for (int idx = tid; idx < Nx * TZ; idx += XG_THREADS) {
        const int mx = idx / TZ, cc = idx - mx * TZ;
        if (mz0 + cc < Mz) base[(size_t)mx * Mz + mz0 + cc] = res[idx];
    }
}

// lines = all (i, ny, nx); CTA handles 2*TP consecutive lines (TP complex columns)
__global__ void __launch_bounds__(XG_THREADS) zgen_c2r_kernel(double* __restrict__ r, long nlines, int Nz, int TP, FftPlanDev plan) {
    const int Mz = Nz / 2 + 1, Nzpad = 2 * Mz;
    double2* A = dyn_smem<double2>();
    double2* B = A + (size_t)Nz * TP;
    const int tid = threadIdx.x;
    const long l0 = (long)blockIdx.x * 2 * TP;
    for (int idx = tid; idx < Nz * TP; idx += XG_THREADS) A[idx] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int idx = tid; idx < TP * Mz; idx += XG_THREADS) {
        const int pr = idx / Mz, k = idx - pr * Mz;
        const long la = l0 + 2 * pr, lb = la + 1;
        if (la >= nlines) continue;
        const double2* ca = reinterpret_cast<const double2*>(r + la * Nzpad);
        double2 a = ca[k];
        double2 b = make_double2(0.0, 0.0);
        if (lb < nlines) b = reinterpret_cast<const double2*>(r + lb * Nzpad)[k];
        if (k == 0 || 2 * k == Nz) { a.y = 0.0; b.y = 0.0; }
        A[(size_t)k * TP + pr] = make_double2(a.x - b.y, a.y + b.x);
        if (k > 0 && 2 * k != Nz) A[(size_t)(Nz - k) * TP + pr] = make_double2(a.x + b.y, b.x - a.y);
    }
    __syncthreads();
    const double2* res = fft_smem<+1>(A, B, plan, plan.tw, TP, tid, XG_THREADS);
    for (int idx = tid; idx < TP * Nz; idx += XG_THREADS) {
        const int pr = idx / Nz, z = idx - pr * Nz;
        const long la = l0 + 2 * pr, lb = la + 1;
        if (la >= nlines) continue;
        const double2 v = res[(size_t)z * TP + pr];
        r[la * Nzpad + z] = v.x;
        if (lb < nlines) r[lb * Nzpad + z] = v.y;
    }
    // the two padding reals of each line are left as they are (FFTW leaves them unspecified)
}

__global__ void __launch_bounds__(XG_THREADS) zgen_r2c_kernel(double* __restrict__ r, long nlines, int Nz, int TP, double scale, FftPlanDev plan) {
    const int Mz = Nz / 2 + 1, Nzpad = 2 * Mz;
    double2* A = dyn_smem<double2>();
    double2* B = A + (size_t)Nz * TP;
    const int tid = threadIdx.x;
    const long l0 = (long)blockIdx.x * 2 * TP;
    for (int idx = tid; idx < TP * Nz; idx += XG_THREADS) {
        const int pr = idx / Nz, z = idx - pr * Nz;
        const long la = l0 + 2 * pr, lb = la + 1;
        double x = 0.0, y = 0.0;
        if (la < nlines) x = r[la * Nzpad + z];
        if (lb < nlines) y = r[lb * Nzpad + z];
        A[(size_t)z * TP + pr] = make_double2(x, y);
    }
    __syncthreads();
    const double2* g = fft_smem<-1>(A, B, plan, plan.tw, TP, tid, XG_THREADS);
    const double hs = 0.5 * scale;
    for (int idx = tid; idx < TP * Mz; idx += XG_THREADS) {
        const int pr = idx / Mz, k = idx - pr * Mz;
        const long la = l0 + 2 * pr, lb = la + 1;
        if (la >= nlines) continue;
        const int kn = (k == 0) ? 0 : Nz - k;
        const double2 gk = g[(size_t)k * TP + pr], gn = g[(size_t)kn * TP + pr];
        reinterpret_cast<double2*>(r + la * Nzpad)[k] = make_double2(hs * (gk.x + gn.x), hs * (gk.y - gn.y));
        if (lb < nlines) reinterpret_cast<double2*>(r + lb * Nzpad)[k] = make_double2(hs * (gk.y + gn.y), -hs * (gk.x - gn.x));
    }
}
}  // namespace
}  // namespace cfgpu

using namespace cfgpu;

extern "C" int cfgpu_xz_generic(cfgpu_field f, int to_physical) {
    cfgpu_ctx ctx = f->ctx;
    const FftPlanDev *fx, *fz;
    CF_TRY(get_fftplan(ctx, f->Nx, &fx));
    CF_TRY(get_fftplan(ctx, f->Nz, &fz));
    const int Mz = f->Mz();
    int TZ = 2304 / f->Nx;
    { int p = 16; while (p > TZ && p > 1) p >>= 1; TZ = p; }  // a single column per CTA for Nx > 2304
    if (getenv("CF_XGEN_TZ")) TZ = atoi(getenv("CF_XGEN_TZ"));
    int TP = 8;
    while (TP > 1 && (size_t)2 * f->Nz * TP * 16 > 96 * 1024) TP >>= 1;
    const size_t smx = 2 * (size_t)f->Nx * TZ * sizeof(double2);
    const size_t smz = 2 * (size_t)f->Nz * TP * sizeof(double2);
    const long nlines = (long)f->Nd * f->Ny * f->Nx;
    dim3 gx((Mz + TZ - 1) / TZ, f->Nd * f->Ny);
    dim3 gz((unsigned)((nlines + 2 * TP - 1) / (2 * TP)));
    static size_t cfg_xi = 0, cfg_xf = 0, cfg_zc = 0, cfg_zr = 0;
    if (to_physical) {
        auto kx = xgen_kernel<+1>;
        if (smx > cfg_xi) { CF_CUDA(cudaFuncSetAttribute(kx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smx)); cfg_xi = smx; }
        CF_LAUNCH(kx, gx, dim3(XG_THREADS), smx, ctx->stream, reinterpret_cast<double2*>(f->dser), f->Nx, Mz, TZ, *fx);
        CF_KERNEL_CHECK();
        auto kz = zgen_c2r_kernel;
        if (smz > cfg_zc) { CF_CUDA(cudaFuncSetAttribute(kz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smz)); cfg_zc = smz; }
        CF_LAUNCH(kz, gz, dim3(XG_THREADS), smz, ctx->stream, f->dser, nlines, f->Nz, TP, *fz);
        CF_KERNEL_CHECK();
    } else {
        auto kz = zgen_r2c_kernel;
        if (smz > cfg_zr) { CF_CUDA(cudaFuncSetAttribute(kz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smz)); cfg_zr = smz; }
        const double scale = 1.0 / ((double)f->Nx * (double)f->Nz);
        CF_LAUNCH(kz, gz, dim3(XG_THREADS), smz, ctx->stream, f->dser, nlines, f->Nz, TP, scale, *fz);
        CF_KERNEL_CHECK();
        auto kx = xgen_kernel<-1>;
        if (smx > cfg_xf) { CF_CUDA(cudaFuncSetAttribute(kx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smx)); cfg_xf = smx; }
        CF_LAUNCH(kx, gx, dim3(XG_THREADS), smx, ctx->stream, reinterpret_cast<double2*>(f->dser), f->Nx, Mz, TZ, *fx);
        CF_KERNEL_CHECK();
    }
    return 0;
}
