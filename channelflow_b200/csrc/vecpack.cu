// Device side of nsolver's state vectors:
//
//  * field2vector / vector2field (reference flowfield.cpp:4448-4760, utilfuncs.cpp:712-765): the map between a
//    divergence-free, no-slip velocity field and the vector of its linearly independent real coefficients (Gibson,
//    Halcrow & Cvitanovic 2008, table 1).  One WARP per Fourier mode: the pack is a pure gather; the unpack rebuilds the
//    boundary conditions (fixDiri / fixDiriMean: reductions over the profile) and v from continuity (Chebyshev integration:
//    a local three-term stencil) in shared memory and writes the three profiles of the mode (and the conjugate kx < 0
//    partner on the kz = 0 plane).
//  * Krylov-vector algebra (cfbasics.h:711-780 L2IP/L2Norm of VectorXd, the modified Gram-Schmidt loop of
//    nsolver/gmres.cpp:37-102): dot, nrm2, axpy, scal, copy on device-resident vectors; reductions are two-pass
//    (per-CTA partial sums in a fixed order, then one CTA), so results are deterministic run to run.
//
// Vector layout (the packing ORDER is a data format shared with the reference): [ (0,0): u re, w re (ny >= 2) ]
// [ (kx,0), kx = 1..Kx ] [ (0,kz), kz = 1..Kz ] [ (kx,kz), kx = -Kx..Kx without 0 (outer), kz = 1..Kz (inner) ]; per mode
// 2(Ny-2) + 2(Ny-4) reals: the "free" component's coefficients 2..Ny-1, then the constrained one's 3..Ny-2, (re, im).
// Roofline: HBM (every retained coefficient read / written once).
#include "vecpack.cuh"

namespace cfgpu {

namespace {
constexpr int VP_WARPS = 4;
constexpr int VP_THREADS = 32 * VP_WARPS;
constexpr double TWO_PI = 6.283185307179586476925286766559;

// mode j of the vector's enumeration -> (kx, kz) and the offset of its block
__device__ __forceinline__ void vp_mode(int j, const PackGeom& g, int& kx, int& kz, long& off) {
    const long S0 = 2L * (g.Ny - 2), S = 2L * (g.Ny - 2) + 2L * (g.Ny - 4);
    if (j == 0) { kx = 0; kz = 0; off = 0; return; }
    off = S0 + (long)(j - 1) * S;
    if (j <= g.Kx) { kx = j; kz = 0; return; }
    if (j <= g.Kx + g.Kz) { kx = 0; kz = j - g.Kx; return; }
    const int gidx = j - (1 + g.Kx + g.Kz);
    const int kxi = gidx / g.Kz;
    kz = gidx - kxi * g.Kz + 1;
    kx = kxi < g.Kx ? kxi - g.Kx : kxi - g.Kx + 1;
}
__device__ __forceinline__ long vp_addr(const PackGeom& g, int kx, int kz, int ny, int i) {  // complex index, serial layout
    const int mx = kx >= 0 ? kx : g.Nx + kx;
    return kz + (long)(g.Nz / 2 + 1) * (mx + (long)g.Nx * (ny + (long)g.Ny * i));
}

__global__ void __launch_bounds__(VP_THREADS) field2vector_kernel(const double2* __restrict__ u, double* __restrict__ a, const PackGeom g, int nmodes) {
    const int lane = threadIdx.x & 31, j = blockIdx.x * VP_WARPS + (threadIdx.x >> 5);
    if (j >= nmodes) return;
    int kx, kz;
    long off;
    vp_mode(j, g, kx, kz, off);
    const int Ny = g.Ny;
    if (j == 0) {
        for (int ny = 2 + lane; ny < Ny; ny += 32) {
            a[off + ny - 2] = u[vp_addr(g, 0, 0, ny, 0)].x;
            a[off + (Ny - 2) + ny - 2] = u[vp_addr(g, 0, 0, ny, 2)].x;
        }
        return;
    }
    const int first = (kz == 0) ? 2 : 0, second = (kz == 0) ? 0 : 2;  // on the kz = 0 plane w is the free component
    double2* a2 = reinterpret_cast<double2*>(a + off);
    for (int ny = 2 + lane; ny < Ny; ny += 32) a2[ny - 2] = u[vp_addr(g, kx, kz, ny, first)];
    a2 += Ny - 2;
    for (int ny = 3 + lane; ny < Ny - 1; ny += 32) a2[ny - 3] = u[vp_addr(g, kx, kz, ny, second)];
}

// warp-wide helpers on one real profile f[0..N) in shared memory
__device__ __forceinline__ void wsums(const double* f, int N, int lane, double& at_b, double& at_a, double& mean) {
    double sb = 0.0, sa = 0.0, sm = 0.0;
    for (int n = lane; n < N; n += 32) {
        const double v = f[n];
        sb += v;
        sa += (n & 1) ? -v : v;
        if (n >= 2 && !(n & 1)) sm -= v / (double)(n * n - 1);
    }
    at_b = warp_sum(sb);
    at_a = warp_sum(sa);
    mean = f[0] + warp_sum(sm);  // chebyshev.cpp:505-512
}
__device__ __forceinline__ void fix_diri(double* f, int N, int lane) {  // utilfuncs.cpp:712-719
    double fb, fa, fm;
    wsums(f, N, lane, fb, fa, fm);
    __syncwarp();
    if (lane == 0) { f[0] -= 0.5 * (fb + fa); f[1] -= 0.5 * (fb - fa); }
    __syncwarp();
}
__device__ __forceinline__ void fix_diri_mean(double* f, int N, int lane) {  // utilfuncs.cpp:721-728
    double fb, fa, fm;
    wsums(f, N, lane, fb, fa, fm);
    __syncwarp();
    if (lane == 0) {
        f[0] -= 0.125 * (fa + fb) + 0.75 * fm;
        f[1] -= 0.5 * (fb - fa);
        f[2] -= 0.375 * (fa + fb) - 0.75 * fm;
    }
    __syncwarp();
}
// u = integral of d (chebyshev.cpp:637-664, u[0] chosen so that mean(u) = 0), then u[0] -= (u(a) + u(b))/2
__device__ __forceinline__ void integrate_fix(const double* d, double* u, int N, double h2, int lane) {
    for (int n = 1 + lane; n < N; n += 32) {
        double v;
        if (n == 1) v = h2 * (d[0] - d[2] / 2);
        else if (n < N - 1) v = h2 * (d[n - 1] - d[n + 1]) / (2 * n);
        else v = h2 * d[N - 2] / (2 * (N - 1));
        u[n] = v;
    }
    if (lane == 0) u[0] = 0.0;
    __syncwarp();
    double fb, fa, fm;
    wsums(u, N, lane, fb, fa, fm);   // u[0] = 0: fm = -sum_{n even >= 2} u_n/(n^2-1)
    __syncwarp();
    if (lane == 0) {
        const double u0 = 0.0 - fm;                 // u[0] -= u.mean()
        u[0] = u0 - 0.5 * ((fa + u0) + (fb + u0));  // f1.sub(0, (f1(a) + f1(b))/2)
    }
    __syncwarp();
}

// shared memory per warp: f0, f1, f2 as re/im pairs of real profiles: 6 x Ny doubles
__global__ void __launch_bounds__(VP_THREADS) vector2field_kernel(const double* __restrict__ a, double2* __restrict__ u, const PackGeom g, int nmodes) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = blockIdx.x * VP_WARPS + warp;
    if (j >= nmodes) return;
    const int Ny = g.Ny;
    double* f0r = dyn_smem<double>() + (size_t)warp * 6 * Ny;
    double* f0i = f0r + Ny; double* f1r = f0i + Ny; double* f1i = f1r + Ny; double* f2r = f1i + Ny; double* f2i = f2r + Ny;
    int kx, kz;
    long off;
    vp_mode(j, g, kx, kz, off);
    const double h2 = (g.b - g.a) / 2;
    for (int n = lane; n < 6 * Ny; n += 32) f0r[n] = 0.0;
    __syncwarp();
    if (j == 0) {
        for (int ny = 2 + lane; ny < Ny; ny += 32) { f0r[ny] = a[off + ny - 2]; f2r[ny] = a[off + (Ny - 2) + ny - 2]; }
        __syncwarp();
        fix_diri(f0r, Ny, lane);
        fix_diri(f2r, Ny, lane);
        for (int ny = lane; ny < Ny; ny += 32) {
            u[vp_addr(g, 0, 0, ny, 0)] = make_double2(f0r[ny], 0.0);
            u[vp_addr(g, 0, 0, ny, 2)] = make_double2(f2r[ny], 0.0);
        }
        return;
    }
    const double2* a2 = reinterpret_cast<const double2*>(a + off);
    // A = free component (rows 2..Ny-1), B = constrained component (rows 3..Ny-2): on the kz = 0 plane A = w, B = u; else A = u, B = w
    double *Ar, *Ai, *Br, *Bi;
    if (kz == 0) { Ar = f2r; Ai = f2i; Br = f0r; Bi = f0i; }
    else { Ar = f0r; Ai = f0i; Br = f2r; Bi = f2i; }
    for (int ny = 2 + lane; ny < Ny; ny += 32) { const double2 v = a2[ny - 2]; Ar[ny] = v.x; Ai[ny] = v.y; }
    for (int ny = 3 + lane; ny < Ny - 1; ny += 32) { const double2 v = a2[(Ny - 2) + ny - 3]; Br[ny] = v.x; Bi[ny] = v.y; }
    __syncwarp();
    fix_diri(Ar, Ny, lane);
    fix_diri(Ai, Ny, lane);
    if (kx == 0 || kz == 0) {
        // one wavenumber vanishes: the constrained component alone balances dv/dy (flowfield.cpp:4594-4680)
        fix_diri_mean(Br, Ny, lane);
        fix_diri_mean(Bi, Ny, lane);
        integrate_fix(Br, f1r, Ny, h2, lane);
        integrate_fix(Bi, f1i, Ny, h2, lane);
        const double k = kz == 0 ? -(TWO_PI * kx) / g.Lx : -(TWO_PI * kz) / g.Lz;   // v = (0 + i k) * integral
        for (int ny = lane; ny < Ny; ny += 32) {
            const double re = f1r[ny], im = f1i[ny];
            f1r[ny] = 0.0 * re - k * im;
            f1i[ny] = 0.0 * im + k * re;
        }
    } else {
        // general mode (flowfield.cpp:4682-4745): last coefficient of w from that of u, then coefficients 0,1,2 of w so that
        // w(+-1) = 0 and kz/Lz mean(w) + kx/Lx mean(u) = 0
        const double ratio = (kx * g.Lz) / (kz * g.Lx);
        if (lane == 0) { f2r[Ny - 1] = -f0r[Ny - 1] * ratio; f2i[Ny - 1] = -f0i[Ny - 1] * ratio; }
        __syncwarp();
        for (int part = 0; part < 2; ++part) {
            double* f2 = part ? f2i : f2r;
            const double* f0 = part ? f0i : f0r;
            double b2, a2_, m2, b0, a0, m0;
            wsums(f2, Ny, lane, b2, a2_, m2);
            wsums(f0, Ny, lane, b0, a0, m0);
            __syncwarp();
            if (lane == 0) {
                const double fm = m2 + ratio * m0;
                f2[0] -= 0.125 * (a2_ + b2) + 0.75 * fm;
                f2[1] -= 0.5 * (b2 - a2_);
                f2[2] -= 0.375 * (a2_ + b2) - 0.75 * fm;
            }
            __syncwarp();
        }
        // dv/dy = -(i kxx u + i kzz w): g = u * (0 - i kxx) + w * (0 - i kzz), written over the f1 arrays' sources in registers
        const double cx = -TWO_PI * kx / g.Lx, cz = -TWO_PI * kz / g.Lz;
        // g_re = -cx u_im - cz w_im ; g_im = cx u_re + cz w_re     (complex products (0 + i c) * f as the reference forms them)
        double* gr = f1r;  // reuse: integrate needs g in its own array; keep g in f1 and integrate into a scratch pass below
        double* gi = f1i;
        for (int ny = lane; ny < Ny; ny += 32) {
            const double ur = f0r[ny], ui = f0i[ny], wr = f2r[ny], wi = f2i[ny];
            gr[ny] = (0.0 * ur - cx * ui) + (0.0 * wr - cz * wi);
            gi[ny] = (0.0 * ui + cx * ur) + (0.0 * wi + cz * wr);
        }
        __syncwarp();
        // integrate in place is not possible (three-term stencil): stage g in registers, lane-strided
        // (Ny <= 32 * VP_MAXR is checked by the launcher)
        constexpr int MAXR = 24;
        double tr[MAXR], ti[MAXR];
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            const int n = 1 + lane + 32 * r;
            tr[r] = ti[r] = 0.0;
            if (n < Ny) {
                if (n == 1) { tr[r] = h2 * (gr[0] - gr[2] / 2); ti[r] = h2 * (gi[0] - gi[2] / 2); }
                else if (n < Ny - 1) { tr[r] = h2 * (gr[n - 1] - gr[n + 1]) / (2 * n); ti[r] = h2 * (gi[n - 1] - gi[n + 1]) / (2 * n); }
                else { tr[r] = h2 * gr[Ny - 2] / (2 * (Ny - 1)); ti[r] = h2 * gi[Ny - 2] / (2 * (Ny - 1)); }
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            const int n = 1 + lane + 32 * r;
            if (n < Ny) { f1r[n] = tr[r]; f1i[n] = ti[r]; }
        }
        if (lane == 0) { f1r[0] = 0.0; f1i[0] = 0.0; }
        __syncwarp();
        for (int part = 0; part < 2; ++part) {
            double* f1 = part ? f1i : f1r;
            double fb, fa, fm;
            wsums(f1, Ny, lane, fb, fa, fm);
            __syncwarp();
            if (lane == 0) {
                const double u0 = 0.0 - fm;
                f1[0] = u0 - 0.5 * ((fa + u0) + (fb + u0));
            }
            __syncwarp();
        }
    }
    __syncwarp();
    for (int ny = lane; ny < Ny; ny += 32) {
        const double2 v0 = make_double2(f0r[ny], f0i[ny]), v1 = make_double2(f1r[ny], f1i[ny]), v2 = make_double2(f2r[ny], f2i[ny]);
        u[vp_addr(g, kx, kz, ny, 0)] = v0;
        u[vp_addr(g, kx, kz, ny, 1)] = v1;
        u[vp_addr(g, kx, kz, ny, 2)] = v2;
        if (kz == 0) {  // conjugate partner (-kx, 0)
            u[vp_addr(g, -kx, 0, ny, 0)] = make_double2(v0.x, -v0.y);
            u[vp_addr(g, -kx, 0, ny, 1)] = make_double2(v1.x, -v1.y);
            u[vp_addr(g, -kx, 0, ny, 2)] = make_double2(v2.x, -v2.y);
        }
    }
}

// ------------------------------------------------------------------------------------------------ vector algebra
constexpr int VA_THREADS = 256;
constexpr int VA_MAXBLOCKS = 592;  // 4 x 148 SMs

__global__ void __launch_bounds__(VA_THREADS) vec_dot_partial_kernel(const double* __restrict__ x, const double* __restrict__ y, long n, double* __restrict__ partial) {
    __shared__ double red[VA_THREADS / 32];
    double s = 0.0;
    // fixed assignment of elements to threads: the summation order does not depend on scheduling
    for (long i = (long)blockIdx.x * VA_THREADS + threadIdx.x; i < n; i += (long)gridDim.x * VA_THREADS) s += x[i] * y[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < VA_THREADS / 32 ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}
__global__ void __launch_bounds__(VA_THREADS) vec_dot_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    __shared__ double red[VA_THREADS / 32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += VA_THREADS) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < VA_THREADS / 32 ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) *out = v;
    }
}
// y = a*x + b*y  (b = 1: axpy, a = 0: scal, b = 0: scaled copy)
__global__ void __launch_bounds__(VA_THREADS) vec_axpby_kernel(double a, const double* __restrict__ x, double b, double* __restrict__ y, long n) {
    for (long i = (long)blockIdx.x * VA_THREADS + threadIdx.x; i < n; i += (long)gridDim.x * VA_THREADS) {
        double v = b == 0.0 ? 0.0 : b * y[i];
        if (a != 0.0) v += a * x[i];
        y[i] = v;
    }
}
}  // namespace

int field2vector_launch(const double* u, double* a, const PackGeom& g, cudaStream_t st) {
    const int nm = pack_nmodes(g);
    CF_LAUNCH(field2vector_kernel, dim3((nm + VP_WARPS - 1) / VP_WARPS), dim3(VP_THREADS), 0, st, reinterpret_cast<const double2*>(u), a, g, nm);
    CF_KERNEL_CHECK();
    return 0;
}
int vector2field_launch(const double* a, double* u, const PackGeom& g, cudaStream_t st) {
    if (g.Ny > 32 * 24 || g.Ny < 5) { set_last_error("vector2field: Ny out of range (5..768)"); return 1; }
    const int nm = pack_nmodes(g);
    const size_t smem = (size_t)VP_WARPS * 6 * g.Ny * sizeof(double);
    static size_t configured = 0;
    auto kfn = vector2field_kernel;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    CF_LAUNCH(kfn, dim3((nm + VP_WARPS - 1) / VP_WARPS), dim3(VP_THREADS), smem, st, a, reinterpret_cast<double2*>(u), g, nm);
    CF_KERNEL_CHECK();
    return 0;
}
int vec_dot_launch(const double* x, const double* y, long n, double* partial /* >= VA_MAXBLOCKS */, double* out_dev, cudaStream_t st) {
    long nb = (n + VA_THREADS - 1) / VA_THREADS;
    if (nb > VA_MAXBLOCKS) nb = VA_MAXBLOCKS;
    if (nb < 1) nb = 1;
    CF_LAUNCH(vec_dot_partial_kernel, dim3((unsigned)nb), dim3(VA_THREADS), 0, st, x, y, n, partial);
    CF_LAUNCH(vec_dot_final_kernel, dim3(1), dim3(VA_THREADS), 0, st, partial, (int)nb, out_dev);
    CF_KERNEL_CHECK();
    return 0;
}
int vec_axpby_launch(double a, const double* x, double b, double* y, long n, cudaStream_t st) {
    long nb = (n + VA_THREADS - 1) / VA_THREADS;
    if (nb > VA_MAXBLOCKS) nb = VA_MAXBLOCKS;
    if (nb < 1) nb = 1;
    CF_LAUNCH(vec_axpby_kernel, dim3((unsigned)nb), dim3(VA_THREADS), 0, st, a, x, b, y, n);
    CF_KERNEL_CHECK();
    return 0;
}
int vec_partial_capacity() { return VA_MAXBLOCKS; }

}  // namespace cfgpu
