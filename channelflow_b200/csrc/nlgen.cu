// See nlgen.cuh.  All kernels are HBM-streaming; spectral-y derivatives are the reference's recurrence
// (chebyshev.cpp:672-697) walked by one thread per (component, kx, kz) column, threads adjacent in kz.
#include "nlgen.cuh"

namespace cfgpu {

namespace {
constexpr int NL_THREADS = 256;
constexpr double TWO_PI = 6.283185307179586476925286766559;

__device__ __forceinline__ void wavenumbers(int mx, int mz, const FieldGeom& g, double& dx, double& dz) {
    const int kx = mx <= g.Nx / 2 ? mx : mx - g.Nx;
    // zero_last_mode (flowfield.h:593): odd derivatives of the Nyquist mode vanish
    dx = (kx == g.Nx / 2) ? 0.0 : TWO_PI * kx / g.Lx;
    dz = (mz == g.Nz / 2) ? 0.0 : TWO_PI * mz / g.Lz;
}

// grid covers 3 * Nx * Mz columns
__global__ void __launch_bounds__(NL_THREADS) grad3_kernel(const double2* __restrict__ u, double2* __restrict__ G, const FieldGeom g) {
    const int Mz = g.Nz / 2 + 1;
    const long ncol = 3L * g.Nx * Mz;
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const int mz = (int)(c % Mz), mx = (int)((c / Mz) % g.Nx), i = (int)(c / ((long)Mz * g.Nx));
    const long rs = (long)g.Nx * Mz, cs = rs * g.Ny;
    double dx, dz;
    wavenumbers(mx, mz, g, dx, dz);
    const double2* up = u + i * cs + (long)mx * Mz + mz;
    double2* gx = G + (3 * i) * cs + (long)mx * Mz + mz;
    double2* gy = gx + cs;
    double2* gz = gy + cs;
    const int Nb = g.Ny - 1;
    const double scale = 4.0 / (g.b - g.a);
    double2 run[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};  // d[n+2] of each parity
    double2 above = make_double2(0.0, 0.0);                              // u[n+1]
    for (int n = Nb; n >= 0; --n) {
        const double2 v = up[n * rs];
        gx[n * rs] = make_double2(-dx * v.y, dx * v.x);
        gz[n * rs] = make_double2(-dz * v.y, dz * v.x);
        double2& r = run[n & 1];
        if (n + 1 <= Nb) {
            const double f = scale * (n + 1);
            r.x = r.x + f * above.x;
            r.y = r.y + f * above.y;
        }
        gy[n * rs] = n == 0 ? make_double2(0.5 * r.x, 0.5 * r.y) : r;
        above = v;
    }
}

__global__ void __launch_bounds__(NL_THREADS) pointwise_nl_kernel(const double* __restrict__ u, double* __restrict__ G, double* __restrict__ f,
                                                                  double conv_coef, int do_outer, double rot, long n) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        double v[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] = u[i * n + p];
        if (conv_coef != 0.0) {
            double fi[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s += conv_coef * v[j] * G[(3 * i + j) * n + p];
                fi[i] = s;
            }
            if (rot != 0.0) {
                fi[0] -= rot * v[1];
                fi[1] += rot * v[0];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) f[i * n + p] = fi[i];
        }
        if (do_outer) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) G[(3 * i + j) * n + p] = v[i] * v[j];
        }
    }
}

__global__ void __launch_bounds__(NL_THREADS) div9_kernel(const double2* __restrict__ T, double2* __restrict__ f, double coef, int accumulate,
                                                          const FieldGeom g) {
    const int Mz = g.Nz / 2 + 1;
    const long ncol = 3L * g.Nx * Mz;
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const int mz = (int)(c % Mz), mx = (int)((c / Mz) % g.Nx), i = (int)(c / ((long)Mz * g.Nx));
    const long rs = (long)g.Nx * Mz, cs = rs * g.Ny;
    double dx, dz;
    wavenumbers(mx, mz, g, dx, dz);
    const double2* t0 = T + (3 * i) * cs + (long)mx * Mz + mz;
    const double2* t1 = t0 + cs;
    const double2* t2 = t1 + cs;
    double2* fp = f + i * cs + (long)mx * Mz + mz;
    const int Nb = g.Ny - 1;
    const double scale = 4.0 / (g.b - g.a);
    double2 run[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
    double2 above = make_double2(0.0, 0.0);
    for (int n = Nb; n >= 0; --n) {
        const double2 a0 = t0[n * rs], a1 = t1[n * rs], a2 = t2[n * rs];
        double2& r = run[n & 1];
        if (n + 1 <= Nb) {
            const double fct = scale * (n + 1);
            r.x = r.x + fct * above.x;
            r.y = r.y + fct * above.y;
        }
        const double2 dy = n == 0 ? make_double2(0.5 * r.x, 0.5 * r.y) : r;
        double2 out = accumulate ? fp[n * rs] : make_double2(0.0, 0.0);
        out.x += coef * (-dx * a0.y - dz * a2.y);
        out.y += coef * (dx * a0.x + dz * a2.x);
        out.x += coef * dy.x;
        out.y += coef * dy.y;
        fp[n * rs] = out;
        above = a1;
    }
}

// one thread per complex element of one component plane set; f = (U d/dx + W d/dz) u + v (U' e_x + W' e_z)
__global__ void __launch_bounds__(NL_THREADS) linearized_kernel(const double2* __restrict__ u, double2* __restrict__ f, const double* __restrict__ prof,
                                                                const FieldGeom g) {
    const int Mz = g.Nz / 2 + 1;
    const long rs = (long)g.Nx * Mz, cs = rs * g.Ny;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < cs; p += stride) {
        const int mz = (int)(p % Mz), mx = (int)((p / Mz) % g.Nx), ny = (int)(p / rs);
        double dx, dz;
        wavenumbers(mx, mz, g, dx, dz);
        const double U = prof[ny], Uy = prof[g.Ny + ny], W = prof[2 * g.Ny + ny], Wy = prof[3 * g.Ny + ny];
        const double k = U * dx + W * dz;  // Uddx_Wddz = i k
        const double2 cu = u[p], cv = u[cs + p], cw = u[2 * cs + p];
        f[p] = make_double2(-k * cu.y + cv.x * Uy, k * cu.x + cv.y * Uy);
        f[cs + p] = make_double2(-k * cv.y, k * cv.x);
        f[2 * cs + p] = make_double2(-k * cw.y + cv.x * Wy, k * cw.x + cv.y * Wy);
    }
}

__global__ void add_base00_kernel(double2* __restrict__ u, const double* __restrict__ Ubase, const double* __restrict__ Wbase, double Vsuck,
                                  double s, const FieldGeom g) {
    const int Mz = g.Nz / 2 + 1;
    const long rs = (long)g.Nx * Mz, cs = rs * g.Ny;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= g.Ny) return;
    if (Ubase) u[n * rs].x += s * Ubase[n];
    if (Wbase) u[2 * cs + n * rs].x += s * Wbase[n];
    if (n == 0) u[cs].x -= s * Vsuck;
}

__global__ void __launch_bounds__(NL_THREADS) coriolis_kernel(const double* __restrict__ u, double* __restrict__ f, double rot, long n,
                                                              int Nz, int Nzpad) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        if ((int)(p % Nzpad) >= Nz) continue;  // the reference loops over nz < Nz only
        f[p] -= rot * u[n + p];
        f[n + p] += rot * u[p];
    }
}

inline int ew_grid(long n) {
    long b = (n + NL_THREADS - 1) / NL_THREADS;
    return (int)(b > 148L * 16 ? 148L * 16 : (b < 1 ? 1 : b));
}
}  // namespace

int grad3_launch(const double* u, double* G, const FieldGeom& g, cudaStream_t st) {
    const long ncol = 3L * g.Nx * (g.Nz / 2 + 1);
    CF_LAUNCH(grad3_kernel, dim3((unsigned)((ncol + NL_THREADS - 1) / NL_THREADS)), dim3(NL_THREADS), 0, st,
              reinterpret_cast<const double2*>(u), reinterpret_cast<double2*>(G), g);
    CF_KERNEL_CHECK();
    return 0;
}
int pointwise_nl_launch(const double* u, double* G, double* f, double conv_coef, int do_outer, double rot, long n, cudaStream_t st) {
    CF_LAUNCH(pointwise_nl_kernel, dim3(ew_grid(n)), dim3(NL_THREADS), 0, st, u, G, f, conv_coef, do_outer, rot, n);
    CF_KERNEL_CHECK();
    return 0;
}
int div9_launch(const double* T, double* f, double coef, int accumulate, const FieldGeom& g, cudaStream_t st) {
    const long ncol = 3L * g.Nx * (g.Nz / 2 + 1);
    CF_LAUNCH(div9_kernel, dim3((unsigned)((ncol + NL_THREADS - 1) / NL_THREADS)), dim3(NL_THREADS), 0, st,
              reinterpret_cast<const double2*>(T), reinterpret_cast<double2*>(f), coef, accumulate, g);
    CF_KERNEL_CHECK();
    return 0;
}
int linearized_launch(const double* u, double* f, const double* prof, const FieldGeom& g, cudaStream_t st) {
    const long n = (long)g.Ny * g.Nx * (g.Nz / 2 + 1);
    CF_LAUNCH(linearized_kernel, dim3(ew_grid(n)), dim3(NL_THREADS), 0, st, reinterpret_cast<const double2*>(u),
              reinterpret_cast<double2*>(f), prof, g);
    CF_KERNEL_CHECK();
    return 0;
}
int add_base00_launch(double* u, const double* Ubase, const double* Wbase, double Vsuck, double s, const FieldGeom& g, cudaStream_t st) {
    CF_LAUNCH(add_base00_kernel, dim3((g.Ny + 127) / 128), dim3(128), 0, st, reinterpret_cast<double2*>(u), Ubase, Wbase, Vsuck, s, g);
    CF_KERNEL_CHECK();
    return 0;
}
int coriolis_launch(const double* u, double* f, double rot, long n, int Nz, cudaStream_t st) {
    CF_LAUNCH(coriolis_kernel, dim3(ew_grid(n)), dim3(NL_THREADS), 0, st, u, f, rot, n, Nz, 2 * (Nz / 2 + 1));
    CF_KERNEL_CHECK();
    return 0;
}

}  // namespace cfgpu
