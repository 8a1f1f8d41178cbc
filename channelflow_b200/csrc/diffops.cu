// Whole-field differential operators (reference diffops.cpp: xdiff/ydiff/zdiff :1650-1782, grad :1784-1941, lapl
// :1943-2042, curl :2229-2334, div :2470-2558), pointwise products in the physical state (cross :2560-2611, outer
// :2336-2388, dot :2390-2468, norm/norm2/energy :2613-2700) and the wall-value norm bcNorm2 (:18-48).  Not on the
// time-step hot path (there the operators are fused into the transform passes): diagnostics, initial conditions, tests.
//
// One generic spectral kernel: a thread owns one (output component, kx, kz) column and sweeps n from Ny-1 down to 0 once,
// carrying the Chebyshev derivative recurrences (chebyshev.cpp:672-697; first and second order) of up to three terms in
// registers; x/z derivatives are powers of i k with the odd-order Nyquist rule of flowfield.h:593.  Threads are adjacent
// in kz, so every load and store is a coalesced run.  Roofline: HBM (each input column read once per term, outputs once).
#include "diffops.cuh"

namespace cfgpu {

namespace {
constexpr int DO_THREADS = 256;
constexpr double TWO_PI = 6.283185307179586476925286766559;

__device__ __forceinline__ double2 cmul_ipow(double2 v, double f, int p) {  // v * f * i^p
    switch (p & 3) {
        case 0: return make_double2(f * v.x, f * v.y);
        case 1: return make_double2(-f * v.y, f * v.x);
        case 2: return make_double2(-f * v.x, -f * v.y);
        default: return make_double2(f * v.y, -f * v.x);
    }
}

__global__ void __launch_bounds__(DO_THREADS) diffop_kernel(const double2* __restrict__ in, double2* __restrict__ out, const DiffOpParams p) {
    const FieldGeom& g = p.g;
    const int Mz = g.Nz / 2 + 1;
    const long ncol = (long)p.nout * g.Nx * Mz;
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const int mz = (int)(c % Mz), mx = (int)((c / Mz) % g.Nx), oc = (int)(c / ((long)Mz * g.Nx));
    const long rs = (long)g.Nx * Mz, cs = rs * g.Ny;
    const int kx = mx <= g.Nx / 2 ? mx : mx - g.Nx;
    // this column's terms
    int nt = 0;
    int tin[DIFF_MAXPEROUT], tny[DIFF_MAXPEROUT], tpow[DIFF_MAXPEROUT];
    double tf[DIFF_MAXPEROUT];
    for (int k = 0; k < p.nterms; ++k) {
        if (p.t[k].out != oc || nt >= DIFF_MAXPEROUT) continue;
        const DiffTerm& t = p.t[k];
        // (i 2 pi k/L)^n with the Nyquist mode dropped for odd n
        double f = t.coef;
        const double dx = (kx == g.Nx / 2 && (t.nx & 1)) ? 0.0 : TWO_PI * kx / g.Lx;
        const double dz = (mz == g.Nz / 2 && (t.nz & 1)) ? 0.0 : TWO_PI * mz / g.Lz;
        for (int q = 0; q < t.nx; ++q) f *= dx;
        for (int q = 0; q < t.nz; ++q) f *= dz;
        tin[nt] = t.in; tny[nt] = t.ny; tpow[nt] = t.nx + t.nz; tf[nt] = f;
        ++nt;
    }
    const int Nb = g.Ny - 1;
    const double scale = 4.0 / (g.b - g.a);
    const long col = (long)mx * Mz + mz;
    double2 r1[DIFF_MAXPEROUT][2], r2[DIFF_MAXPEROUT][2], up1[DIFF_MAXPEROUT], upd[DIFF_MAXPEROUT];
#pragma unroll
    for (int k = 0; k < DIFF_MAXPEROUT; ++k) {
        r1[k][0] = r1[k][1] = r2[k][0] = r2[k][1] = up1[k] = upd[k] = make_double2(0.0, 0.0);
    }
    double2* op = out + oc * cs + col;
    for (int n = Nb; n >= 0; --n) {
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < DIFF_MAXPEROUT; ++k) {
            if (k >= nt) break;
            const double2 v = in[tin[k] * cs + n * rs + col];
            double2 val = v;
            if (tny[k] >= 1) {
                // d1[n] = d1[n+2] + scale (n+1) u[n+1], halved at n = 0
                double2& a = r1[k][n & 1];
                if (n + 1 <= Nb) {
                    const double fct = scale * (n + 1);
                    a.x += fct * up1[k].x;
                    a.y += fct * up1[k].y;
                }
                const double2 d1 = n == 0 ? make_double2(0.5 * a.x, 0.5 * a.y) : a;
                val = d1;
                if (tny[k] == 2) {
                    double2& b = r2[k][n & 1];
                    if (n + 1 <= Nb) {
                        const double fct = scale * (n + 1);
                        b.x += fct * upd[k].x;
                        b.y += fct * upd[k].y;
                    }
                    val = n == 0 ? make_double2(0.5 * b.x, 0.5 * b.y) : b;
                    upd[k] = d1;
                }
                up1[k] = v;
            }
            const double2 w = cmul_ipow(val, tf[k], tpow[k]);
            acc.x += w.x;
            acc.y += w.y;
        }
        op[n * rs] = acc;
    }
}

__global__ void __launch_bounds__(DO_THREADS) pointwise_kernel(int op, const double* __restrict__ f, const double* __restrict__ g,
                                                               double* __restrict__ out, int fd, int gd, long n) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        switch (op) {
            case PW_CROSS: {
                const double a0 = f[p], a1 = f[n + p], a2 = f[2 * n + p], b0 = g[p], b1 = g[n + p], b2 = g[2 * n + p];
                out[p] = a1 * b2 - a2 * b1;
                out[n + p] = a2 * b0 - a0 * b2;
                out[2 * n + p] = a0 * b1 - a1 * b0;
                break;
            }
            case PW_OUTER:
                for (int i = 0; i < fd; ++i)
                    for (int j = 0; j < gd; ++j) out[(long)(i * gd + j) * n + p] = f[i * n + p] * g[j * n + p];
                break;
            case PW_DOT: {
                double s = 0.0;
                for (int i = 0; i < fd; ++i) s += f[i * n + p] * g[i * n + p];
                out[p] = s;
                break;
            }
            case PW_NORM2:
            case PW_NORM:
            case PW_ENERGY: {
                double s = 0.0;
                for (int i = 0; i < fd; ++i) s += (op == PW_ENERGY ? 0.5 : 1.0) * f[i * n + p] * f[i * n + p];
                out[p] = op == PW_NORM ? sqrt(s) : s;
                break;
            }
            default:  // PW_MUL: componentwise product
                for (int i = 0; i < fd; ++i) out[i * n + p] = f[i * n + p] * g[i * n + p];
        }
    }
}

// one thread per (component, mx, mz) column
__global__ void __launch_bounds__(DO_THREADS) bcnorm2_kernel(const double2* __restrict__ u, const double2* __restrict__ v, int Nx, int Ny, int Mz,
                                                             int Nd, int yspectral, double* __restrict__ partial) {
    __shared__ double red[DO_THREADS / 32];
    const long ncol = (long)Nd * Nx * Mz;
    const long rs = (long)Nx * Mz, cs = rs * Ny;
    double s = 0.0;
    for (long c = (long)blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += (long)gridDim.x * blockDim.x) {
        const long col = c % rs, i = c / rs;
        const double2* up = u + i * cs + col;
        const double2* vp = v ? v + i * cs + col : nullptr;
        double2 a = make_double2(0.0, 0.0), b = a;
        if (yspectral) {
            for (int n = Ny - 1; n >= 0; --n) {
                double2 w = up[n * rs];
                if (vp) { const double2 q = vp[n * rs]; w.x -= q.x; w.y -= q.y; }
                b.x += w.x; b.y += w.y;
                if (n & 1) { a.x -= w.x; a.y -= w.y; } else { a.x += w.x; a.y += w.y; }
            }
        } else {
            b = up[0]; a = up[(Ny - 1) * rs];
            if (vp) { b.x -= vp[0].x; b.y -= vp[0].y; a.x -= vp[(Ny - 1) * rs].x; a.y -= vp[(Ny - 1) * rs].y; }
        }
        s += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < DO_THREADS / 32 ? red[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) sum_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) *out = v;
    }
}
}  // namespace

int diffop_launch(const double* in, double* out, const DiffOpParams& p, cudaStream_t st) {
    if (p.nterms < 1 || p.nterms > DIFF_MAXTERMS) { set_last_error("diffop: 1..27 terms"); return 1; }
    int per[64] = {0};
    for (int k = 0; k < p.nterms; ++k) {
        const DiffTerm& t = p.t[k];
        if (t.out < 0 || t.out >= p.nout || t.out >= 64 || t.ny < 0 || t.ny > 2 || t.nx < 0 || t.nz < 0 || ++per[t.out] > DIFF_MAXPEROUT) {
            set_last_error("diffop: bad term (at most 3 terms per output component, y order <= 2)");
            return 1;
        }
    }
    const long ncol = (long)p.nout * p.g.Nx * (p.g.Nz / 2 + 1);
    CF_LAUNCH(diffop_kernel, dim3((unsigned)((ncol + DO_THREADS - 1) / DO_THREADS)), dim3(DO_THREADS), 0, st,
              reinterpret_cast<const double2*>(in), reinterpret_cast<double2*>(out), p);
    CF_KERNEL_CHECK();
    return 0;
}
int pointwise_launch(int op, const double* f, const double* g, double* out, int fd, int gd, long n, cudaStream_t st) {
    long nb = (n + DO_THREADS - 1) / DO_THREADS;
    if (nb > 148L * 16) nb = 148L * 16;
    if (nb < 1) nb = 1;
    CF_LAUNCH(pointwise_kernel, dim3((unsigned)nb), dim3(DO_THREADS), 0, st, op, f, g, out, fd, gd, n);
    CF_KERNEL_CHECK();
    return 0;
}
int bcnorm2_launch(const double* u, const double* v, int Nx, int Ny, int Nz, int Nd, int yspectral, double* partial, size_t cap,
                   double* out_dev, cudaStream_t st) {
    const int Mz = Nz / 2 + 1;
    const long ncol = (long)Nd * Nx * Mz;
    long nb = (ncol + DO_THREADS - 1) / DO_THREADS;
    if (nb > 592) nb = 592;
    if ((size_t)nb > cap) { set_last_error("bcnorm2: partial buffer too small"); return 1; }
    CF_LAUNCH(bcnorm2_kernel, dim3((unsigned)nb), dim3(DO_THREADS), 0, st, reinterpret_cast<const double2*>(u),
              reinterpret_cast<const double2*>(v), Nx, Ny, Mz, Nd, yspectral, partial);
    CF_LAUNCH(sum_kernel, dim3(1), dim3(256), 0, st, (const double*)partial, (int)nb, out_dev);
    CF_KERNEL_CHECK();
    return 0;
}

}  // namespace cfgpu
