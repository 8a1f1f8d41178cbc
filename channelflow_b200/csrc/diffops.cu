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
// PressureSolver::solve, step II (poissonsolver.cpp:352-431): the homogeneous correction that turns the Dirichlet solution of
// lapl p = -div N(u) into the one with dp/dy = nu d2v/dy2 at both walls.  Per Fourier mode (not the mean mode):
//   alpha = nu v''(a) - p'(a),  beta = nu v''(b) - p'(b),  mu = sqrt(lambda),  H = b - a,
//   g(y) = c exp(-mu (y-a)) + d exp(mu (y-b)),  c = (-alpha + beta e^{-mu H}) / delta,  d = (beta - alpha e^{-mu H}) / delta,
//   delta = mu (1 - e^{-2 mu H}).
// Wall derivatives from the Chebyshev coefficients in closed form: T_n'(+-1) = (+-1)^{n+1} n^2, T_n''(+-1) = (+-1)^n n^2 (n^2-1)/3.
// One thread per (kx, kz) column, adjacent threads adjacent in kz: coalesced.  g is written at the Gauss-Lobatto points
// (y-physical, xz-spectral); the caller transforms it with the y-GEMM and adds it to p.
__global__ void __launch_bounds__(DO_THREADS) pressure_neumann_kernel(const double2* __restrict__ p, const double2* __restrict__ v, double nu,
                                                                      FieldGeom g, double2* __restrict__ out) {
    const int Mz = g.Nz / 2 + 1, N = g.Ny;
    const long ncol = (long)g.Nx * Mz;
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const int mz = (int)(c % Mz), mx = (int)(c / Mz);
    const int kx = mx <= g.Nx / 2 ? mx : mx - g.Nx;
    const long rs = ncol;
    if (mx == 0 && mz == 0) {
        for (int j = 0; j < N; ++j) out[j * rs + c] = make_double2(0.0, 0.0);
        return;
    }
    const double H = g.b - g.a, s1 = 2.0 / H, s2 = s1 * s1;
    double2 pa = {0, 0}, pb = {0, 0}, va = {0, 0}, vb = {0, 0};
    for (int n = N - 1; n >= 1; --n) {
        const double n2 = (double)n * n, w1 = s1 * n2, w2 = s2 * n2 * (n2 - 1.0) / 3.0;
        const double sg = (n & 1) ? -1.0 : 1.0;  // (-1)^n
        const double2 pc = p[n * rs + c], vc = v[n * rs + c];
        pb.x += w1 * pc.x; pb.y += w1 * pc.y;
        pa.x -= sg * w1 * pc.x; pa.y -= sg * w1 * pc.y;
        vb.x += w2 * vc.x; vb.y += w2 * vc.y;
        va.x += sg * w2 * vc.x; va.y += sg * w2 * vc.y;
    }
    const double2 alpha = make_double2(nu * va.x - pa.x, nu * va.y - pa.y), beta = make_double2(nu * vb.x - pb.x, nu * vb.y - pb.y);
    const double lambda = TWO_PI * TWO_PI * ((kx / g.Lx) * (kx / g.Lx) + (mz / g.Lz) * (mz / g.Lz));
    const double mu = sqrt(lambda), em = exp(-mu * H), delta = mu * (1.0 - exp(-2.0 * mu * H));
    const double2 cc = make_double2((-alpha.x + beta.x * em) / delta, (-alpha.y + beta.y * em) / delta);
    const double2 dd = make_double2((beta.x - alpha.x * em) / delta, (beta.y - alpha.y * em) / delta);
    const double PI_ = 0.5 * TWO_PI;
    for (int j = 0; j < N; ++j) {
        const double y = 0.5 * ((g.b + g.a) + (g.b - g.a) * cos(PI_ * j / (N - 1)));
        const double ea = exp(-mu * (y - g.a)), eb = exp(mu * (y - g.b));
        out[j * rs + c] = make_double2(cc.x * ea + dd.x * eb, cc.y * ea + dd.y * eb);
    }
}
int pressure_neumann_launch(const double* p, const double* v, double nu, const FieldGeom& g, double* out, cudaStream_t st) {
    const long ncol = (long)g.Nx * (g.Nz / 2 + 1);
    CF_LAUNCH(pressure_neumann_kernel, dim3((unsigned)((ncol + DO_THREADS - 1) / DO_THREADS)), dim3(DO_THREADS), 0, st,
              reinterpret_cast<const double2*>(p), reinterpret_cast<const double2*>(v), nu, g, reinterpret_cast<double2*>(out));
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
