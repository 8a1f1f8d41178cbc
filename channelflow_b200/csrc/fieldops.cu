// Elementwise FlowField operations and L2 norms / inner products (all HBM-streaming kernels).
// Replaces FlowField::add / operator*= / setToZero / zeroPaddedModes (flowfield.h:596-615, flowfield.cpp:1459-1469,
// 2229-2255) and L2Norm2 / L2Dist2 / L2InnerProduct (diffops.cpp:353-541 -> chebyshev.cpp:758-802).
#include "fieldops.cuh"

namespace cfgpu {

namespace {
constexpr int EW_THREADS = 256;

__global__ void __launch_bounds__(EW_THREADS) axpby_kernel(double* __restrict__ y, double a, const double* __restrict__ x,
                                                           double b, const double* __restrict__ z, long n2) {
    // n2 = number of double2 elements
    const long stride = (long)gridDim.x * blockDim.x;
    double2* y2 = reinterpret_cast<double2*>(y);
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const double2* z2 = reinterpret_cast<const double2*>(z);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 yy = y2[i];
        const double2 xx = x2[i];
        if (z) {
            const double2 zz = z2[i];
            yy.x += a * xx.x + b * zz.x;
            yy.y += a * xx.y + b * zz.y;
        } else {
            yy.x += a * xx.x;
            yy.y += a * xx.y;
        }
        y2[i] = yy;
    }
}

__global__ void __launch_bounds__(EW_THREADS) scale_kernel(double* __restrict__ y, double s, long n2) {
    const long stride = (long)gridDim.x * blockDim.x;
    double2* y2 = reinterpret_cast<double2*>(y);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 v = y2[i];
        v.x *= s;
        v.y *= s;
        y2[i] = v;
    }
}

// zero every (mx,mz) with |kx| > Kx or kz > Kz, all y, all components. One thread per complex element.
__global__ void __launch_bounds__(EW_THREADS) zero_padded_kernel(double2* __restrict__ c, int Nx, int Mz, int Kx, int Kz, long ncplx) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < ncplx; i += stride) {
        const int mz = (int)(i % Mz);
        const int mx = (int)((i / Mz) % Nx);
        const int kx = mx <= Nx / 2 ? mx : mx - Nx;
        const int akx = kx < 0 ? -kx : kx;
        if (akx > Kx || mz > Kz) c[i] = make_double2(0.0, 0.0);
    }
}

// FlowField *= FieldSymmetry (flowfield.cpp:1274-1433):  (u,v,w)(x,y,z) -> s (sx u, sy v, sz w)(sx x + ax Lx, sy y, sz z + az Lz)
// on a spectral field, in place.  In Fourier-Chebyshev coefficients (kz >= 0 stored, u(-kx,-kz) = conj u(kx,kz)):
//   sx sz = +1 :  u(kx,kz) <- c(kx)  [u or conj u](kx,kz)                      (conj when sx = sz = -1)
//   sx sz = -1 :  u(kx,kz) <- c(kx)  [u or conj u](-kx,kz), both rows of a +-kx pair exchanged by one thread
//   c(kx) = s s_i (sy)^n exp(i 2 pi (ax sx kx + az sz kz)),  s_i = the sign of component i (sx, sy, sz for a vector)
// One thread per (component, kx >= 0 pair, kz) sweeping n; adjacent threads adjacent in kz.  Modes outside [Kxlo..Kxhi] x
// [0..Kz] (the de-aliased box of a padded field, else every mode) are left alone, and so is the unpaired kx = Nx/2 row of
// the exchanging case, as in the reference.
__global__ void __launch_bounds__(EW_THREADS) symmetry_kernel(double2* __restrict__ c, int Nx, int Ny, int Mz, int Nd, int Kxlo, int Kxhi,
                                                              int Kz, int s, int sx, int sy, int sz, double ax, double az) {
    const int nkz = Kz + 1, npos = Kxhi + 1;  // pairs kx = 0 .. Kxhi (kx = 0 pairs with itself)
    const long items = (long)Nd * npos * nkz;
    const long it = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= items) return;
    const int kz = (int)(it % nkz), kx = (int)((it / nkz) % npos), i = (int)(it / ((long)nkz * npos));
    const int si = Nd == 3 ? (i == 0 ? sx : i == 1 ? sy : sz) : 1;  // sign_i (flowfield.cpp:1066-1093), vectors and scalars
    const double TWO_PI_ = 6.283185307179586476925286766559;
    const long rs = (long)Nx * Mz, cs = rs * Ny;
    const bool exchange = sx + sz == 0, conj_ = exchange ? sx == 1 : sx == -1;
    const int kxm = -kx;
    const bool have_m = kx > 0 && kxm >= Kxlo;  // the -kx row exists in the range
    if (exchange && kx > 0 && !have_m) return;  // unpaired Nyquist row
    double2 cp, cm;
    sincos(TWO_PI_ * (ax * sx * kx + az * sz * kz), &cp.y, &cp.x);
    sincos(TWO_PI_ * (ax * sx * kxm + az * sz * kz), &cm.y, &cm.x);
    double2* colp = c + i * cs + (long)kx * Mz + kz;
    double2* colm = c + i * cs + (long)(kxm < 0 ? Nx + kxm : kxm) * Mz + kz;
    for (int n = 0; n < Ny; ++n) {
        const double sg = (double)(s * si * ((sy == -1 && (n & 1)) ? -1 : 1));
        double2 up = colp[n * rs], um = make_double2(0.0, 0.0);
        if (have_m) um = colm[n * rs];
        if (conj_) { up.y = -up.y; um.y = -um.y; }
        const double2 srcp = exchange ? (kx == 0 ? up : um) : up, srcm = exchange ? up : um;
        colp[n * rs] = make_double2(sg * (cp.x * srcp.x - cp.y * srcp.y), sg * (cp.x * srcp.y + cp.y * srcp.x));
        if (have_m) colm[n * rs] = make_double2(sg * (cm.x * srcm.x - cm.y * srcm.y), sg * (cm.x * srcm.y + cm.y * srcm.x));
    }
}

// out[n] = (re, im) of mode (mx,mz) component i  /  add
__global__ void profile_get_kernel(const double2* __restrict__ c, long off0, long rs, int Ny, double2* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < Ny) out[n] = c[off0 + n * rs];
}
__global__ void profile_add_kernel(double2* __restrict__ c, long off0, long rs, int Ny, const double2* __restrict__ in, double s) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < Ny) {
        double2 v = c[off0 + n * rs];
        v.x += s * in[n].x;
        v.y += s * in[n].y;
        c[off0 + n * rs] = v;
    }
}

// ---- L2 quadratic form: sum_q cz(q) sum_i sum_{parts} sum_m x[m] sum_{n = m mod 2} y[n] W[m][n]
// x = u - alpha*v, y = (ip ? v : x).  One CTA per tile of TMn modes and component; profiles staged in smem [n][t].
constexpr int NRM_THREADS = 256;
__global__ void __launch_bounds__(NRM_THREADS) l2form_kernel(const double* __restrict__ u, const double* __restrict__ v, int mode /*0 norm,1 dist,2 ip*/,
                                                             const double* __restrict__ W, int N, int Nx, int Mz, int Kx, int Kz, int fullbox,
                                                             int qlo, int nq, int TMn, long rs, long cs, double czw, double* __restrict__ partial) {
    const int TT = 2 * TMn;
    double* X = dyn_smem<double>();
    double* Y = (mode == 2) ? X + (size_t)N * TT : X;
    __shared__ double red[NRM_THREADS / 32];
    const int tid = threadIdx.x;
    const int comp = blockIdx.y;
    const int q0 = blockIdx.x * TMn;
    const int nkz = fullbox ? Mz : Kz + 1;
    for (int idx = tid; idx < N * TT; idx += NRM_THREADS) {
        const int n = idx / TT, t = idx - n * TT, m = t >> 1;
        const int q = q0 + m;
        double xv = 0.0, yv = 0.0;
        if (q < nq) {
            const int mxi = (qlo + q) / nkz, kz = (qlo + q) - mxi * nkz;
            int mx = mxi;
            if (!fullbox) {
                const int kx = mxi <= Kx ? mxi : mxi - (2 * Kx + 1);
                mx = kx >= 0 ? kx : Nx + kx;
            }
            const long go = comp * cs + n * rs + 2L * (kz + (long)Mz * mx) + (t & 1);
            const double a = u[go];
            if (mode == 0) { xv = a; }
            else if (mode == 1) { xv = a - v[go]; }
            else { xv = a; yv = v[go]; }
        }
        X[idx] = xv;
        if (mode == 2) Y[idx] = yv;
    }
    __syncthreads();
    const int t = tid % TT, ms = tid / TT, nms = NRM_THREADS / TT;
    double sum = 0.0;
    if (ms < nms) {
        for (int m = N - 1 - ms; m >= 0; m -= nms) {
            double psum = 0.0;
            const double* Wm = W + (size_t)m * N;
            for (int n = m % 2; n < N; n += 2) psum += Y[n * TT + t] * __ldg(&Wm[n]);
            sum += X[m * TT + t] * psum;
        }
        const int q = q0 + (t >> 1);
        if (q < nq) {
            const int kz = (qlo + q) % nkz;
            if (kz > 0) sum *= czw;  // 2: the kz < 0 ghost modes (diffops.cpp:417-487); 1: plain sum over stored modes (divNorm2)
        } else sum = 0.0;
    }
    sum = warp_sum(sum);
    if ((tid & 31) == 0) red[tid >> 5] = sum;
    __syncthreads();
    if (tid < 32) {
        double s = tid < NRM_THREADS / 32 ? red[tid] : 0.0;
        s = warp_sum(s);
        if (tid == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ partial, int n, double scale, double* __restrict__ out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) *out = v * scale;
    }
}

int grid_for(long n) {
    long g = (n + EW_THREADS - 1) / EW_THREADS;
    const long cap = 148L * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}
}  // namespace

int axpby_launch(double* y, double a, const double* x, double b, const double* z, long n, cudaStream_t st) {
    const long n2 = n / 2;  // field sizes are always even (Nzpad even)
    CF_LAUNCH(axpby_kernel, dim3(grid_for(n2)), dim3(EW_THREADS), 0, st, y, a, x, b, z, n2);
    CF_KERNEL_CHECK();
    return 0;
}
int symmetry_launch(double* d, int Nx, int Ny, int Nz, int Nd, int Kxlo, int Kxhi, int Kz, int s, int sx, int sy, int sz, double ax, double az,
                    cudaStream_t st) {
    const long items = (long)Nd * (Kxhi + 1) * (Kz + 1);
    CF_LAUNCH(symmetry_kernel, dim3((unsigned)((items + EW_THREADS - 1) / EW_THREADS)), dim3(EW_THREADS), 0, st, reinterpret_cast<double2*>(d),
              Nx, Ny, Nz / 2 + 1, Nd, Kxlo, Kxhi, Kz, s, sx, sy, sz, ax, az);
    CF_KERNEL_CHECK();
    return 0;
}
int scale_launch(double* y, double s, long n, cudaStream_t st) {
    const long n2 = n / 2;
    CF_LAUNCH(scale_kernel, dim3(grid_for(n2)), dim3(EW_THREADS), 0, st, y, s, n2);
    CF_KERNEL_CHECK();
    return 0;
}
int zero_padded_launch(double* d, int Nx, int Ny, int Nz, int Nd, int Kx, int Kz, cudaStream_t st) {
    const int Mz = Nz / 2 + 1;
    const long ncplx = (long)Nx * Mz * Ny * Nd;
    CF_LAUNCH(zero_padded_kernel, dim3(grid_for(ncplx)), dim3(EW_THREADS), 0, st, reinterpret_cast<double2*>(d), Nx, Mz, Kx, Kz, ncplx);
    CF_KERNEL_CHECK();
    return 0;
}
int profile_get_launch(const double* d, long off0_cplx, long rs_cplx, int Ny, double* out_dev, cudaStream_t st) {
    CF_LAUNCH(profile_get_kernel, dim3((Ny + 127) / 128), dim3(128), 0, st, reinterpret_cast<const double2*>(d), off0_cplx, rs_cplx, Ny,
              reinterpret_cast<double2*>(out_dev));
    CF_KERNEL_CHECK();
    return 0;
}
int profile_add_launch(double* d, long off0_cplx, long rs_cplx, int Ny, const double* in_dev, double s, cudaStream_t st) {
    CF_LAUNCH(profile_add_kernel, dim3((Ny + 127) / 128), dim3(128), 0, st, reinterpret_cast<double2*>(d), off0_cplx, rs_cplx, Ny,
              reinterpret_cast<const double2*>(in_dev), s);
    CF_KERNEL_CHECK();
    return 0;
}

__global__ void __launch_bounds__(EW_THREADS) rows_pack_kernel(double2* __restrict__ field, double2* __restrict__ buf, int Nx, int Mz,
                                                               int Kx, int nkz, int x0, int nloc, long total, int dir) {
    const long stride = (long)gridDim.x * blockDim.x;
    const int nmx = 2 * Kx + 1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int kz = (int)(i % nkz);
        const int xl = (int)((i / nkz) % nloc);
        const long row = i / ((long)nkz * nloc);
        const int mxi = x0 + xl;
        const int kx = mxi <= Kx ? mxi : mxi - nmx;
        const int mx = kx >= 0 ? kx : Nx + kx;
        const long fo = (row * Nx + mx) * Mz + kz;
        if (dir == 0) buf[i] = field[fo];
        else field[fo] = buf[i];
    }
}

int rows_pack_launch(double* field, double* buf, int Nx, int Nz, int nrows, int Kx, int Kz, int x0, int x1, int dir, cudaStream_t st) {
    const long total = (long)nrows * (x1 - x0) * (Kz + 1);
    if (total <= 0) return 0;
    CF_LAUNCH(rows_pack_kernel, dim3(grid_for(total)), dim3(EW_THREADS), 0, st, reinterpret_cast<double2*>(field),
              reinterpret_cast<double2*>(buf), Nx, Nz / 2 + 1, Kx, Kz + 1, x0, x1 - x0, total, dir);
    CF_KERNEL_CHECK();
    return 0;
}

// serial [i][ny][mx][mz] <-> tile-major [tile][i][ny][TM] (complex), retained modes q = (mxi - x0)*(Kz+1) + kz of this
// rank in natural order, TM per tile.  dir 0: serial -> tile (slots past nq are zero filled), dir 1: tile -> serial.
__global__ void __launch_bounds__(EW_THREADS) tile_convert_kernel(double2* __restrict__ ser, double2* __restrict__ tile, int Nx, int Ny,
                                                                  int Mz, int Nd, int Kx, int nkz, int x0, int nq, int TM, int dir) {
    const int t = blockIdx.x, comp = blockIdx.y;
    const int nmx = 2 * Kx + 1;
    const long rs = (long)Nx * Mz, cs = rs * Ny;
    double2* tb = tile + ((long)t * Nd + comp) * Ny * TM;
    for (int e = threadIdx.x; e < Ny * TM; e += blockDim.x) {
        const int n = e / TM, m = e - n * TM, q = t * TM + m;
        if (q >= nq) {
            if (dir == 0) tb[e] = make_double2(0.0, 0.0);
            continue;
        }
        const int ml = q / nkz, kz = q - ml * nkz, mxi = x0 + ml;
        const int kx = mxi <= Kx ? mxi : mxi - nmx;
        const int mx = kx >= 0 ? kx : Nx + kx;
        const long so = comp * cs + n * rs + (long)mx * Mz + kz;
        if (dir == 0) tb[e] = ser[so];
        else ser[so] = tb[e];
    }
}

int tile_convert_launch(double* ser, double* tile, int Nx, int Ny, int Nz, int Nd, int Kx, int Kz, int x0, int nq, int TM, int dir,
                        cudaStream_t st) {
    const int ntiles = (nq + TM - 1) / TM;
    if (ntiles <= 0) return 0;
    CF_LAUNCH(tile_convert_kernel, dim3(ntiles, Nd), dim3(EW_THREADS), 0, st, reinterpret_cast<double2*>(ser),
              reinterpret_cast<double2*>(tile), Nx, Ny, Nz / 2 + 1, Nd, Kx, Kz + 1, x0, nq, TM, dir);
    CF_KERNEL_CHECK();
    return 0;
}

int l2form_launch(const double* u, const double* v, int mode, const double* W, int N, int Nx, int Nz, int Nd, int Kx, int Kz, int fullbox,
                  int x0, int x1, double scale, double* partial_dev, size_t partial_cap, double* out_dev, cudaStream_t st, double czw) {
    const int Mz = Nz / 2 + 1;
    const int nkz_ = fullbox ? Mz : Kz + 1;
    const int qlo = x0 * nkz_;
    const int nq = (x1 - x0) * nkz_;
    const int narr = mode == 2 ? 2 : 1;
    int TMn = 16;
    while (TMn > 1 && (size_t)N * 2 * TMn * narr * sizeof(double) > 160 * 1024) TMn >>= 1;
    const size_t smem = (size_t)N * 2 * TMn * narr * sizeof(double);
    static size_t configured = 0;
    auto kfn = l2form_kernel;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((nq + TMn - 1) / TMn, Nd);
    const int nparts = grid.x * grid.y;
    if ((size_t)nparts > partial_cap) {
        set_last_error("l2form: partial buffer too small");
        return 1;
    }
    const long rs = (long)Nx * 2 * Mz, cs = rs * N;
    CF_LAUNCH(kfn, grid, dim3(NRM_THREADS), smem, st, u, v, mode, W, N, Nx, Mz, Kx, Kz, fullbox, qlo, nq, TMn, rs, cs, czw, partial_dev);
    CF_KERNEL_CHECK();
    CF_LAUNCH(sum_partials_kernel, dim3(1), dim3(256), 0, st, (const double*)partial_dev, nparts, scale, out_dev);
    CF_KERNEL_CHECK();
    return 0;
}

}  // namespace cfgpu
