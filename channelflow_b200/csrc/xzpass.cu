// x-pass / z-pass kernels: batched complex FFTs in x, paired real FFTs in z fused with the pointwise nonlinear
// term.  See xzpass.cuh.  All three are HBM-bound streaming kernels (each pencil element read once, written
// once, 16-byte accesses contiguous along kz); the FFT itself runs out of shared memory.
#include "xzpass.cuh"

#include <cstdlib>

namespace cfgpu {

namespace {
constexpr int XZ_THREADS = 256;
constexpr double TWO_PI = 6.283185307179586476925286766559;
}  // namespace
static int set_smem(const void* fn, size_t bytes, size_t& configured);

// ------------------------------------------------------------------------------------------------ x inverse
// grid = (nfields, ceil(nkz/TZ), nyn)   P[src][yl][mxi][kz] -> Q[f][yl][nx][kz], optional d/dx = i 2 pi kx / Lx
// The output field is the FASTEST grid index: the CTAs that share source rows (u, v, w and the three components of curl u
// read 9 source tiles of 5 distinct fields) run next to each other and the re-reads hit L2 instead of HBM.
__global__ void __launch_bounds__(XZ_THREADS, 3) xpass_inverse_kernel(const XPassParams p) {
    const int Nx = p.Nx, Kx = p.Kx, TZ = p.TZ;
    const int nmx = 2 * Kx + 1, nkz = p.Kz + 1;
    double2* a = dyn_smem<double2>();                 // [Nx][TZ], transformed in place
    // the in-place passes with product-tree twiddles only touch the head of the table (64 entries at Nx = 512): small
    // enough that three CTAs fit an SM
    const int ntw = fft_plan_ntw(p.plan);
    double2* tws = a + (size_t)Nx * TZ;                 // twiddle table in shared memory
    double* kxf = reinterpret_cast<double*>(tws + ntw);  // 2 pi kx / Lx of pencil row mxi
    int* rev = reinterpret_cast<int*>(kxf + nmx);       // digit-reversed row of mode row mx (input side of the DIT transform)
    const int tid = threadIdx.x;
    const int f = p.fsel[blockIdx.x], yl = blockIdx.z, kz0 = blockIdx.y * TZ;
    const int s = p.src[f], oa = p.opa[f], sb = p.srcb[f], ob = p.opb[f];

    for (int t = tid; t < Nx; t += XZ_THREADS) {
        if (t < ntw) tws[t] = __ldg(&p.plan.tw[t]);
        rev[t] = __ldg(&p.plan.rev[t]);
    }
    for (int mxi = tid; mxi < nmx; mxi += XZ_THREADS) {
        const int kx = mxi <= Kx ? mxi : mxi - nmx;
        kxf[mxi] = TWO_PI * kx / p.Lx;
    }
    __syncthreads();
    // TZ is a power of two dividing the block size: a thread keeps its column c and walks the pencil rows
    const int tzs = __ffs(TZ) - 1;
    const int c = tid & (TZ - 1), kz = kz0 + c;
    const int mstep = XZ_THREADS >> tzs;
    // zero the aliased rows Kx+1 .. Nx-Kx-1
    for (int r = tid >> tzs; r < Nx - nmx; r += mstep) a[rev[Kx + 1 + r] * TZ + c] = make_double2(0.0, 0.0);
    const double2* __restrict__ in = p.in;
    const double kzf = TWO_PI * kz / p.Lz;
    const bool live = kz < nkz;
    // four rows per thread in flight: all loads are issued before the first use
    for (int m0 = tid >> tzs; m0 < nmx; m0 += 4 * mstep) {
        double2 va[4], vb[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int mxi = m0 + h * mstep;
            va[h] = vb[h] = make_double2(0.0, 0.0);
            if (mxi < nmx && live) {
                va[h] = in[xpass_row_offset(p, s, yl, mxi, nkz) + kz];
                if (sb >= 0) vb[h] = in[xpass_row_offset(p, sb, yl, mxi, nkz) + kz];
            }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int mxi = m0 + h * mstep;
            if (mxi >= nmx) break;
            const int mx = mxi <= Kx ? mxi : Nx - nmx + mxi;
            double2 v = va[h];
            if (oa) {
                const double k = oa == 1 ? kxf[mxi] : kzf;
                v = make_double2(-k * v.y, k * v.x);
            }
            if (sb >= 0) {
                double2 w = vb[h];
                if (ob) {
                    const double k = ob == 1 ? kxf[mxi] : kzf;
                    w = make_double2(-k * w.y, k * w.x);
                }
                v = make_double2(v.x - w.x, v.y - w.y);
            }
            a[rev[mx] * TZ + c] = v;
        }
    }
    __syncthreads();
    fft_smem_inplace<+1, false, true>(a, p.plan, tws, TZ, tid, XZ_THREADS);
    double2* __restrict__ out = p.out + ((size_t)f * p.nyn + yl) * Nx * nkz;
    if (live)
        for (int nx = tid >> tzs; nx < Nx; nx += mstep) out[(size_t)nx * nkz + kz] = a[nx * TZ + c];
}

// ------------------------------------------------------------------------------------------------ x forward
// grid = (ceil(nkz/TZ), nyn, nfields)   Q[f][yl][nx][kz] -> P[f][yl][mxi][kz]   (aliased kx are dropped)
__global__ void __launch_bounds__(XZ_THREADS) xpass_forward_kernel(const XPassParams p) {
    const int Nx = p.Nx, Kx = p.Kx, TZ = p.TZ;
    const int nmx = 2 * Kx + 1, nkz = p.Kz + 1;
    const int tid = threadIdx.x;
    const int f = p.fsel[blockIdx.z], yl = blockIdx.y, kz0 = blockIdx.x * TZ;
    const int sa = p.src[f], sb = p.srcb[f];
    const int C = sb >= 0 ? 2 * TZ : TZ;  // columns: TZ of the first source, then TZ of the second
    double2* a = dyn_smem<double2>();
    double2* b = a + (size_t)Nx * C;
    const double2* __restrict__ ina = p.in + ((size_t)sa * p.nyn + yl) * Nx * nkz;
    const double2* __restrict__ inb = p.in + ((size_t)(sb >= 0 ? sb : 0) * p.nyn + yl) * Nx * nkz;
    for (int idx = tid; idx < Nx * C; idx += XZ_THREADS) {
        const int nx = idx / C, c = idx - nx * C;
        const int kz = kz0 + (c < TZ ? c : c - TZ);
        a[idx] = kz < nkz ? (c < TZ ? ina : inb)[(size_t)nx * nkz + kz] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const double2* res = fft_smem<-1>(a, b, p.plan, p.plan.tw, C, tid, XZ_THREADS);
    for (int idx = tid; idx < nmx * TZ; idx += XZ_THREADS) {
        const int mxi = idx / TZ, c = idx - mxi * TZ;
        const int kz = kz0 + c;
        const int kx = mxi <= Kx ? mxi : mxi - nmx;
        const int mx = kx >= 0 ? kx : Nx + kx;
        if (kz >= nkz) continue;
        double2 v = res[mx * C + c];
        if (sb >= 0) {
            const double2 w = res[mx * C + TZ + c];
            const double k = p.cb * (p.opb[f] == 1 ? TWO_PI * kx / p.Lx : TWO_PI * kz / p.Lz);
            v = make_double2(v.x - k * w.y, v.y + k * w.x);
        }
        int s = -1;
        size_t o = 0;
        if (p.peer_direct) o = xpass_peer_offset(p, f, p.ny0 + yl, mxi, nkz, s);
        if (p.peer_direct == 1 || (p.peer_direct == 2 && s == p.self_rank)) {
            p.peer_out[s][o + kz] = v;
        } else {
            p.out[xpass_row_offset(p, f, yl, mxi, nkz) + kz] = v;
        }
    }
}

// In-place variant for long x-lines (Nx >= 1024): one shared-memory buffer instead of the Stockham pair, hence twice the kz
// columns per CTA for the same footprint -- the global-memory pieces, not the transform, limit the pass there (weak scaling
// along x: 2 -> 4 columns at Nx = 1024, 1 -> 2 at 2048).  Rows are written to their digit-reversed places while loading and
// the decimation-in-time passes run in place, as in the inverse pass.  At Nx = 512 the Stockham kernel is faster (0.67 vs 0.92 ms).
__global__ void __launch_bounds__(XZ_THREADS) xpass_forward_inplace_kernel(const XPassParams p) {
    const int Nx = p.Nx, Kx = p.Kx, TZ = p.TZ;
    const int nmx = 2 * Kx + 1, nkz = p.Kz + 1;
    const int tid = threadIdx.x;
    const int f = p.fsel[blockIdx.z], yl = blockIdx.y, kz0 = blockIdx.x * TZ;
    const int sa = p.src[f], sb = p.srcb[f];
    const int C = sb >= 0 ? 2 * TZ : TZ;  // columns: TZ of the first source, then TZ of the second
    const int ntw = fft_plan_ntw(p.plan);
    double2* a = dyn_smem<double2>();                    // [Nx][C], transformed in place
    double2* tws = a + (size_t)Nx * C;
    int* rev = reinterpret_cast<int*>(tws + ntw);
    for (int t = tid; t < Nx; t += XZ_THREADS) {
        if (t < ntw) tws[t] = __ldg(&p.plan.tw[t]);
        rev[t] = __ldg(&p.plan.rev[t]);
    }
    __syncthreads();
    const double2* __restrict__ ina = p.in + ((size_t)sa * p.nyn + yl) * Nx * nkz;
    const double2* __restrict__ inb = p.in + ((size_t)(sb >= 0 ? sb : 0) * p.nyn + yl) * Nx * nkz;
    for (int idx = tid; idx < Nx * C; idx += XZ_THREADS) {
        const int nx = idx / C, c = idx - nx * C;
        const int kz = kz0 + (c < TZ ? c : c - TZ);
        a[rev[nx] * C + c] = kz < nkz ? (c < TZ ? ina : inb)[(size_t)nx * nkz + kz] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    fft_smem_inplace<-1, false, true>(a, p.plan, tws, C, tid, XZ_THREADS);
    const double2* res = a;
    for (int idx = tid; idx < nmx * TZ; idx += XZ_THREADS) {
        const int mxi = idx / TZ, c = idx - mxi * TZ;
        const int kz = kz0 + c;
        const int kx = mxi <= Kx ? mxi : mxi - nmx;
        const int mx = kx >= 0 ? kx : Nx + kx;
        if (kz >= nkz) continue;
        double2 v = res[mx * C + c];
        if (sb >= 0) {
            const double2 w = res[mx * C + TZ + c];
            const double k = p.cb * (p.opb[f] == 1 ? TWO_PI * kx / p.Lx : TWO_PI * kz / p.Lz);
            v = make_double2(v.x - k * w.y, v.y + k * w.x);
        }
        int s = -1;
        size_t o = 0;
        if (p.peer_direct) o = xpass_peer_offset(p, f, p.ny0 + yl, mxi, nkz, s);
        if (p.peer_direct == 1 || (p.peer_direct == 2 && s == p.self_rank)) {
            p.peer_out[s][o + kz] = v;
        } else {
            p.out[xpass_row_offset(p, f, yl, mxi, nkz) + kz] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ z pass
// grid = (ceil(Nx/TL), nyn).  Two real fields share one complex transform (Z = A + iB with the Hermitian
// extension written explicitly), so the rotational term (u, v, w, curl u) needs 3 inverse + 2 forward complex FFTs per line.
__global__ void __launch_bounds__(512) zpass_kernel(const ZPassParams p) {
    const int Nx = p.Nx, Nz = p.Nz, TL = p.TL;
    const int nkz = p.Kz + 1;
    const bool rot = p.mode == ZP_ROTATIONAL;
    const int npair = rot ? 3 : 2;
    const int CA = npair * TL;
    double2* A = dyn_smem<double2>();
    double2* B = A + (size_t)Nz * CA;
    __shared__ double red[32];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int yl = blockIdx.y, ny = p.ny0 + yl, nx0 = blockIdx.x * TL;
    const size_t fstride = (size_t)p.nyn * Nx * nkz;  // field stride of Q and F

    // rows nkz .. Nz-nkz (the de-aliased band and the Nyquist mode) are not written by the packing loop below
    for (int idx = tid; idx < (Nz - 2 * nkz + 1) * CA; idx += NT) A[(size_t)nkz * CA + idx] = make_double2(0.0, 0.0);

    const double2* __restrict__ Q = p.Q + (size_t)yl * Nx * nkz;
    for (int idx = tid; idx < TL * nkz; idx += NT) {
        const int l = idx / nkz, k = idx - l * nkz;
        const int nx = nx0 + l;
        if (nx >= Nx) continue;
        const size_t off = (size_t)nx * nkz + k;
        double2 fa[3], fb[3];
        if (rot) {
            fa[0] = Q[off];               fb[0] = Q[fstride + off];       // u, v
            fa[1] = Q[2 * fstride + off]; fb[1] = Q[3 * fstride + off];   // w, omega_x
            fa[2] = Q[4 * fstride + off]; fb[2] = Q[5 * fstride + off];   // omega_y, omega_z
        } else {
            fa[0] = Q[off];               fb[0] = Q[fstride + off];
            fa[1] = Q[2 * fstride + off]; fb[1] = make_double2(0.0, 0.0);
        }
        for (int q = 0; q < npair; ++q) {
            double2 a = fa[q], b = fb[q];
            if (k == 0) { a.y = 0.0; b.y = 0.0; }  // c2r ignores the imaginary part of the mean mode
            A[(size_t)k * CA + q * TL + l] = make_double2(a.x - b.y, a.y + b.x);
            if (k > 0) A[(size_t)(Nz - k) * CA + q * TL + l] = make_double2(a.x + b.y, b.x - a.y);
        }
    }
    __syncthreads();
    double2* res = fft_smem<+1>(A, B, p.plan, p.plan.tw, CA, tid, NT);
    double2* other = (res == A) ? B : A;

    // pointwise stage
    double U = 0.0, Uy = 0.0, W = 0.0, Wy = 0.0;
    if (p.Uy) {
        U = p.Uy[ny];
        Uy = p.Uy[p.Ny + ny];
        W = p.Uy[2 * p.Ny + ny];
        Wy = p.Uy[3 * p.Ny + ny];
    }
    const double idx_ = (double)Nx / p.Lx, idz_ = (double)Nz / p.Lz, idy_ = p.inv_dy ? p.inv_dy[ny] : 0.0;
    double cmax = 0.0;
    const int CF = 2 * TL;
    for (int idx = tid; idx < Nz * TL; idx += NT) {
        const int z = idx / TL, l = idx - z * TL;
        if (nx0 + l >= Nx) {
            if (rot) {
                other[(size_t)z * CF + l] = make_double2(0.0, 0.0);
                other[(size_t)z * CF + TL + l] = make_double2(0.0, 0.0);
            }
            continue;
        }
        const double2* r = res + (size_t)z * CA + l;
        if (rot) {
            const double2 z0 = r[0], z1 = r[TL], z2 = r[2 * TL];
            const double u = z0.x, v = z0.y, w = z1.x;
            const double ut = u + U, vt = v - p.Vsuck, wt = w + W;
            const double ox = z1.y + Wy;   // curl of the total velocity: the base flow adds (W', 0, -U')
            const double oy = z2.x;
            const double oz = z2.y - Uy;
            double fx = oy * wt - oz * vt;
            double fy = oz * ut - ox * wt;
            const double fz = ox * vt - oy * ut;
            if (p.rotation != 0.0) {
                fx -= p.rotation * vt;
                fy += p.rotation * ut;
            }
            other[(size_t)z * CF + l] = make_double2(fx, fy);
            other[(size_t)z * CF + TL + l] = make_double2(fz, 0.0);
            double m = ut * idx_;
            const double m2 = v * idy_, m3 = wt * idz_;
            m = m2 > m ? m2 : m;
            m = m3 > m ? m3 : m;
            cmax = m > cmax ? m : cmax;
        } else {
            const double2 z0 = r[0], z1 = r[TL];
            double m = (z0.x + U) * idx_;
            const double m2 = z0.y * idy_, m3 = (z1.x + W) * idz_;
            m = m2 > m ? m2 : m;
            m = m3 > m ? m3 : m;
            cmax = m > cmax ? m : cmax;
        }
    }
    if (p.cfl_max) {
        cmax = warp_max(cmax);
        if ((tid & 31) == 0) red[tid >> 5] = cmax;
        __syncthreads();
        if (tid < 32) {
            double v = tid < (NT >> 5) ? red[tid] : 0.0;
            v = warp_max(v);
            if (tid == 0 && v > 0.0) atomic_max_double(p.cfl_max, v);
        }
    }
    if (!rot) return;
    __syncthreads();
    const double2* g = fft_smem<-1>(other, res, p.plan, p.plan.tw, CF, tid, NT);

    double2* __restrict__ F = p.F + (size_t)yl * Nx * nkz;
    const double sc = p.scale, hs = 0.5 * p.scale;
    for (int idx = tid; idx < TL * nkz; idx += NT) {
        const int l = idx / nkz, k = idx - l * nkz;
        const int nx = nx0 + l;
        if (nx >= Nx) continue;
        const int kn = k == 0 ? 0 : Nz - k;
        const double2 gk = g[(size_t)k * CF + l], gn = g[(size_t)kn * CF + l], hk = g[(size_t)k * CF + TL + l];
        const size_t off = (size_t)nx * nkz + k;
        F[off] = make_double2(hs * (gk.x + gn.x), hs * (gk.y - gn.y));
        F[fstride + off] = make_double2(hs * (gk.y + gn.y), -hs * (gk.x - gn.x));
        F[2 * fstride + off] = make_double2(sc * hk.x, sc * hk.y);
    }
}

// ------------------------------------------------------------------------------------------------ z pass, warp-private FFTs
// Same work as zpass_kernel for Nz <= 512: every complex transform (3 inverse + 2 forward per x-line for the rotational
// form) is owned by ONE warp and runs in place in its own skewed shared-memory line without block barriers
// (fft_smem.cuh: warp_fft); the CTA only synchronises between pack / inverse / pointwise / forward / store.
// grid = (ceil(Nx/TL), nyn), block = 32 * npair * TL threads.  NZ = Nz (power of two, compile-time FFT plan).
// (at Nz = 512 two lines = 6 warps per CTA and three CTAs per SM: the register budget is set for exactly that)
template <int NZ>
__global__ void __launch_bounds__(NZ == 512 ? 192 : 320, NZ == 512 ? 4 : 2) zpass_warp_kernel(const ZPassParams p) {
    const int Nx = p.Nx, Nz = NZ, TL = p.TL;
    const int nkz = p.Kz + 1;
    const bool rot = p.mode == ZP_ROTATIONAL;
    const int npair = rot ? 3 : 2;
    const int njobs = npair * TL;           // job j = q * TL + l : transform q of line l
    const int NP = fft_skew2_len(Nz);
    constexpr int NTW = NZ / (WarpFftShape<NZ>::RL > 1 ? WarpFftShape<NZ>::RL : 8);  // twiddles the in-place passes touch
    double2* buf = dyn_smem<double2>();     // [njobs][NP]
    double2* tws = buf + (size_t)njobs * NP; // twiddle table exp(-2 pi i t / Nz), t < NTW
    double* red = reinterpret_cast<double*>(tws + NTW);  // [warps]
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int yl = blockIdx.y, ny = p.ny0 + yl, nx0 = blockIdx.x * TL;
    const size_t fstride = (size_t)p.nyn * Nx * nkz;  // field stride of Q and F

    for (int t = tid; t < NTW; t += NT) tws[t] = __ldg(&p.plan.tw[t]);
    // rows nkz .. Nz-nkz (the de-aliased band and the Nyquist mode) are not written by the packing loop below
    const int nzero = Nz - 2 * nkz + 1;
    for (int k = nkz + tid; k < nkz + nzero; k += NT) {   // (no runtime divisions in the index arithmetic of this kernel)
        const int a = fft_skew2(fft_digit_rev<NZ>(k));
        for (int j = 0; j < njobs; ++j) buf[(size_t)j * NP + a] = make_double2(0.0, 0.0);
    }
    const double2* __restrict__ Q = p.Q + (size_t)yl * Nx * nkz;
    // item = (line l, mode k): thread t takes mode k = t (+ NT ...) of two lines at a time, so that two items are in flight
    // (all loads issued before the first use) and the digit-reversed slot of k is computed once per pair
    for (int k = tid; k < nkz; k += NT) {
        const int ka = fft_skew2(fft_digit_rev<NZ>(k)), kb = fft_skew2(fft_digit_rev<NZ>(k > 0 ? Nz - k : 0));  // inputs of the DIT transform
        for (int l0 = 0; l0 < TL; l0 += 2) {
            double2 fa[2][3], fb[2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int l = l0 + h, nx = nx0 + l;
#pragma unroll
                for (int q = 0; q < 3; ++q) fa[h][q] = fb[h][q] = make_double2(0.0, 0.0);
                if (l < TL && nx < Nx) {
                    const size_t off = (size_t)nx * nkz + k;
                    if (rot) {
                        fa[h][0] = Q[off];               fb[h][0] = Q[fstride + off];       // u, v
                        fa[h][1] = Q[2 * fstride + off]; fb[h][1] = Q[3 * fstride + off];   // w, omega_x
                        fa[h][2] = Q[4 * fstride + off]; fb[h][2] = Q[5 * fstride + off];   // omega_y, omega_z
                    } else {
                        fa[h][0] = Q[off];               fb[h][0] = Q[fstride + off];
                        fa[h][1] = Q[2 * fstride + off];
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int l = l0 + h;
                if (l >= TL) break;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q >= npair) break;
                    double2 a = fa[h][q], b = fb[h][q];
                    if (k == 0) { a.y = 0.0; b.y = 0.0; }  // c2r ignores the imaginary part of the mean mode
                    double2* line = buf + (size_t)(q * TL + l) * NP;
                    line[ka] = make_double2(a.x - b.y, a.y + b.x);
                    if (k > 0) line[kb] = make_double2(a.x + b.y, b.x - a.y);
                }
            }
        }
    }
    __syncthreads();
    for (int j = warp; j < njobs; j += (NT >> 5)) warp_fft_dit<+1, NZ>(buf + (size_t)j * NP, tws, lane);
    __syncthreads();

    // pointwise stage: (line l, point z); the products overwrite transforms 0 and 1 of the same line
    double U = 0.0, Uy = 0.0, W = 0.0, Wy = 0.0;
    if (p.Uy) {
        U = p.Uy[ny];
        Uy = p.Uy[p.Ny + ny];
        W = p.Uy[2 * p.Ny + ny];
        Wy = p.Uy[3 * p.Ny + ny];
    }
    const double idx_ = (double)Nx / p.Lx, idz_ = (double)Nz / p.Lz, idy_ = p.inv_dy ? p.inv_dy[ny] : 0.0;
    double cmax = 0.0;
    const size_t js = (size_t)TL * NP;  // stride between transforms of one line
    // The third component of the product is real: the f_z lines of TWO neighbouring x-lines share one forward transform
    // (f_z(l) + i f_z(l+1), separated again by the conjugate symmetry in the store), so a pair of lines costs three forward
    // transforms instead of four.  One thread owns point z of both lines of a pair: it reads every input of the two lines
    // before it writes, and the pair's f_z goes to the transform-1 slot of the pair's first line, which only it reads.
    const int NLP = (TL + 1) >> 1;
    if (rot) {
        for (int idx = tid; idx < Nz * NLP; idx += NT) {
            const int lp = idx / Nz, z = idx - lp * Nz;
            const int zs = fft_skew2(z);
            double fzv[2] = {0.0, 0.0};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int l = 2 * lp + h;
                if (l >= TL || nx0 + l >= Nx) continue;
                double2* r = buf + (size_t)l * NP + zs;
                const double2 z0 = r[0], z1 = r[js], z2 = r[2 * js];
                const double u = z0.x, v = z0.y, w = z1.x;
                const double ut = u + U, vt = v - p.Vsuck, wt = w + W;
                const double ox = z1.y + Wy;   // curl of the total velocity: the base flow adds (W', 0, -U')
                const double oy = z2.x;
                const double oz = z2.y - Uy;
                double fx = oy * wt - oz * vt;
                double fy = oz * ut - ox * wt;
                fzv[h] = ox * vt - oy * ut;
                if (p.rotation != 0.0) {
                    fx -= p.rotation * vt;
                    fy += p.rotation * ut;
                }
                r[0] = make_double2(fx, fy);
                double m = ut * idx_;
                const double m2 = v * idy_, m3 = wt * idz_;
                m = m2 > m ? m2 : m;
                m = m3 > m ? m3 : m;
                cmax = m > cmax ? m : cmax;
            }
            buf[(size_t)(TL + 2 * lp) * NP + zs] = make_double2(fzv[0], fzv[1]);
        }
    } else {
        for (int idx = tid; idx < Nz * TL; idx += NT) {
            const int l = idx / Nz, z = idx - l * Nz;
            if (nx0 + l >= Nx) continue;
            const double2* r = buf + (size_t)l * NP + fft_skew2(z);
            const double2 z0 = r[0], z1 = r[js];
            double m = (z0.x + U) * idx_;
            const double m2 = z0.y * idy_, m3 = (z1.x + W) * idz_;
            m = m2 > m ? m2 : m;
            m = m3 > m ? m3 : m;
            cmax = m > cmax ? m : cmax;
        }
    }
    if (p.cfl_max) {
        cmax = warp_max(cmax);
        if (lane == 0) red[warp] = cmax;
        __syncthreads();
        if (tid < 32) {
            double v = tid < (NT >> 5) ? red[tid] : 0.0;
            v = warp_max(v);
            if (tid == 0 && v > 0.0) atomic_max_double(p.cfl_max, v);
        }
    }
    if (!rot) return;
    __syncthreads();
    for (int j = warp; j < TL + NLP; j += (NT >> 5)) {
        const int slot = j < TL ? j : TL + 2 * (j - TL);
        warp_fft_dif<-1, NZ>(buf + (size_t)slot * NP, tws, lane);
    }
    __syncthreads();

    double2* __restrict__ F = p.F + (size_t)yl * Nx * nkz;
    const double hs = 0.5 * p.scale;
    for (int k = tid; k < nkz; k += NT) {
        const int ks = fft_skew2(fft_digit_rev<NZ>(k)), kn = fft_skew2(fft_digit_rev<NZ>(k == 0 ? 0 : Nz - k));  // DIF outputs
        for (int l = 0; l < TL; ++l) {
            const int nx = nx0 + l;
            if (nx >= Nx) break;
            const double2* g = buf + (size_t)l * NP;                        // transform 0 of line l: fx + i fy
            const double2* h = buf + (size_t)(TL + (l & ~1)) * NP;          // f_z of the pair: fz(even line) + i fz(odd line)
            const double2 gk = g[ks], gn = g[kn], hk = h[ks], hn = h[kn];
            const size_t off = (size_t)nx * nkz + k;
            F[off] = make_double2(hs * (gk.x + gn.x), hs * (gk.y - gn.y));
            F[fstride + off] = make_double2(hs * (gk.y + gn.y), -hs * (gk.x - gn.x));
            F[2 * fstride + off] = (l & 1) ? make_double2(hs * (hk.y + hn.y), -hs * (hk.x - hn.x)) : make_double2(hs * (hk.x + hn.x), hs * (hk.y - hn.y));
        }
    }
}

// ------------------------------------------------------------------------------------------------ z pass, other forms
// Convection / divergence / skew-symmetric forms (see xzpass.cuh) for power-of-two Nz <= 512: one x-line per CTA, one
// warp per complex transform (6 inverse c2r pairs for the 12 fields u, grad u; up to 5 forward pairs for u.grad u and
// the 6 products u_i u_j).  grid = (Nx, nyn), block = 192 threads.
template <int NZ>
__global__ void __launch_bounds__(192, 2) zpass_general_kernel(const ZPassParams p) {
    const int Nx = p.Nx, Nz = NZ;
    const int nkz = p.Kz + 1;
    const int mode = p.mode;
    const bool need_grad = mode != ZP_DIVERGENCE, need_prod = mode != ZP_CONVECTION;
    const int ninv = need_grad ? 6 : 2;
    constexpr int NP = NZ + (NZ >> 3) + 1;
    double2* buf = dyn_smem<double2>();  // [6][NP]
    double2* tws = buf + 6 * NP;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
    const int yl = blockIdx.y, ny = p.ny0 + yl, nx = blockIdx.x;
    const size_t fstride = (size_t)p.nyn * Nx * nkz;

    for (int t = tid; t < Nz; t += NT) tws[t] = __ldg(&p.plan.tw[t]);
    const int nzero = Nz - 2 * nkz + 1;
    for (int idx = tid; idx < nzero * ninv; idx += NT) {
        const int j = idx / nzero, k = nkz + (idx - j * nzero);
        buf[j * NP + fft_skew(k)] = make_double2(0.0, 0.0);
    }
    const double2* __restrict__ Q = p.Q + (size_t)yl * Nx * nkz + (size_t)nx * nkz;
    for (int k = tid; k < nkz; k += NT) {
        double2 q[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) q[i] = (i < 3 || need_grad) ? Q[i * fstride + k] : make_double2(0.0, 0.0);
        const int ka = fft_skew(k), kb = fft_skew(k > 0 ? Nz - k : 0);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (j >= ninv) break;
            double2 a = q[2 * j], b = (need_grad || j == 0) ? q[2 * j + 1] : make_double2(0.0, 0.0);
            if (k == 0) { a.y = 0.0; b.y = 0.0; }
            buf[j * NP + ka] = make_double2(a.x - b.y, a.y + b.x);
            if (k > 0) buf[j * NP + kb] = make_double2(a.x + b.y, b.x - a.y);
        }
    }
    __syncthreads();
    for (int j = warp; j < ninv; j += NW) warp_fft_pow2<+1, NZ>(buf + j * NP, tws, lane);
    __syncthreads();

    double U = 0.0, Uy = 0.0, W = 0.0, Wy = 0.0;
    if (p.Uy) {
        U = p.Uy[ny];
        Uy = p.Uy[p.Ny + ny];
        W = p.Uy[2 * p.Ny + ny];
        Wy = p.Uy[3 * p.Ny + ny];
    }
    for (int z = tid; z < Nz; z += NT) {
        const int zs = fft_skew(z);
        double v[3], g[3][3];  // g[i][j] = d u_i / d x_j
        {
            const double2 z0 = buf[zs], z1 = buf[NP + zs];
            v[0] = z0.x + U; v[1] = z0.y - p.Vsuck; v[2] = z1.x + W;
            if (need_grad) {
                const double2 z2 = buf[2 * NP + zs], z3 = buf[3 * NP + zs], z4 = buf[4 * NP + zs], z5 = buf[5 * NP + zs];
                g[0][1] = z1.y + Uy; g[1][1] = z2.x; g[2][1] = z2.y + Wy;   // d/dy of (u + U, v, w + W)
                g[0][0] = z3.x; g[1][0] = z3.y; g[2][0] = z4.x;             // d/dx
                g[0][2] = z4.y; g[1][2] = z5.x; g[2][2] = z5.y;             // d/dz
            }
        }
        double c[3] = {0.0, 0.0, 0.0};
        if (need_grad) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 3; ++j) s += p.cc * v[j] * g[i][j];
                c[i] = s;
            }
            if (p.rotation != 0.0) {
                c[0] -= p.rotation * v[1];
                c[1] += p.rotation * v[0];
            }
        }
        if (!need_prod) {
            buf[zs] = make_double2(c[0], c[1]);
            buf[NP + zs] = make_double2(c[2], 0.0);
        } else {
            buf[zs] = make_double2(c[0], c[1]);
            buf[NP + zs] = make_double2(c[2], v[0] * v[0]);
            buf[2 * NP + zs] = make_double2(v[0] * v[1], v[0] * v[2]);
            buf[3 * NP + zs] = make_double2(v[1] * v[1], v[1] * v[2]);
            buf[4 * NP + zs] = make_double2(v[2] * v[2], 0.0);
        }
    }
    __syncthreads();
    const int nfwd = need_prod ? 5 : 2;
    for (int j = warp; j < nfwd; j += NW) warp_fft_pow2<-1, NZ>(buf + j * NP, tws, lane);
    __syncthreads();

    double2* __restrict__ F = p.F + (size_t)yl * Nx * nkz + (size_t)nx * nkz;
    const double sc = p.scale;
    for (int k = tid; k < nkz; k += NT) {
        const int ks = fft_skew(k), kn = fft_skew(k == 0 ? 0 : Nz - k);
        // transform j holds A + iB of two real lines: A_k = (g_k + conj g_{N-k})/2, B_k = (g_k - conj g_{N-k})/(2i)
        double2 A[5], B[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (j >= nfwd) break;
            const double2 gk = buf[j * NP + ks], gn = buf[j * NP + kn];
            A[j] = make_double2(0.5 * sc * (gk.x + gn.x), 0.5 * sc * (gk.y - gn.y));
            B[j] = make_double2(0.5 * sc * (gk.y + gn.y), -0.5 * sc * (gk.x - gn.x));
        }
        if (!need_prod) {
            F[k] = A[0]; F[fstride + k] = B[0]; F[2 * fstride + k] = A[1];
        } else {
            // conv = A0, B0, A1 ; uu = B1, uv = A2, uw = B2, vv = A3, vw = B3, ww = A4
            const double kzz = p.cd * TWO_PI * k / p.Lz;
            const double2 cv[3] = {A[0], B[0], A[1]};
            const double2 t2[3] = {B[2], B[3], A[4]};  // u_i w
#pragma unroll
            for (int i = 0; i < 3; ++i) F[i * fstride + k] = make_double2(cv[i].x - kzz * t2[i].y, cv[i].y + kzz * t2[i].x);
            F[3 * fstride + k] = B[1]; F[4 * fstride + k] = A[2]; F[5 * fstride + k] = B[2];
            F[6 * fstride + k] = A[3]; F[7 * fstride + k] = B[3];
        }
    }
}

template <int NZ>
static int zpass_general_launch(const ZPassParams& p, cudaStream_t stream) {
    const size_t smem = ((size_t)6 * (NZ + (NZ >> 3) + 1) + NZ) * sizeof(double2);
    static size_t configured = 0;
    auto kfn = zpass_general_kernel<NZ>;
    CF_TRY(set_smem((const void*)kfn, smem, configured));
    dim3 grid(p.Nx, p.nyn);
    CF_LAUNCH(kfn, grid, dim3(192), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}
bool zpass_general_supported(int Nz) { return Nz >= 16 && Nz <= 512 && (Nz & (Nz - 1)) == 0; }

// ------------------------------------------------------------------------------------------------ launchers
static int set_smem(const void* fn, size_t bytes, size_t& configured) {
    if (bytes > configured) {
        CF_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
    return 0;
}

int xpass_inverse_launch(const XPassParams& p, cudaStream_t stream) {
    const int nkz = p.Kz + 1;
    if (p.TZ & (p.TZ - 1) || XZ_THREADS % p.TZ) { set_last_error("xpass_inverse: TZ must be a power of two"); return 1; }
    const int ntw = fft_plan_ntw(p.plan);
    const size_t smem = ((size_t)p.Nx * p.TZ + ntw) * sizeof(double2) + (size_t)(2 * p.Kx + 1) * sizeof(double) + (size_t)p.Nx * sizeof(int);
    static size_t configured = 0;
    auto kfn = xpass_inverse_kernel;
    CF_TRY(set_smem((const void*)kfn, smem, configured));
    dim3 grid(p.nfields, (nkz + p.TZ - 1) / p.TZ, p.nyn);
    CF_LAUNCH(kfn, grid, dim3(XZ_THREADS), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}

int xpass_forward_launch(const XPassParams& p, cudaStream_t stream) {
    const int nkz = p.Kz + 1;
    bool two = false;
    for (int i = 0; i < p.nfields; ++i) two = two || p.srcb[p.fsel[i]] >= 0;
    dim3 grid((nkz + p.TZ - 1) / p.TZ, p.nyn, p.nfields);
    if (p.inplace) {
        const size_t smem = (size_t)p.Nx * p.TZ * (two ? 2 : 1) * sizeof(double2) + (size_t)fft_plan_ntw(p.plan) * sizeof(double2) + (size_t)p.Nx * sizeof(int);
        static size_t configured_ip = 0;
        auto kfn = xpass_forward_inplace_kernel;
        CF_TRY(set_smem((const void*)kfn, smem, configured_ip));
        CF_LAUNCH(kfn, grid, dim3(XZ_THREADS), smem, stream, p);
        CF_KERNEL_CHECK();
        return 0;
    }
    const size_t smem = 2 * (size_t)p.Nx * p.TZ * (two ? 2 : 1) * sizeof(double2);
    static size_t configured = 0;
    auto kfn = xpass_forward_kernel;
    CF_TRY(set_smem((const void*)kfn, smem, configured));
    CF_LAUNCH(kfn, grid, dim3(XZ_THREADS), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}

template <int NZ>
static int zpass_warp_launch(const ZPassParams& p0, cudaStream_t stream) {
    ZPassParams p = p0;
    const int npair = p.mode == ZP_ROTATIONAL ? 3 : 2;
    // lines per CTA: small CTAs (two lines = 6 warps at Nz = 512, three CTAs per SM) interleave their pack / FFT / store
    // phases better than fewer large ones (measured at 512x257x512: 1 line 2.83 ms, 2 lines 2.75 ms, 3 lines 2.97 ms)
    const size_t per_line = (size_t)npair * fft_skew2_len(p.Nz) * sizeof(double2);
    int TL = 1;
    const size_t cap = (size_t)(getenv("CF_ZP_SMEM_KB") ? atoi(getenv("CF_ZP_SMEM_KB")) : 60) * 1024;
    while (TL < 8 && (size_t)(TL + 1) * per_line <= cap && npair * (TL + 1) <= (NZ == 512 ? 6 : 10) && TL + 1 <= p.Nx) ++TL;
    p.TL = TL;
    constexpr int NTW = NZ / (WarpFftShape<NZ>::RL > 1 ? WarpFftShape<NZ>::RL : 8);
    const size_t smem = (size_t)TL * per_line + (size_t)NTW * sizeof(double2) + 16 * sizeof(double);
    static size_t configured = 0;
    auto kfn = zpass_warp_kernel<NZ>;
    CF_TRY(set_smem((const void*)kfn, smem, configured));
    dim3 grid((p.Nx + TL - 1) / TL, p.nyn);
    int nt = 32 * npair * TL;
    if (nt < 128) nt = 128;
    if (getenv("CF_ZP_THREADS")) nt = atoi(getenv("CF_ZP_THREADS"));  // experiment knob (<= the kernel's launch bound)
    CF_LAUNCH(kfn, grid, dim3(nt), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}

int zpass_launch(const ZPassParams& p, cudaStream_t stream) {
    if (p.mode >= ZP_CONVECTION) {
        switch (p.Nz) {
            case 16: return zpass_general_launch<16>(p, stream);
            case 32: return zpass_general_launch<32>(p, stream);
            case 64: return zpass_general_launch<64>(p, stream);
            case 128: return zpass_general_launch<128>(p, stream);
            case 256: return zpass_general_launch<256>(p, stream);
            case 512: return zpass_general_launch<512>(p, stream);
            default: set_last_error("zpass: this nonlinear form needs a power-of-two Nz <= 512"); return 1;
        }
    }
    switch (p.Nz) {  // warp-private transforms for the power-of-two lengths
        case 16: return zpass_warp_launch<16>(p, stream);
        case 32: return zpass_warp_launch<32>(p, stream);
        case 64: return zpass_warp_launch<64>(p, stream);
        case 128: return zpass_warp_launch<128>(p, stream);
        case 256: return zpass_warp_launch<256>(p, stream);
        case 512: return zpass_warp_launch<512>(p, stream);
        default: break;
    }
    const int npair = p.mode == ZP_ROTATIONAL ? 3 : 2;
    const size_t smem = 2 * (size_t)p.Nz * npair * p.TL * sizeof(double2);
    static size_t configured = 0;
    auto kfn = zpass_kernel;
    CF_TRY(set_smem((const void*)kfn, smem, configured));
    dim3 grid((p.Nx + p.TL - 1) / p.TL, p.nyn);
    // one first-pass butterfly per thread: Nz/R0 butterflies for each of the npair*TL columns
    const int R0 = p.plan.npass ? p.plan.radix[0] : 1;
    int nt = ((p.Nz / R0) * npair * p.TL + 31) & ~31;
    nt = nt < 128 ? 128 : (nt > 512 ? 512 : nt);
    CF_LAUNCH(kfn, grid, dim3(nt), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}

}  // namespace cfgpu
