// Batched tau-solver kernels (setup = factorisation + influence matrix, solve = fused RHS + Kleiser-Schumann).
// See tau.cuh for what is replaced and for the tile-major storage of the per-mode factors.
//
// Solve kernel: one CTA owns one tile of TM consecutive retained modes.  The whole CTA streams the history fields,
// accumulates the right-hand side into shared memory as real columns (re or im part of one mode's Chebyshev profile,
// stored skewed so that per-lane contiguous accesses are conflict free) and stages the tile's UL factors; then every
// WARP owns one column: lane l holds E consecutive coefficients in registers and each recurrence of the reference
// (derivative recurrence, UL back substitution, forward elimination) is evaluated as a blocked scan over warp
// shuffles (col_deriv / col_solve below).  The C&H "B" row multiply (helmholtz.cpp:81-85) and the right-hand sides
// of the four Helmholtz problems per mode are formed in registers; the influence-matrix and tau corrections and the
// scatter are data-parallel passes of the whole CTA.
// Setup kernel (rare: once per dt): one thread per (mode, parity) walks the reference's recurrences sequentially.
// Roofline: HBM (history fields + factors are each read once, outputs written once).
#include "tau.cuh"

namespace cfgpu {

namespace {

constexpr int TAU_THREADS = 256;
constexpr int TAU_SETUP_THREADS = 128;
constexpr double PI = 3.14159265358979323846264338327950288;

__host__ __device__ __forceinline__ double cN(int m, int Nb) { return (m == 0 || m == Nb) ? 2.0 : 1.0; }
__host__ __device__ __forceinline__ int betaN(int n, int Nb) { return (n > Nb - 2) ? 0 : 1; }
// C&H 5.1.24 rows n >= 2 of the quasi-tridiagonal systems (helmholtz.cpp:44-56)
// sub-diagonal of A = lambda-weighted B row (helmholtz.cpp:44-52).  Written as -(lambda * B_lo) everywhere -- setup and
// solve -- so that the solve kernel gets it from the B table with one multiply instead of a division per element.
__host__ __device__ __forceinline__ double B_lo(int n, int Nb);
__device__ __forceinline__ double A_lo_from_B(double blo, double lam) { return -(lam * blo); }
__device__ __forceinline__ double A_lo(int n, int Nb, double lam) { return A_lo_from_B(cN(n - 2, Nb) / (double)(4 * n * (n - 1)), lam); }
__device__ __forceinline__ double A_dg(int n, int Nb, double lam, double nus) {
    return nus + (betaN(n, Nb) * lam) / (double)(2 * (n * n - 1));
}
__device__ __forceinline__ double A_up(int n, int Nb, double lam) {
    return betaN(n + 2, Nb) ? -lam / (double)(4 * n * (n + 1)) : 0.0;
}
__host__ __device__ __forceinline__ double B_lo(int n, int Nb) { return cN(n - 2, Nb) / (double)(4 * n * (n - 1)); }
__host__ __device__ __forceinline__ double B_dg(int n, int Nb) { return -((double)betaN(n, Nb)) / (double)(2 * (n * n - 1)); }
__host__ __device__ __forceinline__ double B_up(int n, int Nb) { return betaN(n + 2, Nb) ? 1.0 / (double)(4 * n * (n + 1)) : 0.0; }

// g = B f (helmholtz.cpp:81-85, bandedtridiag.cpp:315-333), boundary rows set to bc0 (n=0) / bc1 (n=1).
__device__ __forceinline__ void bmul(const double* f, double* g, int N, int TT, double bc0, double bc1, int tid, int NT) {
    const int Nb = N - 1;
    for (int idx = tid; idx < N * TT; idx += NT) {
        const int n = idx / TT;
        double v;
        if (n == 0) v = bc0;
        else if (n == 1) v = bc1;
        else {
            v = B_lo(n, Nb) * f[idx - 2 * TT] + B_dg(n, Nb) * f[idx];
            if (n + 2 <= Nb) v += B_up(n, Nb) * f[idx + 2 * TT];
        }
        g[idx] = v;
    }
}

// UL solve of one parity block in place (bandedtridiag.cpp:258-273), reference operation order.  Used by the
// (rare) setup kernel only; factor arrays are the tile's [n][TM] arrays in HBM.
__device__ __forceinline__ void ul_solve_chain(double* g, int N, int TT, int t, int par, const double* up,
                                               const double* inv, const double* band, int TM, int m, double lam) {
    const int Nb = N - 1;
    const int nl = par ? Nb - 1 : Nb;
    for (int n = nl - 2; n >= par + 2; n -= 2) g[n * TT + t] -= up[m * N + n] * g[(n + 2) * TT + t];
    double acc = g[par * TT + t];
    for (int n = par + 2; n <= nl; n += 2) acc -= band[m * N + n] * g[n * TT + t];
    acc /= inv[m * N + par];  // slot `par` of inv holds diag(0) of this parity block
    g[par * TT + t] = acc;
    double prev = acc;
    for (int n = par + 2; n <= nl; n += 2) {
        const double v = (g[n * TT + t] - A_lo(n, Nb, lam) * prev) * inv[m * N + n];
        g[n * TT + t] = v;
        prev = v;
    }
}

// d = du/dy for the entries of parity `par` (chebyshev.cpp:672-697); optional d -= sub.
__device__ __forceinline__ void diff_chain(const double* u, double* d, int N, int TT, int t, int par, double scale,
                                           const double* sub) {
    const int Nb = N - 1;
    const int nl = ((Nb & 1) == par) ? Nb : Nb - 1;
    double run = 0.0;
    for (int n = nl; n >= par; n -= 2) {
        if (n + 1 <= Nb) run = run + scale * (n + 1) * u[(n + 1) * TT + t];
        double v = run;
        if (n == 0) { v *= 0.5; }
        d[n * TT + t] = sub ? v - sub[n * TT + t] : v;
    }
}

// eval_b / eval_a of d = du/dy (chebyshev.cpp:405-430 applied to diff): sums run from n = N-1 down to 0.
__device__ __forceinline__ void dudy_at_walls(const double* u, int N, int TT, int t, double scale, double& at_b, double& at_a) {
    const int Nb = N - 1;
    double de = 0.0, dod = 0.0;  // running d[n+2] for even / odd n
    double sb = 0.0, sa = 0.0;
    for (int n = Nb; n >= 0; --n) {
        double& run = (n & 1) ? dod : de;
        if (n + 1 <= Nb) run = run + scale * (n + 1) * u[(n + 1) * TT + t];
        double v = run;
        if (n == 0) v *= 0.5;
        sb += v;
        sa += v * ((n % 2 == 0) ? 1 : -1);
    }
    at_b = sb;
    at_a = sa;
}

__device__ __forceinline__ void mode_of_q(int q, const ModeGeom& g, int& kx, int& kz, long& off) {
    const int nkz = g.Kz + 1, nmx = 2 * g.Kx + 1;
    const int ml = q / nkz, mxi = g.mx0 + ml;
    kz = q - ml * nkz;
    kx = mxi <= g.Kx ? mxi : mxi - nmx;
    const int mx = kx >= 0 ? kx : g.Nx + kx;
    off = 2L * (kz + (long)(g.Nz / 2 + 1) * mx);
}

// modes of factor tile tl: q0 + m for m in [mfirst, mend); is00: the (0,0) mode's own tile (field data in tile 0, slot 0)
__device__ __forceinline__ void tile_modes(int tl, const TauData& td, int& q0, int& mfirst, int& mend, bool& is00) {
    const int ngen = td.ntiles - td.has00;
    is00 = td.has00 && tl == ngen;
    if (is00) { q0 = 0; mfirst = 0; mend = 1; return; }
    q0 = tl * td.TM;
    mfirst = (td.has00 && tl == 0) ? 1 : 0;
    mend = td.nq - q0 < td.TM ? td.nq - q0 : td.TM;
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-parallel column solver.  One warp owns one real "column" (the real or imaginary part of one mode's Chebyshev
// profile); lane l holds the E consecutive coefficients n = l*E .. l*E+E-1 in registers (E even, so the slot parity
// is the parity of n and the even/odd blocks of the quasi-tridiagonal system interleave in the slots).
// Every recurrence of the reference (derivative recurrence chebyshev.cpp:672-697, UL back/forward substitution
// bandedtridiag.cpp:258-273) is a first-order linear recurrence x_n = a_n x_{n+-2} + b_n, evaluated as a blocked
// scan: compose the lane-local affine map, combine the 32 lane maps with a log-step shuffle scan, then redo the
// local recurrence from the exact incoming value with the reference's own formula.
// Shared-memory columns are stored skewed, addr(n) = n + n/E, i.e. lane l starts at l*(E+1): consecutive lanes are
// an odd number of 8-byte words apart, which makes the per-lane contiguous accesses bank-conflict free.

template <int E>
__device__ __forceinline__ int col_addr(int n) { return n + n / E; }

template <int E>
__device__ __forceinline__ void col_load(const double* __restrict__ col, int lane, int N, double (&v)[E]) {
    const double* p = col + lane * (E + 1);
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = (lane * E + e < N) ? p[e] : 0.0;
}
template <int E>
__device__ __forceinline__ void col_store(double* __restrict__ col, int lane, int N, const double (&v)[E]) {
    double* p = col + lane * (E + 1);
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (lane * E + e < N) p[e] = v[e];
}

__device__ __forceinline__ double shfl_down_or(double v, int d, int lane, double fill) {
    const double t = __shfl_down_sync(0xffffffffu, v, d);
    return lane + d < 32 ? t : fill;
}
__device__ __forceinline__ double shfl_up_or(double v, int d, int lane, double fill) {
    const double t = __shfl_up_sync(0xffffffffu, v, d);
    return lane >= d ? t : fill;
}

// d = du/dy (chebyshev.cpp:672-697): d_n = sum_{m > n, m-n odd} (scale*m) u_m, d_0 *= 1/2.  u, d in lane registers.
template <int E>
__device__ __forceinline__ void col_deriv(const double (&u)[E], double (&d)[E], double scale, int lane) {
    double c[E];
    double tot[2] = {0.0, 0.0};  // lane totals of the even-m / odd-m terms
    const double n0d = (double)(lane * E);
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        c[e] = scale * (n0d + (double)e) * u[e];
        tot[e & 1] = tot[e & 1] + c[e];
    }
    // inclusive suffix sum over lanes, then shift to exclusive
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const double t0 = __shfl_down_sync(0xffffffffu, tot[0], s), t1 = __shfl_down_sync(0xffffffffu, tot[1], s);
        if (lane + s < 32) {
            tot[0] += t0;
            tot[1] += t1;
        }
    }
    double run[2];
    run[0] = shfl_down_or(tot[0], 1, lane, 0.0);
    run[1] = shfl_down_or(tot[1], 1, lane, 0.0);
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        d[e] = run[(e & 1) ^ 1];
        run[e & 1] = run[e & 1] + c[e];
    }
    if (lane == 0) d[0] *= 0.5;
}

// Solve  A x = B r  with boundary-row values bc (both parity blocks of one Helmholtz operator, helmholtz.cpp:79-95):
//   g_n = B_lo r_{n-2} + B_dg r_n + B_up r_{n+2}  (n >= 2)
//   back substitution   x_n = g_n - up_n x_{n+2}         n = nl .. par+2          (bandedtridiag.cpp:258-262)
//   bordered row        x_par = (bc - sum band_n x_n) / diag0                      (:263-268; diag0 in slot `par` of inv)
//   forward elimination x_n = (x_n - lo_n x_{n-2}) inv_n  n = par+2 .. nl          (:269-273), lo_n = A_lo(n, lambda)
// r (in) and x (out) are lane registers; up/inv/band are skewed shared-memory columns (SKEW) or plain rows in global
// memory, bt the B rows (three skewed shared-memory columns), lo_n = -(lambda * B_lo(n)) costs one multiply.  If wall != nullptr the
// sums  S_p = sum_{n = p mod 2} n^2 x_n  are returned in wall[0..1] (all lanes).
template <int E, bool SKEW>
__device__ __forceinline__ void col_solve(const double (&r)[E], double (&x)[E], const double* __restrict__ up,
                                          const double* __restrict__ inv, const double* __restrict__ band,
                                          const double lam, const double* __restrict__ bt, const int N, const int lane,
                                          const double bc0, const double bc1, double* wall) {
    const int Nb = N - 1;
    const int n0 = lane * E, a0 = SKEW ? lane * (E + 1) : lane * E;
    const int btl = lane * (E + 1), btp = tau_col_pitch(N, E);  // B rows: three skewed shared-memory columns
    double g[E], l[E];
    {
        // neighbours two rows away (other lanes at the block edges)
        double lo2[2], hi2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            lo2[e] = shfl_up_or(r[E - 2 + e], 1, lane, 0.0);
            hi2[e] = shfl_down_or(r[e], 1, lane, 0.0);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int n = n0 + e;
            const double rm = e >= 2 ? r[e >= 2 ? e - 2 : 0] : lo2[e & 1];
            const double rp = e < E - 2 ? r[e < E - 2 ? e + 2 : 0] : hi2[e & 1];
            double v = 0.0, lv = 0.0;
            if (n >= 2 && n < N) {
                const double blo = bt[btl + e];
                v = blo * rm + bt[btp + btl + e] * r[e];
                v += bt[2 * btp + btl + e] * rp;
                lv = A_lo_from_B(blo, lam);
            }
            g[e] = v;
            l[e] = lv;
        }
    }
    // ---- back substitution
    double u[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int n = n0 + e;
        u[e] = (n >= 2 && n + 2 <= Nb) ? up[a0 + e] : 0.0;
    }
    {
        double A[2] = {1.0, 1.0}, B[2] = {0.0, 0.0};
#pragma unroll
        for (int e = E - 1; e >= 0; --e) {
            const int p = e & 1;
            B[p] = g[e] - u[e] * B[p];
            A[p] = -(u[e] * A[p]);
        }
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const double A2 = __shfl_down_sync(0xffffffffu, A[p], s), B2 = __shfl_down_sync(0xffffffffu, B[p], s);
                if (lane + s < 32) {
                    B[p] = A[p] * B2 + B[p];
                    A[p] = A[p] * A2;
                }
            }
        }
        double xin[2];
        xin[0] = shfl_down_or(B[0], 1, lane, 0.0);
        xin[1] = shfl_down_or(B[1], 1, lane, 0.0);
#pragma unroll
        for (int e = E - 1; e >= 0; --e) {
            const int p = e & 1;
            x[e] = g[e] - u[e] * xin[p];
            xin[p] = x[e];
        }
    }
    // ---- bordered row
    double xpar[2];
    {
        double s[2] = {0.0, 0.0};
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int n = n0 + e;
            if (n >= 2 && n < N) s[e & 1] += band[a0 + e] * x[e];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
            s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
        }
        xpar[0] = (bc0 - s[0]) / inv[0];
        xpar[1] = (bc1 - s[1]) / inv[1];
    }
    // ---- forward elimination
    {
        double iv[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int n = n0 + e;
            iv[e] = (n >= 2 && n < N) ? inv[a0 + e] : 0.0;
        }
        double A[2] = {1.0, 1.0}, B[2] = {0.0, 0.0};
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int p = e & 1;
            if (lane == 0 && e < 2) {
                A[p] = 0.0;
                B[p] = xpar[p];
            } else {
                B[p] = (x[e] - l[e] * B[p]) * iv[e];
                A[p] = -(l[e] * A[p]) * iv[e];
            }
        }
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const double A2 = __shfl_up_sync(0xffffffffu, A[p], s), B2 = __shfl_up_sync(0xffffffffu, B[p], s);
                if (lane >= s) {
                    B[p] = A[p] * B2 + B[p];
                    A[p] = A[p] * A2;
                }
            }
        }
        double vin[2];
        vin[0] = shfl_up_or(B[0], 1, lane, 0.0);
        vin[1] = shfl_up_or(B[1], 1, lane, 0.0);
        double ws[2] = {0.0, 0.0};
        const double n0d = (double)n0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int p = e & 1;
            double v;
            if (lane == 0 && e < 2) v = xpar[p];
            else v = (x[e] - l[e] * vin[p]) * iv[e];
            x[e] = v;
            vin[p] = v;
            if (wall) ws[p] += (n0d + (double)e) * (n0d + (double)e) * v;
        }
        if (wall) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ws[0] += __shfl_xor_sync(0xffffffffu, ws[0], o);
                ws[1] += __shfl_xor_sync(0xffffffffu, ws[1], o);
            }
            wall[0] = ws[0];
            wall[1] = ws[1];
        }
    }
}

}  // namespace

void tau_btab_host(int N, double* tab) {
    const int Nb = N - 1;
    for (int n = 0; n < N; ++n) {
        tab[n] = n >= 2 ? B_lo(n, Nb) : 0.0;
        tab[N + n] = n >= 2 ? B_dg(n, Nb) : 0.0;
        tab[2 * N + n] = n >= 2 ? B_up(n, Nb) : 0.0;
    }
}

// =================================================================================================== setup
// grid = ntiles; CTA = one tile. Real profiles: smem arrays [n][TM].
__global__ void __launch_bounds__(TAU_SETUP_THREADS) tau_setup_kernel(const TauData td, const ModeGeom g, const double lambda_t) {
    const int N = td.N, Nb = N - 1, TM = td.TM;
    const int tid = threadIdx.x, NT = TAU_SETUP_THREADS;
    const int tl = blockIdx.x;
    int q0, mfirst, mend;
    bool is00_;
    tile_modes(tl, td, q0, mfirst, mend, is00_);   // masked slots are set up as harmless kx = kz = 0 modes
    double* A1 = dyn_smem<double>();
    double* A2 = A1 + (size_t)N * TM;
    double* A3 = A2 + (size_t)N * TM;
    double* s_lamP = A3 + (size_t)N * TM;
    double* s_lamV = s_lamP + TM;
    double* s_w = s_lamV + TM;  // [8][TM]: Ab, Ca, Bb, Da, dplus, dminus, dP0dy_Nb1, spare

    const double scale = 4.0 / (td.b - td.a);
    const double nusP = 1.0 / (((td.b - td.a) / 2) * ((td.b - td.a) / 2));
    const double nusV = td.nu / (((td.b - td.a) / 2) * ((td.b - td.a) / 2));

    double* upP = td.tile_arr(tl, TAR_UPP); double* invP = td.tile_arr(tl, TAR_INVP); double* bandP = td.tile_arr(tl, TAR_BANDP);
    double* upV = td.tile_arr(tl, TAR_UPV); double* invV = td.tile_arr(tl, TAR_INVV); double* bandV = td.tile_arr(tl, TAR_BANDV);
    double* gPp = td.tile_arr(tl, TAR_PP); double* gvp = td.tile_arr(tl, TAR_VP); double* gPm = td.tile_arr(tl, TAR_PM);
    double* gvm = td.tile_arr(tl, TAR_VM); double* gP0 = td.tile_arr(tl, TAR_P0); double* gv0 = td.tile_arr(tl, TAR_V0);

    if (tid < TM) {
        const int q = q0 + tid;
        int kx = 0, kz = 0;
        long off;
        if (tid >= mfirst && tid < mend) mode_of_q(q, g, kx, kz, off);
        const double kxL = kx / g.Lx, kzL = kz / g.Lz;
        const double kappa2 = 4 * (PI * PI) * (kxL * kxL + kzL * kzL);
        const double c = 4.0 * (PI * PI) * td.nu;
        const double lamV = lambda_t + c * (kxL * kxL + kzL * kzL);
        s_lamP[tid] = kappa2;
        s_lamV[tid] = lamV;
        td.tile_sc(tl, TSC_LAMP)[tid] = kappa2;
        td.tile_sc(tl, TSC_LAMV)[tid] = lamV;
        td.tile_sc(tl, TSC_KXX)[tid] = 2 * PI * kx / g.Lx;
        td.tile_sc(tl, TSC_KZZ)[tid] = 2 * PI * kz / g.Lz;
    }
    __syncthreads();

    // ---- UL factorisation of Ae, Ao for both Helmholtz operators (bandedtridiag.cpp:212-229)
    for (int c = tid; c < 4 * TM; c += NT) {
        const int m = c % TM, par = (c / TM) & 1, h = c / (2 * TM);
        const double lam = h ? s_lamV[m] : s_lamP[m];
        const double nus = h ? nusV : nusP;
        double* up = h ? upV : upP;
        double* inv = h ? invV : invP;
        double* band = h ? bandV : bandP;
        const int nl = par ? Nb - 1 : Nb;
        double dgk = A_dg(nl, Nb, lam, nus);
        double bandk = 1.0;
        for (int n = nl; n >= par + 4; n -= 2) {
            const double Akk = dgk;
            inv[m * N + n] = 1.0 / Akk;
            const double w = A_lo(n, Nb, lam);
            const double upm = A_up(n - 2, Nb, lam) / Akk;
            up[m * N + n - 2] = upm;
            const double dprev = A_dg(n - 2, Nb, lam, nus) - w * upm;
            const double bk = bandk / Akk;
            band[m * N + n] = bk;
            bandk = 1.0 - w * bk;
            dgk = dprev;
        }
        const int n1 = par + 2;
        inv[m * N + n1] = 1.0 / dgk;
        const double b1 = bandk / dgk;
        band[m * N + n1] = b1;
        inv[m * N + par] = 1.0 - A_lo(n1, Nb, lam) * b1;  // diag(0) == band(0)
    }
    __syncthreads();

    // ---- P+-, v+- and the influence matrix (tausolver.cpp:117-147)
    for (int pm = 0; pm < 2; ++pm) {
        for (int idx = tid; idx < N * TM; idx += NT) {
            const int n = idx / TM;
            // P(a)=0,P(b)=1 -> g0 = (ub+ua)/2 = .5, g1 = (ub-ua)/2 = .5 ; P(a)=1,P(b)=0 -> .5, -.5
            A1[idx] = n == 0 ? 0.5 : (n == 1 ? (pm == 0 ? 0.5 : -0.5) : 0.0);
        }
        __syncthreads();
        for (int c = tid; c < 2 * TM; c += NT) {
            const int m = c % TM, par = c / TM;
            ul_solve_chain(A1, N, TM, m, par, upP, invP, bandP, TM, m, s_lamP[m]);
        }
        __syncthreads();
        for (int c = tid; c < 2 * TM; c += NT) diff_chain(A1, A2, N, TM, c % TM, c / TM, scale, nullptr);
        __syncthreads();
        bmul(A2, A3, N, TM, 0.0, 0.0, tid, NT);
        __syncthreads();
        for (int c = tid; c < 2 * TM; c += NT) {
            const int m = c % TM, par = c / TM;
            ul_solve_chain(A3, N, TM, m, par, upV, invV, bandV, TM, m, s_lamV[m]);
        }
        __syncthreads();
        double* gP = pm == 0 ? gPp : gPm;
        double* gv = pm == 0 ? gvp : gvm;
        for (int idx = tid; idx < N * TM; idx += NT) {
            const int n = idx / TM, m = idx - n * TM;
            gP[m * N + n] = A1[idx];
            gv[m * N + n] = A3[idx];
        }
        if (tid < TM) {
            double vb, va;
            dudy_at_walls(A3, N, TM, tid, scale, vb, va);
            s_w[(2 * pm) * TM + tid] = vb;      // A (plus) / B (minus)
            s_w[(2 * pm + 1) * TM + tid] = va;  // C (plus) / D (minus)
        }
        __syncthreads();
    }
    if (tid < TM) {
        const double A = s_w[0 * TM + tid], C = s_w[1 * TM + tid], B = s_w[2 * TM + tid], D = s_w[3 * TM + tid];
        const double disc = A * D - B * C;
        td.tile_sc(tl, TSC_I00)[tid] = D / disc;
        td.tile_sc(tl, TSC_I01)[tid] = -B / disc;
        td.tile_sc(tl, TSC_I10)[tid] = -C / disc;
        td.tile_sc(tl, TSC_I11)[tid] = A / disc;
    }
    // ---- tau-correction basis P0, v0, sigma0 (tausolver.cpp:149-175)
    {
        const double cc = 2 / (td.b - td.a);
        for (int idx = tid; idx < N * TM; idx += NT) {
            const int i = idx / TM;
            int nf;
            if (i == 0) nf = Nb - 1;
            else if (i == Nb) nf = 0;
            else if (i % 2 == 0) nf = 2 * (Nb - 1);
            else nf = 2 * Nb;
            A1[idx] = cc * nf;
        }
    }
    __syncthreads();
    bmul(A1, A2, N, TM, 0.0, 0.0, tid, NT);
    __syncthreads();
    for (int c = tid; c < 2 * TM; c += NT) {
        const int m = c % TM, par = c / TM;
        ul_solve_chain(A2, N, TM, m, par, upP, invP, bandP, TM, m, s_lamP[m]);
    }
    __syncthreads();  // A2 = P0 (before influence correction)
    for (int c = tid; c < 2 * TM; c += NT) diff_chain(A2, A3, N, TM, c % TM, c / TM, scale, nullptr);
    __syncthreads();  // A3 = dP0/dy
    if (tid < TM) s_w[6 * TM + tid] = A3[(Nb - 1) * TM + tid];
    bmul(A3, A1, N, TM, 0.0, 0.0, tid, NT);
    __syncthreads();
    for (int c = tid; c < 2 * TM; c += NT) {
        const int m = c % TM, par = c / TM;
        ul_solve_chain(A1, N, TM, m, par, upV, invV, bandV, TM, m, s_lamV[m]);
    }
    __syncthreads();  // A1 = v0
    if (tid < TM) {
        double vb, va;
        dudy_at_walls(A1, N, TM, tid, scale, vb, va);
        s_w[4 * TM + tid] = -td.tile_sc(tl, TSC_I00)[tid] * vb - td.tile_sc(tl, TSC_I01)[tid] * va;
        s_w[5 * TM + tid] = -td.tile_sc(tl, TSC_I10)[tid] * vb - td.tile_sc(tl, TSC_I11)[tid] * va;
    }
    __syncthreads();
    for (int idx = tid; idx < N * TM; idx += NT) {
        const int n = idx / TM, m = idx - n * TM, gi = m * N + n;
        const double dp = s_w[4 * TM + m], dm = s_w[5 * TM + m];
        const double P0 = A2[idx] + (dp * gPp[gi] + dm * gPm[gi]);
        const double v0 = A1[idx] + (dp * gvp[gi] + dm * gvm[gi]);
        A2[idx] = P0;
        A1[idx] = v0;
        gP0[gi] = P0;
        gv0[gi] = v0;
    }
    __syncthreads();
    if (tid < TM) {
        const double lam = s_lamV[tid];
        // v0'' has zero coefficients at Nb-1 and Nb, dP0/dy[Nb] == 0 (chebyshev.cpp:688-689)
        td.tile_sc(tl, TSC_S0NB1)[tid] = lam * A1[(Nb - 1) * TM + tid] + s_w[6 * TM + tid] - td.nu * 0.0;
        td.tile_sc(tl, TSC_S0NB)[tid] = lam * A1[Nb * TM + tid] + 0.0 - td.nu * 0.0;
    }
}

// =================================================================================================== solve
// grid = ntiles ; CTA 0 handles the (0,0) mode (real parts, mean-flow constraint).
// Shared memory: profile columns D[4][TT][NP] (Rx, Ry, Rz, P; column t = 2*mode + re/im, skewed in n), factor columns
// F[3][TM][NP] (up, inv, band of the operator in use), the B rows [3][NP], scalars.
template <int E, int NTERMS>
__global__ void __launch_bounds__(TAU_THREADS, (E <= 10 ? 2 : 1)) tau_solve_kernel(const TauSolveParams p) {
    const TauData& td = p.td;
    const int N = td.N, Nb = N - 1, TM = td.TM, TT = 2 * TM;
    const int tid = threadIdx.x, NT = TAU_THREADS, lane = tid & 31, warp = tid >> 5, NW = TAU_THREADS / 32;
    const int tl = blockIdx.x;
    const double scale = 4.0 / (td.b - td.a);
    const long rs_ser = (long)p.g.Nx * (2 * (p.g.Nz / 2 + 1));  // row (ny) stride in doubles
    const long cs_ser = rs_ser * p.g.Ny;                        // component stride
    const int NP = tau_col_pitch(N, E);                     // column pitch (doubles)
    const int AS = TT * NP;                                 // doubles per profile array
    const int FS = TM * NP;                                 // doubles per factor array
    const int NM = N * TM;

    double* Rx = dyn_smem<double>();
    double* Ry = Rx + AS; double* Rz = Ry + AS; double* Pq = Rz + AS;
    double* Fup = Pq + AS; double* Finv = Fup + FS; double* Fband = Finv + FS;   // velocity-operator factors
    double* s_sc = Fband + FS;                   // [TSC_COUNT][TM]
    double* s_w = s_sc + TSC_COUNT * TM;         // [8][TT]
    long* s_off = reinterpret_cast<long*>(s_w + 8 * TT);  // [TM]
    double* bt = reinterpret_cast<double*>(s_off + TM);   // B rows [3][NP] (mode independent), skewed like the columns
    int q0, mfirst, mend;
    bool is00;
    tile_modes(tl, td, q0, mfirst, mend, is00);
    if (mfirst >= mend) return;        // tile 0 holding nothing but the masked (0,0) slot
    const int ftile = is00 ? 0 : tl;   // tile of the FIELD data (tile-major layout): the (0,0) mode lives in slot 0 of tile 0

    if (tid < TM) {
        long off = -1;
        if (tid >= mfirst && tid < mend) { int kx, kz; mode_of_q(q0 + tid, p.g, kx, kz, off); }
        s_off[tid] = off;
    }
    for (int i = tid; i < TSC_COUNT * TM; i += NT) s_sc[i] = td.tile_sc(tl, 0)[i];
    {   // velocity-operator factors (upV, invV, bandV contiguous, [m][n]) -> skewed shared columns, asynchronously: they
        // are first needed after the pressure solve.  One (array, mode) row per warp at a time, no index divisions.
        const double* fsrc = td.tile_arr(tl, TAR_UPV);
        for (int row = warp; row < 3 * TM; row += NW) {
            const double* src = fsrc + (size_t)row * N;
            double* dst = Fup + row * NP;
            for (int n = lane; n < N; n += 32) cp_async8(dst + col_addr<E>(n), src + n);
        }
        {
            const double* src = td.btab();
            for (int i = tid; i < 3 * N; i += NT) {
                const int r = i >= 2 * N ? 2 : (i >= N ? 1 : 0), n = i - r * N;
                cp_async8(bt + r * NP + col_addr<E>(n), src + i);
            }
        }
        cp_async_commit();
        if (p.tile_layout && p.prefetch_terms) {
            // tile-major history fields: this CTA's 3 x N x TM block of every term is contiguous -- request all of it now
            const size_t bytes = (size_t)3 * NM * 2 * sizeof(double);
            const int nterms = NTERMS > 0 ? NTERMS : p.nterms;
            // (fetching for the CTA some tiles ahead instead was measured: 2.3-2.6 ms against 1.9 ms, the blocks do not survive
            // in L2 until they are used)
            for (int j = 0; j < nterms; ++j) {
                const char* blk = reinterpret_cast<const char*>(p.term[j] + (size_t)ftile * 3 * NM * 2);
                for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)NT * 128) prefetch_l2(blk + o);
            }
        }
        // pull the pressure-operator factors (read straight from global by the solving warps) towards L2
        // ... and the six correction profiles, used last
        const char* blk = reinterpret_cast<const char*>(td.tile_arr(tl, TAR_UPP));
        const size_t bytes = (size_t)3 * NM * sizeof(double);
        for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)NT * 128) {
            prefetch_l2(blk + o);
            prefetch_l2(blk + 2 * bytes + o);
            prefetch_l2(blk + 3 * bytes + o);
        }
    }
    __syncthreads();

    // ---- S1: right-hand side = linear combination of history fields (dnsalgo.cpp:217-224), complex elements.
    // Thread -> (mode m, row slot j): rows n = j, j + NT/TM, ..; all loads of two rows are issued before their use.
    {
        const int m = tid % TM, j0 = tid / TM, JS = NT / TM;   // NT % TM == 0 is guaranteed by the launcher
        long off = s_off[m];
        long rs = rs_ser, cs = cs_ser;
        if (p.tile_layout && off >= 0) { off = ((long)ftile * 3 * N * TM + m) * 2; cs = (long)N * TM * 2; rs = TM * 2; }
        const int nterms = NTERMS > 0 ? NTERMS : p.nterms;
        if (off >= 0 && j0 < JS) {
            for (int comp = 0; comp < 3; ++comp) {
                const long gbase = comp * cs + off;
                double* dre = Rx + comp * AS + (2 * m) * NP;  // Rx,Ry,Rz contiguous
                for (int n0 = j0; n0 < N; n0 += 2 * JS) {
                    const int n1 = n0 + JS;
                    const bool has1 = n1 < N;
                    double2 v0[NTERMS > 0 ? NTERMS : TAU_MAXTERMS], v1[NTERMS > 0 ? NTERMS : TAU_MAXTERMS];
#pragma unroll
                    for (int j = 0; j < (NTERMS > 0 ? NTERMS : TAU_MAXTERMS); ++j) {
                        if (j < nterms) {
                            v0[j] = *reinterpret_cast<const double2*>(p.term[j] + gbase + n0 * rs);
                            if (has1) v1[j] = *reinterpret_cast<const double2*>(p.term[j] + gbase + n1 * rs);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h == 1 && !has1) break;
                        const int n = h ? n1 : n0;
                        double are = 0.0, aim = 0.0;
#pragma unroll
                        for (int j = 0; j < (NTERMS > 0 ? NTERMS : TAU_MAXTERMS); ++j)
                            if (j < nterms) {
                                const double2 v = h ? v1[j] : v0[j];
                                are += p.coef[j] * v.x;
                                aim += p.coef[j] * v.y;
                            }
                        if (is00) {
                            // mean mode: base-flow diffusion and the imposed pressure gradient (nse.cpp:512-530); real parts only
                            if (comp == 0 && p.Ubaseyy) are += td.nu * p.Ubaseyy[n];
                            if (comp == 2 && p.Wbaseyy) are += td.nu * p.Wbaseyy[n];
                            if (p.constraint == 0 && n == 0) {
                                if (comp == 0) are -= p.dPdxRef;
                                if (comp == 2) are -= p.dPdzRef;
                            }
                            aim = 0.0;
                        }
                        const int a = col_addr<E>(n);
                        dre[a] = are;
                        dre[NP + a] = aim;
                    }
                }
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- S2: pressure Helmholtz  P'' - kappa^2 P = dRy/dy + i (kxx Rx + kzz Rz), P(+-1) = 0  (tausolver.cpp:357-366, 193-201)
    // ---- S3: v particular solution  nu v'' - lambda v = P' - Ry, v(+-1) = 0, in place in Ry  (tausolver.cpp:203-210)
    // Both act on column `col` only, so a warp carries P from one to the other in registers.  The pressure factors are
    // used by the two columns of a mode only and come straight from global memory ([m][n] rows, prefetched to L2).
    for (int col = warp; col < TT; col += NW) {
        const int m = col >> 1, ri = col & 1;
        if (m < mfirst || m >= mend) continue;
        double r[E], x[E];
        {
            double y[E], d[E], ox[E], oz[E];
            col_load<E>(Ry + col * NP, lane, N, y);
            col_deriv<E>(y, d, scale, lane);
            col_load<E>(Rx + (col ^ 1) * NP, lane, N, ox);
            col_load<E>(Rz + (col ^ 1) * NP, lane, N, oz);
            const double kxx = s_sc[TSC_KXX * TM + m], kzz = s_sc[TSC_KZZ * TM + m];
#pragma unroll
            for (int e = 0; e < E; ++e) r[e] = ri ? d[e] + (kxx * ox[e] + kzz * oz[e]) : d[e] - (kxx * ox[e] + kzz * oz[e]);
        }
        const double* fP = td.tile_arr(tl, TAR_UPP) + (size_t)m * N;
        col_solve<E, false>(r, x, fP, fP + NM, fP + 2 * NM, s_sc[TSC_LAMP * TM + m], bt, N, lane, 0.0, 0.0, nullptr);
        col_store<E>(Pq + col * NP, lane, N, x);
        if (is00) continue;
        {
            double d[E], y[E];
            col_deriv<E>(x, d, scale, lane);
            col_load<E>(Ry + col * NP, lane, N, y);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                r[e] = d[e] - y[e];
                const int n = lane * E + e;  // Ry[Nb], Ry[Nb-1] are needed again by the tau correction
                if (n == Nb) s_w[4 * TT + col] = y[e];
                if (n == Nb - 1) s_w[5 * TT + col] = y[e];
            }
        }
        double wall[2];
        col_solve<E, true>(r, x, Fup + m * NP, Finv + m * NP, Fband + m * NP, s_sc[TSC_LAMV * TM + m], bt, N, lane, 0.0, 0.0, wall);
        col_store<E>(Ry + col * NP, lane, N, x);
        if (lane == 0) { s_w[6 * TT + col] = wall[0]; s_w[7 * TT + col] = wall[1]; }
    }
    __syncthreads();

    if (!is00) {
        // ---- S4: influence-matrix (tausolver.cpp:178-191) and tau (tausolver.cpp:215-244) correction amplitudes.
        // v'(b) = (2/L) sum n^2 v_n, v'(a) = (2/L) sum (-1)^(n+1) n^2 v_n  (closed form of eval_b/eval_a of diff(v))
        const double* gPp = td.tile_arr(tl, TAR_PP); const double* gvp = td.tile_arr(tl, TAR_VP);
        const double* gPm = td.tile_arr(tl, TAR_PM); const double* gvm = td.tile_arr(tl, TAR_VM);
        const double* gP0 = td.tile_arr(tl, TAR_P0); const double* gv0 = td.tile_arr(tl, TAR_V0);
        if (tid < TT) {
            const int m = tid >> 1;
            if (m >= mfirst && m < mend) {
                const double Se = s_w[6 * TT + tid], So = s_w[7 * TT + tid];
                const double vb = 0.5 * scale * (So + Se), va = 0.5 * scale * (So - Se);
                const double dp = -s_sc[TSC_I00 * TM + m] * vb - s_sc[TSC_I01 * TM + m] * va;
                const double dm = -s_sc[TSC_I10 * TM + m] * vb - s_sc[TSC_I11 * TM + m] * va;
                s_w[tid] = dp;
                s_w[TT + tid] = dm;
                if (p.taucorr) {
                    const double lam = s_sc[TSC_LAMV * TM + m];
                    const int iNb = m * N + Nb, iNb1 = m * N + Nb - 1;
                    const int aNb = tid * NP + col_addr<E>(Nb), aNb1 = tid * NP + col_addr<E>(Nb - 1);
                    const double vNb = Ry[aNb] + (dp * gvp[iNb] + dm * gvm[iNb]);
                    const double vNb1 = Ry[aNb1] + (dp * gvp[iNb1] + dm * gvm[iNb1]);
                    const double pNb = Pq[aNb] + (dp * gPp[iNb] + dm * gPm[iNb]);
                    // v'' has zero coefficients at Nb-1, Nb; P'[Nb] = 0, P'[Nb-1] = scale Nb P[Nb]
                    const double s1nb = lam * vNb - s_w[4 * TT + tid];
                    double s1nb1 = lam * vNb1 - s_w[5 * TT + tid];
                    s1nb1 += scale * Nb * pNb;
                    s_w[2 * TT + tid] = s1nb / (1.0 - s_sc[TSC_S0NB * TM + m]);     // sigmaNb
                    s_w[3 * TT + tid] = s1nb1 / (1.0 - s_sc[TSC_S0NB1 * TM + m]);   // sigmaNb1
                }
            }
        }
        __syncthreads();
        // one (mode, half of the rows) item per warp: rows n = lane + 32*half + 64*k, four of them in flight
        for (int item = warp; item < 2 * TM; item += NW) {
            const int m = item >> 1;
            if (m < mfirst || m >= mend) continue;
            const double dpr = s_w[2 * m], dpi = s_w[2 * m + 1], dmr = s_w[TT + 2 * m], dmi = s_w[TT + 2 * m + 1];
            const double sNbr = s_w[2 * TT + 2 * m], sNbi = s_w[2 * TT + 2 * m + 1];
            const double sNb1r = s_w[3 * TT + 2 * m], sNb1i = s_w[3 * TT + 2 * m + 1];
            const int gm = m * N;
            double* Pc = Pq + (2 * m) * NP;
            double* Vc = Ry + (2 * m) * NP;
            for (int nb = lane + 32 * (item & 1); nb < N; nb += 256) {
                double gq[4][6];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int n = nb + 64 * h;
                    if (n < N) {
                        gq[h][0] = gPp[gm + n]; gq[h][1] = gvp[gm + n]; gq[h][2] = gPm[gm + n]; gq[h][3] = gvm[gm + n];
                        if (p.taucorr) { gq[h][4] = gP0[gm + n]; gq[h][5] = gv0[gm + n]; }
                    }
                }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int n = nb + 64 * h;
                    if (n >= N) break;
                    const double pp = gq[h][0], vp = gq[h][1], pm = gq[h][2], vm = gq[h][3];
                    const int a = col_addr<E>(n);
                    double Pr = Pc[a], Pi = Pc[a + NP], Vr = Vc[a], Vi = Vc[a + NP];
                    Pr += dpr * pp + dmr * pm;
                    Pi += dpi * pp + dmi * pm;
                    Vr += dpr * vp + dmr * vm;
                    Vi += dpi * vp + dmi * vm;
                    if (p.taucorr) {
                        const double p0 = gq[h][4], v0 = gq[h][5];
                        const bool ev = (n & 1) == 0;
                        Pr += (ev ? sNb1r : sNbr) * p0;
                        Pi += (ev ? sNb1i : sNbi) * p0;
                        Vr += (ev ? sNbr : sNb1r) * v0;
                        Vi += (ev ? sNbi : sNb1i) * v0;
                    }
                    Pc[a] = Pr; Pc[a + NP] = Pi;
                    Vc[a] = Vr; Vc[a + NP] = Vi;
                }
            }
        }
        __syncthreads();
    } else if (p.constraint == 1) {
        // mean mode with the bulk-velocity constraint (helmholtz.cpp:158-213): keep the right-hand sides (v = 0 anyway)
        for (int n = tid; n < N; n += NT) {
            const int a = col_addr<E>(n);
            Ry[a] = Rx[a];
            Ry[NP + a] = Rz[a];
        }
        __syncthreads();
    }

    // ---- S5: u, w from the x/z momentum equations  nu u'' - lambda u = i kxx P - Rx, in place in Rx, Rz (tausolver.cpp:368-384).
    // Mean mode + bulk velocity: the imaginary column carries the solve with right-hand side nu*T0 instead ("uc").
    const bool bulk00 = is00 && p.constraint == 1;
    for (int c2 = warp; c2 < 2 * TT; c2 += NW) {
        const int which = c2 / TT, col = c2 - which * TT, m = col >> 1, ri = col & 1;
        if (m < mfirst || m >= mend) continue;
        double* R = which ? Rz : Rx;
        double r[E], x[E];
        {
            double po[E], rr[E];
            col_load<E>(Pq + (col ^ 1) * NP, lane, N, po);
            col_load<E>(R + col * NP, lane, N, rr);
            const double k = s_sc[(which ? TSC_KZZ : TSC_KXX) * TM + m];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                r[e] = ri ? k * po[e] - rr[e] : -k * po[e] - rr[e];
                if (bulk00 && ri) r[e] = (lane * E + e == 0) ? td.nu : 0.0;
            }
        }
        col_solve<E, true>(r, x, Fup + m * NP, Finv + m * NP, Fband + m * NP, s_sc[TSC_LAMV * TM + m], bt, N, lane, 0.0, 0.0, nullptr);
        col_store<E>(R + col * NP, lane, N, x);
    }
    __syncthreads();
    if (bulk00) {
        if (warp < 2) {
            // mean(u) = u_0 - sum_{n even >= 2} u_n/(n^2-1)  (chebyshev.cpp:505-512) of the solution (re) and of uc (im)
            const double* R = warp ? Rz : Rx;
            double ua[E], uc[E];
            col_load<E>(R, lane, N, ua);
            col_load<E>(R + NP, lane, N, uc);
            double sa = 0.0, sc = 0.0;
#pragma unroll
            for (int e = 0; e < E; e += 2) {
                const int n = lane * E + e;
                if (n >= 2 && n < N) { sa += ua[e] / (double)(n * n - 1); sc += uc[e] / (double)(n * n - 1); }
            }
            sa = warp_sum(sa); sc = warp_sum(sc);
            if (lane == 0) {
                const double uam = ua[0] - sa, ucm = uc[0] - sc;
                const double target = warp == 0 ? p.umean_target : p.wmean_target;
                const double mu = td.nu * (target - uam) / ucm;
                s_w[warp] = mu;
                if (p.dPd_act) p.dPd_act[warp] = mu;
            }
        }
        __syncthreads();
        if (warp < 2) {
            double* R = warp ? Rz : Rx;
            double r[E], x[E], sv[E];
            col_load<E>(Ry + warp * NP, lane, N, sv);  // Ry.re = original Rx, Ry.im = original Rz
            const double mu = s_w[warp];
#pragma unroll
            for (int e = 0; e < E; ++e) r[e] = -sv[e] + ((lane * E + e == 0) ? mu : 0.0);
            col_solve<E, true>(r, x, Fup, Finv, Fband, s_sc[TSC_LAMV * TM], bt, N, lane, 0.0, 0.0, nullptr);
            col_store<E>(R, lane, N, x);
#pragma unroll
            for (int e = 0; e < E; ++e) x[e] = 0.0;
            col_store<E>(R + NP, lane, N, x);
        }
        __syncthreads();
    }

    // ---- S6: scatter (nse.cpp:566-572)
    const int tms = __ffs(TM) - 1;  // TM divides TAU_THREADS = 256: a power of two
    for (int e = tid; e < NM; e += NT) {
        const int n = e >> tms, m = e & (TM - 1);
        long off = s_off[m];
        if (off < 0) continue;
        long rs = rs_ser, cs = cs_ser;
        if (p.tile_layout) { off = ((long)ftile * 3 * N * TM + m) * 2; cs = (long)N * TM * 2; rs = TM * 2; }
        const long go = n * rs + off;
        const int a = (2 * m) * NP + col_addr<E>(n);
        double2 V = make_double2(Ry[a], Ry[a + NP]);
        if (is00) V = make_double2(0.0, 0.0);
        *reinterpret_cast<double2*>(&p.uout[go]) = make_double2(Rx[a], Rx[a + NP]);
        *reinterpret_cast<double2*>(&p.uout[cs + go]) = V;
        *reinterpret_cast<double2*>(&p.uout[2 * cs + go]) = make_double2(Rz[a], Rz[a + NP]);
        const long goq = p.tile_layout ? ((long)ftile * N * TM + m) * 2 + n * rs : go;
        *reinterpret_cast<double2*>(&p.qout[goq]) = make_double2(Pq[a], Pq[a + NP]);
    }
}

// =================================================================================================== linear
// NSE::linear (nse.cpp:393-477): L = nu u'' - nu kappa^2 u - grad q  per retained mode (+ mean-mode constants).
// grid = ceil(nq/TM). smem: Pk, Pyk, X, T, R as [n][t].
__global__ void __launch_bounds__(TAU_SETUP_THREADS) linear_kernel(const TauSolveParams p, const double* __restrict__ u,
                                                                   const double* __restrict__ q, double* __restrict__ L) {
    const TauData& td = p.td;
    const int N = td.N;
    const int tid = threadIdx.x, NT = TAU_SETUP_THREADS;
    const double scale = 4.0 / (td.b - td.a);
    // u, q, L in the reference layout, or all three tile-major with td.TM modes per tile (q: one component)
    const bool tiled = p.tile_layout != 0;
    const long rs = tiled ? 2L * td.TM : (long)p.g.Nx * (2 * (p.g.Nz / 2 + 1));
    const long cs = tiled ? (long)N * td.TM * 2 : rs * p.g.Ny;
    const int TM = p.TM_lin, TT = 2 * TM;
    const size_t AS = (size_t)N * TT;
    double* Pk = dyn_smem<double>();
    double* Pyk = Pk + AS; double* X = Pyk + AS; double* T = X + AS; double* R = T + AS;
    double* s_k = R + AS;  // [3][TM]: kappa2, kxx, kzz
    long* s_off = reinterpret_cast<long*>(s_k + 3 * TM);   // mode offset in a 3-component field (u, L); -1: no mode
    long* s_offq = s_off + TM;                              // ... in the 1-component field q
    __shared__ double s_shear[2];
    const int q0 = blockIdx.x * TM;
    if (tid < TM) {
        const int qq = q0 + tid;
        long off = -1, offq = -1;
        if (qq < td.nq) {
            int kx, kz;
            mode_of_q(qq, p.g, kx, kz, off);
            offq = off;
            if (tiled) {
                const long t = qq / td.TM, pos = qq % td.TM;
                off = t * 3 * N * td.TM * 2 + pos * 2;
                offq = t * N * td.TM * 2 + pos * 2;
            }
        }
        s_off[tid] = off;
        s_offq[tid] = offq;
        const int qs = qq < td.nq ? qq : 0;
        s_k[tid] = td.scq(TSC_LAMP, qs);
        s_k[TM + tid] = td.scq(TSC_KXX, qs);
        s_k[2 * TM + tid] = td.scq(TSC_KZZ, qs);
    }
    __syncthreads();
    for (int idx = tid; idx < (int)AS; idx += NT) {
        const int n = idx / TT, t = idx - n * TT;
        const long off = s_offq[t >> 1];
        Pk[idx] = off >= 0 ? q[n * rs + off + (t & 1)] : 0.0;
    }
    __syncthreads();
    for (int c = tid; c < 2 * TT; c += NT) diff_chain(Pk, Pyk, N, TT, c % TT, c / TT, scale, nullptr);
    __syncthreads();
    for (int comp = 0; comp < 3; ++comp) {
        for (int idx = tid; idx < (int)AS; idx += NT) {
            const int n = idx / TT, t = idx - n * TT;
            const long off = s_off[t >> 1];
            X[idx] = off >= 0 ? td.nu * u[comp * cs + n * rs + off + (t & 1)] : 0.0;
        }
        __syncthreads();
        for (int c = tid; c < 2 * TT; c += NT) diff_chain(X, T, N, TT, c % TT, c / TT, scale, nullptr);
        __syncthreads();
        if (td.has00 && blockIdx.x == 0 && p.constraint == 1 && comp != 1 && tid == 0) {  // q = 0 is slot 0 of block 0
            // wall shear of nu*u for the (0,0) mode, real part (t = 0): eval_b - eval_a of d(nu u)/dy
            double sb = 0.0, sa = 0.0;
            for (int n = N - 1; n >= 0; --n) { sb += T[n * TT]; sa += T[n * TT] * ((n % 2 == 0) ? 1 : -1); }
            s_shear[comp / 2] = (sb - sa) / (td.b - td.a);
        }
        for (int c = tid; c < 2 * TT; c += NT) diff_chain(T, R, N, TT, c % TT, c / TT, scale, nullptr);
        __syncthreads();
        for (int idx = tid; idx < (int)AS; idx += NT) {
            const int n = idx / TT, t = idx - n * TT, m = t >> 1;
            const long off = s_off[m];
            if (off < 0) continue;
            const double kap2 = s_k[m], kxx = s_k[TM + m], kzz = s_k[2 * TM + m];
            double g;  // component of grad q
            if (comp == 1) g = Pyk[idx];
            else {
                const double k = comp == 0 ? kxx : kzz;
                g = (t & 1) ? k * Pk[idx - 1] : -k * Pk[idx + 1];
            }
            double v = R[idx] - kap2 * X[idx] - g;
            if (td.has00 && q0 + m == 0 && (t & 1) == 0) {
                if (comp == 0 && p.Ubaseyy) v += td.nu * p.Ubaseyy[n];
                if (comp == 2 && p.Wbaseyy) v += td.nu * p.Wbaseyy[n];
                if (n == 0 && comp != 1) {
                    if (p.constraint == 0) v -= comp == 0 ? p.dPdxRef : p.dPdzRef;
                    else v -= s_shear[comp / 2] + (comp == 0 ? p.lin_base_dPdx : p.lin_base_dPdz);
                }
            }
            L[comp * cs + n * rs + off + (t & 1)] = v;
        }
        __syncthreads();
    }
}

// =================================================================================================== launchers
int tau_pick_TM(int N, int bytes_per_mode_row) {
    const size_t budget = 190 * 1024;
    int TM = 32;
    while (TM > 1 && (size_t)N * TM * bytes_per_mode_row > budget) TM >>= 1;
    return TM;
}

// lane block size of the warp-parallel column solver: smallest instantiated even E with 32*E >= N
int tau_pick_E(int N) {
    const int need = 2 * ((N + 63) / 64);
    const int avail[] = {2, 4, 6, 8, 10, 12, 16, 20};
    for (int e : avail)
        if (e >= need) return e;
    return 0;
}

static size_t solve_smem(int N, int TM) {
    const int E = tau_pick_E(N);
    if (!E) return (size_t)1 << 30;
    const size_t NP = tau_col_pitch(N, E);
    return ((size_t)11 * TM * NP + 3 * NP + TSC_COUNT * TM + 16 * TM) * sizeof(double) + TM * sizeof(long);
}

// modes per tile of the solve kernel: two CTAs per SM when the profiles are long (one streams while the other
// solves), at most 8 modes (128-byte runs of the history fields)
int tau_pick_TM_solve(int N) {
    int TM = 8;
    while (TM > 1 && (solve_smem(N, TM) > 112 * 1024 || TAU_THREADS % TM)) --TM;
    return TM;
}

int tau_setup_launch(const TauData& td, const ModeGeom& g, double lambda_t, cudaStream_t stream) {
    const size_t smem = ((size_t)3 * td.N * td.TM + 10 * td.TM) * sizeof(double);
    static size_t configured = 0;
    auto kfn = tau_setup_kernel;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(td.ntiles);
    CF_LAUNCH(kfn, grid, dim3(TAU_SETUP_THREADS), smem, stream, td, g, lambda_t);
    CF_KERNEL_CHECK();
    return 0;
}

int linear_launch(const TauSolveParams& p, const double* u, const double* q, double* L, cudaStream_t stream) {
    const int TT = 2 * p.TM_lin;
    const size_t smem = ((size_t)5 * p.td.N * TT + 3 * p.TM_lin) * sizeof(double) + 2 * p.TM_lin * sizeof(long);
    static size_t configured = 0;
    auto kfn = linear_kernel;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((p.td.nq + p.TM_lin - 1) / p.TM_lin);
    CF_LAUNCH(kfn, grid, dim3(TAU_SETUP_THREADS), smem, stream, p, u, q, L);
    CF_KERNEL_CHECK();
    return 0;
}

template <int E, int NTERMS>
static int solve_launch_en(const TauSolveParams& p, size_t smem, cudaStream_t stream) {
    static size_t configured = 0;
    auto kfn = tau_solve_kernel<E, NTERMS>;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(p.td.ntiles);
    CF_LAUNCH(kfn, grid, dim3(TAU_THREADS), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}
// term counts of the reference's steppers get their own instantiation: 2k for SBDF-k (6 = SBDF3), 4 for the
// CNAB/SMRK/CNRK substeps; anything else runs the generic (predicated) variant
template <int E>
static int solve_launch_e(const TauSolveParams& p, size_t smem, cudaStream_t stream) {
    switch (p.nterms) {
        case 1: return solve_launch_en<E, 1>(p, smem, stream);
        case 4: return solve_launch_en<E, 4>(p, smem, stream);
        case 6: return solve_launch_en<E, 6>(p, smem, stream);
        default: return solve_launch_en<E, 0>(p, smem, stream);
    }
}

int tau_solve_launch(const TauSolveParams& p, cudaStream_t stream) {
    const size_t smem = solve_smem(p.td.N, p.td.TM);
    if (smem > 227 * 1024 || TAU_THREADS % p.td.TM) {
        set_last_error("tau_solve: Ny too large for the shared-memory tile");
        return 1;
    }
    switch (tau_pick_E(p.td.N)) {
        case 2: return solve_launch_e<2>(p, smem, stream);
        case 4: return solve_launch_e<4>(p, smem, stream);
        case 6: return solve_launch_e<6>(p, smem, stream);
        case 8: return solve_launch_e<8>(p, smem, stream);
        case 10: return solve_launch_e<10>(p, smem, stream);
        case 12: return solve_launch_e<12>(p, smem, stream);
        case 16: return solve_launch_e<16>(p, smem, stream);
        case 20: return solve_launch_e<20>(p, smem, stream);
    }
    set_last_error("tau_solve: unsupported Ny");
    return 1;
}

}  // namespace cfgpu
