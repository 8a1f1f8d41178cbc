// Batched tau-solver kernels (setup = factorisation + influence matrix, solve = fused RHS + Kleiser-Schumann).
// See tau.cuh for what is replaced and for the tile-major storage of the per-mode factors.
//
// Solve kernel: one CTA owns one tile of TM consecutive retained modes.  Their Chebyshev profiles live in shared
// memory as [n][t] (t = 2*mode + re/im, fastest) together with the tile's UL factors [n][mode], all staged with
// coalesced loads by the whole CTA.  The bordered-tridiagonal solves are sequential recurrences in n; a "chain"
// thread owns one (mode, parity) and carries the real and imaginary parts together (same factors, 2-way ILP),
// with the recurrence state in registers and every operand coming from shared memory -- no global-memory access
// sits on a dependent path.  The right-hand side of each Helmholtz problem (Chebyshev derivative recurrence,
// i k P - R combinations) and the C&H "B" row multiply (helmholtz.cpp:81-85) are fused into the backward sweep,
// so each solve is two sweeps over its N/2 rows and works in place.  Data-parallel stages (RHS accumulation from
// the history fields, influence-matrix and tau corrections, scatter) use all threads.
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
__device__ __forceinline__ double A_lo(int n, int Nb, double lam) { return -(cN(n - 2, Nb) * lam) / (double)(4 * n * (n - 1)); }
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
    for (int n = nl - 2; n >= par + 2; n -= 2) g[n * TT + t] -= up[n * TM + m] * g[(n + 2) * TT + t];
    double acc = g[par * TT + t];
    for (int n = par + 2; n <= nl; n += 2) acc -= band[n * TM + m] * g[n * TT + t];
    acc /= inv[par * TM + m];  // slot `par` of inv holds diag(0) of this parity block
    g[par * TT + t] = acc;
    double prev = acc;
    for (int n = par + 2; n <= nl; n += 2) {
        const double v = (g[n * TT + t] - A_lo(n, Nb, lam) * prev) * inv[n * TM + m];
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
    const int mxi = q / nkz;
    kz = q - mxi * nkz;
    kx = mxi <= g.Kx ? mxi : mxi - nmx;
    const int mx = kx >= 0 ? kx : g.Nx + kx;
    off = 2L * (kz + (long)(g.Nz / 2 + 1) * mx);
}

// first mode of a tile and number of valid modes in it
__device__ __forceinline__ void tile_modes(int tl, int TM, int nq, int& q0, int& nvalid) {
    if (tl == 0) { q0 = 0; nvalid = 1; return; }
    q0 = 1 + (tl - 1) * TM;
    nvalid = nq - q0 < TM ? nq - q0 : TM;
}

// ---------------------------------------------------------------------------------------------------------------
// One Helmholtz solve  A x = B r  (+ boundary row value bc) for the real and imaginary parts of one (mode, parity),
// fully fused (helmholtz.cpp:79-95 + bandedtridiag.cpp:232-277):
//   backward sweep, n = nl .. par+2:  g_n = B_lo r_{n-2} + B_dg r_n + B_up r_{n+2} ;  x_n = g_n - up_n x_{n+2} ;
//                                     acc -= band_n x_n            (row 0 of the bordered system)
//   row 0:                            x_par = (bc + acc) / diag0
//   forward sweep, n = par+2 .. nl:   x_n = (x_n - lo_n x_{n-2}) inv_n
// rhs(n, re, im) is called for n = nl, nl-2, .., par in strictly descending order (it may carry running sums) and may
// read X at rows < n' for any n' not yet written (rows are written in descending order, one step behind the reads),
// which is what allows the in-place use.  X points at the (re) column of the mode: X[n*TT], X[n*TT+1].
// Optionally accumulates S = sum_n n^2 x_n (wall derivative of the solution, see the influence-matrix step).
template <class RhsF>
__device__ __forceinline__ void helm_chain(double* __restrict__ X, const int N, const int TT, const int par, const double* __restrict__ up,
                                           const double* __restrict__ inv, const double* __restrict__ band, const double* __restrict__ lo,
                                           const int TM, const double* __restrict__ btab, const double bc_re, const double bc_im,
                                           RhsF rhs, double* wall_re, double* wall_im) {
    const int Nb = N - 1;
    const int nl = par ? Nb - 1 : Nb;
    const double* __restrict__ Blo = btab;
    const double* __restrict__ Bdg = btab + N;
    const double* __restrict__ Bup = btab + 2 * N;
    double rc_re, rc_im, rp_re = 0.0, rp_im = 0.0;
    rhs(nl, rc_re, rc_im);
    double xr = 0.0, xi = 0.0, acc_re = bc_re, acc_im = bc_im;
#pragma unroll 2
    for (int n = nl; n >= par + 2; n -= 2) {
        double rm_re, rm_im;
        rhs(n - 2, rm_re, rm_im);
        const double blo = Blo[n], bdg = Bdg[n];
        double g_re = blo * rm_re + bdg * rc_re;
        double g_im = blo * rm_im + bdg * rc_im;
        if (n + 2 <= Nb) {
            const double bup = Bup[n];
            g_re += bup * rp_re;
            g_im += bup * rp_im;
        }
        if (n != nl) {
            const double u = up[n * TM];
            g_re -= u * xr;
            g_im -= u * xi;
        }
        xr = g_re;
        xi = g_im;
        X[n * TT] = xr;
        X[n * TT + 1] = xi;
        const double bd = band[n * TM];
        acc_re -= bd * xr;
        acc_im -= bd * xi;
        rp_re = rc_re; rp_im = rc_im;
        rc_re = rm_re; rc_im = rm_im;
    }
    const double d0 = inv[par * TM];  // slot `par` of inv holds diag(0) of this parity block
    double pr = acc_re / d0, pi = acc_im / d0;
    X[par * TT] = pr;
    X[par * TT + 1] = pi;
    double sr = (double)(par * par) * pr, si = (double)(par * par) * pi;
#pragma unroll 2
    for (int n = par + 2; n <= nl; n += 2) {
        const double l = lo[n * TM], iv = inv[n * TM];
        const double vr = (X[n * TT] - l * pr) * iv;
        const double vi = (X[n * TT + 1] - l * pi) * iv;
        X[n * TT] = vr;
        X[n * TT + 1] = vi;
        pr = vr;
        pi = vi;
        if (wall_re) {
            const double n2 = (double)(n * n);
            sr += n2 * vr;
            si += n2 * vi;
        }
    }
    if (wall_re) { *wall_re = sr; *wall_im = si; }
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
    int q0, nvalid;
    tile_modes(tl, TM, td.nq, q0, nvalid);
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
        if (tid < nvalid) mode_of_q(q, g, kx, kz, off);
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
            inv[n * TM + m] = 1.0 / Akk;
            const double w = A_lo(n, Nb, lam);
            const double upm = A_up(n - 2, Nb, lam) / Akk;
            up[(n - 2) * TM + m] = upm;
            const double dprev = A_dg(n - 2, Nb, lam, nus) - w * upm;
            const double bk = bandk / Akk;
            band[n * TM + m] = bk;
            bandk = 1.0 - w * bk;
            dgk = dprev;
        }
        const int n1 = par + 2;
        inv[n1 * TM + m] = 1.0 / dgk;
        const double b1 = bandk / dgk;
        band[n1 * TM + m] = b1;
        inv[par * TM + m] = 1.0 - A_lo(n1, Nb, lam) * b1;  // diag(0) == band(0)
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
            gP[idx] = A1[idx];
            gv[idx] = A3[idx];
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
        const int m = idx % TM;
        const double dp = s_w[4 * TM + m], dm = s_w[5 * TM + m];
        const double P0 = A2[idx] + (dp * gPp[idx] + dm * gPm[idx]);
        const double v0 = A1[idx] + (dp * gvp[idx] + dm * gvm[idx]);
        A2[idx] = P0;
        A1[idx] = v0;
        gP0[idx] = P0;
        gv0[idx] = v0;
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
__global__ void __launch_bounds__(TAU_THREADS) tau_solve_kernel(const TauSolveParams p) {
    const TauData& td = p.td;
    const int N = td.N, Nb = N - 1, TM = td.TM, TT = 2 * TM;
    const int tid = threadIdx.x, NT = TAU_THREADS;
    const int tl = blockIdx.x;
    const bool is00 = tl == 0;
    const double scale = 4.0 / (td.b - td.a);
    const long rs = (long)p.g.Nx * (2 * (p.g.Nz / 2 + 1));  // row (ny) stride in doubles
    const long cs = rs * p.g.Ny;                            // component stride
    const int AS = N * TT;                                  // doubles per profile array
    const int FS = N * TM;                                  // doubles per factor array

    double* Rx = dyn_smem<double>();
    double* Ry = Rx + AS; double* Rz = Ry + AS; double* Pq = Rz + AS;
    double* Fup = Pq + AS; double* Finv = Fup + FS; double* Fband = Finv + FS; double* Flo = Fband + FS;
    double* s_bt = Flo + FS;                     // [3][N] B rows
    double* s_sc = s_bt + 3 * N;                 // [TSC_COUNT][TM]
    double* s_w = s_sc + TSC_COUNT * TM;         // [8][TT]
    long* s_off = reinterpret_cast<long*>(s_w + 8 * TT);  // [TM]
    int q0, nvalid;
    tile_modes(tl, TM, td.nq, q0, nvalid);

    if (tid < TM) {
        long off = -1;
        if (tid < nvalid) { int kx, kz; mode_of_q(q0 + tid, p.g, kx, kz, off); }
        s_off[tid] = off;
    }
    for (int i = tid; i < TSC_COUNT * TM; i += NT) s_sc[i] = td.tile_sc(tl, 0)[i];
    for (int i = tid; i < 3 * N; i += NT) s_bt[i] = td.btab()[i];
    __syncthreads();

    // ---- S1: right-hand side = linear combination of history fields (dnsalgo.cpp:217-224), complex elements
    {
        const int nel = 3 * N * TM;
        for (int e = tid; e < nel; e += NT) {
            const int comp = e / FS, r = e - comp * FS;
            const int n = r / TM, m = r - n * TM;
            const long off = s_off[m];
            double are = 0.0, aim = 0.0;
            if (off >= 0) {
                const long go = comp * cs + n * rs + off;
                double2 v[TAU_MAXTERMS];
#pragma unroll
                for (int j = 0; j < TAU_MAXTERMS; ++j)
                    if (j < p.nterms) v[j] = *reinterpret_cast<const double2*>(p.term[j] + go);
#pragma unroll
                for (int j = 0; j < TAU_MAXTERMS; ++j)
                    if (j < p.nterms) { are += p.coef[j] * v[j].x; aim += p.coef[j] * v[j].y; }
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
            }
            *reinterpret_cast<double2*>(&Rx[comp * AS + n * TT + 2 * m]) = make_double2(are, aim);  // Rx,Ry,Rz contiguous
        }
        // pressure-operator factors of this tile (upP, invP, bandP are contiguous) and its sub-diagonal
        const double* fsrc = td.tile_arr(tl, TAR_UPP);
        for (int i = tid; i < 3 * FS; i += NT) Fup[i] = fsrc[i];
        for (int i = tid; i < FS; i += NT) {
            const int n = i / TM, m = i - n * TM;
            Flo[i] = n >= 2 ? A_lo(n, Nb, s_sc[TSC_LAMP * TM + m]) : 0.0;
        }
    }
    __syncthreads();

    // ---- S2: pressure Helmholtz  P'' - kappa^2 P = dRy/dy + i (kxx Rx + kzz Rz), P(+-1) = 0  (tausolver.cpp:357-366, 193-201)
    if (tid < 2 * TM) {
        const int m = tid % TM, par = tid / TM;
        if (m < nvalid) {
            const double kxx = s_sc[TSC_KXX * TM + m], kzz = s_sc[TSC_KZZ * TM + m];
            const double* ry = Ry + 2 * m; const double* rx = Rx + 2 * m; const double* rz = Rz + 2 * m;
            double run_re = 0.0, run_im = 0.0;
            auto rhs = [&](int n, double& re, double& im) {
                if (n + 1 <= Nb) {
                    const double f = scale * (n + 1);
                    run_re = run_re + f * ry[(n + 1) * TT];
                    run_im = run_im + f * ry[(n + 1) * TT + 1];
                }
                double dre = run_re, dim = run_im;
                if (n == 0) { dre *= 0.5; dim *= 0.5; }
                re = dre - (kxx * rx[n * TT + 1] + kzz * rz[n * TT + 1]);
                im = dim + (kxx * rx[n * TT] + kzz * rz[n * TT]);
            };
            helm_chain(Pq + 2 * m, N, TT, par, Fup + m, Finv + m, Fband + m, Flo + m, TM, s_bt, 0.0, 0.0, rhs, nullptr, nullptr);
        }
    }
    __syncthreads();

    // ---- velocity-operator factors replace the pressure ones
    {
        const double* fsrc = td.tile_arr(tl, TAR_UPV);
        for (int i = tid; i < 3 * FS; i += NT) Fup[i] = fsrc[i];
        for (int i = tid; i < FS; i += NT) {
            const int n = i / TM, m = i - n * TM;
            Flo[i] = n >= 2 ? A_lo(n, Nb, s_sc[TSC_LAMV * TM + m]) : 0.0;
        }
    }
    __syncthreads();

    if (!is00) {
        // ---- S3: v particular solution  nu v'' - lambda v = P' - Ry, v(+-1) = 0, in place in Ry  (tausolver.cpp:203-210)
        if (tid < 2 * TM) {
            const int m = tid % TM, par = tid / TM;
            if (m < nvalid) {
                double* ry = Ry + 2 * m; const double* pq = Pq + 2 * m;
                const int ntop = par ? Nb - 1 : Nb;  // Ry[Nb], Ry[Nb-1] are needed again by the tau correction
                s_w[(4 + par) * TT + 2 * m] = ry[ntop * TT];
                s_w[(4 + par) * TT + 2 * m + 1] = ry[ntop * TT + 1];
                double run_re = 0.0, run_im = 0.0;
                auto rhs = [&](int n, double& re, double& im) {
                    if (n + 1 <= Nb) {
                        const double f = scale * (n + 1);
                        run_re = run_re + f * pq[(n + 1) * TT];
                        run_im = run_im + f * pq[(n + 1) * TT + 1];
                    }
                    double dre = run_re, dim = run_im;
                    if (n == 0) { dre *= 0.5; dim *= 0.5; }
                    re = dre - ry[n * TT];
                    im = dim - ry[n * TT + 1];
                };
                helm_chain(ry, N, TT, par, Fup + m, Finv + m, Fband + m, Flo + m, TM, s_bt, 0.0, 0.0, rhs,
                           &s_w[(6 + par) * TT + 2 * m], &s_w[(6 + par) * TT + 2 * m + 1]);
            }
        }
        __syncthreads();
        // ---- S4: influence-matrix (tausolver.cpp:178-191) and tau (tausolver.cpp:215-244) correction amplitudes.
        // v'(b) = (2/L) sum n^2 v_n, v'(a) = (2/L) sum (-1)^(n+1) n^2 v_n  (closed form of eval_b/eval_a of diff(v))
        const double* gPp = td.tile_arr(tl, TAR_PP); const double* gvp = td.tile_arr(tl, TAR_VP);
        const double* gPm = td.tile_arr(tl, TAR_PM); const double* gvm = td.tile_arr(tl, TAR_VM);
        const double* gP0 = td.tile_arr(tl, TAR_P0); const double* gv0 = td.tile_arr(tl, TAR_V0);
        if (tid < TT) {
            const int m = tid >> 1;
            if (m < nvalid) {
                const double Se = s_w[6 * TT + tid], So = s_w[7 * TT + tid];
                const double vb = 0.5 * scale * (So + Se), va = 0.5 * scale * (So - Se);
                const double dp = -s_sc[TSC_I00 * TM + m] * vb - s_sc[TSC_I01 * TM + m] * va;
                const double dm = -s_sc[TSC_I10 * TM + m] * vb - s_sc[TSC_I11 * TM + m] * va;
                s_w[tid] = dp;
                s_w[TT + tid] = dm;
                if (p.taucorr) {
                    const double lam = s_sc[TSC_LAMV * TM + m];
                    const int iNb = Nb * TM + m, iNb1 = (Nb - 1) * TM + m;
                    const double vNb = Ry[Nb * TT + tid] + (dp * gvp[iNb] + dm * gvm[iNb]);
                    const double vNb1 = Ry[(Nb - 1) * TT + tid] + (dp * gvp[iNb1] + dm * gvm[iNb1]);
                    const double pNb = Pq[Nb * TT + tid] + (dp * gPp[iNb] + dm * gPm[iNb]);
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
        for (int e = tid; e < FS; e += NT) {
            const int n = e / TM, m = e - n * TM;
            if (m >= nvalid) continue;
            const double pp = gPp[e], vp = gvp[e], pm = gPm[e], vm = gvm[e];
            double2 P = *reinterpret_cast<double2*>(&Pq[n * TT + 2 * m]);
            double2 V = *reinterpret_cast<double2*>(&Ry[n * TT + 2 * m]);
            const double dpr = s_w[2 * m], dpi = s_w[2 * m + 1], dmr = s_w[TT + 2 * m], dmi = s_w[TT + 2 * m + 1];
            P.x += dpr * pp + dmr * pm;
            P.y += dpi * pp + dmi * pm;
            V.x += dpr * vp + dmr * vm;
            V.y += dpi * vp + dmi * vm;
            if (p.taucorr) {
                const double p0 = gP0[e], v0 = gv0[e];
                const double sNbr = s_w[2 * TT + 2 * m], sNbi = s_w[2 * TT + 2 * m + 1];
                const double sNb1r = s_w[3 * TT + 2 * m], sNb1i = s_w[3 * TT + 2 * m + 1];
                const bool ev = (n & 1) == 0;
                P.x += (ev ? sNb1r : sNbr) * p0;
                P.y += (ev ? sNb1i : sNbi) * p0;
                V.x += (ev ? sNbr : sNb1r) * v0;
                V.y += (ev ? sNbi : sNb1i) * v0;
            }
            *reinterpret_cast<double2*>(&Pq[n * TT + 2 * m]) = P;
            *reinterpret_cast<double2*>(&Ry[n * TT + 2 * m]) = V;
        }
        __syncthreads();
    } else if (p.constraint == 1) {
        // mean mode with the bulk-velocity constraint (helmholtz.cpp:158-213): keep the right-hand sides, v = 0 anyway
        for (int n = tid; n < N; n += NT) {
            Ry[n * TT] = Rx[n * TT];
            Ry[n * TT + 1] = Rz[n * TT];
        }
        __syncthreads();
    }

    // ---- S5: u, w from the x/z momentum equations  nu u'' - lambda u = i kxx P - Rx, in place in Rx, Rz (tausolver.cpp:368-384).
    // Mean mode + bulk velocity: the imaginary slot carries the solve with right-hand side nu*T0 instead ("uc").
    const bool bulk00 = is00 && p.constraint == 1;
    if (tid < 4 * TM) {
        const int m = tid % TM, par = (tid / TM) & 1, which = tid / (2 * TM);
        if (m < nvalid) {
            const double k = s_sc[(which ? TSC_KZZ : TSC_KXX) * TM + m];
            double* rr = (which ? Rz : Rx) + 2 * m; const double* pq = Pq + 2 * m;
            const double nu = td.nu;
            auto rhs = [&](int n, double& re, double& im) {
                re = -k * pq[n * TT + 1] - rr[n * TT];
                im = k * pq[n * TT] - rr[n * TT + 1];
                if (bulk00) im = n == 0 ? nu : 0.0;
            };
            helm_chain(rr, N, TT, par, Fup + m, Finv + m, Fband + m, Flo + m, TM, s_bt, 0.0, 0.0, rhs, nullptr, nullptr);
        }
    }
    __syncthreads();
    if (bulk00) {
        if (tid < 2) {
            const double* ua = tid == 0 ? Rx : Rz;  // re: solution for the actual rhs, im: uc
            double uam = ua[0], ucm = ua[1];
            for (int n = 2; n < N; n += 2) { uam -= ua[n * TT] / (double)(n * n - 1); ucm -= ua[n * TT + 1] / (double)(n * n - 1); }
            const double target = tid == 0 ? p.umean_target : p.wmean_target;
            const double mu = td.nu * (target - uam) / ucm;
            s_w[tid] = mu;
            if (p.dPd_act) p.dPd_act[tid] = mu;
        }
        __syncthreads();
        if (tid < 4) {
            const int par = tid & 1, which = tid >> 1;
            double* rr = which ? Rz : Rx;
            const double* saved = Ry + which;  // Ry.re = original Rx, Ry.im = original Rz
            const double mu = s_w[which];
            auto rhs = [&](int n, double& re, double& im) {
                re = -saved[n * TT] + (n == 0 ? mu : 0.0);
                im = 0.0;
            };
            helm_chain(rr, N, TT, par, Fup, Finv, Fband, Flo, TM, s_bt, 0.0, 0.0, rhs, nullptr, nullptr);
        }
        __syncthreads();
    }

    // ---- S6: scatter (nse.cpp:566-572)
    for (int e = tid; e < FS; e += NT) {
        const int n = e / TM, m = e - n * TM;
        const long off = s_off[m];
        if (off < 0) continue;
        const long go = n * rs + off;
        const double2 U = *reinterpret_cast<double2*>(&Rx[n * TT + 2 * m]);
        double2 V = *reinterpret_cast<double2*>(&Ry[n * TT + 2 * m]);
        const double2 W = *reinterpret_cast<double2*>(&Rz[n * TT + 2 * m]);
        const double2 P = *reinterpret_cast<double2*>(&Pq[n * TT + 2 * m]);
        if (is00) V = make_double2(0.0, 0.0);
        *reinterpret_cast<double2*>(&p.uout[go]) = U;
        *reinterpret_cast<double2*>(&p.uout[cs + go]) = V;
        *reinterpret_cast<double2*>(&p.uout[2 * cs + go]) = W;
        *reinterpret_cast<double2*>(&p.qout[go]) = P;
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
    const long rs = (long)p.g.Nx * (2 * (p.g.Nz / 2 + 1));
    const long cs = rs * p.g.Ny;
    const int TM = p.TM_lin, TT = 2 * TM;
    const size_t AS = (size_t)N * TT;
    double* Pk = dyn_smem<double>();
    double* Pyk = Pk + AS; double* X = Pyk + AS; double* T = X + AS; double* R = T + AS;
    double* s_k = R + AS;  // [3][TM]: kappa2, kxx, kzz
    long* s_off = reinterpret_cast<long*>(s_k + 3 * TM);
    __shared__ double s_shear[2];
    const int q0 = blockIdx.x * TM;
    if (tid < TM) {
        const int qq = q0 + tid;
        long off = -1;
        if (qq < td.nq) { int kx, kz; mode_of_q(qq, p.g, kx, kz, off); }
        s_off[tid] = off;
        const int qs = qq < td.nq ? qq : 0;
        s_k[tid] = td.scq(TSC_LAMP, qs);
        s_k[TM + tid] = td.scq(TSC_KXX, qs);
        s_k[2 * TM + tid] = td.scq(TSC_KZZ, qs);
    }
    __syncthreads();
    for (int idx = tid; idx < (int)AS; idx += NT) {
        const int n = idx / TT, t = idx - n * TT;
        const long off = s_off[t >> 1];
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
        if (blockIdx.x == 0 && p.constraint == 1 && comp != 1 && tid == 0) {
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
            if (q0 + m == 0 && (t & 1) == 0) {
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

static size_t solve_smem(int N, int TM) {
    const int TT = 2 * TM;
    return ((size_t)4 * N * TT + 4 * N * TM + 3 * N + TSC_COUNT * TM + 8 * TT) * sizeof(double) + TM * sizeof(long);
}

// modes per tile of the solve kernel: two CTAs per SM when the profiles are long (one streams while the other
// walks its recurrences), at most 8 modes (128-byte runs of the history fields)
int tau_pick_TM_solve(int N) {
    int TM = 8;
    while (TM > 1 && solve_smem(N, TM) > 110 * 1024) --TM;
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
    const size_t smem = ((size_t)5 * p.td.N * TT + 3 * p.TM_lin) * sizeof(double) + p.TM_lin * sizeof(long);
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

int tau_solve_launch(const TauSolveParams& p, cudaStream_t stream) {
    const size_t smem = solve_smem(p.td.N, p.td.TM);
    if (smem > 227 * 1024) {
        set_last_error("tau_solve: Ny too large for the shared-memory tile");
        return 1;
    }
    static size_t configured = 0;
    auto kfn = tau_solve_kernel;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(p.td.ntiles);
    CF_LAUNCH(kfn, grid, dim3(TAU_THREADS), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}

}  // namespace cfgpu
