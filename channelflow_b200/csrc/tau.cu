// Batched tau-solver kernels (setup = factorisation + influence matrix, solve = fused RHS + Kleiser-Schumann).
// See tau.cuh for what is replaced.  One CTA owns TM consecutive retained modes; their Chebyshev profiles live
// in shared memory as [n][t] (t = 2*mode + re/im, fastest), so every stage is either a coalesced data-parallel
// sweep over (n,t) or a set of independent sequential recurrences ("chains", one thread per
// (mode, re/im, parity)) that walk n through shared memory exactly in the reference's operation order.
// Roofline: HBM (history fields + factors are each read once, outputs written once).
#include "tau.cuh"

namespace cfgpu {

namespace {

constexpr int TAU_THREADS = 128;
constexpr double PI = 3.14159265358979323846264338327950288;

__device__ __forceinline__ double cN(int m, int Nb) { return (m == 0 || m == Nb) ? 2.0 : 1.0; }
__device__ __forceinline__ int betaN(int n, int Nb) { return (n > Nb - 2) ? 0 : 1; }
// C&H 5.1.24 rows n >= 2 of the quasi-tridiagonal systems (helmholtz.cpp:44-56)
__device__ __forceinline__ double A_lo(int n, int Nb, double lam) { return -(cN(n - 2, Nb) * lam) / (double)(4 * n * (n - 1)); }
__device__ __forceinline__ double A_dg(int n, int Nb, double lam, double nus) {
    return nus + (betaN(n, Nb) * lam) / (double)(2 * (n * n - 1));
}
__device__ __forceinline__ double A_up(int n, int Nb, double lam) {
    return betaN(n + 2, Nb) ? -lam / (double)(4 * n * (n + 1)) : 0.0;
}
__device__ __forceinline__ double B_lo(int n, int Nb) { return cN(n - 2, Nb) / (double)(4 * n * (n - 1)); }
__device__ __forceinline__ double B_dg(int n, int Nb) { return -((double)betaN(n, Nb)) / (double)(2 * (n * n - 1)); }
__device__ __forceinline__ double B_up(int n, int Nb) { return betaN(n + 2, Nb) ? 1.0 / (double)(4 * n * (n + 1)) : 0.0; }

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

// UL solve of one parity block in place (bandedtridiag.cpp:258-273). Factor arrays are [n][ldq] in HBM.
__device__ __forceinline__ void ul_solve_chain(double* g, int N, int TT, int t, int par, const double* up,
                                               const double* inv, const double* band, size_t ldq, int q, double lam) {
    const int Nb = N - 1;
    const int nl = par ? Nb - 1 : Nb;
    for (int n = nl - 2; n >= par + 2; n -= 2) g[n * TT + t] -= up[(size_t)n * ldq + q] * g[(n + 2) * TT + t];
    double acc = g[par * TT + t];
    for (int n = par + 2; n <= nl; n += 2) acc -= band[(size_t)n * ldq + q] * g[n * TT + t];
    acc /= inv[(size_t)par * ldq + q];  // slot `par` of inv holds diag(0) of this parity block
    g[par * TT + t] = acc;
    double prev = acc;
    for (int n = par + 2; n <= nl; n += 2) {
        const double v = (g[n * TT + t] - A_lo(n, Nb, lam) * prev) * inv[(size_t)n * ldq + q];
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

}  // namespace

// =================================================================================================== setup
// grid = ceil(nq / TM). Real profiles: smem arrays [n][TM].
__global__ void __launch_bounds__(TAU_THREADS) tau_setup_kernel(const TauData td, const ModeGeom g, const double lambda_t, const int TM) {
    const int N = td.N, Nb = N - 1;
    const size_t ldq = td.ldq;
    const int tid = threadIdx.x, NT = TAU_THREADS;
    const int q0 = blockIdx.x * TM;
    double* A1 = dyn_smem<double>();
    double* A2 = A1 + (size_t)N * TM;
    double* A3 = A2 + (size_t)N * TM;
    double* s_lamP = A3 + (size_t)N * TM;
    double* s_lamV = s_lamP + TM;
    double* s_w = s_lamV + TM;  // [8][TM]: Ab, Ca, Bb, Da, dplus, dminus, dP0dy_Nb1, spare

    const double scale = 4.0 / (td.b - td.a);
    const double nusP = 1.0 / (((td.b - td.a) / 2) * ((td.b - td.a) / 2));
    const double nusV = td.nu / (((td.b - td.a) / 2) * ((td.b - td.a) / 2));

    double* upP = td.arr(0); double* invP = td.arr(1); double* bandP = td.arr(2);
    double* upV = td.arr(3); double* invV = td.arr(4); double* bandV = td.arr(5);
    double* gPp = td.arr(6); double* gvp = td.arr(7); double* gPm = td.arr(8); double* gvm = td.arr(9);
    double* gP0 = td.arr(10); double* gv0 = td.arr(11);

    if (tid < TM) {
        const int q = q0 + tid;
        int kx = 0, kz = 0;
        long off;
        if (q < td.nq) mode_of_q(q, g, kx, kz, off);
        const double kxL = kx / g.Lx, kzL = kz / g.Lz;
        const double kappa2 = 4 * (PI * PI) * (kxL * kxL + kzL * kzL);
        const double c = 4.0 * (PI * PI) * td.nu;
        const double lamV = lambda_t + c * (kxL * kxL + kzL * kzL);
        s_lamP[tid] = kappa2;
        s_lamV[tid] = lamV;
        td.sc(TSC_LAMP)[q] = kappa2;
        td.sc(TSC_LAMV)[q] = lamV;
        td.sc(TSC_KXX)[q] = 2 * PI * kx / g.Lx;
        td.sc(TSC_KZZ)[q] = 2 * PI * kz / g.Lz;
    }
    __syncthreads();

    // ---- UL factorisation of Ae, Ao for both Helmholtz operators (bandedtridiag.cpp:212-229)
    for (int c = tid; c < 4 * TM; c += NT) {
        const int m = c % TM, par = (c / TM) & 1, h = c / (2 * TM);
        const int q = q0 + m;
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
            inv[(size_t)n * ldq + q] = 1.0 / Akk;
            const double w = A_lo(n, Nb, lam);
            const double upm = A_up(n - 2, Nb, lam) / Akk;
            up[(size_t)(n - 2) * ldq + q] = upm;
            const double dprev = A_dg(n - 2, Nb, lam, nus) - w * upm;
            const double bk = bandk / Akk;
            band[(size_t)n * ldq + q] = bk;
            bandk = 1.0 - w * bk;
            dgk = dprev;
        }
        const int n1 = par + 2;
        inv[(size_t)n1 * ldq + q] = 1.0 / dgk;
        const double b1 = bandk / dgk;
        band[(size_t)n1 * ldq + q] = b1;
        inv[(size_t)par * ldq + q] = 1.0 - A_lo(n1, Nb, lam) * b1;  // diag(0) == band(0)
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
            ul_solve_chain(A1, N, TM, m, par, upP, invP, bandP, ldq, q0 + m, s_lamP[m]);
        }
        __syncthreads();
        for (int c = tid; c < 2 * TM; c += NT) diff_chain(A1, A2, N, TM, c % TM, c / TM, scale, nullptr);
        __syncthreads();
        bmul(A2, A3, N, TM, 0.0, 0.0, tid, NT);
        __syncthreads();
        for (int c = tid; c < 2 * TM; c += NT) {
            const int m = c % TM, par = c / TM;
            ul_solve_chain(A3, N, TM, m, par, upV, invV, bandV, ldq, q0 + m, s_lamV[m]);
        }
        __syncthreads();
        double* gP = pm == 0 ? gPp : gPm;
        double* gv = pm == 0 ? gvp : gvm;
        for (int idx = tid; idx < N * TM; idx += NT) {
            const int n = idx / TM, m = idx - n * TM;
            gP[(size_t)n * ldq + q0 + m] = A1[idx];
            gv[(size_t)n * ldq + q0 + m] = A3[idx];
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
        const int q = q0 + tid;
        const double A = s_w[0 * TM + tid], C = s_w[1 * TM + tid], B = s_w[2 * TM + tid], D = s_w[3 * TM + tid];
        const double disc = A * D - B * C;
        td.sc(TSC_I00)[q] = D / disc;
        td.sc(TSC_I01)[q] = -B / disc;
        td.sc(TSC_I10)[q] = -C / disc;
        td.sc(TSC_I11)[q] = A / disc;
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
        ul_solve_chain(A2, N, TM, m, par, upP, invP, bandP, ldq, q0 + m, s_lamP[m]);
    }
    __syncthreads();  // A2 = P0 (before influence correction)
    for (int c = tid; c < 2 * TM; c += NT) diff_chain(A2, A3, N, TM, c % TM, c / TM, scale, nullptr);
    __syncthreads();  // A3 = dP0/dy
    if (tid < TM) s_w[6 * TM + tid] = A3[(Nb - 1) * TM + tid];
    bmul(A3, A1, N, TM, 0.0, 0.0, tid, NT);
    __syncthreads();
    for (int c = tid; c < 2 * TM; c += NT) {
        const int m = c % TM, par = c / TM;
        ul_solve_chain(A1, N, TM, m, par, upV, invV, bandV, ldq, q0 + m, s_lamV[m]);
    }
    __syncthreads();  // A1 = v0
    if (tid < TM) {
        const int q = q0 + tid;
        double vb, va;
        dudy_at_walls(A1, N, TM, tid, scale, vb, va);
        s_w[4 * TM + tid] = -td.sc(TSC_I00)[q] * vb - td.sc(TSC_I01)[q] * va;
        s_w[5 * TM + tid] = -td.sc(TSC_I10)[q] * vb - td.sc(TSC_I11)[q] * va;
    }
    __syncthreads();
    for (int idx = tid; idx < N * TM; idx += NT) {
        const int n = idx / TM, m = idx - n * TM;
        const size_t go = (size_t)n * ldq + q0 + m;
        const double dp = s_w[4 * TM + m], dm = s_w[5 * TM + m];
        const double P0 = A2[idx] + (dp * gPp[go] + dm * gPm[go]);
        const double v0 = A1[idx] + (dp * gvp[go] + dm * gvm[go]);
        A2[idx] = P0;
        A1[idx] = v0;
        gP0[go] = P0;
        gv0[go] = v0;
    }
    __syncthreads();
    if (tid < TM) {
        const int q = q0 + tid;
        const double lam = s_lamV[tid];
        // v0'' has zero coefficients at Nb-1 and Nb, dP0/dy[Nb] == 0 (chebyshev.cpp:688-689)
        td.sc(TSC_S0NB1)[q] = lam * A1[(Nb - 1) * TM + tid] + s_w[6 * TM + tid] - td.nu * 0.0;
        td.sc(TSC_S0NB)[q] = lam * A1[Nb * TM + tid] + 0.0 - td.nu * 0.0;
    }
}

// =================================================================================================== solve
// grid = ceil((nq-1)/TM) + 1 ; the last CTA handles the (0,0) mode (mean-flow constraint).
__global__ void __launch_bounds__(TAU_THREADS) tau_solve_kernel(const TauSolveParams p) {
    const TauData& td = p.td;
    const int N = td.N, Nb = N - 1;
    const size_t ldq = td.ldq;
    const int tid = threadIdx.x, NT = TAU_THREADS;
    const double scale = 4.0 / (td.b - td.a);
    const long rs = (long)p.g.Nx * (2 * (p.g.Nz / 2 + 1));  // row (ny) stride in doubles
    const long cs = rs * p.g.Ny;                            // component stride
    const double* upP = td.arr(0); const double* invP = td.arr(1); const double* bandP = td.arr(2);
    const double* upV = td.arr(3); const double* invV = td.arr(4); const double* bandV = td.arr(5);

    const int ntiles = (td.nq - 1 + p.TM - 1) / p.TM;
    if ((int)blockIdx.x == ntiles) {
        // ------------------------------------------------------------------ (0,0) mode, real parts only
        double* Rx = dyn_smem<double>();
        double* Ry = Rx + N; double* Rz = Ry + N; double* Pq = Rz + N; double* T = Pq + N; double* Uu = T + N;
        double* Ww = Uu + N; double* X1 = Ww + N; double* X2 = X1 + N;
        __shared__ double s_mu[2];
        for (int idx = tid; idx < 3 * N; idx += NT) {
            const int comp = idx / N, n = idx - comp * N;
            double acc = 0.0;
            for (int j = 0; j < p.nterms; ++j) acc += p.coef[j] * p.term[j][comp * cs + n * rs];
            if (comp == 0 && p.Ubaseyy) acc += td.nu * p.Ubaseyy[n];
            if (comp == 2 && p.Wbaseyy) acc += td.nu * p.Wbaseyy[n];
            if (p.constraint == 0 && n == 0) {
                if (comp == 0) acc -= p.dPdxRef;
                if (comp == 2) acc -= p.dPdzRef;
            }
            Rx[idx] = acc;  // Rx,Ry,Rz contiguous
        }
        __syncthreads();
        const double lamP = td.sc(TSC_LAMP)[0], lamV = td.sc(TSC_LAMV)[0];
        if (tid < 2) diff_chain(Ry, T, N, 1, 0, tid, scale, nullptr);  // r = Ry'
        __syncthreads();
        bmul(T, Pq, N, 1, 0.0, 0.0, tid, NT);
        for (int idx = tid; idx < N; idx += NT) { X1[idx] = -Rx[idx]; X2[idx] = -Rz[idx]; }
        __syncthreads();
        bmul(X1, Uu, N, 1, 0.0, 0.0, tid, NT);
        bmul(X2, Ww, N, 1, 0.0, 0.0, tid, NT);
        __syncthreads();
        if (tid < 2) ul_solve_chain(Pq, N, 1, 0, tid, upP, invP, bandP, ldq, 0, lamP);
        else if (tid < 4) ul_solve_chain(Uu, N, 1, 0, tid - 2, upV, invV, bandV, ldq, 0, lamV);
        else if (tid < 6) ul_solve_chain(Ww, N, 1, 0, tid - 4, upV, invV, bandV, ldq, 0, lamV);
        __syncthreads();
        if (p.constraint == 1) {
            // mean-constrained Helmholtz (helmholtz.cpp:158-213): Uu/Ww hold the "eqn1" solutions
            for (int idx = tid; idx < N; idx += NT) T[idx] = idx == 0 ? td.nu : 0.0;
            __syncthreads();
            bmul(T, Ry, N, 1, 0.0, 0.0, tid, NT);  // Ry is free now; becomes uc
            __syncthreads();
            if (tid < 2) ul_solve_chain(Ry, N, 1, 0, tid, upV, invV, bandV, ldq, 0, lamV);
            __syncthreads();
            if (tid < 2) {
                const double* ua = tid == 0 ? Uu : Ww;
                double uam = ua[0], ucm = Ry[0];
                for (int n = 2; n < N; n += 2) { uam -= ua[n] / (double)(n * n - 1); ucm -= Ry[n] / (double)(n * n - 1); }
                const double target = tid == 0 ? p.umean_target : p.wmean_target;
                s_mu[tid] = td.nu * (target - uam) / ucm;
            }
            __syncthreads();
            for (int idx = tid; idx < N; idx += NT) {
                X1[idx] = -Rx[idx] + (idx == 0 ? s_mu[0] : 0.0);
                X2[idx] = -Rz[idx] + (idx == 0 ? s_mu[1] : 0.0);
            }
            __syncthreads();
            bmul(X1, Uu, N, 1, 0.0, 0.0, tid, NT);
            bmul(X2, Ww, N, 1, 0.0, 0.0, tid, NT);
            __syncthreads();
            if (tid < 2) ul_solve_chain(Uu, N, 1, 0, tid, upV, invV, bandV, ldq, 0, lamV);
            else if (tid < 4) ul_solve_chain(Ww, N, 1, 0, tid - 2, upV, invV, bandV, ldq, 0, lamV);
            if (tid == 0 && p.dPd_act) { p.dPd_act[0] = s_mu[0]; p.dPd_act[1] = s_mu[1]; }
            __syncthreads();
        }
        for (int idx = tid; idx < N; idx += NT) {
            const long o = (long)idx * rs;
            *reinterpret_cast<double2*>(&p.uout[o]) = make_double2(Uu[idx], 0.0);
            *reinterpret_cast<double2*>(&p.uout[cs + o]) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2*>(&p.uout[2 * cs + o]) = make_double2(Ww[idx], 0.0);
            *reinterpret_cast<double2*>(&p.qout[o]) = make_double2(Pq[idx], 0.0);
        }
        return;
    }

    // ---------------------------------------------------------------------- general modes
    const int TM = p.TM, TT = 2 * TM;
    const size_t AS = (size_t)N * TT;
    double* Rx = dyn_smem<double>();
    double* Ry = Rx + AS; double* Rz = Ry + AS; double* Pq = Rz + AS; double* V = Pq + AS; double* T = V + AS;
    double* s_sc = T + AS;                       // [10][TM]: lamP lamV kxx kzz i00 i01 i10 i11 s0nb1 s0nb
    double* s_w = s_sc + 10 * TM;                // [2][TT]
    long* s_off = reinterpret_cast<long*>(s_w + 2 * TT);  // [TM]
    const int q0 = 1 + blockIdx.x * TM;

    if (tid < TM) {
        const int q = q0 + tid;
        long off = -1;
        if (q < td.nq) { int kx, kz; mode_of_q(q, p.g, kx, kz, off); }
        s_off[tid] = off;
        const int qq = q < td.nq ? q : 0;
        for (int s = 0; s < 10; ++s) s_sc[s * TM + tid] = td.sc(s)[qq];
    }
    __syncthreads();

    // P0: right-hand side = linear combination of history fields (dnsalgo.cpp:217-224)
    for (int idx = tid; idx < 3 * (int)AS; idx += NT) {
        const int comp = idx / (int)AS, r = idx - comp * (int)AS;
        const int n = r / TT, t = r - n * TT;
        const long off = s_off[t >> 1];
        double acc = 0.0;
        if (off >= 0) {
            const long go = comp * cs + n * rs + off + (t & 1);
            for (int j = 0; j < p.nterms; ++j) acc += p.coef[j] * p.term[j][go];
        }
        Rx[idx] = acc;  // Rx,Ry,Rz contiguous
    }
    __syncthreads();

    // P1: r = dRy/dy + i (kxx Rx + kzz Rz)   (tausolver.cpp:357-366)
    for (int c = tid; c < 2 * TT; c += NT) diff_chain(Ry, T, N, TT, c % TT, c / TT, scale, nullptr);
    __syncthreads();
    for (int idx = tid; idx < (int)AS; idx += NT) {
        const int t = idx % TT, m = t >> 1;
        const double kxx = s_sc[TSC_KXX * TM + m], kzz = s_sc[TSC_KZZ * TM + m];
        if ((t & 1) == 0) T[idx] -= kxx * Rx[idx + 1] + kzz * Rz[idx + 1];
        else T[idx] += kxx * Rx[idx - 1] + kzz * Rz[idx - 1];
    }
    __syncthreads();
    // P2: pressure Helmholtz
    bmul(T, Pq, N, TT, 0.0, 0.0, tid, NT);
    __syncthreads();
    for (int c = tid; c < 2 * TT; c += NT) {
        const int t = c % TT, m = t >> 1;
        ul_solve_chain(Pq, N, TT, t, c / TT, upP, invP, bandP, ldq, q0 + m < td.nq ? q0 + m : 0, s_sc[TSC_LAMP * TM + m]);
    }
    __syncthreads();
    // P3: v particular solution: nu v'' - lambda v = P' - Ry
    for (int c = tid; c < 2 * TT; c += NT) diff_chain(Pq, T, N, TT, c % TT, c / TT, scale, Ry);
    __syncthreads();
    bmul(T, V, N, TT, 0.0, 0.0, tid, NT);
    __syncthreads();
    for (int c = tid; c < 2 * TT; c += NT) {
        const int t = c % TT, m = t >> 1;
        ul_solve_chain(V, N, TT, t, c / TT, upV, invV, bandV, ldq, q0 + m < td.nq ? q0 + m : 0, s_sc[TSC_LAMV * TM + m]);
    }
    __syncthreads();
    // P4: influence-matrix correction (tausolver.cpp:178-191)
    if (tid < TT) {
        const int m = tid >> 1;
        double vb, va;
        dudy_at_walls(V, N, TT, tid, scale, vb, va);
        s_w[tid] = -s_sc[TSC_I00 * TM + m] * vb - s_sc[TSC_I01 * TM + m] * va;
        s_w[TT + tid] = -s_sc[TSC_I10 * TM + m] * vb - s_sc[TSC_I11 * TM + m] * va;
    }
    __syncthreads();
    {
        const double* gPp = td.arr(6); const double* gvp = td.arr(7); const double* gPm = td.arr(8); const double* gvm = td.arr(9);
        for (int idx = tid; idx < (int)AS; idx += NT) {
            const int n = idx / TT, t = idx - n * TT, m = t >> 1;
            const int q = q0 + m < td.nq ? q0 + m : 0;
            const size_t go = (size_t)n * ldq + q;
            const double dp = s_w[t], dm = s_w[TT + t];
            Pq[idx] += dp * gPp[go] + dm * gPm[go];
            V[idx] += dp * gvp[go] + dm * gvm[go];
        }
    }
    __syncthreads();
    // P5: tau correction (tausolver.cpp:215-244). v'' has zero coefficients at Nb-1, Nb; P'[Nb] = 0.
    if (p.taucorr) {
        if (tid < TT) {
            const int m = tid >> 1;
            const double lam = s_sc[TSC_LAMV * TM + m];
            double s1nb = lam * V[Nb * TT + tid] - td.nu * 0.0 - Ry[Nb * TT + tid];
            double s1nb1 = lam * V[(Nb - 1) * TT + tid] - td.nu * 0.0 - Ry[(Nb - 1) * TT + tid];
            s1nb += 0.0;
            s1nb1 += scale * Nb * Pq[Nb * TT + tid];
            s_w[tid] = s1nb / (1.0 - s_sc[TSC_S0NB * TM + m]);          // sigmaNb
            s_w[TT + tid] = s1nb1 / (1.0 - s_sc[TSC_S0NB1 * TM + m]);   // sigmaNb1
        }
        __syncthreads();
        const double* gP0 = td.arr(10); const double* gv0 = td.arr(11);
        for (int idx = tid; idx < (int)AS; idx += NT) {
            const int n = idx / TT, t = idx - n * TT, m = t >> 1;
            const int q = q0 + m < td.nq ? q0 + m : 0;
            const size_t go = (size_t)n * ldq + q;
            const double sNb = s_w[t], sNb1 = s_w[TT + t];
            Pq[idx] += ((n % 2 == 0) ? sNb1 : sNb) * gP0[go];
            V[idx] += ((n % 2 == 0) ? sNb : sNb1) * gv0[go];
        }
        __syncthreads();
    }
    // P6: u, w from the x/z momentum equations: nu u'' - lambda u = i kxx P - Rx  (tausolver.cpp:368-384)
    for (int idx = tid; idx < (int)AS; idx += NT) {
        const int t = idx % TT, m = t >> 1;
        const double kxx = s_sc[TSC_KXX * TM + m], kzz = s_sc[TSC_KZZ * TM + m];
        if ((t & 1) == 0) {
            Rx[idx] = -kxx * Pq[idx + 1] - Rx[idx];
            Rz[idx] = -kzz * Pq[idx + 1] - Rz[idx];
        } else {
            Rx[idx] = kxx * Pq[idx - 1] - Rx[idx];
            Rz[idx] = kzz * Pq[idx - 1] - Rz[idx];
        }
    }
    __syncthreads();
    bmul(Rx, T, N, TT, 0.0, 0.0, tid, NT);
    bmul(Rz, Ry, N, TT, 0.0, 0.0, tid, NT);
    __syncthreads();
    for (int c = tid; c < 4 * TT; c += NT) {
        const int t = c % TT, m = t >> 1, par = (c / TT) & 1, which = c / (2 * TT);
        ul_solve_chain(which ? Ry : T, N, TT, t, par, upV, invV, bandV, ldq, q0 + m < td.nq ? q0 + m : 0, s_sc[TSC_LAMV * TM + m]);
    }
    __syncthreads();
    // P7: scatter (nse.cpp:566-572)
    for (int idx = tid; idx < (int)AS; idx += NT) {
        const int n = idx / TT, t = idx - n * TT;
        const long off = s_off[t >> 1];
        if (off < 0) continue;
        const long go = n * rs + off + (t & 1);
        p.uout[go] = T[idx];
        p.uout[cs + go] = V[idx];
        p.uout[2 * cs + go] = Ry[idx];
        p.qout[go] = Pq[idx];
    }
}

// =================================================================================================== linear
// NSE::linear (nse.cpp:393-477): L = nu u'' - nu kappa^2 u - grad q  per retained mode (+ mean-mode constants).
// grid = ceil(nq/TM). smem: Pk, Pyk, X, T, R as [n][t].
__global__ void __launch_bounds__(TAU_THREADS) linear_kernel(const TauSolveParams p, const double* __restrict__ u,
                                                             const double* __restrict__ q, double* __restrict__ L) {
    const TauData& td = p.td;
    const int N = td.N;
    const int tid = threadIdx.x, NT = TAU_THREADS;
    const double scale = 4.0 / (td.b - td.a);
    const long rs = (long)p.g.Nx * (2 * (p.g.Nz / 2 + 1));
    const long cs = rs * p.g.Ny;
    const int TM = p.TM, TT = 2 * TM;
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
        s_k[tid] = td.sc(TSC_LAMP)[qs];
        s_k[TM + tid] = td.sc(TSC_KXX)[qs];
        s_k[2 * TM + tid] = td.sc(TSC_KZZ)[qs];
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

int tau_setup_launch(const TauData& td, const ModeGeom& g, double lambda_t, int TM, cudaStream_t stream) {
    const size_t smem = ((size_t)3 * td.N * TM + 10 * TM) * sizeof(double);
    static size_t configured = 0;
    auto kfn = tau_setup_kernel;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((td.nq + TM - 1) / TM);
    CF_LAUNCH(kfn, grid, dim3(TAU_THREADS), smem, stream, td, g, lambda_t, TM);
    CF_KERNEL_CHECK();
    return 0;
}

int linear_launch(const TauSolveParams& p, const double* u, const double* q, double* L, cudaStream_t stream) {
    const int TT = 2 * p.TM;
    const size_t smem = ((size_t)5 * p.td.N * TT + 3 * p.TM) * sizeof(double) + p.TM * sizeof(long);
    static size_t configured = 0;
    auto kfn = linear_kernel;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((p.td.nq + p.TM - 1) / p.TM);
    CF_LAUNCH(kfn, grid, dim3(TAU_THREADS), smem, stream, p, u, q, L);
    CF_KERNEL_CHECK();
    return 0;
}

int tau_solve_launch(const TauSolveParams& p, cudaStream_t stream) {
    const int TT = 2 * p.TM;
    size_t smem = ((size_t)6 * p.td.N * TT + 10 * p.TM + 2 * TT) * sizeof(double) + p.TM * sizeof(long);
    const size_t smem00 = (size_t)9 * p.td.N * sizeof(double);
    if (smem00 > smem) smem = smem00;
    static size_t configured = 0;
    auto kfn = tau_solve_kernel;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int ntiles = (p.td.nq - 1 + p.TM - 1) / p.TM;
    dim3 grid(ntiles + 1);
    CF_LAUNCH(kfn, grid, dim3(TAU_THREADS), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}

}  // namespace cfgpu
