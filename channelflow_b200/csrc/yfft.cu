// Chebyshev y-transform as a shared-memory FFT: the same jobs as ygemm.cu (YGemmParams), for the profile lengths whose
// even extension 2(Ny-1) factors into 2, 3, 5.
//
// The transform of the reference (flowfield.cpp:1888-1987, FFTW REDFT00) is a DCT-I of length Ny = M+1.  The two real
// profiles of one Fourier mode (re, im) form one complex sequence z_n; its even extension e (length L = 2M, e_n = e_{L-n})
// has the transform  FFT(e)_j = z_0 + (-1)^j z_M + 2 sum_{0<n<M} z_n cos(pi j n / M),  real part = DCT of re, imaginary part
// = DCT of im.  So one complex FFT of length L per mode does both profiles; only its outputs 0..M are kept.
//   inverse : u_j = sum_n c_n cos(pi j n/M)                    -> e_0 = c_0, e_M = c_M, e_n = c_n/2 otherwise; u = FFT(e)
//   forward : c_n = w_n sum_j g_j x_j cos(pi j n/M)            -> e = x (even extension);  c_n = w_n FFT(e)_n,
//             w_n = 1/M (1/(2M) at n = 0, M), g_j = 2 (1 at the ends)      (flowfield.cpp:1913-1934)
//   d/dy    : coefficients d_n = (4/(b-a)) (1/2 at n = 0) sum_{m>n, m-n odd} m c_m   (chebyshev.cpp:672-697), formed in shared
//             memory by a chunked suffix sum over n before the second FFT of the same input.
// Against the DMMA contraction this is O(L log L) instead of O(Ny^2/2) flops per profile: at Ny = 257 the contraction is
// bound by the FP64 tensor pipe (1.65 ms inverse, 1.14 ms forward at 512x257x512), the FFT by memory latency (1.11 / 0.59 ms).
// Two kernels: yfft_kernel transforms the full even extension (length 2M); yfft_half_kernel -- the default -- reduces it to ONE
// complex FFT of length M per mode (see its comment) and also serves the forward jobs with a second input.
//
// CTA = C complex columns (modes) of one job; shared memory a[L][C] (in-place transform, digit-reversed input rows) plus the
// plan's twiddle and row tables.  A row piece of C = 8 columns is 128 contiguous bytes in HBM.
#include "fft_smem.cuh"
#include "ygemm.cuh"

namespace cfgpu {

namespace {

constexpr int YF_THREADS = 256;

// CH = rows per thread: thread (c, t) owns column c and the rows [t ch, t ch + ch), ch = ceil(N / (YF_THREADS / C)) <= CH.  It
// loads them into registers (a warp reads 32/C row pieces of 16 C contiguous bytes per instruction); the suffix sums of
// m c_m for the derivative need each coefficient twice.
// One CTA = one transform: the units of work are (job, value) and (job, derivative) pairs, numbered fastest along the grid
// so that the two CTAs that read the same coefficients run side by side (the second read hits L2).
struct YfftUnits {
    int n;
    unsigned char job[2 * YG_MAXJOB], der[2 * YG_MAXJOB];
};

template <int C, int CH>
__global__ void __launch_bounds__(YF_THREADS, 3) yfft_kernel(const YGemmParams p, const FftPlanDev pl, const double dscale,
                                                             const YfftUnits un) {
    const int N = p.N, M = N - 1, L = 2 * M;
    const int tid = threadIdx.x;
    const int unit = blockIdx.x % un.n;
    const long cblock = blockIdx.x / un.n;
    const YGemmJob& jb = p.job[un.job[unit]];
    const bool is_der = un.der[unit] != 0;
    constexpr int TPC = YF_THREADS / C;  // threads per column
    const int ntw = fft_plan_ntw(pl);
    double2* a = dyn_smem<double2>();      // [L][C]
    double2* stw = a + (size_t)L * C;      // [ntw] twiddles the passes touch
    double2* part = stw + ntw;             // [2][TPC][C] chunk sums of the derivative
    int* rev = reinterpret_cast<int*>(part + 2 * TPC * C);   // [L] digit-reversed rows
    const int c = tid % C, t = tid / C;
    const long col = cblock * C + c;                  // complex column
    const bool cvalid = 2 * col < p.ncols;
    const long dc = 2 * col;
    const long inoff = !cvalid ? 0 : (p.in_runstart ? p.in_runstart[dc / p.in_runlen] + dc % p.in_runlen : dc);
    const long outoff = !cvalid ? 0 : (p.out_runstart ? p.out_runstart[dc / p.out_runlen] + dc % p.out_runlen : dc);
    const double2 zero = make_double2(0.0, 0.0);
    const int ch = (N + TPC - 1) / TPC;
    const int n0 = t * ch;

    // the thread's rows first (the longest latency), the tables behind them
    double2 v[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int n = n0 + i;
        v[i] = zero;
        if (i < ch && n < N && cvalid) {
            const double* src = jb.in + (size_t)n * p.in_ld + inoff;
            v[i] = make_double2(src[0], src[1]);
        }
    }
    for (int i = tid; i < L; i += YF_THREADS) rev[i] = pl.rev[i];
    for (int i = tid; i < ntw; i += YF_THREADS) stw[i] = pl.tw[i];
    __syncthreads();

    auto store = [&](int m, int r, double2 x) {
        if (!cvalid) return;
        double* dst = (jb.out_rows[m] ? jb.out_rows[m][r] : jb.out[m] + (size_t)r * p.out_ld) + outoff;
        dst[0] = x.x;
        dst[1] = x.y;
    };
    // even extension of the thread's rows, halved away from the ends (inverse) or as they are (forward)
    auto scatter = [&](const double2 (&x)[CH], bool halve) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int n = n0 + i;
            if (i < ch && n < N) {
                const double h = (halve && n != 0 && n != M) ? 0.5 : 1.0;
                const double2 e = make_double2(h * x[i].x, h * x[i].y);
                a[(size_t)rev[n] * C + c] = e;
                if (n > 0 && n < M) a[(size_t)rev[L - n] * C + c] = e;
            }
        }
    };

    if (p.mode == 1) {
        // ---- forward: c_n = w_n FFT_n of the even extension of the physical profile
        scatter(v, false);
        __syncthreads();
        fft_smem_inplace<-1, false, true>(a, pl, stw, C, tid, YF_THREADS);
        const double w = 1.0 / M;
        for (int n = t; n < N; n += TPC) {
            const double wn = (n == 0 || n == M) ? 0.5 * w : w;
            const double2 x = a[(size_t)n * C + c];
            store(0, n, make_double2(wn * x.x, wn * x.y));
        }
        return;
    }

    // ---- inverse: values (matrix 0) or y-derivative (matrix 1)
    if (!is_der) {
        scatter(v, true);
        __syncthreads();
        fft_smem_inplace<-1, false, true>(a, pl, stw, C, tid, YF_THREADS);
        for (int j = t; j < N; j += TPC) store(0, j, a[(size_t)j * C + c]);
        return;
    }
    {
        // sums of m c_m over the even and the odd m of the thread's rows, the totals of the chunks above, the walk down
        double2 se = zero, so = zero;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int m = n0 + i;   // (rows past the end hold zeros)
            if (m & 1) { so.x += m * v[i].x; so.y += m * v[i].y; }
            else { se.x += m * v[i].x; se.y += m * v[i].y; }
        }
        part[(size_t)t * C + c] = se;
        part[(size_t)(TPC + t) * C + c] = so;
        __syncthreads();
        se = zero; so = zero;
        for (int tt = TPC - 1; tt > t; --tt) {
            const double2 pe = part[(size_t)tt * C + c], po = part[(size_t)(TPC + tt) * C + c];
            se.x += pe.x; se.y += pe.y;
            so.x += po.x; so.y += po.y;
        }
#pragma unroll
        for (int i = CH - 1; i >= 0; --i) {
            const int n = n0 + i;
            const double2 sm = (n & 1) ? se : so;   // sum over m > n of the other parity
            const double f = (n == 0) ? 0.5 * dscale : dscale;
            if (n & 1) { so.x += n * v[i].x; so.y += n * v[i].y; }
            else { se.x += n * v[i].x; se.y += n * v[i].y; }
            v[i] = make_double2(f * sm.x, f * sm.y);
        }
        scatter(v, true);
        __syncthreads();
        fft_smem_inplace<-1, false, true>(a, pl, stw, C, tid, YF_THREADS);
        const int mder = jb.mat0 == 0 ? 1 : 0;   // output slot of the derivative
        for (int j = t; j < N; j += TPC) store(mder, j, a[(size_t)j * C + c]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Half-length variant.  For the even extension above only M of the L = 2M transform points carry information; the classic
// reduction (a real-even sequence needs a real transform of half the length) carries over to the packed complex profiles
// because every step is linear:
//   y_j = (x_j + x_{M-j}) - 2 sin(pi j/M) (x_j - x_{M-j}),  j = 0 .. M-1           Y = FFT_M(y)
//   T_{2k}   = (Y_k + Y_{M-k}) / 2                                                    k = 0 .. M/2
//   T_{2k+1} = T_{2k-1} - D_k,  D_k = -i (Y_k - Y_{M-k}) / 2,                         k = 1 .. M/2-1
//   T_1      = x_0 - x_M + 2 sum_{0<j<M/2} cos(pi j/M) (x_j - x_{M-j})
// where T_k = x_0 + (-1)^k x_M + 2 sum_{0<j<M} x_j cos(pi j k/M) is what the full-length transform returns at k <= M.
// Thread (c, t) loads the row PAIRS (j, M-j) of its chunk of j <= M/2, so y_j, y_{M-j} and its share of T_1 come from its own
// registers; after the transform it owns a chunk of k: T_{2k}, D_k from a[k], a[M-k], the odd rows by a chunked prefix sum.
template <int C, int CH, bool TWO>   // CH >= pairs per thread; TWO: the launch has forward jobs with a second input
__global__ void __launch_bounds__(YF_THREADS, 3) yfft_half_kernel(const YGemmParams p, const FftPlanDev pl, const double2* __restrict__ twL,
                                                                  const double dscale, const YfftUnits un) {
    const int N = p.N, M = N - 1, H = M / 2;
    const int tid = threadIdx.x;
    const int unit = blockIdx.x % un.n;
    const long cblock = blockIdx.x / un.n;
    const YGemmJob& jb = p.job[un.job[unit]];
    const bool is_der = un.der[unit] != 0;
    constexpr int TPC = YF_THREADS / C;
    const int ntw = fft_plan_ntw(pl);
    double2* a = dyn_smem<double2>();      // [M][C]
    double2* stw = a + (size_t)M * C;      // [ntw] twiddles of the length-M plan
    double2* part = stw + ntw;             // [4 TPC][C] chunk sums
    double2* cbuf = part + 4 * TPC * C;    // [N][C] d/dy coefficients of the second input (forward with in2 only)
    int* rev = reinterpret_cast<int*>(cbuf + (p.two_inputs ? (size_t)N * C : 0));   // [M]
    const int c = tid % C, t = tid / C;
    const long col = cblock * C + c;
    const bool cvalid = 2 * col < p.ncols;
    const unsigned dc = (unsigned)(2 * col);   // (the launcher checks ncols < 2^31: 32-bit divisions)
    const long inoff = !cvalid ? 0 : (p.in_runstart ? p.in_runstart[dc / (unsigned)p.in_runlen] + dc % (unsigned)p.in_runlen : (long)dc);
    const long outoff = !cvalid ? 0 : (p.out_runstart ? p.out_runstart[dc / (unsigned)p.out_runlen] + dc % (unsigned)p.out_runlen : (long)dc);
    const double2 zero = make_double2(0.0, 0.0);
    const int ch = (H + 1 + TPC - 1) / TPC;   // pairs per thread
    const int j0 = t * ch;

    // Forward jobs with a second input (f = F x + cd D F x2, the divergence / skew-symmetric forms): pass 0 transforms x2 and
    // leaves the coefficients of its y-derivative in cbuf, pass 1 transforms x and adds them.  Everything else: pass 1 only.
    const bool two = TWO && p.mode == 1 && jb.in2 != nullptr;   // (TWO = false: one pass, no trace of the second in the code)
    for (int pass = two ? 0 : 1; pass < 2; ++pass) {
    const double* __restrict__ inp = pass == 0 ? jb.in2 : jb.in;
    // rows j (lo) and M-j (hi) of the thread's pairs; the self-paired row M/2 is loaded once
    double2 lo[CH], hi[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int j = j0 + i;
        lo[i] = zero; hi[i] = zero;
        if (i < ch && j <= H && cvalid) {
            const double* s0 = inp + (size_t)j * p.in_ld + inoff;
            lo[i] = make_double2(s0[0], s0[1]);
            if (j < H) {
                const double* s1 = inp + (size_t)(M - j) * p.in_ld + inoff;
                hi[i] = make_double2(s1[0], s1[1]);
            }
        }
    }
    if (pass == (two ? 0 : 1)) {   // first pass: the plan's tables
        for (int i = tid; i < M; i += YF_THREADS) rev[i] = pl.rev[i];
        for (int i = tid; i < ntw; i += YF_THREADS) stw[i] = pl.tw[i];
    }

    if (is_der) {
        // coefficients of the derivative in place of the input: suffix sums of m c_m by parity (see the full-length kernel).
        // The thread's rows form two runs of n: [j0, j0+ch) and (M-j0-ch, M-j0]; chunk index in n order: t and 2 TPC - 1 - t.
        double2 se = zero, so = zero, he = zero, ho = zero;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int m = j0 + i, mh = M - m;
            if (m & 1) { so.x += m * lo[i].x; so.y += m * lo[i].y; } else { se.x += m * lo[i].x; se.y += m * lo[i].y; }
            if (mh & 1) { ho.x += mh * hi[i].x; ho.y += mh * hi[i].y; } else { he.x += mh * hi[i].x; he.y += mh * hi[i].y; }
        }
        // part[(2 q) C + c] even sum of chunk q, part[(2 q + 1) C + c] odd sum; 2 TPC chunks  (needs [4 TPC][C])
        part[(size_t)(2 * t) * C + c] = se;
        part[(size_t)(2 * t + 1) * C + c] = so;
        part[(size_t)(2 * (2 * TPC - 1 - t)) * C + c] = he;
        part[(size_t)(2 * (2 * TPC - 1 - t) + 1) * C + c] = ho;
        __syncthreads();
        // totals above the high run, then the walk down the high run (n descending = i ascending)
        se = zero; so = zero;
        for (int q = 2 * TPC - 1; q > 2 * TPC - 1 - t; --q) {
            const double2 pe = part[(size_t)(2 * q) * C + c], po = part[(size_t)(2 * q + 1) * C + c];
            se.x += pe.x; se.y += pe.y; so.x += po.x; so.y += po.y;
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int n = M - (j0 + i);
            const double2 sm = (n & 1) ? se : so;
            const double2 v = hi[i];
            if (n & 1) { so.x += n * v.x; so.y += n * v.y; } else { se.x += n * v.x; se.y += n * v.y; }
            hi[i] = make_double2(dscale * sm.x, dscale * sm.y);   // (n >= M/2 > 0: no halving)
        }
        // between the runs: chunks 2 TPC - 1 - t - 1 ... t + 1
        for (int q = 2 * TPC - 2 - t; q > t; --q) {
            const double2 pe = part[(size_t)(2 * q) * C + c], po = part[(size_t)(2 * q + 1) * C + c];
            se.x += pe.x; se.y += pe.y; so.x += po.x; so.y += po.y;
        }
#pragma unroll
        for (int i = CH - 1; i >= 0; --i) {
            const int n = j0 + i;
            const double2 sm = (n & 1) ? se : so;
            const double f = (n == 0) ? 0.5 * dscale : dscale;
            const double2 v = lo[i];
            if (n & 1) { so.x += n * v.x; so.y += n * v.y; } else { se.x += n * v.x; se.y += n * v.y; }
            lo[i] = make_double2(f * sm.x, f * sm.y);
        }
        __syncthreads();   // part is reused below
    }

    // inverse: coefficients halved away from the ends (the even extension of the full-length kernel)
    if (p.mode == 0) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int j = j0 + i;
            if (j != 0) { lo[i].x *= 0.5; lo[i].y *= 0.5; hi[i].x *= 0.5; hi[i].y *= 0.5; }
        }
    }
    __syncthreads();   // rev, stw in place
    // y_j, y_{M-j} and the thread's share of T_1
    double2 t1 = zero;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int j = j0 + i;
        if (i < ch && j <= H) {
            if (j == H) {
                a[rev[j] * C + c] = make_double2(2.0 * lo[i].x, 2.0 * lo[i].y);
            } else {
                const double2 w = twL[j];   // cos(pi j/M) - i sin(pi j/M)
                const double2 S = make_double2(lo[i].x + hi[i].x, lo[i].y + hi[i].y), D = make_double2(lo[i].x - hi[i].x, lo[i].y - hi[i].y);
                const double s2 = -2.0 * w.y;   // 2 sin(pi j/M)
                a[rev[j] * C + c] = make_double2(S.x - s2 * D.x, S.y - s2 * D.y);
                if (j > 0) a[rev[M - j] * C + c] = make_double2(S.x + s2 * D.x, S.y + s2 * D.y);
                const double cw = (j == 0) ? 1.0 : 2.0 * w.x;
                t1.x += cw * D.x; t1.y += cw * D.y;
            }
        }
    }
    constexpr int NWARP = YF_THREADS / 32;
    const int warp = tid >> 5;
#pragma unroll
    for (int o = C; o < 32; o <<= 1) {   // the threads of a column sit C lanes apart
        t1.x += __shfl_xor_sync(0xffffffffu, t1.x, o);
        t1.y += __shfl_xor_sync(0xffffffffu, t1.y, o);
    }
    if ((tid & 31) < C) part[warp * C + c] = t1;
    __syncthreads();
    fft_smem_inplace<-1, false, true>(a, pl, stw, C, tid, YF_THREADS);
    // T_1 of the column
    t1 = zero;
#pragma unroll
    for (int q = 0; q < NWARP; ++q) { const double2 v = part[q * C + c]; t1.x += v.x; t1.y += v.y; }
    __syncthreads();   // everybody has read part
    // the thread's chunk of k in [1, H): D_k and their sum
    const int ck = (H + TPC - 1) / TPC;
    const int k0 = t * ck;
    double2 ds = zero;
    const int kend = (k0 + ck < H) ? k0 + ck : H;
    for (int k = (k0 > 1 ? k0 : 1); k < kend; ++k) {
        const double2 yk = a[k * C + c], ym = a[(M - k) * C + c];
        ds.x += 0.5 * (yk.y - ym.y);    // D_k = -i (Y_k - Y_{M-k}) / 2
        ds.y -= 0.5 * (yk.x - ym.x);
    }
    part[t * C + c] = ds;
    __syncthreads();
    double2 odd = t1;   // T_{2k-1} entering the chunk: T_1 - sum_{k' < k0} D_k'
    for (int q = 0; q < t; ++q) { const double2 v = part[q * C + c]; odd.x -= v.x; odd.y -= v.y; }
    const int mslot = (p.mode == 0 && is_der && jb.mat0 == 0) ? 1 : 0;
    const double w = 1.0 / M;
    double* const* const orows = jb.out_rows[mslot];
    double* const obase = jb.out[mslot] + outoff;
    const bool fwd = p.mode == 1;
    auto store = [&](int r, double2 x) {
        if (fwd) { const double wn = (r == 0 || r == M) ? 0.5 * w : w; x.x *= wn; x.y *= wn; }
        if (pass == 0) { cbuf[r * C + c] = x; return; }
        if (two) { const double2 d = cbuf[r * C + c]; x.x += d.x; x.y += d.y; }
        if (!cvalid) return;
        double* dst = orows ? orows[r] + outoff : obase + (size_t)r * p.out_ld;
        dst[0] = x.x;
        dst[1] = x.y;
    };
    for (int k = k0; k < kend; ++k) {
        const double2 yk = a[k * C + c], ym = a[(k == 0 ? 0 : M - k) * C + c];
        store(2 * k, make_double2(0.5 * (yk.x + ym.x), 0.5 * (yk.y + ym.y)));
        if (k > 0) { odd.x -= 0.5 * (yk.y - ym.y); odd.y += 0.5 * (yk.x - ym.x); }
        store(2 * k + 1, odd);
    }
    if (t == TPC - 1) store(M, a[H * C + c]);   // T_M = Y_{M/2}
    if (pass == 0) {
        // cbuf: coefficients c of x2 -> d_n = in2_scale (4/(b-a)) (1/2 at n = 0) sum_{m>n, m-n odd} m c_m, in place
        __syncthreads();
        const double d2 = p.in2_scale * dscale;
        const int chn = (N + TPC - 1) / TPC;
        const int n0 = t * chn, n1 = (n0 + chn < N) ? n0 + chn : N;
        double2 se = zero, so = zero;
        for (int m = n0; m < n1; ++m) {
            const double2 v = cbuf[m * C + c];
            if (m & 1) { so.x += m * v.x; so.y += m * v.y; } else { se.x += m * v.x; se.y += m * v.y; }
        }
        part[(2 * t) * C + c] = se;
        part[(2 * t + 1) * C + c] = so;
        __syncthreads();
        se = zero; so = zero;
        for (int q = TPC - 1; q > t; --q) {
            const double2 pe = part[(2 * q) * C + c], po = part[(2 * q + 1) * C + c];
            se.x += pe.x; se.y += pe.y; so.x += po.x; so.y += po.y;
        }
        for (int n = n1 - 1; n >= n0; --n) {
            const double2 sm = (n & 1) ? se : so;
            const double f = (n == 0) ? 0.5 * d2 : d2;
            const double2 v = cbuf[n * C + c];
            if (n & 1) { so.x += n * v.x; so.y += n * v.y; } else { se.x += n * v.x; se.y += n * v.y; }
            cbuf[n * C + c] = make_double2(f * sm.x, f * sm.y);
        }
        __syncthreads();
    }
    }   // pass
}

template <int C, int CH>
int launch_half_c(const YGemmParams& p, const FftPlanDev& plM, const double2* twL, double dscale, cudaStream_t stream) {
    const int M = p.N - 1;
    const size_t smem = ((size_t)M * C + fft_plan_ntw(plM) + 4 * YF_THREADS + (p.two_inputs ? (size_t)p.N * C : 0)) * sizeof(double2) +
                        (size_t)M * sizeof(int);
    if (smem > 200 * 1024) return -1;
    auto kfn = p.two_inputs ? yfft_half_kernel<C, CH, true> : yfft_half_kernel<C, CH, false>;
    static size_t configured[2] = {0, 0};
    if (smem > configured[p.two_inputs ? 1 : 0]) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[p.two_inputs ? 1 : 0] = smem;
    }
    YfftUnits un;
    un.n = 0;
    for (int j = 0; j < p.njobs; ++j) {
        if (p.mode == 1 || p.job[j].mat0 == 0) { un.job[un.n] = (unsigned char)j; un.der[un.n++] = 0; }
        if (p.mode == 0 && p.job[j].mat0 + p.job[j].nmat > 1) { un.job[un.n] = (unsigned char)j; un.der[un.n++] = 1; }
    }
    const long ccols = p.ncols / 2;
    const long nblk = ((ccols + C - 1) / C) * un.n;
    if (nblk > 0x7fffffffL) return -1;
    CF_LAUNCH(kfn, dim3((unsigned)nblk), dim3(YF_THREADS), smem, stream, p, plM, twL, dscale, un);
    CF_KERNEL_CHECK();
    return 0;
}

template <int C>
int launch_half_ch(const YGemmParams& p, const FftPlanDev& plM, const double2* twL, double dscale, cudaStream_t stream) {
    const int TPC = YF_THREADS / C;
    const int ch = ((p.N - 1) / 2 + 1 + TPC - 1) / TPC;
    if (ch <= 2) return launch_half_c<C, 2>(p, plM, twL, dscale, stream);
    if (ch <= 3) return launch_half_c<C, 3>(p, plM, twL, dscale, stream);
    if (ch <= 5) return launch_half_c<C, 5>(p, plM, twL, dscale, stream);
    if (ch <= 9) return launch_half_c<C, 9>(p, plM, twL, dscale, stream);
    return -1;
}

template <int C, int CH>
int launch_c(const YGemmParams& p, const FftPlanDev& pl, double dscale, cudaStream_t stream) {
    const int N = p.N, L = 2 * (N - 1);
    const size_t smem = ((size_t)L * C + fft_plan_ntw(pl) + 2 * YF_THREADS) * sizeof(double2) + (size_t)L * sizeof(int);
    auto kfn = yfft_kernel<C, CH>;
    static size_t configured = 0;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    YfftUnits un;
    un.n = 0;
    for (int j = 0; j < p.njobs; ++j) {
        if (p.mode == 1 || p.job[j].mat0 == 0) { un.job[un.n] = (unsigned char)j; un.der[un.n++] = 0; }
        if (p.mode == 0 && p.job[j].mat0 + p.job[j].nmat > 1) { un.job[un.n] = (unsigned char)j; un.der[un.n++] = 1; }
    }
    const long ccols = p.ncols / 2;
    const long nblk = ((ccols + C - 1) / C) * un.n;
    if (nblk > 0x7fffffffL) return -1;
    CF_LAUNCH(kfn, dim3((unsigned)nblk), dim3(YF_THREADS), smem, stream, p, pl, dscale, un);
    CF_KERNEL_CHECK();
    return 0;
}

template <int C>
int launch_ch(const YGemmParams& p, const FftPlanDev& pl, double dscale, cudaStream_t stream) {
    const int ch = (p.N + YF_THREADS / C - 1) / (YF_THREADS / C);
    if (ch <= 3) return launch_c<C, 3>(p, pl, dscale, stream);
    if (ch <= 5) return launch_c<C, 5>(p, pl, dscale, stream);
    if (ch <= 9) return launch_c<C, 9>(p, pl, dscale, stream);
    if (ch <= 17) return launch_c<C, 17>(p, pl, dscale, stream);
    return -1;
}

}  // namespace

bool yfft_length_supported(int N) {
    if (N < 3) return false;
    int m = 2 * (N - 1);
    for (int r : {2, 3, 5})
        while (m % r == 0) m /= r;
    return m == 1;
}

// p as for ygemm_launch.  Returns -1 when the parameters need the contraction.
int yfft_launch(const YGemmParams& p_in, const FftPlanDev& pl, double a, double b, cudaStream_t stream) {
    if (pl.N != 2 * (p_in.N - 1)) return -1;
    YGemmParams p = p_in;
    p.two_inputs = 0;
    for (int j = 0; j < p.njobs; ++j) {
        if (p.mode == 1 && (p.job[j].nmat != 1 || p.job[j].mat0 != 0)) return -1;
        if (p.job[j].in2) p.two_inputs = 1;
    }
    if (p.two_inputs && (p.mode != 1 || p.in2_scale == 0.0)) return -1;
    if ((p.in_runstart && (p.in_runlen & 1)) || (p.out_runstart && (p.out_runlen & 1)) || (p.in_ld & 1) || (p.out_ld & 1) || (p.ncols & 1)) return -1;
    if (p.ncols >= 0x7fffffffL) return -1;
    if (p.ncols <= 0 || p.njobs <= 0) return 0;
    static const int cw = getenv("CF_YFFT_C") ? atoi(getenv("CF_YFFT_C")) : 8;
    const double dscale = 4.0 / (b - a);
    // half-length transform (CF_YFFT_HALF=0: the full even extension); needs the plan of length Ny-1 next to pl's twiddles
    static const int half = getenv("CF_YFFT_HALF") ? atoi(getenv("CF_YFFT_HALF")) : 1;
    if (half && p.fft_half && p.fft_half->N == p.N - 1 && (p.N - 1) % 2 == 0) {
        int rc = -1;
        if (cw == 16) rc = launch_half_ch<16>(p, *p.fft_half, pl.tw, dscale, stream);
        if (rc < 0 && cw != 4) rc = launch_half_ch<8>(p, *p.fft_half, pl.tw, dscale, stream);
        if (rc < 0) rc = launch_half_ch<4>(p, *p.fft_half, pl.tw, dscale, stream);
        if (rc >= 0) return rc;
    }
    if (p.two_inputs) return -1;   // (the full-length kernel has no second-input path)
    const size_t need8 = ((size_t)pl.N * 8 + pl.N + 2 * YF_THREADS) * sizeof(double2);
    if (cw != 4 && need8 <= 200 * 1024) {
        const int rc = launch_ch<8>(p, pl, dscale, stream);
        if (rc >= 0) return rc;
    }
    const size_t need4 = ((size_t)pl.N * 4 + pl.N + 2 * YF_THREADS) * sizeof(double2);
    if (need4 > 200 * 1024) return -1;
    return launch_ch<4>(p, pl, dscale, stream);
}

}  // namespace cfgpu
