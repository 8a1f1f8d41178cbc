// Shared-memory Stockham FFT used by the x- and z-pass kernels (replaces the FFTW rank-2 r2c/c2r plans of the
// reference, flowfield.cpp:624-649, executed at :1853 and :1883).
//
// Data layout in shared memory: buf[n * C + c], n = position along the transformed direction, c = one of C
// independent transforms handled by the CTA ("columns").  Threads are mapped (butterfly j, column c) with c
// fastest, so every load/store of a pass touches consecutive 16-byte words.  Mixed radix 8/4/2/3/5, autosort
// (no bit reversal), ping-pong between two buffers.  Twiddles come from a table exp(-2 pi i t / N) computed on
// the host in long double (read through the read-only cache); the inverse transform conjugates them.
// Unnormalised, FFTW sign convention: DIR=-1 forward (exp(-i..)), DIR=+1 backward.
#pragma once
#include "cf_common.cuh"

namespace cfgpu {

constexpr int FFT_MAXPASS = 16;

struct FftPlanDev {
    int N;
    int npass;
    int radix[FFT_MAXPASS];
    const double2* tw;  // N entries
    const int* rev;     // N entries: fft_plan_rev(n), the digit-reversed row of the in-place transforms
};

template <int DIR>
__device__ __forceinline__ double2 tw_mul(double2 v, double2 w) {
    // v * w (forward) or v * conj(w) (backward)
    if (DIR < 0) return make_double2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
    return make_double2(v.x * w.x + v.y * w.y, v.y * w.x - v.x * w.y);
}
// w[r] = w1^r, r = 1..R-1, by a shallow product tree: one table load per butterfly instead of R-1 (the loads, not
// the FP64 pipe, are what the shared-memory transforms are short of); costs a few ulp in the higher powers
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <int R>
__device__ __forceinline__ void twiddle_powers(double2 w1, double2 (&w)[R]) {
    w[1] = w1;
    if (R > 2) w[2] = cmul(w1, w1);
    if (R > 3) w[3] = cmul(w[2], w1);
    if (R > 4) w[4] = cmul(w[2], w[2]);
#pragma unroll
    for (int r = 5; r < R; ++r) w[r] = cmul(w[4], w[r - 4]);
}
// multiply by -i (forward) / +i (backward)
template <int DIR>
__device__ __forceinline__ double2 rot90(double2 v) {
    if (DIR < 0) return make_double2(v.y, -v.x);
    return make_double2(-v.y, v.x);
}
__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 operator-(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 operator*(double s, double2 a) { return make_double2(s * a.x, s * a.y); }

// In-register radix-R butterfly: v[q] <- sum_r v[r] W_R^{rq} (W_R = exp(-+2 pi i / R) for DIR = -1 / +1).
template <int DIR, int R>
__device__ __forceinline__ void radix_butterfly(double2 (&v)[R]) {
    if (R == 2) {
        double2 t = v[0] - v[1];
        v[0] = v[0] + v[1];
        v[1] = t;
    } else if (R == 4) {
        double2 s02 = v[0] + v[2], d02 = v[0] - v[2], s13 = v[1] + v[3], d13 = rot90<DIR>(v[1] - v[3]);
        v[0] = s02 + s13;
        v[1] = d02 + d13;
        v[2] = s02 - s13;
        v[3] = d02 - d13;
    } else if (R == 8) {
        // radix-2 split (r, r+4), twiddles W8^r on the odd half, then two radix-4 butterflies
        const double h = 0.70710678118654752440084436210485;
        double2 a[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            a[r] = v[r] + v[r + 4];
            a[r + 4] = v[r] - v[r + 4];
        }
        {
            const double2 t5 = a[5], t7 = a[7];
            if (DIR < 0) {
                a[5] = make_double2(h * (t5.x + t5.y), h * (t5.y - t5.x));
                a[7] = make_double2(h * (t7.y - t7.x), -h * (t7.x + t7.y));
            } else {
                a[5] = make_double2(h * (t5.x - t5.y), h * (t5.x + t5.y));
                a[7] = make_double2(-h * (t7.x + t7.y), h * (t7.x - t7.y));
            }
            a[6] = rot90<DIR>(a[6]);
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const double2 b0 = a[4 * o], b1 = a[4 * o + 1], b2 = a[4 * o + 2], b3 = a[4 * o + 3];
            const double2 s02 = b0 + b2, d02 = b0 - b2, s13 = b1 + b3, d13 = rot90<DIR>(b1 - b3);
            v[o] = s02 + s13;
            v[o + 2] = d02 + d13;
            v[o + 4] = s02 - s13;
            v[o + 6] = d02 - d13;
        }
    } else if (R == 3) {
        const double wi = (DIR < 0 ? -1.0 : 1.0) * 0.86602540378443864676372317075294;  // sin(2 pi/3)
        double2 s = v[1] + v[2], d = v[1] - v[2];
        double2 t = make_double2(v[0].x - 0.5 * s.x, v[0].y - 0.5 * s.y);
        double2 u = make_double2(-wi * d.y, wi * d.x);
        v[0] = v[0] + s;
        v[1] = t + u;
        v[2] = t - u;
    } else if (R == 5) {
        const double c1 = 0.30901699437494742410229341718282, c2 = -0.80901699437494742410229341718282;
        const double sg = (DIR < 0 ? -1.0 : 1.0);
        const double s1 = sg * 0.95105651629515357211643933337938, s2 = sg * 0.58778525229247312916870595463907;
        double2 s14 = v[1] + v[4], d14 = v[1] - v[4], s23 = v[2] + v[3], d23 = v[2] - v[3];
        double2 t1 = make_double2(v[0].x + c1 * s14.x + c2 * s23.x, v[0].y + c1 * s14.y + c2 * s23.y);
        double2 t2 = make_double2(v[0].x + c2 * s14.x + c1 * s23.x, v[0].y + c2 * s14.y + c1 * s23.y);
        double2 q1 = make_double2(s1 * d14.x + s2 * d23.x, s1 * d14.y + s2 * d23.y);
        double2 q2 = make_double2(s2 * d14.x - s1 * d23.x, s2 * d14.y - s1 * d23.y);
        double2 u1 = make_double2(-q1.y, q1.x), u2 = make_double2(-q2.y, q2.x);
        v[0] = v[0] + s14 + s23;
        v[1] = t1 + u1;
        v[4] = t1 - u1;
        v[2] = t2 + u2;
        v[3] = t2 - u2;
    }
}

// One Stockham pass of radix R over all C columns. Ns = product of the radices of the previous passes.
// TWPOW: twiddles w^(r k) from one table load and a product tree (pays when the table is in global memory / L1;
// with a shared-memory table and few columns per CTA the R-1 loads are the cheaper way -- both measured, x-pass at 512)
template <int DIR, int R, bool TWPOW>
__device__ __forceinline__ void fft_pass(const double2* __restrict__ a, double2* __restrict__ b, const FftPlanDev& pl,
                                         const double2* __restrict__ tw, int Ns, int C, int tid, int nthreads) {
    const int N = pl.N;
    const int nb = N / R;
    const int tstep = N / (Ns * R);
    const int total = nb * C;
    for (int idx = tid; idx < total; idx += nthreads) {
        const int j = idx / C, c = idx - j * C;
        const int k = j % Ns;
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = a[(j + r * nb) * C + c];
        if (Ns > 1) {
            if (TWPOW) {
                double2 w[R];
                twiddle_powers<R>(tw[k * tstep], w);
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = tw_mul<DIR>(v[r], w[r]);
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = tw_mul<DIR>(v[r], tw[r * k * tstep]);
            }
        }
        radix_butterfly<DIR, R>(v);
        const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) b[(j0 + r * Ns) * C + c] = v[r];
    }
}

// Full transform of C columns. `a` holds the input; returns the buffer (a or b) holding the result.
// Ends with a __syncthreads(); the caller must have synchronised after filling `a`.
// `tw` = twiddle table (pl.tw in HBM/L1, or a copy the caller staged in shared memory).
template <int DIR, bool TWPOW = true>
__device__ __forceinline__ double2* fft_smem(double2* a, double2* b, const FftPlanDev& pl, const double2* tw, int C, int tid,
                                             int nthreads) {
    int Ns = 1;
    for (int p = 0; p < pl.npass; ++p) {
        const int R = pl.radix[p];
        if (R == 8) fft_pass<DIR, 8, TWPOW>(a, b, pl, tw, Ns, C, tid, nthreads);
        else if (R == 4) fft_pass<DIR, 4, TWPOW>(a, b, pl, tw, Ns, C, tid, nthreads);
        else if (R == 2) fft_pass<DIR, 2, TWPOW>(a, b, pl, tw, Ns, C, tid, nthreads);
        else if (R == 3) fft_pass<DIR, 3, TWPOW>(a, b, pl, tw, Ns, C, tid, nthreads);
        else fft_pass<DIR, 5, TWPOW>(a, b, pl, tw, Ns, C, tid, nthreads);
        __syncthreads();
        double2* t = a;
        a = b;
        b = t;
        Ns *= R;
    }
    return a;
}

// ---------------------------------------------------------------------------------------------------------------
// In-place block transform of C columns, a[n * C + c], any plan (radix 8/4/2/3/5): a butterfly reads and writes the same
// R rows, so one buffer is enough (half the shared memory of the Stockham version: twice the columns per CTA).
//   DIF = false (decimation in time):      input row n at fft_plan_rev(pl, n), output in natural order
//   DIF = true  (decimation in frequency): input in natural order, output row k at fft_plan_rev(pl, k)
// With C a multiple of 8 every quarter-warp touches one contiguous 128-byte piece of a row: no bank conflicts at any stride.
// (Used by the inverse x-pass.  The DIF variant in the forward x-pass measured slower than the Stockham transform there
// -- 0.92 against 0.67 ms at 512x257x512 -- so that kernel keeps two buffers; the variant stays for the next attempt.)
// number of leading twiddle-table entries the product-tree passes touch: tw[k * tstep] with k * tstep < N / R_p
__host__ __device__ inline int fft_plan_ntw(const FftPlanDev& pl) {
    int rmin = 0;
    for (int p = 1; p < pl.npass; ++p)
        if (rmin == 0 || pl.radix[p] < rmin) rmin = pl.radix[p];
    return rmin ? pl.N / rmin : 1;
}
__host__ __device__ inline int fft_plan_rev(const FftPlanDev& pl, int n) {
    int pos = 0, span = pl.N;
    for (int p = pl.npass - 1; p >= 0; --p) {
        const int R = pl.radix[p];
        span /= R;
        pos += (n % R) * span;
        n /= R;
    }
    return pos;
}

template <int DIR, int R, bool DIF, bool TWPOW>
__device__ __forceinline__ void fft_inplace_pass(double2* __restrict__ a, const FftPlanDev& pl, const double2* __restrict__ tw,
                                                 int Ns, int C, int tid, int nthreads) {
    const int N = pl.N;
    const int nb = N / R;
    const int tstep = N / (Ns * R);
    const int total = nb * C;
    const bool pow2 = ((C & (C - 1)) | (Ns & (Ns - 1))) == 0;   // CTA-uniform: shifts instead of integer divisions
    const int cs = __ffs(C) - 1, nss = __ffs(Ns) - 1;
    for (int idx = tid; idx < total; idx += nthreads) {
        int j, c, k, jb;
        if (pow2) { j = idx >> cs; c = idx & (C - 1); k = j & (Ns - 1); jb = j >> nss; }
        else { j = idx / C; c = idx - j * C; k = j % Ns; jb = j / Ns; }
        double2* e = a + (size_t)(jb * Ns * R + k) * C + c;
        const int es = Ns * C;
        double2 v[R], w[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = e[r * es];
        if (Ns > 1) {
            if (TWPOW) twiddle_powers<R>(tw[k * tstep], w);
            else {
#pragma unroll
                for (int r = 1; r < R; ++r) w[r] = tw[r * k * tstep];
            }
            if (!DIF) {
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = tw_mul<DIR>(v[r], w[r]);
            }
        }
        radix_butterfly<DIR, R>(v);
        if (DIF && Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = tw_mul<DIR>(v[r], w[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) e[r * es] = v[r];
    }
}

// Ends with a __syncthreads(); the caller must have synchronised after filling `a`.
template <int DIR, bool DIF, bool TWPOW>
__device__ __forceinline__ void fft_smem_inplace(double2* a, const FftPlanDev& pl, const double2* tw, int C, int tid, int nthreads) {
    int Ns = 1;
    if (DIF)
        for (int p = 0; p < pl.npass; ++p) Ns *= pl.radix[p];
    for (int q = 0; q < pl.npass; ++q) {
        const int p = DIF ? pl.npass - 1 - q : q;
        const int R = pl.radix[p];
        if (DIF) Ns /= R;
        if (R == 8) fft_inplace_pass<DIR, 8, DIF, TWPOW>(a, pl, tw, Ns, C, tid, nthreads);
        else if (R == 4) fft_inplace_pass<DIR, 4, DIF, TWPOW>(a, pl, tw, Ns, C, tid, nthreads);
        else if (R == 2) fft_inplace_pass<DIR, 2, DIF, TWPOW>(a, pl, tw, Ns, C, tid, nthreads);
        else if (R == 3) fft_inplace_pass<DIR, 3, DIF, TWPOW>(a, pl, tw, Ns, C, tid, nthreads);
        else fft_inplace_pass<DIR, 5, DIF, TWPOW>(a, pl, tw, Ns, C, tid, nthreads);
        __syncthreads();
        if (!DIF) Ns *= R;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-private transform: one warp owns one length-N line in shared memory (skewed, addr(i) = i + i/8, so that the
// stride-R stores of the first pass are bank-conflict free), in place: every lane reads the inputs of its butterflies
// into registers, __syncwarp, writes the outputs.  No block barrier inside the transform.  Needs N/R <= 32*(16/R)
// butterflies per pass (N <= 512 for radix 8/4/2, <= 480 with radix 3/5), see warp_fft_supported.
__host__ __device__ inline int fft_skew(int i) { return i + (i >> 3); }
__host__ __device__ inline int fft_skew_len(int N) { return N + (N >> 3) + 1; }

template <int DIR, int R>
__device__ __forceinline__ void warp_fft_pass(double2* __restrict__ buf, const double2* __restrict__ tw, int N, int Ns, int lane) {
    constexpr int NBMAX = 16 / R;
    const int nb = N / R;
    const int tstep = N / (Ns * R);
    double2 v[NBMAX][R];
#pragma unroll
    for (int i = 0; i < NBMAX; ++i) {
        const int b = lane + 32 * i;
        if (b < nb) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[i][r] = buf[fft_skew(b + r * nb)];
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NBMAX; ++i) {
        const int b = lane + 32 * i;
        if (b < nb) {
            const int k = b % Ns;
            if (Ns > 1) {
                double2 w[R];
                twiddle_powers<R>(tw[k * tstep], w);
#pragma unroll
                for (int r = 1; r < R; ++r) v[i][r] = tw_mul<DIR>(v[i][r], w[r]);
            }
            radix_butterfly<DIR, R>(v[i]);
            const int j0 = (b / Ns) * Ns * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) buf[fft_skew(j0 + r * Ns)] = v[i][r];
        }
    }
    __syncwarp();
}

template <int DIR>
__device__ __forceinline__ void warp_fft(double2* buf, const FftPlanDev& pl, int lane) {
    int Ns = 1;
    for (int p = 0; p < pl.npass; ++p) {
        const int R = pl.radix[p];
        if (R == 8) warp_fft_pass<DIR, 8>(buf, pl.tw, pl.N, Ns, lane);
        else if (R == 4) warp_fft_pass<DIR, 4>(buf, pl.tw, pl.N, Ns, lane);
        else if (R == 2) warp_fft_pass<DIR, 2>(buf, pl.tw, pl.N, Ns, lane);
        else if (R == 3) warp_fft_pass<DIR, 3>(buf, pl.tw, pl.N, Ns, lane);
        else warp_fft_pass<DIR, 5>(buf, pl.tw, pl.N, Ns, lane);
        Ns *= R;
    }
}

// Compile-time plan for the power-of-two lengths (radix 8 passes, then one radix 4 or 2): straight-line code, the
// butterflies' register arrays of the different passes are never live together.
template <int DIR, int N>
__device__ __forceinline__ void warp_fft_pow2(double2* buf, const double2* tw, int lane) {
    static_assert(N >= 8 && N <= 512 && (N & (N - 1)) == 0, "power of two, 8..512");
    int Ns = 1;
    constexpr int L = N == 512 ? 9 : N == 256 ? 8 : N == 128 ? 7 : N == 64 ? 6 : N == 32 ? 5 : N == 16 ? 4 : 3;
    constexpr int N8 = L / 3, REM = L % 3;
#pragma unroll
    for (int p = 0; p < N8; ++p) {
        warp_fft_pass<DIR, 8>(buf, tw, N, Ns, lane);
        Ns *= 8;
    }
    if (REM == 2) warp_fft_pass<DIR, 4>(buf, tw, N, Ns, lane);
    if (REM == 1) warp_fft_pass<DIR, 2>(buf, tw, N, Ns, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// In-place warp-private transforms with digit-reversed data on one side (power-of-two N, radix-8 passes of stride
// 1, 8, 64 and a last radix-4/2 pass): a butterfly reads and writes the same R locations, so a lane needs one
// butterfly in registers at a time (the Stockham variant above holds all of its butterflies of a pass).
//   warp_fft_dit: input at fft_digit_rev<N>(n), output in natural order  (decimation in time, twiddles before)
//   warp_fft_dif: input in natural order, output k at fft_digit_rev<N>(k) (decimation in frequency, twiddles after)
// Callers write/read the reversed side with the index they compute anyway.  Skew addr(i) = i + i/8 + i/64 keeps every
// quarter-warp of the stride-1/8/64 accesses AND of consecutive reversed indices (stride 64) on distinct banks.
__host__ __device__ inline int fft_skew2(int i) { return i + (i >> 3) + (i >> 6); }
__host__ __device__ inline int fft_skew2_len(int N) { return fft_skew2(N - 1) + 1; }

template <int N>
struct WarpFftShape {
    static constexpr int L = N == 512 ? 9 : N == 256 ? 8 : N == 128 ? 7 : N == 64 ? 6 : N == 32 ? 5 : N == 16 ? 4 : 3;
    static constexpr int N8 = L / 3, REM = L % 3;
    static constexpr int RL = REM == 2 ? 4 : (REM == 1 ? 2 : 1);  // radix of the last pass (1: none)
    static constexpr int NS_LAST = N / RL;                        // its stride = 8^N8
};

template <int N>
__host__ __device__ inline int fft_digit_rev(int n) {
    using S = WarpFftShape<N>;
    int pos = 0, span = N;
    if (S::RL > 1) { span /= S::RL; pos += (n % S::RL) * span; n /= S::RL; }
#pragma unroll
    for (int p = 0; p < S::N8; ++p) { span >>= 3; pos += (n & 7) * span; n >>= 3; }
    return pos;
}

template <int DIR, int R, int N, int NS, bool DIF>
__device__ __forceinline__ void warp_fft_inplace_pass(double2* __restrict__ buf, const double2* __restrict__ tw, int lane) {
    constexpr int nb = N / R;
    constexpr int tstep = N / (NS * R);
#pragma unroll
    for (int i = 0; i < (nb + 31) / 32; ++i) {
        const int b = lane + 32 * i;
        if ((nb % 32 == 0) || b < nb) {
            const int k = b % NS;
            const int base = (b / NS) * NS * R + k;
            double2 v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = buf[fft_skew2(base + r * NS)];
            double2 w[R];
            if (NS > 1) twiddle_powers<R>(tw[k * tstep], w);
            if (!DIF && NS > 1) {
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = tw_mul<DIR>(v[r], w[r]);
            }
            radix_butterfly<DIR, R>(v);
            if (DIF && NS > 1) {
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = tw_mul<DIR>(v[r], w[r]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) buf[fft_skew2(base + r * NS)] = v[r];
        }
    }
    __syncwarp();
}

template <int DIR, int N>
__device__ __forceinline__ void warp_fft_dit(double2* buf, const double2* tw, int lane) {
    using S = WarpFftShape<N>;
    if (S::N8 >= 1) warp_fft_inplace_pass<DIR, 8, N, 1, false>(buf, tw, lane);
    if (S::N8 >= 2) warp_fft_inplace_pass<DIR, 8, N, 8, false>(buf, tw, lane);
    if (S::N8 >= 3) warp_fft_inplace_pass<DIR, 8, N, 64, false>(buf, tw, lane);
    if (S::RL == 4) warp_fft_inplace_pass<DIR, 4, N, S::NS_LAST, false>(buf, tw, lane);
    if (S::RL == 2) warp_fft_inplace_pass<DIR, 2, N, S::NS_LAST, false>(buf, tw, lane);
}
template <int DIR, int N>
__device__ __forceinline__ void warp_fft_dif(double2* buf, const double2* tw, int lane) {
    using S = WarpFftShape<N>;
    if (S::RL == 4) warp_fft_inplace_pass<DIR, 4, N, S::NS_LAST, true>(buf, tw, lane);
    if (S::RL == 2) warp_fft_inplace_pass<DIR, 2, N, S::NS_LAST, true>(buf, tw, lane);
    if (S::N8 >= 3) warp_fft_inplace_pass<DIR, 8, N, 64, true>(buf, tw, lane);
    if (S::N8 >= 2) warp_fft_inplace_pass<DIR, 8, N, 8, true>(buf, tw, lane);
    if (S::N8 >= 1) warp_fft_inplace_pass<DIR, 8, N, 1, true>(buf, tw, lane);
}

inline bool warp_fft_supported(const FftPlanDev& pl) {
    for (int p = 0; p < pl.npass; ++p)
        if (pl.N / pl.radix[p] > 32 * (16 / pl.radix[p])) return false;
    return pl.npass > 0;
}
inline bool warp_fft_pow2_supported(int N) { return N >= 8 && N <= 512 && (N & (N - 1)) == 0; }

}  // namespace cfgpu
