// DMMA (FP64 tensor-core) kernel for the Chebyshev y-transform.  See ygemm.cuh for the maths.
//
// Tiling: one CTA owns BN (64/32/16) columns (= consecutive doubles: re/im of consecutive kz modes) of one field
// component and ALL Ny rows: it stages the even/odd (or sum/difference) operand tiles in shared memory once
// (each HBM element is read exactly once, 512-byte runs; inverse mode uses cp.async so that the whole tile is in
// flight at once), then its 8 warps -- 4 along the output rows, 2 along the columns -- each own a 32-row x (BN/2)-column
// output tile of BOTH parity products (E and O accumulators in registers) and issue mma.sync.m16n8k8.f64 (ptxas lowers
// it to four DMMA.8x8x4 on sm_100a -- cuobjdump -sass, profiles/r02_sass_opcodes.txt -- the wide PTX shape only saves
// fragment bookkeeping): per k-step of 8 a warp loads 8 A and BN/8 B fragment registers for BN/8 MMAs.  A fragments come
// from L1/L2 (the matrices are a few hundred KB, shared by every CTA) and are register double-buffered one k-step
// ahead; B fragments come from shared memory (row pitch BN+4 doubles => conflict-free fragment loads).
// A trailing remainder of at most 2 rows (Ny = 2^k+1 gives 32*m + 1 rows: the self-paired middle point / last
// coefficient) is not worth a 32-row tile and is evaluated as plain dot products instead.
// Each output element is written exactly once.
// Roofline: tensor (FP64) for Ny >~ 49, HBM below; algorithmic flops = 2*Ny*ceil(Ny/2)*2 per column per output.
#include "fft_smem.cuh"
#include "ygemm.cuh"

#include <cstdlib>

namespace cfgpu {

constexpr bool YG_SPLIT_DEFAULT = true;   // A/B at 512x257x512 (profiles/r02c_*): inverse 1.71 -> 1.65 ms, forward 1.18 -> 1.14 ms

// BN columns per CTA, WN warps along the columns (4 warps along the rows): <64,2> is one 256-thread CTA per SM, <32,1> two
// independent 128-thread CTAs per SM with the same 32 x 32 warp tile -- one CTA's staging and epilogue overlap the other's
// DMMA loop.
template <int BN, int WN, int MINB>
__global__ void __launch_bounds__(128 * WN, MINB) ygemm_kernel(const YGemmParams p) {
    constexpr int YG_THREADS = 128 * WN;
    constexpr int LD = BN + 4;
    constexpr int NT = BN / (8 * WN);   // 8-column n-tiles per warp
    const YGemmJob& job = p.job[blockIdx.y];
    const bool two = p.two_inputs != 0;  // launch-wide: second pair of tiles present in shared memory
    double* B1 = dyn_smem<double>();
    double* B2 = B1 + (size_t)p.K1p * LD;
    double* B1b = B2 + (size_t)p.K2p * LD;
    double* B2b = B1b + (two ? (size_t)p.K1p * LD : 0);
    long* cin = reinterpret_cast<long*>(B2b + (two ? (size_t)p.K2p * LD : 0));
    long* cout = cin + BN;

    const int tid = threadIdx.x;
    const long c0 = (long)blockIdx.x * BN;

    if (tid < BN) {
        long c = c0 + tid;
        long oi = -1, oo = -1;
        if (c < p.ncols) {
            oi = p.in_runstart ? p.in_runstart[c / p.in_runlen] + (c % p.in_runlen) : c;
            oo = p.out_runstart ? p.out_runstart[c / p.out_runlen] + (c % p.out_runlen) : c;
        }
        cin[tid] = oi;
        cout[tid] = oo;
    }
    __syncthreads();

    const int N = p.N, Nb = N - 1;
    const double* __restrict__ in = job.in;
    if (p.mode == 0) {
        // rows 0..K1p-1 of B1 hold even n = 2r, rows of B2 hold odd n = 2r+1; zero padded.  Columns come in (re,im)
        // pairs (all run lengths are even), so 16-byte cp.async copies are aligned on both sides.
        const int HB = BN / 2;
        const int total = (p.K1p + p.K2p) * HB;
        for (int idx = tid; idx < total; idx += YG_THREADS) {
            const int r = idx / HB, c = 2 * (idx - r * HB);
            const long off = cin[c];
            const bool odd = r >= p.K1p;
            const int rr = odd ? r - p.K1p : r;
            double* dst = (odd ? B2 : B1) + rr * LD + c;
            const int n = odd ? 2 * rr + 1 : 2 * rr;
            if (rr < (odd ? p.K2 : p.K1) && off >= 0) cp_async16(dst, in + (long)n * p.in_ld + off);
            else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
        }
        cp_async_wait_all();
    } else {
        // forward: B1[j] = x[j] + x[Nb-j], B2[j] = x[j] - x[Nb-j]  (self-paired middle row: B1 = x, B2 = 0)
        const int HB = BN / 2;
        const int total = p.K1p * HB;
        for (int which = 0; which < ((two && job.in2) ? 2 : 1); ++which) {
        const double* __restrict__ in = which ? job.in2 : job.in;
        double* B1 = which ? B1b : (dyn_smem<double>());
        double* B2 = which ? B2b : (dyn_smem<double>() + (size_t)p.K1p * LD);
        for (int i0 = tid; i0 < total; i0 += 4 * YG_THREADS) {
            double2 a[4], b[4];
            int jj[4], cc[4];
            bool ok[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int idx = i0 + h * YG_THREADS;
                const int j = idx / HB, c = 2 * (idx - j * HB);
                jj[h] = j; cc[h] = c;
                const long off = idx < total ? cin[c] : -1;
                ok[h] = idx < total && j < p.K1 && off >= 0;
                a[h] = b[h] = make_double2(0.0, 0.0);
                if (ok[h]) {
                    a[h] = *reinterpret_cast<const double2*>(in + (long)j * p.in_ld + off);
                    if (Nb - j != j) b[h] = *reinterpret_cast<const double2*>(in + (long)(Nb - j) * p.in_ld + off);
                }
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (i0 + h * YG_THREADS >= total) break;
                const int j = jj[h], c = cc[h];
                double2 s = make_double2(0.0, 0.0), d = s;
                if (ok[h]) {
                    if (Nb - j != j) {
                        s = make_double2(a[h].x + b[h].x, a[h].y + b[h].y);
                        d = make_double2(a[h].x - b[h].x, a[h].y - b[h].y);
                    } else {
                        s = a[h];
                    }
                }
                *reinterpret_cast<double2*>(&B1[j * LD + c]) = s;
                if (j < p.K2p) *reinterpret_cast<double2*>(&B2[j * LD + c]) = d;
            }
        }
        }
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int wn = warp % WN, wm = warp / WN;  // WN warps along the columns, 4 along the rows
    const int lr = lane >> 2, lk = lane & 3;  // fragment row / k (A), n / k (B)
    const int Mmax = p.M > p.M2 ? p.M : p.M2;
    const int rem = Mmax & 31;
    const int Mgemm = (rem >= 1 && rem <= 2) ? Mmax - rem : Mmax;  // rows done on the tensor pipe
    const int Mtiles = (Mgemm + 31) / 32;
    const int ncb = wn * (BN / WN);

    for (int mi = 0; mi < job.nmat; ++mi) {
        const int mat = job.mat0 + mi;
        const double* __restrict__ A1 = p.A1[mat];
        const double* __restrict__ A2 = p.A2[mat];
        double* __restrict__ out = job.out[mi];
        double* const* __restrict__ rows = job.out_rows[mi];
        auto row_ptr = [&](int r) -> double* { return rows ? rows[r] : out + (long)r * p.out_ld; };
        const double sgn = p.sgn[mat];
        for (int mt = wm; mt < Mtiles; mt += 4) {
            const int row0 = mt * 32;
            double e[2][NT][4], o[2][NT][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int q = 0; q < 4; ++q) e[i][t][q] = o[i][t][q] = 0.0;
            // acc += A[row0 .. row0+32][0 .. Kp) * Bt[0 .. Kp)[this warp's columns]; A fragments of one k-step:
            // [row block i][a0..a3] = rows lr / lr+8 of block i, columns lk / lk+4, prefetched one k-step ahead
            auto accumulate = [&](double (&acc)[2][NT][4], const double* __restrict__ A, const int Kp, const double* __restrict__ Bt) {
                const double* __restrict__ ap = A + (size_t)(row0 + lr) * Kp + lk;
                const double* __restrict__ bp = Bt + lk * LD + ncb + lr;
                const int nk = Kp / 8;
                double a[2][4], an[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) a[i][q] = __ldg(ap + (size_t)(16 * i + 8 * (q & 1)) * Kp + 4 * (q >> 1));
                for (int ks = 0; ks < nk; ++ks) {
                    if (ks + 1 < nk) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                an[i][q] = __ldg(ap + (size_t)(16 * i + 8 * (q & 1)) * Kp + 4 * (q >> 1) + 8 * (ks + 1));
                    }
                    double b[NT][2];
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        b[t][0] = bp[(8 * ks) * LD + t * 8];
                        b[t][1] = bp[(8 * ks + 4) * LD + t * 8];
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int t = 0; t < NT; ++t) dmma_m16n8k8(acc[i][t], a[i], b[t]);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int q = 0; q < 4; ++q) a[i][q] = an[i][q];
                }
            };
            accumulate(e, A1, p.K1p, B1);
            accumulate(o, A2, p.K2p, B2);
            if (two && job.in2 && p.mode == 1) {
                // derivative matrices of the second input: its difference tile feeds the even rows, its sum tile the odd rows
                accumulate(e, p.A1b, p.K1p, B2b);
                accumulate(o, p.A2b, p.K1p, B1b);
            }

            // epilogue: c0,c1 = C[lr][2*lk + {0,1}], c2,c3 = C[lr+8][..] of each 16-row block
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int r = row0 + i * 16 + hh * 8 + lr;
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const int cc = ncb + t * 8 + 2 * lk;
                        const long off = cout[cc];
                        if (off < 0 || r >= Mgemm) continue;
                        const double E0 = e[i][t][2 * hh], E1 = e[i][t][2 * hh + 1], O0 = o[i][t][2 * hh], O1 = o[i][t][2 * hh + 1];
                        if (p.mode == 0) {
                            if (r < p.M) {
                                *reinterpret_cast<double2*>(row_ptr(r) + off) = make_double2(E0 + O0, E1 + O1);
                                const int rr = Nb - r;
                                if (rr != r)
                                    *reinterpret_cast<double2*>(row_ptr(rr) + off) = make_double2(sgn * (E0 - O0), sgn * (E1 - O1));
                            }
                        } else {
                            if (r < p.M) *reinterpret_cast<double2*>(row_ptr(2 * r) + off) = make_double2(E0, E1);
                            if (r < p.M2) *reinterpret_cast<double2*>(row_ptr(2 * r + 1) + off) = make_double2(O0, O1);
                        }
                    }
                }
            }
        }
        // remainder rows (at most 2) as dot products.  Every warp takes BN/8 columns and splits K over its lanes
        // (lane = slice * CW + column), partial sums combined with shuffles: no block barrier, all 8 warps busy, a quarter
        // (BN = 64) of the K loop per lane.  (One thread per column on two warps kept the other six -- and the tensor
        // pipe of the SM, one CTA being resident -- waiting for ~45 % of the CTA's life time: ncu r01f.)
        {
            constexpr int CW = BN / (YG_THREADS / 32);   // columns per warp
            constexpr int NSL = 32 / CW;                 // K slices per column
            static_assert(CW >= 1 && CW * NSL == 32, "BN must be a multiple of the warp count and divide 32 per warp");
            const int cc = lane % CW, sl = lane / CW;
            const int c = warp * CW + cc;
            const long off = cout[c];
            for (int r = Mgemm; r < Mmax; ++r) {
                double E = 0.0, O = 0.0;
                {   // both parity products in one loop: two independent chains, twice the loads in flight
                    const bool hasE = r < p.M, hasO = r < p.M2;
                    const double* __restrict__ a1 = A1 + (size_t)r * p.K1p;
                    const double* __restrict__ a2 = A2 + (size_t)r * p.K2p;
                    const int K1 = hasE ? p.K1 : 0, K2 = hasO ? p.K2 : 0;
                    const int Kc = K1 < K2 ? K1 : K2;
                    int k = sl;
#pragma unroll 8
                    for (; k < Kc; k += NSL) {
                        E += __ldg(a1 + k) * B1[k * LD + c];
                        O += __ldg(a2 + k) * B2[k * LD + c];
                    }
                    for (int k1 = k; k1 < K1; k1 += NSL) E += __ldg(a1 + k1) * B1[k1 * LD + c];
                    for (int k2 = k; k2 < K2; k2 += NSL) O += __ldg(a2 + k2) * B2[k2 * LD + c];
                }
                if (two && job.in2 && p.mode == 1) {
                    if (r < p.M) {
                        const double* __restrict__ ar = p.A1b + (size_t)r * p.K1p;
#pragma unroll 4
                        for (int k = sl; k < p.K1; k += NSL) E += __ldg(ar + k) * B2b[k * LD + c];
                    }
                    if (r < p.M2) {
                        const double* __restrict__ ar = p.A2b + (size_t)r * p.K1p;
#pragma unroll 4
                        for (int k = sl; k < p.K1; k += NSL) O += __ldg(ar + k) * B1b[k * LD + c];
                    }
                }
#pragma unroll
                for (int d = CW; d < 32; d <<= 1) {
                    E += __shfl_xor_sync(0xffffffffu, E, d);
                    O += __shfl_xor_sync(0xffffffffu, O, d);
                }
                if (sl != 0 || off < 0) continue;
                if (p.mode == 0) {
                    if (r < p.M) {
                        row_ptr(r)[off] = E + O;
                        const int rr = Nb - r;
                        if (rr != r) row_ptr(rr)[off] = sgn * (E - O);
                    }
                } else {
                    if (r < p.M) row_ptr(2 * r)[off] = E;
                    if (r < p.M2) row_ptr(2 * r + 1)[off] = O;
                }
            }
        }
    }
}

template <int BN, int WN = 2, int MINB = 1>
static int launch_bn(const YGemmParams& p, cudaStream_t stream) {
    constexpr int YG_THREADS = 128 * WN;
    const size_t smem = (size_t)(p.two_inputs ? 2 : 1) * (p.K1p + p.K2p) * (BN + 4) * sizeof(double) + 2 * BN * sizeof(long);
    auto kfn = ygemm_kernel<BN, WN, MINB>;
    static size_t configured = 0;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)((p.ncols + BN - 1) / BN), (unsigned)p.njobs);
    CF_LAUNCH(kfn, grid, dim3(YG_THREADS), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}

int ygemm_launch(const YGemmParams& p0, cudaStream_t stream) {
    YGemmParams p = p0;
    if (p.ncols <= 0 || p.njobs <= 0) return 0;
    if (p.fft) {
        const int rc = yfft_launch(p, *p.fft, p.ya, p.yb, stream);
        if (rc >= 0) return rc;
        // CF_YFFT_STRICT=1 (tests): a job the FFT kernels do not cover is an error instead of a contraction
        static const int strict = getenv("CF_YFFT_STRICT") ? atoi(getenv("CF_YFFT_STRICT")) : 0;
        if (strict) { set_last_error("ygemm: y-transform parameters not covered by the FFT kernels (CF_YFFT_STRICT)"); return 1; }
    }
    p.two_inputs = 0;
    for (int j = 0; j < p.njobs; ++j)
        if (p.job[j].in2) p.two_inputs = 1;
    if (p.two_inputs && (p.mode != 1 || !p.A1b || !p.A2b)) {
        set_last_error("ygemm: a second input needs the forward mode and the derivative matrices");
        return 1;
    }
    if ((p.in_runstart && (p.in_runlen & 1)) || (p.out_runstart && (p.out_runlen & 1)) || (p.in_ld & 1) || (p.out_ld & 1)) {
        set_last_error("ygemm: column runs must be (re,im) pairs");
        return 1;
    }
    const size_t rows = (size_t)(p.K1p + p.K2p) * (p.two_inputs ? 2 : 1);
    const size_t limit = 220 * 1024;
    // two half-width CTAs per SM when one full-width tile would own the SM alone (long profiles): CF_YG_SPLIT=0/1 overrides
    static const int split = getenv("CF_YG_SPLIT") ? atoi(getenv("CF_YG_SPLIT")) : -1;
    const bool one_per_sm = rows * 68 * 8 + 1024 > 110 * 1024;
    if (rows * 36 * 8 + 512 <= 110 * 1024 && (split == 1 || (split < 0 && one_per_sm && YG_SPLIT_DEFAULT))) return launch_bn<32, 1, 2>(p, stream);
    if (rows * 36 * 8 + 512 <= 110 * 1024 && split == 2) return launch_bn<32, 2, 2>(p, stream);  // 16 warps per SM, 32 x 16 warp tiles
    if (rows * 68 * 8 + 1024 <= limit) return launch_bn<64>(p, stream);
    if (rows * 36 * 8 + 512 <= limit) return launch_bn<32>(p, stream);
    if (rows * 20 * 8 + 256 <= limit) return launch_bn<16>(p, stream);
    set_last_error("ygemm: Ny too large for the shared-memory tile");
    return 1;
}

}  // namespace cfgpu
