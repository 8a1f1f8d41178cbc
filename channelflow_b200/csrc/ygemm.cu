// DMMA (FP64 tensor-core) kernel for the Chebyshev y-transform.  See ygemm.cuh for the maths.
//
// Tiling: one CTA owns BN (64/32/16) columns (= consecutive doubles: re/im of consecutive kz modes) of one field
// component and ALL Ny rows: it stages the even/odd (or sum/difference) operand tiles in shared memory once
// (each HBM element is read exactly once, coalesced 512-byte rows), then its 8 warps sweep the output rows in
// 32-row tiles issuing mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) with A fragments served from L1/L2 (the matrices
// are a few hundred KB and shared by every CTA) and B fragments from shared memory (row pitch BN+4 doubles =>
// conflict-free fragment loads).  Each output element is written exactly once.
// Roofline: tensor (FP64) for Ny >~ 49, HBM below; algorithmic flops = 2*Ny*ceil(Ny/2)*2 per column per output.
#include "ygemm.cuh"

namespace cfgpu {

template <int BN>
__global__ void __launch_bounds__(256) ygemm_kernel(const YGemmParams p) {
    constexpr int LD = BN + 4;
    constexpr int NWN = BN / 16;  // warps along the column direction (16 columns each)
    constexpr int NWM = 8 / NWN;  // warps along the row direction
    double* B1 = dyn_smem<double>();
    double* B2 = B1 + (size_t)p.K1p * LD;
    long* cin = reinterpret_cast<long*>(B2 + (size_t)p.K2p * LD);
    long* cout = cin + BN;

    const YGemmJob& job = p.job[blockIdx.y];
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
        // rows 0..K1p-1 of B1 hold even n = 2r, rows of B2 hold odd n = 2r+1; zero padded
        const int total = (p.K1p + p.K2p) * BN;
        for (int idx = tid; idx < total; idx += 256) {
            const int r = idx / BN, c = idx % BN;
            const long off = cin[c];
            double v = 0.0;
            if (r < p.K1p) {
                if (r < p.K1 && off >= 0) v = in[(long)(2 * r) * p.in_ld + off];
                B1[r * LD + c] = v;
            } else {
                const int r2 = r - p.K1p;
                if (r2 < p.K2 && off >= 0) v = in[(long)(2 * r2 + 1) * p.in_ld + off];
                B2[r2 * LD + c] = v;
            }
        }
    } else {
        // forward: B1[j] = x[j] + x[Nb-j], B2[j] = x[j] - x[Nb-j]  (self-paired middle row: B1 = x, B2 = 0)
        const int total = p.K1p * BN;
        for (int idx = tid; idx < total; idx += 256) {
            const int j = idx / BN, c = idx % BN;
            const long off = cin[c];
            double s = 0.0, d = 0.0;
            if (j < p.K1 && off >= 0) {
                const int jj = Nb - j;
                const double a = in[(long)j * p.in_ld + off];
                if (jj != j) {
                    const double b = in[(long)jj * p.in_ld + off];
                    s = a + b;
                    d = a - b;
                } else {
                    s = a;
                }
            }
            B1[j * LD + c] = s;
            if (j < p.K2p) B2[j * LD + c] = d;
        }
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int wn = warp % NWN, wm = warp / NWN;
    const int lr = lane >> 2, lk = lane & 3;  // fragment row / k (A), n / k (B)
    const int Mmax = p.M > p.M2 ? p.M : p.M2;
    const int Mp = (Mmax + 7) & ~7;
    const int Mtiles = (Mp + 31) / 32;
    const int ncb = wn * 16;

    for (int mi = 0; mi < job.nmat; ++mi) {
        const int mat = job.mat0 + mi;
        const double* __restrict__ A1 = p.A1[mat];
        const double* __restrict__ A2 = p.A2[mat];
        double* __restrict__ out = job.out[mi];
        const double sgn = p.sgn[mat];
        for (int mt = wm; mt < Mtiles; mt += NWM) {
            const int row0 = mt * 32;
            double e[4][2][2], o[4][2][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int t = 0; t < 2; ++t) e[i][t][0] = e[i][t][1] = o[i][t][0] = o[i][t][1] = 0.0;

            for (int k = 0; k < p.K1p; k += 4) {
                double b[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) b[t] = B1[(k + lk) * LD + ncb + t * 8 + lr];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (row0 + i * 8 < Mp) {
                        const double a = __ldg(&A1[(size_t)(row0 + i * 8 + lr) * p.K1p + k + lk]);
#pragma unroll
                        for (int t = 0; t < 2; ++t) dmma_m8n8k4(e[i][t][0], e[i][t][1], a, b[t]);
                    }
                }
            }
            for (int k = 0; k < p.K2p; k += 4) {
                double b[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) b[t] = B2[(k + lk) * LD + ncb + t * 8 + lr];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (row0 + i * 8 < Mp) {
                        const double a = __ldg(&A2[(size_t)(row0 + i * 8 + lr) * p.K2p + k + lk]);
#pragma unroll
                        for (int t = 0; t < 2; ++t) dmma_m8n8k4(o[i][t][0], o[i][t][1], a, b[t]);
                    }
                }
            }

            // epilogue: c fragment = C[lr][2*lk + {0,1}]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = row0 + i * 8 + lr;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int cc = ncb + t * 8 + 2 * lk;
                    const long off = cout[cc];
                    if (off < 0) continue;
                    if (p.mode == 0) {
                        if (r < p.M) {
                            const double E0 = e[i][t][0], E1 = e[i][t][1], O0 = o[i][t][0], O1 = o[i][t][1];
                            *reinterpret_cast<double2*>(&out[(long)r * p.out_ld + off]) = make_double2(E0 + O0, E1 + O1);
                            const int rr = Nb - r;
                            if (rr != r)
                                *reinterpret_cast<double2*>(&out[(long)rr * p.out_ld + off]) =
                                    make_double2(sgn * (E0 - O0), sgn * (E1 - O1));
                        }
                    } else {
                        if (r < p.M)
                            *reinterpret_cast<double2*>(&out[(long)(2 * r) * p.out_ld + off]) =
                                make_double2(e[i][t][0], e[i][t][1]);
                        if (r < p.M2)
                            *reinterpret_cast<double2*>(&out[(long)(2 * r + 1) * p.out_ld + off]) =
                                make_double2(o[i][t][0], o[i][t][1]);
                    }
                }
            }
        }
    }
}

template <int BN>
static int launch_bn(const YGemmParams& p, cudaStream_t stream) {
    const size_t smem = (size_t)(p.K1p + p.K2p) * (BN + 4) * sizeof(double) + 2 * BN * sizeof(long);
    auto kfn = ygemm_kernel<BN>;
    static size_t configured = 0;
    if (smem > configured) {
        CF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)((p.ncols + BN - 1) / BN), (unsigned)p.njobs);
    CF_LAUNCH(kfn, grid, dim3(256), smem, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}

int ygemm_launch(const YGemmParams& p, cudaStream_t stream) {
    if (p.ncols <= 0 || p.njobs <= 0) return 0;
    const size_t rows = (size_t)(p.K1p + p.K2p);
    const size_t limit = 220 * 1024;
    if (rows * 68 * 8 + 1024 <= limit) return launch_bn<64>(p, stream);
    if (rows * 36 * 8 + 512 <= limit) return launch_bn<32>(p, stream);
    if (rows * 20 * 8 + 256 <= limit) return launch_bn<16>(p, stream);
    set_last_error("ygemm: Ny too large for the shared-memory tile");
    return 1;
}

}  // namespace cfgpu
