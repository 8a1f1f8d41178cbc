// Chebyshev y-transform (DCT-I of length Ny) as a dense FP64 contraction on the DMMA tensor pipe.
//
// Replaces FlowField::makePhysical_y / makeSpectral_y (reference flowfield.cpp:1888-1987: gather a strided
// column, FFTW REDFT00, scatter) and the y-derivative recurrence used by curl (diffops.cpp:2266-2277,
// chebyshev.cpp:672-697), which is folded into a second matrix (value and d/dy come out of one pass).
//
// The cosine matrix has the reflection symmetry C[Ny-1-j][n] = (-1)^n C[j][n], so each transform is split into
// an even-n and an odd-n contraction of half the size (2x fewer flops):
//   inverse : E = Ce * X_even, O = Co * X_odd ;  u[j] = E + O ,  u[Ny-1-j] = s (E - O)   (s=+1 value, -1 d/dy)
//   forward : S = x[j] + x[Ny-1-j], D = x[j] - x[Ny-1-j] ;  c[2r] = Fe * S ,  c[2r+1] = Fo * D
// Normalisations of the reference (1/(Ny-1), the 1/2 end factors) are folded into the matrices.
#pragma once
#include "cf_common.cuh"

namespace cfgpu {

struct FftPlanDev;
constexpr int YG_MAXJOB = 9;
constexpr int YG_MAXMAT = 2;

struct YGemmJob {
    const double* in;
    const double* in2;   // forward mode only, may be null: second input whose y-DERIVATIVE coefficients are added to the
                         // output (matrices A1b/A2b act on the swapped difference / sum tiles)
    double* out[YG_MAXMAT];
    double* const* out_rows[YG_MAXMAT];  // optional device table [N]: address of output row r (rows may live in another GPU's
                                         // memory: the slab all-to-all fused into the epilogue); null: out + r*out_ld
    int nmat;        // how many of the plan's matrices to apply to this input (inverse: 1 = value, 2 = value + d/dy)
    int mat0;        // first matrix index (inverse: 0 = value, 1 = derivative)
    double add00[YG_MAXMAT];  // unused (reserved)
};

struct YGemmParams {
    int N;      // Ny
    int mode;   // 0 = inverse (spectral -> physical), 1 = forward
    int M;      // rows of A1/A2 actually used (inverse: Nh ; forward: Ne)
    int M2;     // rows of A2 used (inverse: Nh ; forward: No)
    int K1, K2;        // true inner dims
    int K1p, K2p;      // inner dims padded to multiples of 4 (lead dims of A1/A2; B tiles zero padded)
    const double* A1[YG_MAXMAT];  // [Mp x K1p] row major, Mp = M rounded up to 8, zero padded
    const double* A2[YG_MAXMAT];  // [Mp x K2p]
    double sgn[YG_MAXMAT];        // inverse: sign of the reflected row
    double in2_scale;             // factor of the in2 term (what A1b/A2b carry: 1, or 1/2 for the skew-symmetric average)
    const double* A1b;            // [Mp x K1p]: even output rows from the difference tile of in2
    const double* A2b;            // [Mp x K1p]: odd output rows from the sum tile of in2
    long ncols;                   // number of (double) columns
    int in_runlen;                // column c lives at in_runstart[c / in_runlen] + c % in_runlen  (nullptr: identity)
    const long* in_runstart;
    long in_ld;                   // row stride (doubles)
    int out_runlen;
    const long* out_runstart;
    long out_ld;
    int two_inputs;               // some job has in2 (set by the launcher: doubles the operand tiles in shared memory)
    const FftPlanDev* fft_half;   // host pointer, may be null: plan of length N-1 for the half-length variant of the same
    const FftPlanDev* fft;        // host pointer, may be null: FFT plan of length 2(N-1) -- the jobs then run as shared-memory
    double ya, yb;                // FFTs (yfft.cu) where that kernel covers them; wall positions for its d/dy scale
    int njobs;
    YGemmJob job[YG_MAXJOB];
};

// launches on `stream`; returns 0 on success
int ygemm_launch(const YGemmParams& p, cudaStream_t stream);
// yfft.cu: the same jobs as FFTs; -1 = not covered (second inputs, unsupported length), the caller falls back to the contraction
bool yfft_length_supported(int N);
int yfft_launch(const YGemmParams& p, const FftPlanDev& pl, double a, double b, cudaStream_t stream);

}  // namespace cfgpu
