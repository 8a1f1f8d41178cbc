// Batched Kleiser-Schumann tau / influence-matrix solver.
//
// Replaces, for all retained (kx,kz) Fourier modes at once, the reference's per-mode heap objects
//   TauSolver (tausolver.cpp:81-176 setup, :178-251 P/v solve + influence + tau correction, :347-450 solve),
//   HelmholtzSolver (helmholtz.cpp:18-95, :158-213 mean-constrained), BandedTridiag (bandedtridiag.cpp:212-277),
// the gather/scatter + mode loop of NSE::solve (nse.cpp:479-575), NSE::reset_lambda (nse.cpp:673-705) and the
// multistep right-hand-side accumulation (dnsalgo.cpp:217-224, FlowField::add flowfield.h:606-615), which is fused
// into the solve kernel as a linear combination of history fields.
//
// Storage (HBM): tile-major.  The retained modes q = (mxi - mx0)*(Kz+1) + kz of this rank are grouped, in that natural
// order, into tiles of TM consecutive modes (the unit of work of one CTA of the solve kernel): tile t holds
// q = t*TM .. t*TM+TM-1.  The (0,0) mode (q = 0 on the rank that owns kx row 0) needs the mean-flow treatment and gets a
// tile of its own at the END (index ngen = ceil(nq/TM)); its slot in tile 0 is masked.  Per tile: 12 arrays [m][n] (UL
// factors of the pressure and velocity Helmholtz operators, the six precomputed profiles P+-, v+-, P0, v0) followed by
// TSC_COUNT scalars [TM] -- one contiguous block, so a CTA streams its factors with fully coalesced loads.  A
// mode-independent table of the C&H 5.1.24 "B" rows follows.  The same tiling is the tile-major FIELD layout
// (TileGeom, cfgpu_internal.h): [tile][component][n][TM] complex, in which a CTA's history data is contiguous.
#pragma once
#include "cf_common.cuh"

namespace cfgpu {

enum { TSC_LAMP = 0, TSC_LAMV, TSC_KXX, TSC_KZZ, TSC_I00, TSC_I01, TSC_I10, TSC_I11, TSC_S0NB1, TSC_S0NB, TSC_COUNT };
// which: 0 upP 1 invP 2 bandP 3 upV 4 invV 5 bandV 6 Pp 7 vp 8 Pm 9 vm 10 P0 11 v0
enum { TAR_UPP = 0, TAR_INVP, TAR_BANDP, TAR_UPV, TAR_INVV, TAR_BANDV, TAR_PP, TAR_VP, TAR_PM, TAR_VM, TAR_P0, TAR_V0, TAR_COUNT };

struct TauData {
    int N;       // number of Chebyshev modes in the solve (Nyd)
    int nq;      // retained modes of this rank, q = (mxi - mx0)*(Kz+1) + kz ; with has00, q = 0 is the (0,0) mode
    int has00;   // this rank owns kx row 0 (always true on a single GPU)
    int TM;      // modes per tile
    int ntiles;  // ceil(nq/TM) + has00
    double nu, a, b;
    double* base;  // single allocation
    __host__ __device__ size_t tile_doubles() const { return (size_t)(TAR_COUNT * N + TSC_COUNT) * TM; }
    __host__ __device__ double* tile(int tl) const { return base + (size_t)tl * tile_doubles(); }
    __host__ __device__ double* tile_arr(int tl, int which) const { return tile(tl) + (size_t)which * N * TM; }   // [m][n]
    __host__ __device__ double* tile_sc(int tl, int which) const { return tile(tl) + (size_t)TAR_COUNT * N * TM + (size_t)which * TM; }
    __host__ __device__ double* btab() const { return base + (size_t)ntiles * tile_doubles(); }  // [3][N]: B_lo, B_dg, B_up
    __host__ __device__ int ngen() const { return ntiles - has00; }   // general tiles; tile ngen() is the (0,0) mode's
    __host__ __device__ int tile_of(int q) const { return (has00 && q == 0) ? ngen() : q / TM; }
    __host__ __device__ int pos_of(int q) const { return (has00 && q == 0) ? 0 : q % TM; }
    __host__ __device__ double& scq(int which, int q) const { return tile_sc(tile_of(q), which)[pos_of(q)]; }
    static int num_tiles(int nq, int TM, int has00) { return (nq + TM - 1) / TM + has00; }
    static size_t doubles(int N, int nq, int TM, int has00) {
        return (size_t)num_tiles(nq, TM, has00) * (TAR_COUNT * N + TSC_COUNT) * TM + 3 * (size_t)N;
    }
};

// pitch (doubles) of a skewed shared-memory column of the solve kernel: addr(n) = n + n/E, made odd so that the
// columns of a tile start in different banks
__host__ __device__ inline int tau_col_pitch(int N, int E) { return ((N - 1) + (N - 1) / E + 1) | 1; }

struct ModeGeom {
    int Nx, Ny, Nz, Kx, Kz;  // field grid and retained box
    int mx0;                 // first kx row (mxi) owned by this rank
    double Lx, Lz;
};

constexpr int TAU_MAXTERMS = 10;

struct TauSolveParams {
    TauData td;
    ModeGeom g;
    int TM_lin;        // modes per CTA of the linear-term kernel
    int taucorr;
    int nterms;
    const double* term[TAU_MAXTERMS];  // 3-component fields, reference layout
    double coef[TAU_MAXTERMS];
    double* uout;      // 3 components
    double* qout;      // 1 component
    // (0,0)-mode extras (nse.cpp:512-548)
    const double* Ubaseyy;  // nu is applied inside; may be null
    const double* Wbaseyy;
    int constraint;    // 0 pressure gradient, 1 bulk velocity
    double dPdxRef, dPdzRef, umean_target, wmean_target;
    double* dPd_act;   // device [2]: dPdxAct, dPdzAct written by the bulk-velocity solve
    // NSE::linear, bulk-velocity branch (nse.cpp:456-472): nu*(Ubase'(b)-Ubase'(a))/Ly and same for W
    double lin_base_dPdx, lin_base_dPdz;
    // 1: the history fields and the outputs are tile-major, [tile][component][n][TM] complex (TileGeom, cfgpu_internal.h;
    // qout has one component); 0: the reference's serial layout
    int tile_layout;
    int prefetch_terms;   // tile layout: L2-prefetch the CTA's whole history block at kernel start
};

int tau_setup_launch(const TauData& td, const ModeGeom& g, double lambda_t, cudaStream_t stream);
void tau_btab_host(int N, double* tab /* [3*N] */);
int tau_solve_launch(const TauSolveParams& p, cudaStream_t stream);
// batched real Helmholtz solves with one operator (HelmholtzSolver::solve, helmholtz.cpp:79-95); all pointers device
int helmholtz_batch_launch(int N, double a, double b, double lambda, double nu, int ncols, const double* f, const double* ua,
                           const double* ub, double* u, cudaStream_t stream);
// BandedTridiag::ULdecomp / ULsolveStrided / multiplyStrided (bandedtridiag.cpp:212-333): op 0 / 1 / 2; device pointers
int tridiag_launch(int op, int M, double* a, double* invdiag, double* x, double* y, int offset, int stride, cudaStream_t stream);
// PoissonSolver::solve (poissonsolver.cpp:146-202): lapl u = f mode by mode, Dirichlet data zero (bc == nullptr) or the wall
// values of bc; serial-layout spectral fields [Nd][N][Nx][Nz/2+1] complex, device pointers
int poisson_launch(int Nx, int N, int Nz, int Nd, double Lx, double Lz, double a, double b, const double* f, const double* bc, double* u,
                   cudaStream_t stream);
int tau_pick_TM(int N, int narrays_bytes_per_mode_row);
int tau_pick_TM_solve(int N);
int tau_pick_E(int N);
int tau_set_min_E(int e);  // returns the previous setting (0 = most parallel)

}  // namespace cfgpu
