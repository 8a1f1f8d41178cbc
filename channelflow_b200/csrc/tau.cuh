// Batched Kleiser-Schumann tau / influence-matrix solver.
//
// Replaces, for all retained (kx,kz) Fourier modes at once, the reference's per-mode heap objects
//   TauSolver (tausolver.cpp:81-176 setup, :178-251 P/v solve + influence + tau correction, :347-450 solve),
//   HelmholtzSolver (helmholtz.cpp:18-95, :158-213 mean-constrained), BandedTridiag (bandedtridiag.cpp:212-277),
// the gather/scatter + mode loop of NSE::solve (nse.cpp:479-575), NSE::reset_lambda (nse.cpp:673-705) and the
// multistep right-hand-side accumulation (dnsalgo.cpp:217-224, FlowField::add flowfield.h:606-615), which is fused
// into the solve kernel as a linear combination of history fields.
//
// Storage (HBM): every per-mode array is [row n][mode q] with q fastest ("SoA"), so a warp touching one row reads
// consecutive doubles.  Per mode and substep: 12*N + NSC doubles (UL factors of the pressure and velocity
// Helmholtz operators, the 6 precomputed profiles P+-, v+-, P0, v0, and scalars).
#pragma once
#include "cf_common.cuh"

namespace cfgpu {

enum { TSC_LAMP = 0, TSC_LAMV, TSC_KXX, TSC_KZZ, TSC_I00, TSC_I01, TSC_I10, TSC_I11, TSC_S0NB1, TSC_S0NB, TSC_COUNT };

struct TauData {
    int N;       // number of Chebyshev modes in the solve (Nyd)
    int nq;      // retained modes, q = mxi*(Kz+1) + kz ; q = 0 is the (0,0) mode
    int ldq;     // nq rounded up to 32
    double nu, a, b;
    double* base;  // single allocation
    __host__ __device__ double* arr(int which) const { return base + (size_t)which * N * ldq; }
    // which: 0 upP 1 invP 2 bandP 3 upV 4 invV 5 bandV 6 Pp 7 vp 8 Pm 9 vm 10 P0 11 v0
    __host__ __device__ double* sc(int which) const { return base + (size_t)12 * N * ldq + (size_t)which * ldq; }
    static size_t doubles(int N, int ldq) { return (size_t)(12 * N + TSC_COUNT) * ldq; }
};

struct ModeGeom {
    int Nx, Ny, Nz, Kx, Kz;  // field grid and retained box
    double Lx, Lz;
};

constexpr int TAU_MAXTERMS = 10;

struct TauSolveParams {
    TauData td;
    ModeGeom g;
    int TM;            // modes per CTA
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
};

int tau_setup_launch(const TauData& td, const ModeGeom& g, double lambda_t, int TM, cudaStream_t stream);
int tau_solve_launch(const TauSolveParams& p, cudaStream_t stream);
int tau_pick_TM(int N, int narrays_bytes_per_mode_row);

}  // namespace cfgpu
