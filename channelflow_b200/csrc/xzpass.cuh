// x-pass and z-pass kernels of the fused nonlinear-term pipeline (see DESIGN.md "Pipeline").
//
// Intermediate ("pencil") arrays are compact: only the retained (de-aliased) modes are stored.
//   P[f][ny][mxi][kz]   complex, mxi = 0..2Kx   (kx = 0..Kx, -Kx..-1), kz = 0..Kz   (x,z spectral, y physical)
//   Q[f][ny][nx ][kz]   complex, nx  = 0..Nx-1                                       (x physical, z spectral)
// Rotational form: the vorticity is formed in spectral space while the x-pass loads its operands
// (w_x = dw/dy - i kz v, w_y = i kz u - i kx w, w_z = i kx v - du/dy), so six fields (u, curl u) go through the inverse
// x and z passes -- three paired c2r transforms per line -- and three (f) come back.
// Replaces FlowField::makePhysical_xz / makeSpectral_xz (flowfield.cpp:1850-1886) for the DNS path, the x/z
// derivative factors of curl (diffops.cpp:2318-2332), cross (diffops.cpp:2588-2606), the Coriolis term and
// base-flow handling of navierstokesNL (nse.cpp:28-36,63-88), zeroPaddedModes (flowfield.cpp:2235-2255; aliased
// modes are simply never produced) and the CFL maximum (flowfield.cpp:4035-4068).
#pragma once
#include "cf_common.cuh"
#include "fft_smem.cuh"

namespace cfgpu {

struct XPassParams {
    int Nx, Ny, Kx, Kz;   // Kx,Kz: retained |kx|<=Kx, kz<=Kz ; nmx = 2Kx+1, nkz = Kz+1
    int TZ;               // kz columns per CTA
    int inplace;          // forward pass only: one-buffer in-place transform (long x-lines), see xpass_forward_inplace_kernel
    double Lx;
    FftPlanDev plan;      // length Nx
    int nfields;          // fields this launch handles: output slots fsel[0..nfields)
    int fsel[12];
    // inverse pass, indexed by output slot:  out = op_a(P[src]) - op_b(P[srcb])   (srcb < 0: no second term);
    // op: 0 identity, 1 d/dx = i 2 pi kx/Lx, 2 d/dz = i 2 pi kz/Lz  -- this is where curl u is formed (diffops.cpp:2318-2332)
    int src[12], opa[12];
    int srcb[12], opb[12];
    double Lz;
    // forward pass, indexed by output slot:  out = FFT(Q[src]) + cb * op_b(FFT(Q[srcb]))  (srcb < 0: one source)
    double cb;
    const double2* in;    // inverse: P (field stride Ny*nmx*nkz); forward: Q (field stride Ny*Nx*nkz)
    double2* out;         // inverse: Q ; forward: P
    int ny0, nyn;         // y range handled (physical slab)
    // P-side layout: blocks by owner rank s of the kx rows, [s][field][yl][mxi - xsplit[s]][kz] (the all-to-all staging
    // of the slab decomposition, comm.cuh); with one rank this is plain P[field][yl][mxi][kz]
    int nranks;
    int nstage;           // number of fields in the staging buffer
    int xsplit[17];       // xsplit[s] .. xsplit[s+1] = kx rows of rank s
    // forward pass with peer memory: rank s's pencil buffer P_s[f][Ny][mxi - xsplit[s]][kz] mapped into this process; the
    // x-pass stores each kx row straight into its owner's memory (the all-to-all fused into the store). null: staged.
    double2* peer_out[16];
    int peer_direct;      // 1: every row to its owner; 2: only the rows owned by self_rank (push mode), the others are staged
    int self_rank;
};
// offset (complex elements) of (field f, global plane y, mxi) inside owner rank s's kx-slab pencil buffer
__host__ __device__ inline size_t xpass_peer_offset(const XPassParams& p, int f, int y, int mxi, int nkz, int& s) {
    s = 0;
    while (s + 1 < p.nranks && mxi >= p.xsplit[s + 1]) ++s;
    const int x0 = p.xsplit[s], nloc = p.xsplit[s + 1] - x0;
    return (size_t)nkz * (((size_t)f * p.Ny + y) * nloc + (mxi - x0));
}
// offset (complex elements) of row (field f, plane yl, mxi) in the staged P-side layout
__host__ __device__ inline size_t xpass_row_offset(const XPassParams& p, int f, int yl, int mxi, int nkz) {
    int s = 0;
    while (s + 1 < p.nranks && mxi >= p.xsplit[s + 1]) ++s;
    const int x0 = p.xsplit[s], nloc = p.xsplit[s + 1] - x0;
    return (size_t)nkz * ((size_t)p.nstage * p.nyn * x0 + ((size_t)f * p.nyn + yl) * nloc + (mxi - x0));
}
int xpass_inverse_launch(const XPassParams& p, cudaStream_t stream);
int xpass_forward_launch(const XPassParams& p, cudaStream_t stream);

// ZP_CONVECTION / ZP_DIVERGENCE / ZP_SKEW (zpass_general_kernel, power-of-two Nz <= 512):
//   inputs  Q[0..2] = u,v,w; for convection and skew also Q[3..5] = d/dy, Q[6..8] = d/dx, Q[9..11] = d/dz of u,v,w
//   outputs convection: F[0..2] = u_j d_j u_i;
//           divergence / skew: F[0..2] = G_i = cc u_j d_j u_i + cd d/dz (u_i w), F[3..5] = u u, u v, u w, F[6..7] = v v, v w
//   (u = total velocity incl. base flow; the x- and y-derivative parts of d_j (u_i u_j) are added by the forward x-pass
//   and y-GEMM; diffops.cpp:3110-3134, 3142-3286, 3586-3643)
enum ZPassMode { ZP_ROTATIONAL = 0, ZP_CFL = 1, ZP_CONVECTION = 2, ZP_DIVERGENCE = 3, ZP_SKEW = 4 };

struct ZPassParams {
    int Nx, Ny, Nz, Kz;
    int TL;               // x-lines per CTA
    int mode;
    double Lx, Lz;
    double scale;         // 1/(Nx*Nz)
    double Vsuck, rotation;
    double cc, cd;        // weights of the convective and divergence parts (1,0 / 0,1 / 1/2,1/2)
    FftPlanDev plan;      // length Nz
    const double2* Q;     // inputs  [nin][ny][nx][kz]
    double2* F;           // outputs [3][ny][nx][kz]
    const double* Uy;     // physical-space base profiles at the Ny Gauss-Lobatto points: U, U', W, W' (4*Ny) or null
    const double* inv_dy; // 1/dy[ny] for the CFL maximum (Ny)
    double* cfl_max;      // device scalar: max over grid of (u_i + U_i)/dx_i  (signed, as in the reference)
    int ny0, nyn;
};
int zpass_launch(const ZPassParams& p, cudaStream_t stream);
bool zpass_general_supported(int Nz);

}  // namespace cfgpu
