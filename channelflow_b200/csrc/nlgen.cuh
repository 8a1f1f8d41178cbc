// Kernels for the non-rotational nonlinear terms (convection, divergence, skew-symmetric, alternating, linearized
// about a profile).  They follow the reference's own sequence of operations (diffops.cpp:1784-1941 grad, :2336-2388
// outer, :2470-2558 div, :3586-3643 dotgrad, :3110-3134 divergenceNL, :3142-3286 skewsymmetricNL, :3288-3365
// linearizedNL) on full-layout device fields, with the generic FlowField transforms in between; each kernel fuses
// the reference's several sweeps over the array into one HBM-streaming pass.
#pragma once
#include "cf_common.cuh"

namespace cfgpu {

struct FieldGeom {
    int Nx, Ny, Nz;
    double Lx, Lz, a, b;
};

// gradf[3i+j] = d u_i / d x_j, spectral state, u has 3 components (diffops.cpp:1866-1935)
int grad3_launch(const double* u, double* gradu, const FieldGeom& g, cudaStream_t st);
// pointwise, physical state.  conv_coef != 0: f_i = conv_coef * sum_j u_j G_ij (+ Coriolis term f_x -= rot*v, f_y += rot*u);
// do_outer: G_ij <- u_i u_j afterwards (in place).  n = Ny*Nx*Nzpad reals per component.
int pointwise_nl_launch(const double* u, double* G, double* f, double conv_coef, int do_outer, double rot, long n, cudaStream_t st);
// f_i (+)= coef * d/dx_j T_ij, spectral state (diffops.cpp:2517-2549, 3222-3279)
int div9_launch(const double* T, double* f, double coef, int accumulate, const FieldGeom& g, cudaStream_t st);
// linearizedNL in the (Spectral xz, Physical y) state; prof = physical U, U', W, W' (4*Ny)
int linearized_launch(const double* u, double* f, const double* prof, const FieldGeom& g, cudaStream_t st);
// u(0,0 mode) += s*(Ubase e_x + Wbase e_z) - s*Vsuck e_y on coefficient 0 (nse.cpp:28-36, 81-88); spectral state
int add_base00_launch(double* u, const double* Ubase, const double* Wbase, double Vsuck, double s, const FieldGeom& g, cudaStream_t st);
// f_x -= rot*v, f_y += rot*u over the raw real arrays for nz < Nz (nse.cpp:63-79)
int coriolis_launch(const double* u, double* f, double rot, long n, int Nz, cudaStream_t st);

}  // namespace cfgpu
