// Elementwise FlowField kernels and L2 forms -- see fieldops.cu.
#pragma once
#include "cf_common.cuh"

namespace cfgpu {
int axpby_launch(double* y, double a, const double* x, double b, const double* z, long n, cudaStream_t st);
int scale_launch(double* y, double s, long n, cudaStream_t st);
// FlowField *= FieldSymmetry on a spectral field in the reference layout, modes kx in [Kxlo, Kxhi], kz in [0, Kz]
int symmetry_launch(double* d, int Nx, int Ny, int Nz, int Nd, int Kxlo, int Kxhi, int Kz, int s, int sx, int sy, int sz, double ax, double az,
                    cudaStream_t st);
int zero_padded_launch(double* d, int Nx, int Ny, int Nz, int Nd, int Kx, int Kz, cudaStream_t st);
int profile_get_launch(const double* d, long off0_cplx, long rs_cplx, int Ny, double* out_dev, cudaStream_t st);
int profile_add_launch(double* d, long off0_cplx, long rs_cplx, int Ny, const double* in_dev, double s, cudaStream_t st);
// mode: 0 = sum |u|^2, 1 = sum |u-v|^2, 2 = <u,v>; result (times scale) written to out_dev
// rows mxi in [x0,x1) only (multi-GPU: the rank's own kx rows)
int l2form_launch(const double* u, const double* v, int mode, const double* W, int N, int Nx, int Nz, int Nd, int Kx, int Kz, int fullbox,
                  int x0, int x1, double scale, double* partial_dev, size_t partial_cap, double* out_dev, cudaStream_t st, double czw = 2.0);
// copy the retained rows mxi in [x0,x1), kz <= Kz of every (component, my) plane between a field (reference layout) and
// a compact buffer [row][mxi-x0][kz]; dir 0: field -> buffer, 1: buffer -> field
int tile_convert_launch(double* ser, double* tile, int Nx, int Ny, int Nz, int Nd, int Kx, int Kz, int x0, int nq, int TM, int dir,
                        cudaStream_t st);
int rows_pack_launch(double* field, double* buf, int Nx, int Nz, int nrows, int Kx, int Kz, int x0, int x1, int dir, cudaStream_t st);
}  // namespace cfgpu
