// Differential and pointwise operators on whole FlowFields -- see diffops.cu.
#pragma once
#include "cf_common.cuh"
#include "nlgen.cuh"

namespace cfgpu {

constexpr int DIFF_MAXTERMS = 27;
constexpr int DIFF_MAXPEROUT = 3;
struct DiffTerm {
    int out, in;     // component indices
    int nx, ny, nz;  // derivative orders; ny in {0,1,2}
    double coef;
};
struct DiffOpParams {
    FieldGeom g;
    int nout, nterms;
    DiffTerm t[DIFF_MAXTERMS];  // at most DIFF_MAXPEROUT terms per output component
};
// out_c = sum over the terms with t.out == c of coef * d^nx/dx^nx d^ny/dy^ny d^nz/dz^nz in_{t.in}; spectral state, out != in
int diffop_launch(const double* in, double* out, const DiffOpParams& p, cudaStream_t st);

enum PointOp { PW_CROSS = 0, PW_OUTER = 1, PW_DOT = 2, PW_NORM2 = 3, PW_NORM = 4, PW_ENERGY = 5, PW_MUL = 6 };
// physical state, n reals per component.  fd, gd: components of f and g
int pointwise_launch(int op, const double* f, const double* g, double* out, int fd, int gd, long n, cudaStream_t st);
// sum over components and all (mx, mz) of |u(a) - v(a)|^2 + |u(b) - v(b)|^2 (v may be null); ystate: 1 spectral, 0 physical
int bcnorm2_launch(const double* u, const double* v, int Nx, int Ny, int Nz, int Nd, int yspectral, double* partial, size_t cap,
                   double* out_dev, cudaStream_t st);
// homogeneous Neumann correction of the pressure Poisson problem (poissonsolver.cpp:352-431): p, v spectral scalar fields
// (serial layout), out = g at the Gauss-Lobatto points (xz-spectral, y-physical)
int pressure_neumann_launch(const double* p, const double* v, double nu, const FieldGeom& g, double* out, cudaStream_t st);

}  // namespace cfgpu
