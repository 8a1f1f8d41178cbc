// field2vector / vector2field and Krylov-vector algebra on the device -- see vecpack.cu.
#pragma once
#include "cf_common.cuh"

namespace cfgpu {

struct PackGeom {
    int Nx, Ny, Nz, Kx, Kz;   // grid and the de-aliased box |kx| <= Kx, kz <= Kz (flowfield.h:578-584)
    double Lx, Lz, a, b;
};
inline int pack_nmodes(const PackGeom& g) { return 1 + g.Kx + g.Kz + 2 * g.Kx * g.Kz; }
// field2vector_size (flowfield.cpp:4448-4479)
inline long pack_size(const PackGeom& g) {
    return 2L * (g.Ny - 2) + (long)(g.Kx + g.Kz + 2 * g.Kx * g.Kz) * (2L * (g.Ny - 2) + 2L * (g.Ny - 4));
}

int field2vector_launch(const double* u_serial, double* a, const PackGeom& g, cudaStream_t st);
int vector2field_launch(const double* a, double* u_serial /* zeroed by the caller */, const PackGeom& g, cudaStream_t st);
int vec_dot_launch(const double* x, const double* y, long n, double* partial, double* out_dev, cudaStream_t st);
int vec_axpby_launch(double a, const double* x, double b, double* y, long n, cudaStream_t st);  // y = a x + b y
int vec_partial_capacity();

}  // namespace cfgpu
