// C-ABI of the device-resident state vectors (include/cfgpu.h, "vec" section): storage, field <-> vector maps,
// dot / nrm2 / axpy / scal.  Kernels in vecpack.cu.
#include <cmath>

#include "cfgpu_internal.h"
#include "vecpack.cuh"

using namespace cfgpu;

#define CF_ARG(cond, msg)        \
    do {                         \
        if (!(cond)) {           \
            set_last_error(msg); \
            return 1;            \
        }                        \
    } while (0)

struct cfgpu_vec_s {
    cfgpu_ctx ctx = nullptr;
    long long n = 0;
    double* d = nullptr;
};

static int pack_geom(cfgpu_field u, PackGeom& g) {
    CF_ARG(u->Nd == 3, "field2vector / vector2field: the field must have 3 components");
    CF_ARG(u->Ny >= 5, "field2vector / vector2field: Ny >= 5");
    g.Nx = u->Nx; g.Ny = u->Ny; g.Nz = u->Nz;
    g.Kx = u->Nx / 3 - 1; g.Kz = u->Nz / 3 - 1;
    CF_ARG(g.Kx >= 0 && g.Kz >= 0, "field2vector / vector2field: grid too small");
    g.Lx = u->Lx; g.Lz = u->Lz; g.a = u->a; g.b = u->b;
    return 0;
}
static int scalar_out(cfgpu_ctx ctx, double* dev, double* out_h) {
    CF_CUDA(cudaMemcpyAsync(out_h, dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" {

int cfgpu_vec_create(cfgpu_ctx ctx, long long n, cfgpu_vec* out) {
    CF_ARG(ctx && out && n >= 0, "cfgpu_vec_create: bad argument");
    cfgpu_vec v = new cfgpu_vec_s();
    v->ctx = ctx; v->n = n;
    if (n > 0) {
        if (dev_alloc(ctx, (void**)&v->d, (size_t)n * sizeof(double))) {
            delete v;
            set_last_error("cfgpu_vec_create: cudaMalloc failed");
            return 1;
        }
        CF_CUDA(cudaMemsetAsync(v->d, 0, (size_t)n * sizeof(double), ctx->stream));
    }
    *out = v;
    return 0;
}
int cfgpu_vec_destroy(cfgpu_vec v) {
    if (!v) return 0;
    dev_free(v->ctx, v->d);
    delete v;
    return 0;
}
int cfgpu_vec_size(cfgpu_vec v, long long* n) { *n = v->n; return 0; }
int cfgpu_vec_upload(cfgpu_vec v, const double* x_h) {
    CF_CUDA(cudaMemcpyAsync(v->d, x_h, (size_t)v->n * sizeof(double), cudaMemcpyHostToDevice, v->ctx->stream));
    CF_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return 0;
}
int cfgpu_vec_download(cfgpu_vec v, double* x_h) {
    CF_CUDA(cudaMemcpyAsync(x_h, v->d, (size_t)v->n * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
    CF_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return 0;
}
int cfgpu_vec_copy(cfgpu_vec dst, cfgpu_vec src) {
    CF_ARG(dst->n == src->n, "cfgpu_vec_copy: size mismatch");
    CF_CUDA(cudaMemcpyAsync(dst->d, src->d, (size_t)src->n * sizeof(double), cudaMemcpyDeviceToDevice, dst->ctx->stream));
    return 0;
}
int cfgpu_vec_zero(cfgpu_vec v) {
    if (v->n) CF_CUDA(cudaMemsetAsync(v->d, 0, (size_t)v->n * sizeof(double), v->ctx->stream));
    return 0;
}
int cfgpu_vec_dot(cfgpu_vec x, cfgpu_vec y, double* out_h) {
    CF_ARG(x->n == y->n, "cfgpu_vec_dot: size mismatch");
    cfgpu_ctx ctx = x->ctx;
    CF_ARG(ctx->comm.nranks == 1, "device state vectors are not distributed: one GPU per vector (replicas / one shot per GPU)");
    CF_TRY(ws_reserve(ctx->ws_red, 1 << 20));
    CF_TRY(vec_dot_launch(x->d, y->d, (long)x->n, ctx->ws_red.ptr + 8, ctx->ws_red.ptr, ctx->stream));
    return scalar_out(ctx, ctx->ws_red.ptr, out_h);
}
int cfgpu_vec_nrm2(cfgpu_vec x, double* out_h) {
    double s = 0;
    CF_TRY(cfgpu_vec_dot(x, x, &s));
    *out_h = std::sqrt(s);
    return 0;
}
int cfgpu_vec_axpy(cfgpu_vec y, double a, cfgpu_vec x) {
    CF_ARG(x->n == y->n, "cfgpu_vec_axpy: size mismatch");
    return vec_axpby_launch(a, x->d, 1.0, y->d, (long)y->n, y->ctx->stream);
}
int cfgpu_vec_axpby(cfgpu_vec y, double a, cfgpu_vec x, double b) {
    CF_ARG(x->n == y->n, "cfgpu_vec_axpby: size mismatch");
    return vec_axpby_launch(a, x->d, b, y->d, (long)y->n, y->ctx->stream);
}
int cfgpu_vec_scal(cfgpu_vec y, double s) { return vec_axpby_launch(0.0, y->d, s, y->d, (long)y->n, y->ctx->stream); }

int cfgpu_field2vector_size(cfgpu_field u, long long* n) {
    PackGeom g;
    CF_TRY(pack_geom(u, g));
    *n = pack_size(g);
    return 0;
}
int cfgpu_field2vector(cfgpu_field u, cfgpu_vec x) {
    PackGeom g;
    CF_TRY(pack_geom(u, g));
    CF_ARG(u->xzstate == CFGPU_SPECTRAL && u->ystate == CFGPU_SPECTRAL, "cfgpu_field2vector: the field must be spectral");
    CF_ARG(x->n >= pack_size(g), "cfgpu_field2vector: vector too short");
    CF_ARG(u->ctx->comm.nranks == 1, "cfgpu_field2vector: device state vectors are not distributed (one GPU per vector)");
    CF_TRY(field_serial(u));
    return field2vector_launch(u->dser, x->d, g, u->ctx->stream);
}
int cfgpu_vector2field(cfgpu_vec x, cfgpu_field u) {
    PackGeom g;
    CF_TRY(pack_geom(u, g));
    CF_ARG(x->n >= pack_size(g), "cfgpu_vector2field: vector too short");
    CF_ARG(u->ctx->comm.nranks == 1, "cfgpu_vector2field: device state vectors are not distributed (one GPU per vector)");
    CF_TRY(cfgpu_field_zero(u));          // flowfield.cpp:4572 (setToZero); the kernel writes the retained modes only
    CF_TRY(field_ser_alloc(u));
    CF_TRY(vector2field_launch(x->d, u->dser, g, u->ctx->stream));
    u->xzstate = u->ystate = CFGPU_SPECTRAL;
    u->padded = 1;
    u->clean_Kx = g.Kx; u->clean_Kz = g.Kz;
    return 0;
}

}  // extern "C"
