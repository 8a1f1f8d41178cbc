/* cfgpu.h -- C-ABI of the B200 (sm_100a) device layer behind Channelflow's FlowField / NSE / DNS classes.
 *
 * The reference (epfl-ecps/channelflow, pure C++/FFTW/MPI) has no FFI boundary of its own; its "plugin" surface
 * is the C++ class API.  This header is the boundary a maintainer binds instead of FFTW + the scalar loops:
 * each entry point names the reference code it replaces (paths relative to the reference root).
 * Conventions: extern "C", opaque handles, plain pointers and sizes, `int` status (0 = ok, message via
 * cfgpu_last_error()).  All field data stay resident in HBM in FP64; host pointers are suffixed _h.  Calls are
 * asynchronous on the context's stream unless they return data to the host.
 *
 * Field storage is the reference's serial layout (channelflow/flowfield.h:370-402):
 *   real    rdata[nz + Nzpad*(nx + Nx*(ny + Ny*i))],  Nzpad = 2*(Nz/2+1)
 *   complex cdata[mz + Mz  *(mx + Nx*(my + Ny*i))],   Mz = Nz/2+1
 * so upload/download are plain copies of FlowField::rdata_.
 */
#ifndef CFGPU_H
#define CFGPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfgpu_ctx_s* cfgpu_ctx;
typedef struct cfgpu_field_s* cfgpu_field;
typedef struct cfgpu_nse_s* cfgpu_nse;
typedef struct cfgpu_vec_s* cfgpu_vec;

enum { CFGPU_PHYSICAL = 0, CFGPU_SPECTRAL = 1 }; /* cfbasics/mathdefs.h: enum fieldstate */

/* ---------------------------------------------------------------- context */
const char* cfgpu_last_error(void);
const char* cfgpu_version(void);
/* replaces CfMPI::getInstance + FFTW planner state (channelflow/cfmpi.cpp:66-127, flowfield.cpp:577-667) */
int cfgpu_init(int device, cfgpu_ctx* ctx);
int cfgpu_finalize(cfgpu_ctx ctx);
int cfgpu_sync(cfgpu_ctx ctx);
/* number of kernels this library has launched since cfgpu_init (bench.py's gpu_launches) */
int cfgpu_launch_count(cfgpu_ctx ctx, long long* n);
/* time a region on the context's stream with CUDA events */
int cfgpu_timer_start(cfgpu_ctx ctx);
int cfgpu_timer_stop(cfgpu_ctx ctx, double* ms);
/* per-stage device timing of the DNS pipeline (CUDA events on the launching stream, accumulated per stage):
 * stages: 0 inverse y-GEMM, 1 inverse x-pass, 2 z-pass+nonlinear, 3 forward x-pass, 4 forward y-GEMM,
 *         5 tau solve (+fused RHS), 6 linear term, 7 tau setup, 8 slab all-to-all (multi-GPU) */
#define CFGPU_NSTAGES 9
int cfgpu_profile_enable(cfgpu_ctx ctx, int on);
int cfgpu_profile_read(cfgpu_ctx ctx, double* ms_h /* [CFGPU_NSTAGES] */, long long* calls_h /* [CFGPU_NSTAGES] */, int reset);
/* CUDA-graph replay of a fixed sequence of calls on the context's stream (launch-bound small grids: `order` consecutive
 * SBDF steps, dnsalgo.cpp:195-260, return every history buffer to its role).  Between begin and end the calls are recorded,
 * not executed; a call that allocates or synchronises fails and the capture must be dropped with cfgpu_graph_abort.
 * Single-GPU contexts only. */
int cfgpu_graph_begin(cfgpu_ctx ctx);
int cfgpu_graph_end(cfgpu_ctx ctx, int* graph_id);
int cfgpu_graph_abort(cfgpu_ctx ctx);
int cfgpu_graph_launch(cfgpu_ctx ctx, int graph_id);
int cfgpu_graph_destroy(cfgpu_ctx ctx, int graph_id);

/* ---------------------------------------------------------------- multi-GPU (one process per GPU)
 * replaces CfMPI (cfmpi.cpp:66-127) and the FFTW-MPI transposes inside makePhysical/makeSpectral (flowfield.cpp:577-667,
 * 1850-1997): slab decomposition, kx rows distributed in the spectral state and y planes in the physical state, one
 * all-to-all per direction inside cfgpu_nse_nonlinear / cfgpu_nse_cflfactor, all-reduce inside norms and CFL.
 * Fields keep the full (serial) shape on every rank; a rank owns -- and keeps valid -- the kx rows of its range. */
int cfgpu_comm_unique_id(void* id128_h /* 128 bytes, from rank 0, to be broadcast by the launcher */);
int cfgpu_comm_init_nccl(cfgpu_ctx ctx, int rank, int nranks, const void* id128_h);
/* host-supplied collectives (CPU tests of the decomposition logic: gloo); pointers are the library's device pointers */
typedef int (*cfgpu_exchange_fn)(void* user, int nmsg, const int* peer, const void* const* sendbuf, const long long* sendbytes,
                                 void* const* recvbuf, const long long* recvbytes);
typedef int (*cfgpu_allreduce_fn)(void* user, double* buf, int n, int op /* 0 sum, 1 max */);
int cfgpu_comm_init_external(cfgpu_ctx ctx, int rank, int nranks, cfgpu_exchange_fn ex, cfgpu_allreduce_fn ar, void* user);
int cfgpu_comm_rank(cfgpu_ctx ctx, int* rank, int* nranks);
/* owned ranges: retained kx rows mxi in [x0,x1) of 2Kx+1 (kx = 0..Kx, -Kx..-1) and y planes [y0,y1) */
int cfgpu_comm_ranges(cfgpu_ctx ctx, int nmx, int Ny, int rank, int* x0, int* x1, int* y0, int* y1);
/* make every rank's copy of a spectral, de-aliased field complete (all-gather of the owned kx rows): I/O, tests */
int cfgpu_field_allgather(cfgpu_field f);

/* ---------------------------------------------------------------- FlowField storage
 * FlowField ctor/resize/copy/assign/swap/setToZero (flowfield.cpp:466-575, 96-102, 452-460, 4076-4090, 2229-2233) */
int cfgpu_field_create(cfgpu_ctx ctx, int Nx, int Ny, int Nz, int Nd, double Lx, double Lz, double a, double b,
                       cfgpu_field* out);
int cfgpu_field_destroy(cfgpu_field f);
/* page-locked host memory for the FlowField host mirror (replaces fftw_malloc of the host array, flowfield.cpp:495-497):
 * transfers from/to it run at full PCIe rate and can be asynchronous */
int cfgpu_host_alloc(void** ptr_h, unsigned long long bytes);
int cfgpu_host_free(void* ptr_h);
int cfgpu_field_upload(cfgpu_field f, const double* data_h, int xzstate, int ystate);
int cfgpu_field_download(cfgpu_field f, double* data_h);
/* Same for an xz-spectral, de-aliased ("padded", flowfield.h:578-584) field: only the retained box |kx| <= Nx/3-1,
 * kz <= Nz/3-1 travels (44 % of the array; two pitched 3-D copies), the rest of the device field is zero / the rest of
 * the host array is left untouched.  With several GPUs only the rank's own kx rows travel. */
int cfgpu_field_upload_padded(cfgpu_field f, const double* data_h, int ystate);
int cfgpu_field_download_padded(cfgpu_field f, double* data_h);
int cfgpu_field_copy(cfgpu_field dst, cfgpu_field src);
int cfgpu_field_swap(cfgpu_field a, cfgpu_field b);
/* component js of src -> component jd of dst (FlowField::operator[](int), flowfield.cpp:1532-1560) */
int cfgpu_field_copy_component(cfgpu_field dst, int jd, cfgpu_field src, int js);
int cfgpu_field_zero(cfgpu_field f);
int cfgpu_field_set_state(cfgpu_field f, int xzstate, int ystate);
int cfgpu_field_get_state(cfgpu_field f, int* xzstate, int* ystate);
int cfgpu_field_set_padded(cfgpu_field f, int padded);
int cfgpu_field_get_padded(cfgpu_field f, int* padded);
/* Device pointer to the data in the reference's serial layout (flowfield.h:370-402).  Fields produced by the time-stepping
 * calls (cfgpu_nse_solve / _nonlinear / _linear) may live in an internal tile-major layout; this call -- like every
 * call that is not one of those three -- converts back first.  The pointer is valid until the next cfgpu_nse_* call on
 * the field. */
int cfgpu_field_device_ptr(cfgpu_field f, double** dptr, long long* ndoubles);
/* FlowField::add / operator*= / += / -= (flowfield.h:596-615, flowfield.cpp:1459-1469, 1759-1771):  y += a*x + b*z */
int cfgpu_field_axpby(cfgpu_field y, double a, cfgpu_field x, double b, cfgpu_field z /* may be NULL */);
int cfgpu_field_scale(cfgpu_field y, double s);
/* one Fourier mode's Chebyshev profile: FlowField::profile / operator+=(ChebyCoeff) (flowfield.cpp:1473-1530) */
int cfgpu_field_get_profile(cfgpu_field f, int mx, int mz, int i, double* re_im_h /* 2*Ny, interleaved */);
int cfgpu_field_add_profile(cfgpu_field f, int mx, int mz, int i, const double* re_im_h, double scale);
/* FlowField::zeroPaddedModes (flowfield.cpp:2235-2255) */
int cfgpu_field_zero_padded_modes(cfgpu_field f);

/* ---------------------------------------------------------------- transforms
 * FlowField::makePhysical_y / makeSpectral_y (flowfield.cpp:1888-1987): DMMA DCT-I contraction
 * FlowField::makePhysical_xz / makeSpectral_xz (flowfield.cpp:1850-1886): batched c2c-x + paired c2r/r2c-z */
int cfgpu_field_make_physical_y(cfgpu_field f);
int cfgpu_field_make_spectral_y(cfgpu_field f);
int cfgpu_field_make_physical_xz(cfgpu_field f);
int cfgpu_field_make_spectral_xz(cfgpu_field f);
int cfgpu_field_make_physical(cfgpu_field f); /* y then xz (flowfield.cpp:1993-1997) */
int cfgpu_field_make_spectral(cfgpu_field f); /* xz then y */

/* ---------------------------------------------------------------- norms
 * L2Norm2 / L2Dist2 / L2InnerProduct of FlowFields (diffops.cpp:417-487, 353-413, 489-541) with the Chebyshev
 * Gram weights of chebyshev.cpp:758-802 (weights evaluated in FP64) */
int cfgpu_l2norm2(cfgpu_field u, int normalize, double* out_h);
/* L2Norm2_3d (diffops.cpp:700-740): the same sum without the kx = 0 modes */
int cfgpu_l2norm2_3d(cfgpu_field u, int normalize, double* out_h);
int cfgpu_l2dist2(cfgpu_field u, cfgpu_field v, int normalize, double* out_h);
int cfgpu_l2ip(cfgpu_field u, cfgpu_field v, int normalize, double* out_h);
/* chebyNorm2 / chebyDist2 / chebyInnerProduct (diffops.cpp:259-350): mode 0 / 1 / 2 with the Chebyshev weight in y */
int cfgpu_chebyform(cfgpu_field u, cfgpu_field v, int mode, int normalize, double* out_h);

/* L2Norm2 / L2Dist2 / L2InnerProduct (mode 0 / 1 / 2) over the modes |kx| <= kxmax, kz <= kzmax (diffops.cpp:543-700);
 * cz = 0 drops the factor 2 of the kz > 0 modes (divNorm2's convention, diffops.cpp:91-116) */
int cfgpu_l2form_box(cfgpu_field u, cfgpu_field v, int mode, int kxmax, int kzmax, int cz, int normalize, double* out_h);
/* bcNorm2 / bcDist2 (diffops.cpp:18-83); v may be NULL */
int cfgpu_bcnorm2(cfgpu_field u, cfgpu_field v, int normalize, double* out_h);

/* ---------------------------------------------------------------- whole-field operators (diagnostics, initial data)
 * Linear differential operator on a spectral field: out[out_c[k]] += coef[k] d^nx/dx^nx d^ny/dy^ny d^nz/dz^nz in[in_c[k]]
 * (at most 3 terms per output component, ny <= 2): xdiff/ydiff/zdiff/diff, grad, lapl, curl, div of diffops.cpp:1650-2558 */
int cfgpu_field_diffop(cfgpu_field out, cfgpu_field in, int nterms, const int* out_c, const int* in_c, const int* nx, const int* ny,
                       const int* nz, const double* coef);
/* Pointwise products of physical fields: op 0 cross, 1 outer (f_i g_j -> component i*gd+j), 2 dot, 3 |f|^2, 4 |f|, 5 energy
 * 1/2 |f|^2, 6 componentwise product (diffops.cpp:2336-2700) */
int cfgpu_field_pointwise(int op, cfgpu_field out, cfgpu_field f, cfgpu_field g /* NULL for ops 3-5 */);

/* FlowField::operator*=(const FieldSymmetry&) (flowfield.cpp:1274-1433), spectral field, in place:
 * (u,v,w)(x,y,z) -> s (sx u, sy v, sz w)(sx x + ax Lx, sy y, sz z + az Lz) */
int cfgpu_field_symmetry(cfgpu_field f, int s, int sx, int sy, int sz, double ax, double az);
/* PoissonSolver::solve (poissonsolver.cpp:146-202): lapl u = f for every stored Fourier mode of every component, Dirichlet
 * data zero (bc == NULL) or the wall values of bc (poissonsolver.cpp:173-202) */
int cfgpu_poisson_solve(cfgpu_field u, cfgpu_field f, cfgpu_field bc);
/* PressureSolver::solve, step II (poissonsolver.cpp:352-431): g (xz-spectral, y-physical scalar field) = the homogeneous
 * solution whose addition to the Dirichlet pressure p gives dp/dy = nu d2v/dy2 at both walls */
int cfgpu_pressure_neumann(cfgpu_field g, cfgpu_field p, cfgpu_field u, double nu);

/* ---------------------------------------------------------------- NSE operator (channelflow/nse.h:23-141)
 * enums follow channelflow/dnsflags.h:24-41 */
typedef struct {
    double nu;
    double Vsuck;
    double rotation;
    int nonlinearity;  /* NonlinearMethod: 0 Rotational, 1 Convection, 2 Divergence, 3 SkewSymmetric, 4/5 Alternating, 6 Linearized */
    int dealias_xz;    /* DNSFlags::dealias_xz() */
    int dealias_y;     /* DNSFlags::dealias_y()  */
    int taucorrection;
    int constraint;    /* MeanConstraint: 0 PressureGradient, 1 BulkVelocity */
    double dPdxRef, dPdzRef;           /* NSE::dPdxRef_, dPdzRef_ */
    double UbulkRef_minus_base;        /* UbulkRef_ - UbulkBase_ (nse.cpp:536-537) */
    double WbulkRef_minus_base;
} cfgpu_nse_config;

/* NSE::NSE(fields, flags) (nse.cpp:212-289): geometry, base flow (Chebyshev coefficients, length Ny), work space */
int cfgpu_nse_create(cfgpu_ctx ctx, int Nx, int Ny, int Nz, double Lx, double Lz, double a, double b,
                     const cfgpu_nse_config* cfg, const double* Ubase_h, const double* Wbase_h, cfgpu_nse* out);
int cfgpu_nse_destroy(cfgpu_nse nse);
int cfgpu_nse_set_constraint(cfgpu_nse nse, int constraint, double dPdxRef, double dPdzRef,
                             double UbulkRef_minus_base, double WbulkRef_minus_base);
/* NSE::reset_lambda (nse.cpp:673-705): batched TauSolver/HelmholtzSolver/BandedTridiag setup for every retained mode */
int cfgpu_nse_reset_lambda(cfgpu_nse nse, const double* lambda_t_h, int nsub);
/* NSE::nonlinear = navierstokesNL + zeroPaddedModes (nse.cpp:12-91, 383-391); u is not modified */
int cfgpu_nse_nonlinear(cfgpu_nse nse, cfgpu_field u, cfgpu_field f);
/* NSE::solve (nse.cpp:479-575) fused with the RHS accumulation of the time steppers (dnsalgo.cpp:217-224, 446-449,
 * 668-673):  rhs = sum_j coef[j]*term[j] ;  solve nu u'' - lambda_s u - grad q = -rhs, div u = 0 per mode */
int cfgpu_nse_solve(cfgpu_nse nse, int s, int nterms, const double* coef_h, const cfgpu_field* terms,
                    cfgpu_field uout, cfgpu_field qout);
/* NSE::linear (nse.cpp:393-477):  L = nu u'' - nu kappa^2 u - grad q (+ mean-mode constants) */
int cfgpu_nse_linear(cfgpu_nse nse, cfgpu_field u, cfgpu_field q, cfgpu_field L);
/* FlowField::CFLfactor(Ubase,Wbase) (flowfield.cpp:4035-4068): max over grid of (u_i+U_i)/dx_i (no abs) */
int cfgpu_nse_cflfactor(cfgpu_nse nse, cfgpu_field u, double* out_h);
/* dPdxAct_/dPdzAct_ computed by the bulk-velocity-constrained solve (nse.cpp:536) */
int cfgpu_nse_get_dPd(cfgpu_nse nse, double* dPdx_h, double* dPdz_h);

/* ---------------------------------------------------------------- 1-d solver classes (tests, tools; host arrays in and out)
 * HelmholtzSolver::solve (helmholtz.cpp:79-95): ncols real systems nu u'' - lambda u = f, u(a) = ua, u(b) = ub */
int cfgpu_helmholtz_solve(cfgpu_ctx ctx, int N, double a, double b, double lambda, double nu, int ncols, const double* f_h /* [ncols][N] */,
                          const double* ua_h, const double* ub_h, double* u_h);
/* BandedTridiag (bandedtridiag.cpp:212-333) on its own storage a[4M-2], invdiag[M]: op 0 ULdecomp (in place), 1 ULsolveStrided
 * (x in place), 2 multiplyStrided (y = A x) */
int cfgpu_tridiag(cfgpu_ctx ctx, int op, int M, double* a_h, double* invdiag_h, double* x_h, double* y_h, int nx, int offset, int stride);
/* TauSolver::solve for one Fourier mode (tausolver.cpp:347-450): R_h = Rx, Ry, Rz as [3][N] complex, out_h = u, v, w, P as
 * [4][N] complex; constraint 1 = mean mode with bulk velocities umean, wmean, returning dPdx, dPdz in dPd_h[2] */
int cfgpu_tausolve_mode(cfgpu_ctx ctx, int N, int kx, int kz, double Lx, double Lz, double a, double b, double lambda, double nu,
                        int taucorrection, int constraint, double umean, double wmean, const double* R_h, double* out_h, double* dPd_h);

/* ---------------------------------------------------------------- state vectors of nsolver (device resident)
 * field2vector / vector2field (channelflow/flowfield.cpp:4448-4760 with fixDiri / fixDiriMean, utilfuncs.cpp:712-765): the
 * map between a divergence-free, no-slip velocity field and the vector of its independent real coefficients, one warp
 * per Fourier mode; the vector stays in HBM.  cfgpu_vec_dot / nrm2 / axpy / scal replace the Eigen VectorXd algebra of
 * the Krylov iterations (cfbasics/cfbasics.h:711-780 L2IP / L2Norm, nsolver/gmres.cpp:37-102 modified Gram-Schmidt,
 * nsolver/arnoldi.cpp).  Vectors are not distributed: one GPU per vector (replicas / one shot per GPU). */
int cfgpu_vec_create(cfgpu_ctx ctx, long long n, cfgpu_vec* out);
int cfgpu_vec_destroy(cfgpu_vec v);
int cfgpu_vec_size(cfgpu_vec v, long long* n);
int cfgpu_vec_upload(cfgpu_vec v, const double* x_h);
int cfgpu_vec_download(cfgpu_vec v, double* x_h);
int cfgpu_vec_copy(cfgpu_vec dst, cfgpu_vec src);
int cfgpu_vec_zero(cfgpu_vec v);
int cfgpu_vec_dot(cfgpu_vec x, cfgpu_vec y, double* out_h);
int cfgpu_vec_nrm2(cfgpu_vec x, double* out_h);
int cfgpu_vec_axpy(cfgpu_vec y, double a, cfgpu_vec x);           /* y += a x */
int cfgpu_vec_axpby(cfgpu_vec y, double a, cfgpu_vec x, double b); /* y = a x + b y */
int cfgpu_vec_scal(cfgpu_vec y, double s);
int cfgpu_field2vector_size(cfgpu_field u, long long* n);         /* flowfield.cpp:4448-4479 */
int cfgpu_field2vector(cfgpu_field u, cfgpu_vec x);               /* flowfield.cpp:4481-4563 */
int cfgpu_vector2field(cfgpu_vec x, cfgpu_field u);               /* flowfield.cpp:4565-4752 */

#ifdef __cplusplus
}
#endif
#endif
