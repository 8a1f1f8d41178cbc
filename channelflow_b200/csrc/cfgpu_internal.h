// Internal structures behind the opaque C-ABI handles of include/cfgpu.h.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/cfgpu.h"
#include "cf_common.cuh"
#include "comm.cuh"
#include "fft_smem.cuh"
#include "tau.cuh"
#include "xzpass.cuh"
#include "ygemm.cuh"

namespace cfgpu {

struct YPlan {
    int N = 0;
    double a = 0, b = 0;
    int Nh = 0, Ne = 0, No = 0;
    int invMp = 0, invK1p = 0, invK2p = 0;
    int fwdMp = 0, fwdKp = 0;
    double *Ce = nullptr, *Co = nullptr, *CDe = nullptr, *CDo = nullptr, *Fe = nullptr, *Fo = nullptr;
    // forward transform followed by d/dy of the coefficients: even rows act on the difference tile, odd rows on the sum
    // tile ([0]: times 1, [1]: times 1/2 for the skew-symmetric form)
    double *GDe[2] = {nullptr, nullptr}, *GDo[2] = {nullptr, nullptr};
    double* Wgram = nullptr;  // [N][N] Chebyshev Gram weights for the L2 norms
    double* Wcheb = nullptr;  // [N][N] diagonal weights of the Chebyshev-weighted norm (chebyNorm2, diffops.cpp:260-300)
};

struct FftPlanHost {
    FftPlanDev dev;
    double2* tw = nullptr;
};

// retained-mode bookkeeping for one (Nx,Nz,Kx,Kz)
struct ModeBox {
    int Nx, Nz, Kx, Kz;
    long* runstart_full = nullptr;  // device [2Kx+1]: offset (doubles) of the kz-run of retained mx in a field row
};

// Tile-major field layout of de-aliased spectral fields: [tile][component][ny][TM] complex over the retained modes
// q = (mxi - x0)*(Kz+1) + kz of this rank in natural order, TM per tile (the tiling of the tau solver, tau.cuh).  What a CTA
// of the tau solve reads of the six history fields is then contiguous instead of 64-byte pieces a row apart.
struct TileGeom {
    int TM = 0, Kx = 0, Kz = 0, x0 = 0, x1 = 0;
    int nq() const { return (x1 - x0) * (Kz + 1); }
    long ntiles() const { return (nq() + TM - 1) / TM; }
    bool same(const TileGeom& o) const { return TM == o.TM && Kx == o.Kx && Kz == o.Kz && x0 == o.x0 && x1 == o.x1; }
};

struct Workspace {
    size_t bytes = 0;
    double* ptr = nullptr;
    unsigned long long gen = 0;  // bumped on every (re)allocation
    bool exported = false;       // mapped by other ranks through CUDA IPC: resized only collectively (ensure_peers)
};

}  // namespace cfgpu

struct cfgpu_ctx_s {
    int device = 0;
    cudaStream_t stream = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t comm_stream = 0;                 // all-to-all of the slab decomposition (overlaps the transforms)
    cudaEvent_t ev_cmp[4] = {nullptr, nullptr, nullptr, nullptr}, ev_com[4] = {nullptr, nullptr, nullptr, nullptr};
    std::map<std::tuple<int, double, double>, cfgpu::YPlan> yplans;
    std::map<int, cfgpu::FftPlanHost> fftplans;
    std::map<std::tuple<int, int, int, int>, cfgpu::ModeBox> boxes;
    cfgpu::Workspace ws_P, ws_Q, ws_S, ws_red, ws_G;  // ws_G: cfgpu_field_allgather staging
    cfgpu::Comm comm;  // rank / world size / collectives (single rank by default)
    // peer mappings of the other ranks' ws_P / ws_S (CUDA IPC), valid for the local base pointers recorded beside them
    void* peerP[cfgpu::COMM_MAXRANKS] = {nullptr};
    void* peerS[cfgpu::COMM_MAXRANKS] = {nullptr};
    unsigned long long peerP_gen = 0, peerS_gen = 0;  // generation of ws_P / ws_S the mappings belong to (0: none)
    // push all-to-all (comm.cuh): flag words [PUSH_SLOTS][COMM_MAXRANKS], a completion counter and an error word per rank
    cfgpu::Workspace ws_F;
    void* peerF[cfgpu::COMM_MAXRANKS] = {nullptr};
    unsigned long long peerF_gen = 0;
    unsigned long long push_seq[cfgpu::PUSH_SLOTS] = {0};
    std::vector<void*> graphs;  // GraphSlot*
    // Block cache of the small device allocations (dev_alloc / dev_free below): exact-size free lists.  A Newton-Krylov
    // search builds and drops a DNS (fields, tau tables, vectors: ~50 allocations) per Krylov vector; cudaMalloc / cudaFree
    // are device-synchronising and cost more than the 320 time steps between them on the grids those searches use.
    std::multimap<size_t, void*> pool_free_blocks;
    std::map<void*, size_t> pool_sizes;   // every live or cached pooled block
    size_t pool_cached_bytes = 0;
    bool capturing = false;
    // stage profiler
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[CFGPU_NSTAGES];
    size_t prof_used[CFGPU_NSTAGES] = {0};
    double prof_ms[CFGPU_NSTAGES] = {0};
    long long prof_calls[CFGPU_NSTAGES] = {0};
};

struct cfgpu_field_s {
    cfgpu_ctx ctx = nullptr;
    int Nx = 0, Ny = 0, Nz = 0, Nd = 0;
    double Lx = 0, Lz = 0, a = 0, b = 0;
    int xzstate = CFGPU_SPECTRAL, ystate = CFGPU_SPECTRAL;
    int padded = 0;
    int clean_Kx = -1, clean_Kz = -1;  // all modes outside this box are known to be exactly zero (-1: unknown)
    // The data lives in one of two buffers.  `dser` is the reference's serial layout [i][my][mx][mz] (allocated on first use: a field
    // without it is zero wherever its tile buffer does not say otherwise);
    // `dtile` (lazily allocated) is the tile-major layout of the retained box, produced and consumed by the DNS hot path
    // (cfgpu_nse_solve / cfgpu_nse_nonlinear / cfgpu_nse_linear).  layout says which one is current for the retained box; outside the box
    // the field is zero if tile_outside_zero, else whatever dser holds there.  Everything but the hot path goes through
    // field_serial(), which converts back on demand.
    double* dser = nullptr;
    long long n = 0;  // doubles
    double* dtile = nullptr;
    long long ntile = 0;
    int layout = 0;   // 0: dser current, 1: dtile current
    bool tile_outside_zero = false;
    cfgpu::TileGeom tg;
    long long tile_compstride() const { return (long long)Ny * tg.TM * 2; }
    long long tile_stride() const { return tile_compstride() * Nd; }
    int Nzpad() const { return 2 * (Nz / 2 + 1); }
    int Mz() const { return Nz / 2 + 1; }
    long long rowstride() const { return (long long)Nx * Nzpad(); }
    long long compstride() const { return rowstride() * Ny; }
};

struct cfgpu_nse_s {
    cfgpu_ctx ctx = nullptr;
    int Nx = 0, Ny = 0, Nz = 0;
    double Lx = 0, Lz = 0, a = 0, b = 0;
    cfgpu_nse_config cfg;
    int Nyd = 0, Kx = 0, Kz = 0;
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;  // owned kx rows (mxi) and y planes of this rank (whole ranges on one GPU)
    int nq = 0;                          // owned retained modes
    cfgpu::ModeGeom geom;
    double* d_base = nullptr;   // device: Ubaseyy[Ny], Wbaseyy[Ny], phys U,U',W,W' [4*Ny], inv_dy[Ny], Ubase[Ny], Wbase[Ny]
    bool has_Ubaseyy = false, has_Wbaseyy = false;
    double lin_base_dPdx = 0, lin_base_dPdz = 0;  // nu*(Ubase'(b)-Ubase'(a))/Ly (nse.cpp:464-469)
    double* d_scal = nullptr;   // device scalars: [0] cfl max, [1] dPdxAct, [2] dPdzAct
    std::vector<double> lambda_t;
    std::vector<cfgpu::TauData> tau;  // one per substep
    int TM_solve = 8, TM_lin = 8;
    bool use_tile = false;       // hot-path fields (solve outputs, nonlinear term) are kept tile-major
    cfgpu::TileGeom tg;
    long* d_tilestart = nullptr; // device [ntiles]: offset of tile t of a 3-component tile-major field
    double** d_rows[2] = {nullptr, nullptr};  // peer row tables of the inverse y-GEMM outputs (5- and 3-field staging)
    double** d_rows_self[2] = {nullptr, nullptr};  // push mode: own planes -> own staging, other rows -> local pencil buffer
    unsigned long long rows_genS = 0, rows_genP = 0;  // generations of ws_S / ws_P the row tables were built for
    int rows_nranks = 0;
    cfgpu_field s_u = nullptr, s_t = nullptr;  // scratch fields of the non-rotational nonlinear terms (3 and 9 components)
};

namespace cfgpu {
int field_serial(cfgpu_field f);                                           // make dser current (tile -> serial if needed)
int field_ser_alloc(cfgpu_field f);                                        // allocate the (zero-filled) serial buffer if absent
// Device memory for fields, vectors and operator tables.  Blocks of at most POOL_MAX_BLOCK bytes of a single-GPU context are
// recycled through the context's free lists (stream order on ctx->stream makes reuse safe); larger ones, and everything in a
// multi-GPU context (peer-mapped, several streams), go straight to cudaMalloc / cudaFree.
int dev_alloc(cfgpu_ctx ctx, void** p, size_t bytes);
void dev_free(cfgpu_ctx ctx, void* p);
int field_serial_output(cfgpu_field f);                                    // dser becomes current, contents to be written
int field_tile(cfgpu_field f, const TileGeom& g);                          // make dtile current (serial -> tile if needed)
int field_tile_output(cfgpu_field f, const TileGeom& g, bool outside_zero);  // dtile becomes current, contents to be written
int get_yplan(cfgpu_ctx ctx, int N, double a, double b, const YPlan** out);
int get_fftplan(cfgpu_ctx ctx, int N, const FftPlanDev** out);
const FftPlanDev* yfft_plan(cfgpu_ctx ctx, int Ny);
const FftPlanDev* yfft_half_plan(cfgpu_ctx ctx, int Ny);
int get_box(cfgpu_ctx ctx, int Nx, int Nz, int Kx, int Kz, const ModeBox** out);
int ws_reserve(Workspace& w, size_t bytes);
int stage_begin(cfgpu_ctx ctx, int stage, cudaStream_t stream = 0);  // 0: the context's compute stream
int stage_end(cfgpu_ctx ctx, int stage, cudaStream_t stream = 0);
struct StageTimer {  // RAII bracket around the launches of one pipeline stage
    cfgpu_ctx ctx; int stage; cudaStream_t stream;
    StageTimer(cfgpu_ctx c, int s, cudaStream_t st = 0) : ctx(c), stage(s), stream(st) { stage_begin(ctx, stage, stream); }
    ~StageTimer() { stage_end(ctx, stage, stream); }
};
// host-side Chebyshev helpers (long double internally)
void cheb_diff_host(const std::vector<double>& u, std::vector<double>& d, double a, double b);
void cheb_to_physical_host(const std::vector<double>& c, std::vector<double>& u);
}  // namespace cfgpu
