// NSE operator entry points of the C-ABI (include/cfgpu.h): nonlinear term pipeline, batched tau solve,
// linear term, CFL.  Host orchestration only; the kernels are in ygemm.cu, xzpass.cu, tau.cu.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cfgpu_internal.h"
#include "fieldops.cuh"
#include "nlgen.cuh"

using namespace cfgpu;

#define CF_ARG(cond, msg)        \
    do {                         \
        if (!(cond)) {           \
            set_last_error(msg); \
            return 1;            \
        }                        \
    } while (0)

// every field handed to an NSE entry point must have the operator's grid: the kernels index with the operator's strides
static bool shape_ok(const cfgpu_nse nse, const cfgpu_field f, int Nd) {
    return f && f->Nx == nse->Nx && f->Ny == nse->Ny && f->Nz == nse->Nz && f->Nd == Nd;
}

namespace cfgpu {
int linear_launch(const TauSolveParams& p, const double* u, const double* q, double* L, cudaStream_t stream);
}

// kz columns per x-pass CTA.  The transforms run in place (one shared-memory buffer), so 8 columns = 128-byte pieces
// of the pencil rows fit three CTAs per SM at Nx = 512.
static int pick_TZ(int Nx) {
    static const int forced = getenv("CF_XP_TZ") ? atoi(getenv("CF_XP_TZ")) : 0;
    if (forced > 0) return forced;
    int tz = 4608 / Nx;
    int p = 16;
    while (p > tz && p > 2) p >>= 1;
    return p;
}
// ... the forward x-pass keeps the two-buffer Stockham transform (its in-place variant measured slower: 0.92 against
// 0.67 ms) and therefore half the columns (measured at Nx = 512: 2 -> 1.93 ms, 4 -> 1.77 ms, 6 -> 1.87 ms, 8 -> 2.70 ms)
// long lines: the in-place kernel (one buffer) takes twice the columns of the Stockham one
static bool forward_inplace(int Nx) {
    static const int forced = getenv("CF_XPF_INPLACE") ? atoi(getenv("CF_XPF_INPLACE")) : -1;
    return forced >= 0 ? forced != 0 : Nx >= 1024;
}
static int pick_TZ_forward(int Nx) {
    static const int forced = getenv("CF_XPF_TZ") ? atoi(getenv("CF_XPF_TZ")) : 0;
    if (forced > 0) return forced;
    int tz = (forward_inplace(Nx) ? 4608 : 2304) / Nx;
    int p = 16;
    while (p > tz && p > 1) p >>= 1;  // one column per CTA for Nx > 2304 (weak scaling of the C4 grid: Nx = 4096 at 8 GPUs)
    return p;
}
static int pick_TL(int Nx, int Nz, int npair) {
    // 2 buffers * Nz * npair*TL * 16 B <= ~96 KB
    int tl = 16;
    while (tl > 1 && (size_t)2 * Nz * npair * tl * 16 > 96 * 1024) tl >>= 1;
    while (tl > 1 && Nx % tl) tl >>= 1;
    return tl;
}

// all-to-all between the kx-slab pencils P[f][Ny][mxi in X_rank][kz] and the y-slab staging
// S[s][f][y in Y_rank][mxi in X_s][kz] (comm.cuh); dir 0: P -> S (inverse transform), 1: S -> P (forward)
static int slab_exchange(cfgpu_nse nse, int nf, int dir, const int* fields, int nsel, cudaStream_t stream) {
    cfgpu_ctx ctx = nse->ctx;
    Comm& cm = ctx->comm;
    if (cm.nranks == 1) return 0;
    const int nmx = 2 * nse->Kx + 1, nkz = nse->Kz + 1;
    const int nxl = nse->x1 - nse->x0, nyl = nse->y1 - nse->y0;
    double2* P = reinterpret_cast<double2*>(ctx->ws_P.ptr);
    double2* S = reinterpret_cast<double2*>(ctx->ws_S.ptr);
    std::vector<ExchangeMsg> msgs;
    for (int r = 0; r < cm.nranks; ++r) {
        int xa, xb, ya, yb;
        part_range(nmx, cm.nranks, r, xa, xb);
        part_range(nse->Ny, cm.nranks, r, ya, yb);
        for (int k = 0; k < nsel; ++k) {
            const int f = fields[k];
            double2* pp = P + ((size_t)f * nse->Ny + ya) * nxl * nkz;                                     // my rows, r's planes
            const long long pb = (long long)(yb - ya) * nxl * nkz * 16;
            double2* sp = S + (size_t)nkz * ((size_t)nf * nyl * xa + (size_t)f * nyl * (xb - xa));          // r's rows, my planes
            const long long sb = (long long)nyl * (xb - xa) * nkz * 16;
            if (dir == 0) msgs.push_back({r, pp, pb, sp, sb});
            else msgs.push_back({r, sp, sb, pp, pb});
        }
    }
    return comm_exchange(cm, msgs.data(), (int)msgs.size(), stream);
}

// How the all-to-all travels when the ranks can map each other's buffers (CFGPU_PEER_MODE):
//   push  the transform kernels write locally at full speed; per velocity component a push kernel on a
//         high-priority stream copies the blocks into the receivers' buffers while the next component is being transformed,
//         completion is signalled through device-side flags (comm.cuh) -- no NCCL call, no barrier
//   fused (default: measured faster at 2 and 8 GPUs, profiles/r02_*) the y-GEMM epilogue / forward x-pass store every row
//         straight into the owning rank's buffer; completion travels through the same device-side flags
enum { PEER_PUSH = 0, PEER_FUSED = 1 };
static int peer_mode() {  // read per call (like CFGPU_NO_PEER): all ranks must of course agree
    const char* e = getenv("CFGPU_PEER_MODE");
    return (e && !strcmp(e, "push")) ? PEER_PUSH : PEER_FUSED;
}
static unsigned long long* flag_words(void* base) { return reinterpret_cast<unsigned long long*>(base); }
static unsigned int* flag_counter(void* base) { return reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(base) + 2048); }
static int* flag_error(void* base) { return reinterpret_cast<int*>(reinterpret_cast<char*>(base) + 2048 + 64); }

// push the blocks of the selected fields to every rank (same geometry as slab_exchange) and publish completion in `slot`
static int slab_push(cfgpu_nse nse, int nf, int dir, const int* fields, int nsel, int slot, cudaStream_t stream) {
    cfgpu_ctx ctx = nse->ctx;
    Comm& cm = ctx->comm;
    const int nmx = 2 * nse->Kx + 1, nkz = nse->Kz + 1;
    const int nxl = nse->x1 - nse->x0, nyl = nse->y1 - nse->y0;
    PushParams pp;
    memset(&pp, 0, sizeof pp);
    const double2* P = reinterpret_cast<const double2*>(ctx->ws_P.ptr);
    const double2* S = reinterpret_cast<const double2*>(ctx->ws_S.ptr);
    for (int r = 0; r < cm.nranks; ++r) {
        int xa, xb, ya, yb;
        part_range(nmx, cm.nranks, r, xa, xb);
        part_range(nse->Ny, cm.nranks, r, ya, yb);
        if (r == cm.rank) continue;  // the producers write this rank's own block in place (row tables / self-direct rows)
        for (int k = 0; k < nsel; ++k) {
            const int f = fields[k];
            PushMsg& m = pp.msg[pp.nmsg++];
            if (dir == 0) {  // my rows of r's planes: P -> r's staging S_r[me][f]
                m.src = P + ((size_t)f * nse->Ny + ya) * nxl * nkz;
                m.dst = reinterpret_cast<double2*>(ctx->peerS[r]) + (size_t)nkz * ((size_t)nf * (yb - ya) * nse->x0 + (size_t)f * (yb - ya) * nxl);
                m.n = (long)(yb - ya) * nxl * nkz;
            } else {         // r's rows of my planes: S[r][f] -> r's pencils P_r[f][my planes]
                m.src = S + (size_t)nkz * ((size_t)nf * nyl * xa + (size_t)f * nyl * (xb - xa));
                m.dst = reinterpret_cast<double2*>(ctx->peerP[r]) + ((size_t)f * nse->Ny + nse->y0) * (xb - xa) * nkz;
                m.n = (long)nyl * (xb - xa) * nkz;
            }
        }
    }
    for (int r = 0; r < cm.nranks; ++r) pp.flags[r] = flag_words(ctx->peerF[r]);
    pp.nranks = cm.nranks; pp.rank = cm.rank; pp.slot = slot;
    pp.seq = ++ctx->push_seq[slot];
    pp.done_counter = flag_counter(ctx->ws_F.ptr);
    static const int nctas = getenv("CF_PUSH_CTAS") ? atoi(getenv("CF_PUSH_CTAS")) : 24;
    return slab_push_launch(pp, nctas, stream);
}
// fused mode: this rank's producer kernel (whose stores went straight into the peers' buffers) precedes this call in
// stream order; publish its completion to every rank
static int slab_signal(cfgpu_nse nse, int slot, cudaStream_t stream) {
    cfgpu_ctx ctx = nse->ctx;
    PushParams pp;
    memset(&pp, 0, sizeof pp);
    for (int r = 0; r < ctx->comm.nranks; ++r) pp.flags[r] = flag_words(ctx->peerF[r]);
    pp.nranks = ctx->comm.nranks; pp.rank = ctx->comm.rank; pp.slot = slot;
    pp.seq = ++ctx->push_seq[slot];
    pp.done_counter = flag_counter(ctx->ws_F.ptr);
    return slab_push_launch(pp, 0, stream);  // no messages: just the flag kernel
}
// fused exchange pipelined per velocity component over two streams (CFGPU_PIPELINE=1).  Off by default: measured slower
// than the single-stream form at 2 and 8 GPUs (4.95 vs 4.63 ms and 1.75 vs 1.60 ms per step at 512x257x512,
// profiles/r02i_*, r02j_*): three launches per transform quantise worse and the co-scheduled kernels share the SMs'
// store path to the peers, which is what the producers wait on.
static bool fused_pipelined() { return getenv("CFGPU_PIPELINE") && atoi(getenv("CFGPU_PIPELINE")) == 1; }
static int slab_wait(cfgpu_nse nse, int slot, cudaStream_t stream) {
    cfgpu_ctx ctx = nse->ctx;
    return slab_wait_launch(flag_words(ctx->ws_F.ptr), slot, ctx->comm.nranks, ctx->push_seq[slot], flag_error(ctx->ws_F.ptr), stream);
}

// Peer-memory path (NCCL backend): map every rank's pencil (P) and staging (S) buffers, and build the tables that send
// each output row of the inverse y-GEMM to the rank owning that y plane.
// A peer-mapped buffer is only ever replaced collectively: every rank has finished with the old one (barrier), all
// mappings of it are closed (second barrier), then it is freed, re-allocated and mapped again.  The decision is the same
// on every rank because the sizes depend on the geometry only.
static int peer_workspace(cfgpu_ctx ctx, Workspace& w, size_t bytes, void** peers, unsigned long long& mapped_gen) {
    Comm& cm = ctx->comm;
    if (bytes > w.bytes) {
        if (mapped_gen) {
            CF_TRY(comm_barrier(cm, ctx->stream));
            CF_CUDA(cudaStreamSynchronize(ctx->stream));
            CF_CUDA(cudaStreamSynchronize(ctx->comm_stream));
            CF_TRY(comm_close_peers(cm, peers));
            mapped_gen = 0;
            CF_TRY(comm_barrier(cm, ctx->stream));
            CF_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        w.exported = false;
        CF_TRY(ws_reserve(w, bytes));
    }
    if (mapped_gen != w.gen) {
        if (mapped_gen) CF_TRY(comm_close_peers(cm, peers));
        mapped_gen = 0;
        CF_TRY(comm_open_peers(cm, w.ptr, peers, ctx->stream));
        if (cm.peer_failed) return 0;
        mapped_gen = w.gen;
        w.exported = true;
    }
    return 0;
}
static int ensure_peers(cfgpu_nse nse, size_t Pbytes, size_t Sbytes) {
    cfgpu_ctx ctx = nse->ctx;
    Comm& cm = ctx->comm;
    if (!ctx->peerF_gen) {
        CF_TRY(peer_workspace(ctx, ctx->ws_F, 4096, ctx->peerF, ctx->peerF_gen));
        if (cm.peer_failed) return 0;
        CF_CUDA(cudaMemsetAsync(ctx->ws_F.ptr, 0, 4096, ctx->stream));
        for (int k = 0; k < PUSH_SLOTS; ++k) ctx->push_seq[k] = 0;
        CF_TRY(comm_barrier(cm, ctx->stream));  // nobody publishes a flag before everybody has cleared its buffer
        CF_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    CF_TRY(peer_workspace(ctx, ctx->ws_P, Pbytes, ctx->peerP, ctx->peerP_gen));
    if (cm.peer_failed) return 0;
    CF_TRY(peer_workspace(ctx, ctx->ws_S, Sbytes, ctx->peerS, ctx->peerS_gen));
    if (cm.peer_failed) {
        comm_close_peers(cm, ctx->peerP);
        ctx->peerP_gen = 0;
        ctx->ws_P.exported = false;
        return 0;
    }
    if (nse->rows_genS != ctx->ws_S.gen || nse->rows_genP != ctx->ws_P.gen || nse->rows_nranks != cm.nranks) {
        const int nkz = nse->Kz + 1, nxl = nse->x1 - nse->x0;
        for (int v = 0; v < 2; ++v) {
            const int nf = v == 0 ? 5 : 3;
            std::vector<double*> tab((size_t)nf * nse->Ny);
            for (int r = 0; r < cm.nranks; ++r) {
                int ya, yb;
                part_range(nse->Ny, cm.nranks, r, ya, yb);
                const int nyl = yb - ya;
                double* Sr = reinterpret_cast<double*>(ctx->peerS[r]);
                for (int f = 0; f < nf; ++f)
                    for (int y = ya; y < yb; ++y)
                        tab[(size_t)f * nse->Ny + y] = Sr + 2 * (size_t)nkz * ((size_t)nf * nyl * nse->x0 + ((size_t)f * nyl + (y - ya)) * nxl);
            }
            if (!nse->d_rows[v]) CF_CUDA(cudaMalloc((void**)&nse->d_rows[v], tab.size() * sizeof(double*)));
            CF_CUDA(cudaMemcpyAsync(nse->d_rows[v], tab.data(), tab.size() * sizeof(double*), cudaMemcpyHostToDevice, ctx->stream));
            CF_CUDA(cudaStreamSynchronize(ctx->stream));
            // push mode: only the rows of this rank's own planes go straight to (its own) staging, the others stay in the
            // local pencil buffer P[f][Ny][nxl][nkz] from where the push kernel sends them
            const size_t Pf = (size_t)nse->Ny * nxl * nkz * 2;
            for (int f = 0; f < nf; ++f)
                for (int y = 0; y < nse->Ny; ++y)
                    if (y < nse->y0 || y >= nse->y1) tab[(size_t)f * nse->Ny + y] = ctx->ws_P.ptr + (size_t)f * Pf + (size_t)y * nxl * nkz * 2;
            if (!nse->d_rows_self[v]) CF_CUDA(cudaMalloc((void**)&nse->d_rows_self[v], tab.size() * sizeof(double*)));
            CF_CUDA(cudaMemcpyAsync(nse->d_rows_self[v], tab.data(), tab.size() * sizeof(double*), cudaMemcpyHostToDevice, ctx->stream));
            CF_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        nse->rows_genS = ctx->ws_S.gen;
        nse->rows_genP = ctx->ws_P.gen;
        nse->rows_nranks = cm.nranks;
    }
    return 0;
}

static void fill_xsplit(cfgpu_nse nse, XPassParams& xp, int nstage) {
    const Comm& cm = nse->ctx->comm;
    xp.nranks = cm.nranks;
    xp.nstage = nstage;
    for (int r = 0; r < cm.nranks; ++r) {
        int a, b;
        part_range(2 * nse->Kx + 1, cm.nranks, r, a, b);
        xp.xsplit[r] = a;
        xp.xsplit[r + 1] = b;
    }
}

// Spectral side of a y-GEMM on a 3-component field in either layout: columns are the rank's retained modes in natural
// order; the serial layout has one run of kz per kx row, the tile-major layout one run of TM modes per tile.
struct SpecAddr { int runlen; const long* runstart; long ld; long compstride; double* base; };
static SpecAddr spec_addr(cfgpu_nse nse, cfgpu_field f, const ModeBox* bx) {
    if (f->layout == 1)
        return {2 * f->tg.TM, nse->d_tilestart, 2L * f->tg.TM, (long)f->tile_compstride(), f->dtile};
    return {2 * (nse->Kz + 1), bx->runstart_full + nse->x0, (long)f->rowstride(), (long)f->compstride(), f->dser};
}
// input of the hot path: tile-major data of another geometry goes back to the serial layout first
static int spec_input(cfgpu_nse nse, cfgpu_field u) {
    if (u->layout == 1 && !u->tg.same(nse->tg)) return field_serial(u);
    if (u->layout == 0) return field_ser_alloc(u);
    return 0;
}
// output of the hot path (all retained modes are written, everything else is zero): tile-major when enabled
static int spec_output(cfgpu_nse nse, cfgpu_field f) {
    if (nse->use_tile) return field_tile_output(f, nse->tg, true);
    CF_TRY(field_serial_output(f));
    if (!(f->clean_Kx >= 0 && f->clean_Kx <= nse->Kx && f->clean_Kz >= 0 && f->clean_Kz <= nse->Kz))
        CF_CUDA(cudaMemsetAsync(f->dser, 0, f->n * sizeof(double), nse->ctx->stream));
    return 0;
}

// spectral u (3 comps, reference layout; this rank's kx rows) -> compact pencils P[0..nout) (y physical), all-to-all,
// then Q (x physical) for this rank's y planes
static int inverse_to_Q(cfgpu_nse nse, cfgpu_field u, bool with_derivs) {
    cfgpu_ctx ctx = nse->ctx;
    const YPlan* yp;
    const ModeBox* bx;
    const FftPlanDev* fx;
    CF_TRY(get_yplan(ctx, nse->Ny, nse->a, nse->b, &yp));
    CF_TRY(get_box(ctx, nse->Nx, nse->Nz, nse->Kx, nse->Kz, &bx));
    CF_TRY(get_fftplan(ctx, nse->Nx, &fx));
    const int nmx = 2 * nse->Kx + 1, nkz = nse->Kz + 1;
    const int nxl = nse->x1 - nse->x0, nyl = nse->y1 - nse->y0;
    const bool multi = ctx->comm.nranks > 1;
    const size_t Pf = (size_t)nse->Ny * nxl * nkz * 2;     // doubles per pencil field P (kx slab)
    const size_t Sf = (size_t)nyl * nmx * nkz * 2;          // doubles per staged field (y slab)
    const size_t Qf = (size_t)nyl * nse->Nx * nkz * 2;      // doubles per pencil field Q (y slab)
    bool peer = multi && comm_peer_capable(ctx->comm);
    if (peer) {
        CF_TRY(ensure_peers(nse, 5 * Pf * sizeof(double), 5 * Sf * sizeof(double)));
        peer = comm_peer_capable(ctx->comm);  // mapping may have failed (collectively): staged exchange instead
    }
    CF_TRY(ws_reserve(ctx->ws_P, 5 * Pf * sizeof(double)));
    if (multi) CF_TRY(ws_reserve(ctx->ws_S, 5 * Sf * sizeof(double)));
    CF_TRY(ws_reserve(ctx->ws_Q, 6 * Qf * sizeof(double)));
    double* P = ctx->ws_P.ptr;
    const int nfP = with_derivs ? 5 : 3;

    YGemmParams p;
    memset(&p, 0, sizeof p);
    p.N = nse->Ny; p.mode = 0;
    p.fft = yfft_plan(ctx, nse->Ny); p.fft_half = yfft_half_plan(ctx, nse->Ny); p.ya = nse->a; p.yb = nse->b;
    p.M = p.M2 = yp->Nh; p.K1 = yp->Ne; p.K2 = yp->No; p.K1p = yp->invK1p; p.K2p = yp->invK2p;
    p.A1[0] = yp->Ce; p.A2[0] = yp->Co; p.sgn[0] = 1.0;
    p.A1[1] = yp->CDe; p.A2[1] = yp->CDo; p.sgn[1] = -1.0;
    p.ncols = (long)nxl * nkz * 2;
    CF_TRY(spec_input(nse, u));
    const SpecAddr ua = spec_addr(nse, u, bx);
    p.in_runlen = ua.runlen; p.in_runstart = ua.runstart; p.in_ld = ua.ld;
    p.out_runlen = 1; p.out_runstart = nullptr; p.out_ld = (long)nxl * nkz * 2;
    p.njobs = 3;
    for (int i = 0; i < 3; ++i) {
        p.job[i].in = ua.base + i * ua.compstride;
        p.job[i].out[0] = P + i * Pf;
        p.job[i].nmat = 1; p.job[i].mat0 = 0;
    }
    if (with_derivs) {
        p.job[0].nmat = 2; p.job[0].out[1] = P + 3 * Pf;  // du/dy
        p.job[2].nmat = 2; p.job[2].out[1] = P + 4 * Pf;  // dw/dy
    }
    const bool fused = peer && peer_mode() == PEER_FUSED;
    if (fused) {
        // every output row goes straight to the staging buffer of the rank that owns its y plane (stores over NVLink);
        // barriers: the consumers are done with the previous contents / all rows have arrived
        double** tab = nse->d_rows[with_derivs ? 0 : 1];
        for (int i = 0; i < 3; ++i) {
            p.job[i].out_rows[0] = tab + (size_t)i * nse->Ny;
            if (p.job[i].nmat == 2) p.job[i].out_rows[1] = tab + (size_t)(i == 0 ? 3 : 4) * nse->Ny;
        }
        // No barrier in front: a rank reaches this point only after it has received the previous exchange from everybody,
        // and everybody sent it only after having consumed what these stores overwrite (see comm.cuh).  Completion travels
        // through device-side flags (CFGPU_FLAG_BARRIER=0: the round-1 NCCL barriers).
        const bool flagbar = !(getenv("CFGPU_FLAG_BARRIER") && atoi(getenv("CFGPU_FLAG_BARRIER")) == 0);
        if (fused_pipelined()) {
            // per velocity component: the y-GEMM (NVLink-bound through its remote stores) on the compute stream, its
            // completion flag; the x-passes of the components that have arrived run beside it on the second stream
            for (int c = 0; c < 3; ++c) {
                YGemmParams pc = p;
                pc.njobs = 1;
                pc.job[0] = p.job[c];
                { StageTimer _t(ctx, 0); CF_TRY(ygemm_launch(pc, ctx->stream)); }
                CF_TRY(slab_signal(nse, c, ctx->stream));
            }
        } else {
            if (!flagbar) CF_TRY(comm_barrier(ctx->comm, ctx->stream));
            { StageTimer _t(ctx, 0); CF_TRY(ygemm_launch(p, ctx->stream)); }
            {
                StageTimer _t(ctx, 8);
                if (flagbar) { CF_TRY(slab_signal(nse, 6, ctx->stream)); CF_TRY(slab_wait(nse, 6, ctx->stream)); }
                else CF_TRY(comm_barrier(ctx->comm, ctx->stream));
            }
        }
    } else if (!multi) {
        StageTimer _t(ctx, 0);
        CF_TRY(ygemm_launch(p, ctx->stream));
    } else {
        // per velocity component: y-GEMM on the compute stream, its all-to-all on the communication stream while the
        // next component's y-GEMM runs
        for (int c = 0; c < 3; ++c) {
            YGemmParams pc = p;
            pc.njobs = 1;
            pc.job[0] = p.job[c];
            if (peer) {
                double** tab = nse->d_rows_self[with_derivs ? 0 : 1];
                pc.job[0].out_rows[0] = tab + (size_t)c * nse->Ny;
                if (pc.job[0].nmat == 2) pc.job[0].out_rows[1] = tab + (size_t)(c == 0 ? 3 : 4) * nse->Ny;
            }
            { StageTimer _t(ctx, 0); CF_TRY(ygemm_launch(pc, ctx->stream)); }
            CF_CUDA(cudaEventRecord(ctx->ev_cmp[c], ctx->stream));
            CF_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_cmp[c], 0));
            const int fl[2] = {c, c == 0 ? 3 : 4};
            if (peer) CF_TRY(slab_push(nse, nfP, 0, fl, (with_derivs && c != 1) ? 2 : 1, c, ctx->comm_stream));
            else {
                CF_TRY(slab_exchange(nse, nfP, 0, fl, (with_derivs && c != 1) ? 2 : 1, ctx->comm_stream));
                CF_CUDA(cudaEventRecord(ctx->ev_com[c], ctx->comm_stream));
            }
        }
    }

    XPassParams xp;
    memset(&xp, 0, sizeof xp);
    xp.Nx = nse->Nx; xp.Ny = nse->Ny; xp.Kx = nse->Kx; xp.Kz = nse->Kz;
    xp.TZ = pick_TZ(nse->Nx);
    xp.Lx = nse->Lx;
    xp.plan = *fx;
    xp.in = reinterpret_cast<const double2*>(multi ? ctx->ws_S.ptr : P);
    xp.out = reinterpret_cast<double2*>(ctx->ws_Q.ptr);
    xp.ny0 = nse->y0; xp.nyn = nyl;
    fill_xsplit(nse, xp, nfP);
    xp.Lz = nse->Lz;
    if (with_derivs) {
        // P fields: 0 u, 1 v, 2 w, 3 du/dy, 4 dw/dy.  Q fields: u, v, w, omega_x = dw/dy - dv/dz,
        // omega_y = du/dz - dw/dx, omega_z = dv/dx - du/dy  (curl, diffops.cpp:2229-2334)
        const int src[6] = {0, 1, 2, 4, 0, 1}, opa[6] = {0, 0, 0, 0, 2, 1};
        const int srcb[6] = {-1, -1, -1, 1, 2, 3}, opb[6] = {0, 0, 0, 2, 1, 0};
        xp.nfields = 6;
        for (int i = 0; i < 6; ++i) { xp.src[i] = src[i]; xp.opa[i] = opa[i]; xp.srcb[i] = srcb[i]; xp.opb[i] = opb[i]; xp.fsel[i] = i; }
    } else {
        xp.nfields = 3;
        for (int i = 0; i < 3; ++i) { xp.src[i] = i; xp.opa[i] = 0; xp.srcb[i] = -1; xp.opb[i] = 0; xp.fsel[i] = i; }
    }
    if (fused && fused_pipelined()) {
        const int sel[3][3] = {{0, -1, -1}, {1, 5, -1}, {2, 3, 4}};
        cudaStream_t X = ctx->comm_stream;
        for (int c = 0; c < 3; ++c) {
            { StageTimer _t(ctx, 8, X); CF_TRY(slab_wait(nse, c, X)); }
            XPassParams xc = xp;
            xc.nfields = 0;
            if (with_derivs) { for (int k = 0; k < 3; ++k) if (sel[c][k] >= 0) xc.fsel[xc.nfields++] = sel[c][k]; }
            else xc.fsel[xc.nfields++] = c;
            StageTimer _t(ctx, 1, X);
            CF_TRY(xpass_inverse_launch(xc, X));
        }
        CF_CUDA(cudaEventRecord(ctx->ev_com[3], X));
        CF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_com[3], 0));
    } else if (!multi || fused) {
        StageTimer _t(ctx, 1);
        CF_TRY(xpass_inverse_launch(xp, ctx->stream));
    } else {
        // output slots that become computable when component c has arrived: u (+du/dy) -> {u}; v -> {v, omega_z};
        // w (+dw/dy) -> {w, omega_x, omega_y}
        const int sel[3][3] = {{0, -1, -1}, {1, 5, -1}, {2, 3, 4}};
        for (int c = 0; c < 3; ++c) {
            if (peer) { StageTimer _t(ctx, 8); CF_TRY(slab_wait(nse, c, ctx->stream)); }
            else CF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_com[c], 0));
            XPassParams xc = xp;
            xc.nfields = 0;
            if (with_derivs) { for (int k = 0; k < 3; ++k) if (sel[c][k] >= 0) xc.fsel[xc.nfields++] = sel[c][k]; }
            else xc.fsel[xc.nfields++] = c;
            StageTimer _t(ctx, 1);
            CF_TRY(xpass_inverse_launch(xc, ctx->stream));
        }
    }
    return 0;
}

static int fill_zpass(cfgpu_nse nse, ZPassParams& zp, int mode) {
    const FftPlanDev* fz;
    CF_TRY(get_fftplan(nse->ctx, nse->Nz, &fz));
    memset(&zp, 0, sizeof zp);
    zp.Nx = nse->Nx; zp.Ny = nse->Ny; zp.Nz = nse->Nz; zp.Kz = nse->Kz;
    zp.mode = mode;
    zp.TL = pick_TL(nse->Nx, nse->Nz, mode == ZP_ROTATIONAL ? 3 : 2);
    zp.Lx = nse->Lx; zp.Lz = nse->Lz;
    zp.scale = 1.0 / ((double)nse->Nx * (double)nse->Nz);
    zp.Vsuck = nse->cfg.Vsuck;
    zp.rotation = nse->cfg.rotation;
    zp.plan = *fz;
    zp.Q = reinterpret_cast<const double2*>(nse->ctx->ws_Q.ptr);
    zp.F = reinterpret_cast<double2*>(nse->ctx->ws_Q.ptr);  // in place: a CTA overwrites only lines it has consumed
    zp.Uy = nse->d_base + 2 * nse->Ny;
    zp.inv_dy = nse->d_base + 6 * nse->Ny;
    zp.cfl_max = nse->d_scal;
    zp.ny0 = nse->y0; zp.nyn = nse->y1 - nse->y0;
    return 0;
}

extern "C" {

int cfgpu_nse_create(cfgpu_ctx ctx, int Nx, int Ny, int Nz, double Lx, double Lz, double a, double b,
                     const cfgpu_nse_config* cfg, const double* Ubase_h, const double* Wbase_h, cfgpu_nse* out) {
    CF_ARG(ctx && cfg && out, "cfgpu_nse_create: null argument");
    CF_ARG(Ny % 2 == 1 && Ny >= 5, "cfgpu_nse_create: Ny must be odd and >= 5 (helmholtz.cpp:31)");
    CF_ARG(cfg->nonlinearity >= 0 && cfg->nonlinearity <= 6, "cfgpu_nse_create: unknown nonlinearity (LinearAboutField is not supported)");
    cfgpu_nse nse = new cfgpu_nse_s();
    nse->ctx = ctx;
    nse->Nx = Nx; nse->Ny = Ny; nse->Nz = Nz; nse->Lx = Lx; nse->Lz = Lz; nse->a = a; nse->b = b;
    nse->cfg = *cfg;
    nse->Nyd = cfg->dealias_y ? 2 * (Ny - 1) / 3 + 1 : Ny;              // nse.cpp:228
    nse->Kx = cfg->dealias_xz ? Nx / 3 - 1 : Nx / 2 - 1;                // nse.cpp:229-230, 498 (kxmax mode is skipped)
    nse->Kz = cfg->dealias_xz ? Nz / 3 - 1 : Nz / 2 - 1;
    // (odd Nx without de-aliasing: the reference skips kx == kxmax only and still advances kx = -(Nx-1)/2; here both rows
    // lie outside the symmetric box and are left untouched -- documented difference)
    CF_ARG(ctx->comm.nranks == 1 || cfg->dealias_xz,
           "cfgpu_nse_create: multi-GPU runs need 2/3 de-aliasing in x,z (the kx-slab partition, the gathers and the padded "
           "transfers are defined on the de-aliased box)");
    CF_ARG(nse->Kx >= 0 && nse->Kz >= 0, "cfgpu_nse_create: grid too small");
    CF_ARG(nse->Nyd % 2 == 1, "cfgpu_nse_create: dealiased Ny must be odd");
    part_range(2 * nse->Kx + 1, ctx->comm.nranks, ctx->comm.rank, nse->x0, nse->x1);
    part_range(Ny, ctx->comm.nranks, ctx->comm.rank, nse->y0, nse->y1);
    CF_ARG(nse->x1 > nse->x0 && nse->y1 > nse->y0, "cfgpu_nse_create: more ranks than kx rows or y planes");
    nse->nq = (nse->x1 - nse->x0) * (nse->Kz + 1);
    nse->geom.mx0 = nse->x0;
    nse->geom.Nx = Nx; nse->geom.Ny = Ny; nse->geom.Nz = Nz; nse->geom.Kx = nse->Kx; nse->geom.Kz = nse->Kz;
    nse->geom.Lx = Lx; nse->geom.Lz = Lz;
    nse->TM_solve = tau_pick_TM_solve(nse->Nyd);
    nse->TM_lin = tau_pick_TM(nse->Nyd, 5 * 2 * 8);
    // tile-major layout of the hot-path fields: de-aliased runs only (the retained box is then all a field holds)
    nse->tg.TM = nse->TM_solve; nse->tg.Kx = nse->Kx; nse->tg.Kz = nse->Kz; nse->tg.x0 = nse->x0; nse->tg.x1 = nse->x1;
    nse->use_tile = cfg->dealias_xz && nse->Nyd == Ny && !getenv("CFGPU_SERIAL_LAYOUT");
    if (nse->use_tile) {
        std::vector<long> ts((size_t)nse->tg.ntiles());
        for (size_t t = 0; t < ts.size(); ++t) ts[t] = (long)t * 3 * Ny * nse->tg.TM * 2;
        CF_CUDA(cudaMalloc((void**)&nse->d_tilestart, ts.size() * sizeof(long)));
        CF_CUDA(cudaMemcpy(nse->d_tilestart, ts.data(), ts.size() * sizeof(long), cudaMemcpyHostToDevice));
    }

    // base-flow data: Ubaseyy, Wbaseyy (spectral), physical U,U',W,W', 1/dy
    std::vector<double> U(Ny, 0.0), W(Ny, 0.0), Uy, Uyy, Wy, Wyy, t;
    if (Ubase_h) U.assign(Ubase_h, Ubase_h + Ny);
    if (Wbase_h) W.assign(Wbase_h, Wbase_h + Ny);
    cheb_diff_host(U, Uy, a, b); cheb_diff_host(Uy, Uyy, a, b);
    cheb_diff_host(W, Wy, a, b); cheb_diff_host(Wy, Wyy, a, b);
    {
        double ub = 0, ua = 0, wb = 0, wa = 0;
        for (int n = Ny - 1; n >= 0; --n) {
            ub += Uy[n]; ua += Uy[n] * ((n % 2 == 0) ? 1 : -1);
            wb += Wy[n]; wa += Wy[n] * ((n % 2 == 0) ? 1 : -1);
        }
        nse->lin_base_dPdx = Ubase_h ? cfg->nu * (ub - ua) / (b - a) : 0.0;
        nse->lin_base_dPdz = Wbase_h ? cfg->nu * (wb - wa) / (b - a) : 0.0;
    }
    nse->has_Ubaseyy = Ubase_h != nullptr;
    nse->has_Wbaseyy = Wbase_h != nullptr;
    std::vector<double> hb(9 * (size_t)Ny, 0.0);
    for (int n = 0; n < Ny; ++n) { hb[n] = Uyy[n]; hb[Ny + n] = Wyy[n]; hb[7 * Ny + n] = U[n]; hb[8 * Ny + n] = W[n]; }
    cheb_to_physical_host(U, t);  for (int n = 0; n < Ny; ++n) hb[2 * Ny + n] = t[n];
    cheb_to_physical_host(Uy, t); for (int n = 0; n < Ny; ++n) hb[3 * Ny + n] = t[n];
    cheb_to_physical_host(W, t);  for (int n = 0; n < Ny; ++n) hb[4 * Ny + n] = t[n];
    cheb_to_physical_host(Wy, t); for (int n = 0; n < Ny; ++n) hb[5 * Ny + n] = t[n];
    {
        const double pi = 3.14159265358979323846;
        std::vector<double> y(Ny);
        for (int n = 0; n < Ny; ++n) y[n] = 0.5 * ((b + a) + (b - a) * cos(pi * n / (Ny - 1)));  // flowfield.h:483
        for (int n = 0; n < Ny; ++n) {
            const double dy = (n == 0 || n == Ny - 1) ? y[0] - y[1] : (y[n - 1] - y[n + 1]) / 2.0;  // flowfield.cpp:4051
            hb[6 * Ny + n] = 1.0 / dy;
        }
    }
    if (cudaMalloc((void**)&nse->d_base, hb.size() * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&nse->d_scal, 8 * sizeof(double)) != cudaSuccess) {
        set_last_error("cfgpu_nse_create: cudaMalloc failed");
        delete nse;
        return 1;
    }
    CF_CUDA(cudaMemcpy(nse->d_base, hb.data(), hb.size() * sizeof(double), cudaMemcpyHostToDevice));
    CF_CUDA(cudaMemset(nse->d_scal, 0, 8 * sizeof(double)));
    *out = nse;
    return 0;
}

int cfgpu_nse_destroy(cfgpu_nse nse) {
    if (!nse) return 0;
    cudaStreamSynchronize(nse->ctx->stream);
    for (auto& t : nse->tau) dev_free(nse->ctx, t.base);
    for (int v = 0; v < 2; ++v) if (nse->d_rows[v]) cudaFree(nse->d_rows[v]);
    for (int v = 0; v < 2; ++v) if (nse->d_rows_self[v]) cudaFree(nse->d_rows_self[v]);
    if (nse->s_u) cfgpu_field_destroy(nse->s_u);
    if (nse->s_t) cfgpu_field_destroy(nse->s_t);
    cudaFree(nse->d_base);
    cudaFree(nse->d_scal);
    if (nse->d_tilestart) cudaFree(nse->d_tilestart);
    delete nse;
    return 0;
}

int cfgpu_nse_set_constraint(cfgpu_nse nse, int constraint, double dPdxRef, double dPdzRef, double UbR, double WbR) {
    nse->cfg.constraint = constraint;
    nse->cfg.dPdxRef = dPdxRef; nse->cfg.dPdzRef = dPdzRef;
    nse->cfg.UbulkRef_minus_base = UbR; nse->cfg.WbulkRef_minus_base = WbR;
    return 0;
}

int cfgpu_nse_reset_lambda(cfgpu_nse nse, const double* lambda_t_h, int nsub) {
    CF_ARG(nsub > 0, "cfgpu_nse_reset_lambda: nsub");
    cfgpu_ctx ctx = nse->ctx;
    while ((int)nse->tau.size() < nsub) {
        TauData td;
        td.N = nse->Nyd; td.nq = nse->nq; td.has00 = nse->x0 == 0 ? 1 : 0; td.TM = nse->TM_solve;
        td.ntiles = TauData::num_tiles(td.nq, td.TM, td.has00);
        td.nu = nse->cfg.nu; td.a = nse->a; td.b = nse->b;
        const size_t n = TauData::doubles(td.N, td.nq, td.TM, td.has00);
        if (dev_alloc(ctx, (void**)&td.base, n * sizeof(double))) {
            set_last_error("cfgpu_nse_reset_lambda: cudaMalloc failed");
            return 1;
        }
        CF_CUDA(cudaMemsetAsync(td.base, 0, n * sizeof(double), ctx->stream));
        std::vector<double> bt(3 * (size_t)td.N);
        tau_btab_host(td.N, bt.data());
        CF_CUDA(cudaMemcpyAsync(td.btab(), bt.data(), bt.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CF_CUDA(cudaStreamSynchronize(ctx->stream));
        nse->tau.push_back(td);
    }
    nse->lambda_t.assign(lambda_t_h, lambda_t_h + nsub);
    for (int s = 0; s < nsub; ++s) {
        nse->tau[s].nu = nse->cfg.nu;
        { StageTimer _t(ctx, 7); CF_TRY(tau_setup_launch(nse->tau[s], nse->geom, lambda_t_h[s], ctx->stream)); }
    }
    return 0;
}

// Convection / divergence / skew-symmetric forms through the compact-pencil pipeline (single GPU, power-of-two Nz):
//   y-GEMM (u and du/dy) -> x-pass (+ d/dx, d/dz factors) -> z-pass (c2r, u.grad u and/or the products u_i u_j, r2c,
//   d/dz of u_i w) -> x-pass (+ d/dx of u_i u) -> y-GEMM (+ d/dy of u_i v through the derivative matrices GD).
// method: 1 convection (dotgrad, diffops.cpp:3586-3643), 2 divergence (:3110-3134), 3 skew-symmetric (:3142-3286)
static int nonlinear_fused_general(cfgpu_nse nse, cfgpu_field u, cfgpu_field f, int method) {
    cfgpu_ctx ctx = nse->ctx;
    const YPlan* yp;
    const ModeBox* bx;
    const FftPlanDev *fx, *fz;
    CF_TRY(get_yplan(ctx, nse->Ny, nse->a, nse->b, &yp));
    CF_TRY(get_box(ctx, nse->Nx, nse->Nz, nse->Kx, nse->Kz, &bx));
    CF_TRY(get_fftplan(ctx, nse->Nx, &fx));
    CF_TRY(get_fftplan(ctx, nse->Nz, &fz));
    const bool grad = method != 2, prod = method != 1;
    const double cc = method == 1 ? 1.0 : (method == 3 ? 0.5 : 0.0), cd = method == 2 ? 1.0 : (method == 3 ? 0.5 : 0.0);
    const int nmx = 2 * nse->Kx + 1, nkz = nse->Kz + 1;
    const size_t Pf = (size_t)nse->Ny * nmx * nkz * 2, Qf = (size_t)nse->Ny * nse->Nx * nkz * 2;
    CF_TRY(ws_reserve(ctx->ws_P, 6 * Pf * sizeof(double)));
    CF_TRY(ws_reserve(ctx->ws_Q, 12 * Qf * sizeof(double)));
    double* P = ctx->ws_P.ptr;

    {   // inverse y: u_i and (if needed) du_i/dy at the Gauss-Lobatto points
        YGemmParams p;
        memset(&p, 0, sizeof p);
        p.N = nse->Ny; p.mode = 0;
        p.fft = yfft_plan(ctx, nse->Ny); p.fft_half = yfft_half_plan(ctx, nse->Ny); p.ya = nse->a; p.yb = nse->b;
        p.M = p.M2 = yp->Nh; p.K1 = yp->Ne; p.K2 = yp->No; p.K1p = yp->invK1p; p.K2p = yp->invK2p;
        p.A1[0] = yp->Ce; p.A2[0] = yp->Co; p.sgn[0] = 1.0;
        p.A1[1] = yp->CDe; p.A2[1] = yp->CDo; p.sgn[1] = -1.0;
        p.ncols = (long)nmx * nkz * 2;
        CF_TRY(spec_input(nse, u));
        const SpecAddr ua = spec_addr(nse, u, bx);
        p.in_runlen = ua.runlen; p.in_runstart = ua.runstart; p.in_ld = ua.ld;
        p.out_runlen = 1; p.out_runstart = nullptr; p.out_ld = (long)nmx * nkz * 2;
        p.njobs = 3;
        for (int i = 0; i < 3; ++i) {
            p.job[i].in = ua.base + i * ua.compstride;
            p.job[i].out[0] = P + i * Pf;
            p.job[i].out[1] = P + (3 + i) * Pf;
            p.job[i].nmat = grad ? 2 : 1; p.job[i].mat0 = 0;
        }
        StageTimer _t(ctx, 0);
        CF_TRY(ygemm_launch(p, ctx->stream));
    }
    XPassParams xp;
    memset(&xp, 0, sizeof xp);
    xp.Nx = nse->Nx; xp.Ny = nse->Ny; xp.Kx = nse->Kx; xp.Kz = nse->Kz;
    xp.TZ = pick_TZ(nse->Nx);
    xp.Lx = nse->Lx; xp.Lz = nse->Lz;
    xp.plan = *fx;
    xp.ny0 = 0; xp.nyn = nse->Ny;
    {   // inverse x: Q = u (3), du/dy (3), du/dx (3), du/dz (3)
        xp.in = reinterpret_cast<const double2*>(P);
        xp.out = reinterpret_cast<double2*>(ctx->ws_Q.ptr);
        fill_xsplit(nse, xp, grad ? 6 : 3);
        xp.nfields = grad ? 12 : 3;
        for (int i = 0; i < 12; ++i) {
            xp.fsel[i] = i;
            xp.src[i] = i < 6 ? i : i % 3;
            xp.opa[i] = i < 6 ? 0 : (i < 9 ? 1 : 2);
            xp.srcb[i] = -1; xp.opb[i] = 0;
        }
        StageTimer _t(ctx, 1);
        CF_TRY(xpass_inverse_launch(xp, ctx->stream));
    }
    {
        ZPassParams zp;
        CF_TRY(fill_zpass(nse, zp, method == 1 ? ZP_CONVECTION : (method == 2 ? ZP_DIVERGENCE : ZP_SKEW)));
        zp.cc = cc; zp.cd = cd;
        zp.cfl_max = nullptr;
        StageTimer _t(ctx, 2);
        CF_TRY(zpass_launch(zp, ctx->stream));
    }
    {   // forward x: H_i = FFT(G_i) + cd d/dx FFT(u_i u)  and  C_i = FFT(u_i v)
        xp.in = reinterpret_cast<const double2*>(ctx->ws_Q.ptr);
        xp.out = reinterpret_cast<double2*>(P);
        fill_xsplit(nse, xp, prod ? 6 : 3);
        xp.nfields = prod ? 6 : 3;
        xp.cb = cd;
        const int t1src[3] = {4, 6, 7};  // u v, v v, v w
        for (int i = 0; i < 6; ++i) {
            xp.fsel[i] = i;
            xp.src[i] = i < 3 ? i : t1src[i - 3];
            xp.srcb[i] = (prod && i < 3) ? 3 + i : -1;
            xp.opb[i] = 1;
        }
        xp.TZ = pick_TZ_forward(nse->Nx);
        xp.inplace = forward_inplace(nse->Nx) ? 1 : 0;
        if (prod) xp.TZ = xp.TZ > 1 ? xp.TZ / 2 : 1;
        StageTimer _t(ctx, 3);
        CF_TRY(xpass_forward_launch(xp, ctx->stream));
    }
    CF_TRY(spec_output(nse, f));
    const SpecAddr fa = spec_addr(nse, f, bx);
    {   // forward y: f_i = F H_i + GD C_i
        YGemmParams p;
        memset(&p, 0, sizeof p);
        p.N = nse->Ny; p.mode = 1;
        p.fft = yfft_plan(ctx, nse->Ny); p.fft_half = yfft_half_plan(ctx, nse->Ny); p.ya = nse->a; p.yb = nse->b;
        p.M = yp->Ne; p.M2 = yp->No; p.K1 = p.K2 = yp->Nh; p.K1p = p.K2p = yp->fwdKp;
        p.A1[0] = yp->Fe; p.A2[0] = yp->Fo; p.sgn[0] = 1.0;
        p.A1b = yp->GDe[cd == 0.5 ? 1 : 0]; p.A2b = yp->GDo[cd == 0.5 ? 1 : 0]; p.in2_scale = cd == 0.5 ? 0.5 : 1.0;
        p.ncols = (long)nmx * nkz * 2;
        p.in_runlen = 1; p.in_runstart = nullptr; p.in_ld = (long)nmx * nkz * 2;
        p.out_runlen = fa.runlen; p.out_runstart = fa.runstart; p.out_ld = fa.ld;
        p.njobs = 3;
        for (int i = 0; i < 3; ++i) {
            p.job[i].in = P + i * Pf;
            p.job[i].in2 = prod ? P + (3 + i) * Pf : nullptr;
            p.job[i].out[0] = fa.base + i * fa.compstride;
            p.job[i].nmat = 1; p.job[i].mat0 = 0;
        }
        StageTimer _t(ctx, 4);
        CF_TRY(ygemm_launch(p, ctx->stream));
    }
    f->xzstate = CFGPU_SPECTRAL; f->ystate = CFGPU_SPECTRAL;
    if (f->layout == 0) { f->clean_Kx = nse->Kx; f->clean_Kz = nse->Kz; }  // (the flags describe the serial buffer)
    if (nse->cfg.dealias_xz) f->padded = 1;
    return 0;
}

// Convection / Divergence / SkewSymmetric / Alternating / LinearAboutProfile (nse.cpp:12-91 -> diffops.cpp): the
// reference's own sequence on scratch copies (u itself is never modified), generic full-grid transforms in between.
static int nonlinear_generic(cfgpu_nse nse, cfgpu_field u, cfgpu_field f) {
    cfgpu_ctx ctx = nse->ctx;
    CF_ARG(ctx->comm.nranks == 1, "only the rotational nonlinearity is distributed over several GPUs in this build");
    {
        const int method = nse->cfg.nonlinearity;
        if (method >= 1 && method <= 5 && zpass_general_supported(nse->Nz) && !getenv("CFGPU_UNFUSED_NL")) {
            // fused pipeline, unless the form/rotation combination is one of the reference's quirks reproduced below
            const int m = method == 4 ? 2 : (method == 5 ? 1 : method);
            if (nse->cfg.rotation == 0.0 || m == 3) {
                if (method == 4) nse->cfg.nonlinearity = 5;
                else if (method == 5) nse->cfg.nonlinearity = 4;
                return nonlinear_fused_general(nse, u, f, m);
            }
        }
    }
    CF_TRY(field_serial(u));
    CF_TRY(field_serial_output(f));  // written in the serial layout below
    if (!nse->s_u) CF_TRY(cfgpu_field_create(ctx, nse->Nx, nse->Ny, nse->Nz, 3, nse->Lx, nse->Lz, nse->a, nse->b, &nse->s_u));
    cfgpu_field su = nse->s_u;
    FieldGeom g{nse->Nx, nse->Ny, nse->Nz, nse->Lx, nse->Lz, nse->a, nse->b};
    const long nreal = (long)su->compstride();
    CF_TRY(cfgpu_field_copy(su, u));
    CF_TRY(field_ser_alloc(su));
    int method = nse->cfg.nonlinearity;
    if (method == 6) {  // LinearAboutProfile (diffops.cpp:3288-3365)
        CF_TRY(cfgpu_field_make_physical_y(su));
        CF_TRY(linearized_launch(su->dser, f->dser, nse->d_base + 2 * nse->Ny, g, ctx->stream));
        f->xzstate = CFGPU_SPECTRAL; f->ystate = CFGPU_PHYSICAL; f->clean_Kx = f->clean_Kz = -1;
        CF_TRY(cfgpu_field_make_spectral_y(f));
    } else {
        if (method == 4) { method = 2; nse->cfg.nonlinearity = 5; }       // Alternating: divergence now, convection next
        else if (method == 5) { method = 1; nse->cfg.nonlinearity = 4; }  // Alternating_
        if (!nse->s_t) CF_TRY(cfgpu_field_create(ctx, nse->Nx, nse->Ny, nse->Nz, 9, nse->Lx, nse->Lz, nse->a, nse->b, &nse->s_t));
        cfgpu_field st = nse->s_t;
        CF_TRY(field_serial_output(st));
        const double rot = nse->cfg.rotation;
        // u_tot = u + Ubase e_x + Wbase e_z - Vsuck e_y on the (0,0) mode (nse.cpp:28-36)
        CF_TRY(add_base00_launch(su->dser, nse->has_Ubaseyy ? nse->d_base + 7 * nse->Ny : nullptr,
                                 nse->has_Wbaseyy ? nse->d_base + 8 * nse->Ny : nullptr, nse->cfg.Vsuck, 1.0, g, ctx->stream));
        if (method == 1 || method == 3) {  // grad(u) (diffops.cpp:3169-3173, 3606-3609)
            CF_TRY(grad3_launch(su->dser, st->dser, g, ctx->stream));
            st->xzstate = st->ystate = CFGPU_SPECTRAL; st->clean_Kx = st->clean_Kz = -1;
            CF_TRY(cfgpu_field_make_physical(st));
        }
        CF_TRY(cfgpu_field_make_physical(su));
        if (method == 1) {         // convectionNL = dotgrad(u,u) (diffops.cpp:2883-2885, 3586-3643)
            CF_TRY(pointwise_nl_launch(su->dser, st->dser, f->dser, 1.0, 0, 0.0, nreal, ctx->stream));
            f->xzstate = f->ystate = CFGPU_PHYSICAL; f->clean_Kx = f->clean_Kz = -1;
            CF_TRY(cfgpu_field_make_spectral(f));
            // REFERENCE QUIRK, reproduced: dotgrad ignores `finalstate` and returns f spectral, yet navierstokesNL then adds
            // the Coriolis term of the physical u to f's raw array and its closing f.makeSpectral() is a no-op
            // (nse.cpp:63-79 after diffops.cpp:3640).  Only the rotational and skew-symmetric forms honour finalstate.
            if (rot != 0.0) CF_TRY(coriolis_launch(su->dser, f->dser, rot, nreal, nse->Nz, ctx->stream));
        } else if (method == 3) {  // skewsymmetricNL (diffops.cpp:3142-3286)
            CF_TRY(pointwise_nl_launch(su->dser, st->dser, f->dser, 0.5, 1, rot, nreal, ctx->stream));
            f->xzstate = f->ystate = CFGPU_PHYSICAL; f->clean_Kx = f->clean_Kz = -1;
            st->xzstate = st->ystate = CFGPU_PHYSICAL;
            CF_TRY(cfgpu_field_make_spectral(st));
            CF_TRY(cfgpu_field_make_spectral(f));
            CF_TRY(div9_launch(st->dser, f->dser, 0.5, 1, g, ctx->stream));
        } else {                   // divergenceNL (diffops.cpp:3110-3134)
            CF_TRY(pointwise_nl_launch(su->dser, st->dser, f->dser, 0.0, 1, 0.0, nreal, ctx->stream));
            st->xzstate = st->ystate = CFGPU_PHYSICAL; st->clean_Kx = st->clean_Kz = -1;
            CF_TRY(cfgpu_field_make_spectral(st));
            CF_TRY(div9_launch(st->dser, f->dser, 1.0, 0, g, ctx->stream));
            f->xzstate = f->ystate = CFGPU_SPECTRAL; f->clean_Kx = f->clean_Kz = -1;
            // same reference quirk as for the convection form: div(uu, f, Physical) leaves f spectral (diffops.cpp:2555-2557)
            if (rot != 0.0) CF_TRY(coriolis_launch(su->dser, f->dser, rot, nreal, nse->Nz, ctx->stream));
        }
    }
    f->xzstate = f->ystate = CFGPU_SPECTRAL;
    if (nse->cfg.dealias_xz) CF_TRY(cfgpu_field_zero_padded_modes(f));  // nse.cpp:389-390
    return 0;
}

int cfgpu_nse_nonlinear(cfgpu_nse nse, cfgpu_field u, cfgpu_field f) {
    CF_ARG(nse && shape_ok(nse, u, 3) && shape_ok(nse, f, 3), "cfgpu_nse_nonlinear: u and f must be 3-component fields on the operator's grid");
    CF_ARG(u->xzstate == CFGPU_SPECTRAL && u->ystate == CFGPU_SPECTRAL, "cfgpu_nse_nonlinear: u must be spectral");
    if (nse->cfg.nonlinearity != 0) return nonlinear_generic(nse, u, f);
    cfgpu_ctx ctx = nse->ctx;
    CF_TRY(inverse_to_Q(nse, u, true));

    ZPassParams zp;
    CF_TRY(fill_zpass(nse, zp, ZP_ROTATIONAL));
    CF_CUDA(cudaMemsetAsync(nse->d_scal, 0, sizeof(double), ctx->stream));
    { StageTimer _t(ctx, 2); CF_TRY(zpass_launch(zp, ctx->stream)); }

    const FftPlanDev* fx;
    CF_TRY(get_fftplan(ctx, nse->Nx, &fx));
    XPassParams xp;
    memset(&xp, 0, sizeof xp);
    xp.Nx = nse->Nx; xp.Ny = nse->Ny; xp.Kx = nse->Kx; xp.Kz = nse->Kz;
    xp.TZ = pick_TZ_forward(nse->Nx);
    xp.inplace = forward_inplace(nse->Nx) ? 1 : 0;
    xp.Lx = nse->Lx;
    xp.plan = *fx;
    xp.nfields = 3;
    const bool multi = ctx->comm.nranks > 1;
    xp.in = reinterpret_cast<const double2*>(ctx->ws_Q.ptr);
    xp.out = reinterpret_cast<double2*>(multi ? ctx->ws_S.ptr : ctx->ws_P.ptr);
    xp.ny0 = nse->y0; xp.nyn = nse->y1 - nse->y0;
    fill_xsplit(nse, xp, 3);
    for (int i = 0; i < 3; ++i) { xp.fsel[i] = i; xp.src[i] = i; xp.srcb[i] = -1; }
    const bool peer = multi && comm_peer_capable(ctx->comm);
    const bool fused = peer && peer_mode() == PEER_FUSED;
    bool f_prepared = false;
    if (fused) {
        // each kx row is stored straight into its owner's pencil buffer
        xp.peer_direct = 1;
        for (int r = 0; r < ctx->comm.nranks; ++r) xp.peer_out[r] = reinterpret_cast<double2*>(ctx->peerP[r]);
        if (fused_pipelined()) {
            // the output field is prepared first so that the forward y-GEMMs on the second stream only depend on the flags
            CF_TRY(spec_output(nse, f));
            f_prepared = true;
            CF_CUDA(cudaEventRecord(ctx->ev_cmp[3], ctx->stream));
            CF_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_cmp[3], 0));
            for (int c = 0; c < 3; ++c) {
                XPassParams xc = xp;
                xc.nfields = 1;
                xc.fsel[0] = c;
                { StageTimer _t(ctx, 3); CF_TRY(xpass_forward_launch(xc, ctx->stream)); }
                CF_TRY(slab_signal(nse, 3 + c, ctx->stream));
            }
        } else {
            { StageTimer _t(ctx, 3); CF_TRY(xpass_forward_launch(xp, ctx->stream)); }
            StageTimer _t(ctx, 8);
            if (!(getenv("CFGPU_FLAG_BARRIER") && atoi(getenv("CFGPU_FLAG_BARRIER")) == 0)) {
                CF_TRY(slab_signal(nse, 7, ctx->stream));
                CF_TRY(slab_wait(nse, 7, ctx->stream));
            } else CF_TRY(comm_barrier(ctx->comm, ctx->stream));
        }
    } else if (!multi) {
        StageTimer _t(ctx, 3);
        CF_TRY(xpass_forward_launch(xp, ctx->stream));
    } else {
        for (int c = 0; c < 3; ++c) {  // x-pass of component c, then its all-to-all overlapping the next x-pass / y-GEMM
            XPassParams xc = xp;
            xc.nfields = 1;
            xc.fsel[0] = c;
            if (peer) {  // this rank's own kx rows go straight into its pencil buffer, the others into the staging
                xc.peer_direct = 2;
                xc.self_rank = ctx->comm.rank;
                xc.peer_out[ctx->comm.rank] = reinterpret_cast<double2*>(ctx->ws_P.ptr);
            }
            { StageTimer _t(ctx, 3); CF_TRY(xpass_forward_launch(xc, ctx->stream)); }
            CF_CUDA(cudaEventRecord(ctx->ev_cmp[c], ctx->stream));
            CF_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_cmp[c], 0));
            if (peer) CF_TRY(slab_push(nse, 3, 1, &c, 1, 3 + c, ctx->comm_stream));
            else {
                CF_TRY(slab_exchange(nse, 3, 1, &c, 1, ctx->comm_stream));
                CF_CUDA(cudaEventRecord(ctx->ev_com[c], ctx->comm_stream));
            }
        }
    }

    // aliased modes of f must be exactly zero (FlowField::zeroPaddedModes, nse.cpp:389-390); the kernels below
    // only ever write retained modes
    if (!f_prepared) CF_TRY(spec_output(nse, f));

    const YPlan* yp;
    const ModeBox* bx;
    CF_TRY(get_yplan(ctx, nse->Ny, nse->a, nse->b, &yp));
    CF_TRY(get_box(ctx, nse->Nx, nse->Nz, nse->Kx, nse->Kz, &bx));
    const int nxl = nse->x1 - nse->x0, nkz = nse->Kz + 1;
    const size_t Pf = (size_t)nse->Ny * nxl * nkz * 2;
    YGemmParams p;
    memset(&p, 0, sizeof p);
    p.N = nse->Ny; p.mode = 1;
    p.fft = yfft_plan(ctx, nse->Ny); p.fft_half = yfft_half_plan(ctx, nse->Ny); p.ya = nse->a; p.yb = nse->b;
    p.M = yp->Ne; p.M2 = yp->No; p.K1 = p.K2 = yp->Nh; p.K1p = p.K2p = yp->fwdKp;
    p.A1[0] = yp->Fe; p.A2[0] = yp->Fo; p.sgn[0] = 1.0;
    p.ncols = (long)nxl * nkz * 2;
    p.in_runlen = 1; p.in_runstart = nullptr; p.in_ld = (long)nxl * nkz * 2;
    const SpecAddr fa = spec_addr(nse, f, bx);
    p.out_runlen = fa.runlen; p.out_runstart = fa.runstart; p.out_ld = fa.ld;
    p.njobs = 3;
    for (int i = 0; i < 3; ++i) {
        p.job[i].in = ctx->ws_P.ptr + i * Pf;
        p.job[i].out[0] = fa.base + i * fa.compstride;
        p.job[i].nmat = 1; p.job[i].mat0 = 0;
    }
    if (fused && fused_pipelined()) {
        cudaStream_t X = ctx->comm_stream;
        for (int c = 0; c < 3; ++c) {
            { StageTimer _t(ctx, 8, X); CF_TRY(slab_wait(nse, 3 + c, X)); }
            YGemmParams pc = p;
            pc.njobs = 1;
            pc.job[0] = p.job[c];
            StageTimer _t(ctx, 4, X);
            CF_TRY(ygemm_launch(pc, X));
        }
        CF_CUDA(cudaEventRecord(ctx->ev_com[3], X));
        CF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_com[3], 0));
    } else if (!multi || fused) {
        StageTimer _t(ctx, 4);
        CF_TRY(ygemm_launch(p, ctx->stream));
    } else {
        for (int c = 0; c < 3; ++c) {
            if (peer) { StageTimer _t(ctx, 8); CF_TRY(slab_wait(nse, 3 + c, ctx->stream)); }
            else CF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_com[c], 0));
            YGemmParams pc = p;
            pc.njobs = 1;
            pc.job[0] = p.job[c];
            StageTimer _t(ctx, 4);
            CF_TRY(ygemm_launch(pc, ctx->stream));
        }
    }
    f->xzstate = CFGPU_SPECTRAL; f->ystate = CFGPU_SPECTRAL;
    if (f->layout == 0) { f->clean_Kx = nse->Kx; f->clean_Kz = nse->Kz; }  // (the flags describe the serial buffer)
    if (nse->cfg.dealias_xz) f->padded = 1;
    return 0;
}

static int fill_tau_params(cfgpu_nse nse, int s, TauSolveParams& tp) {
    CF_ARG(s >= 0 && s < (int)nse->tau.size() && s < (int)nse->lambda_t.size(), "substep index out of range (call reset_lambda first)");
    memset(&tp, 0, sizeof tp);
    tp.td = nse->tau[s];
    tp.g = nse->geom;
    tp.TM_lin = nse->TM_lin;
    tp.taucorr = nse->cfg.taucorrection;
    tp.Ubaseyy = nse->has_Ubaseyy ? nse->d_base : nullptr;
    tp.Wbaseyy = nse->has_Wbaseyy ? nse->d_base + nse->Ny : nullptr;
    tp.constraint = nse->cfg.constraint;
    tp.dPdxRef = nse->cfg.dPdxRef; tp.dPdzRef = nse->cfg.dPdzRef;
    tp.umean_target = nse->cfg.UbulkRef_minus_base; tp.wmean_target = nse->cfg.WbulkRef_minus_base;
    tp.dPd_act = nse->d_scal + 1;
    tp.lin_base_dPdx = nse->lin_base_dPdx; tp.lin_base_dPdz = nse->lin_base_dPdz;
    return 0;
}

int cfgpu_nse_solve(cfgpu_nse nse, int s, int nterms, const double* coef_h, const cfgpu_field* terms, cfgpu_field uout,
                    cfgpu_field qout) {
    CF_ARG(nterms >= 1 && nterms <= TAU_MAXTERMS, "cfgpu_nse_solve: 1..10 terms");
    CF_ARG(nse && shape_ok(nse, uout, 3) && shape_ok(nse, qout, 1), "cfgpu_nse_solve: uout (3 components) and qout (1) must be on the operator's grid");
    TauSolveParams tp;
    CF_TRY(fill_tau_params(nse, s, tp));
    tp.nterms = nterms;
    for (int j = 0; j < nterms; ++j) {
        CF_ARG(shape_ok(nse, terms[j], 3), "cfgpu_nse_solve: term shape mismatch");
        if (nse->use_tile) CF_TRY(field_tile(terms[j], nse->tg));
        else CF_TRY(field_serial(terms[j]));
        tp.term[j] = nse->use_tile ? terms[j]->dtile : terms[j]->dser;
        tp.coef[j] = coef_h[j];
    }
    // only the retained box of the outputs is written; whatever they hold outside it stays (nse.cpp:566-572)
    if (nse->use_tile) {
        CF_TRY(field_tile_output(uout, nse->tg, false));
        CF_TRY(field_tile_output(qout, nse->tg, false));
    } else {
        CF_TRY(field_serial(uout));
        CF_TRY(field_serial(qout));
    }
    tp.uout = nse->use_tile ? uout->dtile : uout->dser;
    tp.qout = nse->use_tile ? qout->dtile : qout->dser;
    tp.tile_layout = nse->use_tile ? 1 : 0;
    {
        static const int pf = getenv("CF_TAU_PREFETCH") ? atoi(getenv("CF_TAU_PREFETCH")) : 1;
        tp.prefetch_terms = pf;
    }
    { StageTimer _t(nse->ctx, 5); CF_TRY(tau_solve_launch(tp, nse->ctx->stream)); }
    uout->xzstate = uout->ystate = qout->xzstate = qout->ystate = CFGPU_SPECTRAL;
    return 0;
}

int cfgpu_nse_linear(cfgpu_nse nse, cfgpu_field u, cfgpu_field q, cfgpu_field L) {
    CF_ARG(nse && shape_ok(nse, u, 3) && shape_ok(nse, q, 1) && shape_ok(nse, L, 3), "cfgpu_nse_linear: u, L (3 components) and q (1) must be on the operator's grid");
    CF_ARG(!nse->tau.empty(), "cfgpu_nse_linear: call reset_lambda first");
    TauSolveParams tp;
    CF_TRY(fill_tau_params(nse, 0, tp));
    if (nse->use_tile) {
        // same policy as the solve: hot-path fields are tile-major; only the retained box of L is written (nse.cpp:393-477)
        CF_TRY(field_tile(u, nse->tg)); CF_TRY(field_tile(q, nse->tg)); CF_TRY(field_tile_output(L, nse->tg, false));
        tp.tile_layout = 1;
        StageTimer _t(nse->ctx, 6);
        CF_TRY(linear_launch(tp, u->dtile, q->dtile, L->dtile, nse->ctx->stream));
    } else {
        CF_TRY(field_serial(u)); CF_TRY(field_serial(q)); CF_TRY(field_serial(L));
        StageTimer _t(nse->ctx, 6);
        CF_TRY(linear_launch(tp, u->dser, q->dser, L->dser, nse->ctx->stream));
    }
    L->xzstate = L->ystate = CFGPU_SPECTRAL;
    return 0;
}

int cfgpu_nse_cflfactor(cfgpu_nse nse, cfgpu_field u, double* out_h) {
    CF_ARG(nse && shape_ok(nse, u, 3), "cfgpu_nse_cflfactor: u must be a 3-component field on the operator's grid");
    CF_ARG(u->xzstate == CFGPU_SPECTRAL && u->ystate == CFGPU_SPECTRAL, "cfgpu_nse_cflfactor: u must be spectral");
    cfgpu_ctx ctx = nse->ctx;
    CF_TRY(inverse_to_Q(nse, u, false));
    ZPassParams zp;
    CF_TRY(fill_zpass(nse, zp, ZP_CFL));
    CF_CUDA(cudaMemsetAsync(nse->d_scal, 0, sizeof(double), ctx->stream));
    CF_TRY(zpass_launch(zp, ctx->stream));
    CF_TRY(comm_allreduce(ctx->comm, nse->d_scal, 1, 1, ctx->stream));
    CF_CUDA(cudaMemcpyAsync(out_h, nse->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- single systems for the host classes HelmholtzSolver / BandedTridiag / TauSolver (tests, tools): host arrays in and
// out, the arithmetic on the device
struct MinLaneBlock {  // see tau_pick_E: longer per-lane chains for the single-system classes, restored on scope exit
    int old;
    explicit MinLaneBlock(int e) : old(tau_set_min_E(e)) {}
    ~MinLaneBlock() { tau_set_min_E(old); }
};
static int scratch(cfgpu_ctx ctx, size_t doubles, double** p) {
    CF_TRY(ws_reserve(ctx->ws_red, (1 << 20) > doubles * 8 ? (1 << 20) : doubles * 8));
    *p = ctx->ws_red.ptr;
    return 0;
}
int cfgpu_helmholtz_solve(cfgpu_ctx ctx, int N, double a, double b, double lambda, double nu, int ncols, const double* f_h,
                          const double* ua_h, const double* ub_h, double* u_h) {
    CF_ARG(ctx && f_h && ua_h && ub_h && u_h && ncols >= 1, "cfgpu_helmholtz_solve: bad argument");
    double* d;
    const size_t nf = (size_t)ncols * N;
    MinLaneBlock sequential_chains(8);
    CF_TRY(scratch(ctx, 2 * nf + 2 * ncols, &d));
    double *df = d, *du = d + nf, *dua = du + nf, *dub = dua + ncols;
    CF_CUDA(cudaMemcpyAsync(df, f_h, nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CF_CUDA(cudaMemcpyAsync(dua, ua_h, ncols * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CF_CUDA(cudaMemcpyAsync(dub, ub_h, ncols * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CF_TRY(helmholtz_batch_launch(N, a, b, lambda, nu, ncols, df, dua, dub, du, ctx->stream));
    CF_CUDA(cudaMemcpyAsync(u_h, du, nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cfgpu_tridiag(cfgpu_ctx ctx, int op, int M, double* a_h, double* invdiag_h, double* x_h, double* y_h, int nx, int offset, int stride) {
    CF_ARG(ctx && a_h && invdiag_h && M >= 2 && op >= 0 && op <= 2, "cfgpu_tridiag: bad argument");
    CF_ARG(op == 0 || (x_h && nx >= offset + stride * (M - 1) + 1 && (offset == 0 || offset == 1) && (stride == 1 || stride == 2)),
           "cfgpu_tridiag: offset must be 0 or 1, stride 1 or 2");
    double* d;
    const size_t na = 4 * (size_t)M - 2;
    CF_TRY(scratch(ctx, na + M + 2 * (size_t)(nx > 0 ? nx : 1), &d));
    double *da = d, *di = d + na, *dx = di + M, *dy = dx + (nx > 0 ? nx : 1);
    CF_CUDA(cudaMemcpyAsync(da, a_h, na * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CF_CUDA(cudaMemcpyAsync(di, invdiag_h, M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (op != 0) CF_CUDA(cudaMemcpyAsync(dx, x_h, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (op == 2) CF_CUDA(cudaMemcpyAsync(dy, y_h, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));  // untouched entries survive
    CF_TRY(tridiag_launch(op, M, da, di, dx, dy, offset, stride, ctx->stream));
    if (op == 0) {
        CF_CUDA(cudaMemcpyAsync(a_h, da, na * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CF_CUDA(cudaMemcpyAsync(invdiag_h, di, M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    } else if (op == 1) CF_CUDA(cudaMemcpyAsync(x_h, dx, nx * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    else CF_CUDA(cudaMemcpyAsync(y_h, dy, nx * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// TauSolver::solve for ONE Fourier mode (tausolver.cpp:347-450) through the batched kernels: a throw-away operator on the
// smallest grid that holds the mode (un-dealiased: every mode below the Nyquist ones is solved), the right-hand side in
// that mode.  constraint 1: mean mode with prescribed bulk velocities (helmholtz.cpp:158-213), returns dPdx, dPdz.
int cfgpu_tausolve_mode(cfgpu_ctx ctx, int N, int kx, int kz, double Lx, double Lz, double a, double b, double lambda, double nu,
                        int taucorrection, int constraint, double umean, double wmean, const double* R_h, double* out_h, double* dPd_h) {
    CF_ARG(ctx && R_h && out_h && kz >= 0, "cfgpu_tausolve_mode: bad argument (kz >= 0: the stored half of the spectrum)");
    CF_ARG(ctx->comm.nranks == 1, "cfgpu_tausolve_mode: single-GPU call");
    const int ak = kx < 0 ? -kx : kx;
    const int Nx = 2 * ak + 4, Nz = 2 * kz + 4;
    MinLaneBlock sequential_chains(8);
    cfgpu_nse_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.nu = nu; cfg.nonlinearity = 0; cfg.dealias_xz = 0; cfg.dealias_y = 0; cfg.taucorrection = taucorrection;
    cfg.constraint = constraint; cfg.UbulkRef_minus_base = umean; cfg.WbulkRef_minus_base = wmean;
    cfgpu_nse nse = nullptr;
    cfgpu_field R = nullptr, U = nullptr, Q = nullptr;
    int rc = 1;
    do {
        if (cfgpu_nse_create(ctx, Nx, N, Nz, Lx, Lz, a, b, &cfg, nullptr, nullptr, &nse)) break;
        const double PI = 3.14159265358979323846264338327950288;
        const double lam_t = lambda - 4.0 * (PI * PI) * nu * ((kx / Lx) * (kx / Lx) + (kz / Lz) * (kz / Lz));
        if (cfgpu_nse_reset_lambda(nse, &lam_t, 1)) break;
        if (cfgpu_field_create(ctx, Nx, N, Nz, 3, Lx, Lz, a, b, &R) || cfgpu_field_create(ctx, Nx, N, Nz, 3, Lx, Lz, a, b, &U) ||
            cfgpu_field_create(ctx, Nx, N, Nz, 1, Lx, Lz, a, b, &Q))
            break;
        const int mx = kx >= 0 ? kx : Nx + kx;
        bool ok = true;
        for (int i = 0; i < 3 && ok; ++i) ok = cfgpu_field_add_profile(R, mx, kz, i, R_h + (size_t)i * 2 * N, 1.0) == 0;
        if (!ok) break;
        const double one = 1.0;
        if (cfgpu_nse_solve(nse, 0, 1, &one, &R, U, Q)) break;
        for (int i = 0; i < 3 && ok; ++i) ok = cfgpu_field_get_profile(U, mx, kz, i, out_h + (size_t)i * 2 * N) == 0;
        ok = ok && cfgpu_field_get_profile(Q, mx, kz, 0, out_h + (size_t)3 * 2 * N) == 0;
        if (!ok) break;
        if (dPd_h) { if (cfgpu_nse_get_dPd(nse, dPd_h, dPd_h + 1)) break; }
        rc = 0;
    } while (false);
    if (R) cfgpu_field_destroy(R);
    if (U) cfgpu_field_destroy(U);
    if (Q) cfgpu_field_destroy(Q);
    if (nse) cfgpu_nse_destroy(nse);
    return rc;
}

int cfgpu_nse_get_dPd(cfgpu_nse nse, double* dPdx_h, double* dPdz_h) {
    double v[2];
    if (nse->ctx->comm.nranks > 1) {  // only the owner of the (0,0) mode computed them; the others hold zeros
        CF_CUDA(cudaMemcpyAsync(nse->d_scal + 4, nse->d_scal + 1, 2 * sizeof(double), cudaMemcpyDeviceToDevice, nse->ctx->stream));
        CF_TRY(comm_allreduce(nse->ctx->comm, nse->d_scal + 4, 2, 0, nse->ctx->stream));
        CF_CUDA(cudaMemcpyAsync(v, nse->d_scal + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, nse->ctx->stream));
    } else
    CF_CUDA(cudaMemcpyAsync(v, nse->d_scal + 1, 2 * sizeof(double), cudaMemcpyDeviceToHost, nse->ctx->stream));
    CF_CUDA(cudaStreamSynchronize(nse->ctx->stream));
    *dPdx_h = v[0];
    *dPdz_h = v[1];
    return 0;
}

}  // extern "C"
