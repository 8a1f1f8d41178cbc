// Implementation of the C-ABI declared in include/cfgpu.h (context, FlowField storage, transforms, norms).
// The NSE operator entry points are in cfgpu_nse.cu.
#include <cmath>
#include <cstring>

#include "cfgpu_internal.h"
#include "diffops.cuh"
#include "fieldops.cuh"
#include "tau.cuh"

namespace cfgpu {

long long g_launches = 0;
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

static const long double PIl = 3.141592653589793238462643383279502884L;

int ws_reserve(Workspace& w, size_t bytes) {
    if (bytes <= w.bytes) return 0;
    if (w.exported) {
        // other ranks hold CUDA-IPC mappings of this buffer: it may only be replaced through ensure_peers (cfgpu_nse.cu),
        // which closes the mappings collectively first
        set_last_error("internal: a peer-mapped workspace must be resized collectively");
        return 1;
    }
    if (w.ptr) CF_CUDA(cudaFree(w.ptr));
    w.ptr = nullptr;
    w.bytes = 0;
    CF_CUDA(cudaMalloc((void**)&w.ptr, bytes));
    w.bytes = bytes;
    ++w.gen;  // every (re)allocation gets a new generation: cached peer mappings / row tables are keyed on it, not on the address
    return 0;
}

// ------------------------------------------------------------------------------------------- stage profiler
int stage_begin(cfgpu_ctx ctx, int st, cudaStream_t stream) {
    if (!ctx->profiling || ctx->capturing) return 0;
    if (!stream) stream = ctx->stream;
    if (ctx->prof_used[st] == ctx->prof_ev[st].size()) {
        cudaEvent_t a, b;
        CF_CUDA(cudaEventCreate(&a));
        CF_CUDA(cudaEventCreate(&b));
        ctx->prof_ev[st].push_back({a, b});
    }
    CF_CUDA(cudaEventRecord(ctx->prof_ev[st][ctx->prof_used[st]].first, stream));
    return 0;
}
int stage_end(cfgpu_ctx ctx, int st, cudaStream_t stream) {
    if (!ctx->profiling || ctx->capturing) return 0;
    if (!stream) stream = ctx->stream;
    CF_CUDA(cudaEventRecord(ctx->prof_ev[st][ctx->prof_used[st]].second, stream));
    ctx->prof_used[st]++;
    ctx->prof_calls[st]++;
    return 0;
}
static int prof_collect(cfgpu_ctx ctx) {
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    for (int st = 0; st < CFGPU_NSTAGES; ++st) {
        for (size_t i = 0; i < ctx->prof_used[st]; ++i) {
            float ms = 0;
            CF_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[st][i].first, ctx->prof_ev[st][i].second));
            ctx->prof_ms[st] += ms;
        }
        ctx->prof_used[st] = 0;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------- plan builders
static int upload(const std::vector<double>& h, double** d) {
    CF_CUDA(cudaMalloc((void**)d, h.size() * sizeof(double)));
    CF_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

static long double cos_pi_frac(long num, long den) {
    // cos(pi * num / den) with exact argument reduction
    num %= 2 * den;
    if (num < 0) num += 2 * den;
    if (num > den) num = 2 * den - num;  // cos(2pi - x) = cos x
    if (2 * num == den) return 0.0L;
    if (2 * num > den) return -cosl(PIl * (long double)(den - num) / (long double)den);
    return cosl(PIl * (long double)num / (long double)den);
}

int get_yplan(cfgpu_ctx ctx, int N, double a, double b, const YPlan** out) {
    auto key = std::make_tuple(N, a, b);
    auto it = ctx->yplans.find(key);
    if (it != ctx->yplans.end()) {
        *out = &it->second;
        return 0;
    }
    if (N < 2) {
        set_last_error("y transform needs Ny >= 2");
        return 1;
    }
    YPlan pl;
    pl.N = N; pl.a = a; pl.b = b;
    const int Nb = N - 1;
    pl.Nh = (N + 1) / 2;
    pl.Ne = Nb / 2 + 1;
    pl.No = N - pl.Ne;
    const int Nh = pl.Nh, Ne = pl.Ne, No = pl.No;
    auto r8 = [](int x) { return (x + 31) & ~31; };  // matrices are zero padded to whole 32-row warp tiles
    auto r4 = [](int x) { return (x + 7) & ~7; };  // inner dimensions padded to whole k-steps of the 16x8x8 MMA
    pl.invMp = r8(Nh); pl.invK1p = r4(Ne); pl.invK2p = r4(No > 0 ? No : 1);
    pl.fwdMp = r8(Ne > No ? Ne : No); pl.fwdKp = r4(Nh);

    // C[j][n] = cos(pi j n / Nb) ; derivative operator on coefficients (chebyshev.cpp:672-697)
    std::vector<long double> C((size_t)N * N), CD((size_t)N * N, 0.0L);
    for (int j = 0; j < N; ++j)
        for (int n = 0; n < N; ++n) C[(size_t)j * N + n] = cos_pi_frac((long)j * n, Nb);
    const long double scale = 4.0L / ((long double)b - (long double)a);
    for (int j = 0; j < N; ++j)
        for (int m = 0; m < N; ++m) {
            long double s = 0.0L;
            for (int n = (m - 1); n >= 0; n -= 2) s += C[(size_t)j * N + n] * (n == 0 ? 0.5L : 1.0L);
            CD[(size_t)j * N + m] = s * scale * (long double)m;
        }
    std::vector<double> Ce((size_t)pl.invMp * pl.invK1p, 0.0), Co((size_t)pl.invMp * pl.invK2p, 0.0);
    std::vector<double> CDe(Ce.size(), 0.0), CDo(Co.size(), 0.0);
    for (int j = 0; j < Nh; ++j) {
        for (int r = 0; r < Ne; ++r) {
            Ce[(size_t)j * pl.invK1p + r] = (double)C[(size_t)j * N + 2 * r];
            CDe[(size_t)j * pl.invK1p + r] = (double)CD[(size_t)j * N + 2 * r];
        }
        for (int r = 0; r < No; ++r) {
            Co[(size_t)j * pl.invK2p + r] = (double)C[(size_t)j * N + 2 * r + 1];
            CDo[(size_t)j * pl.invK2p + r] = (double)CD[(size_t)j * N + 2 * r + 1];
        }
    }
    // forward: c_n = w_n sum_j g_j cos(pi j n/Nb) x_j  (flowfield.cpp:1913-1934)
    std::vector<double> Fe((size_t)pl.fwdMp * pl.fwdKp, 0.0), Fo(Fe.size(), 0.0);
    for (int n = 0; n < N; ++n) {
        const long double wn = ((n == 0 || n == Nb) ? 0.5L : 1.0L) / (long double)Nb;
        for (int j = 0; j < Nh; ++j) {
            const long double gj = (j == 0 || j == Nb) ? 1.0L : 2.0L;
            const double v = (double)(wn * gj * C[(size_t)j * N + n]);
            if (n % 2 == 0) Fe[(size_t)(n / 2) * pl.fwdKp + j] = v;
            else Fo[(size_t)(n / 2) * pl.fwdKp + j] = v;
        }
    }
    // GD = (d/dy on coefficients, chebyshev.cpp:672-697) o (forward transform): physical profile -> coefficients of its
    // y-derivative, used by the divergence / skew-symmetric forms (diffops.cpp:2540-2549, 3266-3277)
    std::vector<double> GDe[2], GDo[2];
    {
        std::vector<long double> Ff((size_t)N * N), GD((size_t)N * N, 0.0L);
        for (int n = 0; n < N; ++n) {
            const long double wn = ((n == 0 || n == Nb) ? 0.5L : 1.0L) / (long double)Nb;
            for (int j = 0; j < N; ++j) Ff[(size_t)n * N + j] = wn * ((j == 0 || j == Nb) ? 1.0L : 2.0L) * C[(size_t)j * N + n];
        }
        for (int n = 0; n < N; ++n)
            for (int m = n + 1; m < N; m += 2) {
                const long double dm = scale * (long double)m * (n == 0 ? 0.5L : 1.0L);
                for (int j = 0; j < N; ++j) GD[(size_t)n * N + j] += dm * Ff[(size_t)m * N + j];
            }
        for (int h = 0; h < 2; ++h) {
            const long double cf = h ? 0.5L : 1.0L;
            GDe[h].assign((size_t)pl.fwdMp * pl.fwdKp, 0.0);
            GDo[h].assign((size_t)pl.fwdMp * pl.fwdKp, 0.0);
            for (int n = 0; n < N; ++n)
                for (int j = 0; j < Nh; ++j) {
                    const bool mid = (2 * j == Nb);
                    if (n % 2 == 0) { if (!mid) GDe[h][(size_t)(n / 2) * pl.fwdKp + j] = (double)(cf * GD[(size_t)n * N + j]); }
                    else GDo[h][(size_t)(n / 2) * pl.fwdKp + j] = (double)(cf * GD[(size_t)n * N + j]);
                }
        }
    }
    // Gram weights <T_m,T_n> (chebyshev.cpp:758-802), FP64 arithmetic (the reference's int version overflows for Ny > 215)
    std::vector<double> W((size_t)N * N, 0.0);
    for (int m = 0; m < N; ++m)
        for (int n = m % 2; n < N; n += 2) {
            const double e = 1.0, dm = m, dn = n;
            W[(size_t)m * N + n] = (e - dm * dm - dn * dn) / ((e + dm - dn) * (e - dm + dn) * (e + dm + dn) * (e - dm - dn));
        }
    CF_TRY(upload(Ce, &pl.Ce)); CF_TRY(upload(Co, &pl.Co)); CF_TRY(upload(CDe, &pl.CDe)); CF_TRY(upload(CDo, &pl.CDo));
    CF_TRY(upload(Fe, &pl.Fe)); CF_TRY(upload(Fo, &pl.Fo)); CF_TRY(upload(W, &pl.Wgram));
    {   // pi/2 c_n on the diagonal, c_0 = 2 (the weight-1/sqrt(1-y^2) inner product of T_m, T_n)
        std::vector<double> Wc((size_t)N * N, 0.0);
        for (int n = 0; n < N; ++n) Wc[(size_t)n * N + n] = 0.5 * 3.14159265358979323846264338327950288 * (n == 0 ? 2.0 : 1.0);
        CF_TRY(upload(Wc, &pl.Wcheb));
    }
    for (int h = 0; h < 2; ++h) { CF_TRY(upload(GDe[h], &pl.GDe[h])); CF_TRY(upload(GDo[h], &pl.GDo[h])); }
    auto res = ctx->yplans.emplace(key, pl);
    *out = &res.first->second;
    return 0;
}

// FFT plan for the y-transform of Ny points (yfft.cu) or null: unsupported length, or CF_YFFT=0 (the DMMA contraction)
const FftPlanDev* yfft_plan(cfgpu_ctx ctx, int Ny) {
    static const int on = getenv("CF_YFFT") ? atoi(getenv("CF_YFFT")) : 1;
    if (!on || !yfft_length_supported(Ny)) return nullptr;
    const FftPlanDev* pl = nullptr;
    if (get_fftplan(ctx, 2 * (Ny - 1), &pl)) return nullptr;
    return pl;
}

const FftPlanDev* yfft_half_plan(cfgpu_ctx ctx, int Ny) {
    if (!yfft_plan(ctx, Ny) || (Ny - 1) % 2) return nullptr;
    const FftPlanDev* pl = nullptr;
    if (get_fftplan(ctx, Ny - 1, &pl)) return nullptr;
    return pl;
}

int get_fftplan(cfgpu_ctx ctx, int N, const FftPlanDev** out) {
    auto it = ctx->fftplans.find(N);
    if (it != ctx->fftplans.end()) {
        *out = &it->second.dev;
        return 0;
    }
    FftPlanHost pl;
    pl.dev.N = N;
    pl.dev.npass = 0;
    int m = N;
    while (m > 1) {
        int r;
        if (m % 8 == 0) r = 8;
        else if (m % 4 == 0) r = 4;
        else if (m % 2 == 0) r = 2;
        else if (m % 3 == 0) r = 3;
        else if (m % 5 == 0) r = 5;
        else {
            set_last_error("FFT length must factor into 2,3,5: " + std::to_string(N));
            return 1;
        }
        if (pl.dev.npass >= FFT_MAXPASS) { set_last_error("FFT length too large"); return 1; }
        pl.dev.radix[pl.dev.npass++] = r;
        m /= r;
    }
    std::vector<double> tw(2 * (size_t)N);
    for (int t = 0; t < N; ++t) {
        // exp(-2 pi i t / N) = cos(pi 2t/N) - i sin(pi 2t/N); sin x = cos(x - pi/2)
        tw[2 * t] = (double)cos_pi_frac(2L * t * 2, 2L * N);
        tw[2 * t + 1] = (double)(-cos_pi_frac(2L * t * 2 - N, 2L * N));
    }
    double* d = nullptr;
    CF_TRY(upload(tw, &d));
    pl.tw = reinterpret_cast<double2*>(d);
    pl.dev.tw = pl.tw;
    {
        pl.dev.rev = nullptr;
        std::vector<int> rev((size_t)N);
        for (int n = 0; n < N; ++n) rev[n] = fft_plan_rev(pl.dev, n);
        int* dr = nullptr;
        CF_CUDA(cudaMalloc((void**)&dr, rev.size() * sizeof(int)));
        CF_CUDA(cudaMemcpy(dr, rev.data(), rev.size() * sizeof(int), cudaMemcpyHostToDevice));
        pl.dev.rev = dr;
    }
    auto res = ctx->fftplans.emplace(N, pl);
    *out = &res.first->second.dev;
    return 0;
}

int get_box(cfgpu_ctx ctx, int Nx, int Nz, int Kx, int Kz, const ModeBox** out) {
    auto key = std::make_tuple(Nx, Nz, Kx, Kz);
    auto it = ctx->boxes.find(key);
    if (it != ctx->boxes.end()) {
        *out = &it->second;
        return 0;
    }
    ModeBox bx;
    bx.Nx = Nx; bx.Nz = Nz; bx.Kx = Kx; bx.Kz = Kz;
    const int nmx = 2 * Kx + 1, Nzpad = 2 * (Nz / 2 + 1);
    std::vector<long> rs(nmx);
    for (int mxi = 0; mxi < nmx; ++mxi) {
        const int kx = mxi <= Kx ? mxi : mxi - nmx;
        const int mx = kx >= 0 ? kx : Nx + kx;
        rs[mxi] = (long)mx * Nzpad;
    }
    CF_CUDA(cudaMalloc((void**)&bx.runstart_full, nmx * sizeof(long)));
    CF_CUDA(cudaMemcpy(bx.runstart_full, rs.data(), nmx * sizeof(long), cudaMemcpyHostToDevice));
    auto res = ctx->boxes.emplace(key, bx);
    *out = &res.first->second;
    return 0;
}

void cheb_diff_host(const std::vector<double>& u, std::vector<double>& d, double a, double b) {
    const int N = (int)u.size(), Nb = N - 1;
    d.assign(N, 0.0);
    if (Nb <= 0) return;
    const double scale = 4.0 / (b - a);
    d[Nb] = 0.0;
    d[Nb - 1] = scale * Nb * u[Nb];
    for (int n = Nb - 2; n >= 0; --n) d[n] = d[n + 2] + scale * (n + 1) * u[n + 1];
    d[0] *= 0.5;
}
void cheb_to_physical_host(const std::vector<double>& c, std::vector<double>& u) {
    const int N = (int)c.size(), Nb = N - 1;
    u.assign(N, 0.0);
    for (int j = 0; j < N; ++j) {
        long double s = 0.0L;
        for (int n = 0; n < N; ++n) s += (long double)c[n] * cos_pi_frac((long)j * n, Nb > 0 ? Nb : 1);
        u[j] = (double)s;
    }
}

}  // namespace cfgpu

using namespace cfgpu;

#define CF_ARG(cond, msg)            \
    do {                             \
        if (!(cond)) {               \
            set_last_error(msg);     \
            return 1;                \
        }                            \
    } while (0)

extern "C" {

const char* cfgpu_last_error(void) { return g_last_error.c_str(); }
const char* cfgpu_version(void) {
#ifdef CF_EMU
    return "cfgpu 0.1 (CPU emulation build: tests only)";
#else
    return "cfgpu 0.1 (sm_100a)";
#endif
}

int cfgpu_init(int device, cfgpu_ctx* out) {
    CF_ARG(out, "cfgpu_init: null out");
    int ndev = 0;
    CF_CUDA(cudaGetDeviceCount(&ndev));
    CF_ARG(ndev > 0, "cfgpu_init: no CUDA device visible (this library has no CPU fallback)");
    CF_ARG(device >= 0 && device < ndev, "cfgpu_init: bad device index");
    CF_CUDA(cudaSetDevice(device));
    cfgpu_ctx ctx = new cfgpu_ctx_s();
    ctx->device = device;
    CF_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CF_CUDA(cudaEventCreate(&ctx->ev0));
    CF_CUDA(cudaEventCreate(&ctx->ev1));
    {   // the exchange stream outranks the compute stream: its few CTAs are scheduled ahead of the queued transform CTAs
        int lo = 0, hi = 0;
        CF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CF_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    }
    for (int i = 0; i < 4; ++i) {
        CF_CUDA(cudaEventCreateWithFlags(&ctx->ev_cmp[i], cudaEventDisableTiming));
        CF_CUDA(cudaEventCreateWithFlags(&ctx->ev_com[i], cudaEventDisableTiming));
    }
    *out = ctx;
    return 0;
}

int cfgpu_finalize(cfgpu_ctx ctx) {
    if (!ctx) return 0;
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->yplans) {
        YPlan& p = kv.second;
        cudaFree(p.Ce); cudaFree(p.Co); cudaFree(p.CDe); cudaFree(p.CDo); cudaFree(p.Fe); cudaFree(p.Fo); cudaFree(p.Wgram); cudaFree(p.Wcheb);
        for (int h = 0; h < 2; ++h) { cudaFree(p.GDe[h]); cudaFree(p.GDo[h]); }
    }
    for (auto& kv : ctx->fftplans) cudaFree(kv.second.tw);
    for (auto& kv : ctx->boxes) cudaFree(kv.second.runstart_full);
    for (auto& kv : ctx->pool_free_blocks) cudaFree(kv.second);
    ctx->pool_free_blocks.clear();
    if (ctx->ws_P.ptr) cudaFree(ctx->ws_P.ptr);
    if (ctx->ws_Q.ptr) cudaFree(ctx->ws_Q.ptr);
    if (ctx->ws_red.ptr) cudaFree(ctx->ws_red.ptr);
    if (ctx->ws_G.ptr) cudaFree(ctx->ws_G.ptr);
    if (ctx->peerP_gen) comm_close_peers(ctx->comm, ctx->peerP);
    if (ctx->peerS_gen) comm_close_peers(ctx->comm, ctx->peerS);
    if (ctx->ws_S.ptr) cudaFree(ctx->ws_S.ptr);
    if (ctx->peerF_gen) comm_close_peers(ctx->comm, ctx->peerF);
    if (ctx->ws_F.ptr) cudaFree(ctx->ws_F.ptr);
    comm_destroy(ctx->comm);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 4; ++i) { cudaEventDestroy(ctx->ev_cmp[i]); cudaEventDestroy(ctx->ev_com[i]); }
    cudaStreamSynchronize(ctx->comm_stream);
    cudaStreamDestroy(ctx->comm_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int cfgpu_sync(cfgpu_ctx ctx) {
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cfgpu_launch_count(cfgpu_ctx, long long* n) {
    *n = g_launches;
    return 0;
}
int cfgpu_timer_start(cfgpu_ctx ctx) {
    CF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return 0;
}
int cfgpu_timer_stop(cfgpu_ctx ctx, double* ms) {
    CF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    CF_CUDA(cudaEventSynchronize(ctx->ev1));
    float f = 0;
    CF_CUDA(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
    *ms = f;
    return 0;
}

int cfgpu_profile_enable(cfgpu_ctx ctx, int on) {
    CF_TRY(prof_collect(ctx));
    ctx->profiling = on != 0;
    return 0;
}
int cfgpu_profile_read(cfgpu_ctx ctx, double* ms_h, long long* calls_h, int reset) {
    CF_TRY(prof_collect(ctx));
    for (int st = 0; st < CFGPU_NSTAGES; ++st) {
        ms_h[st] = ctx->prof_ms[st];
        calls_h[st] = ctx->prof_calls[st];
        if (reset) { ctx->prof_ms[st] = 0; ctx->prof_calls[st] = 0; }
    }
    return 0;
}

#ifndef CF_EMU
// CUDA-graph replay of a fixed launch sequence (the launch-bound small grids: `order` consecutive SBDF steps return every
// history buffer to its role, so the captured sequence can be replayed; host/dnsalgo.cpp:MultistepDNS::advance).  The
// caller runs the sequence once eagerly first (work spaces get allocated), then once between begin/end: nothing executes
// during capture, cfgpu_graph_launch executes it.  A synchronising or allocating call inside the capture fails with the
// CUDA capture error and the capture is discarded.
struct GraphSlot { cudaGraphExec_t exec; long long kernels; };
static long long g_capture_launches0 = 0;
int cfgpu_graph_begin(cfgpu_ctx ctx) {
    CF_ARG(ctx && !ctx->capturing, "graph capture already active");
    CF_ARG(ctx->comm.nranks == 1, "cfgpu_graph_begin: single-GPU contexts only (the slab exchange runs on several streams)");
    CF_ARG(!ctx->profiling, "cfgpu_graph_begin: per-stage profiling is on (its event timers cannot be replayed)");
    CF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    g_capture_launches0 = g_launches;
    return 0;
}
int cfgpu_graph_end(cfgpu_ctx ctx, int* graph_id) {
    CF_ARG(ctx && ctx->capturing && graph_id, "no graph capture active");
    cudaGraph_t g = nullptr;
    ctx->capturing = false;
    const long long kernels = g_launches - g_capture_launches0;
    g_launches = g_capture_launches0;  // nothing ran: the launches are counted when the graph is launched
    CF_CUDA(cudaStreamEndCapture(ctx->stream, &g));
    cudaGraphExec_t ge = nullptr;
    const cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    CF_CUDA(e);
    GraphSlot* slot = new GraphSlot{ge, kernels};
    for (size_t i = 0; i < ctx->graphs.size(); ++i)
        if (!ctx->graphs[i]) { ctx->graphs[i] = slot; *graph_id = (int)i; return 0; }
    ctx->graphs.push_back((void*)slot);
    *graph_id = (int)ctx->graphs.size() - 1;
    return 0;
}
int cfgpu_graph_abort(cfgpu_ctx ctx) {
    CF_ARG(ctx, "cfgpu_graph_abort: bad argument");
    if (!ctx->capturing) return 0;
    ctx->capturing = false;
    g_launches = g_capture_launches0;
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(ctx->stream, &g);  // invalidated captures return an error and no graph: both fine here
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    return 0;
}
int cfgpu_graph_launch(cfgpu_ctx ctx, int graph_id) {
    CF_ARG(ctx && graph_id >= 0 && graph_id < (int)ctx->graphs.size() && ctx->graphs[graph_id], "bad graph id");
    GraphSlot* slot = (GraphSlot*)ctx->graphs[graph_id];
    CF_CUDA(cudaGraphLaunch(slot->exec, ctx->stream));
    g_launches += slot->kernels;
    return 0;
}
int cfgpu_graph_destroy(cfgpu_ctx ctx, int graph_id) {
    CF_ARG(ctx && graph_id >= 0 && graph_id < (int)ctx->graphs.size() && ctx->graphs[graph_id], "bad graph id");
    GraphSlot* slot = (GraphSlot*)ctx->graphs[graph_id];
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaGraphExecDestroy(slot->exec);
    delete slot;
    ctx->graphs[graph_id] = nullptr;
    return 0;
}
#else
int cfgpu_graph_begin(cfgpu_ctx) { set_last_error("graphs unavailable in the emulation build"); return 1; }
int cfgpu_graph_end(cfgpu_ctx, int*) { set_last_error("graphs unavailable in the emulation build"); return 1; }
int cfgpu_graph_launch(cfgpu_ctx, int) { set_last_error("graphs unavailable in the emulation build"); return 1; }
int cfgpu_graph_abort(cfgpu_ctx) { return 0; }
int cfgpu_graph_destroy(cfgpu_ctx, int) { set_last_error("graphs unavailable in the emulation build"); return 1; }
#endif

// ------------------------------------------------------------------------------------------------ multi-GPU
int cfgpu_comm_unique_id(void* id128_h) { return comm_unique_id(id128_h); }
int cfgpu_comm_init_nccl(cfgpu_ctx ctx, int rank, int nranks, const void* id128_h) {
    CF_ARG(ctx && id128_h, "cfgpu_comm_init_nccl: null argument");
    return comm_init_nccl(ctx->comm, rank, nranks, id128_h);
}
int cfgpu_comm_init_external(cfgpu_ctx ctx, int rank, int nranks, cfgpu_exchange_fn ex, cfgpu_allreduce_fn ar, void* user) {
    CF_ARG(ctx && ex && ar, "cfgpu_comm_init_external: null argument");
    CF_ARG(nranks >= 1 && nranks <= COMM_MAXRANKS && rank >= 0 && rank < nranks, "cfgpu_comm_init_external: bad rank / world size");
    ctx->comm.rank = rank; ctx->comm.nranks = nranks;
    ctx->comm.ext_exchange = ex; ctx->comm.ext_allreduce = ar; ctx->comm.ext_user = user;
    return 0;
}
int cfgpu_comm_rank(cfgpu_ctx ctx, int* rank, int* nranks) {
    if (rank) *rank = ctx->comm.rank;
    if (nranks) *nranks = ctx->comm.nranks;
    return 0;
}
int cfgpu_comm_ranges(cfgpu_ctx ctx, int nmx, int Ny, int rank, int* x0, int* x1, int* y0, int* y1) {
    CF_ARG(rank >= 0 && rank < ctx->comm.nranks, "cfgpu_comm_ranges: bad rank");
    int a, b;
    part_range(nmx, ctx->comm.nranks, rank, a, b);
    if (x0) *x0 = a;
    if (x1) *x1 = b;
    part_range(Ny, ctx->comm.nranks, rank, a, b);
    if (y0) *y0 = a;
    if (y1) *y1 = b;
    return 0;
}
// All-gather of the owned kx rows of a spectral, de-aliased field.  Staging layout per owner s: [i][my][mxi in X_s][kz<=Kz].
}  // extern "C"

namespace cfgpu {
// ------------------------------------------------------------------------------------------------ pooled device memory
static constexpr size_t POOL_MAX_BLOCK = (size_t)64 << 20, POOL_MAX_CACHED = (size_t)2 << 30;
int dev_alloc(cfgpu_ctx ctx, void** p, size_t bytes) {
    if (bytes == 0) bytes = 8;
    const bool pooled = ctx->comm.nranks == 1 && bytes <= POOL_MAX_BLOCK && !getenv("CFGPU_NO_POOL");
    if (pooled) {
        auto it = ctx->pool_free_blocks.find(bytes);
        if (it != ctx->pool_free_blocks.end()) {
            *p = it->second;
            ctx->pool_free_blocks.erase(it);
            ctx->pool_cached_bytes -= bytes;
            return 0;
        }
    }
    if (cudaMalloc(p, bytes) != cudaSuccess) {
        cudaGetLastError();
        // give the cache back and try once more
        for (auto& kv : ctx->pool_free_blocks) { ctx->pool_sizes.erase(kv.second); cudaFree(kv.second); }
        ctx->pool_free_blocks.clear();
        ctx->pool_cached_bytes = 0;
        if (cudaMalloc(p, bytes) != cudaSuccess) {
            cudaGetLastError();
            *p = nullptr;
            return 1;
        }
    }
    if (pooled) ctx->pool_sizes[*p] = bytes;
    return 0;
}
void dev_free(cfgpu_ctx ctx, void* p) {
    if (!p) return;
    auto it = ctx->pool_sizes.find(p);
    if (it != ctx->pool_sizes.end() && ctx->pool_cached_bytes + it->second <= POOL_MAX_CACHED) {
        ctx->pool_free_blocks.emplace(it->second, p);
        ctx->pool_cached_bytes += it->second;
        return;
    }
    if (it != ctx->pool_sizes.end()) ctx->pool_sizes.erase(it);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(p);
}

// ------------------------------------------------------------------------------------------------ field layouts
// The serial buffer is allocated on first use: hot-path fields of a multi-GPU run live in their tile-major buffer only
// (this rank's modes), so a rank's footprint scales with 1/nranks.  A field without a serial buffer is all zero outside
// whatever its tile buffer holds.
int field_ser_alloc(cfgpu_field f) {
    if (f->dser) return 0;
    if (dev_alloc(f->ctx, (void**)&f->dser, f->n * sizeof(double))) {
        set_last_error("field: cudaMalloc of the serial buffer failed");
        return 1;
    }
    CF_CUDA(cudaMemsetAsync(f->dser, 0, f->n * sizeof(double), f->ctx->stream));
    f->clean_Kx = 0; f->clean_Kz = 0;  // all zero
    return 0;
}
int field_serial_output(cfgpu_field f) {
    CF_TRY(field_ser_alloc(f));
    f->layout = 0;
    return 0;
}
static int tile_alloc(cfgpu_field f, const TileGeom& g) {
    const long long need = g.ntiles() * (long long)f->Ny * g.TM * 2 * f->Nd;
    if (!f->dtile || f->ntile < need) {
        if (f->dtile) { dev_free(f->ctx, f->dtile); f->dtile = nullptr; }
        if (dev_alloc(f->ctx, (void**)&f->dtile, need * sizeof(double))) {
            set_last_error("field: cudaMalloc of the tile-major buffer failed");
            return 1;
        }
        f->ntile = need;
        CF_CUDA(cudaMemsetAsync(f->dtile, 0, need * sizeof(double), f->ctx->stream));
    }
    f->tg = g;
    return 0;
}
static bool layout_trace() {
    static const bool on = getenv("CFGPU_LAYOUT_TRACE") != nullptr;  // report every layout conversion (they should not
    return on;                                                      // occur inside the time-stepping loop)
}
int field_serial(cfgpu_field f) {
    CF_TRY(field_ser_alloc(f));
    if (f->layout == 0) return 0;
    if (layout_trace()) fprintf(stderr, "[cfgpu] field %p: tile-major -> serial\n", (void*)f);
    const TileGeom& g = f->tg;
    if (f->tile_outside_zero) {
        if (!(f->clean_Kx >= 0 && f->clean_Kx <= g.Kx && f->clean_Kz >= 0 && f->clean_Kz <= g.Kz))
            CF_CUDA(cudaMemsetAsync(f->dser, 0, f->n * sizeof(double), f->ctx->stream));
        f->clean_Kx = g.Kx; f->clean_Kz = g.Kz;
    }
    CF_TRY(tile_convert_launch(f->dser, f->dtile, f->Nx, f->Ny, f->Nz, f->Nd, g.Kx, g.Kz, g.x0, g.nq(), g.TM, 1, f->ctx->stream));
    f->layout = 0;
    return 0;
}
int field_tile(cfgpu_field f, const TileGeom& g) {
    if (f->layout == 1 && f->tg.same(g)) return 0;
    if (f->layout == 0 && !f->dser) {  // never written: all zero in any layout
        CF_TRY(tile_alloc(f, g));
        CF_CUDA(cudaMemsetAsync(f->dtile, 0, f->ntile * sizeof(double), f->ctx->stream));
        f->layout = 1;
        f->tile_outside_zero = true;
        return 0;
    }
    CF_TRY(field_serial(f));
    CF_TRY(tile_alloc(f, g));
    if (layout_trace()) fprintf(stderr, "[cfgpu] field %p: serial -> tile-major\n", (void*)f);
    CF_TRY(tile_convert_launch(f->dser, f->dtile, f->Nx, f->Ny, f->Nz, f->Nd, g.Kx, g.Kz, g.x0, g.nq(), g.TM, 0, f->ctx->stream));
    f->layout = 1;
    f->tile_outside_zero = false;  // outside the box the serial buffer stays authoritative
    return 0;
}
int field_tile_output(cfgpu_field f, const TileGeom& g, bool outside_zero) {
    if (!(f->layout == 1 && f->tg.same(g))) {
        // the serial buffer keeps defining the field outside the box unless the caller overwrites that with zeros
        if (f->layout == 1 && !outside_zero) CF_TRY(field_serial(f));
        CF_TRY(tile_alloc(f, g));
        f->tile_outside_zero = false;
    }
    f->layout = 1;
    if (outside_zero || !f->dser) f->tile_outside_zero = true;  // no serial buffer: nothing but zeros outside the box
    return 0;
}
}  // namespace cfgpu

extern "C" {
int cfgpu_field_allgather(cfgpu_field f) {
    CF_TRY(field_serial(f));
    cfgpu_ctx ctx = f->ctx;
    Comm& cm = ctx->comm;
    if (cm.nranks == 1) return 0;
    CF_ARG(f->xzstate == CFGPU_SPECTRAL, "cfgpu_field_allgather: field must be xz-spectral");
    CF_ARG(f->padded, "cfgpu_field_allgather: the kx-slab partition is defined on the de-aliased box; the field must be padded");
    const int Kx = f->Nx / 3 - 1, Kz = f->Nz / 3 - 1, nmx = 2 * Kx + 1, nkz = Kz + 1;
    const size_t rows = (size_t)f->Nd * f->Ny;
    // a workspace of its own: ws_S may be mapped by the other ranks (peer-memory all-to-all) and must not be replaced here
    CF_TRY(ws_reserve(ctx->ws_G, rows * nmx * nkz * 2 * sizeof(double)));
    double2* S = reinterpret_cast<double2*>(ctx->ws_G.ptr);
    int x0, x1;
    part_range(nmx, cm.nranks, cm.rank, x0, x1);
    CF_TRY(rows_pack_launch(f->dser, reinterpret_cast<double*>(S + rows * nkz * x0), f->Nx, f->Nz, (int)rows, Kx, Kz, x0, x1, 0, ctx->stream));
    std::vector<ExchangeMsg> msgs;
    for (int r = 0; r < cm.nranks; ++r) {
        if (r == cm.rank) continue;
        int a, b;
        part_range(nmx, cm.nranks, r, a, b);
        msgs.push_back({r, S + rows * nkz * x0, (long long)(rows * nkz * (x1 - x0) * 16), S + rows * nkz * a, (long long)(rows * nkz * (b - a) * 16)});
    }
    CF_TRY(comm_exchange(cm, msgs.data(), (int)msgs.size(), ctx->stream));
    for (int r = 0; r < cm.nranks; ++r) {
        if (r == cm.rank) continue;
        int a, b;
        part_range(nmx, cm.nranks, r, a, b);
        CF_TRY(rows_pack_launch(f->dser, reinterpret_cast<double*>(S + rows * nkz * a), f->Nx, f->Nz, (int)rows, Kx, Kz, a, b, 1, ctx->stream));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ fields
int cfgpu_field_create(cfgpu_ctx ctx, int Nx, int Ny, int Nz, int Nd, double Lx, double Lz, double a, double b,
                       cfgpu_field* out) {
    CF_ARG(ctx && out, "cfgpu_field_create: null argument");
    CF_ARG(Nx > 0 && Ny > 0 && Nz > 0 && Nd > 0, "cfgpu_field_create: bad dimensions");
    cfgpu_field f = new cfgpu_field_s();
    f->ctx = ctx;
    f->Nx = Nx; f->Ny = Ny; f->Nz = Nz; f->Nd = Nd;
    f->Lx = Lx; f->Lz = Lz; f->a = a; f->b = b;
    f->n = (long long)Nx * Ny * f->Nzpad() * Nd;
    f->clean_Kx = 0; f->clean_Kz = 0;  // all zero; the serial buffer is allocated on first use (field_ser_alloc)
    *out = f;
    return 0;
}
int cfgpu_field_destroy(cfgpu_field f) {
    if (!f) return 0;
    dev_free(f->ctx, f->dser);
    dev_free(f->ctx, f->dtile);
    delete f;
    return 0;
}
int cfgpu_host_alloc(void** p, unsigned long long bytes) {
    CF_ARG(p, "cfgpu_host_alloc: null argument");
    CF_CUDA(cudaMallocHost(p, (size_t)(bytes ? bytes : 8)));
    return 0;
}
int cfgpu_host_free(void* p) {
    if (p) CF_CUDA(cudaFreeHost(p));
    return 0;
}
int cfgpu_field_upload(cfgpu_field f, const double* h, int xz, int y) {
    CF_TRY(field_serial_output(f));  // everything is overwritten
    CF_CUDA(cudaMemcpyAsync(f->dser, h, f->n * sizeof(double), cudaMemcpyHostToDevice, f->ctx->stream));
    CF_CUDA(cudaStreamSynchronize(f->ctx->stream));
    f->xzstate = xz; f->ystate = y;
    f->clean_Kx = f->clean_Kz = -1;
    return 0;
}
int cfgpu_field_download(cfgpu_field f, double* h) {
    CF_TRY(field_serial(f));
    CF_CUDA(cudaMemcpyAsync(h, f->dser, f->n * sizeof(double), cudaMemcpyDeviceToHost, f->ctx->stream));
    CF_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return 0;
}
// pitched copies of the retained box (rows of this rank) between a host array and the device field, both in the
// reference layout [i][my][mx][mz] complex
static int box_copy(cfgpu_field f, double* h, bool to_device) {
    cfgpu_ctx ctx = f->ctx;
    const int Kx = f->Nx / 3 - 1, Kz = f->Nz / 3 - 1, nmx = 2 * Kx + 1;
    int x0 = 0, x1 = nmx;
    if (ctx->comm.nranks > 1) part_range(nmx, ctx->comm.nranks, ctx->comm.rank, x0, x1);
    const size_t pitch = (size_t)f->Mz() * 16;
    for (int part = 0; part < 2; ++part) {
        // part 0: kx >= 0 (mxi = mx <= Kx); part 1: kx < 0 (mxi > Kx, mx = Nx - nmx + mxi)
        const int a = part == 0 ? x0 : (x0 > Kx + 1 ? x0 : Kx + 1);
        const int b = part == 0 ? (x1 < Kx + 1 ? x1 : Kx + 1) : x1;
        if (b <= a) continue;
        const int mx0 = part == 0 ? a : f->Nx - nmx + a;
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof p);
        cudaPitchedPtr hp = make_cudaPitchedPtr((void*)h, pitch, pitch, (size_t)f->Nx);
        cudaPitchedPtr dp = make_cudaPitchedPtr((void*)f->dser, pitch, pitch, (size_t)f->Nx);
        p.srcPtr = to_device ? hp : dp;
        p.dstPtr = to_device ? dp : hp;
        p.srcPos = make_cudaPos(0, (size_t)mx0, 0);
        p.dstPos = p.srcPos;
        p.extent = make_cudaExtent((size_t)(Kz + 1) * 16, (size_t)(b - a), (size_t)f->Nd * f->Ny);
        p.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        CF_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
    }
    return 0;
}
int cfgpu_field_upload_padded(cfgpu_field f, const double* h, int ystate) {
    CF_TRY(field_serial_output(f));  // the box is overwritten, everything else zeroed (below, unless the serial buffer is known clean)
    const int Kx = f->Nx / 3 - 1, Kz = f->Nz / 3 - 1;
    CF_ARG(Kx >= 0 && Kz >= 0, "cfgpu_field_upload_padded: grid too small");
    if (!(f->clean_Kx >= 0 && f->clean_Kx <= Kx && f->clean_Kz >= 0 && f->clean_Kz <= Kz))
        CF_CUDA(cudaMemsetAsync(f->dser, 0, f->n * sizeof(double), f->ctx->stream));
    CF_TRY(box_copy(f, const_cast<double*>(h), true));
    CF_CUDA(cudaStreamSynchronize(f->ctx->stream));
    f->xzstate = CFGPU_SPECTRAL; f->ystate = ystate; f->padded = 1;
    f->clean_Kx = Kx; f->clean_Kz = Kz;
    return 0;
}
int cfgpu_field_download_padded(cfgpu_field f, double* h) {
    CF_ARG(f->xzstate == CFGPU_SPECTRAL, "cfgpu_field_download_padded: field must be xz-spectral");
    CF_TRY(field_serial(f));
    CF_TRY(box_copy(f, h, false));
    CF_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return 0;
}
static bool same_shape(cfgpu_field a, cfgpu_field b) {
    return a->Nx == b->Nx && a->Ny == b->Ny && a->Nz == b->Nz && a->Nd == b->Nd;
}
int cfgpu_field_copy(cfgpu_field dst, cfgpu_field src) {
    CF_ARG(same_shape(dst, src), "cfgpu_field_copy: shape mismatch");
    if (dst == src) return 0;
    if (src->layout == 1) {
        // the retained box lives in the tile buffer; outside it the field is zero (by flag, or because the serial buffer is
        // known to be clean there) or what the serial buffer holds
        const bool zero_out = src->tile_outside_zero || (src->clean_Kx >= 0 && src->clean_Kx <= src->tg.Kx &&
                                                         src->clean_Kz >= 0 && src->clean_Kz <= src->tg.Kz);
        CF_TRY(field_tile_output(dst, src->tg, zero_out));
        CF_CUDA(cudaMemcpyAsync(dst->dtile, src->dtile, src->tg.ntiles() * src->tile_stride() * sizeof(double),
                                cudaMemcpyDeviceToDevice, dst->ctx->stream));
        dst->tile_outside_zero = zero_out;
        if (!zero_out) {
            CF_TRY(field_ser_alloc(dst));
            CF_CUDA(cudaMemcpyAsync(dst->dser, src->dser, src->n * sizeof(double), cudaMemcpyDeviceToDevice, dst->ctx->stream));
            dst->clean_Kx = src->clean_Kx; dst->clean_Kz = src->clean_Kz;
        }
    } else if (!src->dser) {  // never written: all zero
        CF_TRY(cfgpu_field_zero(dst));
    } else {
        CF_TRY(field_serial_output(dst));
        CF_CUDA(cudaMemcpyAsync(dst->dser, src->dser, src->n * sizeof(double), cudaMemcpyDeviceToDevice, dst->ctx->stream));
        dst->clean_Kx = src->clean_Kx; dst->clean_Kz = src->clean_Kz;
    }
    dst->xzstate = src->xzstate; dst->ystate = src->ystate; dst->padded = src->padded;
    dst->Lx = src->Lx; dst->Lz = src->Lz; dst->a = src->a; dst->b = src->b;
    return 0;
}
// dst component jd <- src component js (FlowField::operator[](int), flowfield.cpp:1532-1560, and its inverse)
int cfgpu_field_copy_component(cfgpu_field dst, int jd, cfgpu_field src, int js) {
    CF_ARG(dst && src && dst != src, "cfgpu_field_copy_component: bad argument");
    CF_ARG(dst->Nx == src->Nx && dst->Ny == src->Ny && dst->Nz == src->Nz, "cfgpu_field_copy_component: grid mismatch");
    CF_ARG(jd >= 0 && jd < dst->Nd && js >= 0 && js < src->Nd, "cfgpu_field_copy_component: component out of range");
    CF_TRY(field_serial(src));
    CF_TRY(field_serial(dst));
    CF_CUDA(cudaMemcpyAsync(dst->dser + jd * dst->compstride(), src->dser + js * src->compstride(), src->compstride() * sizeof(double),
                            cudaMemcpyDeviceToDevice, dst->ctx->stream));
    auto mrg = [](int p, int q) { return (p < 0 || q < 0) ? -1 : (p > q ? p : q); };
    dst->clean_Kx = mrg(dst->clean_Kx, src->clean_Kx); dst->clean_Kz = mrg(dst->clean_Kz, src->clean_Kz);
    return 0;
}
int cfgpu_field_swap(cfgpu_field a, cfgpu_field b) {
    CF_ARG(same_shape(a, b), "cfgpu_field_swap: shape mismatch");
    std::swap(a->dser, b->dser);
    std::swap(a->dtile, b->dtile); std::swap(a->ntile, b->ntile); std::swap(a->layout, b->layout);
    std::swap(a->tile_outside_zero, b->tile_outside_zero); std::swap(a->tg, b->tg);
    std::swap(a->xzstate, b->xzstate); std::swap(a->ystate, b->ystate); std::swap(a->padded, b->padded);
    std::swap(a->clean_Kx, b->clean_Kx); std::swap(a->clean_Kz, b->clean_Kz);
    return 0;
}
int cfgpu_field_zero(cfgpu_field f) {
    f->layout = 0;
    if (f->dser) CF_CUDA(cudaMemsetAsync(f->dser, 0, f->n * sizeof(double), f->ctx->stream));
    f->clean_Kx = 0; f->clean_Kz = 0;
    return 0;
}
int cfgpu_field_set_state(cfgpu_field f, int xz, int y) { f->xzstate = xz; f->ystate = y; return 0; }
int cfgpu_field_get_state(cfgpu_field f, int* xz, int* y) { *xz = f->xzstate; *y = f->ystate; return 0; }
int cfgpu_field_set_padded(cfgpu_field f, int p) { f->padded = p; return 0; }
int cfgpu_field_get_padded(cfgpu_field f, int* p) { *p = f->padded; return 0; }
int cfgpu_field_device_ptr(cfgpu_field f, double** d, long long* n) { CF_TRY(field_serial(f)); *d = f->dser; if (n) *n = f->n; return 0; }

int cfgpu_field_axpby(cfgpu_field y, double a, cfgpu_field x, double b, cfgpu_field z) {
    CF_ARG(same_shape(y, x) && (!z || same_shape(y, z)), "cfgpu_field_axpby: shape mismatch");
    // Operands that live tile-major and are zero outside their box (nonlinear terms, linear terms: what the RK / CNAB
    // steppers combine) are combined in place in that layout: no conversion, only the retained modes are touched.
    auto tiled0 = [](cfgpu_field f) { return f->layout == 1 && f->tile_outside_zero; };
    if (tiled0(x) && (!z || (tiled0(z) && z->tg.same(x->tg))) && y != x && y != z) {
        CF_TRY(field_tile(y, x->tg));  // no-op if y is already in this layout; a never-written y becomes an all-zero tile field
        const long n = (long)(x->tg.ntiles() * x->tile_stride());
        CF_TRY(axpby_launch(y->dtile, a, x->dtile, b, z ? z->dtile : nullptr, n, y->ctx->stream));
        return 0;
    }
    CF_TRY(field_serial(y)); CF_TRY(field_serial(x)); if (z) CF_TRY(field_serial(z));
    CF_TRY(axpby_launch(y->dser, a, x->dser, b, z ? z->dser : nullptr, (long)y->n, y->ctx->stream));
    auto mrg = [](int p, int q) { return (p < 0 || q < 0) ? -1 : (p > q ? p : q); };
    y->clean_Kx = mrg(y->clean_Kx, x->clean_Kx); y->clean_Kz = mrg(y->clean_Kz, x->clean_Kz);
    if (z) { y->clean_Kx = mrg(y->clean_Kx, z->clean_Kx); y->clean_Kz = mrg(y->clean_Kz, z->clean_Kz); }
    return 0;
}
int cfgpu_field_scale(cfgpu_field y, double s) {
    if (y->layout == 1) {
        // the box lives in the tile buffer; whatever the field holds outside it is in the serial buffer
        CF_TRY(scale_launch(y->dtile, s, (long)(y->tg.ntiles() * y->tile_stride()), y->ctx->stream));
        if (!y->tile_outside_zero && y->dser) CF_TRY(scale_launch(y->dser, s, (long)y->n, y->ctx->stream));
        return 0;
    }
    if (!y->dser) return 0;  // never written: all zero
    CF_TRY(scale_launch(y->dser, s, (long)y->n, y->ctx->stream));
    return 0;
}

int cfgpu_field_get_profile(cfgpu_field f, int mx, int mz, int i, double* out_h) {
    CF_ARG(mx >= 0 && mx < f->Nx && mz >= 0 && mz < f->Mz() && i >= 0 && i < f->Nd, "cfgpu_field_get_profile: index");
    CF_TRY(field_serial(f));
    cfgpu_ctx ctx = f->ctx;
    CF_TRY(ws_reserve(ctx->ws_red, 1 << 20));
    const long off0 = (long)i * f->Ny * f->Nx * f->Mz() + mz + (long)f->Mz() * mx;
    CF_TRY(profile_get_launch(f->dser, off0, (long)f->Nx * f->Mz(), f->Ny, ctx->ws_red.ptr, ctx->stream));
    CF_CUDA(cudaMemcpyAsync(out_h, ctx->ws_red.ptr, 2 * f->Ny * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cfgpu_field_add_profile(cfgpu_field f, int mx, int mz, int i, const double* in_h, double scale) {
    CF_ARG(mx >= 0 && mx < f->Nx && mz >= 0 && mz < f->Mz() && i >= 0 && i < f->Nd, "cfgpu_field_add_profile: index");
    CF_TRY(field_serial(f));
    cfgpu_ctx ctx = f->ctx;
    CF_TRY(ws_reserve(ctx->ws_red, 1 << 20));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    CF_CUDA(cudaMemcpyAsync(ctx->ws_red.ptr, in_h, 2 * f->Ny * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const long off0 = (long)i * f->Ny * f->Nx * f->Mz() + mz + (long)f->Mz() * mx;
    CF_TRY(profile_add_launch(f->dser, off0, (long)f->Nx * f->Mz(), f->Ny, ctx->ws_red.ptr, scale, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// FlowField::operator*=(const FieldSymmetry&) (flowfield.cpp:1274-1433); spectral state, in place
int cfgpu_field_symmetry(cfgpu_field f, int s, int sx, int sy, int sz, double ax, double az) {
    CF_ARG(f && (s == 1 || s == -1) && (sx == 1 || sx == -1) && (sy == 1 || sy == -1) && (sz == 1 || sz == -1), "cfgpu_field_symmetry: signs must be +-1");
    CF_ARG(f->xzstate == CFGPU_SPECTRAL && f->ystate == CFGPU_SPECTRAL, "cfgpu_field_symmetry: field must be spectral");
    CF_ARG(f->ctx->comm.nranks == 1, "cfgpu_field_symmetry: single-GPU call");
    CF_ARG(f->Nd == 1 || f->Nd == 3, "cfgpu_field_symmetry: scalar and 3-vector fields (tensor sign rules are not carried)");
    CF_TRY(field_serial(f));
    const int Kxhi = f->padded ? f->Nx / 3 - 1 : f->Nx / 2, Kxlo = f->padded ? -(f->Nx / 3 - 1) : f->Nx / 2 + 1 - f->Nx;
    const int Kz = f->padded ? f->Nz / 3 - 1 : f->Nz / 2;
    CF_TRY(symmetry_launch(f->dser, f->Nx, f->Ny, f->Nz, f->Nd, Kxlo, Kxhi, Kz, s, sx, sy, sz, ax, az, f->ctx->stream));
    return 0;
}
int cfgpu_field_zero_padded_modes(cfgpu_field f) {
    CF_TRY(field_serial(f));
    const int Kx = f->Nx / 3 - 1, Kz = f->Nz / 3 - 1;  // flowfield.h:578-584
    CF_TRY(zero_padded_launch(f->dser, f->Nx, f->Ny, f->Nz, f->Nd, Kx, Kz, f->ctx->stream));
    f->padded = 1;
    f->clean_Kx = Kx; f->clean_Kz = Kz;
    return 0;
}

// ------------------------------------------------------------------------------------------------ generic y transform
static int y_transform(cfgpu_field f, int mode) {
    CF_TRY(field_serial(f));
    const YPlan* pl;
    CF_TRY(get_yplan(f->ctx, f->Ny, f->a, f->b, &pl));
    const long ncols = f->rowstride();
    for (int i0 = 0; i0 < f->Nd; i0 += YG_MAXJOB) {
        YGemmParams p;
        memset(&p, 0, sizeof p);
        p.N = f->Ny; p.mode = mode;
        p.fft = yfft_plan(f->ctx, f->Ny); p.fft_half = yfft_half_plan(f->ctx, f->Ny); p.ya = f->a; p.yb = f->b;
        if (mode == 0) {
            p.M = p.M2 = pl->Nh; p.K1 = pl->Ne; p.K2 = pl->No; p.K1p = pl->invK1p; p.K2p = pl->invK2p;
            p.A1[0] = pl->Ce; p.A2[0] = pl->Co; p.sgn[0] = 1.0;
        } else {
            p.M = pl->Ne; p.M2 = pl->No; p.K1 = p.K2 = pl->Nh; p.K1p = p.K2p = pl->fwdKp;
            p.A1[0] = pl->Fe; p.A2[0] = pl->Fo; p.sgn[0] = 1.0;
        }
        p.ncols = ncols;
        p.in_runlen = 1; p.in_runstart = nullptr; p.in_ld = ncols;
        p.out_runlen = 1; p.out_runstart = nullptr; p.out_ld = ncols;
        p.njobs = 0;
        for (int i = i0; i < f->Nd && p.njobs < YG_MAXJOB; ++i) {
            YGemmJob& j = p.job[p.njobs++];
            j.in = f->dser + i * f->compstride();
            j.out[0] = f->dser + i * f->compstride();
            j.nmat = 1; j.mat0 = 0;
        }
        CF_TRY(ygemm_launch(p, f->ctx->stream));
    }
    f->clean_Kx = f->clean_Kz = -1;
    return 0;
}
int cfgpu_field_make_physical_y(cfgpu_field f) {
    if (f->ystate == CFGPU_PHYSICAL) return 0;
    if (f->Ny >= 2) CF_TRY(y_transform(f, 0));
    f->ystate = CFGPU_PHYSICAL;
    return 0;
}
int cfgpu_field_make_spectral_y(cfgpu_field f) {
    if (f->ystate == CFGPU_SPECTRAL) return 0;
    if (f->Ny >= 2) CF_TRY(y_transform(f, 1));
    f->ystate = CFGPU_SPECTRAL;
    return 0;
}

// generic xz transforms live in cfgpu_xzgen.cu
int cfgpu_xz_generic(cfgpu_field f, int to_physical);
int cfgpu_field_make_physical_xz(cfgpu_field f) {
    CF_TRY(field_serial(f));
    if (f->xzstate == CFGPU_PHYSICAL) return 0;
    CF_TRY(cfgpu_xz_generic(f, 1));
    f->xzstate = CFGPU_PHYSICAL;
    f->clean_Kx = f->clean_Kz = -1;
    return 0;
}
int cfgpu_field_make_spectral_xz(cfgpu_field f) {
    CF_TRY(field_serial(f));
    if (f->xzstate == CFGPU_SPECTRAL) return 0;
    CF_TRY(cfgpu_xz_generic(f, 0));
    f->xzstate = CFGPU_SPECTRAL;
    f->clean_Kx = f->clean_Kz = -1;
    return 0;
}
int cfgpu_field_make_physical(cfgpu_field f) {
    CF_TRY(cfgpu_field_make_physical_y(f));
    return cfgpu_field_make_physical_xz(f);
}
int cfgpu_field_make_spectral(cfgpu_field f) {
    CF_TRY(cfgpu_field_make_spectral_xz(f));
    return cfgpu_field_make_spectral_y(f);
}

// ------------------------------------------------------------------------------------------------ norms
static int l2form(cfgpu_field u, cfgpu_field v, int mode, int normalize, bool padded, double* out_h, bool skip_kx0 = false, bool cheby = false) {
    CF_TRY(field_serial(u)); if (v) CF_TRY(field_serial(v));
    CF_ARG(u->xzstate == CFGPU_SPECTRAL && u->ystate == CFGPU_SPECTRAL, "L2 norm: field must be spectral");
    cfgpu_ctx ctx = u->ctx;
    const YPlan* pl;
    CF_TRY(get_yplan(ctx, u->Ny, u->a, u->b, &pl));
    {   // one partial sum per (tile of <= 16 modes, component)
        const size_t modes = (size_t)u->Nx * u->Mz() * u->Nd;
        const size_t need = (modes + 64) * sizeof(double);  // bound for one mode per tile
        CF_TRY(ws_reserve(ctx->ws_red, need > ((size_t)1 << 20) ? need : ((size_t)1 << 20)));
    }
    double scale = 1.0;
    if (!normalize) scale = (u->b - u->a) * u->Lx * u->Lz;
    const int Kx = u->Nx / 3 - 1, Kz = u->Nz / 3 - 1;
    double* out_dev = ctx->ws_red.ptr;
    double* partial = ctx->ws_red.ptr + 8;
    const size_t cap = ctx->ws_red.bytes / sizeof(double) - 8;
    int x0 = 0, x1 = padded ? 2 * Kx + 1 : u->Nx;
    if (ctx->comm.nranks > 1) {
        CF_ARG(padded, "L2 norm: multi-GPU norms need de-aliased (padded) fields");
        part_range(2 * Kx + 1, ctx->comm.nranks, ctx->comm.rank, x0, x1);
    }
    if (skip_kx0 && x0 < 1) x0 = 1;  // row 0 is kx = 0 in both enumerations
    CF_TRY(l2form_launch(u->dser, v ? v->dser : nullptr, mode, cheby ? pl->Wcheb : pl->Wgram, u->Ny, u->Nx, u->Nz, u->Nd, Kx, Kz, padded ? 0 : 1, x0, x1, scale,
                         partial, cap, out_dev, ctx->stream));
    CF_TRY(comm_allreduce(ctx->comm, out_dev, 1, 0, ctx->stream));
    CF_CUDA(cudaMemcpyAsync(out_h, out_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int cfgpu_l2norm2(cfgpu_field u, int normalize, double* out_h) { return l2form(u, nullptr, 0, normalize, u->padded != 0, out_h); }
// chebyNorm2 / chebyDist2 / chebyInnerProduct (diffops.cpp:259-350): mode 0 / 1 / 2, v NULL for the norm
int cfgpu_chebyform(cfgpu_field u, cfgpu_field v, int mode, int normalize, double* out_h) {
    CF_ARG(u && out_h && mode >= 0 && mode <= 2 && (mode == 0 || v), "cfgpu_chebyform: bad argument");
    return l2form(u, mode == 0 ? nullptr : v, mode, normalize, u->padded != 0, out_h, false, true);
}

// L2Norm2 / L2Dist2 / L2InnerProduct restricted to |kx| <= kxmax, kz <= kzmax (diffops.cpp:543-700); cz = 0 sums the stored
// modes without the factor 2 for kz > 0 (the convention of divNorm2, diffops.cpp:91-116)
int cfgpu_l2form_box(cfgpu_field u, cfgpu_field v, int mode, int kxmax, int kzmax, int cz, int normalize, double* out_h) {
    CF_ARG(u && out_h && mode >= 0 && mode <= 2 && (mode == 0 || v), "cfgpu_l2form_box: bad argument");
    CF_ARG(!v || same_shape(u, v), "cfgpu_l2form_box: shape mismatch");
    CF_TRY(field_serial(u)); if (v) CF_TRY(field_serial(v));
    CF_ARG(u->xzstate == CFGPU_SPECTRAL && u->ystate == CFGPU_SPECTRAL, "L2 norm: field must be spectral");
    cfgpu_ctx ctx = u->ctx;
    CF_ARG(ctx->comm.nranks == 1, "cfgpu_l2form_box: single-GPU call (use the default norms in multi-GPU runs)");
    const YPlan* pl;
    CF_TRY(get_yplan(ctx, u->Ny, u->a, u->b, &pl));
    CF_TRY(ws_reserve(ctx->ws_red, 1 << 20));
    const double scale = normalize ? 1.0 : (u->b - u->a) * u->Lx * u->Lz;
    const bool full = kxmax >= u->Nx / 2 && kzmax >= u->Nz / 2;
    if (!full) CF_ARG(kxmax >= 0 && kxmax < (u->Nx + 1) / 2 && kzmax >= 0 && kzmax <= u->Nz / 2, "cfgpu_l2form_box: box out of range");
    double* out_dev = ctx->ws_red.ptr;
    CF_TRY(l2form_launch(u->dser, v ? v->dser : nullptr, mode, pl->Wgram, u->Ny, u->Nx, u->Nz, u->Nd, kxmax, kzmax, full ? 1 : 0, 0,
                         full ? u->Nx : 2 * kxmax + 1, scale, ctx->ws_red.ptr + 8, ctx->ws_red.bytes / sizeof(double) - 8, out_dev, ctx->stream,
                         cz ? 2.0 : 1.0));
    CF_CUDA(cudaMemcpyAsync(out_h, out_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// bcNorm2 / bcDist2 (diffops.cpp:18-83): sum over components and all modes of |f(a)|^2 + |f(b)|^2
int cfgpu_bcnorm2(cfgpu_field u, cfgpu_field v, int normalize, double* out_h) {
    CF_ARG(u && out_h && (!v || same_shape(u, v)), "cfgpu_bcnorm2: bad argument");
    CF_TRY(field_serial(u)); if (v) CF_TRY(field_serial(v));
    CF_ARG(u->xzstate == CFGPU_SPECTRAL, "cfgpu_bcnorm2: field must be xz-spectral");
    cfgpu_ctx ctx = u->ctx;
    CF_ARG(ctx->comm.nranks == 1, "cfgpu_bcnorm2: single-GPU call");
    CF_TRY(ws_reserve(ctx->ws_red, 1 << 20));
    CF_TRY(bcnorm2_launch(u->dser, v ? v->dser : nullptr, u->Nx, u->Ny, u->Nz, u->Nd, u->ystate == CFGPU_SPECTRAL, ctx->ws_red.ptr + 8,
                          ctx->ws_red.bytes / sizeof(double) - 8, ctx->ws_red.ptr, ctx->stream));
    double r = 0;
    CF_CUDA(cudaMemcpyAsync(&r, ctx->ws_red.ptr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CF_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_h = normalize ? r : r * u->Lx * u->Lz;
    return 0;
}
// PoissonSolver::solve (poissonsolver.cpp:146-202)
int cfgpu_poisson_solve(cfgpu_field u, cfgpu_field f, cfgpu_field bc) {
    CF_ARG(u && f && u != f && same_shape(u, f) && (!bc || same_shape(bc, f)), "cfgpu_poisson_solve: bad argument");
    CF_ARG(f->xzstate == CFGPU_SPECTRAL && f->ystate == CFGPU_SPECTRAL, "cfgpu_poisson_solve: f must be spectral");
    CF_ARG(!bc || (bc->xzstate == CFGPU_SPECTRAL && bc->ystate == CFGPU_SPECTRAL), "cfgpu_poisson_solve: bc must be spectral");
    CF_ARG(u->ctx->comm.nranks == 1, "cfgpu_poisson_solve: single-GPU call");
    CF_TRY(field_serial(f)); if (bc) CF_TRY(field_serial(bc));
    CF_TRY(field_serial_output(u));
    CF_TRY(poisson_launch(f->Nx, f->Ny, f->Nz, f->Nd, f->Lx, f->Lz, f->a, f->b, f->dser, bc ? bc->dser : nullptr, u->dser, u->ctx->stream));
    u->xzstate = u->ystate = CFGPU_SPECTRAL;
    u->padded = 0;
    u->clean_Kx = u->clean_Kz = -1;
    return 0;
}
// PressureSolver::solve step II (poissonsolver.cpp:352-431): g = the homogeneous correction for dp/dy = nu v_yy at the walls
int cfgpu_pressure_neumann(cfgpu_field g, cfgpu_field p, cfgpu_field u, double nu) {
    CF_ARG(g && p && u && g != p && g->Nd == 1 && p->Nd == 1 && u->Nd == 3, "cfgpu_pressure_neumann: g, p scalar fields, u a 3-vector field");
    CF_ARG(g->Nx == p->Nx && g->Ny == p->Ny && g->Nz == p->Nz && u->Nx == p->Nx && u->Ny == p->Ny && u->Nz == p->Nz,
           "cfgpu_pressure_neumann: grid mismatch");
    CF_ARG(p->xzstate == CFGPU_SPECTRAL && p->ystate == CFGPU_SPECTRAL && u->xzstate == CFGPU_SPECTRAL && u->ystate == CFGPU_SPECTRAL,
           "cfgpu_pressure_neumann: p and u must be spectral");
    CF_ARG(p->ctx->comm.nranks == 1, "cfgpu_pressure_neumann: single-GPU call");
    CF_TRY(field_serial(p)); CF_TRY(field_serial(u));
    CF_TRY(field_serial_output(g));
    const FieldGeom fg{p->Nx, p->Ny, p->Nz, p->Lx, p->Lz, p->a, p->b};
    CF_TRY(pressure_neumann_launch(p->dser, u->dser + u->compstride(), nu, fg, g->dser, p->ctx->stream));
    g->xzstate = CFGPU_SPECTRAL; g->ystate = CFGPU_PHYSICAL;
    g->padded = 0;
    g->clean_Kx = g->clean_Kz = -1;
    return 0;
}
// generic linear differential operator on a spectral field (xdiff/ydiff/zdiff/grad/lapl/curl/div of diffops.cpp)
int cfgpu_field_diffop(cfgpu_field out, cfgpu_field in, int nterms, const int* out_c, const int* in_c, const int* nx, const int* ny,
                       const int* nz, const double* coef) {
    CF_ARG(out && in && out != in && nterms >= 1 && nterms <= DIFF_MAXTERMS, "cfgpu_field_diffop: bad argument");
    CF_ARG(out->Nx == in->Nx && out->Ny == in->Ny && out->Nz == in->Nz, "cfgpu_field_diffop: grid mismatch");
    CF_ARG(in->xzstate == CFGPU_SPECTRAL && in->ystate == CFGPU_SPECTRAL, "cfgpu_field_diffop: input must be spectral");
    CF_TRY(field_serial(in));
    CF_TRY(field_serial_output(out));
    DiffOpParams p;
    memset(&p, 0, sizeof p);
    p.g = FieldGeom{in->Nx, in->Ny, in->Nz, in->Lx, in->Lz, in->a, in->b};
    p.nout = out->Nd; p.nterms = nterms;
    for (int k = 0; k < nterms; ++k) {
        CF_ARG(in_c[k] >= 0 && in_c[k] < in->Nd && out_c[k] >= 0 && out_c[k] < out->Nd, "cfgpu_field_diffop: component out of range");
        p.t[k] = DiffTerm{out_c[k], in_c[k], nx[k], ny[k], nz[k], coef[k]};
    }
    {   // components without a term must come out zero
        bool has[64] = {false};
        for (int k = 0; k < nterms; ++k) has[out_c[k]] = true;
        for (int c = 0; c < out->Nd; ++c)
            if (!has[c]) CF_CUDA(cudaMemsetAsync(out->dser + c * out->compstride(), 0, out->compstride() * sizeof(double), out->ctx->stream));
    }
    CF_TRY(diffop_launch(in->dser, out->dser, p, in->ctx->stream));
    out->xzstate = out->ystate = CFGPU_SPECTRAL;
    out->clean_Kx = in->clean_Kx; out->clean_Kz = in->clean_Kz;
    out->padded = in->padded;
    return 0;
}
// pointwise products of physical fields (cross/outer/dot/norm/norm2/energy of diffops.cpp); op: PointOp of diffops.cuh
int cfgpu_field_pointwise(int op, cfgpu_field out, cfgpu_field f, cfgpu_field g) {
    CF_ARG(out && f && op >= PW_CROSS && op <= PW_MUL, "cfgpu_field_pointwise: bad argument");
    const bool two = op == PW_CROSS || op == PW_OUTER || op == PW_DOT || op == PW_MUL;
    CF_ARG(!two || g, "cfgpu_field_pointwise: this operation needs two inputs");
    CF_ARG(f->xzstate == CFGPU_PHYSICAL && f->ystate == CFGPU_PHYSICAL && (!two || (g->xzstate == CFGPU_PHYSICAL && g->ystate == CFGPU_PHYSICAL)),
           "cfgpu_field_pointwise: inputs must be physical");
    const int need = op == PW_CROSS ? 3 : (op == PW_OUTER ? f->Nd * g->Nd : (op == PW_MUL ? f->Nd : 1));
    CF_ARG(out->Nd == need && out->Nx == f->Nx && out->Ny == f->Ny && out->Nz == f->Nz, "cfgpu_field_pointwise: output shape");
    CF_ARG(op != PW_CROSS || (f->Nd == 3 && g->Nd == 3), "cfgpu_field_pointwise: cross needs 3-component fields");
    CF_ARG(!(op == PW_DOT || op == PW_MUL) || f->Nd == g->Nd, "cfgpu_field_pointwise: component mismatch");
    CF_ARG(out != f && out != g, "cfgpu_field_pointwise: output aliases an input");
    CF_TRY(field_serial(f)); if (two) CF_TRY(field_serial(g));
    CF_TRY(field_serial_output(out));
    CF_TRY(pointwise_launch(op, f->dser, two ? g->dser : nullptr, out->dser, f->Nd, two ? g->Nd : 0, (long)f->compstride(), f->ctx->stream));
    out->xzstate = out->ystate = CFGPU_PHYSICAL;
    out->clean_Kx = out->clean_Kz = -1;
    out->padded = 0;
    return 0;
}
int cfgpu_l2norm2_3d(cfgpu_field u, int normalize, double* out_h) { return l2form(u, nullptr, 0, normalize, u->padded != 0, out_h, true); }
int cfgpu_l2dist2(cfgpu_field u, cfgpu_field v, int normalize, double* out_h) {
    CF_ARG(same_shape(u, v), "L2Dist2: shape mismatch");
    return l2form(u, v, 1, normalize, u->padded && v->padded, out_h);
}
int cfgpu_l2ip(cfgpu_field u, cfgpu_field v, int normalize, double* out_h) {
    CF_ARG(same_shape(u, v), "L2InnerProduct: shape mismatch");
    return l2form(u, v, 2, normalize, u->padded || v->padded, out_h);
}

}  // extern "C"
