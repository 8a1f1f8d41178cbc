// See comm.cuh.
#include "comm.cuh"

#include <cstdlib>
#include <vector>
#ifndef CF_EMU
#include <dlfcn.h>
#endif

namespace cfgpu {

#ifndef CF_EMU
namespace {
// the handful of NCCL entry points used, bound with dlsym (public API of nccl.h, NCCL >= 2.7)
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t*);
typedef int (*fn_CommInitRank)(void**, int, ncclUniqueId_t, int);
typedef int (*fn_CommDestroy)(void*);
typedef int (*fn_Group)(void);
typedef int (*fn_Send)(const void*, size_t, int /*datatype*/, int, void*, cudaStream_t);
typedef int (*fn_Recv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int /*op*/, void*, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);
constexpr int NCCL_INT8 = 0, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;

struct NcclApi {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_Group GroupStart = nullptr, GroupEnd = nullptr;
    fn_Send Send = nullptr;
    fn_Recv Recv = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.lib) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy torch (or anyone) already loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_last_error("NCCL not found (libnccl.so.2): multi-GPU runs need it");
        return 1;
    }
    g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
    g_nccl.GroupStart = (fn_Group)dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (fn_Group)dlsym(h, "ncclGroupEnd");
    g_nccl.Send = (fn_Send)dlsym(h, "ncclSend");
    g_nccl.Recv = (fn_Recv)dlsym(h, "ncclRecv");
    g_nccl.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.Send || !g_nccl.Recv ||
        !g_nccl.AllReduce) {
        set_last_error("libnccl.so.2 lacks a required symbol");
        return 1;
    }
    g_nccl.lib = h;
    return 0;
}
int nccl_check(int r, const char* what) {
    if (r == 0) return 0;
    set_last_error(std::string(what) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"));
    return 1;
}
#define CF_NCCL(expr) CF_TRY(nccl_check((expr), #expr))
}  // namespace
#endif

int comm_unique_id(void* out128) {
#ifdef CF_EMU
    set_last_error("NCCL is unavailable in the emulation build");
    return 1;
#else
    CF_TRY(nccl_load());
    ncclUniqueId_t id;
    CF_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return 0;
#endif
}

int comm_init_nccl(Comm& c, int rank, int nranks, const void* id128) {
#ifdef CF_EMU
    set_last_error("NCCL is unavailable in the emulation build");
    return 1;
#else
    if (nranks < 1 || nranks > COMM_MAXRANKS || rank < 0 || rank >= nranks) {
        set_last_error("comm_init: bad rank / world size");
        return 1;
    }
    CF_TRY(nccl_load());
    ncclUniqueId_t id;
    memcpy(&id, id128, 128);
    void* comm = nullptr;
    CF_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    c.rank = rank; c.nranks = nranks; c.nccl_comm = comm;
    return 0;
#endif
}

int comm_destroy(Comm& c) {
#ifndef CF_EMU
    if (c.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c.nccl_comm);
#endif
    c.nccl_comm = nullptr;
    if (c.d_bar) cudaFree(c.d_bar);
    c.d_bar = nullptr;
    return 0;
}

int comm_exchange(Comm& c, const ExchangeMsg* msgs, int nmsg, cudaStream_t stream) {
    for (int i = 0; i < nmsg; ++i) {
        if (msgs[i].peer == c.rank) {
            if (msgs[i].sendbytes != msgs[i].recvbytes) {
                set_last_error("comm_exchange: self message size mismatch");
                return 1;
            }
            if (msgs[i].sendbytes > 0 && msgs[i].send != msgs[i].recv)
                CF_CUDA(cudaMemcpyAsync(msgs[i].recv, msgs[i].send, (size_t)msgs[i].sendbytes, cudaMemcpyDeviceToDevice, stream));
        }
    }
    if (c.nranks == 1) return 0;
    if (c.ext_exchange) {
        std::vector<int> peer;
        std::vector<const void*> sp;
        std::vector<void*> rp;
        std::vector<long long> sb, rb;
        for (int i = 0; i < nmsg; ++i) {
            if (msgs[i].peer == c.rank) continue;
            peer.push_back(msgs[i].peer);
            sp.push_back(msgs[i].send); sb.push_back(msgs[i].sendbytes);
            rp.push_back(msgs[i].recv); rb.push_back(msgs[i].recvbytes);
        }
        CF_CUDA(cudaStreamSynchronize(stream));
        if (c.ext_exchange(c.ext_user, (int)peer.size(), peer.data(), sp.data(), sb.data(), rp.data(), rb.data()) != 0) {
            set_last_error("comm_exchange: external exchange callback failed");
            return 1;
        }
        return 0;
    }
#ifndef CF_EMU
    if (!c.nccl_comm) {
        set_last_error("comm_exchange: no communicator (call cfgpu_comm_init_nccl)");
        return 1;
    }
    CF_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < nmsg; ++i) {
        if (msgs[i].peer == c.rank) continue;
        if (msgs[i].sendbytes > 0) CF_NCCL(g_nccl.Send(msgs[i].send, (size_t)msgs[i].sendbytes, NCCL_INT8, msgs[i].peer, c.nccl_comm, stream));
        if (msgs[i].recvbytes > 0) CF_NCCL(g_nccl.Recv(msgs[i].recv, (size_t)msgs[i].recvbytes, NCCL_INT8, msgs[i].peer, c.nccl_comm, stream));
    }
    CF_NCCL(g_nccl.GroupEnd());
    return 0;
#else
    set_last_error("comm_exchange: no exchange callback installed");
    return 1;
#endif
}

bool comm_peer_capable(const Comm& c) {
#ifdef CF_EMU
    return false;
#else
    return c.nranks > 1 && c.nccl_comm != nullptr && !c.peer_failed && !getenv("CFGPU_NO_PEER");
#endif
}

int comm_open_peers(Comm& c, void* local, void** peers, cudaStream_t stream) {
#ifdef CF_EMU
    set_last_error("peer memory is unavailable in the emulation build");
    return 1;
#else
    // all-gather of the 64-byte IPC handles through the communicator itself (device staging, grouped send/recv)
    cudaIpcMemHandle_t mine;
    CF_CUDA(cudaIpcGetMemHandle(&mine, local));
    char* d = nullptr;
    CF_CUDA(cudaMalloc((void**)&d, (size_t)c.nranks * sizeof mine));
    CF_CUDA(cudaMemcpyAsync(d + (size_t)c.rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice, stream));
    std::vector<ExchangeMsg> msgs;
    for (int r = 0; r < c.nranks; ++r)
        if (r != c.rank)
            msgs.push_back({r, d + (size_t)c.rank * sizeof mine, (long long)sizeof mine, d + (size_t)r * sizeof mine, (long long)sizeof mine});
    CF_TRY(comm_exchange(c, msgs.data(), (int)msgs.size(), stream));
    std::vector<cudaIpcMemHandle_t> all(c.nranks);
    CF_CUDA(cudaMemcpyAsync(all.data(), d, (size_t)c.nranks * sizeof mine, cudaMemcpyDeviceToHost, stream));
    CF_CUDA(cudaStreamSynchronize(stream));
    CF_CUDA(cudaFree(d));
    // map; whether it worked is agreed on collectively so that all ranks take the same path afterwards
    double ok = 1.0;
    for (int r = 0; r < c.nranks; ++r) {
        if (r == c.rank) { peers[r] = local; continue; }
        peers[r] = nullptr;
        if (cudaIpcOpenMemHandle(&peers[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            peers[r] = nullptr;
            ok = 0.0;
        }
    }
    if (!c.d_bar) {
        CF_CUDA(cudaMalloc((void**)&c.d_bar, sizeof(double)));
    }
    ok = -ok;  // all-reduce MAX of -ok == -min(ok)
    CF_CUDA(cudaMemcpyAsync(c.d_bar, &ok, sizeof(double), cudaMemcpyHostToDevice, stream));
    CF_TRY(comm_allreduce(c, c.d_bar, 1, 1, stream));
    CF_CUDA(cudaMemcpyAsync(&ok, c.d_bar, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CF_CUDA(cudaStreamSynchronize(stream));
    CF_CUDA(cudaMemsetAsync(c.d_bar, 0, sizeof(double), stream));
    if (ok > -0.5) {
        for (int r = 0; r < c.nranks; ++r) {
            if (r != c.rank && peers[r]) cudaIpcCloseMemHandle(peers[r]);
            peers[r] = nullptr;
        }
        c.peer_failed = true;
    }
    return 0;
#endif
}

int comm_close_peers(Comm& c, void** peers) {
#ifndef CF_EMU
    for (int r = 0; r < c.nranks; ++r) {
        if (r != c.rank && peers[r]) cudaIpcCloseMemHandle(peers[r]);
        peers[r] = nullptr;
    }
#endif
    return 0;
}

int comm_barrier(Comm& c, cudaStream_t stream) {
    if (c.nranks == 1) return 0;
    if (!c.d_bar) {
        CF_CUDA(cudaMalloc((void**)&c.d_bar, sizeof(double)));
        CF_CUDA(cudaMemset(c.d_bar, 0, sizeof(double)));
    }
    return comm_allreduce(c, c.d_bar, 1, 1, stream);
}

// ------------------------------------------------------------------------------------------------ push all-to-all
namespace {
constexpr int PUSH_THREADS = 256;
constexpr long PUSH_CHUNK = 4096;  // complex elements per work item (64 KB)

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
#ifdef CF_EMU
    *p = v;
#else
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
#ifdef CF_EMU
    return *p;
#else
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ unsigned long long global_ns() {
#ifdef CF_EMU
    return 0;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}

__global__ void __launch_bounds__(PUSH_THREADS) slab_push_kernel(const PushParams p) {
    // work items = 64 KB chunks of all messages, dealt round-robin to the CTAs
    long item = blockIdx.x;
    for (int m = 0; m < p.nmsg; ++m) {
        const long nchunks = (p.msg[m].n + PUSH_CHUNK - 1) / PUSH_CHUNK;
        const double2* __restrict__ src = p.msg[m].src;
        double2* __restrict__ dst = p.msg[m].dst;
        for (; item < nchunks; item += gridDim.x) {
            const long i0 = item * PUSH_CHUNK;
            const long i1 = i0 + PUSH_CHUNK < p.msg[m].n ? i0 + PUSH_CHUNK : p.msg[m].n;
            // four 16-byte elements per thread in flight
            for (long i = i0 + threadIdx.x; i < i1; i += 4 * PUSH_THREADS) {
                double2 v[4];
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    if (i + h * PUSH_THREADS < i1) v[h] = src[i + h * PUSH_THREADS];
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    if (i + h * PUSH_THREADS < i1) dst[i + h * PUSH_THREADS] = v[h];
            }
        }
        item -= nchunks;
    }
    // completion: every CTA makes its stores visible system-wide, the last one publishes the sequence number
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(p.done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *p.done_counter = 0;
            __threadfence_system();
            for (int r = 0; r < p.nranks; ++r) st_release_sys(p.flags[r] + (size_t)p.slot * COMM_MAXRANKS + p.rank, p.seq);
        }
    }
}

// completion flags after copy-engine transfers (the copies precede this kernel in stream order)
__global__ void __launch_bounds__(32) slab_signal_kernel(const PushParams p) {
    __threadfence_system();
    const int r = threadIdx.x;
    if (r < p.nranks) st_release_sys(p.flags[r] + (size_t)p.slot * COMM_MAXRANKS + p.rank, p.seq);
}

__global__ void __launch_bounds__(32) slab_wait_kernel(const unsigned long long* flags, int slot, int nranks, unsigned long long seq, int* err) {
    const int r = threadIdx.x;
    if (r < nranks) {
        const unsigned long long* w = flags + (size_t)slot * COMM_MAXRANKS + r;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(w) < seq) {
            if (global_ns() - t0 > 20000000000ull) {  // 20 s: a peer died or the call sequences diverged
                *err = 1;
                break;
            }
        }
    }
    __syncwarp();
    __threadfence_system();
}
}  // namespace

int slab_push_launch(const PushParams& p, int nctas, cudaStream_t stream) {
    if (p.nmsg < 0 || p.nmsg > PUSH_MAXMSG || p.slot < 0 || p.slot >= PUSH_SLOTS) {
        set_last_error("slab_push: bad message count / slot");
        return 1;
    }
    if (nctas <= 0) {
        // copy engines: one DMA transfer per block (they are contiguous on both sides), no SM is taken from the transforms
        for (int m = 0; m < p.nmsg; ++m)
            if (p.msg[m].n > 0)
                CF_CUDA(cudaMemcpyAsync(p.msg[m].dst, p.msg[m].src, (size_t)p.msg[m].n * sizeof(double2), cudaMemcpyDeviceToDevice, stream));
        CF_LAUNCH(slab_signal_kernel, dim3(1), dim3(32), 0, stream, p);
        CF_KERNEL_CHECK();
        return 0;
    }
    CF_LAUNCH(slab_push_kernel, dim3(nctas), dim3(PUSH_THREADS), 0, stream, p);
    CF_KERNEL_CHECK();
    return 0;
}
int slab_wait_launch(const unsigned long long* myflags, int slot, int nranks, unsigned long long seq, int* err_dev, cudaStream_t stream) {
    CF_LAUNCH(slab_wait_kernel, dim3(1), dim3(32), 0, stream, myflags, slot, nranks, seq, err_dev);
    CF_KERNEL_CHECK();
    return 0;
}

int comm_allreduce(Comm& c, double* dev, int n, int op, cudaStream_t stream) {
    if (c.nranks == 1) return 0;
    if (c.ext_allreduce) {
        CF_CUDA(cudaStreamSynchronize(stream));
        if (c.ext_allreduce(c.ext_user, dev, n, op) != 0) {
            set_last_error("comm_allreduce: external callback failed");
            return 1;
        }
        return 0;
    }
#ifndef CF_EMU
    if (!c.nccl_comm) {
        set_last_error("comm_allreduce: no communicator");
        return 1;
    }
    CF_NCCL(g_nccl.AllReduce(dev, dev, (size_t)n, NCCL_FLOAT64, op == 1 ? NCCL_MAX : NCCL_SUM, c.nccl_comm, stream));
    return 0;
#else
    set_last_error("comm_allreduce: no callback installed");
    return 1;
#endif
}

}  // namespace cfgpu
