// Multi-GPU plumbing of the slab decomposition (replaces CfMPI + the FFTW-MPI transposes of the reference:
// cfmpi.cpp:66-127, flowfield.cpp:577-667 plans, the Mzloc/Nyloc transposes inside makePhysical/makeSpectral).
//
// One process per GPU.  Spectral state: the retained kx rows (mxi = 0..2Kx) are split into contiguous ranges, one per
// rank (y and kz local => y-transform, tau solve, norms are local); physical state: the Ny Gauss-Lobatto planes are
// split into contiguous ranges (x, z local => x/z FFT passes and the pointwise nonlinear term are local).  Between
// the two there is exactly one personalised all-to-all per direction, issued as grouped ncclSend/ncclRecv on the
// context's stream straight from/into the pencil buffers (no packing kernels: the pencil layouts are chosen so that
// every message is one contiguous block).  NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by
// torch, else the system one), so single-GPU use has no NCCL dependency.  For the CPU tests (gloo, world size 2) the
// exchange and all-reduce can instead be supplied by the host through callbacks (the emulation build's "device"
// pointers are host pointers).
#pragma once
#include "../../include/cfgpu.h"
#include "cf_common.cuh"

namespace cfgpu {

constexpr int COMM_MAXRANKS = 16;

struct Comm {
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    double* d_bar = nullptr;  // one double for the stream-ordered barrier
    bool peer_failed = false; // mapping peer memory failed on some rank: every rank uses the staged exchange instead
    cfgpu_exchange_fn ext_exchange = nullptr;
    cfgpu_allreduce_fn ext_allreduce = nullptr;
    void* ext_user = nullptr;
};

// balanced contiguous split of n items: rank r owns [lo, hi)
inline void part_range(int n, int nranks, int r, int& lo, int& hi) {
    lo = (int)((long)n * r / nranks);
    hi = (int)((long)n * (r + 1) / nranks);
}

struct ExchangeMsg {
    int peer;
    const void* send;
    long long sendbytes;
    void* recv;
    long long recvbytes;
};

int comm_unique_id(void* out128);
int comm_init_nccl(Comm& c, int rank, int nranks, const void* id128);
int comm_destroy(Comm& c);
// personalised exchange on `stream`; messages to self are device-to-device copies
int comm_exchange(Comm& c, const ExchangeMsg* msgs, int nmsg, cudaStream_t stream);
// in-place all-reduce of n doubles in device memory; op: 0 = sum, 1 = max
int comm_allreduce(Comm& c, double* dev, int n, int op, cudaStream_t stream);

// ---- peer memory (NVLink / NVSwitch): with the NCCL backend every rank can map the other ranks' pencil buffers
// (CUDA IPC) and the transform kernels store straight into the consumer's memory -- the all-to-all is fused into the
// y-GEMM epilogue and the forward x-pass store; only a stream-ordered barrier separates producer and consumer.
bool comm_peer_capable(const Comm& c);
// map `local` (a cudaMalloc'ed buffer of this rank) on every rank: peers[r] = address of rank r's buffer in this process
int comm_open_peers(Comm& c, void* local, void** peers /* [nranks] */, cudaStream_t stream);
int comm_close_peers(Comm& c, void** peers);
// all ranks have executed everything enqueued on `stream` before this point (and their stores are visible)
int comm_barrier(Comm& c, cudaStream_t stream);

// ---- all-to-all as a push over peer memory with device-side completion flags (no NCCL call, no host round trip):
// every rank copies its contiguous blocks straight into the other ranks' buffers (16-byte loads / stores, 512 bytes per
// warp instruction, a few CTAs on a high-priority stream beside the transform kernels) and the last CTA to finish
// publishes a sequence number in a flag word of every receiver (st.release.sys after __threadfence_system); the consumer
// stream runs a one-CTA kernel that spins (ld.acquire.sys, bounded by a time-out) until all senders' flags have reached the
// expected sequence number.  Double buffering is not needed: a rank can only reach the next exchange after it has received
// the previous one from everybody, and everybody sends only after having consumed (cfgpu_nse.cu).
constexpr int PUSH_MAXMSG = 2 * COMM_MAXRANKS;
constexpr int PUSH_SLOTS = 8;
struct PushMsg {
    const double2* src;
    double2* dst;
    long n;  // complex elements
};
struct PushParams {
    int nmsg;
    PushMsg msg[PUSH_MAXMSG];
    unsigned long long* flags[COMM_MAXRANKS];  // every rank's flag buffer as mapped here ([PUSH_SLOTS][COMM_MAXRANKS] words)
    int nranks, rank, slot;
    unsigned long long seq;
    unsigned int* done_counter;  // local, zero between launches
};
// nctas > 0: SM push kernel with that many CTAs; nctas <= 0: copy engines (cudaMemcpyAsync per block) + a flag kernel
int slab_push_launch(const PushParams& p, int nctas, cudaStream_t stream);
// wait until flag word (slot, r) >= seq for every rank r; *err_dev is set to 1 on time-out (never hangs the GPU)
int slab_wait_launch(const unsigned long long* myflags, int slot, int nranks, unsigned long long seq, int* err_dev, cudaStream_t stream);

}  // namespace cfgpu
