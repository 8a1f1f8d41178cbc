/* TEST INFRASTRUCTURE -- runtime of the CPU emulation declared in tests/emu/cuda_runtime.h.
 *
 * One CUDA thread = one fiber with its own stack; fibers of a block are scheduled round-robin on one host
 * thread and only switch at barriers (__syncthreads, warp shuffles, MMA), so execution is deterministic.
 * Blocks of a grid are distributed over host threads.
 */
#include "cuda_runtime.h"

#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

extern "C" void cfemu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl cfemu_switch
.type cfemu_switch,@function
cfemu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cfemu_switch,.-cfemu_switch
)");

namespace cfemu {

thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local unsigned char* t_dyn_smem;

namespace {

const size_t STACK_BYTES = 256 * 1024;

struct Barrier {
    int count = 0;
    unsigned gen = 0;
};

struct Fiber {
    void* sp = nullptr;
    unsigned char* stack = nullptr;
    bool done = false;
};

struct Worker {
    std::vector<Fiber> fibers;  // stacks are reused between blocks
    void* sched_sp = nullptr;
    int cur = -1;
    int nthreads = 0;
    int alive = 0;
    Barrier block_bar;
    std::vector<Barrier> warp_bar;
    std::vector<int> warp_alive;
    std::vector<double> slot_a, slot_b;
    const std::function<void()>* body = nullptr;
    std::vector<unsigned char> smem;
};

thread_local Worker* t_w = nullptr;

void set_thread_index(int tid) {
    const unsigned bx = t_blockDim.x, by = t_blockDim.y;
    t_threadIdx.x = tid % bx;
    t_threadIdx.y = (tid / bx) % by;
    t_threadIdx.z = tid / (bx * by);
}

void yield_to_sched() {
    Worker* w = t_w;
    int me = w->cur;
    cfemu_switch(&w->fibers[me].sp, w->sched_sp);
    // resumed
    set_thread_index(me);
}

void fiber_main() {
    Worker* w = t_w;
    int me = w->cur;
    set_thread_index(me);
    (*w->body)();
    w->fibers[me].done = true;
    w->alive--;
    w->warp_alive[me / 32]--;
    for (;;) cfemu_switch(&w->fibers[me].sp, w->sched_sp);
}

void prepare_fiber(Fiber& f) {
    if (!f.stack) {
        f.stack = (unsigned char*)mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (f.stack == MAP_FAILED) { perror("cfemu mmap"); abort(); }
    }
    uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;              // alignment filler
    *--sp = (void*)&fiber_main;   // return address for the first switch
    for (int i = 0; i < 6; ++i) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
    f.sp = sp;
    f.done = false;
}

void barrier_wait(Barrier& b, const int& group_alive) {
    unsigned gen = b.gen;
    b.count++;
    for (;;) {
        if (b.gen != gen) return;
        if (b.count >= group_alive) {
            b.count = 0;
            b.gen++;
            return;
        }
        yield_to_sched();
    }
}

void run_block(Worker& w, dim3 grid, dim3 block, size_t smem, unsigned bid, const std::function<void()>& body) {
    t_w = &w;
    t_blockDim = block;
    t_gridDim = grid;
    t_blockIdx.x = bid % grid.x;
    t_blockIdx.y = (bid / grid.x) % grid.y;
    t_blockIdx.z = bid / (grid.x * grid.y);
    const int nt = block.x * block.y * block.z;
    if ((int)w.fibers.size() < nt) w.fibers.resize(nt);
    const int nwarps = (nt + 31) / 32;
    w.nthreads = nt;
    w.alive = nt;
    w.block_bar = Barrier();
    w.warp_bar.assign(nwarps, Barrier());
    w.warp_alive.assign(nwarps, 32);
    if (nt % 32) w.warp_alive[nwarps - 1] = nt % 32;
    w.slot_a.assign((size_t)nwarps * 32, 0.0);
    w.slot_b.assign((size_t)nwarps * 32, 0.0);
    w.body = &body;
    if (w.smem.size() < smem + 64) w.smem.resize(smem + 64);
    t_dyn_smem = (unsigned char*)(((uintptr_t)w.smem.data() + 63) & ~(uintptr_t)63);
    for (int i = 0; i < nt; ++i) prepare_fiber(w.fibers[i]);
    while (w.alive > 0) {
        for (int i = 0; i < nt; ++i) {
            if (w.fibers[i].done) continue;
            w.cur = i;
            cfemu_switch(&w.sched_sp, w.fibers[i].sp);
        }
    }
    w.cur = -1;
}

}  // namespace

void sync_block() {
    Worker* w = t_w;
    barrier_wait(w->block_bar, w->alive);
}
void sync_warp() {
    Worker* w = t_w;
    int wid = w->cur / 32;
    barrier_wait(w->warp_bar[wid], w->warp_alive[wid]);
}
double shfl_exchange_d(double v, int src_lane) {
    Worker* w = t_w;
    int wid = w->cur / 32, lane = w->cur % 32;
    w->slot_a[wid * 32 + lane] = v;
    barrier_wait(w->warp_bar[wid], w->warp_alive[wid]);
    double r = w->slot_a[wid * 32 + src_lane];
    barrier_wait(w->warp_bar[wid], w->warp_alive[wid]);
    return r;
}
void dmma884(double& c0, double& c1, double a, double b) {
    Worker* w = t_w;
    int wid = w->cur / 32, lane = w->cur % 32;
    double* A = &w->slot_a[wid * 32];
    double* B = &w->slot_b[wid * 32];
    A[lane] = a;  // A[row = lane/4][k = lane%4]
    B[lane] = b;  // B[k = lane%4][n = lane/4]
    barrier_wait(w->warp_bar[wid], w->warp_alive[wid]);
    const int row = lane / 4, col = 2 * (lane % 4);
    double s0 = c0, s1 = c1;
    for (int k = 0; k < 4; ++k) {
        s0 = std::fma(A[row * 4 + k], B[col * 4 + k], s0);
        s1 = std::fma(A[row * 4 + k], B[(col + 1) * 4 + k], s1);
    }
    barrier_wait(w->warp_bar[wid], w->warp_alive[wid]);
    c0 = s0;
    c1 = s1;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned nblocks = grid.x * grid.y * grid.z;
    if (nblocks == 0) return;
    static int nhw = [] {
        const char* e = getenv("CFEMU_THREADS");
        int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        return n > 0 ? n : 1;
    }();
    const unsigned nthr = std::min<unsigned>(nhw, nblocks);
    static thread_local Worker main_worker;
    if (nthr <= 1) {
        for (unsigned b = 0; b < nblocks; ++b) run_block(main_worker, grid, block, smem, b, body);
        return;
    }
    std::atomic<unsigned> next(0);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthr; ++t)
        pool.emplace_back([&]() {
            static thread_local Worker w;
            for (;;) {
                unsigned b = next.fetch_add(1);
                if (b >= nblocks) break;
                run_block(w, grid, block, smem, b, body);
            }
            for (auto& f : w.fibers)
                if (f.stack) { munmap(f.stack, STACK_BYTES); f.stack = nullptr; }
            w.fibers.clear();
        });
    for (auto& th : pool) th.join();
}

}  // namespace cfemu

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cfemu_event(); (*e)->t = 0; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
