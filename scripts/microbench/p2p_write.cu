// Micro-benchmark (development aid, not product): achieved NVLink write bandwidth GPU0 -> GPU1 for
//  (a) plain 16-byte stores from G CTAs, (b) cudaMemcpyAsync (copy engine), (c) cp.async.bulk shared->global (TMA bulk
//  store) from G CTAs, (d) 16-byte stores in 64-byte scattered pieces (the fused-epilogue pattern).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_write p2p_write.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

__global__ void copy_st(const double2* __restrict__ src, double2* __restrict__ dst, long n) {
    const long stride = (long)gridDim.x * blockDim.x;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double2 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}
// 64-byte pieces: thread t of a 4-thread group writes 16 B; groups are `pitch` elements apart (scattered rows)
__global__ void copy_scatter64(const double2* __restrict__ src, double2* __restrict__ dst, long n, long pitch) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long g = i >> 2, t = i & 3;
        const long rows = n / pitch;           // g -> (col group, row): consecutive groups go to different rows
        const long r = g % rows, cg = g / rows;
        const long o = r * pitch + cg * 4 + t;
        if (o < n) dst[o] = src[o];
    }
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// bulk: each CTA moves CH-byte chunks: global(local) -> smem via cp.async.bulk + mbarrier, smem -> global(peer) via bulk store
template <int CH, int STAGES>
__global__ void copy_bulk(const char* __restrict__ src, char* __restrict__ dst, long bytes) {
    extern __shared__ __align__(128) char sm[];
    __shared__ __align__(8) unsigned long long bar[STAGES];
    const long nch = bytes / CH;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int phase[STAGES];
    for (int s = 0; s < STAGES; ++s) phase[s] = 0;
    long issued = 0, stored = 0;
    long c = blockIdx.x;
    long mine = 0;
    for (long k = c; k < nch; k += gridDim.x) ++mine;
    // prologue
    auto issue_load = [&](long j) {
        const int s = (int)(j % STAGES);
        const long chunk = c + j * gridDim.x;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)s * CH)),
                     "l"(src + chunk * CH), "r"(CH), "r"(smem_u32(&bar[s])) : "memory");
    };
    for (; issued < mine && issued < STAGES; ++issued) issue_load(issued);
    for (; stored < mine; ++stored) {
        const int s = (int)(stored % STAGES);
        // wait for the load of this stage
        asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(&bar[s])), "r"(phase[s]) : "memory");
        phase[s] ^= 1;
        const long chunk = c + stored * gridDim.x;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + chunk * CH), "r"(smem_u32(sm + (size_t)s * CH)), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (issued < mine) {
            // the stage we are about to refill is the one stored STAGES-1 iterations ago... wait until at most STAGES-1 stores pending reads
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
            // refill the oldest fully-read stage: stage of (issued % STAGES) == stage of (stored+1) % STAGES only if issued == stored + STAGES... keep it simple:
            // wait for ALL reads of the stage to be refilled
            if ((issued % STAGES) == s) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue_load(issued);
            ++issued;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const long bytes = 256L << 20;
    char *src, *dst, *loc;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&dst, bytes)); CK(cudaMemset(dst, 0, bytes));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&src, bytes)); CK(cudaMemset(src, 1, bytes)); CK(cudaMalloc(&loc, bytes));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char* name, int G, auto fn) {
        fn(); CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(e0, st));
        const int reps = 5;
        for (int r = 0; r < reps; ++r) fn();
        CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-34s G=%4d  %8.1f GB/s\n", name, G, bytes / (ms / reps * 1e-3) / 1e9);
        CK(cudaGetLastError());
    };
    const long n = bytes / 16;
    for (int G : {8, 16, 24, 48, 74, 148, 296, 592}) {
        timeit("st.16B remote", G, [&] { copy_st<<<G, 256, 0, st>>>((const double2*)src, (double2*)dst, n); });
    }
    for (int G : {148, 592}) timeit("st.16B local", G, [&] { copy_st<<<G, 256, 0, st>>>((const double2*)src, (double2*)loc, n); });
    for (int G : {148, 592, 1184}) timeit("st.16B remote 64B scattered", G, [&] { copy_scatter64<<<G, 256, 0, st>>>((const double2*)src, (double2*)dst, n, 2736 / 16 * 16); });
    timeit("cudaMemcpyAsync (CE) remote", 0, [&] { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st)); });
    {   // 4 concurrent CE copies on 4 streams
        cudaStream_t s4[4]; for (auto& s : s4) CK(cudaStreamCreate(&s));
        cudaEvent_t ev[4]; for (auto& evk : ev) CK(cudaEventCreateWithFlags(&evk, cudaEventDisableTiming));
        timeit("cudaMemcpyAsync x4 streams remote", 0, [&] {
            cudaEvent_t go; CK(cudaEventCreateWithFlags(&go, cudaEventDisableTiming)); CK(cudaEventRecord(go, st));
            for (int k = 0; k < 4; ++k) {
                CK(cudaStreamWaitEvent(s4[k], go, 0));
                CK(cudaMemcpyAsync(dst + k * (bytes / 4), src + k * (bytes / 4), bytes / 4, cudaMemcpyDeviceToDevice, s4[k]));
                CK(cudaEventRecord(ev[k], s4[k])); CK(cudaStreamWaitEvent(st, ev[k], 0));
            }
            CK(cudaEventDestroy(go));
        });
    }
    for (int G : {4, 8, 16, 24, 48, 148}) {
        constexpr int CH = 16384, STG = 4;
        CK(cudaFuncSetAttribute(copy_bulk<CH, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * STG));
        timeit("cp.async.bulk 16KB x4 remote", G, [&] { copy_bulk<CH, STG><<<G, 32, CH * STG, st>>>(src, dst, bytes); });
    }
    for (int G : {8, 24, 148}) {
        constexpr int CH = 32768, STG = 6;
        CK(cudaFuncSetAttribute(copy_bulk<CH, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * STG));
        timeit("cp.async.bulk 32KB x6 remote", G, [&] { copy_bulk<CH, STG><<<G, 32, CH * STG, st>>>(src, dst, bytes); });
    }
    // verify the bulk copy wrote the data
    CK(cudaSetDevice(1));
    unsigned char h[64];
    CK(cudaMemcpy(h, dst + bytes - 64, 64, cudaMemcpyDeviceToHost));
    printf("tail byte %d (expect 1)\n", h[63]);
    return 0;
}
