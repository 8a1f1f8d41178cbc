// Common device/host helpers for the cfgpu kernels (sm_100a).
//
// All kernels are launched through CF_LAUNCH so that the same sources can also be compiled by g++ with
// -DCF_EMU -Itests/emu (fiber-based CPU emulation, test infrastructure only -- see tests/emu/cuda_runtime.h).
// The product build is nvcc -gencode arch=compute_100a,code=sm_100a; there is no CPU code path in it.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

namespace cfgpu {
extern long long g_launches;  // kernels launched by this library (reported by cfgpu_launch_count)
}

#ifdef CF_EMU
#define CF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (++cfgpu::g_launches, cfemu::launch(grid, block, smem, [=]() { kernel(__VA_ARGS__); }))
#else
#define CF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (++cfgpu::g_launches, kernel<<<grid, block, smem, stream>>>(__VA_ARGS__))
#endif

// Launch with a thread-block cluster of `cluster` CTAs along x (co-scheduled on one GPC, started together).
#ifdef CF_EMU
#define CF_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, cluster, ...) CF_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)
#else
namespace cfgpu {
template <class... KArgs, class... Args>
inline void launch_cluster(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
}
}  // namespace cfgpu
#define CF_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, cluster, ...) \
    (++cfgpu::g_launches, cfgpu::launch_cluster(kernel, grid, block, smem, stream, cluster, __VA_ARGS__))
#endif

namespace cfgpu {

// ---- error plumbing (C-ABI returns int status; message via cfgpu_last_error) ----
void set_last_error(const std::string& msg);
inline int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return 0;
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    set_last_error(buf);
    return 1;
}
#define CF_CUDA(expr)                                                        \
    do {                                                                     \
        if (cfgpu::check_cuda((expr), #expr, __FILE__, __LINE__)) return 1;  \
    } while (0)
#define CF_KERNEL_CHECK() CF_CUDA(cudaGetLastError())
#define CF_TRY(expr)             \
    do {                         \
        if ((expr) != 0) return 1; \
    } while (0)

// ---- dynamic shared memory base ----
template <class T>
__device__ __forceinline__ T* dyn_smem() {
#ifdef CF_EMU
    return reinterpret_cast<T*>(cfemu::t_dyn_smem);
#else
    extern __shared__ __align__(16) unsigned char cf_dyn_smem_raw[];
    return reinterpret_cast<T*>(cf_dyn_smem_raw);
#endif
}

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 ----
// fragment layout (PTX ISA, mma.m8n8k4 .f64): a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c0,c1 = C[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
#ifdef CF_EMU
    cfemu::dmma884(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

// ---- cp.async (SASS LDGSTS): 16-byte global -> shared copies that bypass registers ----
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
#ifdef CF_EMU
    memcpy(smem_dst, gmem_src, 16);
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
#ifdef CF_EMU
    memcpy(smem_dst, gmem_src, 8);
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef CF_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {  // at most N committed groups still pending
#ifndef CF_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef CF_EMU
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}

// ---- L2 prefetch of one 128-byte line ----
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef CF_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ---- FP64 tensor-core MMA, the wide shape: D(16x8) += A(16x8, row) * B(8x8, col).  SASS on sm_100a: 4 x DMMA.8x8x4 ----
// fragment layout (PTX ISA, mma.m16n8k8 .f64), g = lane/4, t = lane%4:
//   a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4];  b0 = B[t][g], b1 = B[t+4][g];
//   c0,c1 = C[g][2t + {0,1}], c2,c3 = C[g+8][2t + {0,1}]
__device__ __forceinline__ void dmma_m16n8k8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
#ifdef CF_EMU
    cfemu::dmma884(c[0], c[1], a[0], b[0]);
    cfemu::dmma884(c[0], c[1], a[2], b[1]);
    cfemu::dmma884(c[2], c[3], a[1], b[0]);
    cfemu::dmma884(c[2], c[3], a[3], b[1]);
#else
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
#endif
}

// ---- atomic max on a double (signed compare, as the reference's CFL max has no abs) ----
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) < v) {
        unsigned long long assumed = old;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace cfgpu
