/* TEST INFRASTRUCTURE -- CPU emulation of the small CUDA subset the kernels in channelflow_b200/csrc use.
 *
 * Purpose: this container has no GPU, and GPU time is scarce, so the *logic* of every kernel (indexing, shared
 * memory staging, barriers, warp shuffles, DMMA fragment layouts) is exercised here on the CPU by compiling the
 * unmodified .cu sources with g++ -DCF_EMU -Itests/emu into tests/_emu/libcfgpu_emu.so.  One CUDA thread = one
 * fiber (ucontext); __syncthreads / warp shuffles are cooperative barriers; blocks of a grid are spread over host
 * threads.  This library is loaded only by tests (explicit path); the product package never falls back to it.
 */
#ifndef CF_EMU_CUDA_RUNTIME_H
#define CF_EMU_CUDA_RUNTIME_H
#ifndef CF_EMU
#error "tests/emu/cuda_runtime.h is only for -DCF_EMU builds"
#endif
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) alignas(n)
#define __constant__ static
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

namespace cfemu {
struct ThreadCtx;
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern thread_local unsigned char* t_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void sync_block();
void sync_warp();
double shfl_exchange_d(double v, int src_lane);  // returns value held by src_lane (absolute lane in warp)
// MMA m8n8k4 emulation
void dmma884(double& c0, double& c1, double a, double b);
}  // namespace cfemu

#define threadIdx (cfemu::t_threadIdx)
#define blockIdx (cfemu::t_blockIdx)
#define blockDim (cfemu::t_blockDim)
#define gridDim (cfemu::t_gridDim)
#define warpSize 32

static inline void __syncthreads() { cfemu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cfemu::sync_warp(); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline double __shfl_sync(unsigned, double v, int src, int = 32) { return cfemu::shfl_exchange_d(v, src & 31); }
static inline double __shfl_xor_sync(unsigned, double v, int m, int = 32) {
    return cfemu::shfl_exchange_d(v, ((int)(threadIdx.x & 31)) ^ m);
}
static inline double __shfl_down_sync(unsigned, double v, int d, int = 32) {
    int l = (int)(threadIdx.x & 31);
    return cfemu::shfl_exchange_d(v, l + d < 32 ? l + d : l);
}
static inline double __shfl_up_sync(unsigned, double v, int d, int = 32) {
    int l = (int)(threadIdx.x & 31);
    return cfemu::shfl_exchange_d(v, l - d >= 0 ? l - d : l);
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }
static inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }

// atomics: blocks may run on different host threads
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long cmp, unsigned long long val) {
    __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline double atomicAdd(double* a, double v) {
    unsigned long long* p = (unsigned long long*)a;
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST), nw;
    double o;
    do {
        memcpy(&o, &old, 8);
        double n = o + v;
        memcpy(&nw, &n, 8);
    } while (!__atomic_compare_exchange_n(p, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return o;
}
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }

// ---- host runtime subset ----
typedef int cudaError_t;
typedef int cudaStream_t;
struct cfemu_event { double t; };
typedef cfemu_event* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNoDevice = 100 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaHostAllocDefault = 0 };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emu error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 4; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { return posix_memalign(p, 256, n ? n : 256) == 0 ? cudaSuccess : 2; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
struct cudaPitchedPtr { void* ptr; size_t pitch, xsize, ysize; };
struct cudaPos { size_t x, y, z; };
struct cudaExtent { size_t width, height, depth; };
struct cudaMemcpy3DParms { cudaPitchedPtr srcPtr, dstPtr; cudaPos srcPos, dstPos; cudaExtent extent; cudaMemcpyKind kind; };
static inline cudaPitchedPtr make_cudaPitchedPtr(void* p, size_t pitch, size_t xs, size_t ys) { cudaPitchedPtr r = {p, pitch, xs, ys}; return r; }
static inline cudaPos make_cudaPos(size_t x, size_t y, size_t z) { cudaPos r = {x, y, z}; return r; }
static inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { cudaExtent r = {w, h, d}; return r; }
static inline cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms* p, cudaStream_t = 0) {
    for (size_t z = 0; z < p->extent.depth; ++z)
        for (size_t y = 0; y < p->extent.height; ++y) {
            const char* s = (const char*)p->srcPtr.ptr + ((p->srcPos.z + z) * p->srcPtr.ysize + p->srcPos.y + y) * p->srcPtr.pitch + p->srcPos.x;
            char* d = (char*)p->dstPtr.ptr + ((p->dstPos.z + z) * p->dstPtr.ysize + p->dstPos.y + y) * p->dstPtr.pitch + p->dstPos.x;
            memmove(d, s, p->extent.width);
        }
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = 1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e);
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);

#endif
