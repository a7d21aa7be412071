// TEST INFRASTRUCTURE ONLY.  Host emulation of the small CUDA subset used by the *general* path of
// 21cmmc_b200/csrc so that the `-m "not gpu"` test-suite can execute the kernel logic (index maths,
// FFT stages, binning, plan bookkeeping) against the oracle without a GPU.  One CUDA block at a time;
// the threads of a block are ucontext fibers, `__syncthreads()` yields to a round-robin scheduler.
// It is compiled only into tests/emu/libps21_emu.so (g++ -DPS21_EMU); the product library
// (21cmmc_b200/libps21.so) is built by nvcc for sm_100a and never contains this code.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct float2 {
    float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct alignas(16) float4 {
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint3 {
    unsigned x, y, z;
};
struct uint4 {
    unsigned x, y, z, w;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace ps21emu {
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
extern char* g_dyn_smem;
void syncthreads();
void syncwarp();
void spin_yield();  // stay runnable, let the other fibers advance (used by the emulated mbarrier wait)
unsigned long long* shfl_slot(int tid);
int linear_tid();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace ps21emu
using ps21emu::blockDim;
using ps21emu::blockIdx;
using ps21emu::gridDim;
using ps21emu::threadIdx;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __syncthreads() ps21emu::syncthreads()
#define __syncwarp() ps21emu::syncwarp()
#define __launch_bounds__(...)
#define __align__(n)
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lanemask) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    const int me = ps21emu::linear_tid();
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    *ps21emu::shfl_slot(me) = raw;
    ps21emu::syncwarp();
    raw = *ps21emu::shfl_slot((me & ~31) | ((me ^ lanemask) & 31));
    ps21emu::syncwarp();
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}
static inline int __any_sync(unsigned, int pred) {
    const int me = ps21emu::linear_tid();
    *ps21emu::shfl_slot(me) = pred ? 1ull : 0ull;
    ps21emu::syncwarp();
    int any = 0;
    for (int l = 0; l < 32; ++l) any |= (int)*ps21emu::shfl_slot((me & ~31) | l);
    ps21emu::syncwarp();
    return any;
}
static inline int __all_sync(unsigned m, int pred) { return !__any_sync(m, !pred); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    const int me = ps21emu::linear_tid();
    *ps21emu::shfl_slot(me) = pred ? 1ull : 0ull;
    ps21emu::syncwarp();
    unsigned out = 0;
    for (int l = 0; l < 32; ++l) out |= (unsigned)(*ps21emu::shfl_slot((me & ~31) | l) & 1ull) << l;
    ps21emu::syncwarp();
    return out;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    const int me = ps21emu::linear_tid();
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    *ps21emu::shfl_slot(me) = raw;
    ps21emu::syncwarp();
    raw = *ps21emu::shfl_slot((me & ~31) | (src & 31));
    ps21emu::syncwarp();
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F>
static inline int cudaFuncSetAttribute(F, int, int) { return 0; }
#define PS21_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(ps21emu::g_dyn_smem)

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline double __dadd_rn(double a, double b) { return a + b; }
template <class T>
static inline T atomicAdd(T* p, T v) {
    T old = *p;
    *p = old + v;
    return old;
}
static inline void __threadfence_block() {}
static inline void __threadfence() {}
using std::max;
using std::min;

// runtime API
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = std::malloc(n ? n : 1);
    return *p ? 0 : 2;
}
static inline cudaError_t cudaFree(void* p) {
    std::free(p);
    return 0;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    std::memcpy(d, s, n);
    return 0;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memcpy(d, s, n);
    return 0;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) {
    std::memset(d, v, n);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
    std::memset(d, v, n);
    return 0;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) {
    *s = (void*)1;
    return 0;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) {
    *e = (void*)1;
    return 0;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
