// TEST INFRASTRUCTURE ONLY.  A minimal CPU stand-in for the CUDA execution model, so that the kernels of
// lagrangian_microbes_b200/csrc/*.cu can be EXECUTED by the CPU test suite (tests/test_kernels_emulated.py): one
// std::thread per CUDA thread, one pthread barrier per CTA (__syncthreads) and per warp (__syncwarp and the
// warp collectives), CTAs one after the other.  Data races, byte stores and atomics are the real thing; clocks,
// memory spaces and scheduling are not modelled.  The product never includes this header: liblm_b200.so is built by
// nvcc from the untouched sources, this shim is found first on the include path only by tests/cuda_emu/emu_build.py.
#pragma once
#include <chrono>
#include <pthread.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static          // function-local static: shared by the threads of the CTA (CTAs run one at a time)

struct uint2 { unsigned int x, y; };
struct int2 { int x, y; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct uint3 { unsigned int x, y, z; };
static inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
constexpr cudaError_t cudaSuccess = 0;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// the rest of the runtime API the C ABI layer (csrc/api.cu) uses: memory is host memory, every "stream" is synchronous,
// events are inert (elapsed time 0)
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2 };
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// CUDA IPC: not available on the emulator (one process); lm_strip_peer_export falls back to plain pointers
struct cudaIpcMemHandle_t { char reserved[64]; };
constexpr unsigned int cudaIpcMemLazyEnablePeerAccess = 1;
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return 1; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned int) { return 1; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned int) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned int) { *d = h; return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline const char *cudaGetErrorName(cudaError_t) { return "emulated"; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }

namespace emu {
struct Warp { pthread_barrier_t bar; unsigned long long slot[32]; };
struct Cta { pthread_barrier_t bar; std::vector<Warp> warps; };
extern thread_local Warp *t_warp;
extern thread_local Cta *t_cta;
extern thread_local int t_lane;
extern void *dyn_smem;
void launch(unsigned int grid, unsigned int block, size_t smem, const std::function<void()> &body);
}  // namespace emu
extern thread_local uint3 threadIdx, blockIdx, blockDim, gridDim;

static inline void __syncthreads() { pthread_barrier_wait(&emu::t_cta->bar); }
static inline void __syncwarp(unsigned int = 0xffffffffu) { pthread_barrier_wait(&emu::t_warp->bar); }

template <class T> static inline T emu_exchange(T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    emu::t_warp->slot[emu::t_lane] = bits;
    pthread_barrier_wait(&emu::t_warp->bar);
    const unsigned long long got = emu::t_warp->slot[src & 31];
    pthread_barrier_wait(&emu::t_warp->bar);
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_sync(unsigned int, T v, int src) { return emu_exchange(v, src); }
template <class T> static inline T __shfl_up_sync(unsigned int, T v, unsigned int d) { return emu_exchange(v, emu::t_lane >= (int)d ? emu::t_lane - (int)d : emu::t_lane); }
template <class T> static inline T __shfl_xor_sync(unsigned int, T v, int m) { return emu_exchange(v, emu::t_lane ^ m); }
static inline unsigned int __ballot_sync(unsigned int, int pred)
{
    emu::t_warp->slot[emu::t_lane] = pred ? 1ull : 0ull;
    pthread_barrier_wait(&emu::t_warp->bar);
    unsigned int m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned int)emu::t_warp->slot[l] << l;
    pthread_barrier_wait(&emu::t_warp->bar);
    return m;
}
static inline int __any_sync(unsigned int mask, int pred) { return __ballot_sync(mask, pred) != 0u; }

template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned int atomicOr(unsigned int *p, unsigned int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned int atomicMax(unsigned int *p, unsigned int v)
{
    unsigned int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicMax(int *p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

static inline int __popc(unsigned int x) { return __builtin_popcount(x); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline long long clock64() { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned int)x) : 32; }
static inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
static inline unsigned int __byte_perm(unsigned int a, unsigned int b, unsigned int sel)
{
    const unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned int r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned int)((src >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);   // selector msb (sign replication) unused
    return r;
}
// explicit-rounding intrinsics: plain IEEE operations (the emulator is compiled with -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float emu_fast_cosf(float a) { return cosf(a); }
#define __cosf emu_fast_cosf
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned int __float_as_uint(float f) { unsigned int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int __float2int_rn(float x)
{
    if (std::isnan(x)) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)nearbyintf(x);
}
#define __log2f(x) log2f(x)          // glibc declares a private __log2f of its own
static inline float sinpif(float x) { return (float)sin(M_PI * (double)x); }
static inline float cospif(float x) { return (float)cos(M_PI * (double)x); }
using std::max;
using std::min;
static inline unsigned int min(unsigned int a, int b) { return a < (unsigned int)b ? a : (unsigned int)b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
