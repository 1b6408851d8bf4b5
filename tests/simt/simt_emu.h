/* TEST INFRASTRUCTURE — a small SIMT emulator: compiles the CUDA sources of real_robots_b200/csrc with g++ and runs
 * every thread of a block as a fiber, so that the `-m "not gpu"` tests can execute the SHIPPED kernel source (warp
 * shuffles, ballots, block barriers, shared memory, atomics) on the CPU and compare it with the oracle before any GPU
 * time is spent.  It is NOT a CPU fallback: nothing under real_robots_b200/ can load the library built from it (it lives in
 * tests/simt/_build, has no torch front end and exports the C ABI under the same names only for tests/simt/emu.py).
 *
 * Model: a launch runs its blocks one after another; inside a block every CUDA thread is a ucontext fiber, scheduled
 * round-robin.  A fiber runs until it reaches a collective (warp shuffle / vote / reduce, __syncwarp, __syncthreads),
 * publishes its operand and yields until all participating threads have arrived.  Lanes that call a *_sync collective
 * with a mask they are not converged on deadlock here (the emulator aborts with a message) — on hardware that is
 * undefined behaviour, so the emulator doubles as a convergence checker.  Atomics are plain read-modify-writes (one
 * fiber runs at a time).  fp semantics are the host's: build with -ffp-contract=off to mirror nvcc -fmad=false. */
#pragma once
#include <ucontext.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#define RR_SIMT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __constant__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

using std::isfinite;
using std::max;
using std::min;

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
struct int4 { int x, y, z, w; } __attribute__((aligned(16)));
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
struct uint2 { unsigned x, y; } __attribute__((aligned(8)));
struct double2 { double x, y; } __attribute__((aligned(16)));
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct ulonglong2 { unsigned long long x, y; } __attribute__((aligned(16)));
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r; r.x = x; r.y = y; return r; }

namespace simt {
struct Fiber { ucontext_t ctx; char* stack = nullptr; bool done = false; };
struct Warp { int count = 0, gen = 0; unsigned long long buf[2][32]; };
struct Block {
    int nthreads = 0, cur = 0, live = 0;
    int bar_count = 0, bar_gen = 0;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    std::vector<int> warp_live;      // live lanes per warp
    ucontext_t main_ctx;
    const std::function<void()>* body = nullptr;
    long spins = 0;
};
extern Block g_blk;
extern unsigned char* g_dyn_smem;
extern dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
static const size_t STACK_BYTES = 1024 * 1024;

inline void switch_to(int next) {
    Block& b = g_blk;
    const int prev = b.cur;
    b.cur = next;
    g_threadIdx.x = (unsigned)next;
    swapcontext(&b.fibers[prev].ctx, &b.fibers[next].ctx);
    g_threadIdx.x = (unsigned)prev;      // resumed
}
inline void yield() {
    Block& b = g_blk;
    if (++b.spins > 60000000L) { std::fprintf(stderr, "simt_emu: deadlock (a collective was reached by only some of its threads)\n"); std::abort(); }
    int n = b.cur;
    for (int k = 0; k < b.nthreads; k++) {
        n = (n + 1 == b.nthreads) ? 0 : n + 1;
        if (!b.fibers[n].done) break;
    }
    if (n != b.cur) switch_to(n);
}
inline int lane_id() { return (int)(g_threadIdx.x & 31u); }
inline int warp_id() { return (int)(g_threadIdx.x >> 5); }
inline unsigned warp_live_mask(int w) {      // lanes of warp w that have not returned
    unsigned m = 0;
    for (int l = 0; l < 32; l++) { int t = w * 32 + l; if (t < g_blk.nthreads && !g_blk.fibers[t].done) m |= 1u << l; }
    return m;
}
// wait until every live lane named in `mask` has arrived; returns the buffer index this collective used
inline int warp_arrive(unsigned mask, unsigned long long operand) {
    Warp& w = g_blk.warps[warp_id()];
    const int g = w.gen, slot = g & 1;
    w.buf[slot][lane_id()] = operand;
    w.count++;
    for (;;) {
        if (w.gen != g) break;
        const int expected = __builtin_popcount(mask & warp_live_mask(warp_id()));
        if (w.count >= expected) { w.count = 0; w.gen++; g_blk.spins = 0; break; }
        yield();
    }
    return slot;
}
inline void block_barrier() {
    Block& b = g_blk;
    const int g = b.bar_gen;
    b.bar_count++;
    for (;;) {
        if (b.bar_gen != g) break;
        if (b.bar_count >= b.live) { b.bar_count = 0; b.bar_gen++; b.spins = 0; break; }
        yield();
    }
}
template <typename T> inline unsigned long long to_bits(T v) { unsigned long long u = 0; static_assert(sizeof(T) <= 8, "operand too wide"); std::memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> inline T from_bits(unsigned long long u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }

void launch(int grid, int block, size_t smem, const std::function<void()>& body);
long launches();
// RR_LAUNCH(kernel, grid, block, smem, stream)(args...) in the sources becomes Launcher(grid, block, smem) << kernel-lambda, called with args
struct Launcher {
    int g, b; size_t s;
    Launcher(int g_, int b_, size_t s_) : g(g_), b(b_), s(s_) {}
    template <typename F> struct Bound {
        int g, b; size_t s; F f;
        template <typename... A> void operator()(A&&... a) { const std::function<void()> body = [&]() { f(a...); }; launch(g, b, s, body); }
    };
    template <typename F> Bound<F> operator<<(F f) { return Bound<F>{g, b, s, f}; }
};
}  // namespace simt

#define threadIdx simt::g_threadIdx
#define blockIdx simt::g_blockIdx
#define blockDim simt::g_blockDim
#define gridDim simt::g_gridDim

// ---- warp collectives
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_arrive(mask, 0); }
static inline void __syncthreads() { simt::block_barrier(); }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src) {
    const int s = simt::warp_arrive(mask, simt::to_bits(v));
    return simt::from_bits<T>(simt::g_blk.warps[simt::warp_id()].buf[s][src & 31]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask) {
    const int s = simt::warp_arrive(mask, simt::to_bits(v));
    return simt::from_bits<T>(simt::g_blk.warps[simt::warp_id()].buf[s][(simt::lane_id() ^ lanemask) & 31]);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, int delta) {
    const int s = simt::warp_arrive(mask, simt::to_bits(v));
    const int src = simt::lane_id() - delta;
    return simt::from_bits<T>(simt::g_blk.warps[simt::warp_id()].buf[s][src >= 0 ? src : simt::lane_id()]);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, int delta) {
    const int s = simt::warp_arrive(mask, simt::to_bits(v));
    const int src = simt::lane_id() + delta;
    return simt::from_bits<T>(simt::g_blk.warps[simt::warp_id()].buf[s][src < 32 ? src : simt::lane_id()]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const int s = simt::warp_arrive(mask, pred ? 1ull : 0ull);
    const unsigned live = mask & simt::warp_live_mask(simt::warp_id());
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if (((live >> l) & 1u) && simt::g_blk.warps[simt::warp_id()].buf[s][l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __activemask() { return simt::warp_live_mask(simt::warp_id()); }
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    const int s = simt::warp_arrive(mask, simt::to_bits(v));
    const unsigned live = mask & simt::warp_live_mask(simt::warp_id());
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if (((live >> l) & 1u) && simt::g_blk.warps[simt::warp_id()].buf[s][l] == simt::to_bits(v)) r |= 1u << l;
    return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    const int s = simt::warp_arrive(mask, v);
    const unsigned live = mask & simt::warp_live_mask(simt::warp_id());
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = std::max(r, (unsigned)simt::g_blk.warps[simt::warp_id()].buf[s][l]);
    return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    const int s = simt::warp_arrive(mask, v);
    const unsigned live = mask & simt::warp_live_mask(simt::warp_id());
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = std::min(r, (unsigned)simt::g_blk.warps[simt::warp_id()].buf[s][l]);
    return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    const int s = simt::warp_arrive(mask, v);
    const unsigned live = mask & simt::warp_live_mask(simt::warp_id());
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r |= (unsigned)simt::g_blk.warps[simt::warp_id()].buf[s][l];
    return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    const int s = simt::warp_arrive(mask, v);
    const unsigned live = mask & simt::warp_live_mask(simt::warp_id());
    unsigned r = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r += (unsigned)simt::g_blk.warps[simt::warp_id()].buf[s][l];
    return r;
}

// ---- atomics (one fiber runs at a time)
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }

// ---- scalar intrinsics
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; std::memcpy(&d, &u, 8); return d; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

// ---- the slice of the CUDA runtime API the library uses ("device" memory is host memory)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof *p); p->multiProcessorCount = 2; std::strcpy(p->name, "simt_emu"); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)std::malloc(n ? n : 1); if (*p) std::memset((void*)*p, 0xCB, n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
template <typename S> static inline cudaError_t cudaMemcpyToSymbol(S& sym, const void* s, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) { std::memcpy((char*)&sym + off, s, n); return cudaSuccess; }
template <typename S> static inline cudaError_t cudaMemcpyToSymbolAsync(S& sym, const void* s, size_t n, size_t off, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy((char*)&sym + off, s, n); return cudaSuccess; }
template <typename S> static inline cudaError_t cudaMemcpyFromSymbol(void* d, const S& sym, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyDeviceToHost) { std::memcpy(d, (const char*)&sym + off, n); return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
