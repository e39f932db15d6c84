// tests/emu/cuda_runtime.h — TEST INFRASTRUCTURE: a minimal SIMT emulator that lets the kernel headers of
// akuaengine_b200/csrc/ (pbf_kernels.cuh, radix_sort.cuh, slab_kernels.cuh, list_build.cuh) be compiled by g++ and run on the
// CPU, thread by thread, so that kernel LOGIC can be checked in the CPU-only container (tests/test_emu_*.py).
//
// It is found instead of the CUDA toolkit's <cuda_runtime.h> only when a translation unit is compiled with
// `g++ -DAKUA_HOST_EMU -I tests/emu` (tests/emu/emu_harness.cpp). The product library is always built by nvcc from the same
// headers and never sees this file; nothing under akuaengine_b200/ can load the emulated code.
//
// Model: one CTA at a time; every CUDA thread of the CTA is a ucontext fiber. Fibers run until they reach a synchronising
// intrinsic (__syncthreads, __syncwarp, warp shuffles / ballots / match / reduce), where they park until the other
// participants have arrived — so warp-synchronous code (the radix sort's match-any ranking, the migration compaction) runs
// with CUDA semantics. Exited threads count as arrived. Atomics are plain read-modify-writes (fibers are cooperative).
// Not modelled: concurrency between CTAs or kernels of ONE device, memory-model races, timing.
// All emulator state is thread_local (and __shared__ is `static thread_local`): one OS thread = one emulated device, so the
// multi-rank x-slab step runs with one thread per rank, peer-to-peer stores being plain stores into the other thread's arrays
// and the in-kernel flag waits real waits (tests/test_emu_slab.py). cuda_host_shim.h adds the host runtime API on top.
#pragma once
#ifndef AKUA_HOST_EMU
#error "tests/emu/cuda_runtime.h is the host emulation shim; compile with -DAKUA_HOST_EMU (tests only)"
#endif
#include <stdint.h>
#include <setjmp.h>
#include <ucontext.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- qualifiers ----
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

// ---- vector types ----
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int3 { int x, y, z; };
struct alignas(8) uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct dim3 { uint32_t x = 1, y = 1, z = 1; dim3() = default; dim3(uint32_t a, uint32_t b = 1, uint32_t c = 1) : x(a), y(b), z(c) {} };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline int3 make_int3(int x, int y, int z) { return {x, y, z}; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }

// ---- runtime API fragments the headers mention ----
typedef int cudaError_t;
typedef void* cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 1 };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; struct { int programmaticStreamSerializationAllowed; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes = 0; cudaStream_t stream = nullptr; cudaLaunchAttribute* attrs = nullptr; unsigned numAttrs = 0; };

namespace emu {

struct Warp {
    uint64_t val[2][32];
    uint32_t arrived[2] = {0, 0};
    uint64_t bufGen[2] = {~0ull, ~0ull};
};
struct Thread {
    ucontext_t ctx;            // entry point only: switches after the first go through _setjmp / _longjmp (no signal-mask syscalls)
    jmp_buf jb;
    bool started = false;
    bool done = false;
    uint64_t gen = 0;          // warp collectives this lane has completed
    uint64_t barrierGen = 0;   // __syncthreads this thread has passed
    bool atBarrier = false;
};
struct Cta {
    std::vector<Thread> threads;
    std::vector<Warp> warps;
    std::vector<char> stacks;
    ucontext_t sched;
    jmp_buf schedJb;
    int current = -1;
    uint64_t barrierGen = 0;
    std::function<void()> body;
};
extern thread_local Cta* g_cta;
extern thread_local long long g_clock;
void yield();
uint32_t live_mask(int warp);
// Gathers every participating lane's 64-bit payload; returns when all lanes named in `mask` (that are still alive) arrived.
void warp_exchange(uint32_t mask, uint64_t mine, uint64_t out[32], uint32_t* participants);
void launch(dim3 grid, dim3 block, const std::function<void()>& body);

}  // namespace emu

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

template <typename... KArgs, typename... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args... args) {
    emu::launch(cfg->gridDim, cfg->blockDim, [&]() { kernel(KArgs(args)...); });
    return cudaSuccess;
}

// ---- synchronisation ----
void __syncthreads();
inline void __syncwarp(uint32_t mask = 0xffffffffu) { uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, 0, o, &p); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
void emu_pause();   // a spinning emulated thread lets the other ranks' OS threads run
inline void __nanosleep(unsigned) { emu_pause(); }
inline long long clock64() { return ++emu::g_clock; }

// ---- warp collectives ----
template <typename T> inline uint64_t emu_pack(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> inline T emu_unpack(uint64_t u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }
inline int emu_lane() { return (int)(threadIdx.x & 31u); }
inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, pred ? 1 : 0, o, &p);
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if (((p >> l) & 1u) && o[l]) r |= 1u << l;
    return r;
}
template <typename T> inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, emu_pack(v), o, &p);
    const int lane = emu_lane(), base = lane & ~(width - 1), s = base + (src & (width - 1));
    return ((p >> s) & 1u) ? emu_unpack<T>(o[s]) : v;
}
template <typename T> inline T __shfl_up_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, emu_pack(v), o, &p);
    const int lane = emu_lane(), base = lane & ~(width - 1), s = lane - (int)delta;
    return (s >= base && ((p >> s) & 1u)) ? emu_unpack<T>(o[s]) : v;
}
template <typename T> inline T __shfl_down_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, emu_pack(v), o, &p);
    const int lane = emu_lane(), base = lane & ~(width - 1), s = lane + (int)delta;
    return (s < base + width && ((p >> s) & 1u)) ? emu_unpack<T>(o[s]) : v;
}
template <typename T> inline T __shfl_xor_sync(uint32_t mask, T v, int laneMask, int width = 32) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, emu_pack(v), o, &p);
    const int lane = emu_lane(), s = lane ^ laneMask;
    return ((s & ~(width - 1)) == (lane & ~(width - 1)) && ((p >> s) & 1u)) ? emu_unpack<T>(o[s]) : v;
}
template <typename T> inline uint32_t __match_any_sync(uint32_t mask, T v) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, emu_pack(v), o, &p);
    const uint64_t mine = emu_pack(v);
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if (((p >> l) & 1u) && o[l] == mine) r |= 1u << l;
    return r;
}
inline uint32_t __reduce_min_sync(uint32_t mask, uint32_t v) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, v, o, &p);
    uint32_t r = v;
    for (int l = 0; l < 32; l++) if ((p >> l) & 1u) r = (uint32_t)o[l] < r ? (uint32_t)o[l] : r;
    return r;
}
inline uint32_t __reduce_add_sync(uint32_t mask, uint32_t v) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, v, o, &p);
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if ((p >> l) & 1u) r += (uint32_t)o[l];
    return r;
}
inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) {
    uint64_t o[32]; uint32_t p; emu::warp_exchange(mask, v, o, &p);
    uint32_t r = v;
    for (int l = 0; l < 32; l++) if ((p >> l) & 1u) r = (uint32_t)o[l] > r ? (uint32_t)o[l] : r;
    return r;
}

// ---- atomics (fibers are cooperative: a plain read-modify-write is atomic) ----
template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }

// ---- math / bit intrinsics ----
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }

#include "cuda_host_shim.h"
