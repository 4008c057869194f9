// wfb_rt.h — thin execution layer under the kernels.
//
// Product build (nvcc, sm_100a): the macros below are the CUDA built-ins.
// Test-only build (-DWFB_EMU, plain g++): every kernel body runs as ONE thread per block, blocks
// run one after another on the host. This exists only so that the kernels' control flow, indexing
// and tie-breaking can be debugged in a container without a GPU (tests/emu/); it is never compiled
// into libwfmash_b200.so and is not a fallback: the product library fails with WFB_ENODEV when no
// device is present.
#pragma once
#include <stdint.h>
#include <limits.h>

#ifndef WFB_EMU
#include <cuda_runtime.h>
#define WFB_DEV __device__ __forceinline__
#define WFB_DEV_MEMBER __device__ __forceinline__
#define WFB_DEV_NOINLINE __device__ __noinline__
#define WFB_SHARED __shared__
#define WFB_TID ((int)threadIdx.x)
#define WFB_NT ((int)blockDim.x)
#define WFB_SYNC() __syncthreads()
#define WFB_KERNEL_PROLOGUE const int bid = (int)blockIdx.x; const int nblocks = (int)gridDim.x; (void)bid; (void)nblocks;
#define WFB_KERNEL(name, ...) __global__ void name(__VA_ARGS__)
#define WFB_KERNEL_LB(name, maxthreads, minblocks, ...) __global__ void __launch_bounds__(maxthreads, minblocks) name(__VA_ARGS__)
WFB_DEV int wfb_warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }
WFB_DEV int wfb_warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
WFB_DEV unsigned wfb_warp_add(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }
WFB_DEV int wfb_lane() { return (int)(threadIdx.x & 31); }
WFB_DEV void wfb_smem_min(int* p, int v) { atomicMin(p, v); }
WFB_DEV void wfb_smem_max(int* p, int v) { atomicMax(p, v); }
WFB_DEV int wfb_atomic_add(int* p, int v) { return atomicAdd(p, v); }
WFB_DEV void wfb_atomic_add64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
WFB_DEV unsigned long long atomicAdd_compat(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
#ifndef WFB_CACHE_HINTS
#define WFB_CACHE_HINTS 0
#endif
#if WFB_CACHE_HINTS
/* The wavefront rows stream through L1 once per score step and would push the (re-read) sequence bytes out of it:
 * rows are loaded evict-first, sequence bytes evict-last. */
WFB_DEV uint32_t wfb_ldg32(const uint32_t* p) { uint32_t v; asm volatile("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
WFB_DEV uint8_t wfb_ldg8(const uint8_t* p) { uint32_t v; asm volatile("ld.global.nc.L1::evict_last.u8 %0, [%1];" : "=r"(v) : "l"(p)); return (uint8_t)v; }
WFB_DEV int4 wfb_ld_row4(const int32_t* p) {
  int4 v;
  asm volatile("ld.global.L1::evict_first.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
WFB_DEV int32_t wfb_ld_row1(const int32_t* p) { int32_t v; asm volatile("ld.global.L1::evict_first.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
#else
WFB_DEV uint32_t wfb_ldg32(const uint32_t* p) { return __ldg(p); }
WFB_DEV uint8_t wfb_ldg8(const uint8_t* p) { return __ldg(p); }
WFB_DEV int4 wfb_ld_row4(const int32_t* p) { return *(const int4*)p; }
WFB_DEV int32_t wfb_ld_row1(const int32_t* p) { return *p; }
#endif
#else
#include <string.h>
#define WFB_DEV static inline
#define WFB_DEV_MEMBER inline
#define WFB_DEV_NOINLINE static
#define WFB_SHARED static
#define WFB_TID 0
#define WFB_NT 1
#define WFB_SYNC() ((void)0)
#define WFB_KERNEL_PROLOGUE
#define WFB_KERNEL(name, ...) static void name(int bid, int nblocks, __VA_ARGS__)
#define WFB_KERNEL_LB(name, maxthreads, minblocks, ...) static void name(int bid, int nblocks, __VA_ARGS__)
WFB_DEV int wfb_warp_min(int v) { return v; }
WFB_DEV int wfb_warp_max(int v) { return v; }
WFB_DEV unsigned wfb_warp_add(unsigned v) { return v; }
WFB_DEV int wfb_lane() { return 0; }
WFB_DEV void wfb_smem_min(int* p, int v) { if (v < *p) *p = v; }
WFB_DEV void wfb_smem_max(int* p, int v) { if (v > *p) *p = v; }
WFB_DEV int wfb_atomic_add(int* p, int v) { const int o = *p; *p = o + v; return o; }
WFB_DEV void wfb_atomic_add64(unsigned long long* p, unsigned long long v) { *p += v; }
WFB_DEV unsigned long long atomicAdd_compat(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
WFB_DEV uint32_t wfb_ldg32(const uint32_t* p) { return *p; }
WFB_DEV uint8_t wfb_ldg8(const uint8_t* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
/* enough of the CUDA vocabulary for the 4-diagonal group path and the word-wise extend to run under emulation */
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, unsigned shift) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (shift & 31));
}
static inline int __ffs(int v) { return v ? __builtin_ctz((unsigned)v) + 1 : 0; }
static inline int4 wfb_ld_row4(const int32_t* p) { return *(const int4*)p; }
static inline int32_t wfb_ld_row1(const int32_t* p) { return *p; }
#endif
