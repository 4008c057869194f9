// wfb_pool.h — device-memory pool for the phase-level temporaries of the mapping path (index build, ANI sketches, L1 / L2 scratch).
//
// wfb_map_phase allocates and frees ~80 device buffers per call. cudaMalloc / cudaFree cost grows with what the process already has
// mapped (the aligner next door holds a 30+ GB wavefront workspace): inside the bench step the same calls that take 45 ms in a
// process of their own took 90 - 730 ms (WFB_TRACE=stages). Freed blocks are kept per device, bucketed by size (rounded up to 1/8 of
// their power of two), and handed out again; the cache is capped (WFB_POOL_MAX_GB, default 16) and dropped when a real cudaMalloc
// fails. All users launch on the legacy default stream, so a block that is reused is reused in stream order.
//
// Include AFTER the CUDA / CUB headers of a host file: it redirects that file's cudaMalloc / cudaFree calls.
#pragma once
#ifndef WFB_EMU
#include <cuda_runtime.h>
cudaError_t wfb_pool_malloc_(void** p, size_t bytes); /* wfa_host.cu */
cudaError_t wfb_pool_free_(void* p);
void wfb_pool_trim_(); /* returns every cached block of the current device to the driver */
template <class T>
static inline cudaError_t wfb_pool_malloc_t_(T** p, size_t bytes) { return wfb_pool_malloc_((void**)p, bytes); }
#define cudaMalloc(p, n) wfb_pool_malloc_t_((p), (n))
#define cudaFree(p) wfb_pool_free_((void*)(p))
#endif
