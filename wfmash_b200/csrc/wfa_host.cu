// wfa_host.cu — host side of the batched biWFA aligner behind the C ABI (include/wfmash_b200.h).
//
// Drives the kernels of wfa_kernels.h: stages the sequences in HBM (forward + reversed copies),
// seeds the task queues with one task per mapping record, then drains breakpoint / base tasks level
// by level (the reference's recursion wavefront_bialign_alignment, wavefront_bialign.c:1144-1221,
// unrolled breadth-first), compacts the operation strings and returns them.
//
// Compiled by nvcc for sm_100a into libwfmash_b200.so. With -DWFB_EMU (tests/emu only) the same file
// compiles with g++ and runs every kernel body as one host thread per CTA; see wfb_rt.h.
#include "wfa_kernels.h"
#include "../../include/wfmash_b200.h"

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};
void wfb_set_last_error_(const std::string& s) { g_last_error = s; }
void wfb_count_launch_() { g_launches.fetch_add(1); }
/* WFB_TRACE: host-side stage timeline ("[wfb] t+<ms since the previous mark> <tag>") */
void wfb_trace_mark_(const char* tag) {
  if (!getenv("WFB_TRACE")) return; /* read every time: callers switch it on between phases */
  static thread_local double last = 0;
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  fprintf(stderr, "[wfb] t+%9.1f ms  %s\n", last == 0 ? 0.0 : now - last, tag);
  last = now;
}

#ifndef WFB_EMU
/* ---- device-memory pool of the mapping path's temporaries (wfb_pool.h) ---- */
#include <map>
#include <mutex>
#include <unordered_map>
namespace {
struct DevPool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;      /* size -> block */
  std::unordered_map<void*, size_t> block_size;  /* every live or cached block this pool handed out */
  size_t cached = 0;
};
DevPool g_pool[16];
size_t pool_round(size_t n) {
  if (n < 256) n = 256;
  size_t p2 = 256;
  while (p2 < n) p2 <<= 1;
  const size_t step = std::max<size_t>(p2 >> 4, 256); /* sizes within a power of two: 8 steps of 1/16 of the upper bound */
  return (n + step - 1) / step * step;
}
size_t pool_cap_bytes() {
  static const size_t cap = (size_t)(getenv("WFB_POOL_MAX_GB") ? atof(getenv("WFB_POOL_MAX_GB")) : 16.0) * (1ull << 30);
  return cap;
}
void pool_drop_all(DevPool& P) { /* P.mu held */
  for (auto& kv : P.free_blocks) { P.block_size.erase(kv.second); cudaFree(kv.second); }
  P.free_blocks.clear();
  P.cached = 0;
}
}  // namespace
cudaError_t wfb_pool_malloc_(void** p, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  DevPool& P = g_pool[dev & 15];
  const size_t want = pool_round(bytes);
  {
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.free_blocks.lower_bound(want);
    if (it != P.free_blocks.end() && it->first <= want + want / 4) {
      *p = it->second;
      P.cached -= it->first;
      P.free_blocks.erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) { /* give the cache back to the driver and try once more */
    cudaGetLastError();
    { std::lock_guard<std::mutex> lk(P.mu); pool_drop_all(P); }
    e = cudaMalloc(p, want);
  }
  if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(P.mu); P.block_size[*p] = want; }
  return e;
}
cudaError_t wfb_pool_free_(void* p) {
  if (!p) return cudaSuccess;
  int dev = 0;
  cudaGetDevice(&dev);
  DevPool& P = g_pool[dev & 15];
  {
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.block_size.find(p);
    if (it != P.block_size.end()) {
      if (P.cached + it->second <= pool_cap_bytes()) {
        P.free_blocks.emplace(it->second, p);
        P.cached += it->second;
        return cudaSuccess;
      }
      P.block_size.erase(it);
    }
  }
  return cudaFree(p); /* not ours, or the cache is full */
}
void wfb_pool_trim_() {
  int dev = 0;
  cudaGetDevice(&dev);
  DevPool& P = g_pool[dev & 15];
  std::lock_guard<std::mutex> lk(P.mu);
  pool_drop_all(P);
}
#endif

#ifndef WFB_EMU
#define WFB_CHECK(call)                                                                              \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      g_last_error = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
      return (e_ == cudaErrorMemoryAllocation) ? WFB_ENOMEM : WFB_ECUDA;                             \
    }                                                                                                \
  } while (0)
#define WFB_LAUNCH(kernel, grid, block, stream, ...)                                                 \
  do {                                                                                               \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                                           \
    g_launches.fetch_add(1);                                                                         \
  } while (0)
#define WFB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                                      \
  do {                                                                                               \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                      \
    g_launches.fetch_add(1);                                                                         \
  } while (0)
typedef cudaStream_t wfb_stream_t;
static int dev_malloc(void** p, size_t bytes) {
  if (cudaMalloc(p, bytes) == cudaSuccess) return 0;
  cudaGetLastError();
  wfb_pool_trim_(); /* the mapping path's cached temporaries go back to the driver before the aligner gives up */
  if (cudaMalloc(p, bytes) == cudaSuccess) return 0;
  cudaGetLastError();
  return -1;
}
static void dev_free(void* p) { if (p) cudaFree(p); }
static int host_malloc(void** p, size_t bytes) { return cudaMallocHost(p, bytes) == cudaSuccess ? 0 : -1; }
static void host_free(void* p) { if (p) cudaFreeHost(p); }
#define WFB_H2D(dst, src, bytes, s) WFB_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s))
#define WFB_D2H(dst, src, bytes, s) WFB_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s))
#define WFB_D2D(dst, src, bytes, s) WFB_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s))
#define WFB_MEMSET(dst, v, bytes, s) WFB_CHECK(cudaMemsetAsync(dst, v, bytes, s))
#define WFB_STREAM_SYNC(s) WFB_CHECK(cudaStreamSynchronize(s))
#else
#define WFB_CHECK(call) do { (void)(call); } while (0)
#define WFB_LAUNCH(kernel, grid, block, stream, ...)                                                 \
  do {                                                                                               \
    for (int bid_ = 0; bid_ < (int)(grid); ++bid_) kernel(bid_, (int)(grid), __VA_ARGS__);           \
    g_launches.fetch_add(1);                                                                         \
  } while (0)
#define WFB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) WFB_LAUNCH(kernel, grid, block, stream, __VA_ARGS__)
typedef int wfb_stream_t;
static int dev_malloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : -1; }
static void dev_free(void* p) { free(p); }
static int host_malloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : -1; }
static void host_free(void* p) { free(p); }
#define WFB_H2D(dst, src, bytes, s) memcpy(dst, src, bytes)
#define WFB_D2H(dst, src, bytes, s) memcpy(dst, src, bytes)
#define WFB_D2D(dst, src, bytes, s) memcpy(dst, src, bytes)
#define WFB_MEMSET(dst, v, bytes, s) memset(dst, v, bytes)
#define WFB_STREAM_SYNC(s) ((void)0)
#endif

namespace {

// Grow-only device / pinned-host buffers owned by the aligner.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    dev_free(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (dev_malloc(&p, want) != 0) { p = nullptr; return -1; }
    cap = want;
    return 0;
  }
  void release() { dev_free(p); p = nullptr; cap = 0; }
};
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    host_free(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (host_malloc(&p, want) != 0) { p = nullptr; return -1; }
    cap = want;
    return 0;
  }
  void release() { host_free(p); p = nullptr; cap = 0; }
};

inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

}  // namespace

struct wfb_aligner {
  double endsfree_kernel_ms = 0; /* CUDA-event time of the ends-free kernel launches since the last reset (wfb_biwfa_paf_batch reads it) */
  uint64_t endsfree_h2d = 0, endsfree_d2h = 0;
  int device = 0;
  WfbPen pen{};
  uint64_t workspace_bytes = 0;
  int sm_count = 148;
  wfb_stream_t stream{};
  DevBuf d_seq, d_pairs, d_slots, d_dense, d_len, d_status, d_counters, d_ctrl;
  DevBuf d_q[4]; /* break[0], break[1], base[0], base[1] */
  DevBuf d_ws, d_arena, d_log, d_runs, d_srcoff, d_tasklog, d_team, d_flags;
  HostBuf h_seq, h_dense, h_misc;
#ifndef WFB_EMU
  cudaEvent_t ev[4]{};
#endif
};

static int env_int(const char* name, int def) {
  const char* v = getenv(name);
  if (!v || !*v) return def;
  const int x = atoi(v);
  return x > 0 ? x : def;
}
/* CTA shapes; the breakpoint kernel's can be tuned without rebuilding (WFB_BREAK_THREADS), bounded by
 * the __launch_bounds__ the library was compiled with. */
static int break_threads() {
  int t = env_int("WFB_BREAK_THREADS", 256);
  t = (t / 32) * 32;
  if (t < 32) t = 32;
  if (t > WFB_BREAK_MAXTHREADS) t = WFB_BREAK_MAXTHREADS;
  return t;
}
static const int kBaseThreads = 64;

extern "C" const char* wfb_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* wfb_version(void) {
#ifdef WFB_EMU
  return "wfmash_b200 0.1 host-emulation (tests only)";
#else
  return "wfmash_b200 0.1 sm_100a";
#endif
}
extern "C" uint64_t wfb_launch_count(void) { return g_launches.load(); }

extern "C" int wfb_device_count(void) {
#ifndef WFB_EMU
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
#else
  return 1;
#endif
}

extern "C" wfb_aligner_t* wfb_aligner_create(int device, const wfb_penalties_t* p, uint64_t workspace_bytes) {
  if (!p) { g_last_error = "penalties == NULL"; return nullptr; }
  /* the kernels rely on 0 <= o1 <= o2 (see wfb_overlap) and positive extensions */
  if (p->mismatch <= 0 || p->gap_extension1 <= 0 || p->gap_extension2 <= 0 || p->gap_opening1 < 0 ||
      p->gap_opening2 < p->gap_opening1) {
    g_last_error = "unsupported penalties (need x>0, e1>0, e2>0, 0<=o1<=o2)";
    return nullptr;
  }
  WfbPen pen;
  pen.x = p->mismatch; pen.o1 = p->gap_opening1; pen.e1 = p->gap_extension1;
  pen.o2 = p->gap_opening2; pen.e2 = p->gap_extension2;
  pen.scope = std::max(std::max(pen.o2 + pen.e2, pen.o1 + pen.e1), pen.x) + 1;
  /* scope + 1 slots are live during a step (the scope window + the slot being written); one more so that the slot a
   * step resets for its successor is outside the window too (a speculative step may be dropped, see wfb_break_task) */
  pen.R = pen.scope + 2;
  if (pen.R > WFB_RMAX) { g_last_error = "penalties too large for the wavefront ring (scope+2 > 40)"; return nullptr; }
#ifndef WFB_EMU
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    g_last_error = "no CUDA device (this library has no CPU path)";
    return nullptr;
  }
  if (device < 0 || device >= n) { g_last_error = "bad device index"; return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { g_last_error = "cudaSetDevice failed"; return nullptr; }
#endif
  wfb_aligner* a = new wfb_aligner();
  a->device = device;
  a->pen = pen;
#ifndef WFB_EMU
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) a->sm_count = prop.multiProcessorCount;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if (workspace_bytes == 0) workspace_bytes = (uint64_t)(free_b * 0.45);
  if (cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking) != cudaSuccess) { delete a; g_last_error = "stream"; return nullptr; }
  for (int i = 0; i < 4; ++i) cudaEventCreate(&a->ev[i]);
#else
  if (workspace_bytes == 0) workspace_bytes = 1ull << 30;
  a->sm_count = 2;
#endif
  a->workspace_bytes = workspace_bytes;
  return a;
}

extern "C" void wfb_aligner_destroy(wfb_aligner_t* a) {
  if (!a) return;
#ifndef WFB_EMU
  cudaSetDevice(a->device);
  cudaStreamSynchronize(a->stream);
  for (int i = 0; i < 4; ++i) cudaEventDestroy(a->ev[i]);
  cudaStreamDestroy(a->stream);
#endif
  DevBuf* bufs[] = {&a->d_seq, &a->d_pairs, &a->d_slots, &a->d_dense, &a->d_len, &a->d_status, &a->d_counters, &a->d_ctrl,
                    &a->d_q[0], &a->d_q[1], &a->d_q[2], &a->d_q[3], &a->d_ws, &a->d_arena, &a->d_log, &a->d_runs, &a->d_srcoff, &a->d_tasklog, &a->d_team, &a->d_flags};
  for (DevBuf* b : bufs) b->release();
  a->h_seq.release();
  a->h_dense.release();
  a->h_misc.release();
  delete a;
}

/* gather sequences that already live in device memory into the aligner's staging layout */
WFB_KERNEL(wfb_gather_kernel, const WfbPairDesc* pairs, int npairs, const uint8_t* src, const long long* src_p_off,
           const long long* src_t_off, uint8_t* seq) {
  WFB_KERNEL_PROLOGUE
  for (int i = bid; i < npairs; i += nblocks) {
    const WfbPairDesc pd = pairs[i];
    const uint8_t* sp = src + src_p_off[i];
    const uint8_t* st = src + src_t_off[i];
    for (int j = WFB_TID; j < pd.plen; j += WFB_NT) seq[pd.p_off + j] = sp[j];
    for (int j = WFB_TID; j < pd.tlen; j += WFB_NT) seq[pd.t_off + j] = st[j];
  }
}

static int gap_affine2p_score(const char* ops, int n, const WfbPen& p) {
  /* cigar_score_gap_affine2p, deps/WFA2-lib/alignment/cigar.c:304-342 (match = 0) */
  int score = 0, i = 0;
  while (i < n) {
    int j = i;
    while (j < n && ops[j] == ops[i]) ++j;
    const int len = j - i;
    if (ops[i] == 'X') score += p.x * len;
    else if (ops[i] == 'I' || ops[i] == 'D') score += std::min(p.o1 + p.e1 * len, p.o2 + p.e2 * len);
    i = j;
  }
  return -score;
}

/* Core: sequences are described by (host pointers) or (device buffer + offsets). */
static int align_impl(wfb_aligner* a, int32_t n, const wfb_pair_t* hpairs, const char* d_src, const int64_t* src_p_off,
                      const int32_t* src_p_len, const int64_t* src_t_off, const int32_t* src_t_len, char* ops,
                      int64_t ops_cap, wfb_aln_result_t* results, wfb_align_stats_t* stats, const float* cost_hint = nullptr,
                      const char** dense_view = nullptr /* library-internal: leave the operation strings in the aligner's pinned buffer (valid until
                                                           its next call) instead of copying them to `ops`; results[i].ops_offset points into it */) {
  if (!a || n < 0 || (n > 0 && (!results || (!ops && !dense_view)))) { g_last_error = "bad argument"; return WFB_EINVAL; }
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n == 0) return WFB_OK;
#ifndef WFB_EMU
  WFB_CHECK(cudaSetDevice(a->device));
#endif
  const WfbPen pen = a->pen;
  wfb_stream_t s = a->stream;
  /* ---- layout ---- */
  std::vector<WfbPairDesc> pd((size_t)n);
  long long seq_bytes = 16, slot_bytes = 0;
  int maxP = 0, maxT = 0;
  for (int i = 0; i < n; ++i) {
    const int plen = hpairs ? hpairs[i].pattern_len : src_p_len[i];
    const int tlen = hpairs ? hpairs[i].text_len : src_t_len[i];
    if (plen < 0 || tlen < 0 || (hpairs && ((plen && !hpairs[i].pattern) || (tlen && !hpairs[i].text)))) {
      g_last_error = "bad pair";
      return WFB_EINVAL;
    }
    pd[i].plen = plen; pd[i].tlen = tlen;
    pd[i].p_off = seq_bytes;    seq_bytes = align_up(seq_bytes + plen + 16, 16);
    pd[i].t_off = seq_bytes;    seq_bytes = align_up(seq_bytes + tlen + 16, 16);
    pd[i].ops_off = slot_bytes; slot_bytes = align_up(slot_bytes + plen + tlen + 8, 8);
    maxP = std::max(maxP, plen); maxT = std::max(maxT, tlen);
  }
  /* all forward copies first (the only part that crosses PCIe), then the reversed copies at the same offsets one region further:
   * the device fills those (wfb_reverse_kernel) */
  const long long fwd_bytes = seq_bytes;
  for (int i = 0; i < n; ++i) { pd[i].prev_off = fwd_bytes + pd[i].p_off; pd[i].trev_off = fwd_bytes + pd[i].t_off; }
  seq_bytes = 2 * fwd_bytes;
  {
    /* worst case a pair needs plen+tlen ops; demand that much so no result can be truncated */
    long long need = 0;
    for (int i = 0; i < n; ++i) need += pd[i].plen + pd[i].tlen;
    if (!dense_view && need > ops_cap) { g_last_error = "ops buffer too small (need sum(pattern_len+text_len))"; return WFB_ECAP; }
  }
  /* ---- device buffers ---- */
  if (a->d_seq.ensure((size_t)seq_bytes + 64) || a->d_pairs.ensure(sizeof(WfbPairDesc) * (size_t)n) ||
      a->d_slots.ensure((size_t)slot_bytes + 64) || a->d_dense.ensure((size_t)slot_bytes + 64) ||
      a->d_len.ensure(sizeof(int) * (size_t)n) || a->d_status.ensure(sizeof(int) * (size_t)n) ||
      a->d_counters.ensure(sizeof(WfbCounters)) || a->d_ctrl.ensure(sizeof(int) * 16)) {
    g_last_error = "device allocation failed (inputs)";
    return WFB_ENOMEM;
  }
  uint8_t* d_seq = (uint8_t*)a->d_seq.p;
  WfbPairDesc* d_pairs = (WfbPairDesc*)a->d_pairs.p;
  char* d_slots = (char*)a->d_slots.p;
  char* d_dense = (char*)a->d_dense.p;
  int* d_len = (int*)a->d_len.p;
  int* d_status = (int*)a->d_status.p;
  WfbCounters* d_counters = (WfbCounters*)a->d_counters.p;
  int* d_ctrl = (int*)a->d_ctrl.p; /* [0] break task counter, [1] base task counter, [2] n_break_next, [3] n_base_next */
#ifndef WFB_EMU
  WFB_CHECK(cudaEventRecord(a->ev[0], s));
#endif
  wfb_trace_mark_("align: layout");
  WFB_H2D(d_pairs, pd.data(), sizeof(WfbPairDesc) * (size_t)n, s);
  if (hpairs) {
    /* pack the forward copies in pinned memory, one H2D */
    if (a->h_seq.ensure((size_t)fwd_bytes)) { g_last_error = "pinned allocation failed"; return WFB_ENOMEM; }
    uint8_t* hs = (uint8_t*)a->h_seq.p;
    /* only the forward regions are written; the device-side reversed regions are filled by a kernel.
     * Copy the span [first p_off, last t_off+tlen) in one go. */
    {
      int nt = (int)std::min<long long>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u), n / 4);
      if (fwd_bytes < (1 << 20)) nt = 1;
      std::atomic<int> next(0);
      auto worker = [&]() {
        for (;;) {
          const int b = next.fetch_add(16);
          if (b >= n) return;
          const int e = std::min(n, b + 16);
          for (int i = b; i < e; ++i) {
            memcpy(hs + pd[i].p_off, hpairs[i].pattern, (size_t)pd[i].plen);
            memset(hs + pd[i].p_off + pd[i].plen, 0, 16);
            memcpy(hs + pd[i].t_off, hpairs[i].text, (size_t)pd[i].tlen);
            memset(hs + pd[i].t_off + pd[i].tlen, 0, 16);
          }
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(worker);
      worker();
      for (auto& t : th) t.join();
    }
    wfb_trace_mark_("align: sequences packed in pinned memory");
    WFB_H2D(d_seq, hs, (size_t)fwd_bytes, s);
    WFB_MEMSET(d_seq + fwd_bytes, 0, (size_t)fwd_bytes, s);
  } else {
    if (a->d_srcoff.ensure(sizeof(long long) * 2 * (size_t)n)) { g_last_error = "device allocation failed"; return WFB_ENOMEM; }
    long long* d_off = (long long*)a->d_srcoff.p;
    std::vector<long long> tmp((size_t)2 * n);
    for (int i = 0; i < n; ++i) { tmp[i] = src_p_off[i]; tmp[n + i] = src_t_off[i]; }
    WFB_H2D(d_off, tmp.data(), sizeof(long long) * 2 * (size_t)n, s);
    WFB_STREAM_SYNC(s); /* tmp goes out of scope */
    WFB_MEMSET(d_seq, 0, (size_t)seq_bytes, s);
    WFB_LAUNCH(wfb_gather_kernel, std::min(n, a->sm_count * 8), 256, s, d_pairs, n, (const uint8_t*)d_src, d_off, d_off + n, d_seq);
  }
  if (a->d_flags.ensure(sizeof(int) * (size_t)n)) { g_last_error = "device allocation failed (pair flags)"; return WFB_ENOMEM; }
  int* d_flags = (int*)a->d_flags.p;
  WFB_MEMSET(d_flags, 0, sizeof(int) * (size_t)n, s);
  WFB_LAUNCH(wfb_reverse_kernel, std::min(n, a->sm_count * 8), 256, s, d_pairs, n, d_seq, d_flags);
  WFB_MEMSET(d_slots, 0, (size_t)slot_bytes, s);
  WFB_MEMSET(d_status, 0, sizeof(int) * (size_t)n, s);
  WFB_MEMSET(d_counters, 0, sizeof(WfbCounters), s);
  /* ---- initial tasks (wavefront_bialign :1279-1283) ---- */
  std::vector<WfbTask> t_break, t_base;
  for (int i = 0; i < n; ++i) {
    WfbTask t;
    t.pair = i; t.pb = 0; t.pe = pd[i].plen; t.tb = 0; t.te = pd[i].tlen;
    t.cbegin = WFB_M; t.cend = WFB_M;
    const bool min_length = std::max(pd[i].plen, pd[i].tlen) <= WFB_FALLBACK_MIN_LENGTH;
    t.score_remaining = min_length ? 0 : INT_MAX;
    if (pd[i].plen == 0 || pd[i].tlen == 0 || !min_length) t_break.push_back(t);
    else t_base.push_back(t);
  }
  /* most expensive first (LPT): the makespan is bounded by the last heavy root a CTA picks up. Cost ~ score^2; the caller's hint
   * (expected edits, e.g. (1 - mapping identity) * length) orders roots of similar length, the length alone otherwise. */
  if (cost_hint && !(getenv("WFB_COST_SORT") && atoi(getenv("WFB_COST_SORT")) == 0))
    std::stable_sort(t_break.begin(), t_break.end(), [&](const WfbTask& x, const WfbTask& y) { return cost_hint[x.pair] > cost_hint[y.pair]; });
  else
    std::stable_sort(t_break.begin(), t_break.end(), [](const WfbTask& x, const WfbTask& y) {
      return (long long)(x.pe + x.te) > (long long)(y.pe + y.te);
    });
  /* ---- workspaces ---- */
  const int W = (int)align_up(maxP + maxT + 8, 8);
  const long long ws_stride = align_up(2LL * pen.R * 5 * W + 2LL * pen.R * 5 * ((W + 63) / 64) + 64, 8); /* rows + their block maxima (wfb_overlap); every CTA's rows 16-byte aligned */
  const int score_cap = std::max(WFB_RECOVERY_MIN_SCORE, pen.x * WFB_FALLBACK_MIN_LENGTH) + 12;
  const long long arena_stride = 5LL * (score_cap + 2) * (score_cap + 2) + 64;
  const int maxruns = 2 * score_cap + 16;
  const size_t base_cta_bytes = (size_t)arena_stride * 4 + (size_t)(score_cap + 1) * 5 * sizeof(WfbBaseMeta) + (size_t)maxruns * sizeof(WfbRun);
  const int kBreakThreads = break_threads();
  int cta_base = a->sm_count * 8;
  int cta_break = a->sm_count * 2;
#ifndef WFB_EMU
  {
    /* persistent CTAs pulling tasks from a queue: exactly as many as can be resident */
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wfb_break_kernel, kBreakThreads, 0) == cudaSuccess && nb > 0)
      cta_break = a->sm_count * nb;
    int nb2 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, wfb_base_kernel, kBaseThreads, 0) == cudaSuccess && nb2 > 0)
      cta_base = a->sm_count * nb2;
  }
#endif
  {
    const uint64_t budget = a->workspace_bytes;
    const uint64_t base_budget = std::min<uint64_t>(budget / 4, (uint64_t)cta_base * base_cta_bytes);
    cta_base = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)cta_base, base_budget / base_cta_bytes));
    const uint64_t break_budget = budget - (uint64_t)cta_base * base_cta_bytes;
    const uint64_t per = (uint64_t)ws_stride * 4;
    cta_break = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)cta_break, break_budget / per));
    /* indices inside one CTA's workspace are 32-bit */
    if (ws_stride >= (1LL << 31)) { g_last_error = "pair too long for 32-bit wavefront indexing"; return WFB_EINVAL; }
  }
  cta_break = std::max(1, std::min<int>(cta_break, (int)std::max<size_t>(1, std::max(t_break.size(), (size_t)n))));
  if (a->d_ws.ensure((size_t)cta_break * (size_t)ws_stride * 4) || a->d_arena.ensure((size_t)cta_base * (size_t)arena_stride * 4) ||
      a->d_log.ensure((size_t)cta_base * (size_t)(score_cap + 1) * 5 * sizeof(WfbBaseMeta)) ||
      a->d_runs.ensure((size_t)cta_base * (size_t)maxruns * sizeof(WfbRun))) {
    g_last_error = "device allocation failed (workspace)";
    return WFB_ENOMEM;
  }
  /* ---- persistent single-launch driver (default): no level barriers ---- */
  bool persist_done = false;
  double persist_ms = 0.0;
  {
    const char* sched = getenv("WFB_SCHED"); /* "level" selects the level-synchronous driver */
    const bool want_persist = !(sched && strcmp(sched, "level") == 0);
    int ctas = cta_break;
    int seq_smem_words = 0; /* 2-bit packed sequence windows of the running task (wfa_kernels.h): enough for 2 x (52 kb + 52 kb) by default */
#ifndef WFB_EMU
    {
      /* packed sequences: 52 KB hold 2 x (52 kb + 52 kb); the size must leave 2 CTAs per SM, and what the CTAs do not take stays L1
       * (the rows are read through it: 98 KB here cost 11 % of the kernel time on scerevisiae8, profiles/r02_row_staging_experiment.md) */
      /* exactly what the batch's largest pair needs (forward + reversed pattern and text at 16 bases per word): every KB not taken
       * can stay L1, and the driver picks the smallest shared-memory carve-out that fits two CTAs (scerevisiae8: 50 KB -> the 132 KB
       * carve-out instead of 164 KB, -2.6 % kernel time; 36 / 24 KB, which leave the 100 kb root tasks unpacked: +8 / +11 %) */
      const long long need_words = 2LL * (((long long)maxP + 30) / 16 + 3) + 2LL * (((long long)maxT + 30) / 16 + 3);
      const int need_kb = (int)std::min<long long>(56, (need_words * 4 + 1023) / 1024);
      const int want = getenv("WFB_SEQ_SMEM_KB") ? atoi(getenv("WFB_SEQ_SMEM_KB")) : need_kb;
      const int tries[3] = {want, std::min(want, 56), 0};
      int nb0 = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb0, wfb_persist_kernel, kBreakThreads, 0) != cudaSuccess || nb0 <= 0) { cudaGetLastError(); nb0 = 0; }
      for (int ti = 0; ti < 3; ++ti) {
        const int kb = tries[ti];
        seq_smem_words = kb > 0 ? kb * 256 : 0;
        if (seq_smem_words > 0 &&
            cudaFuncSetAttribute(wfb_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, seq_smem_words * 4) != cudaSuccess) {
          cudaGetLastError();
          continue;
        }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wfb_persist_kernel, kBreakThreads, (size_t)seq_smem_words * 4) != cudaSuccess) { cudaGetLastError(); nb = 0; }
        if (nb > 0 && (nb >= nb0 || kb == 0)) { ctas = a->sm_count * nb; break; } /* never trade resident CTAs for shared memory */
      }
      if (getenv("WFB_TRACE")) fprintf(stderr, "[wfb] persist: dynamic shared memory %d KB, %d CTAs (%d per SM without it)\n", seq_smem_words / 256, ctas, nb0);
    }
#endif
    const size_t per_cta = (size_t)ws_stride * 4 + base_cta_bytes;
    ctas = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ctas, a->workspace_bytes / per_cta));
    const size_t n0 = t_break.size() + t_base.size();
    size_t qcap = std::min<size_t>(std::max<size_t>((size_t)n * 1024, (size_t)1 << 16), (size_t)1 << 25);
    if (want_persist && n0 <= qcap / 2) {
      if (a->d_ws.ensure((size_t)ctas * (size_t)ws_stride * 4) || a->d_arena.ensure((size_t)ctas * (size_t)arena_stride * 4) ||
          a->d_log.ensure((size_t)ctas * (size_t)(score_cap + 1) * 5 * sizeof(WfbBaseMeta)) || a->d_runs.ensure((size_t)ctas * (size_t)maxruns * sizeof(WfbRun)) ||
          a->d_q[0].ensure(sizeof(WfbTask) * qcap) || a->d_q[1].ensure(sizeof(int) * qcap)) {
        g_last_error = "device allocation failed (persistent queue / workspace)";
        return WFB_ENOMEM;
      }
      std::vector<WfbTask> init(t_break);
      init.insert(init.end(), t_base.begin(), t_base.end());
      std::vector<int> ones(n0, 1);
      WFB_MEMSET(a->d_q[1].p, 0, sizeof(int) * qcap, s);
      WFB_H2D(a->d_q[0].p, init.data(), sizeof(WfbTask) * n0, s);
      WFB_H2D(a->d_q[1].p, ones.data(), sizeof(int) * n0, s);
      int ctrl0[16] = {0};
      ctrl0[8] = 0;            /* head */
      ctrl0[9] = (int)n0;      /* tail */
      ctrl0[10] = (int)n0;     /* outstanding */
      ctrl0[11] = 0;           /* error */
      WFB_H2D(d_ctrl, ctrl0, sizeof(int) * 16, s);
      WFB_STREAM_SYNC(s);
      WfbPQueue pq;
      pq.tasks = (WfbTask*)a->d_q[0].p; pq.ready = (int*)a->d_q[1].p; pq.head = d_ctrl + 8; pq.tail = d_ctrl + 9;
      pq.outstanding = d_ctrl + 10; pq.error = d_ctrl + 11; pq.cap = (int)qcap;
      long long* d_cta_log = nullptr;
      if (getenv("WFB_TRACE") && strcmp(getenv("WFB_TRACE"), "stages") != 0) { /* WFB_TRACE=stages: the host-stage timeline only, no in-kernel per-CTA / per-pair log */
        if (a->d_tasklog.ensure(sizeof(long long) * (4 * (size_t)ctas + 2 * (size_t)n))) { g_last_error = "device allocation failed (cta log)"; return WFB_ENOMEM; }
        d_cta_log = (long long*)a->d_tasklog.p;
        WFB_MEMSET(d_cta_log, 0, sizeof(long long) * (4 * (size_t)ctas + 2 * (size_t)n), s);
      }
      /* team mode: idle CTAs help the owners of wide wavefronts (WFB_TEAM=0 turns it off) */
      WfbTeamSlot* d_team = nullptr;
      int* d_team_list = nullptr;
#ifndef WFB_EMU
      if (!(getenv("WFB_TEAM") && atoi(getenv("WFB_TEAM")) == 0)) {
        const size_t tb = sizeof(WfbTeamSlot) * (size_t)ctas + sizeof(int) * WFB_TEAM_LIST;
        if (a->d_team.ensure(tb)) { g_last_error = "device allocation failed (team slots)"; return WFB_ENOMEM; }
        WFB_MEMSET(a->d_team.p, 0, tb, s);
        d_team = (WfbTeamSlot*)a->d_team.p;
        d_team_list = (int*)((char*)a->d_team.p + sizeof(WfbTeamSlot) * (size_t)ctas);
      }
#endif
#ifndef WFB_EMU
      WFB_CHECK(cudaEventRecord(a->ev[2], s));
#endif
      WFB_LAUNCH_SMEM(wfb_persist_kernel, ctas, kBreakThreads, (size_t)seq_smem_words * 4, s, pq, (const WfbPairDesc*)d_pairs, (const uint8_t*)d_seq, (int32_t*)a->d_ws.p, ws_stride,
                 W, (int32_t*)a->d_arena.p, arena_stride, (WfbBaseMeta*)a->d_log.p, score_cap, (WfbRun*)a->d_runs.p, maxruns, pen, d_slots,
                 d_status, d_counters, d_cta_log, d_team, d_team_list, (const int*)d_flags, seq_smem_words);
#ifndef WFB_EMU
      WFB_CHECK(cudaEventRecord(a->ev[3], s));
#endif
      int ctrl1[16] = {0};
      WFB_D2H(ctrl1, d_ctrl, sizeof(int) * 16, s);
      WFB_STREAM_SYNC(s);
#ifndef WFB_EMU
      {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { g_last_error = std::string("kernel: ") + cudaGetErrorString(e); return WFB_ECUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a->ev[2], a->ev[3]);
        persist_ms = ms;
      }
#endif
      if (d_cta_log) { /* WFB_TRACE: how busy were the CTAs, and when did they run out of work? */
        std::vector<long long> cl((size_t)4 * ctas + 2 * (size_t)n);
        WFB_D2H(cl.data(), d_cta_log, sizeof(long long) * cl.size(), s);
        WFB_STREAM_SYNC(s);
        long long t0 = LLONG_MAX, t1 = 0, busy = 0, help = 0;
        std::vector<long long> ends;
        for (int c = 0; c < ctas; ++c) { if (cl[4 * c + 2]) t0 = std::min(t0, cl[4 * c + 2]); t1 = std::max(t1, cl[4 * c + 1]); busy += cl[4 * c]; help += cl[4 * c + 3]; ends.push_back(cl[4 * c + 1]); }
        fprintf(stderr, "[wfb] persist helper CTA-ms %.0f\n", help * 1e-6);
        std::sort(ends.begin(), ends.end());
        const double span = (double)(t1 - t0) * 1e-6;
        fprintf(stderr, "[wfb] persist ctas=%d span_ms=%.1f busy=%.3f  CTA exit times (ms): p10=%.1f p50=%.1f p90=%.1f p99=%.1f max=%.1f\n", ctas, span,
                span > 0 ? (double)busy * 1e-6 / (span * ctas) : 0.0, (ends[ends.size() / 10] - t0) * 1e-6, (ends[ends.size() / 2] - t0) * 1e-6,
                (ends[ends.size() * 9 / 10] - t0) * 1e-6, (ends[ends.size() * 99 / 100] - t0) * 1e-6, span);
      }
      if (d_cta_log) {
        std::vector<long long> cl((size_t)4 * ctas + 2 * (size_t)n);
        WFB_D2H(cl.data(), d_cta_log, sizeof(long long) * cl.size(), s);
        WFB_STREAM_SYNC(s);
        long long t0 = LLONG_MAX;
        for (int c = 0; c < ctas; ++c) if (cl[4 * c + 2]) t0 = std::min(t0, cl[4 * c + 2]);
        std::vector<int> idx((size_t)n);
        for (int i = 0; i < n; ++i) idx[i] = i;
        const long long* pl = cl.data() + 4 * ctas;
        std::sort(idx.begin(), idx.end(), [&](int x, int y) { return pl[2 * x + 1] > pl[2 * y + 1]; });
        for (int j = 0; j < std::min(n, 12); ++j) {
          const int i = idx[j];
          fprintf(stderr, "[wfb]   late pair %d plen=%d tlen=%d done_at_ms=%.1f cta_ms=%.1f hint=%.0f\n", i, pd[i].plen, pd[i].tlen, (pl[2 * i + 1] - t0) * 1e-6,
                  pl[2 * i] * 1e-6, cost_hint ? cost_hint[i] : -1.f);
        }
        std::sort(idx.begin(), idx.end(), [&](int x, int y) { return pl[2 * x] > pl[2 * y]; });
        double tot = 0, top = 0;
        for (int i = 0; i < n; ++i) tot += pl[2 * i] * 1e-6;
        for (int j = 0; j < std::min(n, 100); ++j) top += pl[2 * idx[j]] * 1e-6;
        fprintf(stderr, "[wfb]   CTA-ms over all pairs %.0f; the 100 costliest pairs hold %.0f\n", tot, top);
        for (int j = 0; j < std::min(n, 12); ++j) {
          const int i = idx[j];
          fprintf(stderr, "[wfb]   heavy pair %d plen=%d tlen=%d done_at_ms=%.1f cta_ms=%.1f hint=%.0f\n", i, pd[i].plen, pd[i].tlen, (pl[2 * i + 1] - t0) * 1e-6,
                  pl[2 * i] * 1e-6, cost_hint ? cost_hint[i] : -1.f);
        }
      }
      if (ctrl1[11] == 0 && ctrl1[10] == 0) {
        persist_done = true;
      } else {
        /* queue overflow or watchdog: start over with the level-synchronous driver */
        WFB_MEMSET(d_slots, 0, (size_t)slot_bytes, s);
        WFB_MEMSET(d_status, 0, sizeof(int) * (size_t)n, s);
        WFB_MEMSET(d_counters, 0, sizeof(WfbCounters), s);
      }
    }
  }
  /* ---- level loop (fallback / WFB_SCHED=level) ---- */
  size_t n_break = persist_done ? 0 : t_break.size(), n_base = persist_done ? 0 : t_base.size();
  int cur = 0;
  {
    if (a->d_q[0].ensure(sizeof(WfbTask) * std::max<size_t>(n_break, 1)) || a->d_q[2].ensure(sizeof(WfbTask) * std::max<size_t>(n_base, 1))) {
      g_last_error = "device allocation failed (queues)";
      return WFB_ENOMEM;
    }
    if (n_break) WFB_H2D(a->d_q[0].p, t_break.data(), sizeof(WfbTask) * n_break, s);
    if (n_base) WFB_H2D(a->d_q[2].p, t_base.data(), sizeof(WfbTask) * n_base, s);
    WFB_STREAM_SYNC(s); /* host vectors + pd must stay alive until copied */
  }
  double break_ms = persist_ms;
  uint64_t levels = persist_done ? 1 : 0;
  const char* dbg_path = getenv("WFB_DEBUG_TASKS"); /* per-task timeline dump (tuning only) */
  FILE* dbg_f = dbg_path ? fopen(dbg_path, "w") : nullptr;
  if (!dbg_f) dbg_path = nullptr;
  if (dbg_f) fprintf(dbg_f, "level\ttask\tt0_ns\tt1_ns\tsmid\tsteps\tscore_f\tscore_r\tplen\ttlen\tstatus\tlevel_ms\n");
  while (n_break || n_base) {
    ++levels;
    const int nxt = cur ^ 1;
    const size_t cap_next = std::max<size_t>(2 * n_break, 1);
    if (a->d_q[nxt].ensure(sizeof(WfbTask) * cap_next) || a->d_q[2 + nxt].ensure(sizeof(WfbTask) * cap_next)) {
      g_last_error = "device allocation failed (queues)";
      return WFB_ENOMEM;
    }
    WFB_MEMSET(d_ctrl, 0, sizeof(int) * 16, s);
    WfbQueue qb, qs;
    qb.tasks = (WfbTask*)a->d_q[nxt].p; qb.count = d_ctrl + 2; qb.cap = (int)cap_next;
    qs.tasks = (WfbTask*)a->d_q[2 + nxt].p; qs.count = d_ctrl + 3; qs.cap = (int)cap_next;
    WfbTaskLog* d_tasklog = nullptr;
    if (dbg_path && n_break) {
      if (a->d_tasklog.ensure(sizeof(WfbTaskLog) * n_break)) { g_last_error = "device allocation failed (tasklog)"; return WFB_ENOMEM; }
      d_tasklog = (WfbTaskLog*)a->d_tasklog.p;
      WFB_MEMSET(d_tasklog, 0, sizeof(WfbTaskLog) * n_break, s);
    }
    if (n_break) {
#ifndef WFB_EMU
      WFB_CHECK(cudaEventRecord(a->ev[2], s));
#endif
      WFB_LAUNCH(wfb_break_kernel, (int)std::min<size_t>(n_break, (size_t)cta_break), kBreakThreads, s,
                 (const WfbTask*)a->d_q[cur].p, (int)n_break, d_ctrl + 0, (const WfbPairDesc*)d_pairs, (const uint8_t*)d_seq,
                 (int32_t*)a->d_ws.p, ws_stride, W, pen, qb, qs, d_slots, d_status, d_counters, d_tasklog);
#ifndef WFB_EMU
      WFB_CHECK(cudaEventRecord(a->ev[3], s));
#endif
    }
    if (n_base) {
      WFB_LAUNCH(wfb_base_kernel, (int)std::min<size_t>(n_base, (size_t)cta_base), kBaseThreads, s,
                 (const WfbTask*)a->d_q[2 + cur].p, (int)n_base, d_ctrl + 1, (const WfbPairDesc*)d_pairs, (const uint8_t*)d_seq,
                 (int32_t*)a->d_arena.p, arena_stride, (WfbBaseMeta*)a->d_log.p, score_cap, (WfbRun*)a->d_runs.p, maxruns, pen,
                 d_slots, d_status, d_counters);
    }
    int ctrl[4] = {0, 0, 0, 0};
    WFB_D2H(ctrl, d_ctrl, sizeof(int) * 4, s);
    WFB_STREAM_SYNC(s);
#ifndef WFB_EMU
    {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { g_last_error = std::string("kernel: ") + cudaGetErrorString(e); return WFB_ECUDA; }
      if (n_break) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a->ev[2], a->ev[3]);
        break_ms += ms;
      }
    }
#endif
    if (dbg_f && n_break) {
      std::vector<WfbTaskLog> tl(n_break);
      WFB_D2H(tl.data(), d_tasklog, sizeof(WfbTaskLog) * n_break, s);
      WFB_STREAM_SYNC(s);
      float lms = 0.f;
#ifndef WFB_EMU
      cudaEventElapsedTime(&lms, a->ev[2], a->ev[3]);
#endif
      for (size_t i = 0; i < n_break; ++i)
        fprintf(dbg_f, "%llu\t%zu\t%lld\t%lld\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.3f\n", (unsigned long long)levels, i, tl[i].t0, tl[i].t1,
                tl[i].smid, tl[i].steps, tl[i].score_f, tl[i].score_r, tl[i].plen, tl[i].tlen, tl[i].status, lms);
    }
    n_break = (size_t)std::min<long long>(ctrl[2], (long long)cap_next);
    n_base = (size_t)std::min<long long>(ctrl[3], (long long)cap_next);
    cur = nxt;
    if (levels > 4096) { g_last_error = "task recursion did not converge"; return WFB_ECUDA; }
  }
  if (dbg_f) fclose(dbg_f);
  /* ---- compaction + results ---- */
  WFB_LAUNCH(wfb_compact_kernel, std::min(n, a->sm_count * 8), 256, s, (const WfbPairDesc*)d_pairs, n, (const char*)d_slots, d_dense, d_len);
#ifndef WFB_EMU
  WFB_CHECK(cudaEventRecord(a->ev[1], s));
#endif
  if (a->h_dense.ensure((size_t)slot_bytes + 64) || a->h_misc.ensure(sizeof(int) * 2 * (size_t)n + sizeof(WfbCounters))) {
    g_last_error = "pinned allocation failed";
    return WFB_ENOMEM;
  }
  char* h_dense = (char*)a->h_dense.p;
  int* h_len = (int*)a->h_misc.p;
  int* h_status = h_len + n;
  WfbCounters* h_cnt = (WfbCounters*)(h_status + n);
  WFB_D2H(h_len, d_len, sizeof(int) * (size_t)n, s);
  WFB_D2H(h_status, d_status, sizeof(int) * (size_t)n, s);
  WFB_D2H(h_cnt, d_counters, sizeof(WfbCounters), s);
  wfb_trace_mark_("align: kernels enqueued");
  WFB_D2H(h_dense, d_dense, (size_t)slot_bytes, s);
  WFB_STREAM_SYNC(s);
  wfb_trace_mark_("align: kernels + D2H done");
  if (dense_view) { /* no copy, no score scan: the caller reads the strings where the D2H left them */
    *dense_view = h_dense;
    for (int i = 0; i < n; ++i) {
      wfb_aln_result_t& r = results[i];
      r.status = h_status[i]; r.ops_offset = pd[i].ops_off; r.ops_len = r.status == 0 ? h_len[i] : 0; r.score = 0; r.reserved_ = 0;
    }
  } else {
  int64_t out_off = 0;
  for (int i = 0; i < n; ++i) {
    wfb_aln_result_t& r = results[i];
    r.status = h_status[i];
    r.ops_offset = out_off;
    r.ops_len = 0;
    r.score = 0;
    r.reserved_ = 0;
    if (r.status == 0) {
      const int len = h_len[i];
      if (out_off + len > ops_cap) { g_last_error = "ops buffer too small"; return WFB_ECAP; }
      r.ops_len = len;
      out_off += len;
    }
  }
  { /* the copies and the score scans are independent per pair (a GB of operations per large batch): over the host cores */
    int nt = (int)std::min<long long>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u), n / 4);
    if (out_off < (1 << 20)) nt = 1;
    std::atomic<int> next(0);
    auto worker = [&]() {
      for (;;) {
        const int b = next.fetch_add(16);
        if (b >= n) return;
        const int e = std::min(n, b + 16);
        for (int i = b; i < e; ++i) {
          wfb_aln_result_t& r = results[i];
          if (r.status != 0) continue;
          memcpy(ops + r.ops_offset, h_dense + pd[i].ops_off, (size_t)r.ops_len);
          r.score = gap_affine2p_score(ops + r.ops_offset, r.ops_len, pen);
        }
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
  }
  }
  wfb_trace_mark_("align: ops copied out + scores");
  if (getenv("WFB_TRACE")) {
    float ms = 0.f;
#ifndef WFB_EMU
    cudaEventElapsedTime(&ms, a->ev[0], a->ev[1]);
#endif
    fprintf(stderr, "[wfb] main n=%d W=%d persist_ms=%.1f total_ms=%.1f levels=%llu cells=%.3g steps=%.3g ext=%.3g ovl=%.3g break_tasks=%llu base_tasks=%llu base_cells=%.3g base_steps=%.3g\n",
            n, W, persist_ms, ms, (unsigned long long)levels, (double)h_cnt->cells, (double)h_cnt->score_steps, (double)h_cnt->extend_matches,
            (double)h_cnt->overlap_tests, (unsigned long long)h_cnt->break_tasks, (unsigned long long)h_cnt->base_tasks, (double)h_cnt->base_cells,
            (double)h_cnt->base_score_steps);
  }
  if (stats) {
    stats->cells = h_cnt->cells;
    stats->extend_matches = h_cnt->extend_matches;
    stats->overlap_tests = h_cnt->overlap_tests;
    stats->score_steps = h_cnt->score_steps;
    stats->break_tasks = h_cnt->break_tasks;
    stats->base_tasks = h_cnt->base_tasks;
    stats->base_cells = h_cnt->base_cells;
    stats->base_extend_matches = h_cnt->base_extend_matches;
    stats->base_score_steps = h_cnt->base_score_steps;
    stats->levels = levels;
    stats->break_kernel_ms = break_ms;
    stats->h2d_bytes = (uint64_t)(hpairs ? fwd_bytes : 0) + sizeof(WfbPairDesc) * (uint64_t)n;
    stats->d2h_bytes = (uint64_t)slot_bytes + sizeof(int) * 2 * (uint64_t)n;
#ifndef WFB_EMU
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a->ev[0], a->ev[1]);
    stats->kernel_ms = ms;
#endif
  }
  return WFB_OK;
}

/* Head / tail patch alignments (wflign.cpp:280-305, 368-397): ends-free gap-affine-2p WFA, one CTA per pair. */
extern "C" int wfb_align_endsfree_batch(wfb_aligner_t* a, const wfb_endsfree_pair_t* hp, int32_t n, int32_t term_group, char* ops,
                                        int64_t ops_cap, wfb_aln_result_t* results) {
  if (!a || n < 0 || (n > 0 && (!hp || !results || !ops))) { g_last_error = "bad argument"; return WFB_EINVAL; }
  if (term_group != 1 && term_group != 8 && term_group != 16) { g_last_error = "term_group must be 1, 8 or 16"; return WFB_EINVAL; }
  if (n == 0) return WFB_OK;
#ifndef WFB_EMU
  WFB_CHECK(cudaSetDevice(a->device));
#endif
  const WfbPen pen = a->pen;
  wfb_stream_t s = a->stream;
  std::vector<WfbPairDesc> pd((size_t)n);
  std::vector<WfbTask> tasks((size_t)n);
  std::vector<WfbEndsFree> efs((size_t)n);
  long long seq_bytes = 16, slot_bytes = 0, need = 0;
  int maxPT = 0;
  for (int i = 0; i < n; ++i) {
    const int plen = hp[i].pattern_len, tlen = hp[i].text_len;
    if (plen <= 0 || tlen <= 0 || !hp[i].pattern || !hp[i].text || hp[i].pattern_begin_free < 0 || hp[i].pattern_begin_free > plen ||
        hp[i].pattern_end_free < 0 || hp[i].pattern_end_free > plen || hp[i].text_begin_free < 0 || hp[i].text_begin_free > tlen ||
        hp[i].text_end_free < 0 || hp[i].text_end_free > tlen) {
      g_last_error = "bad ends-free pair (wavefront_align.c checks P0<=|P|, Pf<=|P|, T0<=|T|, Tf<=|T|)";
      return WFB_EINVAL;
    }
    pd[i].plen = plen; pd[i].tlen = tlen;
    pd[i].p_off = seq_bytes; seq_bytes = align_up(seq_bytes + plen + 16, 16);
    pd[i].t_off = seq_bytes; seq_bytes = align_up(seq_bytes + tlen + 16, 16);
    pd[i].prev_off = pd[i].p_off; pd[i].trev_off = pd[i].t_off; /* no reverse aligner here */
    pd[i].ops_off = slot_bytes; slot_bytes = align_up(slot_bytes + plen + tlen + 8, 8);
    need += plen + tlen;
    maxPT = std::max(maxPT, plen + tlen);
    WfbTask t; t.pair = i; t.pb = 0; t.pe = plen; t.tb = 0; t.te = tlen; t.cbegin = WFB_M; t.cend = WFB_M; t.score_remaining = 0;
    tasks[i] = t;
    WfbEndsFree e; e.pbf = hp[i].pattern_begin_free; e.pef = hp[i].pattern_end_free; e.tbf = hp[i].text_begin_free; e.tef = hp[i].text_end_free;
    efs[i] = e;
  }
  if (need > ops_cap) { g_last_error = "ops buffer too small (need sum(pattern_len+text_len))"; return WFB_ECAP; }
  if (a->d_seq.ensure((size_t)seq_bytes + 64) || a->d_pairs.ensure(sizeof(WfbPairDesc) * (size_t)n) || a->d_slots.ensure((size_t)slot_bytes + 64) ||
      a->d_dense.ensure((size_t)slot_bytes + 64) || a->d_len.ensure(sizeof(int) * (size_t)n) || a->d_status.ensure(sizeof(int) * (size_t)n) ||
      a->d_ctrl.ensure(sizeof(int) * 16) || a->d_q[0].ensure(sizeof(WfbTask) * (size_t)n) || a->d_srcoff.ensure(sizeof(WfbEndsFree) * (size_t)n) ||
      a->h_seq.ensure((size_t)seq_bytes)) {
    g_last_error = "allocation failed";
    return WFB_ENOMEM;
  }
  uint8_t* hs = (uint8_t*)a->h_seq.p;
  for (int i = 0; i < n; ++i) {
    memcpy(hs + pd[i].p_off, hp[i].pattern, (size_t)pd[i].plen); memset(hs + pd[i].p_off + pd[i].plen, 0, 16);
    memcpy(hs + pd[i].t_off, hp[i].text, (size_t)pd[i].tlen);    memset(hs + pd[i].t_off + pd[i].tlen, 0, 16);
  }
  WFB_H2D(a->d_seq.p, hs, (size_t)seq_bytes, s);
  WFB_H2D(a->d_pairs.p, pd.data(), sizeof(WfbPairDesc) * (size_t)n, s);
  WFB_H2D(a->d_srcoff.p, efs.data(), sizeof(WfbEndsFree) * (size_t)n, s);
  WFB_MEMSET(a->d_slots.p, 0, (size_t)slot_bytes, s);
  WFB_MEMSET(a->d_status.p, 0, sizeof(int) * (size_t)n, s);
  WFB_STREAM_SYNC(s);
  int* d_ctrl = (int*)a->d_ctrl.p;
  int* d_status = (int*)a->d_status.p;
  std::vector<int> h_status((size_t)n, 0);
  /* pass 0: every pair, modest arena; later passes: only the pairs that ran out of arena / score slots */
  std::vector<int> todo((size_t)n);
  for (int i = 0; i < n; ++i) todo[i] = i;
  const long long arena_sizes[3] = {4LL << 20, 64LL << 20, 512LL << 20}; /* ints per CTA */
  const int score_caps[3] = {1024, 6000, 12000};
  for (int pass = 0; pass < 3 && !todo.empty(); ++pass) {
    const long long arena_stride = arena_sizes[pass];
    const int score_cap = score_caps[pass];
    const int maxruns = 2 * score_cap + 16;
    const int runflag_stride = (int)align_up(maxPT + 16, 16);
    const size_t per_cta = (size_t)arena_stride * 4 + (size_t)(score_cap + 1) * 5 * sizeof(WfbBaseMeta) + (size_t)maxruns * sizeof(WfbRun) + (size_t)runflag_stride;
    /* later passes hold the few patches of score > 1024: as many CTAs as the workspace allows (two per SM), largest patch first */
    int ctas = (int)std::min<size_t>(todo.size(), (size_t)a->sm_count * (pass == 0 ? 4 : pass == 1 ? 2 : 1));
    if (pass > 0)
      std::stable_sort(todo.begin(), todo.end(), [&](int x, int y) { return (long long)pd[x].plen + pd[x].tlen > (long long)pd[y].plen + pd[y].tlen; });
    ctas = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctas, (size_t)(a->workspace_bytes / per_cta)));
    if (a->d_arena.ensure((size_t)ctas * (size_t)arena_stride * 4) || a->d_log.ensure((size_t)ctas * (size_t)(score_cap + 1) * 5 * sizeof(WfbBaseMeta)) ||
        a->d_runs.ensure((size_t)ctas * (size_t)maxruns * sizeof(WfbRun)) || a->d_ws.ensure((size_t)ctas * (size_t)runflag_stride)) {
      g_last_error = "device allocation failed (ends-free workspace)";
      return WFB_ENOMEM;
    }
    std::vector<WfbTask> tk(todo.size());
    std::vector<WfbEndsFree> ek(todo.size());
    for (size_t j = 0; j < todo.size(); ++j) { tk[j] = tasks[todo[j]]; ek[j] = efs[todo[j]]; }
    WFB_H2D(a->d_q[0].p, tk.data(), sizeof(WfbTask) * tk.size(), s);
    WFB_H2D(a->d_srcoff.p, ek.data(), sizeof(WfbEndsFree) * ek.size(), s);
    WFB_MEMSET(d_ctrl, 0, sizeof(int) * 16, s);
    for (size_t j = 0; j < todo.size(); ++j) h_status[todo[j]] = 0;
    WFB_H2D(d_status, h_status.data(), sizeof(int) * (size_t)n, s);
    WFB_STREAM_SYNC(s);
    struct timespec tr_ts0; clock_gettime(CLOCK_MONOTONIC, &tr_ts0);
#ifndef WFB_EMU
    WFB_CHECK(cudaEventRecord(a->ev[2], s));
#endif
    /* a patch starts with a wavefront as wide as its free ends (hundreds of diagonals) and gains two diagonals per score: 128 threads
     * for the first pass, 256 for the patches of score > 1024 (the kernel's cells are one diagonal per thread per trip) */
    const int ef_threads = getenv("WFB_EF_THREADS") ? atoi(getenv("WFB_EF_THREADS")) : (pass == 0 ? 128 : 256);
    WFB_LAUNCH(wfb_endsfree_kernel, ctas, ef_threads, s, (const WfbTask*)a->d_q[0].p, (const WfbEndsFree*)a->d_srcoff.p, (int)todo.size(),
               d_ctrl + 0, (const WfbPairDesc*)a->d_pairs.p, (const uint8_t*)a->d_seq.p, (int32_t*)a->d_arena.p, arena_stride,
               (WfbBaseMeta*)a->d_log.p, score_cap, (WfbRun*)a->d_runs.p, maxruns, (unsigned char*)a->d_ws.p, runflag_stride, (int)term_group,
               pen, (char*)a->d_slots.p, d_status);
#ifndef WFB_EMU
    WFB_CHECK(cudaEventRecord(a->ev[3], s));
#endif
    WFB_D2H(h_status.data(), d_status, sizeof(int) * (size_t)n, s);
    WFB_STREAM_SYNC(s);
#ifndef WFB_EMU
    { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { g_last_error = std::string("kernel: ") + cudaGetErrorString(e); return WFB_ECUDA; } }
    { float ms = 0.f; cudaEventElapsedTime(&ms, a->ev[2], a->ev[3]); a->endsfree_kernel_ms += ms; }
#endif
    if (getenv("WFB_TRACE")) {
      struct timespec tr_ts1; clock_gettime(CLOCK_MONOTONIC, &tr_ts1);
      fprintf(stderr, "[wfb] endsfree n=%d pass=%d todo=%zu ctas=%d kernel+sync_ms=%.1f\n", n, pass, todo.size(), ctas,
              (tr_ts1.tv_sec - tr_ts0.tv_sec) * 1e3 + (tr_ts1.tv_nsec - tr_ts0.tv_nsec) * 1e-6);
    }
    std::vector<int> again;
    for (int i : todo) if (h_status[i] == WFB_PAIR_BASE_SCORE_CAP) again.push_back(i);
    todo.swap(again);
  }
  WFB_LAUNCH(wfb_compact_kernel, std::min(n, a->sm_count * 8), 256, s, (const WfbPairDesc*)a->d_pairs.p, n, (const char*)a->d_slots.p,
             (char*)a->d_dense.p, (int*)a->d_len.p);
  if (a->h_dense.ensure((size_t)slot_bytes + 64) || a->h_misc.ensure(sizeof(int) * (size_t)n)) { g_last_error = "pinned allocation failed"; return WFB_ENOMEM; }
  WFB_D2H(a->h_misc.p, a->d_len.p, sizeof(int) * (size_t)n, s);
  WFB_D2H(a->h_dense.p, a->d_dense.p, (size_t)slot_bytes, s);
  WFB_STREAM_SYNC(s);
  a->endsfree_h2d += (uint64_t)seq_bytes; a->endsfree_d2h += (uint64_t)slot_bytes;
  const int* h_len = (const int*)a->h_misc.p;
  const char* h_dense = (const char*)a->h_dense.p;
  int64_t out_off = 0;
  for (int i = 0; i < n; ++i) {
    wfb_aln_result_t& r = results[i];
    r.status = h_status[i]; r.ops_offset = out_off; r.ops_len = 0; r.score = 0; r.reserved_ = 0;
    if (r.status == 0) {
      memcpy(ops + out_off, h_dense + pd[i].ops_off, (size_t)h_len[i]);
      r.ops_len = h_len[i];
      r.score = gap_affine2p_score(ops + out_off, h_len[i], pen);
      out_off += h_len[i];
    }
  }
  return WFB_OK;
}

extern "C" int wfb_align_batch(wfb_aligner_t* a, const wfb_pair_t* pairs, int32_t n, char* ops, int64_t ops_cap,
                               wfb_aln_result_t* results, wfb_align_stats_t* stats) {
  if (n > 0 && !pairs) { g_last_error = "pairs == NULL"; return WFB_EINVAL; }
  return align_impl(a, n, pairs, nullptr, nullptr, nullptr, nullptr, nullptr, ops, ops_cap, results, stats);
}

/* (library-internal) ends-free kernel time / bytes accumulated since the last call; resets them */
void wfb_take_endsfree_counters_(wfb_aligner_t* a, double* kernel_ms, uint64_t* h2d, uint64_t* d2h) {
  *kernel_ms = a->endsfree_kernel_ms; *h2d = a->endsfree_h2d; *d2h = a->endsfree_d2h;
  a->endsfree_kernel_ms = 0; a->endsfree_h2d = 0; a->endsfree_d2h = 0;
}

/* (library-internal, epilogue.cu) the hinted batch with the operation strings left in the aligner's pinned buffer */
int wfb_align_batch_view_(wfb_aligner_t* a, const wfb_pair_t* pairs, int32_t n, const float* cost_hint, const char** dense, wfb_aln_result_t* results,
                          wfb_align_stats_t* stats) {
  if ((n > 0 && !pairs) || !dense) { g_last_error = "bad argument"; return WFB_EINVAL; }
  *dense = nullptr;
  return align_impl(a, n, pairs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, results, stats, cost_hint, dense);
}

extern "C" int wfb_align_batch_hinted(wfb_aligner_t* a, const wfb_pair_t* pairs, int32_t n, const float* cost_hint, char* ops, int64_t ops_cap,
                                      wfb_aln_result_t* results, wfb_align_stats_t* stats) {
  if (n > 0 && !pairs) { g_last_error = "pairs == NULL"; return WFB_EINVAL; }
  return align_impl(a, n, pairs, nullptr, nullptr, nullptr, nullptr, nullptr, ops, ops_cap, results, stats, cost_hint);
}

extern "C" int wfb_align_batch_device(wfb_aligner_t* a, const char* d_seq, const int64_t* pattern_off, const int32_t* pattern_len,
                                      const int64_t* text_off, const int32_t* text_len, int32_t n, char* ops, int64_t ops_cap,
                                      wfb_aln_result_t* results, wfb_align_stats_t* stats) {
  if (n > 0 && (!d_seq || !pattern_off || !pattern_len || !text_off || !text_len)) { g_last_error = "NULL argument"; return WFB_EINVAL; }
  return align_impl(a, n, nullptr, d_seq, pattern_off, pattern_len, text_off, text_len, ops, ops_cap, results, stats);
}

extern "C" void* wfb_device_malloc(int device, uint64_t bytes) {
#ifndef WFB_EMU
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
#endif
  void* p = nullptr;
  if (dev_malloc(&p, (size_t)bytes) != 0) return nullptr;
  return p;
}
extern "C" void wfb_device_free(int device, void* p) {
#ifndef WFB_EMU
  cudaSetDevice(device);
#endif
  (void)device;
  dev_free(p);
}
extern "C" int wfb_memcpy_h2d(int device, void* dst, const void* src, uint64_t bytes) {
#ifndef WFB_EMU
  WFB_CHECK(cudaSetDevice(device));
  WFB_CHECK(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice));
#else
  (void)device;
  memcpy(dst, src, (size_t)bytes);
#endif
  return WFB_OK;
}

#if defined(WFB_PHASE_TIMERS) && !defined(WFB_EMU)
/* tuning builds only (not declared in include/wfmash_b200.h): read and clear the phase timers */
extern "C" int wfb_debug_phase_timers(unsigned long long* out32) {
  if (cudaMemcpyFromSymbol(out32, g_wfb_phase, sizeof(unsigned long long) * 32) != cudaSuccess) return WFB_ECUDA;
  unsigned long long z[32] = {0};
  if (cudaMemcpyToSymbol(g_wfb_phase, z, sizeof(z)) != cudaSuccess) return WFB_ECUDA;
  return WFB_OK;
}
#endif

/* wavefront_align-shaped shim over the batch entry (one pair). */
extern "C" int wfb_wavefront_align(wfb_aligner_t* a, const char* pattern, int32_t pattern_length, const char* text, int32_t text_length,
                                   char* cigar_operations, int32_t cigar_cap, int32_t* cigar_length, int32_t* cigar_score) {
  if (!a || !pattern || !text || pattern_length < 0 || text_length < 0 || !cigar_operations || !cigar_length) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  wfb_pair_t pr;
  pr.pattern = pattern; pr.pattern_len = pattern_length; pr.text = text; pr.text_len = text_length;
  wfb_aln_result_t r;
  const int rc = wfb_align_batch(a, &pr, 1, cigar_operations, (int64_t)cigar_cap, &r, nullptr);
  if (rc != WFB_OK) return rc;
  *cigar_length = r.status == 0 ? r.ops_len : 0;
  if (cigar_score) *cigar_score = r.status == 0 ? r.score : 0;
  return r.status == 0 ? WFB_WF_STATUS_ALG_COMPLETED : WFB_WF_STATUS_UNATTAINABLE;
}
