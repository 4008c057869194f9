// index_host.cu — host side of the GPU-resident reference index and the batched L1 stage (C ABI:
// wfb_index_build / wfb_index_export / wfb_index_free / wfb_l1_batch). Kernels: index_kernels.h,
// minmer_kernels.h, sketch_kernels.h (hand-written); CUB is used for the sorts / scans between them.
#include "l2_kernels.h"

#include <math.h>
#include <cmath>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include <limits>
#include <mutex>

#ifndef WFB_EMU
#include <cub/cub.cuh>
#endif
#include "wfb_pool.h" /* this file's cudaMalloc / cudaFree go through the library's device-memory pool */

void wfb_set_last_error_(const std::string& s);
void wfb_count_launch_();

struct wfb_index {
  int device = 0;
  wfb_index_params_t params{};
  long long n_minmers_all = 0, n_minmers = 0, n_points = 0, n_uniq = 0, n_buckets = 0;
  unsigned long long threshold = 0;
  wfb_minmer_t* d_minmers = nullptr; /* kept minmerIndex, reference order */
  uint64_t* d_points = nullptr;      /* packed interval points, grouped by hash */
  IxSlot* d_table = nullptr;
  unsigned long long* d_uhash = nullptr;
  uint32_t *d_ustart = nullptr, *d_ucount = nullptr;
  /* grow-only device workspace of the batched mapping calls (wfb_l1_batch / wfb_map_fragments_batch): allocated on the
   * first call and reused, so that a steady stream of batches does no cudaMalloc / cudaFree. Calls on one index serialise
   * on ws_mu. */
  mutable std::mutex ws_mu;
  mutable void* ws_ptr[48] = {};
  mutable size_t ws_cap[48] = {};
};

#ifndef WFB_EMU
#define IX_CHECK(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      wfb_set_last_error_(std::string(#call) + ": " + cudaGetErrorString(e_));           \
      rc = (e_ == cudaErrorMemoryAllocation) ? WFB_ENOMEM : WFB_ECUDA;                   \
      goto done;                                                                         \
    }                                                                                    \
  } while (0)
#define IX_LAUNCH(kernel, grid, block, smem, ...)                                        \
  do {                                                                                   \
    kernel<<<(grid), (block), (smem)>>>(__VA_ARGS__);                                    \
    wfb_count_launch_();                                                                 \
  } while (0)

namespace {
struct Tmp { /* grow-only CUB temp storage */
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t b) {
    if (b <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, b + 256);
    if (e == cudaSuccess) cap = b + 256;
    return e;
  }
  ~Tmp() { if (p) cudaFree(p); }
};
cudaError_t incl_scan(Tmp& t, const int* in, long long* out, long long n) {
  size_t b = 0;
  cub::DeviceScan::InclusiveSum(nullptr, b, in, out, (int)n);
  cudaError_t e = t.ensure(b);
  if (e != cudaSuccess) return e;
  b = t.cap;
  return cub::DeviceScan::InclusiveSum(t.p, b, in, out, (int)n);
}
cudaError_t excl_scan(Tmp& t, const int* in, long long* out, long long n) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, in, out, (int)n);
  cudaError_t e = t.ensure(b);
  if (e != cudaSuccess) return e;
  b = t.cap;
  return cub::DeviceScan::ExclusiveSum(t.p, b, in, out, (int)n);
}
}  // namespace
#endif

/* count_threshold of winSketch.hpp:298-349 */
static unsigned long long count_threshold(const std::vector<unsigned long long>& freqs, unsigned long long total_windows, double F) {
  const unsigned long long min_occ = 10;
  unsigned long long thr = (F <= 1.0) ? std::max<unsigned long long>(min_occ, (unsigned long long)(total_windows * F))
                                      : std::max<unsigned long long>(min_occ, (unsigned long long)F);
  unsigned long long wpos = 0, wuniq = 0;
  for (unsigned long long f : freqs) if (f > thr && f > min_occ) { wuniq++; wpos += f; }
  if (wpos > total_windows / 2 || wuniq > freqs.size() * 0.7) {
    std::vector<unsigned long long> all(freqs);
    std::sort(all.begin(), all.end());
    size_t keep = (size_t)(all.size() * 0.999);
    if (keep >= all.size()) keep = all.size() - 1;
    thr = std::max(thr, all[keep]);
  }
  return thr;
}

extern "C" void wfb_index_free(wfb_index_t* ix) {
  if (!ix) return;
#ifndef WFB_EMU
  cudaSetDevice(ix->device);
  cudaFree(ix->d_minmers); cudaFree(ix->d_points); cudaFree(ix->d_table); cudaFree(ix->d_uhash); cudaFree(ix->d_ustart); cudaFree(ix->d_ucount);
  for (int i = 0; i < 48; ++i) cudaFree(ix->ws_ptr[i]);
#else
  free(ix->d_minmers); free(ix->d_points); free(ix->d_table); free(ix->d_uhash); free(ix->d_ustart); free(ix->d_ucount);
#endif
  delete ix;
}

int wfb_minmers_build_impl(int device, const char* const* seq_ptrs, const int64_t* seq_lens, const int32_t* seq_ids, int32_t nseq,
                           int32_t kmer_size, int32_t window_size, int32_t sketch_size, wfb_minmer_t* out, int64_t out_cap,
                           int64_t* out_count, wfb_minmer_stats_t* stats, wfb_minmer_t** d_out_keep); /* minmer_host.cu */

extern "C" wfb_index_t* wfb_index_build(int device, const wfb_index_params_t* prm, const char* const* seq_ptrs, const int64_t* seq_lens,
                                        const int32_t* seq_ids, int32_t nseq, wfb_index_stats_t* stats) {
  if (!prm || nseq <= 0 || !seq_ptrs || !seq_lens || !seq_ids) { wfb_set_last_error_("bad argument"); return nullptr; }
  if (stats) memset(stats, 0, sizeof(*stats));
  const int k = prm->kmer_size, w = prm->window_size, s = prm->sketch_size;
  /* 1. windowed minmers of every target (addMinmers), in Sketch::build's order */
  long long total_len = 0;
  int max_seq_id = 0;
  for (int i = 0; i < nseq; ++i) { total_len += seq_lens[i]; max_seq_id = std::max(max_seq_id, seq_ids[i]); }
  if (max_seq_id >= (1 << 22)) { wfb_set_last_error_("seqId too large for the packed interval points (22 bits)"); return nullptr; }
  int64_t n = 0;
  wfb_minmer_stats_t ms;
  wfb_minmer_t* d_mi_built = nullptr; /* the minmers never leave the device between addMinmers and the index kernels */
#ifdef WFB_EMU
  long long cap = (long long)((double)total_len * (0.01 * s + 0.05)) + 4096;
  std::vector<wfb_minmer_t> mi((size_t)cap);
  int rc = wfb_minmers_build_impl(device, seq_ptrs, seq_lens, seq_ids, nseq, k, w, s, mi.data(), cap, &n, &ms, nullptr);
#else
  int rc = wfb_minmers_build_impl(device, seq_ptrs, seq_lens, seq_ids, nseq, k, w, s, nullptr, 0, &n, &ms, &d_mi_built);
#endif
  if (rc != WFB_OK) return nullptr;
  if (n == 0) {
#ifndef WFB_EMU
    cudaFree(d_mi_built);
#endif
    wfb_set_last_error_("reference sketch is empty (winSketch.hpp:451-456)");
    return nullptr;
  }
  /* worker partitions of Sketch::build (winSketch.hpp:271-277): contiguous ranges of the sequences >= w */
  std::vector<int> part((size_t)max_seq_id + 1, 0);
  {
    int nvalid = 0;
    for (int i = 0; i < nseq; ++i) if (seq_lens[i] >= w) ++nvalid;
    const int threads = std::max(1, prm->index_threads);
    const int chunk = (nvalid + threads - 1) / threads;
    int j = 0;
    for (int i = 0; i < nseq; ++i) if (seq_lens[i] >= w) { part[seq_ids[i]] = j / std::max(1, chunk); ++j; }
  }
  wfb_index* ix = new wfb_index();
  ix->device = device;
  ix->params = *prm;
  ix->n_minmers_all = n;
#ifndef WFB_EMU
  {
    wfb_minmer_t* d_mi = d_mi_built; unsigned long long *d_keys = nullptr, *d_skeys = nullptr, *d_ufreq = nullptr, *d_uhash_all = nullptr;
    int *d_idx = nullptr, *d_sidx = nullptr, *d_head = nullptr, *d_keep_s = nullptr, *d_keep_o = nullptr, *d_pstart = nullptr, *d_hk = nullptr, *d_part = nullptr;
    long long *d_run = nullptr, *d_pair = nullptr, *d_uk = nullptr, *d_off = nullptr;
    int* d_fail = nullptr;
    Tmp tmp;
    long long nuniq_all = 0, npairs = 0, nuk = 0, nkept = 0;
    std::vector<unsigned long long> hfreq;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const int G = 148 * 8, B = 256;
    IX_CHECK(cudaSetDevice(device));
    IX_CHECK(cudaEventCreate(&e0)); IX_CHECK(cudaEventCreate(&e1));
    IX_CHECK(cudaMalloc(&d_part, sizeof(int) * part.size()));
    IX_CHECK(cudaMemcpy(d_part, part.data(), sizeof(int) * part.size(), cudaMemcpyHostToDevice));
    IX_CHECK(cudaMalloc(&d_keys, 8 * (size_t)n)); IX_CHECK(cudaMalloc(&d_skeys, 8 * (size_t)n));
    IX_CHECK(cudaMalloc(&d_idx, 4 * (size_t)n)); IX_CHECK(cudaMalloc(&d_sidx, 4 * (size_t)n));
    IX_CHECK(cudaMalloc(&d_head, 4 * (size_t)n)); IX_CHECK(cudaMalloc(&d_keep_s, 4 * (size_t)n)); IX_CHECK(cudaMalloc(&d_keep_o, 4 * (size_t)n));
    IX_CHECK(cudaMalloc(&d_pstart, 4 * (size_t)n)); IX_CHECK(cudaMalloc(&d_hk, 4 * (size_t)n));
    IX_CHECK(cudaMalloc(&d_run, 8 * (size_t)n)); IX_CHECK(cudaMalloc(&d_pair, 8 * (size_t)n)); IX_CHECK(cudaMalloc(&d_uk, 8 * (size_t)n));
    IX_CHECK(cudaMalloc(&d_off, 8 * (size_t)n));
    IX_CHECK(cudaMalloc(&d_fail, 4)); IX_CHECK(cudaMemset(d_fail, 0, 4));
    IX_CHECK(cudaEventRecord(e0));
    /* 2. stable sort by hash (postings keep the index order inside a hash) */
    IX_LAUNCH(ix_hash_keys_kernel, G, B, 0, d_mi, n, d_keys, d_idx);
    {
      size_t b = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, b, d_keys, d_skeys, d_idx, d_sidx, (int)n);
      IX_CHECK(tmp.ensure(b));
      b = tmp.cap;
      IX_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, b, d_keys, d_skeys, d_idx, d_sidx, (int)n));
    }
    /* 3. frequencies (:266-296) */
    IX_LAUNCH(ix_head_flags_kernel, G, B, 0, d_skeys, n, d_head);
    IX_CHECK(incl_scan(tmp, d_head, d_run, n));
    IX_CHECK(cudaMemcpy(&nuniq_all, d_run + (n - 1), 8, cudaMemcpyDeviceToHost));
    IX_CHECK(cudaMalloc(&d_ufreq, 8 * (size_t)nuniq_all)); IX_CHECK(cudaMalloc(&d_uhash_all, 8 * (size_t)nuniq_all));
    IX_CHECK(cudaMemset(d_ufreq, 0, 8 * (size_t)nuniq_all));
    IX_LAUNCH(ix_run_freq_kernel, G, B, 0, d_head, d_run, n, d_ufreq, d_skeys, d_uhash_all);
    hfreq.resize((size_t)nuniq_all);
    IX_CHECK(cudaMemcpy(hfreq.data(), d_ufreq, 8 * (size_t)nuniq_all, cudaMemcpyDeviceToHost));
    /* 4. threshold (:298-349) and keep flags (:374-377) */
    ix->threshold = count_threshold(hfreq, (unsigned long long)n, prm->max_kmer_freq);
    IX_LAUNCH(ix_keep_kernel, G, B, 0, d_run, d_ufreq, d_sidx, n, ix->threshold, d_keep_s, d_keep_o);
    /* 5. postings (:379-387) */
    IX_LAUNCH(ix_pair_start_kernel, G, B, 0, d_mi, d_sidx, d_skeys, d_keep_s, d_part, n, d_pstart);
    IX_CHECK(incl_scan(tmp, d_pstart, d_pair, n));
    IX_CHECK(cudaMemcpy(&npairs, d_pair + (n - 1), 8, cudaMemcpyDeviceToHost));
    ix->n_points = 2 * npairs;
    if (ix->n_points >= (1LL << 32)) { wfb_set_last_error_("too many interval points for 32-bit postings offsets"); rc = WFB_EINVAL; goto done; }
    IX_CHECK(cudaMalloc(&ix->d_points, 8 * (size_t)std::max<long long>(ix->n_points, 1)));
    IX_LAUNCH(ix_points_kernel, G, B, 0, d_mi, d_sidx, d_keep_s, d_pstart, d_pair, n, ix->d_points);
    /* 6. kept unique hashes -> (start, count) */
    IX_LAUNCH(ix_and_kernel, G, B, 0, d_head, d_keep_s, n, d_hk);
    IX_CHECK(incl_scan(tmp, d_hk, d_uk, n));
    IX_CHECK(cudaMemcpy(&nuk, d_uk + (n - 1), 8, cudaMemcpyDeviceToHost));
    ix->n_uniq = nuk;
    IX_CHECK(cudaMalloc(&ix->d_uhash, 8 * (size_t)std::max<long long>(nuk, 1)));
    IX_CHECK(cudaMalloc(&ix->d_ustart, 4 * (size_t)std::max<long long>(nuk, 1)));
    IX_CHECK(cudaMalloc(&ix->d_ucount, 4 * (size_t)std::max<long long>(nuk, 1)));
    IX_CHECK(cudaMemset(ix->d_ucount, 0, 4 * (size_t)std::max<long long>(nuk, 1)));
    IX_LAUNCH(ix_uniq_kernel, G, B, 0, d_head, d_keep_s, d_pstart, d_pair, d_uk, d_skeys, n, ix->d_uhash, ix->d_ustart, ix->d_ucount);
    /* 7. open-addressing table, load <= 0.5 */
    {
      long long nb = 1;
      while (nb * IX_BUCKET < 2 * std::max<long long>(nuk, 1)) nb <<= 1;
      ix->n_buckets = nb;
      IX_CHECK(cudaMalloc(&ix->d_table, sizeof(IxSlot) * (size_t)nb * IX_BUCKET));
      IX_LAUNCH(ix_fill_kernel, G, B, 0, (unsigned long long*)ix->d_table, nb * IX_BUCKET * 2, (unsigned long long)IX_EMPTY);
      IX_LAUNCH(ix_insert_kernel, G, B, 0, ix->d_table, nb, (const uint64_t*)ix->d_uhash, ix->d_ustart, ix->d_ucount, nuk, d_fail);
      int fail = 0;
      IX_CHECK(cudaMemcpy(&fail, d_fail, 4, cudaMemcpyDeviceToHost));
      if (fail) { wfb_set_last_error_("hash table insert failed"); rc = WFB_ECUDA; goto done; }
    }
    /* 8. kept minmerIndex in reference order (:389,424-429) */
    IX_CHECK(excl_scan(tmp, d_keep_o, d_off, n));
    {
      long long last_off = 0; int last_keep = 0;
      IX_CHECK(cudaMemcpy(&last_off, d_off + (n - 1), 8, cudaMemcpyDeviceToHost));
      IX_CHECK(cudaMemcpy(&last_keep, d_keep_o + (n - 1), 4, cudaMemcpyDeviceToHost));
      nkept = last_off + last_keep;
    }
    ix->n_minmers = nkept;
    IX_CHECK(cudaMalloc(&ix->d_minmers, sizeof(wfb_minmer_t) * (size_t)std::max<long long>(nkept, 1)));
    IX_LAUNCH(ix_compact_minmers_kernel, G, B, 0, d_mi, d_keep_o, d_off, n, ix->d_minmers);
    IX_CHECK(cudaEventRecord(e1));
    IX_CHECK(cudaEventSynchronize(e1));
    IX_CHECK(cudaGetLastError());
    if (stats) {
      float t = 0;
      cudaEventElapsedTime(&t, e0, e1);
      stats->index_kernel_ms = t;
    }
  done:
    if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1);
    cudaFree(d_mi); cudaFree(d_keys); cudaFree(d_skeys); cudaFree(d_ufreq); cudaFree(d_uhash_all); cudaFree(d_idx); cudaFree(d_sidx);
    cudaFree(d_head); cudaFree(d_keep_s); cudaFree(d_keep_o); cudaFree(d_pstart); cudaFree(d_hk); cudaFree(d_part); cudaFree(d_run);
    cudaFree(d_pair); cudaFree(d_uk); cudaFree(d_off); cudaFree(d_fail);
    if (rc != WFB_OK) { wfb_index_free(ix); return nullptr; }
  }
#else
  { wfb_set_last_error_("index build is not part of the host emulation"); wfb_index_free(ix); return nullptr; }
#endif
  if (stats) {
    stats->minmer = ms;
    stats->total_windows = (uint64_t)n; stats->kept_minmers = (uint64_t)ix->n_minmers; stats->interval_points = (uint64_t)ix->n_points;
    stats->unique_hashes = (uint64_t)ix->n_uniq; stats->count_threshold = ix->threshold; stats->table_buckets = (uint64_t)ix->n_buckets;
  }
  return ix;
}

/* An index that already exists in host memory (wfb_index_file_read, or wfb_index_export of another index): upload the
 * kept minmers, the per-hash postings and build the open-addressing table (step 7 of wfb_index_build). */
extern "C" wfb_index_t* wfb_index_import(int device, const wfb_index_params_t* prm, const wfb_index_view_t* v) {
  if (!prm || !v || prm->kmer_size < 1 || prm->window_size < 1 || prm->sketch_size < 1 || v->n_minmers <= 0 || v->n_uniq <= 0 || v->n_points <= 0 ||
      !v->minmers || !v->uhash || !v->ustart || !v->ucount || !v->points) {
    wfb_set_last_error_("bad argument (an empty index cannot be imported: the reference exits on an empty sketch, winSketch.hpp:451-456)");
    return nullptr;
  }
#ifndef WFB_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    wfb_set_last_error_("no CUDA device (this library has no CPU path)");
    return nullptr;
  }
  wfb_index* ix = new wfb_index();
  ix->device = device;
  ix->params = *prm;
  ix->n_minmers_all = ix->n_minmers = v->n_minmers;
  ix->n_points = v->n_points;
  ix->n_uniq = v->n_uniq;
  int rc = WFB_OK;
  int* d_fail = nullptr;
  {
    const long long nuk = v->n_uniq;
    const int B = 256, G = 148 * 8;
    IX_CHECK(cudaSetDevice(device));
    IX_CHECK(cudaMalloc(&ix->d_minmers, sizeof(wfb_minmer_t) * (size_t)v->n_minmers));
    IX_CHECK(cudaMemcpy(ix->d_minmers, v->minmers, sizeof(wfb_minmer_t) * (size_t)v->n_minmers, cudaMemcpyHostToDevice));
    IX_CHECK(cudaMalloc(&ix->d_points, 8 * (size_t)v->n_points));
    IX_CHECK(cudaMemcpy(ix->d_points, v->points, 8 * (size_t)v->n_points, cudaMemcpyHostToDevice));
    IX_CHECK(cudaMalloc(&ix->d_uhash, 8 * (size_t)nuk));
    IX_CHECK(cudaMalloc(&ix->d_ustart, 4 * (size_t)nuk));
    IX_CHECK(cudaMalloc(&ix->d_ucount, 4 * (size_t)nuk));
    IX_CHECK(cudaMemcpy(ix->d_uhash, v->uhash, 8 * (size_t)nuk, cudaMemcpyHostToDevice));
    IX_CHECK(cudaMemcpy(ix->d_ustart, v->ustart, 4 * (size_t)nuk, cudaMemcpyHostToDevice));
    IX_CHECK(cudaMemcpy(ix->d_ucount, v->ucount, 4 * (size_t)nuk, cudaMemcpyHostToDevice));
    IX_CHECK(cudaMalloc(&d_fail, 4));
    IX_CHECK(cudaMemset(d_fail, 0, 4));
    long long nb = 1;
    while (nb * IX_BUCKET < 2 * nuk) nb <<= 1;
    ix->n_buckets = nb;
    IX_CHECK(cudaMalloc(&ix->d_table, sizeof(IxSlot) * (size_t)nb * IX_BUCKET));
    IX_LAUNCH(ix_fill_kernel, G, B, 0, (unsigned long long*)ix->d_table, nb * IX_BUCKET * 2, (unsigned long long)IX_EMPTY);
    IX_LAUNCH(ix_insert_kernel, G, B, 0, ix->d_table, nb, (const uint64_t*)ix->d_uhash, ix->d_ustart, ix->d_ucount, nuk, d_fail);
    int fail = 0;
    IX_CHECK(cudaMemcpy(&fail, d_fail, 4, cudaMemcpyDeviceToHost));
    if (fail) { wfb_set_last_error_("hash table insert failed"); rc = WFB_ECUDA; goto done; }
  }
done:
  cudaFree(d_fail);
  if (rc != WFB_OK) { wfb_index_free(ix); return nullptr; }
  return ix;
#else
  (void)device;
  wfb_set_last_error_("index import is not part of the host emulation");
  return nullptr;
#endif
}

extern "C" int wfb_index_export(const wfb_index_t* ix, wfb_minmer_t* minmers, int64_t minmers_cap, uint64_t* uhash, uint32_t* ustart,
                                uint32_t* ucount, int64_t uniq_cap, uint64_t* points, int64_t points_cap) {
  if (!ix) { wfb_set_last_error_("index == NULL"); return WFB_EINVAL; }
  if ((minmers && minmers_cap < ix->n_minmers) || (uhash && uniq_cap < ix->n_uniq) || (points && points_cap < ix->n_points)) {
    wfb_set_last_error_("export buffer too small");
    return WFB_ECAP;
  }
#ifndef WFB_EMU
  int rc = WFB_OK;
  IX_CHECK(cudaSetDevice(ix->device));
  if (minmers && ix->n_minmers) IX_CHECK(cudaMemcpy(minmers, ix->d_minmers, sizeof(wfb_minmer_t) * (size_t)ix->n_minmers, cudaMemcpyDeviceToHost));
  if (uhash && ix->n_uniq) {
    IX_CHECK(cudaMemcpy(uhash, ix->d_uhash, 8 * (size_t)ix->n_uniq, cudaMemcpyDeviceToHost));
    IX_CHECK(cudaMemcpy(ustart, ix->d_ustart, 4 * (size_t)ix->n_uniq, cudaMemcpyDeviceToHost));
    IX_CHECK(cudaMemcpy(ucount, ix->d_ucount, 4 * (size_t)ix->n_uniq, cudaMemcpyDeviceToHost));
  }
  if (points && ix->n_points) IX_CHECK(cudaMemcpy(points, ix->d_points, 8 * (size_t)ix->n_points, cudaMemcpyDeviceToHost));
done:
  return rc;
#else
  return WFB_ENODEV;
#endif
}

extern "C" int wfb_index_sizes(const wfb_index_t* ix, int64_t* n_minmers, int64_t* n_uniq, int64_t* n_points, uint64_t* threshold) {
  if (!ix) return WFB_EINVAL;
  if (n_minmers) *n_minmers = ix->n_minmers;
  if (n_uniq) *n_uniq = ix->n_uniq;
  if (n_points) *n_points = ix->n_points;
  if (threshold) *threshold = ix->threshold;
  return WFB_OK;
}

/* ---- L1 (+ L2) over a batch of fragments ------------------------------------------------------------------- */
static int l1_validate(const wfb_index_t* ix, const wfb_l1_params_t* lp, const char* seq_base, int64_t seq_bytes, const wfb_frag_t* frags,
                       const wfb_frag_query_t* fq, int32_t n) {
  if (!ix || !lp || n < 0 || (n > 0 && (!seq_base || !frags || !fq))) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  if (!lp->ref_group || !lp->sketch_cutoffs || lp->n_cutoffs <= 0 || lp->minimum_hits < 1) { wfb_set_last_error_("bad L1 parameters"); return WFB_EINVAL; }
  const int w = ix->params.window_size, s = ix->params.sketch_size;
  if (s > 512) { wfb_set_last_error_("sketch_size > 512 unsupported by the L1 kernel"); return WFB_EINVAL; }
  for (int i = 0; i < n; ++i) {
    if (frags[i].len != w) { wfb_set_last_error_("L1 kernel handles fragments of length == window_size (windowLen == 0) only"); return WFB_EINVAL; }
    if (frags[i].seq_offset < 0 || frags[i].seq_offset + frags[i].len > seq_bytes) { wfb_set_last_error_("fragment out of range"); return WFB_EINVAL; }
  }
  return WFB_OK;
}

/* Q.kmerComplexity exactly as MappingCore::getSeedHits computes it (src/map/include/mappingCore.hpp:72-74): the first
 * division is done in long double on the host, like the reference's. */
static float ix_kmer_complexity(unsigned long long max_hash, int qn, int len, int k) {
  if (qn <= 0) return 0.f;
  const double max_hash_01 = (long double)(max_hash) / std::numeric_limits<uint64_t>::max();
  return (double(qn) / max_hash_01) / ((len - k + 1) * 2);
}

#ifndef WFB_EMU
namespace {
/* per-fragment Q.kmerComplexity from what the L1 kernel left on the device */
int l1_host_complexity(const unsigned long long* d_qmax, const int* d_qn, int32_t n, int len, int k, std::vector<float>& kc, std::vector<int32_t>& qn) {
  int rc = WFB_OK;
  std::vector<unsigned long long> mx((size_t)n);
  qn.resize((size_t)n);
  kc.resize((size_t)n);
  IX_CHECK(cudaMemcpy(mx.data(), d_qmax, 8 * (size_t)n, cudaMemcpyDeviceToHost));
  IX_CHECK(cudaMemcpy(qn.data(), d_qn, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  for (int32_t i = 0; i < n; ++i) kc[(size_t)i] = ix_kmer_complexity(mx[(size_t)i], qn[(size_t)i], len, k);
done:
  return rc;
}
/* slot `slot` of the index's grow-only workspace, at least `bytes` big (contents are NOT preserved when it grows) */
template <typename T>
cudaError_t ws_get(const wfb_index_t* ix, int slot, size_t bytes, T** out) {
  if (bytes < 256) bytes = 256;
  if (ix->ws_cap[slot] < bytes) {
    if (ix->ws_ptr[slot]) cudaFree(ix->ws_ptr[slot]);
    ix->ws_ptr[slot] = nullptr; ix->ws_cap[slot] = 0;
    const size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&ix->ws_ptr[slot], want);
    if (e != cudaSuccess) return e;
    ix->ws_cap[slot] = want;
  }
  *out = (T*)ix->ws_ptr[slot];
  return cudaSuccess;
}
enum { WS_SEQ, WS_FRAGS, WS_FQ, WS_GROUP, WS_CUT, WS_Q, WS_QN, WS_FN, WS_FST, WS_LFRAG, WS_KC, WS_GS, WS_LTMP, WS_LOCI, WS_LC, WS_FOFF, WS_QMAX,
       WS_S1, WS_MS, WS_SLAB, WS_MAP, WS_SORTED, WS_CNT, WS_CTR, WS_KP, WS_KP2, WS_KF, WS_KF2, WS_KF3, WS_IDX, WS_IDX2, WS_IDX3, WS_CUBTMP, WS_REDO };

struct L1Dev { /* what ix_l1_kernel leaves in device memory (views into the index's workspace); the L2 kernel reads it in place */
  uint8_t* d_seq = nullptr; wfb_frag_t* d_frags = nullptr; IxFragQuery* d_fq = nullptr; int *d_group = nullptr, *d_cut = nullptr;
  wfb_minmer_t* d_q = nullptr; int *d_qn = nullptr, *d_fn = nullptr, *d_fst = nullptr, *d_lfrag = nullptr; float* d_kc = nullptr;
  uint64_t* d_gs = nullptr; IxL1Locus *d_ltmp = nullptr, *d_loci = nullptr; unsigned long long *d_lc = nullptr, *d_qmax = nullptr; long long* d_foff = nullptr;
  long long loci_cap = 0;
  unsigned long long n_loci = 0;
  double kernel_ms = 0;
  int sm_count = 0;
};

int l1_run(const wfb_index_t* ix, const wfb_l1_params_t* lp, const char* seq_base, int64_t seq_bytes, const wfb_frag_t* frags,
           const wfb_frag_query_t* fq, int32_t n, long long loci_cap, L1Dev& D) {
  int rc = WFB_OK;
  const int k = ix->params.kmer_size, w = ix->params.window_size, s = ix->params.sketch_size;
  int npow2 = 1;
  while (npow2 < w - k + 1) npow2 <<= 1;
  const size_t sketch_smem = (size_t)npow2 * 12 + (((size_t)w + 15) & ~(size_t)15) + 16;
  IxL1Params P;
  P.k = k; P.w = w; P.s = s; P.minimum_hits = lp->minimum_hits;
  P.skip_self = lp->skip_self; P.skip_prefix = lp->skip_prefix; P.lower_triangular = lp->lower_triangular;
  P.ncut = lp->n_cutoffs; P.smem_cap = 4096; P.gcap = 1 << 17; P.max_loci = 256; P.complexity_threshold = lp->kmer_complexity_threshold;
  if (getenv("WFB_L1_FRAG_CAPS")) { /* tests: "gcap,max_loci" of the first pass, to force the redo pass below */
    int g = 0, m = 0;
    if (sscanf(getenv("WFB_L1_FRAG_CAPS"), "%d,%d", &g, &m) == 2 && g >= P.smem_cap && m >= 1) { P.gcap = g; P.max_loci = m; }
  }
  P.par_sweep = getenv("WFB_L1_SERIAL") ? 0 : 1;
  const size_t smem = std::max(std::max(sketch_smem, (size_t)P.smem_cap * 8), (size_t)IX_PAR_CAP * 8 + ix_par_aux_bytes());
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int grid = 0;
  D.loci_cap = loci_cap;
  IX_CHECK(cudaSetDevice(ix->device));
  {
    cudaDeviceProp prop;
    IX_CHECK(cudaGetDeviceProperties(&prop, ix->device));
    D.sm_count = prop.multiProcessorCount;
    grid = std::min(n, prop.multiProcessorCount * 4);
  }
  IX_CHECK(ws_get(ix, WS_SEQ, (size_t)seq_bytes + 16, &D.d_seq));
  IX_CHECK(cudaMemcpy(D.d_seq, seq_base, (size_t)seq_bytes, cudaMemcpyHostToDevice));
  IX_CHECK(ws_get(ix, WS_FRAGS, sizeof(wfb_frag_t) * (size_t)n, &D.d_frags));
  IX_CHECK(cudaMemcpy(D.d_frags, frags, sizeof(wfb_frag_t) * (size_t)n, cudaMemcpyHostToDevice));
  IX_CHECK(ws_get(ix, WS_FQ, sizeof(IxFragQuery) * (size_t)n, &D.d_fq));
  IX_CHECK(cudaMemcpy(D.d_fq, fq, sizeof(IxFragQuery) * (size_t)n, cudaMemcpyHostToDevice));
  IX_CHECK(ws_get(ix, WS_GROUP, 4 * (size_t)lp->n_ref_group, &D.d_group));
  IX_CHECK(cudaMemcpy(D.d_group, lp->ref_group, 4 * (size_t)lp->n_ref_group, cudaMemcpyHostToDevice));
  IX_CHECK(ws_get(ix, WS_CUT, 4 * (size_t)lp->n_cutoffs, &D.d_cut));
  IX_CHECK(cudaMemcpy(D.d_cut, lp->sketch_cutoffs, 4 * (size_t)lp->n_cutoffs, cudaMemcpyHostToDevice));
  IX_CHECK(ws_get(ix, WS_Q, sizeof(wfb_minmer_t) * (size_t)n * s, &D.d_q));
  IX_CHECK(cudaMemset(D.d_q, 0, sizeof(wfb_minmer_t) * (size_t)n * s));
  IX_CHECK(ws_get(ix, WS_QN, 4 * (size_t)n, &D.d_qn)); IX_CHECK(ws_get(ix, WS_FN, 4 * (size_t)n, &D.d_fn)); IX_CHECK(ws_get(ix, WS_FST, 4 * (size_t)n, &D.d_fst));
  IX_CHECK(ws_get(ix, WS_KC, 4 * (size_t)n, &D.d_kc)); IX_CHECK(ws_get(ix, WS_FOFF, 8 * (size_t)n, &D.d_foff)); IX_CHECK(ws_get(ix, WS_QMAX, 8 * (size_t)n, &D.d_qmax));
  IX_CHECK(ws_get(ix, WS_GS, 8 * (size_t)P.gcap * grid, &D.d_gs));
  IX_CHECK(ws_get(ix, WS_LTMP, sizeof(IxL1Locus) * (size_t)2 * P.max_loci * grid, &D.d_ltmp));
  IX_CHECK(ws_get(ix, WS_LOCI, sizeof(IxL1Locus) * (size_t)std::max<long long>(loci_cap, 1), &D.d_loci));
  IX_CHECK(ws_get(ix, WS_LFRAG, 4 * (size_t)std::max<long long>(loci_cap, 1), &D.d_lfrag));
  IX_CHECK(ws_get(ix, WS_LC, 8, &D.d_lc)); IX_CHECK(cudaMemset(D.d_lc, 0, 8));
  IX_CHECK(cudaFuncSetAttribute(ix_l1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  IX_CHECK(cudaEventCreate(&e0)); IX_CHECK(cudaEventCreate(&e1));
  IX_CHECK(cudaEventRecord(e0));
  IX_LAUNCH(ix_l1_kernel, grid, 128, smem, D.d_seq, D.d_frags, D.d_fq, n, npow2, P, ix->d_table, ix->n_buckets, ix->d_points, D.d_group, D.d_cut,
            D.d_q, D.d_qn, D.d_kc, D.d_qmax, D.d_gs, D.d_ltmp, D.d_loci, D.d_lfrag, D.d_lc, loci_cap, D.d_foff, D.d_fn, D.d_fst, (const int*)nullptr, 0);
  IX_CHECK(cudaEventRecord(e1));
  IX_CHECK(cudaEventSynchronize(e1));
  IX_CHECK(cudaGetLastError());
  {
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    D.kernel_ms = t;
  }
  IX_CHECK(cudaMemcpy(&D.n_loci, D.d_lc, 8, cudaMemcpyDeviceToHost));
  if ((long long)D.n_loci <= loci_cap) {
    /* Fragments that outgrew the per-fragment scratch of the first pass (> 2^17 interval points after the gather, or > 256 candidate
     * regions: a repeat-rich fragment, a pangenome of hundreds of haplotypes) are run again, alone, with 32x the scratch — the reference
     * has no such limit (mappingCore.hpp:88-215), so neither has the phase. Only if a fragment outgrows that too does its status stay set. */
    std::vector<int> hst((size_t)n), redo;
    IX_CHECK(cudaMemcpy(hst.data(), D.d_fst, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    for (int f = 0; f < n; ++f) if (hst[(size_t)f] == WFB_ECAP) redo.push_back(f);
    if (!redo.empty()) {
      IxL1Params P2 = P;
      P2.gcap = 1 << 22; P2.max_loci = 8192;
      const int grid2 = (int)std::min<size_t>(redo.size(), 16);
      int* d_list = nullptr;
      IX_CHECK(ws_get(ix, WS_REDO, 4 * redo.size(), &d_list));
      IX_CHECK(cudaMemcpy(d_list, redo.data(), 4 * redo.size(), cudaMemcpyHostToDevice));
      IX_CHECK(ws_get(ix, WS_GS, 8 * (size_t)P2.gcap * grid2, &D.d_gs));
      IX_CHECK(ws_get(ix, WS_LTMP, sizeof(IxL1Locus) * (size_t)2 * P2.max_loci * grid2, &D.d_ltmp));
      IX_CHECK(cudaEventRecord(e0));
      IX_LAUNCH(ix_l1_kernel, grid2, 128, smem, D.d_seq, D.d_frags, D.d_fq, n, npow2, P2, ix->d_table, ix->n_buckets, ix->d_points, D.d_group, D.d_cut,
                D.d_q, D.d_qn, D.d_kc, D.d_qmax, D.d_gs, D.d_ltmp, D.d_loci, D.d_lfrag, D.d_lc, loci_cap, D.d_foff, D.d_fn, D.d_fst, (const int*)d_list,
                (int)redo.size());
      IX_CHECK(cudaEventRecord(e1));
      IX_CHECK(cudaEventSynchronize(e1));
      IX_CHECK(cudaGetLastError());
      float t = 0;
      cudaEventElapsedTime(&t, e0, e1);
      D.kernel_ms += t;
      IX_CHECK(cudaMemcpy(&D.n_loci, D.d_lc, 8, cudaMemcpyDeviceToHost));
    }
  }
  if ((long long)D.n_loci > loci_cap) { wfb_set_last_error_("loci buffer too small"); rc = WFB_ECAP; goto done; }
done:
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  return rc;
}

int l1_copy_out(const L1Dev& D, int32_t n, int s, int w, int k, wfb_l1_out_t* out) {
  int rc = WFB_OK;
  const unsigned long long nl = D.n_loci;
  std::vector<IxL1Locus> hl;
  out->n_loci = (int64_t)nl;
  out->kernel_ms = D.kernel_ms;
  if (out->q_minmers) IX_CHECK(cudaMemcpy(out->q_minmers, D.d_q, sizeof(wfb_minmer_t) * (size_t)n * s, cudaMemcpyDeviceToHost));
  if (out->q_count) IX_CHECK(cudaMemcpy(out->q_count, D.d_qn, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  if (out->q_complexity) {
    std::vector<float> kc; std::vector<int32_t> qn;
    rc = l1_host_complexity(D.d_qmax, D.d_qn, n, w, k, kc, qn);
    if (rc != WFB_OK) return rc;
    memcpy(out->q_complexity, kc.data(), 4 * (size_t)n);
  }
  IX_CHECK(cudaMemcpy(out->frag_loci_offset, D.d_foff, 8 * (size_t)n, cudaMemcpyDeviceToHost));
  IX_CHECK(cudaMemcpy(out->frag_loci_count, D.d_fn, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  IX_CHECK(cudaMemcpy(out->frag_status, D.d_fst, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  hl.resize((size_t)nl);
  if (nl) IX_CHECK(cudaMemcpy(hl.data(), D.d_loci, sizeof(IxL1Locus) * (size_t)nl, cudaMemcpyDeviceToHost));
  for (unsigned long long i = 0; i < nl; ++i) {
    out->loci[i].seqId = hl[i].seqId; out->loci[i].intersectionSize = hl[i].intersectionSize;
    out->loci[i].rangeStartPos = hl[i].rangeStartPos; out->loci[i].rangeEndPos = hl[i].rangeEndPos;
  }
done:
  return rc;
}
}  // namespace
#endif

extern "C" int wfb_l1_batch(const wfb_index_t* ix, const wfb_l1_params_t* lp, const char* seq_base, int64_t seq_bytes, const wfb_frag_t* frags,
                            const wfb_frag_query_t* fq, int32_t n, wfb_l1_out_t* out) {
  if (!out) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  int rc = l1_validate(ix, lp, seq_base, seq_bytes, frags, fq, n);
  if (rc != WFB_OK) return rc;
  out->n_loci = 0;
  out->kernel_ms = 0;
  if (n == 0) return WFB_OK;
#ifndef WFB_EMU
  std::lock_guard<std::mutex> ws_lock(ix->ws_mu);
  L1Dev D;
  rc = l1_run(ix, lp, seq_base, seq_bytes, frags, fq, n, out->loci_cap, D);
  if (rc != WFB_OK) return rc;
  return l1_copy_out(D, n, ix->params.sketch_size, ix->params.window_size, ix->params.kmer_size, out);
#else
  wfb_set_last_error_("L1 is not part of the host emulation");
  return WFB_ENODEV;
#endif
}

/* ---- L2 ---------------------------------------------------------------------------------------------------- */
/* Stat::j2md / md2j (src/map/include/map_stats.hpp:56-80), with the reference's float / double mix */
static float l2_j2md(float j, int k) {
  if (j == 0) return 1.0f;
  if (j == 1) return 0.0f;
  const float mash_dist = 1 - std::pow(2 * j / (1 + j), 1.0 / k);
  return mash_dist;
}
static float l2_md2j(float d, int k) {
  const float sim = 1 - d;
  const float jaccard = std::pow((double)sim, (double)k) / (2 - std::pow((double)sim, (double)k));
  return jaccard;
}

extern "C" int wfb_stage1_min_hits(double hg_numerator, float ani_diff, int32_t kmer_size, int32_t sketch_size, int32_t* out) {
  if (!out || sketch_size < 1 || kmer_size < 1) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  out[0] = 0;
  for (int qs = 1; qs <= sketch_size; ++qs) { /* Map::doL2Mapping, computeMap.hpp:999-1012, for Q.sketchSize == qs */
    const double jaccardSimilarity = hg_numerator / qs;
    const double mash_dist = l2_j2md((float)jaccardSimilarity, kmer_size);
    const double cutoff_ani = std::max(0.0, (1 - mash_dist) - ani_diff);
    const double cutoff_j = l2_md2j((float)(1 - cutoff_ani), kmer_size);
    int v = 0;
    while (v <= qs && static_cast<double>(v) / qs < cutoff_j) ++v; /* first intersectionSize that is not "< cutoff_j" */
    out[qs] = v;
  }
  return WFB_OK;
}

extern "C" int wfb_l2_min_shared(float percentage_identity, int32_t kmer_size, int32_t sketch_size, int32_t* out) {
  if (!out || sketch_size < 1 || kmer_size < 1) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  out[0] = 0;
  for (int qs = 1; qs <= sketch_size; ++qs) { /* computeMap.hpp:1018-1024 with keep_low_pct_id == false */
    int v = 0;
    for (; v <= qs; ++v) {
      const float mash_dist = l2_j2md(1.0 * v / qs, kmer_size);
      const float nucIdentity = (1 - mash_dist);
      if (nucIdentity >= percentage_identity) break;
    }
    out[qs] = v; /* qs + 1 = nothing passes */
  }
  return WFB_OK;
}

/* nucIdentity of every mapping (computeMap.hpp:1018-1019), frag_map_offset CSR */
static void l2_finish_host(wfb_l2_mapping_t* m, int64_t nm, const int32_t* q_count, const float* q_complexity, int32_t n, int k, int s, int64_t* frag_map_offset) {
  std::vector<float> tab((size_t)(s + 1) * (s + 1), -1.f);
  for (int i = 0; i <= n; ++i) frag_map_offset[i] = 0;
  for (int64_t i = 0; i < nm; ++i) {
    const int qn = q_count[m[i].frag], sh = m[i].conservedSketches;
    float& t = tab[(size_t)qn * (s + 1) + std::min(std::max(sh, 0), s)];
    if (t < 0) { const float mash_dist = l2_j2md(1.0 * sh / qn, k); t = (1 - mash_dist); }
    m[i].nucIdentity = t;
    m[i].kmerComplexity = q_complexity[m[i].frag];
    frag_map_offset[m[i].frag + 1]++;
  }
  for (int i = 0; i < n; ++i) frag_map_offset[i + 1] += frag_map_offset[i];
}

extern "C" int wfb_map_fragments_batch(const wfb_index_t* ix, const wfb_l1_params_t* lp, const wfb_l2_params_t* l2p, const char* seq_base,
                                       int64_t seq_bytes, const wfb_frag_t* frags, const wfb_frag_query_t* fq, int32_t n, wfb_map_out_t* out) {
  if (!out || !l2p || !out->mappings || !out->frag_map_offset || !out->frag_status) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  int rc = l1_validate(ix, lp, seq_base, seq_bytes, frags, fq, n);
  if (rc != WFB_OK) return rc;
  const int s = ix->params.sketch_size;
  if ((l2p->stage1_min_hits && l2p->n_stage1_min_hits < s + 1) || (l2p->l2_min_shared && l2p->n_l2_min_shared < s + 1)) {
    wfb_set_last_error_("L2 tables need sketch_size + 1 entries");
    return WFB_EINVAL;
  }
  out->n_mappings = 0; out->l1_kernel_ms = 0; out->l2_kernel_ms = 0; out->sort_kernel_ms = 0; out->n_l1_loci = 0; out->l2_loci = 0; out->l2_steps = 0;
  if (out->l1) { out->l1->n_loci = 0; out->l1->kernel_ms = 0; }
  out->frag_map_offset[0] = 0;
  if (n == 0) return WFB_OK;
#ifndef WFB_EMU
  std::lock_guard<std::mutex> ws_lock(ix->ws_mu);
  L1Dev D;
  long long loci_cap = out->l1 ? out->l1->loci_cap : (64LL * n + 1024);
  if (!out->l1 && getenv("WFB_L1_LOCI_CAP0")) loci_cap = std::max<long long>(1, atoll(getenv("WFB_L1_LOCI_CAP0"))); /* tests: force the growth path */
  for (int attempt = 0;; ++attempt) {
    rc = l1_run(ix, lp, seq_base, seq_bytes, frags, fq, n, loci_cap, D);
    /* the library's own loci buffer grows to what the kernel counted (an all-vs-all of many haplotypes averages far more than
     * 64 candidate regions per fragment); a caller-provided buffer (out->l1) is the caller's to grow */
    if (rc == WFB_ECAP && !out->l1 && attempt < 3 && (long long)D.n_loci > loci_cap) { loci_cap = (long long)D.n_loci + (long long)D.n_loci / 8 + 1024; continue; }
    break;
  }
  if (rc != WFB_OK) return rc;
  if (out->l1) { rc = l1_copy_out(D, n, s, ix->params.window_size, ix->params.kmer_size, out->l1); if (rc != WFB_OK) return rc; }
  out->l1_kernel_ms = D.kernel_ms;
  out->n_l1_loci = D.n_loci;
  int *d_s1 = nullptr, *d_ms = nullptr; L2Entry* d_slab = nullptr; wfb_l2_mapping_t *d_map = nullptr, *d_sorted = nullptr;
  unsigned long long *d_cnt = nullptr, *d_kp = nullptr, *d_kp2 = nullptr; L2Counters* d_ctr = nullptr;
  unsigned int *d_kf = nullptr, *d_kf2 = nullptr, *d_kf3 = nullptr, *d_idx = nullptr, *d_idx2 = nullptr, *d_idx3 = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  void* d_cubtmp = nullptr;
  L2Counters hc{};
  unsigned long long nm = 0;
  std::vector<int32_t> hq;
  std::vector<float> hkc;
  L2Params P;
  P.k = ix->params.kmer_size; P.w = ix->params.window_size; P.s = s; P.vec_cap = 256; P.warps_per_cta = 8;
  const size_t smem = l2_warp_smem(s) * P.warps_per_cta;
  const long long warps_needed = (long long)std::max<unsigned long long>(D.n_loci, 1);
  const int grid = (int)std::min<long long>((warps_needed + P.warps_per_cta - 1) / P.warps_per_cta, (long long)D.sm_count * 4);
  const long long cap = out->mappings_cap;
  if (l2p->stage1_min_hits) {
    IX_CHECK(ws_get(ix, WS_S1, 4 * (size_t)(s + 1), &d_s1));
    IX_CHECK(cudaMemcpy(d_s1, l2p->stage1_min_hits, 4 * (size_t)(s + 1), cudaMemcpyHostToDevice));
  }
  if (l2p->l2_min_shared) {
    IX_CHECK(ws_get(ix, WS_MS, 4 * (size_t)(s + 1), &d_ms));
    IX_CHECK(cudaMemcpy(d_ms, l2p->l2_min_shared, 4 * (size_t)(s + 1), cudaMemcpyHostToDevice));
  }
  IX_CHECK(ws_get(ix, WS_SLAB, sizeof(L2Entry) * (size_t)P.vec_cap * P.warps_per_cta * grid, &d_slab));
  IX_CHECK(ws_get(ix, WS_MAP, sizeof(wfb_l2_mapping_t) * (size_t)std::max<long long>(cap, 1), &d_map));
  IX_CHECK(ws_get(ix, WS_CNT, 8, &d_cnt)); IX_CHECK(cudaMemset(d_cnt, 0, 8));
  IX_CHECK(ws_get(ix, WS_CTR, sizeof(L2Counters), &d_ctr)); IX_CHECK(cudaMemset(d_ctr, 0, sizeof(L2Counters)));
  IX_CHECK(cudaFuncSetAttribute(l2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  IX_CHECK(cudaEventCreate(&e0)); IX_CHECK(cudaEventCreate(&e1)); IX_CHECK(cudaEventCreate(&e2));
  IX_CHECK(cudaEventRecord(e0));
  IX_LAUNCH(l2_kernel, grid, 32 * P.warps_per_cta, smem, ix->d_minmers, ix->n_minmers, D.d_loci, D.d_lfrag, D.d_lc, D.loci_cap, D.d_q, D.d_qn, P,
            d_s1, d_ms, d_slab, d_map, d_cnt, cap, D.d_fst, d_ctr);
  IX_CHECK(cudaEventRecord(e1));
  IX_CHECK(cudaMemcpy(&nm, d_cnt, 8, cudaMemcpyDeviceToHost));
  IX_CHECK(cudaGetLastError());
  if ((long long)nm > cap) { wfb_set_last_error_("mappings buffer too small"); out->n_mappings = (int64_t)nm; rc = WFB_ECAP; goto done; }
  if (nm > 0) { /* order by (frag, refSeqId, refStartPos): two stable LSD radix passes */
    const long long N = (long long)nm;
    const int G = (int)std::min<long long>((N + 255) / 256, 148 * 8);
    int fbits = 1;
    while ((1LL << fbits) < n) ++fbits;
    size_t b1 = 0, b2 = 0;
    IX_CHECK(ws_get(ix, WS_KP, 8 * (size_t)N, &d_kp)); IX_CHECK(ws_get(ix, WS_KP2, 8 * (size_t)N, &d_kp2));
    IX_CHECK(ws_get(ix, WS_KF, 4 * (size_t)N, &d_kf)); IX_CHECK(ws_get(ix, WS_KF2, 4 * (size_t)N, &d_kf2)); IX_CHECK(ws_get(ix, WS_KF3, 4 * (size_t)N, &d_kf3));
    IX_CHECK(ws_get(ix, WS_IDX, 4 * (size_t)N, &d_idx)); IX_CHECK(ws_get(ix, WS_IDX2, 4 * (size_t)N, &d_idx2)); IX_CHECK(ws_get(ix, WS_IDX3, 4 * (size_t)N, &d_idx3));
    IX_CHECK(ws_get(ix, WS_SORTED, sizeof(wfb_l2_mapping_t) * (size_t)N, &d_sorted));
    IX_LAUNCH(l2_keys_kernel, G, 256, 0, d_map, N, d_kp, d_kf, d_idx);
    IX_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, b1, d_kp, d_kp2, d_idx, d_idx2, (int)N));
    IX_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, b2, d_kf2, d_kf3, d_idx2, d_idx3, (int)N, 0, fbits));
    IX_CHECK(ws_get(ix, WS_CUBTMP, std::max(b1, b2), &d_cubtmp));
    IX_CHECK(cub::DeviceRadixSort::SortPairs(d_cubtmp, b1, d_kp, d_kp2, d_idx, d_idx2, (int)N));
    wfb_count_launch_();
    IX_LAUNCH(l2_gather_u32_kernel, G, 256, 0, d_kf, d_idx2, N, d_kf2);
    IX_CHECK(cub::DeviceRadixSort::SortPairs(d_cubtmp, b2, d_kf2, d_kf3, d_idx2, d_idx3, (int)N, 0, fbits));
    wfb_count_launch_();
    IX_LAUNCH(l2_permute_kernel, G, 256, 0, d_map, d_idx3, N, d_sorted);
    IX_CHECK(cudaEventRecord(e2));
    IX_CHECK(cudaMemcpy(out->mappings, d_sorted, sizeof(wfb_l2_mapping_t) * (size_t)N, cudaMemcpyDeviceToHost));
  } else {
    IX_CHECK(cudaEventRecord(e2));
  }
  IX_CHECK(cudaEventSynchronize(e2));
  IX_CHECK(cudaGetLastError());
  {
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1); out->l2_kernel_ms = t;
    cudaEventElapsedTime(&t, e1, e2); out->sort_kernel_ms = t;
  }
  IX_CHECK(cudaMemcpy(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost));
  out->l2_loci = hc.loci; out->l2_steps = hc.steps;
  IX_CHECK(cudaMemcpy(out->frag_status, D.d_fst, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  rc = l1_host_complexity(D.d_qmax, D.d_qn, n, P.w, P.k, hkc, hq);
  if (rc != WFB_OK) goto done;
  out->n_mappings = (int64_t)nm;
  l2_finish_host(out->mappings, (int64_t)nm, hq.data(), hkc.data(), n, P.k, s, out->frag_map_offset);
done:
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e2) cudaEventDestroy(e2);
  return rc;
#else
  wfb_set_last_error_("the mapping batch is not part of the host emulation (see wfb_emu_l2_loci)");
  return WFB_ENODEV;
#endif
}

#ifdef WFB_EMU
/* TEST-ONLY (never in the product library): runs l2_kernel's body under the single-thread emulation over host arrays,
 * so the kernel's bookkeeping can be checked against the oracle in a container without a GPU. */
extern "C" int wfb_emu_l2_loci(const wfb_minmer_t* index, int64_t n_index, const wfb_l1_locus_t* loci, const int32_t* locus_frag, int64_t n_loci,
                               const wfb_minmer_t* q_all, const int32_t* q_count, int32_t n_frags, int32_t k, int32_t w, int32_t s,
                               const int32_t* stage1_min_hits, const int32_t* l2_min_shared, wfb_l2_mapping_t* out, int64_t out_cap,
                               int64_t* n_out, int32_t* frag_status, uint64_t* steps) {
  std::vector<IxL1Locus> L((size_t)n_loci);
  for (int64_t i = 0; i < n_loci; ++i) { L[i].seqId = loci[i].seqId; L[i].intersectionSize = loci[i].intersectionSize; L[i].rangeStartPos = loci[i].rangeStartPos; L[i].rangeEndPos = loci[i].rangeEndPos; }
  L2Params P; P.k = k; P.w = w; P.s = s; P.vec_cap = 256; P.warps_per_cta = 1;
  const int grid = 3;
  std::vector<L2Entry> slab((size_t)P.vec_cap * grid);
  std::vector<unsigned char> smem(l2_warp_smem(s) + 64);
  unsigned long long nl = (unsigned long long)n_loci, cnt = 0;
  L2Counters ctr{};
  for (int b = 0; b < grid; ++b)
    l2_kernel(b, grid, index, n_index, L.data(), locus_frag, &nl, n_loci, q_all, q_count, P, stage1_min_hits, l2_min_shared, slab.data(), out, &cnt,
              out_cap, frag_status, &ctr, smem.data());
  *n_out = (int64_t)cnt;
  if (steps) *steps = ctr.steps;
  if ((int64_t)cnt > out_cap) return WFB_ECAP;
  std::stable_sort(out, out + cnt, [](const wfb_l2_mapping_t& a, const wfb_l2_mapping_t& b) {
    if (a.frag != b.frag) return a.frag < b.frag;
    if (a.refSeqId != b.refSeqId) return a.refSeqId < b.refSeqId;
    return a.refStartPos < b.refStartPos;
  });
  std::vector<int64_t> off((size_t)n_frags + 1);
  std::vector<float> kc((size_t)n_frags);
  for (int32_t i = 0; i < n_frags; ++i) kc[(size_t)i] = q_count[i] > 0 ? ix_kmer_complexity(q_all[(size_t)i * s + q_count[i] - 1].hash, q_count[i], w, k) : 0.f;
  l2_finish_host(out, (int64_t)cnt, q_count, kc.data(), n_frags, k, s, off.data());
  return WFB_OK;
}
#endif

#ifdef WFB_EMU
/* TEST-ONLY: the serial L1 walk (ix_l1_regions per group slice, as the kernel's fallback runs it) and the data-parallel
 * sweep (ix_l1_regions_par) over the same sorted packed keys, under the single-thread emulation. */
extern "C" int wfb_emu_l1_sweeps(const uint64_t* keys, int32_t n, int32_t q_sketch, int32_t w, int32_t s, int32_t minimum_hits, int32_t skip_prefix,
                                 const int32_t* cutoffs, int32_t ncut, const int32_t* ref_group, wfb_l1_locus_t* out_serial, int32_t* n_serial,
                                 wfb_l1_locus_t* out_par, int32_t* n_par) {
  if (n > IX_PAR_CAP) return WFB_EINVAL;
  IxL1Params P{};
  P.k = 15; P.w = w; P.s = s; P.minimum_hits = minimum_hits; P.skip_prefix = skip_prefix; P.ncut = ncut; P.max_loci = 256; P.par_sweep = 1;
  std::vector<IxL1Locus> ltmp((size_t)2 * P.max_loci), ltmp2((size_t)2 * P.max_loci);
  int nout = 0, err = 0, b = 0;
  while (b < n) {
    int e = n;
    if (skip_prefix) { e = b; const int g = ref_group[ix_seq(keys[b])]; while (e < n && ref_group[ix_seq(keys[e])] == g) ++e; }
    ix_l1_regions(keys + b, e - b, minimum_hits, q_sketch, P, cutoffs, ltmp.data(), nout, ltmp.data() + P.max_loci, P.max_loci, err);
    b = e;
  }
  if (err) return WFB_ECAP;
  *n_serial = nout;
  for (int i = 0; i < nout; ++i) { out_serial[i].seqId = ltmp[i].seqId; out_serial[i].intersectionSize = ltmp[i].intersectionSize; out_serial[i].rangeStartPos = ltmp[i].rangeStartPos; out_serial[i].rangeEndPos = ltmp[i].rangeEndPos; }
  std::vector<unsigned char> aux(ix_par_aux_bytes() + 64);
  static IxParShared S;
  const IxParAux A = ix_par_carve(aux.data());
  if (!ix_l1_regions_par(keys, n, q_sketch, P, cutoffs, ref_group, A, S, ltmp2.data(), ltmp2.data() + P.max_loci, P.max_loci)) return WFB_ECAP;
  if (S.err) return WFB_ECAP;
  *n_par = S.nout;
  for (int i = 0; i < S.nout; ++i) { out_par[i].seqId = ltmp2[i].seqId; out_par[i].intersectionSize = ltmp2[i].intersectionSize; out_par[i].rangeStartPos = ltmp2[i].rangeStartPos; out_par[i].rangeEndPos = ltmp2[i].rangeEndPos; }
  return WFB_OK;
}
#endif
