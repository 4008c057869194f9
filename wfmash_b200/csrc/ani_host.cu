// ani_host.cu — ANI auto-identity: skch::Stat::estimate_identity_for_groups
// (src/map/include/map_stats.hpp:325-822) as called by src/interface/main.cpp:75-134 when -p is not given.
//   wfb_ani_group_sketches : per-group bottom-s multiset MinHash of the canonical k-mer hashes on the GPU (ani_kernels.h)
//   wfb_ani_estimate_identity : the pairwise group comparison, percentile and adjustment on the host (a few hundred
//                               merge walks over 4096 hashes; the reference does this part on the host too)
#include "ani_kernels.h"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#ifndef WFB_EMU
#include <cub/cub.cuh>
#endif
#include "wfb_pool.h" /* this file's cudaMalloc / cudaFree go through the library's device-memory pool */

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */
void wfb_count_launch_();

#ifndef WFB_EMU
#define ANI_CHECK(call)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      wfb_set_last_error_(std::string(#call) + ": " + cudaGetErrorString(e_));           \
      rc = (e_ == cudaErrorMemoryAllocation) ? WFB_ENOMEM : WFB_ECUDA;                   \
      goto done;                                                                         \
    }                                                                                    \
  } while (0)
#endif

extern "C" int wfb_ani_group_sketches(int device, const char* const* seq_ptrs, const int64_t* seq_lens, const int32_t* seq_group, int32_t nseq,
                                      int32_t n_groups, int32_t kmer_size, int32_t sketch_size, uint64_t* sketches, int32_t* sketch_count,
                                      wfb_ani_stats_t* stats) {
  if (nseq < 0 || n_groups <= 0 || kmer_size < 1 || kmer_size > 32 || sketch_size < 1 || !sketches || !sketch_count ||
      (nseq > 0 && (!seq_ptrs || !seq_lens || !seq_group))) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  const int k = kmer_size;
  std::vector<AniTile> tiles;
  std::vector<int64_t> off((size_t)nseq + 1, 0);
  std::vector<unsigned long long> npos((size_t)n_groups, 0);
  for (int32_t i = 0; i < nseq; ++i) {
    if (seq_lens[i] < 0 || seq_group[i] < 0 || seq_group[i] >= n_groups) { wfb_set_last_error_("bad sequence length or group"); return WFB_EINVAL; }
    off[(size_t)i + 1] = off[(size_t)i] + seq_lens[i];
    const int64_t n = seq_lens[i] - k + 1;
    if (n <= 0) continue;
    int32_t head_bad = 0; /* map_stats.hpp:565-571: scan of the first min(k, len) bases */
    for (int j = 0; j < k; ++j) {
      char c = seq_ptrs[i][j];
      if (c > 96 && c < 123) c -= 32;
      if (c != 'A' && c != 'C' && c != 'G' && c != 'T') { head_bad = 1; break; }
    }
    for (int64_t st = 0; st < n; st += ANI_TILE)
      tiles.push_back(AniTile{off[(size_t)i], seq_lens[i], st, (int32_t)std::min<int64_t>(ANI_TILE, n - st), seq_group[i], head_bad, 0});
    npos[(size_t)seq_group[i]] += (unsigned long long)n;
  }
  const int64_t blob_bytes = off[(size_t)nseq];
  /* threshold and candidate capacity per group */
  const unsigned long long C = 8ULL * (unsigned long long)sketch_size;
  std::vector<uint64_t> thr((size_t)n_groups);
  std::vector<unsigned long long> cap((size_t)n_groups), base((size_t)n_groups + 1, 0), cnt((size_t)n_groups, 0);
  for (int g = 0; g < n_groups; ++g) {
    if (npos[(size_t)g] <= 4 * C) { thr[(size_t)g] = ~0ULL; cap[(size_t)g] = npos[(size_t)g]; }
    else { thr[(size_t)g] = (uint64_t)ldexpl((long double)C / (long double)npos[(size_t)g], 64); cap[(size_t)g] = 4 * C; }
  }
  int rc = WFB_OK;
  int passes = 0;
  unsigned long long n_valid = 0;
  double hash_ms = 0.0, sort_ms = 0.0;
  memset(sketches, 0, sizeof(uint64_t) * (size_t)n_groups * (size_t)sketch_size);
#ifndef WFB_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    wfb_set_last_error_("no CUDA device (this library has no CPU path)");
    return WFB_ENODEV;
  }
  uint8_t* d_blob = nullptr;
  AniTile* d_tiles = nullptr;
  uint64_t *d_thr = nullptr, *d_cand = nullptr, *d_sorted = nullptr, *d_sk = nullptr;
  unsigned long long *d_base = nullptr, *d_cap = nullptr, *d_cnt = nullptr, *d_end = nullptr, *d_valid = nullptr;
  void* d_tmp = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  ANI_CHECK(cudaSetDevice(device));
  ANI_CHECK(cudaMalloc(&d_blob, (size_t)blob_bytes + 64));
  ANI_CHECK(cudaMemset(d_blob + blob_bytes, 'N', 64));
  for (int32_t i = 0; i < nseq; ++i)
    if (seq_lens[i] > 0) ANI_CHECK(cudaMemcpy(d_blob + off[(size_t)i], seq_ptrs[i], (size_t)seq_lens[i], cudaMemcpyHostToDevice));
  ANI_CHECK(cudaMalloc(&d_tiles, sizeof(AniTile) * std::max<size_t>(tiles.size(), 1)));
  ANI_CHECK(cudaMemcpy(d_tiles, tiles.data(), sizeof(AniTile) * tiles.size(), cudaMemcpyHostToDevice));
  ANI_CHECK(cudaMalloc(&d_thr, 8 * (size_t)n_groups));
  ANI_CHECK(cudaMalloc(&d_base, 8 * ((size_t)n_groups + 1)));
  ANI_CHECK(cudaMalloc(&d_cap, 8 * (size_t)n_groups));
  ANI_CHECK(cudaMalloc(&d_cnt, 8 * (size_t)n_groups));
  ANI_CHECK(cudaMalloc(&d_end, 8 * (size_t)n_groups));
  ANI_CHECK(cudaMalloc(&d_valid, 8));
  ANI_CHECK(cudaMalloc(&d_sk, 8 * (size_t)n_groups * (size_t)sketch_size));
  ANI_CHECK(cudaEventCreate(&e0)); ANI_CHECK(cudaEventCreate(&e1)); ANI_CHECK(cudaEventCreate(&e2));
  {
    cudaDeviceProp prop;
    ANI_CHECK(cudaGetDeviceProperties(&prop, device));
    const int grid = (int)std::max<size_t>(1, std::min<size_t>(tiles.size(), (size_t)prop.multiProcessorCount * 8));
    for (;;) {
      ++passes;
      for (int g = 0; g < n_groups; ++g) base[(size_t)g + 1] = base[(size_t)g] + cap[(size_t)g];
      const size_t total = (size_t)base[(size_t)n_groups];
      cudaFree(d_cand); cudaFree(d_sorted); cudaFree(d_tmp); d_cand = d_sorted = nullptr; d_tmp = nullptr;
      ANI_CHECK(cudaMalloc(&d_cand, 8 * std::max<size_t>(total, 1)));
      ANI_CHECK(cudaMalloc(&d_sorted, 8 * std::max<size_t>(total, 1)));
      ANI_CHECK(cudaMemcpy(d_thr, thr.data(), 8 * (size_t)n_groups, cudaMemcpyHostToDevice));
      ANI_CHECK(cudaMemcpy(d_base, base.data(), 8 * ((size_t)n_groups + 1), cudaMemcpyHostToDevice));
      ANI_CHECK(cudaMemcpy(d_cap, cap.data(), 8 * (size_t)n_groups, cudaMemcpyHostToDevice));
      ANI_CHECK(cudaMemset(d_cnt, 0, 8 * (size_t)n_groups));
      ANI_CHECK(cudaMemset(d_valid, 0, 8));
      ANI_CHECK(cudaEventRecord(e0, 0));
      if (!tiles.empty()) {
        ani_hash_kernel<<<grid, ANI_THREADS, 0, 0>>>(d_blob, d_tiles, (int)tiles.size(), k, d_thr, d_base, d_cap, d_cnt, d_cand, d_valid);
        wfb_count_launch_();
      }
      ANI_CHECK(cudaEventRecord(e1, 0));
      ANI_CHECK(cudaGetLastError());
      ANI_CHECK(cudaMemcpy(cnt.data(), d_cnt, 8 * (size_t)n_groups, cudaMemcpyDeviceToHost));
      ANI_CHECK(cudaMemcpy(&n_valid, d_valid, 8, cudaMemcpyDeviceToHost));
      { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); hash_ms += ms; }
      bool again = false;
      for (int g = 0; g < n_groups; ++g) {
        if (cnt[(size_t)g] > cap[(size_t)g]) { cap[(size_t)g] = cnt[(size_t)g]; again = true; }                 /* duplicated small hashes */
        else if (cnt[(size_t)g] < (unsigned long long)sketch_size && thr[(size_t)g] != ~0ULL) { thr[(size_t)g] = ~0ULL; cap[(size_t)g] = npos[(size_t)g]; again = true; }
      }
      if (again && passes < 4) continue;
      if (again) { wfb_set_last_error_("ANI candidate selection did not converge"); rc = WFB_ECAP; goto done; }
      /* order every group's candidates: one segmented radix sort */
      std::vector<unsigned long long> endv((size_t)n_groups);
      for (int g = 0; g < n_groups; ++g) endv[(size_t)g] = base[(size_t)g] + cnt[(size_t)g];
      ANI_CHECK(cudaMemcpy(d_end, endv.data(), 8 * (size_t)n_groups, cudaMemcpyHostToDevice));
      size_t tmp_bytes = 0;
      ANI_CHECK(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tmp_bytes, d_cand, d_sorted, (int64_t)total, n_groups, d_base, d_end));
      ANI_CHECK(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
      ANI_CHECK(cub::DeviceSegmentedRadixSort::SortKeys(d_tmp, tmp_bytes, d_cand, d_sorted, (int64_t)total, n_groups, d_base, d_end));
      wfb_count_launch_();
      ani_gather_kernel<<<std::min(n_groups, 1024), 256, 0, 0>>>(d_sorted, d_base, d_cnt, n_groups, sketch_size, d_sk);
      wfb_count_launch_();
      ANI_CHECK(cudaEventRecord(e2, 0));
      ANI_CHECK(cudaGetLastError());
      ANI_CHECK(cudaMemcpy(sketches, d_sk, 8 * (size_t)n_groups * (size_t)sketch_size, cudaMemcpyDeviceToHost));
      { float ms = 0; cudaEventElapsedTime(&ms, e1, e2); sort_ms += ms; }
      break;
    }
  }
done:
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e2) cudaEventDestroy(e2);
  cudaFree(d_blob); cudaFree(d_tiles); cudaFree(d_thr); cudaFree(d_cand); cudaFree(d_sorted); cudaFree(d_sk);
  cudaFree(d_base); cudaFree(d_cap); cudaFree(d_cnt); cudaFree(d_end); cudaFree(d_valid); cudaFree(d_tmp);
  if (rc != WFB_OK) return rc;
#else
  (void)device;
  { /* single-thread emulation of the same kernels (tests/emu), std::sort in place of the segmented radix sort */
    std::vector<uint8_t> blob((size_t)blob_bytes + 64, (uint8_t)'N');
    for (int32_t i = 0; i < nseq; ++i) if (seq_lens[i] > 0) memcpy(blob.data() + off[(size_t)i], seq_ptrs[i], (size_t)seq_lens[i]);
    std::vector<unsigned char> sm(ANI_SMEM_BYTES + 64);
    std::vector<uint64_t> cand;
    for (;;) {
      ++passes;
      for (int g = 0; g < n_groups; ++g) base[(size_t)g + 1] = base[(size_t)g] + cap[(size_t)g];
      cand.assign((size_t)base[(size_t)n_groups] + 1, 0);
      std::fill(cnt.begin(), cnt.end(), 0ULL);
      n_valid = 0;
      if (!tiles.empty())
        for (int b = 0; b < 3; ++b)
          ani_hash_kernel(b, 3, blob.data(), tiles.data(), (int)tiles.size(), k, thr.data(), base.data(), cap.data(), cnt.data(), cand.data(), &n_valid, sm.data());
      bool again = false;
      for (int g = 0; g < n_groups; ++g) {
        if (cnt[(size_t)g] > cap[(size_t)g]) { cap[(size_t)g] = cnt[(size_t)g]; again = true; }
        else if (cnt[(size_t)g] < (unsigned long long)sketch_size && thr[(size_t)g] != ~0ULL) { thr[(size_t)g] = ~0ULL; cap[(size_t)g] = npos[(size_t)g]; again = true; }
      }
      if (again && passes < 4) continue;
      if (again) { wfb_set_last_error_("ANI candidate selection did not converge"); return WFB_ECAP; }
      for (int g = 0; g < n_groups; ++g) std::sort(cand.begin() + (ptrdiff_t)base[(size_t)g], cand.begin() + (ptrdiff_t)(base[(size_t)g] + cnt[(size_t)g]));
      for (int b = 0; b < 1; ++b) ani_gather_kernel(b, 1, cand.data(), base.data(), cnt.data(), n_groups, sketch_size, sketches);
      break;
    }
  }
#endif
  for (int g = 0; g < n_groups; ++g) sketch_count[g] = (int32_t)std::min<unsigned long long>(cnt[(size_t)g], (unsigned long long)sketch_size);
  if (stats) {
    stats->hash_kernel_ms = hash_ms; stats->sort_kernel_ms = sort_ms; stats->bases = (uint64_t)blob_bytes; stats->valid_kmers = n_valid;
    stats->tiles = (uint64_t)tiles.size(); stats->passes = passes;
    unsigned long long tc = 0;
    for (int g = 0; g < n_groups; ++g) tc += cnt[(size_t)g];
    stats->candidates = tc;
  }
  return WFB_OK;
}

static float ani_j2md(float j, int k) { /* Stat::j2md, map_stats.hpp:56-66 */
  if (j == 0) return 1.0f;
  if (j == 1) return 0.0f;
  const float mash_dist = 1 - std::pow(2 * j / (1 + j), 1.0 / k);
  return mash_dist;
}

extern "C" double wfb_ani_estimate_identity(const uint64_t* q_sketch, const int32_t* q_count, const int32_t* q_group, int32_t nq, const uint64_t* t_sketch,
                                            const int32_t* t_count, const int32_t* t_group, int32_t nt, int32_t sketch_size, int32_t kmer_size,
                                            int32_t ani_percentile, float ani_adjustment, int32_t* n_comparisons) {
  const double fallback = 0.70f; /* skch::fixed::percentage_identity (a float) */
  if (n_comparisons) *n_comparisons = 0;
  if (nq <= 0 || nt <= 0 || !q_sketch || !t_sketch || !q_count || !t_count || !q_group || !t_group) return fallback;
  /* the reference walks two std::map<int, sketch>: ascending group id on both sides (map_stats.hpp:700-745) */
  std::vector<int> qi((size_t)nq), ti((size_t)nt);
  for (int i = 0; i < nq; ++i) qi[(size_t)i] = i;
  for (int i = 0; i < nt; ++i) ti[(size_t)i] = i;
  std::sort(qi.begin(), qi.end(), [&](int a, int b) { return q_group[a] < q_group[b]; });
  std::sort(ti.begin(), ti.end(), [&](int a, int b) { return t_group[a] < t_group[b]; });
  std::vector<double> anis;
  for (int a : qi)
    for (int b : ti) {
      if (q_group[a] == t_group[b]) continue; /* a group is not compared with itself; `is_self_mode` compares the addresses of two
                                                 different members and is therefore always false: both (A,B) and (B,A) are kept */
      const uint64_t* qs = q_sketch + (size_t)a * sketch_size;
      const uint64_t* ts = t_sketch + (size_t)b * sketch_size;
      const size_t qn = (size_t)q_count[a], tn = (size_t)t_count[b];
      if (qn == 0 || tn == 0) continue;
      size_t inter = 0, i = 0, j = 0;
      while (i < qn && j < tn) {
        if (qs[i] == ts[j]) { inter++; i++; j++; }
        else if (qs[i] < ts[j]) i++;
        else j++;
      }
      if (inter == 0) continue;
      const double jaccard = static_cast<double>(inter) / std::min(qn, tn);
      const double mash_dist = ani_j2md((float)jaccard, kmer_size);
      anis.push_back(1.0 - mash_dist);
    }
  if (n_comparisons) *n_comparisons = (int32_t)anis.size();
  if (anis.empty()) return fallback;
  std::sort(anis.begin(), anis.end());
  size_t idx = ((size_t)ani_percentile * anis.size()) / 100;
  if (idx >= anis.size()) idx = anis.size() - 1;
  double adjusted = anis[idx] + (ani_adjustment / 100.0);
  if (adjusted < 0.0) adjusted = 0.0;
  if (adjusted > 1.0) adjusted = 1.0;
  return adjusted;
}
