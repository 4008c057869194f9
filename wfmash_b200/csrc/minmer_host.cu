// minmer_host.cu — host side of the reference-side minmer pipeline behind the C ABI:
//   wfb_minmers_build : addMinmers for a batch of target sequences (commonFunc.hpp:439-708 via
//                       Sketch::buildHelper, winSketch.hpp:467-499), output in Sketch::build's order.
// Kernels: minmer_kernels.h (hand-written). Sorting / scanning between them uses CUB (library code,
// plumbing between the kernels, not the hot loop).
#include "minmer_kernels.h"
#include "../../include/wfmash_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <string>
#include <vector>

#ifndef WFB_EMU
#include <cub/cub.cuh>
#endif
#ifndef WFB_MM_FSMEM_DEFAULT
#define WFB_MM_FSMEM_DEFAULT 0 /* shared-memory containers for the filtered run: measured 3.6 x SLOWER (40 threads per SM), kept as a switch */
#endif
#ifndef WFB_MM_LCUR_DEFAULT
#define WFB_MM_LCUR_DEFAULT 1 /* departures read from the candidate stream (measured: filtered stream 12.2 -> 11.5 ms on scerevisiae8) */
#endif
#ifndef WFB_MM_FILTER_DEFAULT
#define WFB_MM_FILTER_DEFAULT 1 /* candidate-filtered minmer build (minmer_kernels.h); WFB_MM_FILTER=0/1 overrides */
#endif
#include "wfb_pool.h" /* this file's cudaMalloc / cudaFree go through the library's device-memory pool */

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */
void wfb_count_launch_();

#ifndef WFB_EMU
#define MM_CHECK(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      wfb_set_last_error_(std::string(#call) + ": " + cudaGetErrorString(e_));           \
      rc = (e_ == cudaErrorMemoryAllocation) ? WFB_ENOMEM : WFB_ECUDA;                   \
      goto done;                                                                         \
    }                                                                                    \
  } while (0)
#define MM_LAUNCH(kernel, grid, block, ...)                                              \
  do {                                                                                   \
    kernel<<<(grid), (block)>>>(__VA_ARGS__);                                            \
    wfb_count_launch_();                                                                 \
  } while (0)
#endif

/* upper-case / N-mask the concatenated targets in place (makeUpperCaseAndValidDNA) */
WFB_KERNEL(mm_clean_kernel, uint8_t* buf, long long n) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) buf[i] = mm_clean_base(buf[i]);
}

struct MmFinal { /* post-pass record */
  uint64_t hash;
  long long wpos, wpos_end;
  int seq, strand;
};

/* post pass 1 (:660-693): number of output pieces of every raw record (0 = dropped) */
WFB_KERNEL(mm_pieces_kernel, const MmRecord* recs, long long n, int w, int* pieces, const int* chunk_flag, long long nrec_filtered) {
  WFB_KERNEL_PROLOGUE
  for (long long r = (long long)bid * WFB_NT + WFB_TID; r < n; r += (long long)nblocks * WFB_NT) {
    const MmRecord m = recs[r];
    int p = 1;
    if (r < nrec_filtered && chunk_flag[m.chunk]) p = 0; /* a filtered run that gave up: the exact re-run wrote this chunk's records again */
    else if (m.wpos < 0 || m.wpos_end < 0 || m.wpos == m.wpos_end) p = 0;
    else if (m.wpos_end > m.wpos + w) p = (int)ceilf((float)(m.wpos_end - m.wpos) / (float)w);
    pieces[r] = p;
  }
}
/* post pass 2: write the pieces (strand sign :672, chunking :673-685) */
WFB_KERNEL(mm_expand_kernel, const MmRecord* recs, long long n, int w, const int* pieces, const long long* offs, MmFinal* out) {
  WFB_KERNEL_PROLOGUE
  for (long long r = (long long)bid * WFB_NT + WFB_TID; r < n; r += (long long)nblocks * WFB_NT) {
    const int p = pieces[r];
    if (p == 0) continue;
    const MmRecord m = recs[r];
    MmFinal f;
    f.hash = m.hash; f.seq = m.seq; f.strand = m.strand < 0 ? -1 : 1;
    if (p == 1 && !(m.wpos_end > m.wpos + w)) {
      f.wpos = m.wpos; f.wpos_end = m.wpos_end;
      out[offs[r]] = f;
    } else {
      for (int c = 0; c < p; ++c) {
        f.wpos = m.wpos + (long long)c * w;
        const long long e = m.wpos + (long long)c * w + w;
        f.wpos_end = e < m.wpos_end ? e : m.wpos_end;
        out[offs[r] + c] = f;
      }
    }
  }
}
/* sort-key extraction for the three stable LSD passes: hash, wpos_end, (seq, wpos) */
WFB_KERNEL(mm_keys_kernel, const MmFinal* f, const int* perm, long long n, int which, unsigned long long* keys) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    const MmFinal& m = f[perm ? perm[i] : i];
    keys[i] = which == 0 ? m.hash : which == 1 ? (unsigned long long)m.wpos_end
                                               : (((unsigned long long)(unsigned)m.seq << 40) | (unsigned long long)m.wpos);
  }
}
WFB_KERNEL(mm_iota_kernel, int* p, long long n) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) p[i] = (int)i;
}
/* :698-706 std::unique on (wpos, hash): keep[i] = first of its run in sorted order (per sequence) */
WFB_KERNEL(mm_unique_flag_kernel, const MmFinal* f, const int* perm, long long n, int* keep) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    int kp = 1;
    if (i > 0) {
      const MmFinal& a = f[perm[i - 1]];
      const MmFinal& b = f[perm[i]];
      if (a.seq == b.seq && a.wpos == b.wpos && a.hash == b.hash) kp = 0;
    }
    keep[i] = kp;
  }
}
WFB_KERNEL(mm_gather_out_kernel, const MmFinal* f, const int* perm, const int* keep, const long long* offs, long long n,
           const MmSeq* seqs, wfb_minmer_t* out) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    if (!keep[i]) continue;
    const MmFinal& m = f[perm[i]];
    wfb_minmer_t o;
    o.hash = m.hash; o.wpos = m.wpos; o.wpos_end = m.wpos_end; o.seqId = seqs[m.seq].seq_id; o.strand = (int16_t)m.strand; o.pad_ = 0;
    out[offs[i]] = o;
  }
}

/* Neighbours of the final order that tie on (sequence, wpos, wpos_end): their order is what the reference's unstable std::sort leaves. */
WFB_KERNEL(mm_tie_count_kernel, const wfb_minmer_t* out, long long n, unsigned long long* ties) {
  WFB_KERNEL_PROLOGUE
  unsigned long long t = 0;
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i + 1 < n; i += (long long)nblocks * WFB_NT)
    t += out[i].seqId == out[i + 1].seqId && out[i].wpos == out[i + 1].wpos && out[i].wpos_end == out[i + 1].wpos_end;
  if (t) atomicAdd_compat(ties, t);
}

/* ---- the reference's order inside (wpos, wpos_end) ties (host; runs only for sequences that have such ties) ----
 * addMinmers ends with std::sort on (wpos, wpos_end) (commonFunc.hpp:696). Records that tie on both come out in the order GNU
 * libstdc++'s introsort leaves them, a function of the order in which the loop pushed them — and computeL2MappedRegions evaluates the
 * sketch after every single insertion (mappingCore.hpp:352-384), so that order shows in the mappings of targets barely longer than one
 * window (all minmers of a sequence of w .. 2w bases are opened by window 0 and closed together by the final flush). Real chromosomes
 * have no such ties (0 on scerevisiae8 / LPA). When a sequence has one, its records are put back in the reference's push order —
 * by closing position; at one position: the leaving k-mer's record, the arriving k-mer's, the evicted one (:517-617); the final flush in
 * hash order (:646-658) — and the reference's post passes (:660-706) are applied literally, with the same library's std::sort. */
static uint64_t mmh_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t mmh_fmix64(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }
static uint64_t mmh_murmur3_lo64(const uint8_t* p, int len) { /* MurmurHash3_x64_128, seed 42, low word (murmur3.h:226-303), len <= 32 */
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = 42, h2 = 42;
  const int nblocks = len / 16;
  for (int i = 0; i < nblocks; ++i) {
    uint64_t k1, k2;
    memcpy(&k1, p + 16 * i, 8); memcpy(&k2, p + 16 * i + 8, 8);
    k1 *= c1; k1 = mmh_rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = mmh_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = mmh_rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = mmh_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const uint8_t* tail = p + 16 * nblocks;
  uint64_t k1 = 0, k2 = 0;
  const int rem = len & 15;
  for (int j = rem - 1; j >= 8; --j) k2 = (k2 << 8) | tail[j];
  for (int j = (rem < 8 ? rem : 8) - 1; j >= 0; --j) k1 = (k1 << 8) | tail[j];
  if (rem > 8) { k2 *= c2; k2 = mmh_rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  if (rem > 0) { k1 *= c1; k1 = mmh_rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = mmh_fmix64(h1); h2 = mmh_fmix64(h2);
  h1 += h2;
  return h1;
}
static uint64_t mmh_canonical(const uint8_t* s, int k) { /* s: cleaned bases */
  uint8_t rc[32];
  for (int j = 0; j < k; ++j) { const uint8_t c = s[k - 1 - j]; rc[j] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }
  const uint64_t f = mmh_murmur3_lo64(s, k), b = mmh_murmur3_lo64(rc, k);
  return f < b ? f : b;
}
struct MmhRaw { wfb_minmer_t m; int stage; };
/* cleaned: the cleaned sequence buffer (MmSeq::off offsets); rec: every raw record after the stitch; out: the final array.
 * Returns the rebuilt array in `fixed` (== out when nothing had to change) and the number of sequences rebuilt. */
static long long mm_fix_tie_order_host(const uint8_t* cleaned, const std::vector<MmSeq>& seqs, int k, int w, const MmRecord* rec, long long nrec,
                                       const int* chunk_flag, long long nrec_filtered, const wfb_minmer_t* out, long long nout,
                                       std::vector<wfb_minmer_t>& fixed) {
  const int ns = (int)seqs.size();
  { /* the output is segmented by seqId: ids given twice make the segments ambiguous, leave the order as it is */
    std::vector<int> idsort((size_t)ns);
    for (int q = 0; q < ns; ++q) idsort[q] = seqs[q].seq_id;
    std::sort(idsort.begin(), idsort.end());
    if (std::adjacent_find(idsort.begin(), idsort.end()) != idsort.end()) { fixed.assign(out, out + nout); return 0; }
  }
  std::vector<char> has_tie((size_t)ns, 0);
  std::vector<long long> seg_begin((size_t)ns + 1, 0);
  { /* the output is ordered by sequence INDEX (input order); map seqId -> index through the segment walk */
    long long i = 0;
    for (int q = 0; q < ns; ++q) {
      seg_begin[q] = i;
      while (i < nout && out[i].seqId == seqs[q].seq_id) {
        if (i + 1 < nout && out[i + 1].seqId == out[i].seqId && out[i + 1].wpos == out[i].wpos && out[i + 1].wpos_end == out[i].wpos_end) has_tie[q] = 1;
        ++i;
      }
    }
    seg_begin[ns] = i;
    if (i != nout) { fixed.assign(out, out + nout); return 0; } /* cannot happen (the output is in input order); never lose records over it */
  }
  std::vector<std::vector<MmhRaw>> raws((size_t)ns);
  for (long long r = 0; r < nrec; ++r) {
    const MmRecord& x = rec[r];
    if (!has_tie[x.seq]) continue;
    if (r < nrec_filtered && chunk_flag && chunk_flag[x.chunk]) continue;
    MmhRaw e;
    e.m.hash = x.hash; e.m.wpos = x.wpos; e.m.wpos_end = x.wpos_end; e.m.seqId = seqs[x.seq].seq_id; e.m.strand = (int16_t)(x.strand < 0 ? -1 : 1); e.m.pad_ = 0;
    e.stage = 2;
    raws[x.seq].push_back(e);
  }
  long long rebuilt = 0;
  fixed.clear();
  fixed.reserve((size_t)nout);
  for (int q = 0; q < ns; ++q) {
    if (!has_tie[q]) { fixed.insert(fixed.end(), out + seg_begin[q], out + seg_begin[q + 1]); continue; }
    std::vector<MmhRaw>& v = raws[q];
    const uint8_t* sq = cleaned + seqs[q].off;
    const long long len = seqs[q].len, flush_end = len - k + 1;
    for (MmhRaw& e : v) {
      if (e.m.wpos_end == flush_end) { e.stage = 3; continue; } /* final flush */
      const long long win = e.m.wpos_end, i = win - k + w;       /* closed at loop position i */
      if (win - 1 >= 0 && mmh_canonical(sq + (win - 1), k) == e.m.hash) e.stage = 0;       /* the k-mer that left the window */
      else if (i >= 0 && i <= len - k && mmh_canonical(sq + i, k) == e.m.hash) e.stage = 1; /* the k-mer that entered it */
    }
    std::sort(v.begin(), v.end(), [](const MmhRaw& a, const MmhRaw& b) { /* a total order: the push order of the reference's loop */
      const bool fa = a.stage == 3, fb = b.stage == 3;
      if (fa != fb) return fb;
      if (fa) return a.m.hash < b.m.hash;
      if (a.m.wpos_end != b.m.wpos_end) return a.m.wpos_end < b.m.wpos_end;
      if (a.stage != b.stage) return a.stage < b.stage;
      if (a.m.wpos != b.m.wpos) return a.m.wpos > b.m.wpos; /* same entry emitted twice at one position: the second one is degenerate */
      return a.m.hash < b.m.hash;
    });
    std::vector<wfb_minmer_t> mi, pieces;
    for (const MmhRaw& e : v) { /* :660-694 */
      const wfb_minmer_t& m = e.m;
      if (m.wpos < 0 || m.wpos_end < 0 || m.wpos == m.wpos_end) continue;
      if (m.wpos_end > m.wpos + w) {
        const int nch = (int)ceilf((float)(m.wpos_end - m.wpos) / (float)w);
        for (int c = 0; c < nch; ++c) {
          wfb_minmer_t p = m;
          p.wpos = m.wpos + (int64_t)c * w;
          p.wpos_end = std::min<int64_t>(m.wpos + (int64_t)c * w + w, m.wpos_end);
          pieces.push_back(p);
        }
      } else mi.push_back(m);
    }
    mi.insert(mi.end(), pieces.begin(), pieces.end());
    std::sort(mi.begin(), mi.end(), [](const wfb_minmer_t& l, const wfb_minmer_t& r) { return l.wpos < r.wpos || (l.wpos == r.wpos && l.wpos_end < r.wpos_end); }); /* :696 */
    mi.erase(std::unique(mi.begin(), mi.end(), [](const wfb_minmer_t& l, const wfb_minmer_t& r) { return l.wpos == r.wpos && l.hash == r.hash; }), mi.end()); /* :701-706 */
    fixed.insert(fixed.end(), mi.begin(), mi.end());
    ++rebuilt;
  }
  return rebuilt;
}

/* Shared by wfb_minmers_build (results copied to the caller's host buffer) and wfb_index_build (d_out_keep != NULL:
 * the result stays in device memory, ownership of *d_out_keep passes to the caller, `out` is not touched). */
int wfb_minmers_build_impl(int device, const char* const* seq_ptrs, const int64_t* seq_lens, const int32_t* seq_ids, int32_t nseq,
                           int32_t kmer_size, int32_t window_size, int32_t sketch_size, wfb_minmer_t* out, int64_t out_cap,
                           int64_t* out_count, wfb_minmer_stats_t* stats, wfb_minmer_t** d_out_keep) {
  if (d_out_keep) *d_out_keep = nullptr;
  if (nseq < 0 || kmer_size <= 0 || kmer_size > 32 || window_size <= kmer_size || sketch_size <= 0 || !out_count ||
      (nseq > 0 && (!seq_ptrs || !seq_lens || !seq_ids))) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  *out_count = 0;
  if (stats) memset(stats, 0, sizeof(*stats));
  int rc = WFB_OK;
  const int k = kmer_size, w = window_size, s = sketch_size;
  /* sequences shorter than w are skipped by Sketch::build (winSketch.hpp:218-232) */
  std::vector<MmSeq> seqs;
  std::vector<int> src_index;
  long long total = 16;
  for (int i = 0; i < nseq; ++i) {
    if (seq_lens[i] < w) continue;
    if (seq_lens[i] > (int64_t)INT_MAX) { /* the stream keeps k-mer positions as 32-bit integers (the reference's offset_t is 64-bit) */
      wfb_set_last_error_("a target sequence longer than 2^31 - 1 bases is not supported by the minmer build");
      return WFB_EINVAL;
    }
    MmSeq q;
    q.off = total; q.len = seq_lens[i]; q.seq_id = seq_ids[i]; q.first_chunk = 0; q.n_chunks = 0;
    total += (seq_lens[i] + 63) / 64 * 64 + 64;
    seqs.push_back(q);
    src_index.push_back(i);
  }
  const int ns = (int)seqs.size();
  if (ns == 0) return WFB_OK;
  MmParams P;
  P.k = k; P.w = w; P.s = s;
  /* Filtered build (minmer_kernels.h): only k-mers whose hash is <= T enter the window machine, T such that a window holds
   * lambda = 2s + 40 of them on average (s = 24: the chance that a window holds fewer than s is ~1e-16; a chunk that meets one re-runs
   * exactly). Off when the candidates would be every other k-mer anyway (tiny windows), or with WFB_MM_FILTER=0. */
  const double lambda = 2.0 * s + 40.0;
  const double density = lambda / (double)(w - k + 1);
  int use_filter = WFB_MM_FILTER_DEFAULT;
  { const char* e = getenv("WFB_MM_FILTER"); if (e && *e) use_filter = atoi(e) != 0; }
  if (density > 0.45) use_filter = 0;
  /* the canonical hash is the smaller of two uniform hashes: P(min <= t) = 1 - (1 - t)^2 */
  const uint64_t T = use_filter ? (uint64_t)ldexp(1.0 - sqrt(1.0 - density), 64) : ~0ULL;
  int cand_cap = std::min<int>(MMC_TILE, (int)(density * MMC_TILE * 1.25 + 8.0 * sqrt(density * MMC_TILE)) + 64);
  { const char* e = getenv("WFB_MM_CAND_CAP"); if (e && atoi(e) > 0) cand_cap = std::min<int>(MMC_TILE, atoi(e)); } /* test hook: forces the tile-overflow fallback */
  P.chunk = 1024; P.warm = w; /* full run tuned in profiles/r01_minmer_chunk_sweep.txt: latency-bound, more chunks = more threads */
  { const char* e = getenv(use_filter ? "WFB_MM_FCHUNK" : "WFB_MM_CHUNK"); if (e && atoi(e) > 0) P.chunk = atoi(e); }
  { const char* e = getenv("WFB_MM_WARM"); if (e && atoi(e) >= w) P.warm = atoi(e); }
  P.qcap = w + 2; P.heap_cap = 3 * w + 64; P.pool_cap = 4 * w + 64;
  P.lcur = 0;
  MmParams PF = P; /* the filtered run's containers only ever hold candidates */
  PF.qcap = PF.heap_cap = PF.pool_cap = (int)(3.0 * lambda) + 64;
  PF.lcur = WFB_MM_LCUR_DEFAULT;
  { const char* e = getenv("WFB_MM_LCUR"); if (e && *e) PF.lcur = atoi(e) != 0; }
  if (PF.lcur) PF.qcap = 1; /* the window queue of a filtered run is the candidate list itself */
  /* shared-memory variant of the filtered run (needs lcur): tight containers, a chunk that outgrows them flags itself like any other */
  int fsmem = WFB_MM_FSMEM_DEFAULT;
  { const char* e = getenv("WFB_MM_FSMEM"); if (e && *e) fsmem = atoi(e) != 0; }
  if (!PF.lcur) fsmem = 0;
  if (fsmem) { PF.heap_cap = std::max(64, (int)(1.5 * lambda) + 32); PF.pool_cap = std::max(64, (int)lambda + 32); }
  std::vector<MmChunk> chunks;
  for (int q = 0; q < ns; ++q) {
    const long long npos = seqs[q].len - k + 1;
    seqs[q].first_chunk = (int)chunks.size();
    for (long long cb = 0; cb < npos; cb += P.chunk) {
      MmChunk c;
      c.seq = q; c.body_begin = cb; c.body_end = std::min<long long>(cb + P.chunk, npos);
      c.run_begin = std::max<long long>(0, cb - P.warm);
      c.first_of_seq = cb == 0; c.last_of_seq = c.body_end == npos;
      chunks.push_back(c);
    }
    seqs[q].n_chunks = (int)chunks.size() - seqs[q].first_chunk;
  }
  const int nchunks = (int)chunks.size();
  const long long scratch_stride = ((long long)sizeof(MmKmer) * (P.qcap + P.heap_cap) + (long long)sizeof(MmNode) * P.pool_cap +
                                    (long long)sizeof(MmWent) * (s + 2) + 255) / 256 * 256;
  long long scratch_stride_f = ((long long)sizeof(MmKmer) * (PF.qcap + PF.heap_cap) + (long long)sizeof(MmNode) * PF.pool_cap +
                                (long long)sizeof(MmWent) * (s + 2) + 255) / 256 * 256;
  int fsmem_threads = 0;
  if (fsmem) { /* + 16: consecutive threads' slices start in different banks */
    scratch_stride_f = ((long long)sizeof(MmKmer) * (PF.qcap + PF.heap_cap) + (long long)sizeof(MmNode) * PF.pool_cap +
                        (long long)sizeof(MmWent) * (s + 2) + 127) / 128 * 128 + 16;
    fsmem_threads = (int)std::min<long long>(64, 220 * 1024 / scratch_stride_f) / 8 * 8;
    if (fsmem_threads < 16) fsmem = 0;
  }
  std::vector<MmTile> tiles;
  std::vector<int> seq_tile0((size_t)ns, 0);
  if (use_filter)
    for (int q = 0; q < ns; ++q) {
      const long long npos = seqs[q].len - k + 1;
      seq_tile0[q] = (int)tiles.size();
      for (long long tb = 0; tb < npos; tb += MMC_TILE) {
        MmTile t;
        t.seq = q; t.start = tb; t.npos = (int)std::min<long long>(MMC_TILE, npos - tb);
        tiles.push_back(t);
      }
    }
  const int ntiles = (int)tiles.size();
  long long nrec_filtered = 0, n_redo = 0; /* records written by the filtered run; chunks re-run exactly */
  long long tie_sequences = 0;             /* sequences whose tied records were put in the reference's order on the host */
  /* expected density ~0.0027*s windows per base; generous cap, overflow is detected */
  long long rec_cap = (long long)((double)total * (0.01 * s + 0.05)) + 65536;
#ifndef WFB_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    wfb_set_last_error_("no CUDA device (this library has no CPU path)");
    return WFB_ENODEV;
  }
  uint8_t* d_seq = nullptr; MmSeq* d_seqs = nullptr; MmChunk* d_chunks = nullptr; unsigned char* d_scratch = nullptr;
  MmRecord* d_rec = nullptr; MmEndEnt* d_end = nullptr; int* d_endcount = nullptr; MmCounters* d_cnt = nullptr;
  int* d_pieces = nullptr; long long* d_offs = nullptr; MmFinal* d_fin = nullptr; unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
  int *d_perm = nullptr, *d_perm2 = nullptr, *d_keep = nullptr; wfb_minmer_t* d_out = nullptr; void* d_tmp = nullptr;
  MmTile* d_tiles = nullptr; int *d_tile0 = nullptr, *d_cand_lp = nullptr, *d_cand_cnt = nullptr, *d_flag = nullptr, *d_redo = nullptr;
  uint64_t* d_cand_hash = nullptr;
  MmCandView CV;
  size_t tmp_bytes = 0;
  uint8_t* h_seq = nullptr;
  MmCounters hc;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, ec = nullptr, ef = nullptr; /* ec: candidates done, ef: filtered stream done */
  long long nrec = 0, nfin = 0, nout = 0;
  const int TPB = 64;
  std::vector<long long> tmpll(2);
  MM_CHECK(cudaSetDevice(device));
  MM_CHECK(cudaMallocHost(&h_seq, (size_t)total));
  memset(h_seq, 'N', (size_t)total);
  for (int q = 0; q < ns; ++q) memcpy(h_seq + seqs[q].off, seq_ptrs[src_index[q]], (size_t)seqs[q].len);
  MM_CHECK(cudaMalloc(&d_seq, (size_t)total));
  MM_CHECK(cudaMalloc(&d_seqs, sizeof(MmSeq) * ns));
  MM_CHECK(cudaMalloc(&d_chunks, sizeof(MmChunk) * (size_t)nchunks));
  MM_CHECK(cudaMalloc(&d_scratch, (size_t)(use_filter ? scratch_stride_f : scratch_stride) *
                                       (((size_t)nchunks + MM_LANES - 1) / MM_LANES * MM_LANES))); /* whole warps: the slabs are interleaved */
  if (use_filter) {
    MM_CHECK(cudaMalloc(&d_tiles, sizeof(MmTile) * (size_t)ntiles));
    MM_CHECK(cudaMalloc(&d_tile0, sizeof(int) * (size_t)ns));
    MM_CHECK(cudaMalloc(&d_cand_hash, sizeof(uint64_t) * (size_t)ntiles * cand_cap));
    MM_CHECK(cudaMalloc(&d_cand_lp, sizeof(int) * (size_t)ntiles * cand_cap));
    MM_CHECK(cudaMalloc(&d_cand_cnt, sizeof(int) * (size_t)ntiles));
    MM_CHECK(cudaMalloc(&d_flag, sizeof(int) * (size_t)nchunks));
    MM_CHECK(cudaMalloc(&d_redo, sizeof(int) * (size_t)nchunks));
    MM_CHECK(cudaMemset(d_flag, 0, sizeof(int) * (size_t)nchunks));
    MM_CHECK(cudaMemcpy(d_tiles, tiles.data(), sizeof(MmTile) * (size_t)ntiles, cudaMemcpyHostToDevice));
    MM_CHECK(cudaMemcpy(d_tile0, seq_tile0.data(), sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice));
  }
  MM_CHECK(cudaMalloc(&d_rec, sizeof(MmRecord) * (size_t)rec_cap));
  MM_CHECK(cudaMalloc(&d_end, sizeof(MmEndEnt) * (size_t)nchunks * s));
  MM_CHECK(cudaMalloc(&d_endcount, sizeof(int) * (size_t)nchunks));
  MM_CHECK(cudaMalloc(&d_cnt, sizeof(MmCounters)));
  MM_CHECK(cudaMemset(d_cnt, 0, sizeof(MmCounters)));
  MM_CHECK(cudaEventCreate(&e0)); MM_CHECK(cudaEventCreate(&e1)); MM_CHECK(cudaEventCreate(&e2));
  MM_CHECK(cudaEventCreate(&ec)); MM_CHECK(cudaEventCreate(&ef));
  MM_CHECK(cudaMemcpy(d_seq, h_seq, (size_t)total, cudaMemcpyHostToDevice));
  MM_CHECK(cudaMemcpy(d_seqs, seqs.data(), sizeof(MmSeq) * ns, cudaMemcpyHostToDevice));
  MM_CHECK(cudaMemcpy(d_chunks, chunks.data(), sizeof(MmChunk) * (size_t)nchunks, cudaMemcpyHostToDevice));
  MM_CHECK(cudaEventRecord(e0));
  MM_LAUNCH(mm_clean_kernel, 148 * 8, 256, d_seq, total);
  if (use_filter) {
    CV.hash = d_cand_hash; CV.lp = d_cand_lp; CV.cnt = d_cand_cnt; CV.cap = cand_cap;
    MM_LAUNCH(mm_cand_kernel, std::min(ntiles, 148 * MMC_MINBLOCKS), MMC_THREADS, d_seq, d_seqs, d_tiles, ntiles, k, T, cand_cap, d_cand_hash, d_cand_lp,
              d_cand_cnt, &d_cnt->candidates);
    MM_CHECK(cudaEventRecord(ec));
    if (fsmem) {
      MM_CHECK(cudaFuncSetAttribute(mm_stream_cand_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fsmem_threads * scratch_stride_f)));
      mm_stream_cand_smem_kernel<<<(nchunks + fsmem_threads - 1) / fsmem_threads, fsmem_threads, (size_t)(fsmem_threads * scratch_stride_f)>>>(
          d_seq, d_seqs, d_chunks, nchunks, PF, scratch_stride_f, d_rec, rec_cap, d_end, d_endcount, d_cnt, CV, d_tile0, d_flag, d_redo);
      wfb_count_launch_();
    } else
    MM_LAUNCH(mm_stream_cand_kernel, (nchunks + TPB - 1) / TPB, TPB, d_seq, d_seqs, d_chunks, nchunks, PF, d_scratch, scratch_stride_f, d_rec,
              rec_cap, d_end, d_endcount, d_cnt, CV, d_tile0, d_flag, d_redo);
    MM_CHECK(cudaMemcpy(&hc, d_cnt, sizeof(hc), cudaMemcpyDeviceToHost));
    nrec_filtered = (long long)std::min<unsigned long long>(hc.n_records, (unsigned long long)rec_cap);
    n_redo = (long long)hc.flagged;
    MM_CHECK(cudaEventRecord(ef));
    if (n_redo > 0) { /* exact re-run of the chunks whose filtered run gave up (short windows around N runs, capacity) */
      int redo_smem = 1; /* few chunks: one CTA each, containers in shared memory; many (every tile overflowed): the global slabs */
      { const char* e = getenv("WFB_MM_REDO_SMEM"); if (e && *e) redo_smem = atoi(e) != 0; }
      /* one CTA per SM at 115 KB each: a wave of 148 chunks takes ~2.9 ms, the global-slab kernel ~35 ms whatever the count (measured, profiles/r02_minmer_filtered_run2.json) */
      if (redo_smem && scratch_stride <= 200 * 1024 && n_redo <= 148 * 8) {
        MM_CHECK(cudaFuncSetAttribute(mm_stream_redo_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch_stride));
        mm_stream_redo_smem_kernel<<<(int)n_redo, 32, (size_t)scratch_stride>>>(d_seq, d_seqs, d_chunks, (int)n_redo, P, scratch_stride, d_rec, rec_cap,
                                                                                 d_end, d_endcount, d_cnt, d_redo);
        wfb_count_launch_();
      } else {
        cudaFree(d_scratch); d_scratch = nullptr;
        MM_CHECK(cudaMalloc(&d_scratch, (size_t)scratch_stride * (((size_t)n_redo + MM_LANES - 1) / MM_LANES * MM_LANES)));
        MM_LAUNCH(mm_stream_kernel, (int)((n_redo + TPB - 1) / TPB), TPB, d_seq, d_seqs, d_chunks, (int)n_redo, P, d_scratch, scratch_stride, d_rec,
                  rec_cap, d_end, d_endcount, d_cnt, d_redo);
      }
    }
  } else {
    MM_LAUNCH(mm_stream_kernel, (nchunks + TPB - 1) / TPB, TPB, d_seq, d_seqs, d_chunks, nchunks, P, d_scratch, scratch_stride, d_rec,
              rec_cap, d_end, d_endcount, d_cnt, (const int*)nullptr);
  }
  MM_CHECK(cudaEventRecord(e1));
  { /* stitch 1: parallel passes over all (chunk, entry) pairs until every inherited start is settled */
    int* d_pending = nullptr;
    MM_CHECK(cudaMalloc(&d_pending, sizeof(int)));
    int pending = 1, passes = 0;
    while (pending > 0 && passes <= nchunks) {
      MM_CHECK(cudaMemset(d_pending, 0, sizeof(int)));
      MM_LAUNCH(mm_stitch_ends_pass_kernel, 148 * 8, 256, d_chunks, nchunks, s, d_end, d_endcount, d_cnt, d_pending);
      MM_CHECK(cudaMemcpy(&pending, d_pending, sizeof(int), cudaMemcpyDeviceToHost));
      ++passes;
    }
    cudaFree(d_pending);
    if (pending > 0) { wfb_set_last_error_("minmer stitch did not converge"); rc = WFB_ECUDA; goto done; }
  }
  MM_CHECK(cudaMemcpy(&hc, d_cnt, sizeof(hc), cudaMemcpyDeviceToHost));
  if (hc.overflow || (long long)hc.n_records > rec_cap) {
    wfb_set_last_error_("minmer stream: capacity overflow (records / heap / pool)");
    rc = WFB_ECAP;
    goto done;
  }
  nrec = (long long)hc.n_records;
  MM_LAUNCH(mm_stitch_records_kernel, 148 * 8, 256, d_rec, nrec, d_chunks, s, d_end, d_endcount, d_cnt, d_flag, nrec_filtered);
  /* post passes */
  MM_CHECK(cudaMalloc(&d_pieces, sizeof(int) * (size_t)(nrec + 1)));
  MM_CHECK(cudaMalloc(&d_offs, sizeof(long long) * (size_t)(nrec + 1)));
  MM_LAUNCH(mm_pieces_kernel, 148 * 8, 256, d_rec, nrec, w, d_pieces, d_flag, nrec_filtered);
  MM_CHECK(cudaMemset(d_pieces + nrec, 0, sizeof(int)));
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_pieces, d_offs, nrec + 1);
  MM_CHECK(cudaMalloc(&d_tmp, tmp_bytes + (size_t)nrec * 32 + (1 << 20)));
  tmp_bytes += (size_t)nrec * 32 + (1 << 20);
  {
    size_t tb = tmp_bytes;
    MM_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_pieces, d_offs, nrec + 1));
  }
  MM_CHECK(cudaMemcpy(&nfin, d_offs + nrec, sizeof(long long), cudaMemcpyDeviceToHost));
  if (nfin > 0) {
    MM_CHECK(cudaMalloc(&d_fin, sizeof(MmFinal) * (size_t)nfin));
    MM_CHECK(cudaMalloc(&d_keys, 8 * (size_t)nfin)); MM_CHECK(cudaMalloc(&d_keys2, 8 * (size_t)nfin));
    MM_CHECK(cudaMalloc(&d_perm, 4 * (size_t)nfin)); MM_CHECK(cudaMalloc(&d_perm2, 4 * (size_t)nfin));
    MM_CHECK(cudaMalloc(&d_keep, 4 * (size_t)(nfin + 1)));
    MM_LAUNCH(mm_expand_kernel, 148 * 8, 256, d_rec, nrec, w, d_pieces, d_offs, d_fin);
    MM_LAUNCH(mm_iota_kernel, 148 * 8, 256, d_perm, nfin);
    for (int pass = 0; pass < 3; ++pass) { /* stable LSD: hash, wpos_end, (seq,wpos) => order (seq,wpos,wpos_end,hash) */
      MM_LAUNCH(mm_keys_kernel, 148 * 8, 256, d_fin, d_perm, nfin, pass, d_keys);
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb, d_keys, d_keys2, d_perm, d_perm2, (int)nfin);
      if (tb > tmp_bytes) { cudaFree(d_tmp); d_tmp = nullptr; MM_CHECK(cudaMalloc(&d_tmp, tb)); tmp_bytes = tb; }
      tb = tmp_bytes;
      MM_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys, d_keys2, d_perm, d_perm2, (int)nfin));
      std::swap(d_perm, d_perm2);
    }
    MM_LAUNCH(mm_unique_flag_kernel, 148 * 8, 256, d_fin, d_perm, nfin, d_keep);
    MM_CHECK(cudaMemset(d_keep + nfin, 0, 4));
    cudaFree(d_offs); d_offs = nullptr;
    MM_CHECK(cudaMalloc(&d_offs, sizeof(long long) * (size_t)(nfin + 1)));
    {
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, d_keep, d_offs, nfin + 1);
      if (tb > tmp_bytes) { cudaFree(d_tmp); d_tmp = nullptr; MM_CHECK(cudaMalloc(&d_tmp, tb)); tmp_bytes = tb; }
      tb = tmp_bytes;
      MM_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_keep, d_offs, nfin + 1));
    }
    MM_CHECK(cudaMemcpy(&nout, d_offs + nfin, sizeof(long long), cudaMemcpyDeviceToHost));
    MM_CHECK(cudaMalloc(&d_out, sizeof(wfb_minmer_t) * (size_t)std::max<long long>(nout, 1)));
    MM_LAUNCH(mm_gather_out_kernel, 148 * 8, 256, d_fin, d_perm, d_keep, d_offs, nfin, d_seqs, d_out);
  }
  if (nout > 1) MM_LAUNCH(mm_tie_count_kernel, 148 * 8, 256, d_out, nout, &d_cnt->ties);
  MM_CHECK(cudaEventRecord(e2));
  MM_CHECK(cudaEventSynchronize(e2));
  MM_CHECK(cudaMemcpy(&hc, d_cnt, sizeof(hc), cudaMemcpyDeviceToHost));
  if (hc.ties > 0) { /* rare (tiny targets, degenerate sequence): put the tied records in the reference's own order on the host */
    std::vector<wfb_minmer_t> h_out((size_t)nout), fixed;
    std::vector<MmRecord> h_rec((size_t)nrec);
    std::vector<int> h_flag;
    std::vector<uint8_t> cleaned((size_t)total);
    MM_CHECK(cudaMemcpy(h_out.data(), d_out, sizeof(wfb_minmer_t) * (size_t)nout, cudaMemcpyDeviceToHost));
    MM_CHECK(cudaMemcpy(h_rec.data(), d_rec, sizeof(MmRecord) * (size_t)nrec, cudaMemcpyDeviceToHost));
    MM_CHECK(cudaMemcpy(cleaned.data(), d_seq, (size_t)total, cudaMemcpyDeviceToHost));
    if (use_filter) {
      h_flag.resize((size_t)nchunks);
      MM_CHECK(cudaMemcpy(h_flag.data(), d_flag, sizeof(int) * (size_t)nchunks, cudaMemcpyDeviceToHost));
    }
    tie_sequences = mm_fix_tie_order_host(cleaned.data(), seqs, k, w, h_rec.data(), nrec, use_filter ? h_flag.data() : nullptr, nrec_filtered, h_out.data(), nout, fixed);
    if ((long long)fixed.size() != nout) {
      cudaFree(d_out); d_out = nullptr;
      nout = (long long)fixed.size();
      MM_CHECK(cudaMalloc(&d_out, sizeof(wfb_minmer_t) * (size_t)std::max<long long>(nout, 1)));
    }
    if (nout > 0) MM_CHECK(cudaMemcpy(d_out, fixed.data(), sizeof(wfb_minmer_t) * (size_t)nout, cudaMemcpyHostToDevice));
  }
  *out_count = nout;
  if (stats) {
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, e0, e1);
    cudaEventElapsedTime(&b, e0, e2);
    stats->stream_kernel_ms = a; stats->total_kernel_ms = b;
    if (use_filter) {
      float c1 = 0, c2 = 0, c3 = 0;
      cudaEventElapsedTime(&c1, e0, ec); cudaEventElapsedTime(&c2, ec, ef); cudaEventElapsedTime(&c3, ef, e1);
      stats->cand_kernel_ms = c1; stats->filtered_stream_ms = c2; stats->redo_ms = c3;
    }
    stats->raw_records = (uint64_t)nrec; stats->chunks = (uint64_t)nchunks; stats->stale_absorbed = hc.stale_absorbed;
    stats->stitch_miss = hc.stitch_miss; stats->bases = 0;
    stats->candidates = hc.candidates; stats->redo_chunks = (uint64_t)n_redo; stats->filtered = (uint64_t)use_filter;
    stats->tie_sequences = (uint64_t)tie_sequences;
    for (int q = 0; q < ns; ++q) stats->bases += (uint64_t)seqs[q].len;
  }
  if (d_out_keep) { *d_out_keep = d_out; d_out = nullptr; goto done; }
  if (nout > out_cap) { wfb_set_last_error_("minmer output buffer too small"); rc = WFB_ECAP; goto done; }
  if (nout > 0) MM_CHECK(cudaMemcpy(out, d_out, sizeof(wfb_minmer_t) * (size_t)nout, cudaMemcpyDeviceToHost));
done:
  if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (e2) cudaEventDestroy(e2); if (ec) cudaEventDestroy(ec); if (ef) cudaEventDestroy(ef);
  cudaFree(d_seq); cudaFree(d_seqs); cudaFree(d_chunks); cudaFree(d_scratch); cudaFree(d_rec); cudaFree(d_end); cudaFree(d_endcount);
  cudaFree(d_cnt); cudaFree(d_pieces); cudaFree(d_offs); cudaFree(d_fin); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_perm);
  cudaFree(d_perm2); cudaFree(d_keep); cudaFree(d_out); cudaFree(d_tmp);
  cudaFree(d_tiles); cudaFree(d_tile0); cudaFree(d_cand_hash); cudaFree(d_cand_lp); cudaFree(d_cand_cnt); cudaFree(d_flag); cudaFree(d_redo);
  if (h_seq) cudaFreeHost(h_seq);
  return rc;
#else
  /* host emulation of the same kernels (tests/emu only): one thread per "CTA", serial sorts */
  std::vector<uint8_t> buf((size_t)total, 'N');
  for (int q = 0; q < ns; ++q) memcpy(buf.data() + seqs[q].off, seq_ptrs[src_index[q]], (size_t)seqs[q].len);
  mm_clean_kernel(0, 1, buf.data(), total);
  std::vector<unsigned char> scratch((size_t)scratch_stride);
  std::vector<MmRecord> rec((size_t)rec_cap);
  std::vector<MmEndEnt> endst((size_t)nchunks * s);
  std::vector<int> endcount((size_t)nchunks);
  std::vector<int> flag((size_t)nchunks, 0), redo((size_t)nchunks, 0);
  MmCounters hc;
  memset(&hc, 0, sizeof(hc));
  if (use_filter) {
    std::vector<uint64_t> cand_hash((size_t)ntiles * cand_cap);
    std::vector<int> cand_lp((size_t)ntiles * cand_cap), cand_cnt((size_t)ntiles);
    std::vector<unsigned char> smem((size_t)MMC_SEQ_BYTES + 16 + (size_t)MMC_TILE * 12 + MMC_THREADS * 4 + 64);
    for (int t = 0; t < ntiles; ++t)
      mm_cand_kernel(t, ntiles, buf.data(), seqs.data(), tiles.data(), ntiles, k, T, cand_cap, cand_hash.data(), cand_lp.data(), cand_cnt.data(),
                     &hc.candidates, smem.data());
    MmCandView CV;
    CV.hash = cand_hash.data(); CV.lp = cand_lp.data(); CV.cnt = cand_cnt.data(); CV.cap = cand_cap;
    for (int c = 0; c < nchunks; ++c)
      mm_stream_cand_kernel(c, nchunks, buf.data(), seqs.data(), chunks.data(), nchunks, PF, scratch.data() - (long long)c * scratch_stride_f,
                            scratch_stride_f, rec.data(), rec_cap, endst.data(), endcount.data(), &hc, CV, seq_tile0.data(), flag.data(), redo.data());
    nrec_filtered = (long long)std::min<unsigned long long>(hc.n_records, (unsigned long long)rec_cap);
    n_redo = (long long)hc.flagged;
    for (long long c = 0; c < n_redo; ++c) {
      mm_stream_kernel((int)c, (int)n_redo, buf.data(), seqs.data(), chunks.data(), (int)n_redo, P, scratch.data() - c * scratch_stride, scratch_stride,
                       rec.data(), rec_cap, endst.data(), endcount.data(), &hc, redo.data());
    }
  } else {
    for (int c = 0; c < nchunks; ++c) /* scratch reused: chunk c uses slot 0 */
      mm_stream_kernel(c, nchunks, buf.data(), seqs.data(), chunks.data(), nchunks, P, scratch.data() - (long long)c * scratch_stride,
                       scratch_stride, rec.data(), rec_cap, endst.data(), endcount.data(), &hc, (const int*)nullptr);
  }
  for (int pending = 1, passes = 0; pending > 0 && passes <= nchunks; ++passes) {
    pending = 0;
    mm_stitch_ends_pass_kernel(0, 1, chunks.data(), nchunks, s, endst.data(), endcount.data(), &hc, &pending);
  }
  if (hc.overflow || (long long)hc.n_records > rec_cap) { wfb_set_last_error_("minmer stream: capacity overflow"); return WFB_ECAP; }
  const long long nrec = (long long)hc.n_records;
  mm_stitch_records_kernel(0, 1, rec.data(), nrec, chunks.data(), s, endst.data(), endcount.data(), &hc, flag.data(), nrec_filtered);
  std::vector<int> pieces((size_t)nrec + 1, 0);
  mm_pieces_kernel(0, 1, rec.data(), nrec, w, pieces.data(), flag.data(), nrec_filtered);
  std::vector<long long> offs((size_t)nrec + 1);
  long long acc = 0;
  for (long long i = 0; i <= nrec; ++i) { offs[i] = acc; if (i < nrec) acc += pieces[i]; }
  const long long nfin = acc;
  std::vector<MmFinal> fin((size_t)std::max<long long>(nfin, 1));
  mm_expand_kernel(0, 1, rec.data(), nrec, w, pieces.data(), offs.data(), fin.data());
  std::vector<int> perm((size_t)nfin);
  for (long long i = 0; i < nfin; ++i) perm[i] = (int)i;
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
    const MmFinal &x = fin[a], &y = fin[b];
    if (x.seq != y.seq) return x.seq < y.seq;
    if (x.wpos != y.wpos) return x.wpos < y.wpos;
    if (x.wpos_end != y.wpos_end) return x.wpos_end < y.wpos_end;
    return x.hash < y.hash;
  });
  std::vector<int> keep((size_t)nfin + 1, 0);
  if (nfin) mm_unique_flag_kernel(0, 1, fin.data(), perm.data(), nfin, keep.data());
  std::vector<long long> offs2((size_t)nfin + 1);
  acc = 0;
  for (long long i = 0; i <= nfin; ++i) { offs2[i] = acc; if (i < nfin) acc += keep[i]; }
  {
    std::vector<wfb_minmer_t> tmp((size_t)std::max<long long>(acc, 1)), fixed;
    if (nfin) mm_gather_out_kernel(0, 1, fin.data(), perm.data(), keep.data(), offs2.data(), nfin, seqs.data(), tmp.data());
    if (acc > 1) mm_tie_count_kernel(0, 1, tmp.data(), acc, &hc.ties);
    if (hc.ties > 0) {
      tie_sequences = mm_fix_tie_order_host(buf.data(), seqs, k, w, rec.data(), nrec, use_filter ? flag.data() : nullptr, nrec_filtered, tmp.data(), acc, fixed);
      acc = (long long)fixed.size();
    } else fixed.assign(tmp.begin(), tmp.begin() + acc);
    *out_count = acc;
    if (stats) {
      stats->raw_records = (uint64_t)nrec; stats->chunks = (uint64_t)nchunks; stats->stale_absorbed = hc.stale_absorbed; stats->stitch_miss = hc.stitch_miss;
      stats->candidates = hc.candidates; stats->redo_chunks = (uint64_t)n_redo; stats->filtered = (uint64_t)use_filter;
      stats->tie_sequences = (uint64_t)tie_sequences;
    }
    if (acc > out_cap) { wfb_set_last_error_("minmer output buffer too small"); return WFB_ECAP; }
    if (acc) memcpy(out, fixed.data(), sizeof(wfb_minmer_t) * (size_t)acc);
  }
  (void)device;
  return rc;
#endif
}

extern "C" int wfb_minmers_build(int device, const char* const* seq_ptrs, const int64_t* seq_lens, const int32_t* seq_ids, int32_t nseq,
                                 int32_t kmer_size, int32_t window_size, int32_t sketch_size, wfb_minmer_t* out, int64_t out_cap,
                                 int64_t* out_count, wfb_minmer_stats_t* stats) {
  return wfb_minmers_build_impl(device, seq_ptrs, seq_lens, seq_ids, nseq, kmer_size, window_size, sketch_size, out, out_cap, out_count,
                                stats, nullptr);
}
