/*
 * epilogue.cu — record-level driver of the aligner path: what wflign::wavefront::do_biwfa_alignment does around its
 * WFA calls (src/common/wflign/src/wflign.cpp:108-483, PAF branch), for a BATCH of mapping records.
 *
 * The three alignment stages run on the GPU through this library's own batch entry points:
 *   main end-to-end biWFA      -> wfb_align_batch           (wflign.cpp:136-148)
 *   head patches (ends-free)   -> wfb_align_endsfree_batch  (wflign.cpp:240-305)
 *   tail patches (ends-free)   -> wfb_align_endsfree_batch  (wflign.cpp:309-418; needs the head-patched CIGAR)
 * Everything else here is the per-record CIGAR bookkeeping on run-length vectors (O(#runs) per record):
 * erosion of the ends, junction merge, the D/= swizzles and the PAF metrics. The reference does the same on
 * CIGAR strings; the run vector is formatted to text once, when the PAF line is written.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "wfmash_b200.h"

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */
static inline void wfb_set_last_error(const char* msg) { wfb_set_last_error_(msg); }

void wfb_trace_mark_(const char* tag); /* wfa_host.cu */
int wfb_align_batch_view_(wfb_aligner_t* a, const wfb_pair_t* pairs, int32_t n, const float* cost_hint, const char** dense, wfb_aln_result_t* results,
                          wfb_align_stats_t* stats); /* wfa_host.cu */
void wfb_take_endsfree_counters_(wfb_aligner_t* a, double* kernel_ms, uint64_t* h2d, uint64_t* d2h); /* wfa_host.cu */

namespace {

struct Run {
  int32_t n;
  char op;
};
typedef std::vector<Run> Cigar;

/* wfa_edit_cigar_to_string (wflign_swizzle.cpp:359-382) and the compress_cigar lambda (wflign.cpp:182-208):
 * run-length encode M/X/I/D with M written as '='. */
Cigar rle(const char* ops, int n) {
  Cigar c;
  for (int i = 0; i < n;) {
    const char o = ops[i];
    int j = i + 1;
    while (j < n && ops[j] == o) ++j;
    c.push_back({j - i, o == 'M' ? '=' : o});
    i = j;
  }
  return c;
}

void append_text(std::string& s, const Run* r, size_t n) {
  char buf[16];
  for (size_t i = 0; i < n; ++i) {
    const int k = snprintf(buf, sizeof buf, "%d%c", r[i].n, r[i].op);
    s.append(buf, (size_t)k);
  }
}

/* merge_adjacent_ops (wflign.cpp:211-238): concatenate, fusing only the two runs that meet at the junction. */
Cigar join(const Cigar& a, const Run* b, size_t nb) {
  Cigar r(a);
  size_t i = 0;
  if (!r.empty() && nb > 0 && r.back().op == b[0].op) { r.back().n += b[0].n; i = 1; }
  r.insert(r.end(), b + i, b + nb);
  return r;
}

/* merge_cigar_ops (wflign_swizzle.cpp:7-37): fuse every pair of equal neighbours. */
void fuse_all(Cigar& c) {
  size_t w = 0;
  for (size_t i = 0; i < c.size(); ++i) {
    if (w > 0 && c[w - 1].op == c[i].op) c[w - 1].n += c[i].n;
    else c[w++] = c[i];
  }
  c.resize(w);
}

/* erode_short_matches_in_cigar (wflign.cpp:19-106): an indel / short match / indel triple among the first (head)
 * or last (tail) three runs is folded into the two indels when they are of different kinds and both longer than
 * the match. */
void erode_short_matches(Cigar& c, int max_match, bool head) {
  if (c.size() < 3) return;
  size_t lo = 1, hi = c.size() - 1;
  if (head) hi = std::min(hi, (size_t)3);
  else lo = std::max(lo, c.size() - 3);
  bool modified = false;
  for (size_t i = lo; i < hi; ++i) {
    const char o = c[i].op, a = c[i - 1].op, b = c[i + 1].op;
    const bool is_match = (o == 'M' || o == '=' || o == 'X');
    if (!is_match || c[i].n > max_match) continue;
    if (!((a == 'I' && b == 'D') || (a == 'D' && b == 'I'))) continue;
    if (c[i - 1].n > c[i].n && c[i + 1].n > c[i].n) {
      c[i - 1].n += c[i].n;
      c[i + 1].n += c[i].n;
      c[i].n = 0;
      modified = true;
    }
  }
  if (!modified) return;
  size_t w = 0;
  for (size_t i = 0; i < c.size(); ++i) {
    if (c[i].n <= 0) continue;
    if (w > 0 && c[w - 1].op == c[i].op) c[w - 1].n += c[i].n;
    else c[w++] = c[i];
  }
  c.resize(w);
}

struct Erosion {
  uint64_t q = 0, t = 0; /* query / target bases exposed for patching */
  size_t cut = 0;        /* head: runs [0,cut) are replaced; tail: runs [cut,size) are replaced */
};

const int kMinPatchLength = 128;       /* wflign.cpp:169 */
const int kMaxErodeLength = 4096;      /* wflign.cpp:170 */
const int kMinConsecutiveMatches = 11; /* wflign.cpp:171 */

inline bool erode_step(const Run& r, bool& found, Erosion& e) {
  if (r.op == '=' && r.n >= kMinConsecutiveMatches) found = true;
  if (found && e.q >= (uint64_t)kMinPatchLength && e.t >= (uint64_t)kMinPatchLength) return false;
  if (e.q >= (uint64_t)kMaxErodeLength || e.t >= (uint64_t)kMaxErodeLength) return false;
  if (r.op == 'M' || r.op == 'X' || r.op == '=') { e.q += r.n; e.t += r.n; }
  else if (r.op == 'I') e.q += r.n;
  else if (r.op == 'D') e.t += r.n;
  return true;
}

Erosion erode_head(const Cigar& c) { /* wflign.cpp:241-276 */
  Erosion e;
  bool found = false;
  for (size_t i = 0; i < c.size(); ++i) {
    if (!erode_step(c[i], found, e)) break;
    e.cut = i + 1;
  }
  return e;
}

Erosion erode_tail(const Cigar& c) { /* wflign.cpp:311-350 */
  Erosion e;
  e.cut = c.size();
  bool found = false;
  for (size_t i = c.size(); i-- > 0;) {
    if (!erode_step(c[i], found, e)) break;
    e.cut = i;
  }
  return e;
}

/* try_swap_start_pattern (wflign_swizzle.cpp:220-262): "N= Dlen D ..." -> "Dlen D N= ..." when the query prefix also
 * matches the target Dlen bases further on. query / target are the reference's NUL-terminated strings. */
void swap_start(Cigar& c, const char* q, size_t qn, const char* t, size_t tn) {
  if (c.size() < 2 || c[0].op != '=' || c[1].op != 'D') return;
  const int64_t N = c[0].n, D = c[1].n;
  if (N > (int64_t)qn || D + N > (int64_t)tn) return;
  if (memcmp(q, t + D, (size_t)N) != 0) return;
  std::swap(c[0], c[1]);
  fuse_all(c);
}

/* try_swap_end_pattern (wflign_swizzle.cpp:264-300) with its helpers alignment_end_coords (:197-218, which advances
 * only on '=' and 'D') and verify_cigar_alignment (:61-104, which rejects every other operation): the trailing
 * "Dlen D N=" becomes "N= Dlen D" only for CIGARs made of '=' and 'D' alone whose swapped form still matches. */
void swap_end(Cigar& c, const char* q, size_t qn, const char* t, size_t tn) {
  const size_t m = c.size();
  if (m < 2 || c[m - 2].op != 'D' || c[m - 1].op != '=') return;
  const int64_t N = c[m - 1].n, D = c[m - 2].n;
  int64_t endQ = 0, endT = 0;
  for (const Run& r : c) {
    if (r.op == '=') { endQ += r.n; endT += r.n; }
    else if (r.op == 'D') endT += r.n;
  }
  const int64_t qs = endQ - N, ts = endT - N - D;
  if (qs < 0 || ts < 0 || qs + N > (int64_t)qn || ts + N > (int64_t)tn) return;
  if (memcmp(q + qs, t + ts, (size_t)N) != 0) return;
  Cigar s(c);
  std::swap(s[m - 2], s[m - 1]);
  fuse_all(s);
  int64_t qp = 0, tp = 0;
  for (const Run& r : s) {
    if (r.op == '=') {
      if (qp + r.n > (int64_t)qn || tp + r.n > (int64_t)tn) return;
      if (memcmp(q + qp, t + tp, (size_t)r.n) != 0) return;
      qp += r.n; tp += r.n;
    } else if (r.op == 'D') {
      if (tp + r.n > (int64_t)tn) return;
      tp += r.n;
    } else {
      return;
    }
  }
  c.swap(s);
}

struct Metrics {
  uint64_t matches = 0, mismatches = 0, insertions = 0, inserted_bp = 0, deletions = 0, deleted_bp = 0, ref_len = 0, q_len = 0;
};

/* process_compressed_cigar (wflign_patch.cpp:225-283) */
Metrics measure(const Run* r, size_t n) {
  Metrics m;
  for (size_t i = 0; i < n; ++i) {
    const uint64_t k = (uint64_t)r[i].n;
    switch (r[i].op) {
      case 'M': case '=': m.matches += k; m.ref_len += k; m.q_len += k; break;
      case 'X': m.mismatches += k; m.ref_len += k; m.q_len += k; break;
      case 'I': m.insertions++; m.inserted_bp += k; m.q_len += k; break;
      case 'D': m.deletions++; m.deleted_bp += k; m.ref_len += k; break;
      default: break;
    }
  }
  return m;
}

/* float2phred (wflign_patch.cpp:2726-2734) */
double float2phred(double prob) {
  if (prob == 1) return 255;
  const double p = -10 * log10(prob);
  if (p < 0 || p > 255) return 255;
  return p;
}

/* operator<<(double) of a default-constructed ostream: %g with 6 significant digits */
void put_g(std::string& s, double v) {
  char buf[48];
  const int k = snprintf(buf, sizeof buf, "%g", v);
  s.append(buf, (size_t)k);
}
void put_u(std::string& s, uint64_t v) {
  char buf[24];
  const int k = snprintf(buf, sizeof buf, "%llu", (unsigned long long)v);
  s.append(buf, (size_t)k);
}

/* trim_indels + write_alignment_paf (wflign_patch.cpp:139-222, 2611-2724) with aln.i = aln.j = 0, aln.is_rev = false,
 * as do_biwfa_alignment sets them (wflign.cpp:155-162). Returns true when a line was appended. */
bool write_paf(std::string& out, const Cigar& c, const wfb_record_t& r, const wfb_paf_params_t& pp) {
  size_t b = 0, e = c.size();
  uint64_t ref_start = r.target_offset, q_start0 = r.query_offset;
  while (b < e && (c[b].op == 'I' || c[b].op == 'D')) {
    if (c[b].op == 'I') q_start0 += (uint64_t)c[b].n; else ref_start += (uint64_t)c[b].n;
    ++b;
  }
  if (b == e) return false; /* nothing but indels: the reference's metrics are undefined here; emit nothing */
  while (e > b && (c[e - 1].op == 'I' || c[e - 1].op == 'D')) --e;
  const Metrics m = measure(c.data() + b, e - b);
  const double gap_compressed_identity = (double)m.matches / (double)(m.matches + m.mismatches + m.insertions + m.deletions);
  const double block_identity = (double)m.matches / (double)(m.matches + m.mismatches + m.inserted_bp + m.deleted_bp);
  if (!(gap_compressed_identity >= pp.min_identity && m.q_len >= pp.min_alignment_length && block_identity >= pp.min_block_identity))
    return false;
  uint64_t qs, qe;
  if (r.query_is_rev) {
    qs = r.query_offset + (r.query_length - (q_start0 - r.query_offset) - m.q_len);
    qe = r.query_offset + (r.query_length - (q_start0 - r.query_offset));
  } else {
    qs = q_start0;
    qe = q_start0 + m.q_len;
  }
  out.append(r.query_name ? r.query_name : ""); out.push_back('\t');
  put_u(out, r.query_total_length); out.push_back('\t');
  put_u(out, qs); out.push_back('\t');
  put_u(out, qe); out.push_back('\t');
  out.push_back(r.query_is_rev ? '-' : '+'); out.push_back('\t');
  out.append(r.target_name ? r.target_name : ""); out.push_back('\t');
  put_u(out, r.target_total_length); out.push_back('\t');
  put_u(out, ref_start); out.push_back('\t');
  put_u(out, ref_start + m.ref_len); out.push_back('\t');
  put_u(out, m.matches); out.push_back('\t');
  put_u(out, std::max(m.ref_len, m.q_len)); out.push_back('\t');
  put_g(out, std::round(float2phred(1.0 - block_identity))); out.push_back('\t');
  out.append("gi:f:"); put_g(out, gap_compressed_identity); out.push_back('\t');
  out.append("bi:f:"); put_g(out, block_identity); out.push_back('\t');
  out.append("md:f:"); put_g(out, (double)r.mashmap_estimated_identity); out.push_back('\t');
  if (r.chain_length > 0) {
    char buf[64];
    const int k = snprintf(buf, sizeof buf, "ch:Z:%d.%d.%d\t", r.chain_id, r.chain_length, r.chain_pos);
    out.append(buf, (size_t)k);
  }
  out.append("cg:Z:");
  append_text(out, c.data() + b, e - b);
  out.append("\t\n");
  return true;
}

/* write_tag_and_md_string (wflign_patch.cpp:2397-2478) over the trimmed runs; the reference starts the target walk at
 * offset 0 of the slice even when trim_indels removed leading deletions (wflign_patch.cpp:2596-2600), and so does this. */
void put_md(std::string& out, const Run* c, size_t n, const char* target) {
  out.append("MD:Z:");
  char last_op = '\0';
  int64_t last_len = 0, t_off = 0, l_md = 0;
  for (size_t x = 0; x < n; ++x) {
    const char op = c[x].op;
    int64_t len = c[x].n;
    if (last_len) {
      if (last_op == op) len += last_len;
      else if (last_op == '=' || last_op == 'M') { l_md += last_len; t_off += last_len; }
      else if (last_op == 'X') {
        for (int64_t i = 0; i < last_len; ++i) { put_u(out, (uint64_t)l_md); out.push_back(target[t_off + i]); l_md = 0; }
        t_off += last_len;
      } else if (last_op == 'D') {
        put_u(out, (uint64_t)l_md); out.push_back('^');
        out.append(target + t_off, (size_t)last_len);
        l_md = 0;
        t_off += last_len;
      } /* an insertion in the middle changes nothing */
    }
    last_op = op;
    last_len = len;
  }
  if (last_len) {
    if (last_op == '=' || last_op == 'M') put_u(out, (uint64_t)(last_len + l_md));
    else if (last_op == 'X') {
      for (int64_t i = 0; i < last_len; ++i) { put_u(out, (uint64_t)l_md); out.push_back(target[t_off + i]); l_md = 0; }
      out.push_back('0');
    } else if (last_op == 'I') put_u(out, (uint64_t)l_md);
    else if (last_op == 'D') {
      put_u(out, (uint64_t)l_md); out.push_back('^');
      out.append(target + t_off, (size_t)last_len);
      out.push_back('0');
    }
  }
}

/* trim_indels + write_alignment_sam (wflign_patch.cpp:139-222, 2480-2609) with aln.i = aln.j = 0, aln.is_rev = false */
bool write_sam(std::string& out, const Cigar& c, const wfb_record_t& r, const wfb_paf_params_t& pp) {
  size_t b = 0, e = c.size();
  uint64_t ref_start = r.target_offset, q_start0 = r.query_offset;
  while (b < e && (c[b].op == 'I' || c[b].op == 'D')) {
    if (c[b].op == 'I') q_start0 += (uint64_t)c[b].n; else ref_start += (uint64_t)c[b].n;
    ++b;
  }
  if (b == e) return false;
  while (e > b && (c[e - 1].op == 'I' || c[e - 1].op == 'D')) --e;
  const Metrics m = measure(c.data() + b, e - b);
  const double gap_compressed_identity = (double)m.matches / (double)(m.matches + m.mismatches + m.insertions + m.deletions);
  const double block_identity = (double)m.matches / (double)(m.matches + m.mismatches + m.inserted_bp + m.deleted_bp);
  if (!(gap_compressed_identity >= pp.min_identity && m.q_len >= pp.min_alignment_length && block_identity >= pp.min_block_identity))
    return false;
  out.append(r.query_name ? r.query_name : ""); out.push_back('\t');
  out.append(r.query_is_rev ? "16" : "0"); out.push_back('\t');
  out.append(r.target_name ? r.target_name : ""); out.push_back('\t');
  put_u(out, ref_start + 1); out.push_back('\t');
  put_g(out, std::round(float2phred(1.0 - block_identity))); out.push_back('\t');
  append_text(out, c.data() + b, e - b);
  out.append("\t*\t0\t0\t");
  if (pp.no_seq_in_sam) out.push_back('*');
  else out.append(r.query + (q_start0 - r.query_offset), (size_t)m.q_len);
  out.append("\t*\tNM:i:"); put_u(out, m.mismatches + m.inserted_bp + m.deleted_bp);
  out.append("\tgi:f:"); put_g(out, gap_compressed_identity);
  out.append("\tbi:f:"); put_g(out, block_identity);
  out.append("\tmd:f:"); put_g(out, (double)r.mashmap_estimated_identity);
  if (r.chain_length > 0) {
    char buf[96];
    const int k = snprintf(buf, sizeof buf, "\tci:i:%d\tch:Z:%d.%d.%d", r.chain_id, r.chain_id, r.chain_length, r.chain_pos);
    out.append(buf, (size_t)k);
  }
  if (pp.emit_md_tag) {
    out.push_back('\t');
    put_md(out, c.data() + b, e - b, r.target);
  }
  out.push_back('\n');
  return true;
}

struct Patch {
  int rec;
  Erosion er;
};

/* The per-record host loops (run-length conversion of the kernels' ops, erosion scans, PAF / SAM text) are independent per
 * record: spread them over the host cores the way the reference spreads records over its worker threads
 * (computeAlignments.hpp:695-720). Records are taken in blocks of `grain` from an atomic counter; every result lands in
 * its own slot, so the output does not depend on the schedule. */
#ifndef WFB_HOST_PAR_MIN_BYTES
#define WFB_HOST_PAR_MIN_BYTES (1 << 20) /* tests build the emulation with 0 to force the threaded path on small cases */
#endif
template <class F>
void for_each_record(int64_t n, int64_t work_bytes, F f) {
  int nt = (int)std::min<int64_t>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u), n / 4);
  if (work_bytes < (int64_t)WFB_HOST_PAR_MIN_BYTES) nt = 1; /* small batches: a thread launch costs more than the loop */
  if (nt <= 1) { for (int64_t i = 0; i < n; ++i) f(i); return; }
  const int64_t grain = std::max<int64_t>(1, n / (8 * (int64_t)nt));
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(grain);
      if (b >= n) return;
      const int64_t e = std::min(n, b + grain);
      for (int64_t i = b; i < e; ++i) f(i);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
}

/* One ends-free patch request: record, which end, and the erosion (wflign.cpp:241-276 head, :311-350 tail) computed on the record's
 * main CIGAR. */
struct PatchReq {
  int rec;
  bool head;
  Erosion er;
};

/* Runs one batch of ends-free patch alignments and splices the results into cig[]: a head patch replaces runs [0,cut), a tail patch
 * runs [cut,size). Requests are spliced in the order given; a record with both a head and a tail request in the same batch must list
 * the tail first (the head splice changes the run indices). A patch whose alignment hit a device capacity keeps the main CIGAR, as
 * the reference does when a patch aligner fails (wflign.cpp:299, :385: `if (head_status == 0)` / `if (tail_status == 0)`), and is counted. */
int patch_batch(wfb_aligner_t* a, const wfb_record_t* recs, std::vector<Cigar>& cig, const std::vector<PatchReq>& all, int term_group, int64_t* failed) {
  if (all.empty()) return WFB_OK;
  /* the reference aligner answers a zero-length side with a pure gap (verified against the compiled reference);
   * those need no kernel */
  std::vector<wfb_endsfree_pair_t> pairs;
  std::vector<int> slot(all.size(), -1);
  int64_t cap = 16;
  for (size_t j = 0; j < all.size(); ++j) {
    const wfb_record_t& r = recs[all[j].rec];
    const Erosion& e = all[j].er;
    if (e.q == 0 || e.t == 0) continue;
    wfb_endsfree_pair_t p;
    if (all[j].head) { /* wflign.cpp:286-297: alignEndsFree(head_target, target_eroded, 0, head_query, query_eroded, 0) */
      p.pattern = r.target; p.text = r.query;
      p.pattern_begin_free = (int32_t)e.t; p.pattern_end_free = 0; p.text_begin_free = (int32_t)e.q; p.text_end_free = 0;
    } else {    /* wflign.cpp:368-389: alignEndsFree(tail_target, 0, tail_target_length, tail_query, 0, tail_query_length) */
      p.pattern = r.target + (r.target_length - e.t); p.text = r.query + (r.query_length - e.q);
      p.pattern_begin_free = 0; p.pattern_end_free = (int32_t)e.t; p.text_begin_free = 0; p.text_end_free = (int32_t)e.q;
    }
    p.pattern_len = (int32_t)e.t; p.text_len = (int32_t)e.q;
    slot[j] = (int)pairs.size();
    pairs.push_back(p);
    cap += (int64_t)(e.t + e.q);
  }
  std::vector<char> ops((size_t)cap);
  std::vector<wfb_aln_result_t> res(pairs.size());
  if (!pairs.empty()) {
    const int rc = wfb_align_endsfree_batch(a, pairs.data(), (int32_t)pairs.size(), term_group, ops.data(), cap, res.data());
    if (rc != WFB_OK) return rc;
  }
  /* splice: the requests of a record are adjacent in `all` (tail first) and touch only that record's CIGAR, so the records are spliced on
   * all host cores (a head splice copies the whole main CIGAR: 21 k records x ~20 KB on one thread was 100 ms of the phase) */
  std::vector<size_t> group; /* first request of each record */
  for (size_t j = 0; j < all.size(); ++j) if (j == 0 || all[j].rec != all[j - 1].rec) group.push_back(j);
  group.push_back(all.size());
  std::atomic<int64_t> n_failed(0);
  for_each_record((int64_t)group.size() - 1, cap, [&](int64_t g) {
    for (size_t j = group[(size_t)g]; j < group[(size_t)g + 1]; ++j) {
      const int i = all[j].rec;
      const Erosion& e = all[j].er;
      const bool head = all[j].head;
      Cigar pc;
      if (slot[j] < 0) {
        pc.push_back({(int32_t)(e.q == 0 ? e.t : e.q), e.q == 0 ? 'D' : 'I'});
      } else {
        const wfb_aln_result_t& rr = res[(size_t)slot[j]];
        if (rr.status != 0) { n_failed.fetch_add(1); continue; }
        pc = rle(ops.data() + rr.ops_offset, rr.ops_len);
      }
      erode_short_matches(pc, 3, head);
      Cigar& c = cig[(size_t)i];
      if (head) {
        c = join(pc, c.data() + e.cut, c.size() - e.cut);
      } else { /* join(runs [0,cut), patch) in place */
        c.resize(e.cut);
        size_t i0 = 0;
        if (!c.empty() && !pc.empty() && c.back().op == pc[0].op) { c.back().n += pc[0].n; i0 = 1; }
        c.insert(c.end(), pc.begin() + (long)i0, pc.end());
      }
    }
  });
  if (failed) *failed += n_failed.load();
  return WFB_OK;
}

/* Head and tail patches of a batch of records (wflign.cpp:241-405). The reference patches the head, then erodes the tail of the
 * head-patched CIGAR. The tail erosion examines runs [cut_tail - 1, size) only; when all of them lie behind the run that follows the
 * head's replaced prefix (the one run the head splice can change, by fusing with the patch's last run), it sees the same runs before
 * and after the head patch, so both patches are aligned in ONE device round and spliced tail-first. The remaining records (short
 * CIGARs whose two ends meet) take the reference's order in a second round. */
int patch_records(wfb_aligner_t* a, const wfb_record_t* recs, std::vector<Cigar>& cig, const std::vector<int32_t>& status, int term_group, int64_t* failed) {
  std::vector<PatchReq> round1, round2;
  std::vector<int> dependent;
  const bool fuse = !(getenv("WFB_PATCH_FUSE") && atoi(getenv("WFB_PATCH_FUSE")) == 0);
  for (size_t i = 0; i < cig.size(); ++i) {
    if (status[i] != WFB_REC_WRITTEN) continue;
    const Erosion eh = erode_head(cig[i]); /* O(runs of the eroded end): not worth a thread */
    const bool need_head = eh.q > 3 || eh.t > 3;
    if (!need_head) { /* the CIGAR the tail erosion sees is the main one */
      const Erosion et = erode_tail(cig[i]);
      if (et.q > 3 || et.t > 3) round1.push_back({(int)i, false, et});
      continue;
    }
    const Erosion et = erode_tail(cig[i]);
    const bool independent = fuse && et.cut >= 1 && et.cut - 1 > eh.cut;
    if (independent) {
      if (et.q > 3 || et.t > 3) round1.push_back({(int)i, false, et});
      round1.push_back({(int)i, true, eh});
    } else {
      round1.push_back({(int)i, true, eh});
      dependent.push_back((int)i);
    }
  }
  int rc = patch_batch(a, recs, cig, round1, term_group, failed);
  if (rc != WFB_OK) return rc;
  for (int i : dependent) {
    const Erosion et = erode_tail(cig[(size_t)i]);
    if (et.q > 3 || et.t > 3) round2.push_back({i, false, et});
  }
  return patch_batch(a, recs, cig, round2, term_group, failed);
}

} // namespace

extern "C" int wfb_biwfa_paf_batch(wfb_aligner_t* a, const wfb_record_t* recs, int32_t n, const wfb_paf_params_t* params, char* out,
                                   int64_t out_cap, int64_t* out_len, int64_t* line_offset, int32_t* rec_status, wfb_align_stats_t* stats) {
  if (!a || n < 0 || !params || !out_len || (n > 0 && (!recs || !line_offset || !rec_status))) { wfb_set_last_error("bad argument"); return WFB_EINVAL; }
  int term_group = params->term_group == 0 ? 8 : params->term_group;
  if (term_group < 0) { /* "what a -march=native build of the reference on THIS host does": WFA2-lib compiles its AVX-512 extend kernels under
                           __AVX512CD__ && __AVX512VL__, its AVX2 kernels under __AVX2__ (wavefront_extend_kernels_avx.h:35,58), scalar otherwise */
#if defined(__x86_64__) && defined(__GNUC__)
    term_group = (__builtin_cpu_supports("avx512cd") && __builtin_cpu_supports("avx512vl")) ? 16 : __builtin_cpu_supports("avx2") ? 8 : 1;
#else
    term_group = 1;
#endif
  }
  *out_len = 0;
  if (n == 0) return WFB_OK;
  std::vector<wfb_pair_t> pairs((size_t)n);
  int64_t cap = 16;
  for (int i = 0; i < n; ++i) {
    const wfb_record_t& r = recs[i];
    if (!r.query || !r.target || r.query_length > (uint64_t)INT32_MAX || r.target_length > (uint64_t)INT32_MAX) {
      wfb_set_last_error("bad record (NULL sequence or length > INT32_MAX)");
      return WFB_EINVAL;
    }
    pairs[(size_t)i] = {r.target, (int32_t)r.target_length, r.query, (int32_t)r.query_length}; /* wflign.cpp:148 */
    cap += (int64_t)(r.target_length + r.query_length);
  }
  wfb_trace_mark_("paf_batch: begin");
  std::vector<wfb_aln_result_t> res((size_t)n);
  /* scheduling hint: expected edits from the mapping's identity estimate (the same quantity the reference's own progress / cost
   * heuristics use); it only orders the work, the alignments do not depend on it */
  std::vector<float> hint((size_t)n);
  for (int i = 0; i < n; ++i) {
    const float id = recs[i].mashmap_estimated_identity;
    const float d = (id > 0.f && id <= 1.f) ? std::max(1.f - id, 0.002f) : 0.05f;
    hint[(size_t)i] = d * (float)std::max(recs[i].query_length, recs[i].target_length);
  }
  /* the operation strings stay in the aligner's pinned D2H buffer and are run-length encoded from there (a GB-sized pageable copy and a
   * second pass over it otherwise); the buffer is reused by the patch rounds, so the conversion below comes first */
  const char* ops = nullptr;
  int rc = wfb_align_batch_view_(a, pairs.data(), n, hint.data(), &ops, res.data(), stats);
  if (rc != WFB_OK) return rc;
  wfb_trace_mark_("paf_batch: main alignments (wfb_align_batch_hinted)");
  std::vector<Cigar> cig((size_t)n);
  std::vector<int32_t> status((size_t)n, WFB_REC_WRITTEN);
  std::atomic<uint64_t> main_cap(0);
  for_each_record(n, cap, [&](int64_t i) {
    if (res[(size_t)i].status != 0) { /* wflign.cpp:150-152 */
      status[(size_t)i] = WFB_REC_UNALIGNED;
      if (res[(size_t)i].status != -3 /* WFB_PAIR_UNATTAINABLE: the reference's own failure */) main_cap.fetch_add(1);
      return;
    }
    cig[(size_t)i] = rle(ops + res[(size_t)i].ops_offset, res[(size_t)i].ops_len);
  });
  wfb_trace_mark_("paf_batch: run-length conversion");
  if (stats) stats->main_device_cap = main_cap.load();
  if (!params->disable_chain_patching) {
    double ms = 0; uint64_t h2d = 0, d2h = 0;
    wfb_take_endsfree_counters_(a, &ms, &h2d, &d2h);
    int64_t patch_failed = 0;
    rc = patch_records(a, recs, cig, status, term_group, &patch_failed);
    if (rc != WFB_OK) return rc;
    wfb_trace_mark_("paf_batch: head + tail patches");
    if (stats) stats->patch_cap_kept_main = (uint64_t)patch_failed;
    wfb_take_endsfree_counters_(a, &ms, &h2d, &d2h);
    if (stats) { stats->patch_kernel_ms = ms; stats->h2d_bytes += h2d; stats->d2h_bytes += d2h; }
  }
  std::vector<std::string> line((size_t)n);
  for_each_record(n, cap, [&](int64_t i) {
    rec_status[i] = status[(size_t)i];
    if (status[(size_t)i] != WFB_REC_WRITTEN) return;
    const wfb_record_t& r = recs[i];
    /* the reference builds std::string(query) / std::string(target) from the char* (wflign.cpp:422-430): NUL-terminated views */
    const size_t qn = strnlen(r.query, (size_t)r.query_length), tn = strnlen(r.target, (size_t)r.target_length);
    Cigar& c = cig[(size_t)i];
    swap_start(c, r.query, qn, r.target, tn);
    swap_end(c, r.query, qn, r.target, tn);
    std::string& text = line[(size_t)i];
    if (!(params->sam_format ? write_sam(text, c, r, *params) : write_paf(text, c, r, *params))) { rec_status[i] = WFB_REC_FILTERED; text.clear(); }
  });
  wfb_trace_mark_("paf_batch: swizzle + text");
  int64_t total = 0;
  for (int i = 0; i < n; ++i) { line_offset[i] = total; total += (int64_t)line[(size_t)i].size(); }
  line_offset[n] = total;
  *out_len = total;
  if (total > out_cap || !out) { wfb_set_last_error("PAF output buffer too small (see *out_len)"); return WFB_ECAP; }
  for (int i = 0; i < n; ++i) memcpy(out + line_offset[i], line[(size_t)i].data(), line[(size_t)i].size());
  wfb_trace_mark_("paf_batch: text copied out");
  return WFB_OK;
}
