// chain_host.cu — host stage between the L2 mappings and the aligner's records (first part):
// chaining of the fragment mappings of one query, merge of every chain into pieces of at most max_mapping_length
// (the <= 50 kb records the aligner sees) and the chain tags of the mapping PAF (ch:Z:id.pos.len).
//
// Replaces skch::MappingFilterUtils::mergeMappingsInRangeWithChains (src/map/include/mappingFilter.hpp:381-571) with
// its union-find (src/common/dset64.hpp), called by Map::filterSubsetMappings (src/map/include/computeMap.hpp:1076-1094).
// Pure host C++ (the reference runs this per query on the host as well); a batch of queries is spread over host threads.
// The arithmetic keeps the reference's integer widths (uint32 end positions that wrap, int max_dist, the signed /
// unsigned comparison against max_mapping_length) and its two std::sort calls over index vectors, so that ties fall
// the same way with the same libstdc++.
#include "wfmash_b200.h"

#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <limits>
#include <map>
#include <numeric>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */

namespace {

inline int ch_strand(const wfb_mapping_t& m) { return (m.flags & 0x01) ? -1 : 1; } /* MappingResult::strand(), base_types.hpp:166-168 */
/* refEndPos() / queryEndPos(): uint32 + uint32 evaluated in 32 bits, then widened (base_types.hpp:214-220) */
inline int64_t ch_ref_end(const wfb_mapping_t& m) { return (int64_t)(uint32_t)(m.refStartPos + m.blockLength); }
inline int64_t ch_query_end(const wfb_mapping_t& m) { return (int64_t)(uint32_t)(m.queryStartPos + m.blockLength); }

/* dsets::DisjointSets (src/common/dset64.hpp): union by rank; equal ranks hang the LARGER id under the smaller one.
 * Which element ends up as a root decides the order of the chains, so the rule is kept; path compression is not
 * observable. */
struct ChainSets {
  std::vector<uint64_t> parent, rank;
  explicit ChainSets(size_t n) : parent(n), rank(n, 0) { std::iota(parent.begin(), parent.end(), 0); }
  uint64_t find(uint64_t x) {
    while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
    return x;
  }
  void unite(uint64_t a, uint64_t b) {
    a = find(a); b = find(b);
    if (a == b) return;
    uint64_t ra = rank[a], rb = rank[b];
    if (ra > rb || (ra == rb && a < b)) { std::swap(ra, rb); std::swap(a, b); }
    parent[a] = b; /* a: lower rank, or equal rank and larger id */
    if (ra == rb) rank[b] = rb + 1;
  }
};

template <typename T>
void ch_reorder(std::vector<T>& v, const std::vector<uint32_t>& p) {
  std::vector<T> out(v.size());
  for (size_t i = 0; i < p.size(); ++i) out[i] = v[p[i]];
  v.swap(out);
}

}  // namespace

/* also used by filter_host.cu (the scaffold filter chains with a second gap) */
void wfb_chain_one_query_(const wfb_chain_params_t& P, wfb_mapping_t* io, int64_t n, std::vector<wfb_mapping_t>& merged,
                          std::vector<wfb_chain_info_t>& info) {
  merged.clear();
  info.clear();
  if (!P.split || n < 2) { /* :390-399: every mapping is its own chain */
    for (int64_t i = 0; i < n; ++i) {
      merged.push_back(io[i]);
      info.push_back(wfb_chain_info_t{(uint32_t)i, 1, 1});
    }
    return;
  }
  std::vector<wfb_mapping_t> m(io, io + n);
  const int max_dist = (int)P.chain_gap; /* the reference passes param.chain_gap through an int parameter */
  const double NO_SCORE = std::numeric_limits<double>::max();
  std::vector<int64_t> id(n);                    /* MappingAuxData::splitMappingId */
  std::vector<double> pair_score(n, NO_SCORE);   /* chainPairScore */
  std::vector<int64_t> pair_id(n, std::numeric_limits<int64_t>::min());
  std::iota(id.begin(), id.end(), 0);
  std::vector<uint32_t> p(n);
  std::iota(p.begin(), p.end(), 0);
  std::sort(p.begin(), p.end(), [&](uint32_t i, uint32_t j) { /* :410-419 */
    const wfb_mapping_t &a = m[i], &b = m[j];
    const int16_t as = (int16_t)ch_strand(a), bs = (int16_t)ch_strand(b);
    return std::tie(a.refSeqId, as, a.queryStartPos, a.refStartPos) < std::tie(b.refSeqId, bs, b.queryStartPos, b.refStartPos);
  });
  ch_reorder(m, p); ch_reorder(id, p); ch_reorder(pair_score, p); ch_reorder(pair_id, p);

  ChainSets sets((size_t)n);
  size_t g0 = 0;
  while (g0 < (size_t)n) { /* one (refSeqId, strand) group at a time, :428-474 */
    size_t g1 = g0 + 1;
    while (g1 < (size_t)n && m[g1].refSeqId == m[g0].refSeqId && ch_strand(m[g1]) == ch_strand(m[g0])) ++g1;
    for (size_t i = g0; i < g1; ++i) {
      if (pair_score[i] != NO_SCORE) sets.unite((uint64_t)id[i], (uint64_t)pair_id[i]);
      double best = NO_SCORE;
      size_t best_j = g1;
      const int64_t qe_i = ch_query_end(m[i]), re_i = ch_ref_end(m[i]);
      const bool fwd = ch_strand(m[i]) == 1;
      for (size_t j = i + 1; j < g1; ++j) {
        if ((int64_t)m[j].queryStartPos > qe_i + max_dist) break;
        int64_t q_dist = (int64_t)m[j].queryStartPos - qe_i;
        if (q_dist < 0) q_dist = 0;
        const int64_t r_dist = fwd ? ((int64_t)m[j].refStartPos - re_i) : ((int64_t)m[i].refStartPos - ch_ref_end(m[j]));
        if (q_dist <= max_dist && r_dist >= -P.window_length / 5 && r_dist <= max_dist) {
          const double d2 = (double)q_dist * q_dist + (double)r_dist * r_dist;
          if (d2 < best && d2 < pair_score[j]) { best = d2; best_j = j; }
        }
      }
      if (best_j != g1) { pair_score[best_j] = best; pair_id[best_j] = id[i]; }
    }
    g0 = g1;
  }
  for (int64_t i = 0; i < n; ++i) /* :477-481 */
    if (pair_score[i] != NO_SCORE) sets.unite((uint64_t)id[i], (uint64_t)pair_id[i]);
  for (int64_t i = 0; i < n; ++i) id[i] = (int64_t)sets.find((uint64_t)id[i]);

  std::iota(p.begin(), p.end(), 0);
  std::sort(p.begin(), p.end(), [&](uint32_t i, uint32_t j) { /* :487-492 */
    return std::tie(id[i], m[i].queryStartPos, m[i].refStartPos) < std::tie(id[j], m[j].queryStartPos, m[j].refStartPos);
  });
  ch_reorder(m, p); ch_reorder(id, p);

  std::map<uint32_t, uint32_t> chain_of; /* splitMappingId -> sequential chain id, :498-520 */
  uint32_t next_chain = 0;
  size_t i = 0;
  while (i < (size_t)n) {
    size_t j = i;
    while (j + 1 < (size_t)n && id[j + 1] == id[i]) ++j;
    uint32_t chain_id;
    auto it = chain_of.find((uint32_t)id[i]);
    if (it == chain_of.end()) { chain_of[(uint32_t)id[i]] = next_chain; chain_id = next_chain++; }
    else chain_id = it->second;
    const uint16_t chain_len = (uint16_t)(j - i + 1);
    uint16_t chain_pos = 1;
    size_t f0 = i;
    while (f0 <= j) { /* pieces whose query AND reference span stay below max_mapping_length, :526-566 */
      size_t f1 = f0;
      while (f1 + 1 <= j) {
        const int64_t qspan = ch_query_end(m[f1 + 1]) - (int64_t)m[f0].queryStartPos;
        const int64_t rspan = ch_ref_end(m[f1 + 1]) - (int64_t)m[f0].refStartPos;
        if ((uint64_t)std::max(qspan, rspan) >= P.max_mapping_length) break; /* signed max compared as unsigned, like the reference */
        ++f1;
      }
      wfb_mapping_t out = m[f0];
      const uint32_t q_start = m[f0].queryStartPos, q_end = (uint32_t)ch_query_end(m[f1]);
      uint32_t r_start = m[f0].refStartPos, r_end = (uint32_t)ch_ref_end(m[f1]);
      double total_id = 0, total_comp = 0;
      uint32_t conserved = 0;
      for (size_t k = f0; k <= f1; ++k) {
        total_id += (float)(m[k].nucIdentity / 10000.0f);   /* getNucIdentity() */
        total_comp += (float)(m[k].kmerComplexity / 100.0f); /* getKmerComplexity() */
        conserved += m[k].conservedSketches;
        if (ch_strand(out) == -1) {
          r_start = std::min(r_start, m[k].refStartPos);
          r_end = std::max(r_end, (uint32_t)ch_ref_end(m[k]));
        }
      }
      out.queryStartPos = q_start;
      out.refStartPos = ch_strand(out) == 1 ? r_start : m[f1].refStartPos;
      out.blockLength = std::max(q_end - q_start, r_end - r_start);
      out.n_merged = (uint32_t)(f1 - f0 + 1);
      out.nucIdentity = (uint16_t)roundf((float)(total_id / out.n_merged) * 10000.0f);   /* setNucIdentity(float) */
      out.kmerComplexity = (uint8_t)roundf((float)(total_comp / out.n_merged) * 100.0f); /* setKmerComplexity(float) */
      out.conservedSketches = conserved;
      merged.push_back(out);
      info.push_back(wfb_chain_info_t{chain_id, chain_pos++, chain_len});
      f0 = f1 + 1;
    }
    i = j + 1;
  }
  std::copy(m.begin(), m.end(), io); /* the reference leaves readMappings in this order */
}

extern "C" int wfb_chain_mappings_batch(const wfb_chain_params_t* params, wfb_mapping_t* mappings, const int64_t* query_offset, int32_t n_queries,
                                        wfb_mapping_t* merged, wfb_chain_info_t* chain_info, int64_t merged_cap, int64_t* merged_offset,
                                        int32_t host_threads) {
  if (!params || n_queries < 0 || (n_queries > 0 && (!mappings || !query_offset || !merged || !chain_info || !merged_offset))) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  if (params->max_mapping_length == 0 || params->window_length <= 0) { wfb_set_last_error_("bad chain parameters"); return WFB_EINVAL; }
  for (int32_t q = 0; q < n_queries; ++q)
    if (query_offset[q + 1] < query_offset[q] || query_offset[q + 1] - query_offset[q] > 0xFFFFFFFFLL) { wfb_set_last_error_("bad query_offset"); return WFB_EINVAL; }
  std::vector<std::vector<wfb_mapping_t>> out((size_t)n_queries);
  std::vector<std::vector<wfb_chain_info_t>> inf((size_t)n_queries);
  const int nt = std::max(1, std::min<int>(host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency(), n_queries));
  auto work = [&](int t) {
    for (int32_t q = t; q < n_queries; q += nt)
      wfb_chain_one_query_(*params, mappings + query_offset[q], query_offset[q + 1] - query_offset[q], out[(size_t)q], inf[(size_t)q]);
  };
  if (nt == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  int64_t tot = 0;
  if (n_queries > 0) merged_offset[0] = 0;
  for (int32_t q = 0; q < n_queries; ++q) { tot += (int64_t)out[(size_t)q].size(); merged_offset[q + 1] = tot; }
  if (tot > merged_cap) { wfb_set_last_error_("merged buffer too small"); return WFB_ECAP; }
  for (int32_t q = 0; q < n_queries; ++q) {
    std::copy(out[(size_t)q].begin(), out[(size_t)q].end(), merged + merged_offset[q]);
    std::copy(inf[(size_t)q].begin(), inf[(size_t)q].end(), chain_info + merged_offset[q]);
  }
  return WFB_OK;
}

extern "C" int wfb_l2_to_query_mappings(const wfb_l2_mapping_t* l2, int64_t n, const int32_t* frag_index, int64_t window_length, int64_t query_len,
                                        const int64_t* ref_seq_len, wfb_mapping_t* out) {
  if (n < 0 || (n > 0 && (!l2 || !frag_index || !ref_seq_len || !out)) || window_length <= 0) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  for (int64_t i = 0; i < n; ++i) {
    wfb_mapping_t e;
    /* Map::doL2Mapping, computeMap.hpp:1029-1044 */
    e.refSeqId = (uint32_t)l2[i].refSeqId;
    e.refStartPos = (uint32_t)l2[i].refStartPos;
    e.queryStartPos = 0;
    e.blockLength = (uint32_t)window_length; /* Q.len: fragments are window_length long */
    e.conservedSketches = (uint32_t)l2[i].conservedSketches;
    e.n_merged = 1;
    e.nucIdentity = (uint16_t)roundf(l2[i].nucIdentity * 10000.0f);
    e.kmerComplexity = (uint8_t)roundf(l2[i].kmerComplexity * 100.0f);
    e.flags = l2[i].strand == -1 ? 0x01 : 0x00;
    /* Map::processFragment, computeMap.hpp:123-127 */
    e.queryStartPos += (uint32_t)(frag_index[l2[i].frag] * window_length);
    /* OutputHandler::mappingBoundarySanityCheck, mappingOutput.hpp:31-69 (the "< 0" tests can never fire on uint32) */
    const int64_t rlen = ref_seq_len[e.refSeqId];
    if ((int64_t)e.refStartPos >= rlen) e.refStartPos = (uint32_t)(rlen - 1);
    if (ch_ref_end(e) < (int64_t)e.refStartPos) e.blockLength = 0;
    if (ch_ref_end(e) >= rlen) e.blockLength = (uint32_t)(rlen - 1 - (int64_t)e.refStartPos);
    if ((int64_t)e.queryStartPos >= query_len) e.queryStartPos = (uint32_t)query_len;
    if (ch_query_end(e) < (int64_t)e.queryStartPos) e.blockLength = 0;
    if (ch_query_end(e) >= query_len) e.blockLength = (uint32_t)(query_len - (int64_t)e.queryStartPos);
    out[i] = e;
  }
  return WFB_OK;
}
