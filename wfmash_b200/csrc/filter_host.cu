// filter_host.cu — host stage between the chain merge and the aligner (second part, and b3):
//   * every filter Map::filterSubsetMappings applies to the mappings of one query
//     (src/map/include/computeMap.hpp:1076-1165): weak-mapping filter, group plane sweep, length-mismatch filter,
//     sparsification, scaffold filter (src/map/include/mappingFilter.hpp:154-293,831-1016; src/map/include/filter.hpp),
//   * the reference-axis plane sweep and the regrouping of the one-to-one mode (computeMap.hpp:788-850),
//   * the mapping PAF writer (src/map/include/mappingOutput.hpp:74-139) and its reader on the aligner side
//     (src/align/include/computeAlignments.hpp:195-303,582-660).
// Pure host C++ (the reference runs these per query on the host too); a batch of queries is spread over host threads.
//
// Exactness notes. The reference's results depend on details that are reproduced on purpose:
//   - its unstable std::sort calls: the same comparators run over the same input order with the same libstdc++;
//   - `param.numMappingsForSegment - 1` travels through an `int` parameter, so the default "inf" (UINT32_MAX) arrives
//     as -2 and the plane sweep keeps the best-scoring mapping(s) of every sweep position only;
//   - the sweep status is a std::set keyed by (score, start, refSeqId): two mappings with equal keys are ONE element;
//   - ChainInfo is paired with the filtered mappings by position, not by identity;
//   - the scaffold distance is evaluated in float, uint32 coordinates rounded to 24 bits included.
#include "wfmash_b200.h"

#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <limits>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */
void wfb_chain_one_query_(const wfb_chain_params_t& P, wfb_mapping_t* io, int64_t n, std::vector<wfb_mapping_t>& merged,
                          std::vector<wfb_chain_info_t>& info); /* chain_host.cu */

namespace {

typedef std::vector<wfb_mapping_t> MapVec;

inline int16_t fl_strand(const wfb_mapping_t& m) { return (m.flags & 0x01) ? (int16_t)-1 : (int16_t)1; }
inline bool fl_discard(const wfb_mapping_t& m) { return (m.flags & 0x02) != 0; }
inline bool fl_overlapped(const wfb_mapping_t& m) { return (m.flags & 0x04) != 0; }
inline void fl_set_discard(wfb_mapping_t& m, bool d) { if (d) m.flags |= 0x02; else m.flags &= (uint8_t)~0x02; }
inline void fl_set_overlapped(wfb_mapping_t& m, bool o) { if (o) m.flags |= 0x04; else m.flags &= (uint8_t)~0x04; }
/* uint32 sums that wrap before they are widened (base_types.hpp:214-220) */
inline int64_t fl_ref_end(const wfb_mapping_t& m) { return (int64_t)(uint32_t)(m.refStartPos + m.blockLength); }
inline int64_t fl_query_end(const wfb_mapping_t& m) { return (int64_t)(uint32_t)(m.queryStartPos + m.blockLength); }
inline float fl_identity(const wfb_mapping_t& m) { return m.nucIdentity / 10000.0f; }
inline float fl_complexity(const wfb_mapping_t& m) { return m.kmerComplexity / 100.0f; }

/* MappingResult::hash (base_types.hpp:232-242): boost-style hash_combine over std::hash of the integer fields (= identity) */
inline uint64_t fl_hash(const wfb_mapping_t& m) {
  uint64_t s = 0;
  auto mix = [&s](uint64_t v) { s ^= v + 0x9e3779b9ull + (s << 6) + (s >> 2); };
  mix(m.refSeqId); mix(m.refStartPos); mix(m.queryStartPos); mix(m.blockLength); mix(m.nucIdentity); mix(m.conservedSketches); mix(m.flags);
  return s;
}

/* ---- plane sweep (filter.hpp). AXIS 0 = query (namespace query), 1 = reference (namespace ref). ---- */
template <int AXIS>
struct Sweep {
  MapVec& v;
  const int64_t* ref_len;
  explicit Sweep(MapVec& vec, const int64_t* rl) : v(vec), ref_len(rl) {}

  double score(int x) const {
    if (AXIS == 0) { /* filter.hpp:46-51 */
      if (v[x].blockLength <= 0 || fl_identity(v[x]) <= 0) return std::numeric_limits<double>::lowest();
      return fl_identity(v[x]) * std::log(static_cast<double>(v[x].blockLength));
    }
    return fl_identity(v[x]) * log((double)v[x].blockLength); /* filter.hpp:335: no guards (log 0 = -inf, 0 * -inf = NaN) */
  }
  /* order of the sweep-line status: best score first (filter.hpp:55-65, 339-349) */
  bool operator()(const int x, const int y) const {
    const double xs = score(x), ys = score(y);
    if (AXIS == 0) return std::tie(xs, v[x].queryStartPos, v[x].refSeqId) > std::tie(ys, v[y].queryStartPos, v[y].refSeqId);
    return std::tie(xs, v[x].refStartPos) > std::tie(ys, v[y].refStartPos);
  }
  double overlap(int x, int y) const { /* filter.hpp:79-87, 363-371 */
    const int64_t xs = AXIS == 0 ? v[x].queryStartPos : v[x].refStartPos, ys = AXIS == 0 ? v[y].queryStartPos : v[y].refStartPos;
    const int64_t xe = AXIS == 0 ? fl_query_end(v[x]) : fl_ref_end(v[x]), ye = AXIS == 0 ? fl_query_end(v[y]) : fl_ref_end(v[y]);
    const int64_t o0 = std::max((uint32_t)xs, (uint32_t)ys), o1 = std::min(xe, ye);
    const int64_t olen = std::max(0, static_cast<int>(o1 - o0));
    return static_cast<double>(olen) / std::min(xe - xs, ye - ys);
  }

  template <typename Status>
  void mark_good(Status& L, int secondary, bool drop_rand, double overlap_threshold) { /* filter.hpp:94-163, 378-447 */
    auto first = L.begin();
    int kept = 0;
    auto it = L.begin();
    for (; it != L.end(); ++it) {
      if ((score(*first) > score(*it) || !fl_discard(v[*it])) && kept > secondary) break;
      fl_set_discard(v[*it], false);
      ++kept;
    }
    const auto kept_end = it;
    if (overlap_threshold < 1.0) {
      for (; it != L.end(); ++it) {
        if (it == L.begin()) continue;
        const int idx = *it;
        for (auto k = L.begin(); k != kept_end; ++k)
          if (overlap(idx, *k) > overlap_threshold) {
            fl_set_overlapped(v[idx], true);
            fl_set_discard(v[idx], true);
            break;
          }
      }
    }
    if (kept > secondary && drop_rand) { /* ties broken by the hash of the mapping; the reference's last key is the address */
      std::vector<std::tuple<double, uint64_t, int>> cand;
      for (auto k = L.begin(); k != L.end(); ++k)
        if (!fl_discard(v[*k])) cand.emplace_back(score(*k), fl_hash(v[*k]), *k);
      std::sort(cand.begin(), cand.end(), std::greater<std::tuple<double, uint64_t, int>>());
      kept = 0;
      for (auto& c : cand) fl_set_discard(v[std::get<2>(c)], true);
      for (auto& c : cand) {
        if (kept > secondary) break;
        fl_set_discard(v[std::get<2>(c)], false);
        ++kept;
      }
    }
  }

  /* liFilterAlgorithm (filter.hpp:170-240) / ref::filterMappings (filter.hpp:474-535) */
  void run(int secondary, bool drop_rand, double overlap_threshold) {
    if (v.size() <= 1) return;
    for (auto& e : v) {
      fl_set_discard(e, true);
      if (AXIS == 0) fl_set_overlapped(e, false);
    }
    std::set<int, Sweep<AXIS>> status(*this);
    /* (sequence, position, kind, id); kind 1 = BEGIN, 2 = END (base_types.hpp:108-112). The reference's schedule also
     * holds 2n zero-initialised records (its vector is sized AND appended to): they sort in front of everything and
     * each erases id 0 from the still empty status, but they make position (0, 0) a sweep stop of its own. */
    typedef std::tuple<int32_t, int64_t, int, int> Event;
    std::vector<Event> ev(2 * v.size(), Event(0, 0, 0, 0));
    for (int i = 0; i < (int)v.size(); ++i) {
      if (AXIS == 0) {
        ev.emplace_back(0, (int64_t)v[i].queryStartPos, 1, i);
        ev.emplace_back(0, fl_query_end(v[i]), 2, i);
      } else {
        ev.emplace_back((int32_t)v[i].refSeqId, (int64_t)v[i].refStartPos, 1, i);
        int32_t s = (int32_t)v[i].refSeqId;
        int64_t p = fl_ref_end(v[i]);
        if (p == ref_len[s] - 1) { s += 1; p = 0; } else p += 1; /* refPosDoPlusOne, filter.hpp:454-466 */
        ev.emplace_back(s, p, 2, i);
      }
    }
    std::sort(ev.begin(), ev.end());
    for (size_t a = 0; a < ev.size();) {
      size_t b = a;
      while (b < ev.size() && std::get<0>(ev[b]) == std::get<0>(ev[a]) && std::get<1>(ev[b]) == std::get<1>(ev[a])) ++b;
      for (size_t e = a; e < b; ++e) {
        if (std::get<2>(ev[e]) == 1) status.insert(std::get<3>(ev[e]));
        else status.erase(std::get<3>(ev[e]));
      }
      mark_good(status, secondary, drop_rand, overlap_threshold);
      a = b;
    }
    v.erase(std::remove_if(v.begin(), v.end(), [](const wfb_mapping_t& e) { return fl_discard(e) || (AXIS == 0 && fl_overlapped(e)); }), v.end());
  }
};

struct Ctx {
  const wfb_filter_params_t& P;
  const int32_t* ref_group;
  const int64_t* ref_len;
};

/* MappingFilterUtils::filterByGroup (mappingFilter.hpp:220-293) */
void filter_by_group(const Ctx& C, MapVec& in, MapVec& out, int n_mappings, bool filter_ref, double overlap_threshold) {
  out.reserve(in.size());
  std::sort(in.begin(), in.end(), [](const wfb_mapping_t& a, const wfb_mapping_t& b) { return std::tie(a.refSeqId, a.refStartPos) < std::tie(b.refSeqId, b.refStartPos); });
  if (C.P.filter_mode == WFB_FILTER_MAP || C.P.filter_mode == WFB_FILTER_ONETOONE) {
    size_t g0 = 0, g1 = 0;
    MapVec tmp;
    while (g1 != in.size()) {
      if (C.P.skip_prefix) {
        const int32_t grp = C.ref_group[in[g0].refSeqId];
        g1 = g0;
        while (g1 < in.size() && C.ref_group[in[g1].refSeqId] == grp) ++g1;
      } else g1 = in.size();
      tmp.assign(in.begin() + (ptrdiff_t)g0, in.begin() + (ptrdiff_t)g1);
      std::sort(tmp.begin(), tmp.end(), [](const wfb_mapping_t& a, const wfb_mapping_t& b) {
        return std::tie(a.queryStartPos, a.refSeqId, a.refStartPos) < std::tie(b.queryStartPos, b.refSeqId, b.refStartPos);
      });
      if (filter_ref) Sweep<1>(tmp, C.ref_len).run((int)(uint16_t)n_mappings, C.P.drop_rand != 0, overlap_threshold); /* uint16_t parameter, filter.hpp:474 */
      else Sweep<0>(tmp, C.ref_len).run(n_mappings, C.P.drop_rand != 0, overlap_threshold);
      out.insert(out.end(), tmp.begin(), tmp.end());
      g0 = g1;
    }
  }
  std::sort(out.begin(), out.end(), [](const wfb_mapping_t& a, const wfb_mapping_t& b) {
    const int16_t as = fl_strand(a), bs = fl_strand(b);
    return std::tie(a.queryStartPos, a.refSeqId, a.refStartPos, as) < std::tie(b.queryStartPos, b.refSeqId, b.refStartPos, bs);
  });
}

/* filterWeakMappings (mappingFilter.hpp:154-179) */
void filter_weak(const Ctx& C, MapVec& v, int64_t min_count, int64_t query_len) {
  const int64_t w = C.P.window_length, bl = C.P.block_length;
  v.erase(std::remove_if(v.begin(), v.end(), [&](const wfb_mapping_t& e) {
            const bool boundary = (int64_t)e.queryStartPos < w || fl_query_end(e) > query_len - w || (int64_t)e.refStartPos < w ||
                                  fl_ref_end(e) > C.ref_len[e.refSeqId] - w;
            if (boundary) return (int64_t)e.blockLength < bl / 2 || (int64_t)e.n_merged < min_count / 2;
            return (int64_t)e.blockLength < bl || (int64_t)e.n_merged < min_count;
          }), v.end());
}

/* filterFalseHighIdentity (mappingFilter.hpp:184-198): both spans are blockLength in the compact struct, so delta = 0 and
 * the bound is 1 (or NaN for an empty block, which compares false): kept as the reference evaluates it */
void filter_length_mismatch(const Ctx& C, MapVec& v) {
  const double bound = std::min(0.7, std::pow((double)C.P.percentage_identity, 3.0));
  v.erase(std::remove_if(v.begin(), v.end(), [&](const wfb_mapping_t& e) {
            const int64_t q_l = fl_query_end(e) - (int64_t)e.queryStartPos, r_l = fl_ref_end(e) - (int64_t)e.refStartPos;
            const uint64_t delta = (uint64_t)std::llabs(r_l - q_l);
            return (1.0 - (double)delta / (((double)q_l + r_l) / 2)) < bound;
          }), v.end());
}

void sparsify(const Ctx& C, MapVec& v) { /* mappingFilter.hpp:203-215 */
  if (C.P.sparsity_hash_threshold == std::numeric_limits<uint64_t>::max()) return;
  v.erase(std::remove_if(v.begin(), v.end(), [&](const wfb_mapping_t& e) { return fl_hash(e) > C.P.sparsity_hash_threshold; }), v.end());
}

inline float sc_centre_q(const wfb_mapping_t& m) { return m.queryStartPos + m.blockLength * 0.5f; }
inline float sc_centre_r(const wfb_mapping_t& m) { return m.refStartPos + m.blockLength * 0.5f; }

/* filterByScaffolds (mappingFilter.hpp:831-1016). The reference finds the nearest anchor with a 2-d tree; the distance of
 * the nearest neighbour does not depend on the search structure, so anchors are sorted by x here and the scan leaves a
 * side as soon as |dx| alone reaches the best distance (the same float expression per candidate). */
void filter_by_scaffolds(const Ctx& C, MapVec& v) {
  if (C.P.scaffold_gap <= 0) return;
  MapVec work = v;
  const MapVec original = v;
  wfb_chain_params_t cp;
  cp.split = C.P.split; cp.reserved_ = 0; cp.chain_gap = C.P.scaffold_gap; cp.window_length = C.P.window_length; cp.max_mapping_length = C.P.max_mapping_length;
  MapVec chains;
  std::vector<wfb_chain_info_t> unused;
  wfb_chain_one_query_(cp, work.data(), (int64_t)work.size(), chains, unused); /* mergeMappingsInRange, :575-733 */
  chains.erase(std::remove_if(chains.begin(), chains.end(), [&](const wfb_mapping_t& m) { return (int64_t)m.blockLength < C.P.scaffold_min_length; }), chains.end());
  if (!chains.empty() && (C.P.filter_mode == WFB_FILTER_MAP || C.P.filter_mode == WFB_FILTER_ONETOONE)) {
    MapVec kept;
    filter_by_group(C, chains, kept, (int)(C.P.num_mappings_for_scaffold - 1), false, C.P.scaffold_overlap_threshold);
    chains.swap(kept);
  }
  std::vector<std::pair<float, float>> anchors;
  for (const auto& ch : chains)
    for (const auto& o : original)
      if (o.refSeqId == ch.refSeqId && fl_strand(o) == fl_strand(ch) && o.queryStartPos >= ch.queryStartPos && fl_query_end(o) <= fl_query_end(ch) &&
          o.refStartPos >= ch.refStartPos && fl_ref_end(o) <= fl_ref_end(ch))
        anchors.emplace_back(sc_centre_q(o), sc_centre_r(o));
  if (v.empty()) return;
  if (anchors.empty()) { v.clear(); return; }
  std::sort(anchors.begin(), anchors.end());
  const float max_dist = static_cast<float>(C.P.scaffold_max_deviation);
  MapVec keep;
  for (const auto& m : v) {
    const float x = sc_centre_q(m), y = sc_centre_r(m);
    float best = std::numeric_limits<float>::infinity();
    const size_t mid = (size_t)(std::lower_bound(anchors.begin(), anchors.end(), std::make_pair(x, -std::numeric_limits<float>::infinity())) - anchors.begin());
    auto visit = [&](const std::pair<float, float>& a) {
      const float d = std::sqrt((a.first - x) * (a.first - x) + (a.second - y) * (a.second - y));
      if (d < best) best = d;
    };
    for (size_t k = mid; k < anchors.size() && !(std::abs(anchors[k].first - x) >= best); ++k) visit(anchors[k]);
    for (size_t k = mid; k-- > 0 && !(std::abs(anchors[k].first - x) >= best);) visit(anchors[k]);
    if (best <= max_dist) keep.push_back(m);
  }
  v.swap(keep);
}

/* Map::filterSubsetMappings (computeMap.hpp:1076-1165) for one query */
void filter_one_query(const Ctx& C, const wfb_mapping_t* in, int64_t n, int64_t query_len, MapVec& out, std::vector<wfb_chain_info_t>& out_chain) {
  out.clear();
  out_chain.clear();
  if (n == 0) return;
  MapVec mappings(in, in + n), merged;
  std::vector<wfb_chain_info_t> info;
  wfb_chain_params_t cp;
  cp.split = C.P.split; cp.reserved_ = 0; cp.chain_gap = C.P.chain_gap; cp.window_length = C.P.window_length; cp.max_mapping_length = C.P.max_mapping_length;
  wfb_chain_one_query_(cp, mappings.data(), n, merged, info);
  if (C.P.merge_mappings && C.P.split) {
    filter_weak(C, merged, (int64_t)std::floor((double)(C.P.block_length / C.P.window_length)), query_len);
    if (C.P.filter_mode == WFB_FILTER_MAP || C.P.filter_mode == WFB_FILTER_ONETOONE) {
      MapVec g;
      filter_by_group(C, merged, g, (int)(C.P.num_mappings_for_segment - 1), false, C.P.overlap_threshold);
      merged.swap(g);
    }
    if (C.P.filter_length_mismatches) filter_length_mismatch(C, merged);
    sparsify(C, merged);
    filter_by_scaffolds(C, merged);
    out.swap(merged);
    out_chain.assign(info.begin(), info.begin() + (ptrdiff_t)std::min(info.size(), out.size()));
    out_chain.resize(out.size(), wfb_chain_info_t{0, 0, 0});
  } else {
    if (C.P.filter_mode == WFB_FILTER_MAP || C.P.filter_mode == WFB_FILTER_ONETOONE) {
      MapVec g;
      filter_by_group(C, mappings, g, (int)(C.P.num_mappings_for_segment - 1), false, C.P.overlap_threshold);
      mappings.swap(g);
    }
    filter_by_scaffolds(C, mappings);
    out.swap(mappings);
    out_chain.resize(out.size());
    for (size_t i = 0; i < out.size(); ++i) out_chain[i] = wfb_chain_info_t{(uint32_t)i, 1, 1};
  }
}

bool params_ok(const wfb_filter_params_t* p) {
  return p && p->window_length > 0 && p->max_mapping_length > 0 && p->filter_mode >= WFB_FILTER_MAP && p->filter_mode <= WFB_FILTER_NONE;
}

void append_g(std::string& s, double x) { /* operator<<(float/double) with the default precision 6 and no format flags */
  char b[64];
  snprintf(b, sizeof b, "%g", x);
  s += b;
}

}  // namespace

extern "C" int wfb_filter_mappings_batch(const wfb_filter_params_t* params, const wfb_mapping_t* mappings, const int64_t* query_offset,
                                         const int64_t* query_len, int32_t n_queries, const int32_t* ref_group, const int64_t* ref_seq_len,
                                         wfb_mapping_t* out, wfb_chain_info_t* out_chain, int64_t out_cap, int64_t* out_offset, int32_t host_threads) {
  if (!params_ok(params) || n_queries < 0 || !ref_seq_len || (params->skip_prefix && !ref_group) ||
      (n_queries > 0 && (!query_offset || !query_len || !out || !out_chain || !out_offset)) ||
      (n_queries > 0 && query_offset && query_offset[n_queries] > query_offset[0] && !mappings)) { /* queries without any mapping need no array */
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  for (int32_t q = 0; q < n_queries; ++q)
    if (query_offset[q + 1] < query_offset[q] || query_offset[q + 1] - query_offset[q] > 0x7FFFFFFFLL) { wfb_set_last_error_("bad query_offset"); return WFB_EINVAL; }
  const Ctx C{*params, ref_group, ref_seq_len};
  std::vector<MapVec> res((size_t)n_queries);
  std::vector<std::vector<wfb_chain_info_t>> inf((size_t)n_queries);
  const int nt = std::max(1, std::min<int>(host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency(), n_queries));
  auto work = [&](int t) {
    for (int32_t q = t; q < n_queries; q += nt)
      filter_one_query(C, mappings + query_offset[q], query_offset[q + 1] - query_offset[q], query_len[q], res[(size_t)q], inf[(size_t)q]);
  };
  if (nt == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  int64_t tot = 0;
  if (n_queries > 0) out_offset[0] = 0;
  for (int32_t q = 0; q < n_queries; ++q) { tot += (int64_t)res[(size_t)q].size(); out_offset[q + 1] = tot; }
  if (tot > out_cap) { wfb_set_last_error_("filtered-mapping buffer too small"); return WFB_ECAP; }
  for (int32_t q = 0; q < n_queries; ++q) {
    std::copy(res[(size_t)q].begin(), res[(size_t)q].end(), out + out_offset[q]);
    std::copy(inf[(size_t)q].begin(), inf[(size_t)q].end(), out_chain + out_offset[q]);
  }
  return WFB_OK;
}

extern "C" int64_t wfb_filter_by_group(const wfb_filter_params_t* params, wfb_mapping_t* mappings, int64_t n, int32_t n_mappings, int32_t filter_ref,
                                       const int32_t* ref_group, const int64_t* ref_seq_len, wfb_mapping_t* out, int64_t out_cap) {
  if (!params_ok(params) || n < 0 || !ref_seq_len || (params->skip_prefix && !ref_group) || (n > 0 && (!mappings || !out))) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  const Ctx C{*params, ref_group, ref_seq_len};
  MapVec in(mappings, mappings + n), res;
  filter_by_group(C, in, res, n_mappings, filter_ref != 0, params->overlap_threshold);
  std::copy(in.begin(), in.end(), mappings);
  if ((int64_t)res.size() > out_cap) { wfb_set_last_error_("filtered-mapping buffer too small"); return WFB_ECAP; }
  std::copy(res.begin(), res.end(), out);
  return (int64_t)res.size();
}

extern "C" int64_t wfb_one_to_one_filter(const wfb_filter_params_t* params, const wfb_mapping_t* mappings, const int64_t* query_offset, int32_t n_queries,
                                         const int32_t* ref_group, const int64_t* ref_seq_len, wfb_mapping_t* out, int32_t* out_query, int64_t out_cap) {
  if (!params_ok(params) || n_queries < 0 || !ref_seq_len || (params->skip_prefix && !ref_group) ||
      (n_queries > 0 && (!mappings || !query_offset || !out || !out_query))) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  const Ctx C{*params, ref_group, ref_seq_len};
  const int64_t total = n_queries ? query_offset[n_queries] : 0;
  /* targetMappings[refSeqId] in query order (computeMap.hpp:803-808) */
  std::vector<int64_t> order((size_t)total);
  for (int64_t i = 0; i < total; ++i) order[(size_t)i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return mappings[a].refSeqId < mappings[b].refSeqId; });
  std::vector<MapVec> per_query((size_t)n_queries);
  for (size_t a = 0; a < order.size();) {
    size_t b = a;
    MapVec tgt, kept;
    while (b < order.size() && mappings[order[b]].refSeqId == mappings[order[a]].refSeqId) tgt.push_back(mappings[order[b++]]);
    filter_by_group(C, tgt, kept, (int)(params->num_mappings_for_segment - 1), true, params->overlap_threshold);
    for (const auto& m : kept) /* :820-833: every query holding the same (refSeqId, refStartPos, queryStartPos) receives a copy */
      for (int32_t q = 0; q < n_queries; ++q)
        for (int64_t i = query_offset[q]; i < query_offset[q + 1]; ++i)
          if (mappings[i].refSeqId == m.refSeqId && mappings[i].refStartPos == m.refStartPos && mappings[i].queryStartPos == m.queryStartPos) {
            per_query[(size_t)q].push_back(m);
            break;
          }
    a = b;
  }
  int64_t n = 0;
  for (auto& v : per_query) n += (int64_t)v.size();
  if (n > out_cap) { wfb_set_last_error_("one-to-one buffer too small"); return WFB_ECAP; }
  n = 0;
  for (int32_t q = 0; q < n_queries; ++q)
    for (const auto& m : per_query[(size_t)q]) { out[n] = m; out_query[n++] = q; }
  return n;
}

extern "C" int64_t wfb_mapping_paf_format(const wfb_filter_params_t* params, const wfb_mapping_t* mappings, const wfb_chain_info_t* chain, int64_t n,
                                          const char* query_name, int64_t query_len, const char* const* ref_names, const int64_t* ref_seq_len, char* buf,
                                          int64_t buf_cap, int64_t* needed) {
  if (!params || n < 0 || !query_name || !ref_names || !ref_seq_len || (n > 0 && !mappings) || (buf_cap > 0 && !buf)) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  std::vector<size_t> idx((size_t)n);
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return mappings[a].queryStartPos < mappings[b].queryStartPos; }); /* mappingOutput.hpp:86-91 */
  const char sep = params->legacy_output ? ' ' : '\t';
  std::string s;
  for (size_t i : idx) {
    const wfb_mapping_t& e = mappings[i];
    const float id = fl_identity(e);
    const float mapq = id == 1 ? 255 : std::round(-10.0 * std::log10(1 - id)); /* :99: 1 - float stays float, so this is the float overload of log10 */
    s += query_name; s += sep;
    s += std::to_string(query_len); s += sep;
    s += std::to_string(e.queryStartPos); s += sep;
    s += std::to_string(fl_query_end(e) - (params->legacy_output ? 1 : 0)); s += sep;
    s += fl_strand(e) == 1 ? "+" : "-"; s += sep;
    s += ref_names[e.refSeqId]; s += sep;
    s += std::to_string(ref_seq_len[e.refSeqId]); s += sep;
    s += std::to_string(e.refStartPos); s += sep;
    s += std::to_string(fl_ref_end(e) - (params->legacy_output ? 1 : 0));
    if (!params->legacy_output) {
      s += sep; s += std::to_string(e.conservedSketches);
      s += sep; s += std::to_string(e.blockLength);
      s += sep; append_g(s, mapq);
      s += sep; s += "id:f:"; append_g(s, id);
      s += sep; s += "kc:f:"; append_g(s, fl_complexity(e));
      if (!params->merge_mappings) { s += sep; s += "jc:f:"; append_g(s, 0.0); }
      else {
        const wfb_chain_info_t c = chain ? chain[i] : wfb_chain_info_t{(uint32_t)i, 1, 1};
        s += sep; s += "ch:Z:"; s += std::to_string(c.chainId); s += '.'; s += std::to_string(c.chainPos); s += '.'; s += std::to_string(c.chainLen);
      }
    } else { s += sep; append_g(s, e.nucIdentity * 100.0); }
    s += '\n';
  }
  if (needed) *needed = (int64_t)s.size();
  if ((int64_t)s.size() > buf_cap) { wfb_set_last_error_("mapping PAF buffer too small"); return WFB_ECAP; }
  memcpy(buf, s.data(), s.size());
  return (int64_t)s.size();
}

namespace {
struct Tok { int32_t off, len; };
std::vector<Tok> ws_tokens(const char* s, int64_t n) { /* tokenize_view, computeAlignments.hpp:62-81 */
  std::vector<Tok> t;
  int64_t p = 0;
  while (p < n) {
    while (p < n && isspace((unsigned char)s[p])) ++p;
    if (p >= n) break;
    const int64_t b = p;
    while (p < n && !isspace((unsigned char)s[p])) ++p;
    t.push_back(Tok{(int32_t)b, (int32_t)(p - b)});
  }
  return t;
}
std::vector<std::string> split_on(const std::string& s, char d) { /* split_view, computeAlignments.hpp:43-59 */
  std::vector<std::string> r;
  size_t pos = 0, f;
  while ((f = s.find(d, pos)) != std::string::npos) { r.push_back(s.substr(pos, f - pos)); pos = f + 1; }
  if (pos <= s.size()) r.push_back(s.substr(pos));
  return r;
}
bool is_a_number(const std::string& s) { /* src/common/utils.cpp:9-11 */
  return !s.empty() && s.find_first_not_of("0123456789.") == std::string::npos && std::count(s.begin(), s.end(), '.') < 2;
}
}  // namespace

extern "C" int wfb_mapping_paf_parse(const char* line, int64_t line_len, uint64_t target_padding, uint64_t query_padding, uint64_t wflign_max_len_minor,
                                     wfb_mapping_row_t* row) {
  if (!line || line_len < 0 || !row) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  try {
    const std::vector<Tok> tk = ws_tokens(line, line_len);
    if (tk.size() < 13) throw std::runtime_error("invalid mashmap mapping record (fewer than 13 tokens)");
    auto tok = [&](size_t i) { return std::string(line + tk[i].off, (size_t)tk[i].len); };
    const std::vector<std::string> idv = split_on(tok(12), ':');
    const float mm_id = !idv.empty() && is_a_number(idv.back()) ? std::stof(idv.back()) : 0.70f; /* skch::fixed::percentage_identity */
    int64_t chain_id = -1, chain_length = 1, chain_pos = 1;
    if (tk.size() > 14) {
      const std::vector<std::string> cv = split_on(tok(14), ':');
      if (cv.size() == 3 && cv[0] == "ch" && cv[1] == "Z") {
        const std::vector<std::string> parts = split_on(cv[2], '.');
        if (parts.size() == 3) { chain_id = std::stoll(parts[0]); chain_pos = std::stoll(parts[1]); chain_length = std::stoll(parts[2]); }
      }
    }
    memset(row, 0, sizeof *row);
    row->q_name_off = tk[0].off; row->q_name_len = tk[0].len;
    row->r_name_off = tk[5].off; row->r_name_len = tk[5].len;
    row->q_start = std::stoll(tok(2));
    row->q_end = std::stoll(tok(3));
    row->strand = tok(4) == "+" ? 1 : -1;
    const uint64_t ref_len = std::stoull(tok(6));
    /* MappingBoundaryRow keeps the chain fields as int32 (align_types.hpp:33-35) */
    row->chain_id = (int32_t)chain_id; row->chain_length = (int32_t)chain_length; row->chain_pos = (int32_t)chain_pos;
    uint64_t r0 = (uint64_t)std::stoll(tok(7)), r1 = (uint64_t)std::stoll(tok(8));
    uint64_t q0 = (uint64_t)row->q_start, q1 = (uint64_t)row->q_end;
    const uint64_t query_len = std::stoull(tok(1));
    if (target_padding > 0) { /* :252-264 */
      r0 = r0 >= target_padding ? r0 - target_padding : 0;
      r1 = r1 + target_padding <= ref_len ? r1 + target_padding : ref_len;
    }
    if (query_padding > 0) { /* :267-288: the padded query range is only stored for the LAST piece of a chain */
      if (chain_pos == 1) q0 = q0 >= query_padding ? q0 - query_padding : 0;
      if (chain_pos == chain_length) {
        q1 = q1 + query_padding <= query_len ? q1 + query_padding : query_len;
        row->q_start = (int64_t)q0;
        row->q_end = (int64_t)q1;
      }
    }
    if (r0 >= ref_len || r1 > ref_len) throw std::runtime_error("coordinates exceed reference length: " + std::to_string(r0) + "-" + std::to_string(r1));
    row->r_start = (int64_t)r0; row->r_end = (int64_t)r1;
    row->mashmap_estimated_identity = mm_id;
    row->query_len = (int64_t)query_len; row->ref_len = (int64_t)ref_len;
    /* createSeqRecord, :611-624: room for the head / tail patches around the target range */
    const uint64_t head = r0 >= wflign_max_len_minor ? wflign_max_len_minor : r0;
    const uint64_t tail = ref_len - r1 >= wflign_max_len_minor ? wflign_max_len_minor : ref_len - r1;
    row->ref_fetch_start = (int64_t)(r0 - head);
    row->ref_fetch_len = (int64_t)(r1 + tail - (r0 - head));
  } catch (const std::exception& e) {
    wfb_set_last_error_(std::string("parseMashmapRow: ") + e.what());
    return WFB_EINVAL;
  }
  return WFB_OK;
}
