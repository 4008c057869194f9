/*
 * TEST INFRASTRUCTURE ONLY. Compiles the reference's UNMODIFIED src/map/include/mappingCore.hpp
 * (getSeedIntervalPoints :81-131, computeL1CandidateRegions :136-301) against mock Sketch / id-manager
 * types, so the L1 restatement (map_oracle.c) and the GPU L1 kernel can be pinned against the real code.
 * The headers mappingCore.hpp pulls in for its L2 half (winSketch.hpp -> htslib, map_stats.hpp -> GSL)
 * are not needed by the two functions under test; their include guards are pre-defined here and the few
 * names the (uninstantiated) L2 templates mention are forward-declared.
 */
#include <vector>
#include <string>
#include <cstring>
#include <cstdint>
#include <unordered_map>
#include "map/include/base_types.hpp"
#include "map/include/map_parameters.hpp"
#define WIN_SKETCH_HPP
#define SLIDING_MAP_HPP
#define INDEX_ITERATOR_L2_HPP
#define MAP_STATS_HPP
namespace skch {
template <typename Q> class SlideMapper {
 public:
  SlideMapper(Q&) {}
  template <typename... A> void insert_minmer(A...) {}
  template <typename... A> void delete_minmer(A...) {}
  int sharedSketchElements = 0, intersectionSize = 0, strand_votes = 0;
  int pivRank = 0;
};
namespace Stat {
inline float j2md(float, int) { return 0; }
inline float md_lower_bound(float, int, int, float) { return 0; }
}
}
#include "map/include/mappingCore.hpp"

namespace {
struct MockSketch {
  using MI_Type = std::vector<skch::MinmerInfo>;
  using MIIter_t = MI_Type::const_iterator;
  std::unordered_map<skch::hash_t, std::vector<skch::IntervalPoint>> minmerPosLookupIndex;
  MI_Type minmerIndex;
};
struct MockIds {
  const int32_t* groups;
  int getRefGroup(skch::seqno_t id) const { return groups[id]; }
};
struct MockQ {
  char* seq = nullptr;
  skch::seqno_t seqId = 0;
  skch::offset_t len = 0;
  int sketchSize = 0;
  float kmerComplexity = 1;
  int refGroup = 0;
  std::vector<skch::MinmerInfo> minmerTableQuery;
};
}

extern "C" {
struct ref_ipoint_t { int64_t pos; uint64_t hash; int32_t seqId; int8_t side; int8_t pad_[3]; };
struct ref_l1_locus_t { int32_t seqId; int32_t pad_; int64_t rangeStartPos; int64_t rangeEndPos; int32_t intersectionSize; int32_t pad2_; };

/* points grouped by hash (uhash/ustart/ucount) exactly as the index holds them; query = its sketch hashes. */
int ref_l1_fragment(const uint64_t* uhash, const int64_t* ustart, const int64_t* ucount, int64_t nuniq, const ref_ipoint_t* points,
                    const uint64_t* q_hashes, int q_n, int32_t q_seq_id, int q_group, const int32_t* ref_group, int skip_self,
                    int skip_prefix, int lower_triangular, int minimum_hits, int param_sketch_size, int window_len, const int* cutoffs,
                    int ncut, ref_l1_locus_t* out, int cap) {
  (void)q_group;
  MockSketch sk;
  for (int64_t u = 0; u < nuniq; ++u) {
    auto& v = sk.minmerPosLookupIndex[uhash[u]];
    for (int64_t t = 0; t < ucount[u]; ++t) {
      const ref_ipoint_t& p = points[ustart[u] + t];
      v.push_back(skch::IntervalPoint{p.pos, p.hash, p.seqId, (skch::side_t)p.side});
    }
  }
  MockIds ids{ref_group};
  skch::Parameters param;
  param.skip_self = skip_self; param.skip_prefix = skip_prefix; param.lower_triangular = lower_triangular;
  param.windowLength = window_len; param.sketchSize = param_sketch_size;
  param.stage1_topANI_filter = true; param.stage2_full_scan = true; /* parse_args.hpp:701-702 */
  MockQ Q;
  Q.seqId = q_seq_id; Q.len = window_len; Q.sketchSize = q_n;
  for (int i = 0; i < q_n; ++i) Q.minmerTableQuery.push_back(skch::MinmerInfo{q_hashes[i], 0, 0, q_seq_id, 1});
  using Core = skch::MappingCore<MockSketch, MockIds>;
  std::vector<skch::IntervalPoint> ip;
  Core::getSeedIntervalPoints(Q, ip, &sk, ids, param);
  std::vector<int> cut(cutoffs, cutoffs + ncut);
  std::vector<skch::L1_candidateLocus_t> l1;
  auto b = ip.begin();
  auto e = ip.begin();
  while (e != ip.end()) { /* Map::doL1Mapping's group slicing, computeMap.hpp:964-982 */
    if (param.skip_prefix) {
      const int g = ids.getRefGroup(b->seqId);
      e = std::find_if_not(b, ip.end(), [&](const skch::IntervalPoint& p) { return g == ids.getRefGroup(p.seqId); });
    } else e = ip.end();
    Core::computeL1CandidateRegions(Q, b, e, minimum_hits, param, cut, l1);
    b = e;
  }
  int n = 0;
  for (auto& l : l1) {
    if (n < cap) out[n] = ref_l1_locus_t{l.seqId, 0, l.rangeStartPos, l.rangeEndPos, l.intersectionSize, 0};
    ++n;
  }
  return n;
}
}
