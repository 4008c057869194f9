/*
 * TEST INFRASTRUCTURE ONLY. Compiles the reference's UNMODIFIED src/map/include/slidingMap.hpp (SlideMapper, :28-215)
 * and src/map/include/mappingCore.hpp (computeL2MappedRegions, :306-442) against a mock Sketch type, so the L2
 * restatement (map_oracle.c orc_l2_locus) and the GPU L2 kernel can be pinned against the real code.
 * winSketch.hpp (htslib) and map_stats.hpp (GSL) are not needed by the function under test: their include guards are
 * pre-defined and skch::Sketch is mocked with the two typedefs slidingMap.hpp reads from it.
 */
#include <vector>
#include <string>
#include <cstring>
#include <cstdint>
#include <unordered_map>
#include "map/include/base_types.hpp"
#include "map/include/map_parameters.hpp"
#define WIN_SKETCH_HPP
#define INDEX_ITERATOR_L2_HPP
#define MAP_STATS_HPP
namespace skch {
class Sketch {
 public:
  typedef std::vector<MinmerInfo> MI_Type;
  typedef MI_Type::const_iterator MIIter_t;
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
  int getRefGroup(skch::seqno_t) const { return 0; }
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
static_assert(sizeof(skch::MinmerInfo) == 32, "MinmerInfo layout");
}

extern "C" {
struct ref_l2_locus_t { int32_t seqId; int32_t sharedSketchSize; int64_t meanOptimalPos, optimalStart, optimalEnd; int32_t strand; int32_t pad_; };

/* A handle keeps the mock Sketch (minmerIndex copy) alive across the loci of a test. */
void* ref_l2_open(const skch::MinmerInfo* index, int64_t n) {
  MockSketch* s = new MockSketch();
  s->minmerIndex.assign(index, index + n);
  return s;
}
void ref_l2_close(void* h) { delete (MockSketch*)h; }

int ref_l2_locus(void* h, const skch::MinmerInfo* q, int q_n, int window_len, int32_t seqId, int64_t rangeStartPos,
                 int64_t rangeEndPos, ref_l2_locus_t* out, int cap) {
  MockSketch* sk = (MockSketch*)h;
  skch::Parameters param;
  param.windowLength = window_len;
  MockQ Q;
  Q.len = window_len; Q.sketchSize = q_n;
  Q.minmerTableQuery.assign(q, q + q_n);
  skch::L1_candidateLocus_t loc{seqId, rangeStartPos, rangeEndPos, 0};
  std::vector<skch::L2_mapLocus_t> v;
  skch::MappingCore<MockSketch, MockIds>::computeL2MappedRegions(Q, loc, v, sk, param);
  int n = 0;
  for (auto& l : v) {
    if (n < cap) out[n] = ref_l2_locus_t{l.seqId, l.sharedSketchSize, l.meanOptimalPos, l.optimalStart, l.optimalEnd, l.strand, 0};
    ++n;
  }
  return n;
}
}
