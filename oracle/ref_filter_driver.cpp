/*
 * TEST INFRASTRUCTURE ONLY. Compiles the reference's UNMODIFIED src/map/include/mappingFilter.hpp and filter.hpp
 * (chain merge + split, weak-mapping / group plane-sweep / length-mismatch / scaffold filters) against a mock
 * SequenceIdManager (sequenceIds.hpp reads FASTA index files) and the silent progress-meter shim
 * (oracle/shims/common/progress.hpp), so the product's host restatement (wfmash_b200/csrc/chain_host.cu) can be pinned
 * against the real code.
 */
#include <vector>
#include <string>
#include <cstring>
#include <cstdint>
#include <cassert> /* the reference's headers rely on these coming in through its other includes */
#include <map>
#include "map/include/base_types.hpp"
#include "map/include/map_parameters.hpp"
#define SEQUENCE_ID_MANAGER_HPP
namespace skch {
class SequenceIdManager {
 public:
  const int32_t* groups = nullptr;
  const int64_t* lengths = nullptr;
  int getRefGroup(seqno_t id) const { return groups ? groups[id] : 0; }
  offset_t getSequenceLength(seqno_t id) const { return lengths[id]; }
  std::string getSequenceName(seqno_t id) const { return "s" + std::to_string(id); }
};
}
#include "map/include/mappingFilter.hpp"
#include "map/include/mappingOutput.hpp"

static_assert(sizeof(skch::MappingResult) == 28, "MappingResult layout");
static_assert(sizeof(skch::ChainInfo) == 8, "ChainInfo layout");

extern "C" {
/* FilterUtils::mergeMappingsInRangeWithChains (mappingFilter.hpp:381-571). mappings is reordered in place like the
 * reference's readMappings. Returns the number of merged mappings (may exceed cap; only cap are written). */
int64_t ref_merge_with_chains(skch::MappingResult* mappings, int64_t n, int32_t split, int64_t chain_gap, int64_t window_length,
                              uint64_t max_mapping_length, int32_t query_seq_id, int64_t query_len, skch::MappingResult* merged,
                              skch::ChainInfo* chain_info, int64_t cap) {
  skch::Parameters param;
  param.split = split; param.chain_gap = chain_gap; param.windowLength = window_length; param.max_mapping_length = max_mapping_length;
  progress_meter::ProgressMeter pm;
  std::vector<skch::MappingResult> v(mappings, mappings + n);
  auto r = skch::MappingFilterUtils::mergeMappingsInRangeWithChains(v, (int)param.chain_gap, param, pm, query_seq_id, query_len);
  for (int64_t i = 0; i < n; ++i) mappings[i] = v[i];
  int64_t m = 0;
  for (size_t i = 0; i < r.mappings.size(); ++i) {
    if (m < cap) { merged[m] = r.mappings[i]; chain_info[m] = r.chainInfo[i]; }
    ++m;
  }
  return m;
}

/* OutputHandler::mappingBoundarySanityCheck (mappingOutput.hpp:31-69) */
void ref_boundary_sanity(skch::MappingResult* mappings, int64_t n, int64_t query_len, const int64_t* seq_lengths) {
  progress_meter::ProgressMeter pm;
  skch::InputSeqProgContainer in(std::string((size_t)query_len, 'N'), "q", 0, pm);
  skch::SequenceIdManager ids;
  ids.lengths = seq_lengths;
  std::vector<skch::MappingResult> v(mappings, mappings + n);
  skch::MappingOutput::mappingBoundarySanityCheck(&in, v, ids);
  for (int64_t i = 0; i < n; ++i) mappings[i] = v[i];
}

/* the MappingResult setters doL2Mapping uses (computeMap.hpp:1029-1044) */
void ref_make_mapping(int32_t refSeqId, int64_t meanOptimalPos, int64_t q_len, int32_t sharedSketchSize, float nucIdentity, float kmerComplexity,
                      int32_t strand, skch::MappingResult* out) {
  skch::MappingResult res;
  res.refSeqId = refSeqId; res.refStartPos = meanOptimalPos; res.queryStartPos = 0; res.blockLength = q_len;
  res.conservedSketches = sharedSketchSize; res.n_merged = 1;
  res.setNucIdentity(nucIdentity); res.setKmerComplexity(kmerComplexity); res.setStrand(strand);
  res.setDiscard(false); res.setOverlapped(false);
  *out = res;
}
}
