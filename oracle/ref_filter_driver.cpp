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

/* ---- the filters between the chain merge and the mapping PAF (second part; b3) ---- */
struct ref_filter_params { /* mirrors wfb_filter_params_t field by field (include/wfmash_b200.h) */
  int32_t split, merge_mappings, filter_mode, skip_prefix, filter_length_mismatches, drop_rand, threads, legacy_output;
  int64_t chain_gap, window_length, block_length;
  uint64_t max_mapping_length, sparsity_hash_threshold;
  uint32_t num_mappings_for_segment, num_mappings_for_scaffold;
  double overlap_threshold, scaffold_overlap_threshold;
  int64_t scaffold_gap, scaffold_max_deviation, scaffold_min_length;
  float percentage_identity; int32_t reserved_;
};

static skch::Parameters to_param(const ref_filter_params* p) {
  skch::Parameters q;
  q.split = p->split; q.mergeMappings = p->merge_mappings; q.filterMode = p->filter_mode; q.skip_prefix = p->skip_prefix;
  q.filterLengthMismatches = p->filter_length_mismatches; q.dropRand = p->drop_rand; q.threads = p->threads; q.legacy_output = p->legacy_output;
  q.chain_gap = p->chain_gap; q.windowLength = p->window_length; q.block_length = p->block_length;
  q.max_mapping_length = p->max_mapping_length; q.sparsity_hash_threshold = p->sparsity_hash_threshold;
  q.numMappingsForSegment = p->num_mappings_for_segment; q.numMappingsForScaffold = p->num_mappings_for_scaffold;
  q.overlap_threshold = p->overlap_threshold; q.scaffold_overlap_threshold = p->scaffold_overlap_threshold;
  q.scaffold_gap = p->scaffold_gap; q.scaffold_max_deviation = p->scaffold_max_deviation; q.scaffold_min_length = p->scaffold_min_length;
  q.percentageIdentity = p->percentage_identity;
  return q;
}

/* The body of Map::filterSubsetMappings (computeMap.hpp:1076-1165; Map itself needs htslib + GSL + taskflow and cannot be
 * compiled here): the same calls in the same order on the reference's UNMODIFIED FilterUtils functions. Returns the number
 * of mappings the reference would print for this query (merged branch when mergeMappings && split, else the non-merged one);
 * out / out_chain receive min(count, cap) of them. */
int64_t ref_filter_subset(const ref_filter_params* fp, const skch::MappingResult* in, int64_t n, int32_t query_seq_id, int64_t query_len,
                          const int32_t* ref_groups, const int64_t* ref_lengths, skch::MappingResult* out, skch::ChainInfo* out_chain, int64_t cap) {
  using FU = skch::MappingFilterUtils;
  skch::Parameters param = to_param(fp);
  skch::SequenceIdManager ids; ids.groups = ref_groups; ids.lengths = ref_lengths;
  progress_meter::ProgressMeter progress;
  skch::MappingResultsVector_t mappings(in, in + n), resM; skch::ChainInfoVector_t resC;
  if (!mappings.empty()) {
    skch::MappingResultsVector_t rawMappings = mappings;
    auto mwc = FU::mergeMappingsInRangeWithChains(mappings, param.chain_gap, param, progress, query_seq_id, query_len);
    auto& mm = mwc.mappings;
    if (param.mergeMappings && param.split) {
      FU::filterWeakMappings(mm, std::floor(param.block_length / param.windowLength), param, ids, query_len);
      if (param.filterMode == skch::filter::MAP || param.filterMode == skch::filter::ONETOONE) {
        skch::MappingResultsVector_t g;
        FU::filterByGroup(mm, g, param.numMappingsForSegment - 1, false, ids, param, progress);
        mm = std::move(g);
      }
      if (param.filterLengthMismatches) FU::filterFalseHighIdentity(mm, param);
      FU::sparsifyMappings(mm, param);
      FU::filterByScaffolds(mm, rawMappings, param, ids, progress, query_seq_id, query_len, nullptr, nullptr, nullptr);
      resM = std::move(mm); resC = std::move(mwc.chainInfo);
    } else {
      if (param.filterMode == skch::filter::MAP || param.filterMode == skch::filter::ONETOONE) {
        skch::MappingResultsVector_t g;
        FU::filterByGroup(mappings, g, param.numMappingsForSegment - 1, false, ids, param, progress);
        mappings = std::move(g);
      }
      FU::filterByScaffolds(mappings, rawMappings, param, ids, progress, query_seq_id, query_len, nullptr, nullptr, nullptr);
      resM = std::move(mappings);
      resC.resize(resM.size());
      for (size_t i = 0; i < resM.size(); ++i) resC[i] = {static_cast<uint32_t>(i), 1, 1};
    }
  }
  for (size_t i = 0; i < resM.size() && (int64_t)i < cap; ++i) {
    out[i] = resM[i];
    /* the reference indexes chainInfo by the position in the FILTERED vector (mappingOutput.hpp:96-97); when the filters
     * removed anything the vector is longer than the mappings: only the first resM.size() entries are ever read */
    out_chain[i] = i < resC.size() ? resC[i] : skch::ChainInfo{0, 0, 0};
  }
  return (int64_t)resM.size();
}

/* FilterUtils::filterByGroup (mappingFilter.hpp:220-293) alone: filter_ref = 0 query plane sweep, 1 reference plane sweep
 * (filter.hpp:303-316 / 474-535). in is reordered like the reference's unfilteredMappings. */
int64_t ref_filter_by_group(const ref_filter_params* fp, skch::MappingResult* in, int64_t n, int32_t n_mappings, int32_t filter_ref,
                            const int32_t* ref_groups, const int64_t* ref_lengths, skch::MappingResult* out, int64_t cap) {
  skch::Parameters param = to_param(fp);
  skch::SequenceIdManager ids; ids.groups = ref_groups; ids.lengths = ref_lengths;
  progress_meter::ProgressMeter progress;
  skch::MappingResultsVector_t v(in, in + n), g;
  skch::MappingFilterUtils::filterByGroup(v, g, n_mappings, filter_ref != 0, ids, param, progress);
  for (int64_t i = 0; i < n; ++i) in[i] = v[i];
  for (size_t i = 0; i < g.size() && (int64_t)i < cap; ++i) out[i] = g[i];
  return (int64_t)g.size();
}

/* OutputHandler::reportReadMappings (mappingOutput.hpp:74-139): the mapping PAF text of one query. Sequence names are the mock
 * id manager's "s<refSeqId>". Returns the text length (buf receives min(len, cap) bytes). */
int64_t ref_report_mappings(const ref_filter_params* fp, const skch::MappingResult* in, const skch::ChainInfo* chain, int64_t n, const char* query_name,
                            int64_t query_len, const int64_t* ref_lengths, char* buf, int64_t cap) {
  skch::Parameters param = to_param(fp);
  skch::SequenceIdManager ids; ids.lengths = ref_lengths;
  skch::MappingResultsVector_t v(in, in + n);
  skch::ChainInfoVector_t c(chain, chain + n);
  std::ostringstream os;
  skch::MappingOutput::reportReadMappings(v, c, std::string(query_name), os, ids, param, nullptr, query_len);
  const std::string s = os.str();
  memcpy(buf, s.data(), std::min<size_t>(s.size(), (size_t)cap));
  return (int64_t)s.size();
}
}
