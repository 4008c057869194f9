/*
 * TEST INFRASTRUCTURE ONLY. Compiles the reference's UNMODIFIED src/map/include/computeMap.hpp — skch::Map, the WHOLE mapping
 * phase: SequenceIdManager, Sketch (index build), query fragments, getSeedHits / L1 / L2 per fragment, chain merge, the filters
 * and the mapping PAF writer — over oracle/shims (htslib/faidx.h + common/faigz.h: uncompressed FASTA + .fai; gsl: declarations,
 * the three functions are defined below from their textbook formulas; common/progress.hpp: silent meter). What it writes is what
 * `wfmash -m` writes for the same sequences and parameters.
 */
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "map/include/computeMap.hpp"

static double ln_choose(double n, double m) { return lgamma(n + 1) - lgamma(m + 1) - lgamma(n - m + 1); }
extern "C" double gsl_cdf_binomial_Q(unsigned int k, double p, unsigned int n) {
  if (k >= n || p <= 0) return 0.0;
  if (p >= 1) return 1.0;
  long double q = 0;
  for (unsigned i = k + 1; i <= n; ++i) q += expl((long double)ln_choose(n, i) + i * logl(p) + (n - i) * log1pl(-p));
  return (double)(q > 1 ? 1 : q);
}
extern "C" double gsl_ran_hypergeometric_pdf(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t) {
  if (t > n1 + n2) t = n1 + n2;
  if (k > n1 || k > t) return 0;
  if (t > n2 && k + n2 < t) return 0;
  return exp(ln_choose(n1, k) + ln_choose(n2, t - k) - ln_choose(n1 + n2, t));
}
extern "C" double gsl_cdf_hypergeometric_P(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t) {
  double p = 0;
  for (unsigned i = 0; i <= k; ++i) p += gsl_ran_hypergeometric_pdf(i, n1, n2, t);
  return p > 1 ? 1 : p;
}

static void write_fasta(const std::string& path, const char* const* names, const char* const* seqs, const int64_t* lens, int32_t n) {
  std::ofstream fa(path), fai(path + ".fai");
  int64_t off = 0;
  for (int32_t i = 0; i < n; ++i) {
    const std::string head = std::string(">") + names[i] + "\n";
    fa << head;
    off += (int64_t)head.size();
    fa.write(seqs[i], lens[i]);
    fa << "\n";
    fai << names[i] << "\t" << lens[i] << "\t" << off << "\t" << (lens[i] > 0 ? lens[i] : 1) << "\t" << (lens[i] > 0 ? lens[i] : 1) + 1 << "\n";
    off += lens[i] + 1;
  }
}

struct ref_map_params { /* the CLI-level knobs; everything else keeps the value src/interface/parse_args.hpp gives it by default */
  int32_t kmer_size, sketch_size, threads, filter_mode, skip_self, skip_prefix, lower_triangular, merge_mappings, split, minimum_hits;
  int64_t window_length, block_length, chain_gap, scaffold_gap, scaffold_max_deviation, scaffold_min_length;
  uint64_t max_mapping_length;
  uint32_t num_mappings_for_segment, num_mappings_for_scaffold;
  float percentage_identity;
  int32_t prefix_delim;
  double overlap_threshold, scaffold_overlap_threshold, max_kmer_freq;
};

extern "C" {
/* skch::Map(param) = the whole `wfmash -m` run (computeMap.hpp:147-227 -> mapQuery). same_file: queries == targets (one FASTA).
 * Returns the size of the mapping PAF it wrote (-1: buffer too small). */
int64_t ref_map_phase(const char* dir, const ref_map_params* P, const char* const* t_names, const char* const* t_seqs, const int64_t* t_lens, int32_t nt,
                      const char* const* q_names, const char* const* q_seqs, const int64_t* q_lens, int32_t nq, int32_t same_file, char* out,
                      int64_t out_cap) {
  const std::string d(dir), tf = d + "/t.fa", qf = same_file ? tf : d + "/q.fa", op = d + "/map.paf";
  write_fasta(tf, t_names, t_seqs, t_lens, nt);
  if (!same_file) write_fasta(qf, q_names, q_seqs, q_lens, nq);
  skch::Parameters p;
  p.kmerSize = P->kmer_size; p.windowLength = P->window_length; p.sketchSize = P->sketch_size; p.threads = P->threads;
  p.block_length = P->block_length; p.chain_gap = P->chain_gap; p.max_mapping_length = P->max_mapping_length; p.alphabetSize = 4;
  p.referenceSize = 0; p.percentageIdentity = P->percentage_identity;
  p.stage2_full_scan = true; p.stage1_topANI_filter = true; p.ANIDiff = skch::fixed::ANIDiff; p.ANIDiffConf = skch::fixed::ANIDiffConf; /* parse_args.hpp:700-725 */
  p.filterMode = P->filter_mode; p.numMappingsForSegment = P->num_mappings_for_segment; p.numMappingsForScaffold = P->num_mappings_for_scaffold;
  p.numMappingsForShortSequence = 1; p.dropRand = false;
  p.refSequences = {tf}; p.querySequences = {qf}; p.outFileName = op;
  p.split = P->split != 0; p.lower_triangular = P->lower_triangular != 0; p.skip_self = P->skip_self != 0; p.skip_prefix = P->skip_prefix != 0;
  p.prefix_delim = (char)P->prefix_delim; p.mergeMappings = P->merge_mappings != 0; p.keep_low_pct_id = true; p.report_ANI_percentage = false;
  p.filterLengthMismatches = true; p.kmerComplexityThreshold = 0; p.hgNumerator = 1.0; p.use_spaced_seeds = false; p.world_minimizers = false;
  p.sparsity_hash_threshold = std::numeric_limits<uint64_t>::max(); p.overlap_threshold = P->overlap_threshold;
  p.scaffold_overlap_threshold = P->scaffold_overlap_threshold; p.scaffold_max_deviation = P->scaffold_max_deviation; p.scaffold_gap = P->scaffold_gap;
  p.scaffold_min_length = P->scaffold_min_length; p.legacy_output = false; p.minimum_hits = P->minimum_hits; p.max_kmer_freq = P->max_kmer_freq;
  p.use_progress_bar = false; p.auto_pct_identity = false; p.ani_percentile = 50; p.ani_adjustment = -2.0f; p.use_streaming_minhash = false;
  p.create_index_only = false; p.overwrite_index = false;
  {
    skch::Map mapper(p);
  }
  std::ifstream in(op, std::ios::binary);
  std::string s((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if ((int64_t)s.size() > out_cap) return -1;
  memcpy(out, s.data(), s.size());
  return (int64_t)s.size();
}
}
