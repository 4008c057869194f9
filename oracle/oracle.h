/*
 * oracle.h — CPU restatement of the reference algorithms (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product path (wfmash_b200/csrc, libwfmash_b200.so) never links, loads or
 * calls anything under oracle/.
 *
 * Parity pinning:
 *   - wfa_oracle.c is pinned against the reference's own known-answer vectors
 *     deps/WFA2-lib/tests/wfa.utest.seq + tests/wfa.utest.check/test.biwfa.affine2p.alg
 *     (305 pairs, fixtures committed under tests/golden/) and differentially against
 *     oracle/_ref/libwfa2ref.so (the unmodified reference sources compiled in place).
 *   - map_oracle.c: the reference has no byte-exact tests for src/map ("parity unpinned" by the
 *     reference's own tests); it is pinned differentially against
 *     oracle/_ref/libmapref.so (reference commonFunc.hpp compiled unmodified) and by committed
 *     fixtures generated from it (tests/golden/make_map_golden.py).
 */
#ifndef WFMASH_B200_ORACLE_H
#define WFMASH_B200_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- path 2: gap-affine-2p WFA / biWFA ------------------------------------------------- */

typedef struct {
  int x;   /* mismatch */
  int o1;  /* gap_opening1 */
  int e1;  /* gap_extension1 */
  int o2;  /* gap_opening2 */
  int e2;  /* gap_extension2 */
} orc_penalties_t;

/* Work counters: C = wavefront cells computed, E = matched chars in extend,
 * O = diagonals tested in breakpoint overlap scans. */
typedef struct {
  int64_t cells;
  int64_t extend_matches;
  int64_t overlap_tests;
  int64_t score_steps;
} orc_wfa_counters_t;

/* biWFA end-to-end alignment (restates wavefront_bialign, wavefront_bialign.c:1266).
 * pattern = target slice, text = query slice (wflign.cpp:148).
 * ops_out receives M/X/I/D chars in alignment order (cigar->operations[begin_offset..end_offset)).
 * Returns 0 on success (WF_STATUS_ALG_COMPLETED), negative otherwise. */
int orc_biwfa_align(const char* pattern, int plen, const char* text, int tlen,
                    const orc_penalties_t* pen, char* ops_out, int ops_cap, int* ops_len,
                    int* score, orc_wfa_counters_t* counters);

/* Unidirectional full-memory WFA + backtrace (restates wavefront_unialign.c:242 +
 * wavefront_backtrace.c:320), end-to-end, components M->M. */
int orc_wfa_align(const char* pattern, int plen, const char* text, int tlen,
                  const orc_penalties_t* pen, char* ops_out, int ops_cap, int* ops_len,
                  int* score, orc_wfa_counters_t* counters);

/* Ends-free unidirectional WFA as used by wfmash's head / tail patching (wflign.cpp:280-305,368-397).
 * term_group = 1 (scalar build), 8 (AVX2) or 16 (AVX-512): which terminating cell the reference picks
 * (wavefront_extend_kernels.c:166-193 vs wavefront_extend_kernels_avx.c:296-400,592-691). */
int orc_wfa_endsfree(const char* pattern, int plen, int pbf, int pef, const char* text, int tlen, int tbf, int tef,
                     const orc_penalties_t* pen, int term_group, char* ops_out, int ops_cap, int* ops_len, int* score);

/* Gap-affine-2p score of an ops string (restates cigar_score_gap_affine2p, alignment/cigar.c);
 * returned as a non-negative penalty. */
int orc_cigar_score(const char* ops, int n, const orc_penalties_t* pen);

/* Check that ops is a valid end-to-end transcript of pattern -> text. 1 = valid. */
int orc_cigar_check(const char* pattern, int plen, const char* text, int tlen, const char* ops, int n);

/* ---- path 1: MashMap 3.5 sketching ------------------------------------------------------ */

/* MinmerInfo, base_types.hpp:79-110 (32 bytes). strand: FWD=1, AMBIG=0, REV=-1. */
typedef struct {
  uint64_t hash;
  int64_t wpos;
  int64_t wpos_end;
  int32_t seqId;
  int16_t strand;
  int16_t pad_;
} orc_minmer_t;

/* MurmurHash3_x64_128 low 64 bits, seed 42 (commonFunc.hpp:173-182, murmur3.h:226-303). */
uint64_t orc_kmer_hash(const char* kmer, int k);

/* sketchSequence (commonFunc.hpp:217-323): bottom-s distinct canonical hashes of one fragment.
 * seq must already be upper-cased/N-masked (makeUpperCaseAndValidDNA). Returns count written. */
int orc_sketch_fragment(const char* seq, int len, int k, int s, int32_t seqId, orc_minmer_t* out);

/* addMinmers (commonFunc.hpp:439-708): windowed minmer intervals of one target sequence (seq already
 * upper-cased / N-masked). Returns the number of records (may exceed cap; only cap are written). */
int64_t orc_add_minmers(const char* seq, int64_t len, int k, int w, int s, int32_t seqId, orc_minmer_t* out, int64_t cap);

/* IntervalPoint, base_types.hpp:63-76 (24 bytes). side: OPEN = 1, CLOSE = -1. */
typedef struct {
  int64_t pos;
  uint64_t hash;
  int32_t seqId;
  int8_t side;
  int8_t pad_[3];
} orc_ipoint_t;

/* L1_candidateLocus_t, mappingCore.hpp:24-30 */
typedef struct {
  int32_t seqId;
  int32_t pad_;
  int64_t rangeStartPos;
  int64_t rangeEndPos;
  int32_t intersectionSize;
  int32_t pad2_;
} orc_l1_locus_t;

uint64_t orc_count_threshold(const uint64_t* freqs, int64_t nuniq, uint64_t total_windows, double max_kmer_freq);
int64_t orc_index_build(const orc_minmer_t* mi, int64_t n, const int32_t* partition_of_seq, double max_kmer_freq,
                        orc_minmer_t* kept, orc_ipoint_t* points, int64_t* npoints, uint64_t* uhash, int64_t* ustart,
                        int64_t* ucount, int64_t* nuniq_out, uint64_t* threshold);
int orc_l1_fragment(const uint64_t* uhash, const int64_t* ustart, const int64_t* ucount, int64_t nuniq, const orc_ipoint_t* points,
                    const uint64_t* q_hashes, int q_n, int32_t q_seq_id, int q_group, const int32_t* ref_group, int skip_self,
                    int skip_prefix, int lower_triangular, int minimum_hits, int param_sketch_size, int window_len,
                    const int* cutoffs, int ncut, orc_l1_locus_t* out, int cap);

/* L2_mapLocus_t, mappingCore.hpp:33-41 */
typedef struct {
  int32_t seqId;
  int32_t sharedSketchSize;
  int64_t meanOptimalPos;
  int64_t optimalStart;
  int64_t optimalEnd;
  int32_t strand;
  int32_t pad_;
} orc_l2_locus_t;

/* the MappingResult fields doL2Mapping sets (computeMap.hpp:1029-1044) + the L2_mapLocus_t extras */
typedef struct {
  int32_t frag;
  int32_t refSeqId;
  int64_t refStartPos;
  int64_t optimalStart;
  int64_t optimalEnd;
  int32_t conservedSketches;
  int32_t strand;
  float nucIdentity;
  float kmerComplexity;
} orc_l2_mapping_t;

int orc_l2_locus(const orc_minmer_t* index, int64_t n_index, const orc_minmer_t* q, int q_n, int window_len_param, int32_t seqId,
                 int64_t rangeStartPos, int64_t rangeEndPos, orc_l2_locus_t* out, int cap, int* best_intersection, int64_t* dup_inserts);
float orc_j2md(float j, int k);
float orc_md2j(float d, int k);
int orc_stage1_pass(double hg_numerator, float ani_diff, int kmer_size, int q_sketch_size, int intersection_size);
int orc_l2_fragment(const orc_minmer_t* index, int64_t n_index, const orc_minmer_t* q, int q_n, float kmer_complexity, int kmer_size,
                    int window_len_param, const orc_l1_locus_t* loci, int n_loci, int stage1, double hg_numerator, float ani_diff,
                    int min_shared, orc_l2_mapping_t* out, int cap);

/* ANI auto-identity sketch (map_stats.hpp:563-637, streamingMinHash.hpp:89-99): see map_oracle.c */
int orc_ani_add_sequence(const char* seq, int64_t len, int k, int ssize, uint64_t* heap, int heap_n);
int orc_ani_add_hash(uint64_t h, int ssize, uint64_t* heap, int heap_n);

#ifdef __cplusplus
}
#endif
#endif
