/*
 * wfmash_b200.h — C ABI of libwfmash_b200.so, the B200-native (sm_100a) replacement for wfmash's two
 * data-parallel hot paths. Plain pointers and sizes only; no torch / CUDA types in the signatures.
 * Every entry point names the reference interface it replaces (paths relative to the wfmash tree).
 *
 * All functions return 0 on success or a negative WFB_E* code; they never fall back to a CPU
 * implementation: without a CUDA device they fail with WFB_ENODEV.
 */
#ifndef WFMASH_B200_H
#define WFMASH_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFB_OK 0
#define WFB_ENODEV (-1)   /* no CUDA device / driver */
#define WFB_ECUDA (-2)    /* CUDA runtime error (see wfb_last_error) */
#define WFB_EINVAL (-3)   /* bad argument */
#define WFB_ENOMEM (-4)   /* device or host allocation failed */
#define WFB_ECAP (-5)     /* output capacity too small */

/* Human-readable description of the last error on this thread. */
const char* wfb_last_error(void);
/* Library / build info: "wfmash_b200 <version> sm_100a". */
const char* wfb_version(void);
/* Number of visible CUDA devices (0 when there is none; never negative). */
int wfb_device_count(void);
/* Number of kernels launched by this library since load (bench.py's gpu_launches counter). */
uint64_t wfb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Path 2 — biWFA base-level aligner.
 * Replaces wflign::wavefront::do_biwfa_alignment's call
 *   wfa::WFAlignerGapAffine2Pieces(0,x,o1,e1,o2,e2,Alignment,MemoryUltralow).alignEnd2End(target,query)
 * (src/common/wflign/src/wflign.cpp:136-148; deps/WFA2-lib/bindings/cpp/WFAligner.cpp:68-78,449-468;
 *  C API wavefront_align, deps/WFA2-lib/wavefront/wavefront_align.c:212) for a BATCH of mapping
 * records. The caller (the align::Aligner batch builder, src/align/include/computeAlignments.hpp:661-720)
 * owns every buffer.
 * ---------------------------------------------------------------------------------------------- */

/* wflign_penalties_t (src/common/wflign/src/wflign.hpp) / affine2p_penalties_t; match is always 0. */
typedef struct {
  int32_t mismatch;
  int32_t gap_opening1;
  int32_t gap_extension1;
  int32_t gap_opening2;
  int32_t gap_extension2;
} wfb_penalties_t;

/* One alignment problem: pattern = target slice, text = (strand-corrected) query slice, ASCII,
 * upper-cased / N-masked by the caller exactly as processAlignment does (computeAlignments.hpp:671-681). */
typedef struct {
  const char* pattern;
  int32_t pattern_len;
  const char* text;
  int32_t text_len;
} wfb_pair_t;

/* Per-pair result. ops are the M/X/I/D characters of cigar_t::operations[begin_offset,end_offset)
 * (deps/WFA2-lib/alignment/cigar.h), i.e. what wflign_edit_cigar_copy consumes
 * (src/common/wflign/src/wflign_alignment.cpp:665-678). status mirrors wavefront_align's return:
 * 0 = WF_STATUS_ALG_COMPLETED, negative = unattainable (the reference then writes no record,
 * wflign.cpp:150-152). score is the reference's cigar score convention: -(gap-affine-2p penalty). */
typedef struct {
  int32_t status;
  int32_t score;
  int64_t ops_offset; /* into the ops buffer passed to wfb_align_batch */
  int32_t ops_len;
  int32_t reserved_;
} wfb_aln_result_t;

/* Work counters summed over the batch (C cells, E extended matches, O overlap tests),
 * split by kernel so that the roofline of the dominant (breakpoint) kernel can be computed. */
typedef struct {
  uint64_t cells;           /* breakpoint kernel: wavefront cells computed                */
  uint64_t extend_matches;  /* breakpoint kernel: matched bases in extend                 */
  uint64_t overlap_tests;   /* breakpoint kernel: diagonals scanned by overlap detection  */
  uint64_t score_steps;     /* breakpoint kernel: score steps (both directions)           */
  uint64_t break_tasks;
  uint64_t base_tasks;
  uint64_t base_cells;          /* base kernel share */
  uint64_t base_extend_matches;
  uint64_t base_score_steps;
  uint64_t levels;
  double kernel_ms;      /* CUDA-event time of all kernels of the call (on the call's stream)   */
  double break_kernel_ms; /* ... of the breakpoint (bidirectional wavefront) kernel launches only */
  double patch_kernel_ms; /* wfb_biwfa_paf_batch only: CUDA-event time of the head / tail patch kernel rounds */
  uint64_t h2d_bytes, d2h_bytes; /* bytes the call copied to / from the device                   */
  uint64_t patch_cap_kept_main;  /* wfb_biwfa_paf_batch only: head / tail patches whose ends-free alignment hit a device capacity; those ends
                                    keep the main CIGAR, as the reference does when a patch aligner fails (wflign.cpp:299,385) */
  uint64_t main_device_cap;      /* wfb_biwfa_paf_batch only: records whose MAIN alignment hit a device capacity (task queue, base-case score
                                    cap) — no reference counterpart; they are reported WFB_REC_UNALIGNED and counted here */
} wfb_align_stats_t;

typedef struct wfb_aligner wfb_aligner_t;

/* Create / destroy an aligner bound to one device. workspace_bytes = 0 picks a default
 * (a fraction of free HBM); the workspace bounds how many alignments are resident at once. */
wfb_aligner_t* wfb_aligner_create(int device, const wfb_penalties_t* penalties, uint64_t workspace_bytes);
void wfb_aligner_destroy(wfb_aligner_t*);

/* Align n pairs end-to-end with HOST buffers (copies H2D / D2H inside).
 *   ops / ops_cap : caller buffer receiving all operation strings back to back; needs
 *                   sum(pattern_len+text_len) bytes in the worst case (WFB_ECAP otherwise).
 *   results[n]    : per pair.
 *   stats         : optional. */
int wfb_align_batch(wfb_aligner_t*, const wfb_pair_t* pairs, int32_t n, char* ops, int64_t ops_cap,
                    wfb_aln_result_t* results, wfb_align_stats_t* stats);

/* wfb_align_batch with a scheduling hint: cost_hint[i] ~ the expected number of edits of pair i (e.g. (1 - mapping identity) *
 * length; any monotone proxy of the alignment score). The persistent kernel starts the most expensive pairs first, so the launch
 * does not end with a few CTAs finishing heavy pairs alone. Results are independent of the hint. NULL = order by length. */
int wfb_align_batch_hinted(wfb_aligner_t*, const wfb_pair_t* pairs, int32_t n, const float* cost_hint, char* ops, int64_t ops_cap,
                           wfb_aln_result_t* results, wfb_align_stats_t* stats);

/* Same with the sequences already resident in device memory (used by bench.py's HBM-resident
 * `value` leg). d_seq holds all sequences; pattern_off/text_off are byte offsets into it.
 * The results (ops + wfb_aln_result_t) still come back to the host buffers. */
int wfb_align_batch_device(wfb_aligner_t*, const char* d_seq, const int64_t* pattern_off, const int32_t* pattern_len,
                           const int64_t* text_off, const int32_t* text_len, int32_t n, char* ops, int64_t ops_cap,
                           wfb_aln_result_t* results, wfb_align_stats_t* stats);

/* Compatibility shim with the shape of the C API that wfb_align_batch sits under:
 *   wavefront_aligner_t* wavefront_aligner_new(attr) / int wavefront_align(aligner, pattern, plen, text, tlen)
 * (deps/WFA2-lib/wavefront/wavefront_aligner.c:422, wavefront_align.c:212) with the result the reference leaves in
 * aligner->cigar (deps/WFA2-lib/alignment/cigar.h: operations[begin_offset, end_offset), score). One pair per call, so
 * that align_benchmark-style drivers can target the GPU; throughput needs wfb_align_batch. Returns the reference's
 * status convention: 0 = WF_STATUS_ALG_COMPLETED, -300 = WF_STATUS_UNATTAINABLE (deps/WFA2-lib/wavefront/wfa.h:46-51),
 * or a negative WFB_E* code for an API failure (distinguishable: WFB_E* are > -100). */
#define WFB_WF_STATUS_ALG_COMPLETED 0
#define WFB_WF_STATUS_UNATTAINABLE (-300)
int wfb_wavefront_align(wfb_aligner_t*, const char* pattern, int32_t pattern_length, const char* text, int32_t text_length,
                        char* cigar_operations, int32_t cigar_cap, int32_t* cigar_length, int32_t* cigar_score);

/* Head / tail patch alignments of do_biwfa_alignment (src/common/wflign/src/wflign.cpp:280-305, 368-397):
 *   wfa::WFAlignerGapAffine2Pieces(0,x,o1,e1,o2,e2,Alignment,MemoryMed).alignEndsFree(pattern,
 *       patternBeginFree, patternEndFree, text, textBeginFree, textEndFree)
 * (deps/WFA2-lib/bindings/cpp/WFAligner.cpp:110-135) for a batch. Results as for wfb_align_batch.
 * term_group: the reference picks the terminating cell in an order that depends on how it was compiled
 * (scalar = 1, AVX2 = 8, AVX-512 = 16 lanes; wavefront_extend_kernels.c:166-193,
 * wavefront_extend_kernels_avx.c:296-400,592-691); pass the value matching the reference build to compare with. */
typedef struct {
  const char* pattern;
  int32_t pattern_len;
  const char* text;
  int32_t text_len;
  int32_t pattern_begin_free, pattern_end_free, text_begin_free, text_end_free;
} wfb_endsfree_pair_t;

int wfb_align_endsfree_batch(wfb_aligner_t*, const wfb_endsfree_pair_t* pairs, int32_t n, int32_t term_group, char* ops,
                             int64_t ops_cap, wfb_aln_result_t* results);

/* Whole-record replacement of wflign::wavefront::do_biwfa_alignment (src/common/wflign/src/wflign.cpp:108-483,
 * PAF branch) for a BATCH of mapping records: main end-to-end biWFA (wfb_align_batch), head and tail patching
 * through ends-free alignments (wfb_align_endsfree_batch; wflign.cpp:167-418), try_swap_start/end_pattern
 * (wflign_swizzle.cpp:220-300), trim_indels + process_compressed_cigar + write_alignment_paf
 * (wflign_patch.cpp:139-283, 2611-2724). Argument meaning follows the reference's parameter list (wflign.cpp:108-133);
 * the call site is Aligner::processAlignment (src/align/include/computeAlignments.hpp:695-720). */
typedef struct {
  const char* query_name;
  const char* query;          /* strand-corrected, upper-cased query slice (text)  */
  uint64_t query_total_length;
  uint64_t query_offset;
  uint64_t query_length;
  int32_t query_is_rev;
  int32_t chain_id;
  const char* target_name;
  const char* target;         /* upper-cased target slice (pattern)                */
  uint64_t target_total_length;
  uint64_t target_offset;
  uint64_t target_length;
  int32_t chain_length;
  int32_t chain_pos;
  float mashmap_estimated_identity;
  int32_t reserved_;
} wfb_record_t;

typedef struct {
  int32_t disable_chain_patching; /* wflign.cpp:125 */
  int32_t term_group;             /* 1 / 8 / 16, see wfb_align_endsfree_batch; 0 = 8 (the AVX2 build, what the fixtures were made with);
                                     -1 = what a -march=native build of the reference would do on this host (cpuid: AVX-512 CD+VL -> 16,
                                     AVX2 -> 8, else 1): the reference's head / tail patch CIGARs differ between its ISA builds */
  float min_identity;             /* wflign_patch.cpp:2624-2626 filters */
  float min_block_identity;
  uint64_t min_alignment_length;
  /* the SAM branch of do_biwfa_alignment (wflign.cpp:455-481 -> write_alignment_sam, wflign_patch.cpp:2480-2609) */
  int32_t sam_format;             /* 0 = PAF lines (paf_format_else_sam = true), 1 = SAM records (-a)         */
  int32_t emit_md_tag;            /* -d: append MD:Z: (write_tag_and_md_string, wflign_patch.cpp:2397-2478)    */
  int32_t no_seq_in_sam;          /* SEQ column '*' instead of the aligned query bases                         */
  int32_t reserved_;
} wfb_paf_params_t;

#define WFB_REC_WRITTEN 0       /* a PAF line was produced                                                    */
#define WFB_REC_FILTERED 1      /* aligned but rejected by the identity / length filters (nothing written)     */
#define WFB_REC_UNALIGNED 2     /* main alignment status != 0: the reference returns early (wflign.cpp:150-152) */
#define WFB_REC_PATCH_CAP (-1)  /* (no longer produced: a patch that exceeds the ends-free kernel's caps keeps the main CIGAR and is
                                   counted in wfb_align_stats_t.patch_cap_kept_main) */

/* out receives the PAF lines ('\n'-terminated) of records 0..n-1 back to back; line_offset[n+1] delimits them
 * (empty for records that produce no line); rec_status[n] = WFB_REC_*. WFB_ECAP if out_cap is too small
 * (*out_len then holds the required size). */
int wfb_biwfa_paf_batch(wfb_aligner_t*, const wfb_record_t* recs, int32_t n, const wfb_paf_params_t* params, char* out,
                        int64_t out_cap, int64_t* out_len, int64_t* line_offset, int32_t* rec_status, wfb_align_stats_t* stats);

/* Device memory helpers so callers without a CUDA binding (ctypes) can stage inputs. */
void* wfb_device_malloc(int device, uint64_t bytes);
void wfb_device_free(int device, void* p);
int wfb_memcpy_h2d(int device, void* dst, const void* src, uint64_t bytes);

/* ------------------------------------------------------------------------------------------------
 * Path 1 — MashMap 3.5 query-fragment sketch.
 * Replaces skch::CommonFunc::sketchSequence as called by MappingCore::getSeedHits
 * (src/map/include/commonFunc.hpp:217-323; src/map/include/mappingCore.hpp:61-76) for a BATCH of
 * fragments.
 * ---------------------------------------------------------------------------------------------- */

/* skch::MinmerInfo (src/map/include/base_types.hpp:28-60), 32-byte reference layout. */
typedef struct {
  uint64_t hash;
  int64_t wpos;
  int64_t wpos_end;
  int32_t seqId;
  int16_t strand; /* FWD = 1, AMBIG = 0, REV = -1 (base_types.hpp:101-106) */
  int16_t pad_;
} wfb_minmer_t;

typedef struct {
  int64_t seq_offset; /* fragment start within seq_base                  */
  int32_t len;        /* fragment length (param.windowLength on the CLI path) */
  int32_t seq_id;     /* copied into MinmerInfo::seqId                   */
} wfb_frag_t;

/* Sketch n fragments: out[i*sketch_size .. i*sketch_size+out_count[i]) receives the fragment's
 * minmers sorted ascending by hash, as sketchSequence leaves them. seq_base is raw FASTA bases
 * (any case); upper-casing / N-masking (makeUpperCaseAndValidDNA, commonFunc.hpp:132-142) is fused
 * into the kernel. kmer_size <= 32. */
int wfb_sketch_fragments(int device, const char* seq_base, int64_t seq_bytes, const wfb_frag_t* frags, int32_t n,
                         int32_t kmer_size, int32_t sketch_size, wfb_minmer_t* out, int32_t* out_count,
                         double* kernel_ms);

/* ------------------------------------------------------------------------------------------------
 * Path 1 — reference-side windowed minmers.
 * Replaces skch::CommonFunc::addMinmers as run per target sequence by Sketch::buildHelper
 * (src/map/include/commonFunc.hpp:439-708; src/map/include/winSketch.hpp:467-499) for a BATCH of target
 * sequences. Output = the concatenation Sketch::build makes of the per-sequence results
 * (winSketch.hpp:424-429): ordered by sequence (input order), then (wpos, wpos_end); records with equal
 * (wpos, wpos_end) in the order the reference's unstable std::sort (GNU libstdc++) leaves them — the L2 stage
 * depends on it for targets of one to two windows (such sequences are put in order on the host: tie_sequences).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double stream_kernel_ms; /* CUDA-event time of the streaming-window kernel             */
  double total_kernel_ms;  /* ... of all kernels incl. stitch, post passes and sorts      */
  uint64_t bases;          /* target bases processed                                      */
  uint64_t raw_records;    /* records emitted by the stream before the post passes        */
  uint64_t chunks;         /* independent chunks (threads) of the stream kernel           */
  uint64_t stale_absorbed; /* expired heap entries absorbed into a sketch entry (commonFunc.hpp:635-641):
                              the only situation in which chunking can deviate from the reference  */
  uint64_t stitch_miss;    /* interval starts that could not be inherited across a chunk boundary (0 = exact) */
  uint64_t filtered;       /* 1 = candidate-filtered build: only k-mers below a hash threshold entered the window machine */
  uint64_t candidates;     /* ... how many of them (all sequences)                                                       */
  uint64_t redo_chunks;    /* ... chunks whose filtered run could not vouch for its result and were re-run over every k-mer */
  double cand_kernel_ms;     /* filtered build: CUDA-event time of the cleaning + candidate kernels ...  */
  double filtered_stream_ms; /* ... of the window machine over the candidate stream ...                  */
  double redo_ms;            /* ... of the exact re-run of the flagged chunks (all three inside stream_kernel_ms) */
  uint64_t tie_sequences;    /* sequences with records tying on (wpos, wpos_end), put in the reference's std::sort order on the host */
} wfb_minmer_stats_t;

/* seq_ptrs[i] / seq_lens[i] : raw FASTA bases of target i (any case); seq_ids[i] -> MinmerInfo::seqId.
 * Sequences shorter than window_size are skipped like Sketch::build does (winSketch.hpp:218-232).
 * Environment switches (tuning and tests only; every setting returns the same bytes): WFB_MM_FILTER=0 runs the window machine over every
 * k-mer instead of the candidates below the hash threshold; WFB_MM_FCHUNK / WFB_MM_CHUNK = positions per chunk thread of the filtered /
 * unfiltered run [1024]; WFB_MM_LCUR=0 keeps a window queue in the filtered run; WFB_MM_FSMEM=1 puts the filtered run's containers in
 * shared memory (slower); WFB_MM_REDO_SMEM=0 re-runs flagged chunks from global-memory slabs; WFB_MM_CAND_CAP=n caps a tile's candidate
 * region (forces the overflow path). */
int wfb_minmers_build(int device, const char* const* seq_ptrs, const int64_t* seq_lens, const int32_t* seq_ids, int32_t nseq,
                      int32_t kmer_size, int32_t window_size, int32_t sketch_size, wfb_minmer_t* out, int64_t out_cap,
                      int64_t* out_count, wfb_minmer_stats_t* stats);

/* ------------------------------------------------------------------------------------------------
 * Path 1 — GPU-resident reference index and the batched L1 stage.
 * wfb_index_build replaces skch::Sketch::Sketch(...)/Sketch::build (src/map/include/winSketch.hpp:141-154,
 * 175-457): minmerIndex (kept minmers, reference order) and minmerPosLookupIndex (hash -> interval points)
 * live in HBM; wfb_index_export returns them for the host L2 stage and for parity dumps.
 * wfb_l1_batch replaces MappingCore::getSeedHits / getSeedIntervalPoints / computeL1CandidateRegions as
 * driven by Map::doL1Mapping (src/map/include/mappingCore.hpp:61-301; src/map/include/computeMap.hpp:945-983)
 * for a batch of query fragments of length == window_size.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t kmer_size;      /* param.kmerSize     */
  int32_t window_size;    /* param.windowLength */
  int32_t sketch_size;    /* param.sketchSize   */
  int32_t index_threads;  /* the reference's -t: abutting intervals of one hash merge only inside one worker's
                             contiguous range of sequences (winSketch.hpp:271-277,379-387) */
  double max_kmer_freq;   /* param.max_kmer_freq (-F, default 0.0002) */
} wfb_index_params_t;

typedef struct {
  wfb_minmer_stats_t minmer;
  double index_kernel_ms;
  uint64_t total_windows, kept_minmers, interval_points, unique_hashes, count_threshold, table_buckets;
} wfb_index_stats_t;

typedef struct wfb_index wfb_index_t;

wfb_index_t* wfb_index_build(int device, const wfb_index_params_t* params, const char* const* seq_ptrs, const int64_t* seq_lens,
                             const int32_t* seq_ids, int32_t nseq, wfb_index_stats_t* stats);
void wfb_index_free(wfb_index_t*);
int wfb_index_sizes(const wfb_index_t*, int64_t* n_minmers, int64_t* n_unique_hashes, int64_t* n_points, uint64_t* count_threshold);
/* Interval points come back packed as sort keys: seqId << 41 | pos << 1 | (side == OPEN), grouped by hash:
 * points[ustart[u] .. ustart[u]+ucount[u]) belong to uhash[u] (ascending), in the reference's push order. */
int wfb_index_export(const wfb_index_t*, wfb_minmer_t* minmers, int64_t minmers_cap, uint64_t* uhash, uint32_t* ustart,
                     uint32_t* ucount, int64_t uniq_cap, uint64_t* points, int64_t points_cap);

typedef struct { /* QueryMetaData (src/map/include/base_types.hpp:336-349) of the sequence a fragment comes from */
  int32_t q_seq_id;
  int32_t q_group; /* idManager.getRefGroup(Q.seqId) */
} wfb_frag_query_t;

typedef struct {
  int32_t minimum_hits;           /* Map::cached_minimum_hits (computeMap.hpp:160): computed by the host from
                                     Stat::estimateMinimumHitsRelaxed, max'ed with -H */
  const int32_t* sketch_cutoffs;  /* Map::sketchCutoffs from setProbs() (computeMap.hpp:234-293) */
  int32_t n_cutoffs;
  const int32_t* ref_group;       /* idManager.getRefGroup(seqId) for every target seqId */
  int32_t n_ref_group;
  int32_t skip_self, skip_prefix, lower_triangular; /* map_parameters.hpp:59-61 */
  float kmer_complexity_threshold; /* param.kmerComplexityThreshold (0 on the CLI path) */
} wfb_l1_params_t;

typedef struct { /* L1_candidateLocus_t (src/map/include/mappingCore.hpp:24-30) */
  int32_t seqId;
  int32_t intersectionSize; /* the L1 hit count */
  int64_t rangeStartPos, rangeEndPos;
} wfb_l1_locus_t;

typedef struct {
  /* caller-allocated outputs */
  wfb_minmer_t* q_minmers;   /* [n * sketch_size] Q.minmerTableQuery per fragment (may be NULL) */
  int32_t* q_count;          /* [n] Q.sketchSize (may be NULL) */
  float* q_complexity;       /* [n] Q.kmerComplexity (may be NULL) */
  wfb_l1_locus_t* loci;      /* [loci_cap] */
  int64_t loci_cap;
  int64_t* frag_loci_offset; /* [n] */
  int32_t* frag_loci_count;  /* [n] */
  int32_t* frag_status;      /* [n] 0 or WFB_ECAP (fragment exceeded an internal capacity) */
  /* set by the call */
  int64_t n_loci;
  double kernel_ms;
} wfb_l1_out_t;

int wfb_l1_batch(const wfb_index_t*, const wfb_l1_params_t*, const char* seq_base, int64_t seq_bytes, const wfb_frag_t* frags,
                 const wfb_frag_query_t* frag_queries, int32_t n, wfb_l1_out_t* out);

/* ------------------------------------------------------------------------------------------------
 * Path 1 — L2 stage on the GPU-resident index, fused behind the L1 stage: the L1 loci never leave
 * the device. wfb_map_fragments_batch replaces Map::mapSingleQueryFrag's two stages for a batch of fragments
 * (src/map/include/computeMap.hpp:879-921): doL1Mapping (:945-983) as wfb_l1_batch does, then per PanSN group slice
 * doL2Mapping (:989-1061) = stage-1 top-ANI test on the L1 hit count, MappingCore::computeL2MappedRegions
 * (src/map/include/mappingCore.hpp:306-442) with SlideMapper (src/map/include/slidingMap.hpp:28-215), the identity
 * test, and the final sort by (refSeqId, refStartPos) (:919-920).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  /* [sketch_size + 1], indexed by Q.sketchSize: smallest L1 intersectionSize that passes the stage-1 top-ANI test
   * (computeMap.hpp:999-1012; wfb_stage1_min_hits computes it). NULL = param.stage1_topANI_filter off. */
  const int32_t* stage1_min_hits;
  int32_t n_stage1_min_hits;
  /* [sketch_size + 1], indexed by Q.sketchSize: smallest sharedSketchSize whose mapping is kept by the identity test
   * (computeMap.hpp:1018-1024). With keep_low_pct_id (the CLI default, parse_args.hpp:173) the test uses
   * Stat::md_lower_bound, i.e. GSL's binomial tail: the host, which links GSL, fills the table;
   * wfb_l2_min_shared fills it for keep_low_pct_id == false. NULL = keep every mapping. */
  const int32_t* l2_min_shared;
  int32_t n_l2_min_shared;
} wfb_l2_params_t;

typedef struct { /* the MappingResult fields doL2Mapping sets (computeMap.hpp:1029-1044; queryStartPos = 0,
                    blockLength = Q.len, n_merged = 1) + L2_mapLocus_t's optimalStart / optimalEnd */
  int32_t frag;              /* index of the fragment in the batch */
  int32_t refSeqId;
  int64_t refStartPos;       /* L2_mapLocus_t::meanOptimalPos */
  int64_t optimalStart, optimalEnd;
  int32_t conservedSketches; /* L2_mapLocus_t::sharedSketchSize */
  int32_t strand;            /* FWD = 1, REV = -1 */
  float nucIdentity;         /* 1 - Stat::j2md(conservedSketches / Q.sketchSize, k) */
  float kmerComplexity;      /* Q.kmerComplexity */
} wfb_l2_mapping_t;

typedef struct {
  /* caller-allocated outputs */
  wfb_l2_mapping_t* mappings;  /* [mappings_cap], grouped by fragment, inside a fragment by (refSeqId, refStartPos) */
  int64_t mappings_cap;
  int64_t* frag_map_offset;    /* [n + 1]: fragment i owns mappings[frag_map_offset[i] .. frag_map_offset[i+1]) */
  int32_t* frag_status;        /* [n] 0 or WFB_ECAP */
  wfb_l1_out_t* l1;            /* optional: also return the L1 stage's outputs (parity dumps); NULL on the fast path */
  /* set by the call */
  int64_t n_mappings;
  double l1_kernel_ms, l2_kernel_ms, sort_kernel_ms;
  uint64_t n_l1_loci;          /* L1 loci produced                                   */
  uint64_t l2_loci;            /* ... that passed the stage-1 test and were scanned    */
  uint64_t l2_steps;           /* minmerIndex records the L2 kernel visited (32 B each) */
} wfb_map_out_t;

int wfb_map_fragments_batch(const wfb_index_t*, const wfb_l1_params_t*, const wfb_l2_params_t*, const char* seq_base, int64_t seq_bytes,
                            const wfb_frag_t* frags, const wfb_frag_query_t* frag_queries, int32_t n, wfb_map_out_t* out);
/* out[sketch_size + 1]; hg_numerator = param.hgNumerator (1.0), ani_diff = param.ANIDiff (0.0) */
int wfb_stage1_min_hits(double hg_numerator, float ani_diff, int32_t kmer_size, int32_t sketch_size, int32_t* out);
/* out[sketch_size + 1] for keep_low_pct_id == false; percentage_identity in [0,1] */
int wfb_l2_min_shared(float percentage_identity, int32_t kmer_size, int32_t sketch_size, int32_t* out);

/* ------------------------------------------------------------------------------------------------
 * Path 1 -> path 2 hand-over, host stage (first part): the fragment mappings of a query become chains,
 * and every chain is cut into merged mappings of at most max_mapping_length: the records the aligner receives.
 * Host C++ like the reference (one query per host thread).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { /* skch::MappingResult (src/map/include/base_types.hpp:153-163), same 28-byte layout */
  uint32_t refSeqId;
  uint32_t refStartPos;
  uint32_t queryStartPos;
  uint32_t blockLength;
  uint32_t n_merged;
  uint32_t conservedSketches;
  uint16_t nucIdentity;   /* identity * 10000, rounded (setNucIdentity)          */
  uint8_t flags;          /* bit 0 = REV strand, bit 1 = discard, bit 2 = overlapped */
  uint8_t kmerComplexity; /* complexity * 100, rounded (setKmerComplexity)       */
} wfb_mapping_t;

typedef struct { /* skch::ChainInfo (base_types.hpp:261-265): the ch:Z:<chainId>.<chainPos>.<chainLen> tag */
  uint32_t chainId;
  uint16_t chainPos;
  uint16_t chainLen;
} wfb_chain_info_t;

typedef struct {
  int32_t split;               /* param.split (default true; false = -N: nothing is chained)  */
  int32_t reserved_;
  int64_t chain_gap;           /* param.chain_gap (-c, default 2000)                          */
  int64_t window_length;       /* param.windowLength                                          */
  uint64_t max_mapping_length; /* param.max_mapping_length (-P, default 50000)                */
} wfb_chain_params_t;

/* What Map::processFragment + OutputHandler::mappingBoundarySanityCheck do to the L2 results of ONE query before the
 * filters see them (src/map/include/computeMap.hpp:121-127,1029-1046; src/map/include/mappingOutput.hpp:31-69):
 * MappingResult construction (scaled identity / complexity, strand flag), queryStartPos = fragmentIndex * windowLength,
 * clamping against the query / reference lengths. frag_index[f] = FragmentData::fragmentIndex of batch fragment f
 * (the overlapping tail fragment carries noOverlapFragmentCount, computeMap.hpp:612). out[n]. */
int wfb_l2_to_query_mappings(const wfb_l2_mapping_t* l2, int64_t n, const int32_t* frag_index, int64_t window_length, int64_t query_len,
                             const int64_t* ref_seq_len, wfb_mapping_t* out);

/* skch::MappingFilterUtils::mergeMappingsInRangeWithChains (src/map/include/mappingFilter.hpp:381-571) as called by
 * Map::filterSubsetMappings (src/map/include/computeMap.hpp:1090-1094), for a batch of queries:
 * query q owns mappings[query_offset[q] .. query_offset[q+1]) (reordered in place, as the reference leaves readMappings)
 * and receives merged[merged_offset[q] .. merged_offset[q+1]) with their chain tags. host_threads <= 0: all cores. */
int wfb_chain_mappings_batch(const wfb_chain_params_t* params, wfb_mapping_t* mappings, const int64_t* query_offset, int32_t n_queries,
                             wfb_mapping_t* merged, wfb_chain_info_t* chain_info, int64_t merged_cap, int64_t* merged_offset,
                             int32_t host_threads);

/* ------------------------------------------------------------------------------------------------
 * SURVEY 8 f2, second part + b3: everything Map::filterSubsetMappings does to the mappings of one query after the
 * chain merge (src/map/include/computeMap.hpp:1076-1165), the mapping PAF text (the reference's map -> align
 * dispatch format, src/map/include/mappingOutput.hpp:74-139) and its reader on the aligner side
 * (Aligner::parseMashmapRow + createSeqRecord, src/align/include/computeAlignments.hpp:195-303,582-660).
 * Host C++ like the reference.
 * ---------------------------------------------------------------------------------------------- */
#define WFB_FILTER_MAP 1      /* skch::filter::MAP (default)      */
#define WFB_FILTER_ONETOONE 2 /* skch::filter::ONETOONE (-o)      */
#define WFB_FILTER_NONE 3     /* skch::filter::NONE (-f)          */

typedef struct { /* the skch::Parameters fields the filters read (src/map/include/map_parameters.hpp:31-109) */
  int32_t split;                    /* param.split (true unless -N)                                     */
  int32_t merge_mappings;           /* param.mergeMappings (true unless -M)                             */
  int32_t filter_mode;              /* WFB_FILTER_*                                                     */
  int32_t skip_prefix;              /* param.skip_prefix (-Y given): plane sweep per PanSN target group */
  int32_t filter_length_mismatches; /* param.filterLengthMismatches (true on the CLI path)              */
  int32_t drop_rand;                /* param.dropRand (false on the CLI path)                           */
  int32_t threads;                  /* param.threads (only sizes the scaffold distance loop)            */
  int32_t legacy_output;            /* param.legacy_output (false on the CLI path)                      */
  int64_t chain_gap;                /* -c, default 2000                                                 */
  int64_t window_length;            /* -w                                                               */
  int64_t block_length;             /* -l, default 0                                                    */
  uint64_t max_mapping_length;      /* -P, default 50000                                                */
  uint64_t sparsity_hash_threshold; /* default UINT64_MAX = keep everything                             */
  uint32_t num_mappings_for_segment;  /* -n; default UINT32_MAX ("inf")                                 */
  uint32_t num_mappings_for_scaffold; /* -r; default 1                                                  */
  double overlap_threshold;           /* -O, default 0.95                                               */
  double scaffold_overlap_threshold;  /* --scaffold-overlap, default 0.5                                */
  int64_t scaffold_gap;               /* -j, default 100000; <= 0 disables the scaffold filter          */
  int64_t scaffold_max_deviation;     /* -D, default 100000                                             */
  int64_t scaffold_min_length;        /* -S, default 10000                                              */
  float percentage_identity;          /* param.percentageIdentity in [0,1]                              */
  int32_t reserved_;
} wfb_filter_params_t;

/* Map::filterSubsetMappings for a batch of queries: chain merge (as wfb_chain_mappings_batch), filterWeakMappings,
 * filterByGroup (query plane sweep, src/map/include/filter.hpp:170-240), filterFalseHighIdentity, sparsifyMappings,
 * filterByScaffolds (src/map/include/mappingFilter.hpp:154-293,831-1016). Query q owns mappings[query_offset[q] ..
 * query_offset[q+1]) (not modified) and receives out[out_offset[q] .. out_offset[q+1]) = the mappings the reference
 * prints for it (the merged ones when merge_mappings && split, else the filtered fragment mappings) with the ChainInfo
 * the reference pairs them with (the i-th surviving mapping gets the i-th pre-filter entry, computeMap.hpp:664-667 +
 * mappingOutput.hpp:96-97). ref_group[refSeqId] is only read when skip_prefix. host_threads <= 0: all cores. */
int wfb_filter_mappings_batch(const wfb_filter_params_t* params, const wfb_mapping_t* mappings, const int64_t* query_offset,
                              const int64_t* query_len, int32_t n_queries, const int32_t* ref_group, const int64_t* ref_seq_len,
                              wfb_mapping_t* out, wfb_chain_info_t* out_chain, int64_t out_cap, int64_t* out_offset, int32_t host_threads);

/* skch::MappingFilterUtils::filterByGroup alone (mappingFilter.hpp:220-293): filter_ref = 0 plane sweep over the query
 * axis, 1 over the reference axis (filter.hpp:474-535). mappings is reordered in place like the reference's
 * unfilteredMappings. Returns the number of survivors (<0: WFB_E*); out[out_cap]. */
int64_t wfb_filter_by_group(const wfb_filter_params_t* params, wfb_mapping_t* mappings, int64_t n, int32_t n_mappings, int32_t filter_ref,
                            const int32_t* ref_group, const int64_t* ref_seq_len, wfb_mapping_t* out, int64_t out_cap);

/* The final pass of the one-to-one mode (computeMap.hpp:788-850): the surviving mappings of ALL queries are regrouped by
 * target sequence, plane-swept over the reference axis, and handed back to every query that holds a mapping with the same
 * (refSeqId, refStartPos, queryStartPos) — including the duplicates this rule creates when two queries share such a
 * triple. Queries are visited in ascending q and targets in ascending refSeqId (the reference iterates unordered_maps,
 * so its line ORDER is unspecified; the multiset of lines is what is reproduced). out_query[i] = owner of out[i];
 * out is grouped by query. Returns the count (<0: WFB_E*). */
int64_t wfb_one_to_one_filter(const wfb_filter_params_t* params, const wfb_mapping_t* mappings, const int64_t* query_offset, int32_t n_queries,
                              const int32_t* ref_group, const int64_t* ref_seq_len, wfb_mapping_t* out, int32_t* out_query, int64_t out_cap);

/* OutputHandler::reportReadMappings (mappingOutput.hpp:74-139): the mapping PAF lines of one query, byte-compatible with
 * `wfmash -m` (so they feed the reference aligner through -i and vice versa). ref_names[refSeqId]. chain may be NULL
 * (every mapping its own chain, mappingOutput.hpp:143-160). Returns the text length; WFB_ECAP (with the needed size in
 * *needed, when not NULL) if buf_cap is too small. */
int64_t wfb_mapping_paf_format(const wfb_filter_params_t* params, const wfb_mapping_t* mappings, const wfb_chain_info_t* chain, int64_t n,
                               const char* query_name, int64_t query_len, const char* const* ref_names, const int64_t* ref_seq_len, char* buf,
                               int64_t buf_cap, int64_t* needed);

typedef struct { /* align::MappingBoundaryRow (src/align/include/align_types.hpp) + what createSeqRecord derives from it */
  int64_t q_start, q_end;           /* after query padding                                               */
  int64_t r_start, r_end;           /* after target padding                                              */
  int64_t ref_fetch_start, ref_fetch_len; /* target slice incl. the wflign_max_len_minor head / tail room (computeAlignments.hpp:611-624) */
  int64_t query_len, ref_len;       /* columns 2 and 7                                                   */
  int64_t chain_id, chain_length, chain_pos;
  int32_t strand;                   /* 1 = '+', -1 = '-'                                                 */
  float mashmap_estimated_identity;
  int32_t q_name_off, q_name_len;   /* byte ranges of the two names inside the line                      */
  int32_t r_name_off, r_name_len;
} wfb_mapping_row_t;

/* Aligner::parseMashmapRow for one line (no trailing '\n' needed). WFB_EINVAL (message in wfb_last_error) where the
 * reference throws: fewer than 13 tokens, unparsable numbers, padded coordinates beyond the reference length. */
int wfb_mapping_paf_parse(const char* line, int64_t line_len, uint64_t target_padding, uint64_t query_padding, uint64_t wflign_max_len_minor,
                          wfb_mapping_row_t* row);

/* ------------------------------------------------------------------------------------------------
 * Run-level constants of the mapping path (host; the reference derives them once per run with GNU GSL, which is an
 * un-vendored third-party dependency: restated from the distributions' definitions, see wfmash_b200/csrc/stats_host.cu).
 * ---------------------------------------------------------------------------------------------- */
/* param.sketchSize when -s is not given (src/interface/parse_args.hpp:642-644) */
int32_t wfb_sketch_size(float percentage_identity, int64_t window_length, int32_t kmer_size);
/* skch::Stat::estimateMinimumHitsRelaxed (src/map/include/map_stats.hpp:159-180): Map::cached_minimum_hits
 * (computeMap.hpp:160) = max(param.minimum_hits, this) -> wfb_l1_params_t::minimum_hits. confidence_interval = 0.95
 * (skch::fixed::confidence_interval). */
int32_t wfb_estimate_minimum_hits_relaxed(int32_t sketch_size, int32_t kmer_size, float percentage_identity, float confidence_interval);
/* Map::sketchCutoffs (computeMap.hpp:150,234-293) -> wfb_l1_params_t::sketch_cutoffs. out[min(sketch_size,1000) + 1];
 * all ones unless stage1_top_ani_filter (true on the CLI path). ani_diff = param.ANIDiff (0), ani_diff_conf =
 * param.ANIDiffConf (0.999). */
int wfb_sketch_cutoffs(int32_t sketch_size, int32_t kmer_size, float ani_diff, float ani_diff_conf, int32_t stage1_top_ani_filter, int32_t* out,
                       int32_t out_len);
/* wfb_l2_min_shared for keep_low_pct_id == true (the CLI default, parse_args.hpp:173): the identity test of
 * Map::doL2Mapping (computeMap.hpp:1016-1024) passes when the UPPER bound of the identity's confidence interval
 * (Stat::md_lower_bound, map_stats.hpp:93-126) reaches percentage_identity. out[sketch_size + 1]. */
int wfb_l2_min_shared_relaxed(float percentage_identity, int32_t kmer_size, int32_t sketch_size, float confidence_interval, int32_t* out);
/* the three GSL functions as restated here (exported for the cross-check against scipy.stats) */
double wfb_stat_binomial_Q(uint32_t k, double p, uint32_t n);
double wfb_stat_hypergeometric_pdf(uint32_t k, uint32_t n1, uint32_t n2, uint32_t t);
double wfb_stat_hypergeometric_P(uint32_t k, uint32_t n1, uint32_t n2, uint32_t t);

/* ------------------------------------------------------------------------------------------------
 * SURVEY 8 f3 — ANI auto-identity: skch::Stat::estimate_identity_for_groups (src/map/include/map_stats.hpp:325-822), the
 * `-p` default of the CLI (src/interface/main.cpp:75-134): decides percentageIdentity, hence sketch size and minimum hits.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double hash_kernel_ms;  /* ani_hash_kernel, all passes               */
  double sort_kernel_ms;  /* segmented radix sort + gather              */
  uint64_t bases;         /* sequence bytes streamed                    */
  uint64_t valid_kmers;   /* canonical, non-palindromic, N-free k-mers  */
  uint64_t candidates;    /* hashes that passed the group thresholds    */
  uint64_t tiles;
  int32_t passes;         /* 1 unless a threshold / capacity was revised */
  int32_t reserved_;
} wfb_ani_stats_t;

/* Per-group MinHash sketches on the GPU: group g's sketch = the min(sketch_size, #k-mers) smallest elements of the MULTISET
 * of canonical k-mer hashes (Murmur3, seed 42, min of both strands, palindromes and k-mers touching a non-ACGT base dropped;
 * map_stats.hpp:563-613) of the sequences with seq_group[i] == g — what the reference's per-sequence StreamingMinHash heaps
 * merged per PanSN group hold (streamingMinHash.hpp:89-99, map_stats.hpp:617-637). The reference keeps separate sketches for
 * the query role and the target role of a group: call once per role (or once, when both files are the same).
 * sketches[n_groups * sketch_size] ascending per group, sketch_count[n_groups]. kmer_size 21 and sketch_size 4096 on the CLI. */
int wfb_ani_group_sketches(int device, const char* const* seq_ptrs, const int64_t* seq_lens, const int32_t* seq_group, int32_t nseq,
                           int32_t n_groups, int32_t kmer_size, int32_t sketch_size, uint64_t* sketches, int32_t* sketch_count,
                           wfb_ani_stats_t* stats);

/* The host part (map_stats.hpp:690-800): every query-role sketch against every target-role sketch of another group id,
 * ANI = 1 - j2md(|intersection| / min(sizes)), the ani_percentile-th value of the sorted ANIs plus ani_adjustment / 100,
 * clamped to [0, 1]; 0.70 when nothing overlaps. CLI defaults: ani_percentile 50, ani_adjustment -2.0. */
double wfb_ani_estimate_identity(const uint64_t* q_sketch, const int32_t* q_count, const int32_t* q_group, int32_t nq, const uint64_t* t_sketch,
                                 const int32_t* t_count, const int32_t* t_group, int32_t nt, int32_t sketch_size, int32_t kmer_size,
                                 int32_t ani_percentile, float ani_adjustment, int32_t* n_comparisons);

/* ------------------------------------------------------------------------------------------------
 * The two phases of src/interface/main.cpp as ONE call each over sequences held in host memory (no CLI, no FASTA / index
 * files): what skch::Map (src/map/include/computeMap.hpp:300-860) and align::Aligner
 * (src/align/include/computeAlignments.hpp:318-720) do between "sequences loaded" and "text written". Every step inside is
 * one of the entry points above; the text handed from one to the other is the reference's mapping PAF.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const char* name; /* sequence name (PanSN: sample#hap#contig)                                  */
  const char* seq;  /* bases, any case, not NUL-terminated                                       */
  int64_t len;
} wfb_seq_t;

typedef struct { /* the skch::Parameters fields of the mapping phase; CLI defaults in brackets (src/interface/parse_args.hpp) */
  int32_t kmer_size;            /* -k [15]                                                         */
  int32_t sketch_size;          /* -s; <= 0 = derive from identity / window / k (wfb_sketch_size)   */
  int32_t minimum_hits;         /* -H; < 0 = auto (wfb_estimate_minimum_hits_relaxed)               */
  int32_t index_threads;        /* param.threads of the index build (partitions of the postings)    */
  int64_t window_length;        /* -w [1000]                                                       */
  float percentage_identity;    /* -p as a fraction; <= 0 = estimate it (ANI auto-identity, the CLI default `ani50-2`) */
  float ani_adjustment;         /* [-2.0]                                                          */
  int32_t ani_percentile;       /* [50]                                                            */
  int32_t skip_self;            /* [1] no -X                                                       */
  int32_t skip_prefix;          /* [1] -Y '#'                                                      */
  int32_t lower_triangular;     /* [0] -L                                                          */
  int32_t stage1_top_ani_filter;/* [1]                                                             */
  int32_t keep_low_pct_id;      /* [1]                                                             */
  int32_t prefix_delim;         /* ['#'] as a character code; group = name up to its LAST delimiter */
  int32_t reserved_;
  double max_kmer_freq;         /* -F [0.0002]                                                     */
  double hg_numerator;          /* [1.0]                                                           */
  float ani_diff;               /* [0.0]                                                           */
  float ani_diff_conf;          /* [0.999]                                                         */
  wfb_filter_params_t filter;   /* chain / filter stage; window_length, percentage_identity and skip_prefix are overwritten
                                   with the values above                                            */
} wfb_map_phase_params_t;

typedef struct {
  int64_t fragments, l2_mappings, mappings; /* query fragments, fragment mappings before / after the filters */
  int32_t sketch_size, minimum_hits;        /* as resolved                                                   */
  float percentage_identity;                /* as resolved (estimated when the parameter was <= 0)           */
  int32_t stale_absorbed;                   /* wfb_minmer_stats_t::stale_absorbed of the index build (saturating): 0 = the
                                               windowed minmers are provably those of the reference's addMinmers */
  double index_seconds, map_kernel_ms, filter_seconds, total_seconds;
  double ani_seconds;                       /* ANI auto-identity (0 when -p was given)                        */
  double index_kernel_ms, ani_kernel_ms;    /* CUDA-event time of the index-build kernels / the ANI kernels   */
} wfb_map_phase_stats_t;

/* Mapping phase: ids + PanSN groups (SequenceIdManager, sequenceIds.hpp:284-441; targets first), optional ANI estimate,
 * index build, query fragments (computeMap.hpp:560-630), L1 + L2 kernels, per-query chain merge + filters, mapping PAF text.
 * *paf receives a malloc'ed buffer (free it with wfb_free_text) holding *paf_len bytes of `wfmash -m` output, grouped by
 * query in input order. The one-to-one mode runs its final reference-axis pass over all queries (wfb_one_to_one_filter). */
int wfb_map_phase(int device, const wfb_map_phase_params_t* params, const wfb_seq_t* targets, int32_t n_targets, const wfb_seq_t* queries,
                  int32_t n_queries, char** paf, int64_t* paf_len, wfb_map_phase_stats_t* stats);
/* The same for a SHARE of the queries (one process per GPU: queries are independent, computeMap.hpp:565-599): query_select[q] != 0
 * marks the queries this call maps; sequence ids, PanSN groups, the ANI estimate and the fragment order (hence the ch:Z: tags) still
 * come from ALL queries, so the concatenation of the shares' texts equals the text of one call over everything. NULL = all.
 * Not available with the one-to-one filter (its final pass needs every query's mappings). */
int wfb_map_phase_subset(int device, const wfb_map_phase_params_t* params, const wfb_seq_t* targets, int32_t n_targets, const wfb_seq_t* queries,
                         int32_t n_queries, const uint8_t* query_select, char** paf, int64_t* paf_len, wfb_map_phase_stats_t* stats);

typedef struct { /* the align::Parameters fields of the alignment phase */
  uint64_t target_padding;       /* -E [min(window_length, 5000)]                                   */
  uint64_t query_padding;        /* -U [min(window_length, 5000)]                                   */
  uint64_t wflign_max_len_minor; /* [window_length * 128]                                           */
  int32_t batch_records;         /* records per GPU batch; <= 0 = 32768                             */
  int32_t reserved_;
  wfb_paf_params_t output;       /* filters, patching, PAF / SAM                                    */
} wfb_align_phase_params_t;

typedef struct {
  int64_t records, written, skipped_lines; /* parsable mapping rows, output records, rows skipped like computeAlignments.hpp:368-372 */
  uint64_t aligned_bp;                     /* the reference's processed_alignment_length (computeAlignments.hpp:480,528)  */
  double kernel_ms, total_seconds;         /* kernel_ms: CUDA-event time of the main biWFA rounds                       */
  double persist_kernel_ms, patch_kernel_ms; /* ... of wfb_persist_kernel alone / of the head + tail patch rounds        */
  int64_t batches;                         /* GPU batches (= launches of wfb_persist_kernel)                             */
  uint64_t cells, base_cells, extend_matches, base_extend_matches, overlap_tests, score_steps, base_score_steps; /* wfb_align_stats_t, summed */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t patch_cap_kept_main, main_device_cap; /* wfb_align_stats_t, summed: output that can differ from the reference's is never silent */
} wfb_align_phase_stats_t;

/* Alignment phase over mapping PAF text: parseMashmapRow + padding, slices fetched with faidx's clamping, upper-casing /
 * N-masking, reverse complement of '-' queries (computeAlignments.hpp:195-303,582-686), batched do_biwfa_alignment
 * (wfb_biwfa_paf_batch). *out: malloc'ed text of *out_len bytes (wfb_free_text), records in the order of the rows. */
int wfb_align_phase(wfb_aligner_t* aligner, const wfb_align_phase_params_t* params, const char* mapping_paf, int64_t mapping_paf_len,
                    const wfb_seq_t* targets, int32_t n_targets, const wfb_seq_t* queries, int32_t n_queries, char** out, int64_t* out_len,
                    wfb_align_phase_stats_t* stats);

void wfb_free_text(char* text);

/* ------------------------------------------------------------------------------------------------
 * SURVEY 8 f4 (second half): the reference's on-disk index (`-W` writes it, `-I` reads it; Sketch::writeIndex / readIndex,
 * src/map/include/winSketch.hpp:554-979): one or more subsets, each = header (magic 0xDEADBEEFCAFEBABE, batch index /
 * count, index_by_size, target names, the SequenceIdManager's name -> id map + next id, sequenceIds.hpp:102-115),
 * parameters (windowLength i64, sketchSize i32, kmerSize i32), minmerIndex (count + 32-byte MinmerInfo records) and
 * minmerPosLookupIndex (count; per hash: key, count, 24-byte IntervalPoint records). Files written here load in
 * `wfmash -I`, and files written by `wfmash -W` load here.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { /* an index in host memory, in the layout wfb_index_export fills */
  wfb_minmer_t* minmers;  int64_t n_minmers;   /* minmerIndex, reference order                                  */
  uint64_t* uhash;        int64_t n_uniq;      /* distinct hashes, ascending                                    */
  uint32_t* ustart;       uint32_t* ucount;    /* points[ustart[i] .. ustart[i] + ucount[i]) belong to uhash[i] */
  uint64_t* points;       int64_t n_points;    /* seqId << 41 | pos << 1 | (OPEN ? 1 : 0), reference order per hash */
} wfb_index_view_t;

typedef struct {
  uint64_t batch_idx, total_batches; /* subset number / count (1 subset unless -b splits the targets)            */
  int64_t index_by_size;             /* param.index_by_size                                                        */
  int64_t window_length; int32_t sketch_size, kmer_size;
  int32_t n_targets, n_ids, next_id, reserved_;
  char* target_names;                /* n_targets names, '\n' separated (malloc'ed by the reader)                */
  char* id_names;                    /* n_ids names of the id map, '\n' separated, ids in id_values               */
  int32_t* id_values;
} wfb_index_file_header_t;

/* Appends (append != 0) or writes one subset. The header's names / ids describe the SequenceIdManager of the run
 * (target_names = the subset's targets; id_names / id_values = every known sequence). */
int wfb_index_file_write(const char* path, int32_t append, const wfb_index_file_header_t* header, const wfb_index_view_t* index);
/* Reads the subset that starts at *offset (0 = first) and advances *offset past it. header and index receive malloc'ed
 * arrays: release them with wfb_index_file_release. WFB_EINVAL on a bad magic number / truncated file. */
int wfb_index_file_read(const char* path, int64_t* offset, wfb_index_file_header_t* header, wfb_index_view_t* index);
void wfb_index_file_release(wfb_index_file_header_t* header, wfb_index_view_t* index);
/* Uploads an index held in host memory (read from a file, or exported from another index) and builds the device lookup
 * table: the result maps exactly like an index built from the sequences. */
wfb_index_t* wfb_index_import(int device, const wfb_index_params_t* params, const wfb_index_view_t* index);

#ifdef __cplusplus
}
#endif
#endif
