/*
 * TEST INFRASTRUCTURE ONLY. Thin driver over the UNMODIFIED reference WFA2-lib C API
 * (/root/reference/deps/WFA2-lib, compiled in place by oracle/Makefile into oracle/_ref/).
 * It configures the aligner exactly as wfmash does
 *   - main alignment: wflign.cpp:136-148  (gap-affine-2p, Alignment scope, MemoryUltralow => biWFA,
 *                     heuristic none, end-to-end)
 *   - patch alignments: wflign.cpp:280-305,368-397 (MemoryMed, heuristic none, ends-free)
 * and returns the raw operation string (M/X/I/D, WFA order) and score.
 * Nothing in the product path links or loads this file.
 */
#include <stdio.h>
#include <stdint.h>
#include <stdbool.h>
#include <string.h>
#include <stdlib.h>
#include "wavefront/wfa.h"

/* memory_mode: 0 = high, 1 = med (piggyback), 2 = low, 3 = ultralow (biWFA) */
static wavefront_aligner_t* make_aligner(int x, int o1, int e1, int o2, int e2, int memory_mode) {
  wavefront_aligner_attr_t attr = wavefront_aligner_attr_default;
  attr.distance_metric = gap_affine_2p;
  attr.affine2p_penalties.match = 0;
  attr.affine2p_penalties.mismatch = x;
  attr.affine2p_penalties.gap_opening1 = o1;
  attr.affine2p_penalties.gap_extension1 = e1;
  attr.affine2p_penalties.gap_opening2 = o2;
  attr.affine2p_penalties.gap_extension2 = e2;
  attr.alignment_scope = compute_alignment;
  switch (memory_mode) {
    case 1: attr.memory_mode = wavefront_memory_med; break;
    case 2: attr.memory_mode = wavefront_memory_low; break;
    case 3: attr.memory_mode = wavefront_memory_ultralow; break;
    default: attr.memory_mode = wavefront_memory_high; break;
  }
  attr.heuristic.strategy = wf_heuristic_none;
  wavefront_aligner_t* a = wavefront_aligner_new(&attr);
  wavefront_aligner_set_heuristic_none(a);
  return a;
}

static int collect(wavefront_aligner_t* a, int status, char* ops_out, int ops_cap, int* ops_len, int* score) {
  if (status == 0) {
    const int n = a->cigar->end_offset - a->cigar->begin_offset;
    *ops_len = n;
    *score = a->cigar->score;
    if (n <= ops_cap) memcpy(ops_out, a->cigar->operations + a->cigar->begin_offset, (size_t)n);
    else status = -1000;
  } else {
    *ops_len = 0;
    *score = 0;
  }
  return status;
}

/* End-to-end alignment; returns the reference's status (0 = completed). */
int ref_wfa_end2end(const char* pattern, int plen, const char* text, int tlen,
                    int x, int o1, int e1, int o2, int e2, int memory_mode,
                    char* ops_out, int ops_cap, int* ops_len, int* score) {
  wavefront_aligner_t* a = make_aligner(x, o1, e1, o2, e2, memory_mode);
  wavefront_aligner_set_alignment_end_to_end(a);
  int status = wavefront_align(a, pattern, plen, text, tlen);
  status = collect(a, status, ops_out, ops_cap, ops_len, score);
  wavefront_aligner_delete(a);
  return status;
}

/* Ends-free alignment (patch alignments use memory_mode = 1). */
int ref_wfa_endsfree(const char* pattern, int plen, int pbegin_free, int pend_free,
                     const char* text, int tlen, int tbegin_free, int tend_free,
                     int x, int o1, int e1, int o2, int e2, int memory_mode,
                     char* ops_out, int ops_cap, int* ops_len, int* score) {
  wavefront_aligner_t* a = make_aligner(x, o1, e1, o2, e2, memory_mode);
  wavefront_aligner_set_alignment_free_ends(a, pbegin_free, pend_free, tbegin_free, tend_free);
  int status = wavefront_align(a, pattern, plen, text, tlen);
  status = collect(a, status, ops_out, ops_cap, ops_len, score);
  wavefront_aligner_delete(a);
  return status;
}

/* Batch end-to-end biWFA for CPU-baseline timing: one aligner per record, like wfmash
 * (wflign.cpp:136 constructs the aligner inside do_biwfa_alignment). Returns #completed;
 * sum of text lengths of completed alignments in *aligned_bp. */
int ref_wfa_batch_end2end(int n, const char* const* patterns, const int* plens,
                          const char* const* texts, const int* tlens,
                          int x, int o1, int e1, int o2, int e2, long long* aligned_bp, int* scores) {
  int ok = 0;
  long long bp = 0;
  for (int i = 0; i < n; ++i) {
    wavefront_aligner_t* a = make_aligner(x, o1, e1, o2, e2, 3);
    wavefront_aligner_set_alignment_end_to_end(a);
    const int status = wavefront_align(a, patterns[i], plens[i], texts[i], tlens[i]);
    if (status == 0) { ++ok; bp += tlens[i]; if (scores) scores[i] = a->cigar->score; }
    else if (scores) scores[i] = 1;
    wavefront_aligner_delete(a);
  }
  if (aligned_bp) *aligned_bp = bp;
  return ok;
}
