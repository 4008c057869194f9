/*
 * TEST INFRASTRUCTURE ONLY. Thin driver over the UNMODIFIED reference wflign sources
 * (/root/reference/src/common/wflign/src/{wflign,wflign_swizzle,wflign_patch,wflign_alignment}.cpp +
 * WFA2-lib + its C++ binding, compiled in place by oracle/Makefile into oracle/_ref/libwflignref.so).
 * Calls wflign::wavefront::do_biwfa_alignment (wflign.cpp:108) exactly like Aligner::processAlignment
 * (src/align/include/computeAlignments.hpp:695-720) and returns the PAF text it writes.
 */
#include <sstream>
#include <string>
#include <cstring>
#include "wflign.hpp"

extern "C" int ref_do_biwfa_alignment(const char* qname, const char* query, uint64_t q_total, uint64_t q_off, uint64_t q_len, int q_is_rev,
                                      const char* tname, const char* target, uint64_t t_total, uint64_t t_off, uint64_t t_len,
                                      int x, int o1, int e1, int o2, int e2, int disable_chain_patching, float min_identity,
                                      uint64_t min_aln_len, float min_block_id, uint64_t max_len_minor, float mm_id, int chain_id,
                                      int chain_len, int chain_pos, char* out, int out_cap) {
  wflign_penalties_t pen;
  pen.match = 0; pen.mismatch = x; pen.gap_opening1 = o1; pen.gap_extension1 = e1; pen.gap_opening2 = o2; pen.gap_extension2 = e2;
  std::stringstream ss;
  std::string q(query, q_len), t(target, t_len); /* the reference takes mutable char* buffers */
  wflign::wavefront::do_biwfa_alignment(qname, &q[0], q_total, q_off, q_len, q_is_rev != 0, tname, &t[0], t_total, t_off, t_len, ss, pen,
                                        false /*emit_md_tag*/, true /*paf*/, false /*no_seq_in_sam*/, disable_chain_patching != 0,
                                        min_identity, min_aln_len, min_block_id, max_len_minor, mm_id, chain_id, chain_len, chain_pos);
  const std::string s = ss.str();
  if ((int)s.size() + 1 > out_cap) return -(int)s.size();
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

/* the SAM branch (paf_format_else_sam = false): write_alignment_sam + write_tag_and_md_string */
extern "C" int ref_do_biwfa_alignment_sam(const char* qname, const char* query, uint64_t q_total, uint64_t q_off, uint64_t q_len, int q_is_rev,
                                          const char* tname, const char* target, uint64_t t_total, uint64_t t_off, uint64_t t_len,
                                          int x, int o1, int e1, int o2, int e2, int disable_chain_patching, float min_identity,
                                          uint64_t min_aln_len, float min_block_id, uint64_t max_len_minor, float mm_id, int chain_id,
                                          int chain_len, int chain_pos, int emit_md_tag, int no_seq_in_sam, char* out, int out_cap) {
  wflign_penalties_t pen;
  pen.match = 0; pen.mismatch = x; pen.gap_opening1 = o1; pen.gap_extension1 = e1; pen.gap_opening2 = o2; pen.gap_extension2 = e2;
  std::stringstream ss;
  std::string q(query, q_len), t(target, t_len);
  wflign::wavefront::do_biwfa_alignment(qname, &q[0], q_total, q_off, q_len, q_is_rev != 0, tname, &t[0], t_total, t_off, t_len, ss, pen,
                                        emit_md_tag != 0, false /*SAM*/, no_seq_in_sam != 0, disable_chain_patching != 0,
                                        min_identity, min_aln_len, min_block_id, max_len_minor, mm_id, chain_id, chain_len, chain_pos);
  const std::string s = ss.str();
  if ((int)s.size() + 1 > out_cap) return -(int)s.size();
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}
