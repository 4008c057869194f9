/*
 * TEST INFRASTRUCTURE ONLY. Compiles the reference's UNMODIFIED src/align/include/computeAlignments.hpp — the whole
 * alignment phase: align::Aligner reads the mapping PAF (parseMashmapRow: padding, chain tags), fetches the padded target /
 * query ranges through faidx, upper-cases / N-masks, reverse-complements '-' queries and calls do_biwfa_alignment per record
 * (computeAlignments.hpp:142-742) — over oracle/shims/common/faigz.h (htslib absent: uncompressed FASTA + .fai) and the
 * wflign / WFA2 objects of oracle/_ref/libwflignref.so. Pins SURVEY 8 row b3's reader side and wfb_align_phase as a whole.
 */
#include <cassert>
#include <climits>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "map/include/base_types.hpp" /* the reference includes its map headers first (src/interface/main.cpp) */
#include "align/include/computeAlignments.hpp"

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

static int g_align_threads = 1;
extern "C" void ref_align_set_threads(int32_t n) { g_align_threads = n > 0 ? n : 1; } /* 1 = records in row order (parity tests); more = timing runs */

extern "C" {
/* Writes the sequences as FASTA + .fai and the mapping PAF into `dir`, runs align::Aligner::compute() with one thread (records in
 * row order) and returns the bytes of its output file (PAF or SAM). -1 when the buffer is too small. */
int64_t ref_align_phase(const char* dir, const char* const* t_names, const char* const* t_seqs, const int64_t* t_lens, int32_t nt,
                        const char* const* q_names, const char* const* q_seqs, const int64_t* q_lens, int32_t nq, const char* mapping_paf,
                        int64_t mapping_paf_len, uint64_t target_padding, uint64_t query_padding, uint64_t wflign_max_len_minor, float min_identity,
                        uint64_t min_alignment_length, float min_block_identity, int32_t disable_chain_patching, int32_t sam_format,
                        int32_t emit_md_tag, int32_t no_seq_in_sam, char* out, int64_t out_cap) {
  const std::string d(dir), tf = d + "/t.fa", qf = d + "/q.fa", mp = d + "/map.paf", op = d + "/aln.out";
  write_fasta(tf, t_names, t_seqs, t_lens, nt);
  write_fasta(qf, q_names, q_seqs, q_lens, nq);
  { std::ofstream m(mp, std::ios::binary); m.write(mapping_paf, mapping_paf_len); }
  align::Parameters p{};
  p.threads = g_align_threads;
  p.refSequences = {tf}; p.querySequences = {qf}; p.mashmapPafFile = mp; p.pafOutputFile = op;
  p.target_padding = target_padding; p.query_padding = query_padding; p.wflign_max_len_minor = wflign_max_len_minor;
  p.min_identity = min_identity; p.min_alignment_length = min_alignment_length; p.min_block_identity = min_block_identity;
  p.disable_chain_patching = disable_chain_patching != 0; p.sam_format = sam_format != 0; p.emit_md_tag = emit_md_tag != 0; p.no_seq_in_sam = no_seq_in_sam != 0;
  p.wfa_patching_mismatch_score = 5; p.wfa_patching_gap_opening_score1 = 8; p.wfa_patching_gap_extension_score1 = 2; /* parse_args.hpp defaults */
  p.wfa_patching_gap_opening_score2 = 24; p.wfa_patching_gap_extension_score2 = 1;
  p.use_progress_bar = false; p.split = true;
  {
    align::Aligner a(p);
    a.compute();
  }
  std::ifstream in(op, std::ios::binary);
  std::string s((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if ((int64_t)s.size() > out_cap) return -1;
  memcpy(out, s.data(), s.size());
  return (int64_t)s.size();
}

/* Aligner::parseMashmapRow alone (computeAlignments.hpp:195-303). Returns 0, or 1 when the reference throws. */
int32_t ref_parse_mashmap_row(const char* line, uint64_t target_padding, uint64_t query_padding, int64_t* q_start, int64_t* q_end, int64_t* r_start,
                              int64_t* r_end, int32_t* strand, float* identity, int32_t* chain_id, int32_t* chain_length, int32_t* chain_pos) {
  try {
    align::MappingBoundaryRow r;
    align::Aligner::parseMashmapRow(line, r, target_padding, query_padding);
    *q_start = r.qStartPos; *q_end = r.qEndPos; *r_start = r.rStartPos; *r_end = r.rEndPos; *strand = r.strand; *identity = r.mashmap_estimated_identity;
    *chain_id = r.chain_id; *chain_length = r.chain_length; *chain_pos = r.chain_pos;
    return 0;
  } catch (const std::exception&) {
    return 1;
  }
}
}
