/*
 * TEST INFRASTRUCTURE ONLY — stand-in for the reference's src/common/faigz.h (a thread-safe FASTA reader built on htslib's
 * faidx / bgzf / khash internals; htslib is not installed here) with the handful of calls the reference's
 * src/map/include/map_stats.hpp makes, implemented over oracle/shims/htslib/faidx.h (uncompressed FASTA + .fai).
 * Shadowing this one header lets map_stats.hpp itself compile UNMODIFIED.
 */
#ifndef WFB_SHIM_FAIGZ_H
#define WFB_SHIM_FAIGZ_H
#include <htslib/faidx.h>

enum fai_format_options { FAI_NONE, FAI_FASTA, FAI_FASTQ };
#define FAI_CREATE 0x01
typedef faidx_t faidx_meta_t;
struct faidx_reader_t { const faidx_meta_t* meta; };

static inline faidx_meta_t* faidx_meta_load(const char* fn, enum fai_format_options, int) { return fai_load(fn); }
static inline void faidx_meta_destroy(faidx_meta_t* m) { fai_destroy(m); }
static inline int faidx_meta_nseq(const faidx_meta_t* m) { return faidx_nseq(m); }
static inline const char* faidx_meta_iseq(const faidx_meta_t* m, int i) { return faidx_iseq(m, i); }
static inline hts_pos_t faidx_meta_seq_len(const faidx_meta_t* m, const char* name) {
  auto it = m->by_name.find(name);
  return it == m->by_name.end() ? -1 : m->recs[it->second].len;
}
static inline faidx_reader_t* faidx_reader_create(const faidx_meta_t* m) { return new faidx_reader_t{m}; }
static inline void faidx_reader_destroy(faidx_reader_t* r) { delete r; }
static inline char* faidx_reader_fetch_seq(faidx_reader_t* r, const char* name, hts_pos_t beg, hts_pos_t end, hts_pos_t* len) {
  return faidx_fetch_seq64(r->meta, name, beg, end, len);
}
#endif
