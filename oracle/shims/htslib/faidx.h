/*
 * TEST INFRASTRUCTURE ONLY — stand-in for <htslib/faidx.h> (htslib is not installed in this image and is not vendored in
 * the reference tree) so that the reference's UNMODIFIED src/map/include/winSketch.hpp (through src/common/seqiter.hpp)
 * compiles in place. It implements exactly the calls that header makes — fai_load, fai_destroy, fai_fetch,
 * faidx_fetch_seq64, faidx_nseq, faidx_iseq — over an UNCOMPRESSED FASTA with its samtools-style .fai next to it
 * (name, length, offset, bases per line, bytes per line), which is what oracle/ref_sketch_driver.cpp writes.
 */
#ifndef WFB_SHIM_HTSLIB_FAIDX_H
#define WFB_SHIM_HTSLIB_FAIDX_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

typedef int64_t hts_pos_t;
struct faidx_t {
  struct Rec { std::string name; int64_t len, offset, line_bases, line_bytes; };
  std::string path;
  std::vector<Rec> recs;
  std::unordered_map<std::string, size_t> by_name;
};

static inline faidx_t* fai_load(const char* fn) {
  std::ifstream in(std::string(fn) + ".fai");
  if (!in.good()) return nullptr;
  faidx_t* f = new faidx_t;
  f->path = fn;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    faidx_t::Rec r;
    if (!(ss >> r.name >> r.len >> r.offset >> r.line_bases >> r.line_bytes)) continue;
    f->by_name[r.name] = f->recs.size();
    f->recs.push_back(r);
  }
  return f;
}
static inline void fai_destroy(faidx_t* f) { delete f; }
static inline int faidx_nseq(const faidx_t* f) { return (int)f->recs.size(); }
static inline const char* faidx_iseq(const faidx_t* f, int i) { return f->recs[(size_t)i].name.c_str(); }

/* [beg, end] 0-based inclusive, clamped to the sequence like htslib does; malloc'ed, NUL-terminated */
static inline char* faidx_fetch_seq64(const faidx_t* f, const char* name, hts_pos_t beg, hts_pos_t end, hts_pos_t* len) {
  auto it = f->by_name.find(name);
  if (it == f->by_name.end()) { *len = -2; return nullptr; }
  const faidx_t::Rec& r = f->recs[it->second];
  if (end < beg) beg = end;
  if (beg < 0) beg = 0; else if (r.len <= beg) beg = r.len;
  if (end < 0) end = 0; else if (r.len <= end) end = r.len - 1;
  const int64_t n = end >= beg && r.len > 0 && beg < r.len ? end - beg + 1 : 0;
  char* out = (char*)malloc((size_t)n + 1);
  FILE* fp = fopen(f->path.c_str(), "rb");
  if (!fp || !out) { if (fp) fclose(fp); free(out); *len = -1; return nullptr; }
  int64_t got = 0;
  for (int64_t p = beg; got < n;) {
    const int64_t line = p / r.line_bases, col = p % r.line_bases;
    const int64_t take = (r.line_bases - col) < (n - got) ? (r.line_bases - col) : (n - got);
    fseek(fp, (long)(r.offset + line * r.line_bytes + col), SEEK_SET);
    if ((int64_t)fread(out + got, 1, (size_t)take, fp) != take) break;
    got += take; p += take;
  }
  fclose(fp);
  out[got] = 0;
  *len = got;
  return out;
}
static inline char* fai_fetch(const faidx_t* f, const char* reg, int* len) {
  hts_pos_t l = 0;
  char* s = faidx_fetch_seq64(f, reg, 0, INT64_MAX - 1, &l); /* whole sequence: the reference passes bare names */
  *len = (int)l;
  return s;
}
#endif
