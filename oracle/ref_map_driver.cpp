/*
 * TEST INFRASTRUCTURE ONLY. Thin driver over the UNMODIFIED reference header
 * /root/reference/src/map/include/commonFunc.hpp (compiled in place by oracle/Makefile into
 * oracle/_ref/libmapref.so). Exposes the reference's sketchSequence / addMinmers / getHash to
 * ctypes so that the C restatement (map_oracle.c) and the CUDA kernels can be checked against the
 * real thing. Nothing in the product path links or loads this file.
 */
#include <vector>
#include <string>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <fcntl.h>
#include <unistd.h>
#include "map/include/base_types.hpp"
#include "map/include/commonFunc.hpp"

extern "C" {

struct ref_minmer_t { uint64_t hash; int64_t wpos; int64_t wpos_end; int32_t seqId; int16_t strand; int16_t pad_; };

uint64_t ref_kmer_hash(const char* kmer, int k) { return skch::CommonFunc::getHash(kmer, k); }

/* commonFunc.hpp:217 — note: upper-cases / N-masks seq IN PLACE like the reference. */
int ref_sketch_fragment(char* seq, int64_t len, int k, int s, int32_t seqId, ref_minmer_t* out, int cap) {
  std::vector<skch::MinmerInfo> v;
  skch::CommonFunc::sketchSequence(v, seq, len, k, 4, s, seqId);
  int n = 0;
  for (auto& m : v) {
    if (n >= cap) break;
    out[n++] = ref_minmer_t{m.hash, m.wpos, m.wpos_end, m.seqId, (int16_t)m.strand, 0};
  }
  return (int)v.size();
}

/* commonFunc.hpp:440 */
int64_t ref_add_minmers(char* seq, int64_t len, int k, int w, int s, int32_t seqId, ref_minmer_t* out, int64_t cap) {
  std::vector<skch::MinmerInfo> v;
  /* one shared meter like the reference's Sketch::build (winSketch.hpp:188-204), silenced: its reporter thread
   * (progress.hpp:32-103) is told it has finished and given time to leave before any progress is counted, so nothing
   * is printed from a foreign thread inside the test process; increment() itself stays a plain atomic add */
  static progress_meter::ProgressMeter* pm = [] {
    auto* m = new progress_meter::ProgressMeter((uint64_t)1 << 60, "", false);
    m->is_finished.store(true);
    std::this_thread::sleep_for(std::chrono::milliseconds(300));
    return m;
  }();
  skch::CommonFunc::addMinmers(v, seq, len, k, w, 4, s, seqId, pm);
  int64_t n = 0;
  for (auto& m : v) {
    if (n >= cap) break;
    out[n++] = ref_minmer_t{m.hash, m.wpos, m.wpos_end, m.seqId, (int16_t)m.strand, 0};
  }
  return (int64_t)v.size();
}
}
