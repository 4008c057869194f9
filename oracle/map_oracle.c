/*
 * map_oracle.c — plain-C restatement of the reference's MashMap 3.5 sketching primitives
 * (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Reference anchors (under /root/reference/src/):
 *   getHash               map/include/commonFunc.hpp:173-182, common/murmur3.h:226-303
 *   reverseComplement     map/include/commonFunc.hpp:74-83
 *   makeUpperCase...      map/include/commonFunc.hpp:110-142
 *   sketchSequence        map/include/commonFunc.hpp:217-323
 *   addMinmers            map/include/commonFunc.hpp:439-708
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

/* MurmurHash3_x64_128 (murmur3.h:226-303), seed 42, low 64 bits (commonFunc.hpp:173-182). */
uint64_t orc_kmer_hash(const char* kmer, int len) {
  const uint8_t* data = (const uint8_t*)kmer;
  const int nblocks = len / 16;
  uint64_t h1 = 42, h2 = 42;
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  for (int i = 0; i < nblocks; i++) {
    uint64_t k1, k2;
    memcpy(&k1, data + 16 * i, 8);
    memcpy(&k2, data + 16 * i + 8, 8);
    k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const uint8_t* tail = data + nblocks * 16;
  uint64_t k1 = 0, k2 = 0;
  const int rem = len & 15;
  for (int i = rem - 1; i >= 8; --i) k2 ^= (uint64_t)tail[i] << (8 * (i - 8));
  if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  for (int i = (rem < 8 ? rem : 8) - 1; i >= 0; --i) k1 ^= (uint64_t)tail[i] << (8 * i);
  if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  h1 += h2;
  return h1;
}

static inline char comp_base(char c) {
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}
static void revcomp(const char* src, char* dst, int n) {
  for (int i = 0; i < n; ++i) dst[n - 1 - i] = comp_base(src[i]);
}

static int cmp_minmer_hash(const void* a, const void* b) {
  const orc_minmer_t* x = (const orc_minmer_t*)a; const orc_minmer_t* y = (const orc_minmer_t*)b;
  return x->hash < y->hash ? -1 : (x->hash > y->hash);
}

/* sketchSequence closed form: the min(s, #distinct) smallest distinct canonical hashes,
 * ascending; wpos = first occurrence, wpos_end = last occurrence, strand = sign of the +-1 tally.
 * (The reference's heap / hash-map dance at commonFunc.hpp:277-305 admits exactly these.) */
int orc_sketch_fragment(const char* seq, int len, int k, int s, int32_t seqId, orc_minmer_t* out) {
  const int nk = len - k + 1;
  if (nk <= 0) return 0;
  orc_minmer_t* all = (orc_minmer_t*)malloc((size_t)nk * sizeof(orc_minmer_t));
  char* rc = (char*)malloc((size_t)k);
  int n = 0;
  for (int i = 0; i < nk; ++i) {
    int ambig = 0;
    for (int j = 0; j < k; ++j) if (seq[i + j] == 'N') { ambig = 1; break; }
    revcomp(seq + i, rc, k);
    const uint64_t hf = orc_kmer_hash(seq + i, k), hb = orc_kmer_hash(rc, k);
    if (hf == hb || ambig) continue;
    all[n].hash = hf < hb ? hf : hb;
    all[n].wpos = i; all[n].wpos_end = i; all[n].seqId = seqId;
    all[n].strand = hf < hb ? 1 : -1; all[n].pad_ = 0;
    ++n;
  }
  /* stable by construction: qsort on hash only, then fold runs (positions folded via min/max) */
  qsort(all, (size_t)n, sizeof(orc_minmer_t), cmp_minmer_hash);
  int m = 0;
  for (int i = 0; i < n && m < s;) {
    int j = i; int64_t lo = all[i].wpos, hi = all[i].wpos; int tally = 0;
    while (j < n && all[j].hash == all[i].hash) {
      if (all[j].wpos < lo) lo = all[j].wpos;
      if (all[j].wpos > hi) hi = all[j].wpos;
      tally += all[j].strand; ++j;
    }
    out[m].hash = all[i].hash; out[m].wpos = lo; out[m].wpos_end = hi; out[m].seqId = seqId;
    out[m].strand = tally > 0 ? 1 : (tally == 0 ? 0 : -1); out[m].pad_ = 0;
    ++m; i = j;
  }
  free(all); free(rc);
  return m;
}
