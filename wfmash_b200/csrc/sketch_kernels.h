// sketch_kernels.h — block-level fragment sketch shared by the sketch kernel (sketch.cu) and the L1 kernel
// (index_kernels.h). See sketch.cu for the reference anchors.
#pragma once
#include "wfb_rt.h"
#include "../../include/wfmash_b200.h"

WFB_DEV uint64_t sk_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
WFB_DEV uint64_t sk_fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

/* MurmurHash3_x64_128 (murmur3.h:226-303) of len <= 32 bytes given as four little-endian words
 * w[0..3] (unused high bytes zero), seed 42; returns the low 64 bits (commonFunc.hpp:173-182). */
WFB_DEV uint64_t sk_murmur3_lo64(const uint64_t w[4], int len) {
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = 42, h2 = 42;
  const int nblocks = len >> 4;
  int wi = 0;
  for (int i = 0; i < nblocks; ++i) {
    uint64_t k1 = w[wi], k2 = w[wi + 1];
    wi += 2;
    k1 *= c1; k1 = sk_rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = sk_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = sk_rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = sk_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const int rem = len & 15;
  if (rem > 8) { uint64_t k2 = w[wi + 1]; k2 *= c2; k2 = sk_rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  if (rem > 0) { uint64_t k1 = w[wi]; k1 *= c1; k1 = sk_rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = sk_fmix64(h1); h2 = sk_fmix64(h2);
  h1 += h2;
  return h1;
}

/* makeUpperCaseAndValidDNA (commonFunc.hpp:110-142): a..z -> A..Z, anything but A,C,G,T -> N */
WFB_DEV uint8_t sk_clean_base(uint8_t c) {
  if (c > 96 && c < 123) c -= 32;
  return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : (uint8_t)'N';
}
WFB_DEV uint8_t sk_comp(uint8_t c) { /* reverseComplement LUT (commonFunc.hpp:74-83) on cleaned bases */
  return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
}

#define SK_INVALID_VAL 0xFFFFFFFFu


/* One CTA sketches one fragment. smem layout: keys[npow2_max] u64 | vals[npow2_max] u32 | seq bytes.
 * Writes the fragment's minmers to o[0..count) (ascending hash) and returns count (uniform). Ends with a
 * barrier-free tail: callers must __syncthreads() before reusing smem or reading o[]. */
WFB_DEV int sk_sketch_block(unsigned char* smem, int* sh_warp, const uint8_t* seq_base, const wfb_frag_t fr, int ksize, int ssize,
                            int npow2_max, wfb_minmer_t* o) {
  uint64_t* keys = (uint64_t*)smem;
  uint32_t* vals = (uint32_t*)(smem + (size_t)npow2_max * 8);
  uint8_t* sseq = (uint8_t*)(smem + (size_t)npow2_max * 12);
  int result_count = 0;
  {
    const int len = fr.len;
    const int nk = len - ksize + 1;
    WFB_SYNC();
    if (nk <= 0) return 0;
    int N = 1;
    while (N < nk) N <<= 1;
    const uint8_t* src = seq_base + fr.seq_offset;
    for (int i = WFB_TID; i < len; i += WFB_NT) sseq[i] = sk_clean_base(wfb_ldg8(src + i));
    WFB_SYNC();
    /* hash every k-mer start (commonFunc.hpp:252-311) */
    for (int i = WFB_TID; i < N; i += WFB_NT) {
      uint64_t key = ~0ULL;
      uint32_t val = SK_INVALID_VAL;
      if (i < nk) {
        uint64_t wf[4] = {0, 0, 0, 0}, wr[4] = {0, 0, 0, 0};
        bool ambig = false;
        for (int j = 0; j < ksize; ++j) {
          const uint8_t b = sseq[i + j];
          ambig |= (b == 'N');
          wf[j >> 3] |= (uint64_t)b << ((j & 7) * 8);
          const int jr = ksize - 1 - j; /* revcomp byte index of this base */
          wr[jr >> 3] |= (uint64_t)sk_comp(b) << ((jr & 7) * 8);
        }
        if (!ambig) {
          const uint64_t hf = sk_murmur3_lo64(wf, ksize), hb = sk_murmur3_lo64(wr, ksize);
          if (hf != hb) {
            key = hf < hb ? hf : hb;
            val = ((uint32_t)i << 1) | (hf < hb ? 1u : 0u);
          }
        }
      }
      keys[i] = key;
      vals[i] = val;
    }
    WFB_SYNC();
    /* bitonic sort by (hash, position) */
    for (int k = 2; k <= N; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = WFB_TID; t < (N >> 1); t += WFB_NT) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int p = i | j;
          const bool up = (i & k) == 0;
          const uint64_t ka = keys[i], kb = keys[p];
          const uint32_t va = vals[i], vb = vals[p];
          const bool gt = (ka > kb) || (ka == kb && va > vb);
          if (gt == up) {
            keys[i] = kb; keys[p] = ka;
            vals[i] = vb; vals[p] = va;
          }
        }
        WFB_SYNC();
      }
    }
    /* rank run heads: each thread owns a contiguous chunk */
    const int per = (N + WFB_NT - 1) / WFB_NT;
    const int b0 = WFB_TID * per, b1 = min(N, b0 + per);
    int cnt = 0;
    for (int i = b0; i < b1; ++i) {
      const bool head = vals[i] != SK_INVALID_VAL && (i == 0 || keys[i] != keys[i - 1]);
      cnt += head ? 1 : 0;
    }
    int incl = cnt;
#ifndef WFB_EMU
    {
      const int lane = wfb_lane(), wid = WFB_TID >> 5;
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane == 31) sh_warp[wid] = incl;
      WFB_SYNC();
      if (wid == 0) {
        const int nw = (WFB_NT + 31) >> 5;
        int w = lane < nw ? sh_warp[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += y;
        }
        sh_warp[lane] = w;
      }
      WFB_SYNC();
      incl += wid ? sh_warp[wid - 1] : 0;
    }
    const int total = sh_warp[((WFB_NT + 31) >> 5) - 1];
#else
    (void)sh_warp;
    const int total = incl;
#endif
    int rank = incl - cnt; /* exclusive */
    for (int i = b0; i < b1 && rank < ssize; ++i) {
      if (vals[i] == SK_INVALID_VAL) break; /* invalid entries sort last */
      if (i != 0 && keys[i] == keys[i - 1]) continue;
      /* fold the run (:285,302-303,317): first pos, last pos, sign of the strand tally */
      const uint64_t h = keys[i];
      int tally = 0;
      int e = i;
      while (e < N && keys[e] == h && vals[e] != SK_INVALID_VAL) {
        tally += (vals[e] & 1u) ? 1 : -1;
        ++e;
      }
      wfb_minmer_t m;
      m.hash = h;
      m.wpos = (int64_t)(vals[i] >> 1);
      m.wpos_end = (int64_t)(vals[e - 1] >> 1);
      m.seqId = fr.seq_id;
      m.strand = (int16_t)(tally > 0 ? 1 : (tally == 0 ? 0 : -1));
      m.pad_ = 0;
      o[rank] = m;
      ++rank;
    }
    result_count = min(total, ssize);
  }
  return result_count;
}
