// ani_kernels.h — ANI auto-identity sketching: the third consumer of the canonical k-mer hash.
//
// Reference: skch::Stat::estimate_identity_for_groups (src/map/include/map_stats.hpp:325-822) streams every sequence
// through a max-heap of 4096 hashes (StreamingMinHash::add_unsafe, streamingMinHash.hpp:89-99: duplicates are kept) and
// merges the per-sequence heaps into one heap per PanSN group. Value-wise that is order independent:
//     group sketch = the `s` smallest elements of the MULTISET of canonical k-mer hashes of the group's sequences.
// On the GPU the selection becomes a threshold filter: hashes are uniform on [0, 2^64), so the s smallest of N lie below
// T = (8 s / N) * 2^64 with overwhelming probability; one streaming kernel hashes every position and appends the few
// hashes <= T to the group's candidate list, a segmented radix sort orders the candidates, the first s are the sketch.
// The host checks count >= s (else T is lifted to "everything") and count <= capacity (else the capacity is grown to the
// exact count: heavy duplication of one small hash, e.g. satellite repeats) and re-runs; both are rare.
//
// Kernel shape: one CTA per tile of ANI_TILE k-mer start positions of one sequence; the tile (+ k-1 halo) is brought in
// with 128-bit loads and cleaned (upper-case, non-ACGT -> N) in shared memory; a thread owns ANI_RUN consecutive
// positions and ROLLS the forward and reverse-complement k-mers through four 64-bit byte-shift registers each, so a
// position costs two Murmur3 evaluations and a handful of shifts instead of 2k byte gathers.
// Roofline: 1 B / base of HBM traffic against ~2 x 45 64-bit multiply/rotate steps per base: integer-ALU bound.
#pragma once
#include "sketch_kernels.h"

#define ANI_THREADS 256
#define ANI_RUN 33 /* odd on purpose: lane l starts at byte 33 l, so the 32 lanes of a warp hit 32 different shared-memory banks */
#define ANI_TILE (ANI_RUN * ANI_THREADS)
#define ANI_SMEM_BYTES (ANI_TILE + 32 + 48) /* tile + halo (k <= 32) + alignment slack of the 16-byte loads */

/* sk_murmur3_lo64 with every word index known at compile time (a run-time index into the four k-mer words would push them
 * out of registers into local memory) */
WFB_DEV uint64_t ani_murmur3_lo64(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3, int len) {
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = 42, h2 = 42;
  const int nblocks = len >> 4;
  if (nblocks >= 1) {
    uint64_t k1 = w0, k2 = w1;
    k1 *= c1; k1 = sk_rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = sk_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = sk_rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = sk_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  if (nblocks >= 2) {
    uint64_t k1 = w2, k2 = w3;
    k1 *= c1; k1 = sk_rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = sk_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = sk_rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = sk_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const uint64_t t1 = nblocks == 0 ? w0 : nblocks == 1 ? w2 : 0, t2 = nblocks == 0 ? w1 : nblocks == 1 ? w3 : 0;
  const int rem = len & 15;
  if (rem > 8) { uint64_t k2 = t2; k2 *= c2; k2 = sk_rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  if (rem > 0) { uint64_t k1 = t1; k1 *= c1; k1 = sk_rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = sk_fmix64(h1); h2 = sk_fmix64(h2);
  h1 += h2;
  return h1;
}

struct AniTile {
  int64_t seq_off;   /* sequence start inside the blob                               */
  int64_t seq_len;
  int64_t start;     /* first k-mer start position of the tile                       */
  int32_t npos;      /* k-mer start positions in the tile                            */
  int32_t group;     /* dense group index                                            */
  int32_t head_bad;  /* a non-ACGT base among the first min(k, len) bases (see below) */
  int32_t pad_;
};

/* One tile. The reference's validity rule (map_stats.hpp:565-613): ambig = k when the k-mer's LAST base is not ACGT,
 * decremented once per position, k-mer hashed only while ambig == 0 — i.e. a bad base at p >= k-1 invalidates the k-mers
 * p-k+1..p. Bad bases at p < k-1 are only seen by the initial scan of the first k bases, which sets ambig = k whatever
 * their position: then k-mers 0..k-1 are ALL invalid (head_bad). */
WFB_DEV void ani_tile(unsigned char* smem, const uint8_t* blob, const AniTile t, int k, const uint64_t* thresholds, const unsigned long long* cand_base,
                      const unsigned long long* cand_cap, unsigned long long* cand_count, uint64_t* cand, unsigned long long* n_valid) {
  uint8_t* s = (uint8_t*)smem;
  const int64_t g0 = t.seq_off + t.start;           /* first byte of the tile in the blob */
  const int nbytes = t.npos + k - 1;
  const int64_t a0 = g0 & ~(int64_t)15;             /* 16-byte aligned start              */
  const int lead = (int)(g0 - a0);
  const int nvec = (lead + nbytes + 15) >> 4;
#ifndef WFB_EMU
  for (int v = WFB_TID; v < nvec; v += WFB_NT) {
    uint4 x = __ldg((const uint4*)(blob + a0) + v);
    uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t o = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) o |= (uint32_t)sk_clean_base((uint8_t)(w[j] >> (8 * b))) << (8 * b);
      w[j] = o;
    }
    ((uint4*)s)[v] = make_uint4(w[0], w[1], w[2], w[3]);
  }
#else
  for (int i = 0; i < nvec * 16; ++i) s[i] = sk_clean_base(blob[a0 + i]);
#endif
  WFB_SYNC();
  const uint8_t* q = s + lead;                      /* q[i] = cleaned base at tile-local position i */
  const uint64_t T = thresholds[t.group];
  const unsigned long long base = cand_base[t.group], cap = cand_cap[t.group];
  unsigned valid_here = 0;
  for (int r = WFB_TID; r * ANI_RUN < t.npos; r += WFB_NT) {
    const int p0 = r * ANI_RUN;
    const int p1 = min(t.npos, p0 + ANI_RUN);
    /* the two k-mers as eight scalar 64-bit words (arrays indexed at run time would live in local memory) */
    uint64_t f0 = 0, f1 = 0, f2 = 0, f3 = 0, r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    int nbad = 0;
#define ANI_INIT_WORD(FW, RW, W)                                                          \
    _Pragma("unroll") for (int jj = 0; jj < 8; ++jj) {                                    \
      const int j = (W) * 8 + jj;                                                         \
      if (j < k) {                                                                        \
        const uint8_t b = q[p0 + j], c = q[p0 + k - 1 - j];                               \
        nbad += b == 'N';                                                                 \
        FW |= (uint64_t)b << (jj * 8);                                                    \
        RW |= (uint64_t)sk_comp(c) << (jj * 8);                                           \
      }                                                                                   \
    }
    ANI_INIT_WORD(f0, r0, 0)
    ANI_INIT_WORD(f1, r1, 1)
    ANI_INIT_WORD(f2, r2, 2)
    ANI_INIT_WORD(f3, r3, 3)
#undef ANI_INIT_WORD
    const int top = k - 1, topw = top >> 3, clrw = k >> 3;
    const uint64_t clr = k < 32 ? ~(0xFFull << ((k & 7) * 8)) : ~0ULL; /* clears the base that leaves the reverse k-mer */
    for (int p = p0; p < p1; ++p) {
      const bool head = t.head_bad && (t.start + p) <= (int64_t)(k - 1);
      if (nbad == 0 && !head) {
        const uint64_t hf = ani_murmur3_lo64(f0, f1, f2, f3, k), hb = ani_murmur3_lo64(r0, r1, r2, r3, k);
        if (hf != hb) {
          ++valid_here;
          const uint64_t h = hf < hb ? hf : hb;
          if (h <= T) {
            const unsigned long long idx = atomicAdd_compat(&cand_count[t.group], 1ULL);
            if (idx < cap) cand[base + idx] = h;
          }
        }
      }
      if (p + 1 < p1) { /* roll both k-mers one base to the right */
        const uint8_t out = q[p], in = q[p + k];
        nbad += (in == 'N') - (out == 'N');
        const uint64_t ins = (uint64_t)in << ((top & 7) * 8);
        f0 = (f0 >> 8) | (f1 << 56);
        f1 = (f1 >> 8) | (f2 << 56);
        f2 = (f2 >> 8) | (f3 << 56);
        f3 = f3 >> 8;
        f0 |= topw == 0 ? ins : 0; f1 |= topw == 1 ? ins : 0; f2 |= topw == 2 ? ins : 0; f3 |= topw == 3 ? ins : 0;
        r3 = (r3 << 8) | (r2 >> 56);
        r2 = (r2 << 8) | (r1 >> 56);
        r1 = (r1 << 8) | (r0 >> 56);
        r0 = (r0 << 8) | (uint64_t)sk_comp(in);
        r0 &= clrw == 0 ? clr : ~0ULL; r1 &= clrw == 1 ? clr : ~0ULL; r2 &= clrw == 2 ? clr : ~0ULL; r3 &= clrw == 3 ? clr : ~0ULL;
      }
    }
  }
  { /* one counter update per warp, not per thread: 6 M same-address atomics per 200 Mbp were 10 % of the stall samples */
    const unsigned warp_valid = wfb_warp_add(valid_here);
    if (wfb_lane() == 0 && warp_valid) atomicAdd_compat(n_valid, (unsigned long long)warp_valid);
  }
  WFB_SYNC();
}

WFB_KERNEL_LB(ani_hash_kernel, ANI_THREADS, 4, const uint8_t* blob, const AniTile* tiles, int ntiles, int k, const uint64_t* thresholds,
              const unsigned long long* cand_base, const unsigned long long* cand_cap, unsigned long long* cand_count, uint64_t* cand,
              unsigned long long* n_valid
#ifdef WFB_EMU
              , unsigned char* smem_emu
#endif
) {
  WFB_KERNEL_PROLOGUE
#ifndef WFB_EMU
  __shared__ __align__(16) unsigned char smem[ANI_SMEM_BYTES];
#else
  unsigned char* smem = smem_emu;
#endif
  for (int i = bid; i < ntiles; i += nblocks) ani_tile(smem, blob, tiles[i], k, thresholds, cand_base, cand_cap, cand_count, cand, n_valid);
}

/* sketches[g * s .. ) = the first min(count, s) sorted candidates of group g */
WFB_KERNEL(ani_gather_kernel, const uint64_t* sorted, const unsigned long long* cand_base, const unsigned long long* cand_count, int n_groups, int s,
           uint64_t* sketches) {
  WFB_KERNEL_PROLOGUE
  for (int g = bid; g < n_groups; g += nblocks) {
    const unsigned long long n = cand_count[g] < (unsigned long long)s ? cand_count[g] : (unsigned long long)s;
    for (int i = WFB_TID; i < (int)n; i += WFB_NT) sketches[(size_t)g * s + i] = sorted[cand_base[g] + i];
  }
}
