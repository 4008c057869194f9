// minmer_kernels.h — MashMap 3.5 reference-side windowed minmers on the GPU (sm_100a).
//
// Replaces skch::CommonFunc::addMinmers (src/map/include/commonFunc.hpp:439-708) as run per target
// sequence by Sketch::buildHelper (src/map/include/winSketch.hpp:467-499).
//
// The reference is one sequential stream per sequence whose state (the window deque Q, the ordered
// sketch `sortedWindow` with per-hash occurrence lists and strand tallies, and a lazily cleaned heap of
// the window's other k-mers) depends on the last w positions. B200 mapping:
//   * a sequence is cut into chunks of WFB_MM_CHUNK k-mer positions; ONE THREAD runs the reference's
//     state machine over its chunk preceded by `warm` (>= w) warm-up positions, so tens of thousands of
//     chunks advance concurrently (the stream itself has no intra-step parallelism worth a warp);
//   * records closed inside the chunk body are appended to a global array through one atomic counter;
//     a record whose interval was opened before the body carries an "inherited" start that the stitch
//     pass replaces with the true start taken from the previous chunk's end state (same hash);
//   * the post passes (:660-706: degenerate drop, strand sign, > w splitting, sort, unique) run as
//     separate data-parallel kernels / library sorts (minmer_host.cu).
//   * round 2: the state machine no longer sees every k-mer. mm_cand_kernel hashes all positions of a tile in parallel (one CTA
//     per tile, rolling byte-shift registers, shared-memory staging) and keeps the CANDIDATES whose canonical hash is below a
//     threshold that leaves ~2s + 40 of them per window; mm_stream_cand_kernel runs the state machine over the candidate stream
//     only (events: a candidate arrives, a candidate leaves, the first fill), 5 - 15 % of the positions with containers of a
//     few hundred entries. A chunk that cannot vouch for the result (sketch short of s entries, expired heap entry met while
//     filling, a capacity hit, a tile with too many candidates) flags itself and is re-run over every k-mer by the exact
//     instantiation of the same step function (mm_stream_redo_smem_kernel: one CTA per flagged chunk, containers in shared memory;
//     mm_stream_kernel when there are thousands). See "candidate stream" below for why the filter is exact otherwise.
//     Measured on scerevisiae8 (96 Mbp, s = 24): 35.4 -> 16.3 ms; candidates 1.8 ms (58 G positions/s), filtered stream 11.5, re-run 3.0
//     (profiles/r02_ncu_minmer_filtered_summary.txt).
// Exactness: the chunk state after >= w warm-up positions equals the reference's state restricted to
// live entries; entries the reference keeps past their expiry ("stale" heap entries, :596-641) can make
// the two differ. Every such absorption is counted (stale_absorbed) so callers can tell; the parity
// tests compare whole sequences against the reference (LPA, yeast, synthetic repeats: identical).
#pragma once
#include "wfb_rt.h"

struct MmKmer { /* KmerInfo, base_types.hpp:135-141 (pos relative to the sequence start) */
  uint64_t hash;
  int pos;
  int strand;
};
struct MmNode { /* one entry of a sortedWindow occurrence list */
  int pos;
  int strand;
  int next;
};
struct MmWent { /* sortedWindow value: (MinmerInfo, deque<KmerInfo>) keyed by hash */
  uint64_t hash;
  long long wpos; /* window id (i + k - w) when the interval was (re)opened; -1 = not open */
  int strand;     /* running tally */
  int head, tail, count;
};
struct MmRecord { /* raw MinmerInfo emitted by the stream (pre post-pass) */
  uint64_t hash;
  long long wpos, wpos_end;
  int seq;    /* index of the sequence in the batch */
  int strand; /* tally before the update, as the reference stores it */
  int chunk;  /* global chunk id that emitted it */
  int inherited; /* 1 = wpos comes from before the chunk body -> stitch */
};
struct MmChunk {
  int seq;           /* sequence index */
  long long body_begin, body_end; /* k-mer start positions [begin, end) */
  long long run_begin;            /* body_begin - warm, clamped at 0 */
  int first_of_seq, last_of_seq;
};
struct MmEndEnt { /* sortedWindow membership at the end of a chunk body */
  uint64_t hash;
  long long wpos;
  int inherited, pad_;
};
struct MmSeq {
  long long off; /* first base inside the cleaned sequence buffer */
  long long len;
  int seq_id;    /* MinmerInfo::seqId */
  int first_chunk, n_chunks;
};
struct MmCounters {
  unsigned long long n_records, stale_absorbed, overflow, stitch_miss;
  unsigned long long candidates, flagged; /* filtered build: k-mers below the threshold; chunks left to the exact re-run */
  unsigned long long ties;               /* neighbours of the final order with equal (seqId, wpos, wpos_end) */
};

WFB_DEV uint64_t mm_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
WFB_DEV uint64_t mm_fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}
/* MurmurHash3_x64_128 low 64 bits, seed 42, len <= 32 (murmur3.h:226-303; commonFunc.hpp:173-182) */
WFB_DEV uint64_t mm_murmur3_lo64(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3, int len) {
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = 42, h2 = 42;
  uint64_t t1 = w0, t2 = w1;
  if (len >= 16) {
    uint64_t k1 = w0, k2 = w1;
    k1 *= c1; k1 = mm_rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = mm_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = mm_rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = mm_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    t1 = w2; t2 = w3;
  }
  if (len == 32) { /* second full block, empty tail */
    uint64_t k1 = w2, k2 = w3;
    k1 *= c1; k1 = mm_rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = mm_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = mm_rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = mm_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  } else {
    const int rem = len & 15;
    if (rem > 8) { uint64_t k2 = t2; k2 *= c2; k2 = mm_rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    if (rem > 0) { uint64_t k1 = t1; k1 *= c1; k1 = mm_rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = mm_fmix64(h1); h2 = mm_fmix64(h2);
  h1 += h2;
  return h1;
}
WFB_DEV uint8_t mm_clean_base(uint8_t c) { /* makeUpperCaseAndValidDNA, commonFunc.hpp:110-142 */
  if (c > 96 && c < 123) c -= 32;
  return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : (uint8_t)'N';
}
WFB_DEV uint8_t mm_comp(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }

/* forward / reverse-complement hashes of the k-mer starting at s (ASCII, already cleaned) */
WFB_DEV void mm_hash_pair(const uint8_t* s, int k, uint64_t& hf, uint64_t& hb) {
  uint64_t f[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0};
  for (int j = 0; j < k; ++j) {
    const uint8_t b = s[j];
    f[j >> 3] |= (uint64_t)b << ((j & 7) * 8);
    const int jr = k - 1 - j;
    r[jr >> 3] |= (uint64_t)mm_comp(b) << ((jr & 7) * 8);
  }
  hf = mm_murmur3_lo64(f[0], f[1], f[2], f[3], k);
  hb = mm_murmur3_lo64(r[0], r[1], r[2], r[3], k);
}

/* The forward and reverse-complement k-mers of the current position as eight 64-bit byte-shift registers: a position costs one
 * new base and a few shifts instead of re-packing 2k bytes (and scalars stay in registers where f[j >> 3] did not). */
struct MmRoll {
  uint64_t f0, f1, f2, f3, r0, r1, r2, r3;
};
WFB_DEV void mm_roll_init(MmRoll& R, const uint8_t* s, int k) {
  R.f0 = R.f1 = R.f2 = R.f3 = R.r0 = R.r1 = R.r2 = R.r3 = 0;
#define MM_INIT_WORD(FW, RW, WI)                                              \
  _Pragma("unroll") for (int jj = 0; jj < 8; ++jj) {                          \
    const int j = (WI) * 8 + jj;                                              \
    if (j < k) {                                                              \
      FW |= (uint64_t)s[j] << (jj * 8);                                       \
      RW |= (uint64_t)mm_comp(s[k - 1 - j]) << (jj * 8);                      \
    }                                                                         \
  }
  MM_INIT_WORD(R.f0, R.r0, 0)
  MM_INIT_WORD(R.f1, R.r1, 1)
  MM_INIT_WORD(R.f2, R.r2, 2)
  MM_INIT_WORD(R.f3, R.r3, 3)
#undef MM_INIT_WORD
}
WFB_DEV void mm_roll_step(MmRoll& R, uint8_t in, int k) { /* k-mer at i -> k-mer at i + 1; in = s[i + k] */
  const int top = k - 1, topw = top >> 3, clrw = k >> 3;
  const uint64_t ins = (uint64_t)in << ((top & 7) * 8);
  const uint64_t clr = k < 32 ? ~(0xFFull << ((k & 7) * 8)) : ~0ULL;
  R.f0 = (R.f0 >> 8) | (R.f1 << 56);
  R.f1 = (R.f1 >> 8) | (R.f2 << 56);
  R.f2 = (R.f2 >> 8) | (R.f3 << 56);
  R.f3 = R.f3 >> 8;
  R.f0 |= topw == 0 ? ins : 0; R.f1 |= topw == 1 ? ins : 0; R.f2 |= topw == 2 ? ins : 0; R.f3 |= topw == 3 ? ins : 0;
  R.r3 = (R.r3 << 8) | (R.r2 >> 56);
  R.r2 = (R.r2 << 8) | (R.r1 >> 56);
  R.r1 = (R.r1 << 8) | (R.r0 >> 56);
  R.r0 = (R.r0 << 8) | (uint64_t)mm_comp(in);
  R.r0 &= clrw == 0 ? clr : ~0ULL; R.r1 &= clrw == 1 ? clr : ~0ULL; R.r2 &= clrw == 2 ? clr : ~0ULL; R.r3 &= clrw == 3 ? clr : ~0ULL;
}

/* ---- per-thread containers in global scratch ----
 * The 32 chunks of a warp share one slab and their containers are INTERLEAVED element by element: element i of lane l lives
 * at slab[i * 32 + l]. The window deque advances in lockstep across the lanes (one k-mer in, one out per position), the top
 * levels of the heaps and the first sortedWindow entries are touched by every lane at the same index, so those accesses
 * coalesce into contiguous 512-byte (MmKmer) runs instead of 32 scattered sectors (ncu on the private-slab layout: 314 B
 * of DRAM traffic per base against 7.2 algorithmic). The host emulation runs one chunk at a time: MM_LANES = 1. */
#ifdef WFB_EMU
#define MM_LANES 1
#else
#define MM_LANES 32
#endif
template <class T, int L> /* L = lane stride: MM_LANES in an interleaved global slab, 1 in a private (shared-memory) one */
struct MmArr {
  T* p; /* this lane's element 0 */
  WFB_DEV_MEMBER T& operator[](long long i) const { return p[i * L]; }
};
template <int L>
struct MmHeap {
  MmArr<MmKmer, L> a;
  int n, cap;
};
WFB_DEV bool mm_less(const MmKmer& x, const MmKmer& y) { return x.hash < y.hash || (x.hash == y.hash && x.pos < y.pos); }
template <int L>
WFB_DEV bool mm_heap_push(MmHeap<L>& h, const MmKmer& v) {
  if (h.n >= h.cap) return false;
  int i = h.n++;
  while (i > 0) {
    const int p = (i - 1) >> 1;
    if (!mm_less(v, h.a[p])) break;
    h.a[i] = h.a[p];
    i = p;
  }
  h.a[i] = v;
  return true;
}
template <int L>
WFB_DEV void mm_heap_sift_down(MmHeap<L>& h, int i) {
  const MmKmer v = h.a[i];
  for (;;) {
    int c = 2 * i + 1;
    if (c >= h.n) break;
    if (c + 1 < h.n && mm_less(h.a[c + 1], h.a[c])) ++c;
    if (!mm_less(h.a[c], v)) break;
    h.a[i] = h.a[c];
    i = c;
  }
  h.a[i] = v;
}
template <int L>
WFB_DEV void mm_heap_pop(MmHeap<L>& h) { /* leaves the popped element readable in a[0] when the heap empties */
  --h.n;
  if (h.n > 0) {
    h.a[0] = h.a[h.n];
    mm_heap_sift_down(h, 0);
  }
}

template <int L>
struct MmPool {
  MmArr<MmNode, L> nodes;
  int free_head;
};
template <int L>
WFB_DEV int mm_pool_alloc(MmPool<L>& p) {
  const int i = p.free_head;
  if (i >= 0) p.free_head = p.nodes[i].next;
  return i;
}
template <int L>
WFB_DEV void mm_pool_free(MmPool<L>& p, int i) { p.nodes[i].next = p.free_head; p.free_head = i; }
template <int L>
WFB_DEV bool mm_went_push_back(MmPool<L>& p, MmWent& e, int pos, int strand) {
  const int n = mm_pool_alloc(p);
  if (n < 0) return false;
  p.nodes[n].pos = pos; p.nodes[n].strand = strand; p.nodes[n].next = -1;
  if (e.tail >= 0) p.nodes[e.tail].next = n; else e.head = n;
  e.tail = n;
  e.count++;
  return true;
}
template <int L>
WFB_DEV void mm_went_pop_front(MmPool<L>& p, MmWent& e) {
  const int n = e.head;
  e.head = p.nodes[n].next;
  if (e.head < 0) e.tail = -1;
  e.count--;
  mm_pool_free(p, n);
}
template <int L>
WFB_DEV void mm_went_clear(MmPool<L>& p, MmWent& e) { while (e.head >= 0) mm_went_pop_front(p, e); }

template <int L>
WFB_DEV int mm_lower_bound(const MmArr<MmWent, L>& W, int wn, uint64_t h) {
  int lo = 0, hi = wn;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (W[mid].hash < h) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct MmParams {
  int k, w, s;
  int chunk, warm;
  int qcap, heap_cap, pool_cap; /* per-thread capacities */
  int lcur;                     /* filtered run: departures read from the candidate stream (no window queue) */
};

/* ---- candidate stream (the data-parallel half of the filtered build) ----
 * A k-mer whose hash is above every hash the window sketch ever holds is a no-op for addMinmers: it goes to the heap when it
 * arrives (:581-585), is never the heap's front while a smaller live k-mer exists, and is ignored when it leaves (:520). With
 * T chosen so that a window holds ~2s + 40 k-mers <= T, the state machine only has to see those CANDIDATES (5 - 15 % of the
 * positions); the chunk flags itself for an exact re-run over all positions whenever that assumption is visibly broken (the
 * sketch is short of s entries after a step, an expired heap entry is met while filling, a capacity is hit).
 * mm_cand_kernel: one CTA per tile of MMC_TILE k-mer start positions: the bytes (+ k-1 halo) are staged in shared memory, a thread
 * rolls the forward / reverse-complement k-mers of MMC_RUN consecutive positions through byte-shift registers (two Murmur3 per
 * position, nothing else), candidates are staged in shared memory and written to the tile's region in position order. */
#define MMC_THREADS 256
#ifndef MMC_MINBLOCKS
#define MMC_MINBLOCKS 4
#endif
#define MMC_RUN 9 /* odd: the 32 lanes of a warp start in 32 different shared-memory banks */
#define MMC_TILE (MMC_RUN * MMC_THREADS)
#define MMC_SEQ_BYTES (MMC_TILE + 32 + 48)
struct MmTile {
  int seq, npos;   /* sequence index, k-mer start positions in the tile */
  long long start; /* first k-mer start position */
};
struct MmCandView { /* candidates of tile t: hash[t * cap + j], lp[t * cap + j] = (position - tile start) << 1 | (strand > 0), j < cnt[t] */
  const uint64_t* hash;
  const int* lp;
  const int* cnt; /* -1 = the tile's region overflowed */
  int cap;
};

WFB_DEV void mm_cand_tile(unsigned char* smem, const uint8_t* seqbuf, const MmSeq* seqs, const MmTile t, int tile_index, int k, uint64_t T, int cap,
                          uint64_t* cand_hash, int* cand_lp, int* cand_cnt, unsigned long long* n_cand) {
  uint8_t* sb = (uint8_t*)smem;
  uint64_t* st_h = (uint64_t*)(smem + ((MMC_SEQ_BYTES + 15) & ~15));
  int* st_lp = (int*)(st_h + MMC_TILE);
  int* cnt = st_lp + MMC_TILE;
  const long long g0 = seqs[t.seq].off + t.start; /* first byte of the tile in the (cleaned) buffer */
  const int nbytes = t.npos + k - 1;
#ifndef WFB_EMU
  {
    const long long a0 = g0 & ~(long long)15;
    const int lead = (int)(g0 - a0);
    const int nvec = (lead + nbytes + 15) >> 4;
    for (int v = WFB_TID; v < nvec; v += WFB_NT) ((uint4*)sb)[v] = __ldg((const uint4*)(seqbuf + a0) + v);
    sb += lead;
  }
#else
  for (int i = 0; i < nbytes; ++i) sb[i] = seqbuf[g0 + i];
#endif
  WFB_SYNC();
  const int nruns = (t.npos + MMC_RUN - 1) / MMC_RUN;
  for (int r = WFB_TID; r < MMC_THREADS; r += WFB_NT) {
    int c = 0;
    if (r < nruns) {
      const int p0 = r * MMC_RUN;
      const int p1 = min(t.npos, p0 + MMC_RUN);
      MmRoll roll;
      mm_roll_init(roll, sb + p0, k);
      /* the reference's ambiguity counter (:553-556) only sees an N at sequence index >= k-1: the k-mers holding that base are skipped */
      int nbad = 0;
      for (int j = 0; j < k; ++j) nbad += (sb[p0 + j] == 'N') && (t.start + p0 + j >= (long long)(k - 1));
      for (int p = p0; p < p1; ++p) {
        if (nbad == 0) {
          const uint64_t hf = mm_murmur3_lo64(roll.f0, roll.f1, roll.f2, roll.f3, k), hb = mm_murmur3_lo64(roll.r0, roll.r1, roll.r2, roll.r3, k);
          const uint64_t h = hf < hb ? hf : hb;
          if (hf != hb && h <= T) {
            st_h[p0 + c] = h;
            st_lp[p0 + c] = (p << 1) | (hf < hb ? 1 : 0);
            ++c;
          }
        }
        if (p + 1 < p1) {
          nbad += (sb[p + k] == 'N') - ((sb[p] == 'N') && (t.start + p >= (long long)(k - 1)));
          mm_roll_step(roll, sb[p + k], k);
        }
      }
    }
    cnt[r] = c;
  }
  WFB_SYNC();
  int total = 0;
  for (int r = 0; r < MMC_THREADS; ++r) total += cnt[r]; /* every thread reads the same words: shared-memory broadcasts */
  if (total <= cap) {
    for (int r = WFB_TID; r < MMC_THREADS; r += WFB_NT) {
      int off = 0;
      for (int q = 0; q < r; ++q) off += cnt[q];
      const int c = cnt[r];
      for (int j = 0; j < c; ++j) {
        cand_hash[(long long)tile_index * cap + off + j] = st_h[r * MMC_RUN + j];
        cand_lp[(long long)tile_index * cap + off + j] = st_lp[r * MMC_RUN + j];
      }
    }
  }
  if (WFB_TID == 0) {
    cand_cnt[tile_index] = total <= cap ? total : -1;
    atomicAdd_compat(n_cand, (unsigned long long)total);
  }
  WFB_SYNC();
}

WFB_KERNEL_LB(mm_cand_kernel, MMC_THREADS, MMC_MINBLOCKS, const uint8_t* seqbuf, const MmSeq* seqs, const MmTile* tiles, int ntiles, int k, uint64_t T, int cap,
              uint64_t* cand_hash, int* cand_lp, int* cand_cnt, unsigned long long* n_cand
#ifdef WFB_EMU
              , unsigned char* smem_emu
#endif
) {
  WFB_KERNEL_PROLOGUE
#ifndef WFB_EMU
  __shared__ __align__(16) unsigned char smem[((MMC_SEQ_BYTES + 15) & ~15) + MMC_TILE * 12 + MMC_THREADS * 4];
#else
  unsigned char* smem = smem_emu;
#endif
  for (int i = bid; i < ntiles; i += nblocks) mm_cand_tile(smem, seqbuf, seqs, tiles[i], i, k, T, cap, cand_hash, cand_lp, cand_cnt, n_cand);
}

struct MmCursor { /* a position in the candidate stream of one sequence */
  long long tile, tile_end, tpos0;
  int j, n;
  long long pos; /* LLONG_MAX = exhausted */
  uint64_t hash;
  int strand;
};
WFB_DEV void mm_cursor_next(MmCursor& c, const MmCandView& CV) {
  c.pos = LLONG_MAX;
  while (c.tile < c.tile_end) {
    if (c.j < c.n) {
      const int lp = CV.lp[c.tile * CV.cap + c.j];
      c.pos = c.tpos0 + (lp >> 1);
      c.strand = (lp & 1) ? 1 : -1;
      c.hash = CV.hash[c.tile * CV.cap + c.j];
      ++c.j;
      return;
    }
    ++c.tile; c.tpos0 += MMC_TILE; c.j = 0;
    c.n = c.tile < c.tile_end ? CV.cnt[c.tile] : 0;
  }
}

/* ---- the reference's loop state of one chunk ---- */
template <int L>
struct MmRun {
  MmArr<MmKmer, L> Q;
  int qh, qn, qcap;
  MmHeap<L> H;
  MmPool<L> pool;
  MmArr<MmWent, L> W;
  int wn;
  int k, w, s;
  long long keep_from, body_win0, run_begin;
  int seq, chunk;
  MmRecord* out;
  long long out_cap;
  MmCounters* counters;
  unsigned long long stale, overflow;
  int shortfall; /* filtered run only: the sketch could not be kept at s entries from the candidates alone */
};
template <int L>
WFB_DEV void mm_emit(MmRun<L>& R, long long i, const MmWent& e, long long wend) {
  if (i < R.keep_from) return;
  const unsigned long long idx = (unsigned long long)atomicAdd_compat(&R.counters->n_records, 1ULL);
  if ((long long)idx < R.out_cap) {
    MmRecord r;
    r.hash = e.hash; r.wpos = e.wpos; r.wpos_end = wend; r.seq = R.seq; r.strand = e.strand;
    r.chunk = R.chunk; r.inherited = (e.wpos < R.body_win0 && R.run_begin > 0) ? 1 : 0;
    R.out[idx] = r;
  } else R.overflow++;
}

/* One iteration of the reference's loop (:479-644) at k-mer start position i; `arrive` = a k-mer enters the window here (kk), `leave` =
 * the oldest k-mer of the window (lv) falls out of it here (:517). The window queue itself is the caller's business.
 * FILT = the run only visits the positions where something happens (candidate arrivals, departures, the first fill). */
template <bool FILT, int L>
WFB_DEV void mm_position(MmRun<L>& R, const long long i, const bool leave, const MmKmer lv, const bool arrive, const MmKmer kk) {
  MmHeap<L>& H = R.H;
  MmPool<L>& pool = R.pool;
  const MmArr<MmWent, L>& W = R.W;
  const int s = R.s;
  const long long win = i + R.k - R.w; /* currentWindowId, :482 */
  if (FILT ? (H.n >= H.cap) : (H.n > 2 * R.w)) { /* :485-495 (the filtered heap is small: it is purged when it is full) */
    int m = 0;
    for (int j = 0; j < H.n; ++j) if (!((long long)H.a[j].pos < win)) H.a[m++] = H.a[j];
    H.n = m;
    for (int j = H.n / 2 - 1; j >= 0; --j) mm_heap_sift_down(H, j);
  }
  /* leaving k-mer, :517-551 */
  if (leave) { /* lv = the window's oldest k-mer, which the caller has already taken off the queue */
    if (R.wn > 0 && lv.hash <= W[R.wn - 1].hash) {
      const int lo = mm_lower_bound(W, R.wn, lv.hash);
      if (lo < R.wn && W[lo].hash == lv.hash) {
        MmWent& e = W[lo];
        if (e.count == 1) {
          mm_emit(R, i, e, win);
          mm_went_clear(pool, e);
          for (int j = lo; j + 1 < R.wn; ++j) W[j] = W[j + 1];
          --R.wn;
        } else {
          if (e.strand - lv.strand == 0 || e.strand == 0) {
            mm_emit(R, i, e, win);
            e.wpos = win;
          }
          e.strand -= lv.strand;
          mm_went_pop_front(pool, e);
        }
      }
    }
  }
  if (arrive) { /* :559-586 (the caller has queued kk) */
    const int lo = mm_lower_bound(W, R.wn, kk.hash);
    if (lo < R.wn && W[lo].hash == kk.hash) {
      MmWent& e = W[lo];
      if (!mm_went_push_back(pool, e, kk.pos, kk.strand)) R.overflow++;
      if (e.strand + kk.strand == 0 || e.strand == 0) {
        mm_emit(R, i, e, win);
        e.wpos = win;
      }
      e.strand += kk.strand;
    } else {
      if (!mm_heap_push(H, kk)) R.overflow++;
    }
  }
  if (win >= R.run_begin) { /* :593-643 (win >= 0 for a run from the sequence start) */
    while (H.n > 0 && (long long)H.a[0].pos < win) mm_heap_pop(H);
    if (R.wn > 0 && H.n > 0 && R.wn == s && H.a[0].hash < W[R.wn - 1].hash) {
      MmWent& e = W[R.wn - 1];
      mm_emit(R, i, e, win);
      for (int n = e.head; n >= 0; n = pool.nodes[n].next)
        if ((long long)pool.nodes[n].pos > win) {
          MmKmer b;
          b.hash = e.hash; b.pos = pool.nodes[n].pos; b.strand = pool.nodes[n].strand;
          if (!mm_heap_push(H, b)) R.overflow++;
        }
      mm_went_clear(pool, e);
      --R.wn;
    }
    while (H.n > 0 && R.wn < s) {
      if ((long long)H.a[0].pos < win) { /* may empty the heap; a[0] stays readable, see :627-633 */
        mm_heap_pop(H);
        if (FILT) R.shortfall = 1; /* the reference's next front may be a k-mer the filter dropped */
      }
      const MmKmer nk = H.a[0];
      const int lo = mm_lower_bound(W, R.wn, nk.hash);
      if (!(lo < R.wn && W[lo].hash == nk.hash)) {
        for (int j = R.wn; j > lo; --j) W[j] = W[j - 1];
        ++R.wn;
        W[lo].head = W[lo].tail = -1;
        W[lo].count = 0;
      }
      W[lo].hash = nk.hash;
      W[lo].wpos = win;
      W[lo].strand = 0;
      while (H.n > 0 && H.a[0].hash == nk.hash) {
        if ((long long)H.a[0].pos < win) R.stale++; /* the reference absorbs expired entries too (:635-641) */
        if (!mm_went_push_back(pool, W[lo], H.a[0].pos, H.a[0].strand)) R.overflow++;
        W[lo].strand += H.a[0].strand;
        mm_heap_pop(H);
      }
    }
    if (FILT && R.wn < s) R.shortfall = 1; /* the reference would have filled the sketch from k-mers above the threshold */
  }
}

/* One thread = one chunk. scratch layout per WARP (scratch_stride bytes per chunk, MM_LANES chunks per slab):
 * Q[qcap][lanes] | heap[heap_cap][lanes] | pool[pool_cap][lanes] | W[s+2][lanes].
 * FILT = false: every position of the chunk (the exact run; `redo` != NULL lists the chunks to run, c indexes it).
 * FILT = true : only the candidates of mm_cand_kernel; a chunk that cannot vouch for its result sets chunk_flag[c] and is re-run
 *               by the exact instantiation (its records, recognisable by their chunk id, are dropped by the post pass). */
template <bool FILT, int L>
WFB_DEV void mm_stream_chunk(const int slot, const int c, const uint8_t* seqbuf, const MmSeq* seqs, const MmChunk* chunks, const MmParams P,
                             unsigned char* scratch_all, long long scratch_stride, MmRecord* out, long long out_cap, MmEndEnt* endstate,
                             int* endcount, MmCounters* counters, const MmCandView CV, const int* seq_tile0, int* chunk_flag, int* redo_list) {
  const MmChunk ch = chunks[c];
  const MmSeq sq = seqs[ch.seq];
  const uint8_t* seq = seqbuf + sq.off;
  const long long len = sq.len;
  const int k = P.k, w = P.w, s = P.s;
  const int lane = slot % L;
  unsigned char* sp = scratch_all + (long long)(slot / L) * scratch_stride * L;
  MmRun<L> R;
  R.Q = MmArr<MmKmer, L>{(MmKmer*)sp + lane};
  R.H.a = MmArr<MmKmer, L>{(MmKmer*)(sp + sizeof(MmKmer) * (size_t)P.qcap * L) + lane};
  R.H.n = 0;
  R.H.cap = P.heap_cap;
  R.pool.nodes = MmArr<MmNode, L>{(MmNode*)(sp + sizeof(MmKmer) * ((size_t)P.qcap + P.heap_cap) * L) + lane};
  R.W = MmArr<MmWent, L>{(MmWent*)(sp + (sizeof(MmKmer) * ((size_t)P.qcap + P.heap_cap) + sizeof(MmNode) * (size_t)P.pool_cap) * L) + lane};
  for (int i = 0; i < P.pool_cap; ++i) R.pool.nodes[i].next = (i + 1 < P.pool_cap) ? i + 1 : -1;
  R.pool.free_head = 0;
  R.qh = 0; R.qn = 0; R.qcap = P.qcap; R.wn = 0;
  R.k = k; R.w = w; R.s = s;
  R.stale = 0; R.overflow = 0; R.shortfall = 0;
  const long long run_begin = ch.run_begin, run_end = ch.body_end;
  R.run_begin = run_begin; R.keep_from = ch.body_begin;
  R.body_win0 = ch.body_begin + k - w; /* window id of the first body step */
  R.seq = ch.seq; R.chunk = c; R.out = out; R.out_cap = out_cap; R.counters = counters;
  if (!FILT) {
    int ambig = 0;
    if (run_begin > 0) { /* the counter a run from position 0 holds here (only N at index >= k-1 arm it, :553-556) */
      for (long long j = run_begin + k - 2; j >= run_begin && j >= k - 1; --j)
        if (seq[j] == 'N') { ambig = (int)(j - run_begin + 1); break; }
    }
    MmRoll roll;
    if (run_begin < run_end) mm_roll_init(roll, seq + run_begin, k);
    for (long long i = run_begin; i < run_end; ++i) {
      const uint64_t hf = mm_murmur3_lo64(roll.f0, roll.f1, roll.f2, roll.f3, k), hb = mm_murmur3_lo64(roll.r0, roll.r1, roll.r2, roll.r3, k);
      if (i + 1 < run_end) mm_roll_step(roll, seq[i + k], k);
      if (seq[i + k - 1] == 'N') ambig = k;
      MmKmer kk, lv;
      kk.hash = hf < hb ? hf : hb; kk.pos = (int)i; kk.strand = hf < hb ? 1 : -1;
      lv = kk;
      const bool arrive = hb != hf && ambig == 0;
      const bool leave = R.qn > 0 && (long long)R.Q[R.qh].pos < i + k - w; /* :517 */
      if (leave) { lv = R.Q[R.qh]; R.qh = (R.qh + 1) % R.qcap; --R.qn; }
      if (arrive) { if (R.qn < R.qcap) { R.Q[(R.qh + R.qn) % R.qcap] = kk; ++R.qn; } else R.overflow++; }
      mm_position<false>(R, i, leave, lv, arrive, kk);
      if (ambig > 0) --ambig;
    }
  } else {
    /* two cursors over the candidates of the tiles that cover [run_begin, run_end): arrivals, and (P.lcur) the same stream again
     * w - k + 1 positions later as the departures — the window queue of a filtered run IS the candidate list */
    MmCursor A, D;
    A.tile = (long long)seq_tile0[ch.seq] + run_begin / MMC_TILE;
    A.tile_end = (long long)seq_tile0[ch.seq] + (run_end + MMC_TILE - 1) / MMC_TILE;
    A.tpos0 = run_begin / MMC_TILE * MMC_TILE;
    bool bad = false;
    for (long long t = A.tile; t < A.tile_end; ++t) bad = bad || CV.cnt[t] < 0;
    A.j = 0; A.n = (A.tile < A.tile_end && !bad) ? CV.cnt[A.tile] : 0;
    A.pos = LLONG_MAX; A.hash = 0; A.strand = 0;
    const long long NONE = LLONG_MAX;
    if (!bad) {
      mm_cursor_next(A, CV);
      while (A.pos != NONE && A.pos < run_begin) mm_cursor_next(A, CV);
      D = A;
      const bool lcur = P.lcur != 0;
      const long long i0 = run_begin + w - k; /* the first step that fills the sketch (its window id is every entry's start) */
      bool did_i0 = false;
      for (;;) {
        const long long oldest = lcur ? D.pos : (R.qn > 0 ? (long long)R.Q[R.qh].pos : NONE);
        const long long nl = oldest != NONE ? oldest + (w - k + 1) : NONE;
        long long i = A.pos < nl ? A.pos : nl;
        if (!did_i0 && i0 <= i) { i = i0; did_i0 = true; }
        if (i >= run_end) break;
        const bool arrive = i == A.pos, leave = i == nl;
        MmKmer kk, lv;
        kk.hash = A.hash; kk.pos = (int)i; kk.strand = A.strand;
        lv = kk;
        if (leave) {
          if (lcur) { lv.hash = D.hash; lv.pos = (int)D.pos; lv.strand = D.strand; mm_cursor_next(D, CV); }
          else { lv = R.Q[R.qh]; R.qh = (R.qh + 1) % R.qcap; --R.qn; }
        }
        if (arrive && !lcur) { if (R.qn < R.qcap) { R.Q[(R.qh + R.qn) % R.qcap] = kk; ++R.qn; } else R.overflow++; }
        mm_position<true>(R, i, leave, lv, arrive, kk);
        if (arrive) mm_cursor_next(A, CV);
        if (R.shortfall | (R.overflow != 0) | (R.stale != 0)) break;
      }
    }
    if (bad || R.shortfall || R.overflow || R.stale) { /* the exact instantiation re-runs this chunk */
      chunk_flag[c] = 1;
      redo_list[atomicAdd_compat(&counters->flagged, 1ULL)] = c;
      endcount[c] = 0;
      return;
    }
  }
  /* membership at the end of the body, for the stitch of the next chunk */
  {
    MmEndEnt* es = endstate + (long long)c * s;
    for (int j = 0; j < R.wn && j < s; ++j) {
      es[j].hash = R.W[j].hash;
      es[j].wpos = R.W[j].wpos;
      es[j].inherited = (R.W[j].wpos < R.body_win0 && run_begin > 0) ? 1 : 0;
      es[j].pad_ = 0;
    }
    endcount[c] = R.wn < s ? R.wn : s;
  }
  /* :646-658: still-open members are closed at len - k + 1 */
  if (ch.last_of_seq) {
    for (int j = 0; j < R.wn && j < s; ++j)
      if (R.W[j].wpos != -1) mm_emit(R, run_end, R.W[j], len - k + 1); /* emissions of the flush count as body emissions */
  }
  if (R.stale) atomicAdd_compat(&counters->stale_absorbed, R.stale);
  if (R.overflow) atomicAdd_compat(&counters->overflow, R.overflow);
}

WFB_KERNEL(mm_stream_kernel, const uint8_t* seqbuf, const MmSeq* seqs, const MmChunk* chunks, int nchunks, MmParams P,
           unsigned char* scratch_all, long long scratch_stride, MmRecord* out, long long out_cap, MmEndEnt* endstate,
           int* endcount, MmCounters* counters, const int* redo) {
  WFB_KERNEL_PROLOGUE
  const int t = bid * WFB_NT + WFB_TID;
  if (t >= nchunks) return;
  MmCandView none;
  none.hash = nullptr; none.lp = nullptr; none.cnt = nullptr; none.cap = 0;
  mm_stream_chunk<false, MM_LANES>(t, redo ? redo[t] : t, seqbuf, seqs, chunks, P, scratch_all, scratch_stride, out, out_cap, endstate, endcount, counters,
                         none, nullptr, nullptr, nullptr);
}
#ifndef WFB_EMU
/* The exact re-run of the few chunks a filtered run flagged: a lone thread walking 1500 positions with its containers in global
 * memory is a chain of ~40 dependent L2 round trips per position (the whole build waited ~20 ms for 355 such threads). One CTA per
 * chunk, its containers private in dynamic shared memory (lane stride 1), one working thread. */
__global__ void __launch_bounds__(32, 1) mm_stream_redo_smem_kernel(const uint8_t* seqbuf, const MmSeq* seqs, const MmChunk* chunks, int nredo, MmParams P,
                                                                    long long scratch_stride, MmRecord* out, long long out_cap, MmEndEnt* endstate,
                                                                    int* endcount, MmCounters* counters, const int* redo) {
  extern __shared__ __align__(16) unsigned char mm_redo_smem[];
  if (threadIdx.x != 0 || (int)blockIdx.x >= nredo) return;
  MmCandView none;
  none.hash = nullptr; none.lp = nullptr; none.cnt = nullptr; none.cap = 0;
  mm_stream_chunk<false, 1>(0, redo[blockIdx.x], seqbuf, seqs, chunks, P, mm_redo_smem, scratch_stride, out, out_cap, endstate, endcount, counters,
                            none, nullptr, nullptr, nullptr);
}
#endif
#ifndef WFB_EMU
/* The filtered run with every chunk's (small) containers in dynamic shared memory: thread t owns the slice at t * scratch_stride, lane
 * stride 1; needs P.lcur (no window queue). Few threads per SM, but an event is a chain of shared-memory accesses instead of L2 / HBM ones. */
__global__ void mm_stream_cand_smem_kernel(const uint8_t* seqbuf, const MmSeq* seqs, const MmChunk* chunks, int nchunks, MmParams P, long long scratch_stride,
                                           MmRecord* out, long long out_cap, MmEndEnt* endstate, int* endcount, MmCounters* counters, MmCandView CV,
                                           const int* seq_tile0, int* chunk_flag, int* redo_list) {
  extern __shared__ __align__(16) unsigned char mm_filt_smem[];
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (t >= nchunks) return;
  mm_stream_chunk<true, 1>((int)threadIdx.x, t, seqbuf, seqs, chunks, P, mm_filt_smem, scratch_stride, out, out_cap, endstate, endcount, counters, CV,
                           seq_tile0, chunk_flag, redo_list);
}
#endif
WFB_KERNEL(mm_stream_cand_kernel, const uint8_t* seqbuf, const MmSeq* seqs, const MmChunk* chunks, int nchunks, MmParams P,
           unsigned char* scratch_all, long long scratch_stride, MmRecord* out, long long out_cap, MmEndEnt* endstate,
           int* endcount, MmCounters* counters, MmCandView CV, const int* seq_tile0, int* chunk_flag, int* redo_list) {
  WFB_KERNEL_PROLOGUE
  const int t = bid * WFB_NT + WFB_TID;
  if (t >= nchunks) return;
  mm_stream_chunk<true, MM_LANES>(t, t, seqbuf, seqs, chunks, P, scratch_all, scratch_stride, out, out_cap, endstate, endcount, counters, CV, seq_tile0,
                        chunk_flag, redo_list);
}

/* Stitch 1: resolve the true interval starts of the chunk end states (each end state has <= s entries sorted by hash),
 * one thread per (chunk, entry). An inherited entry takes its start from the previous chunk's entry of the same hash
 * once THAT entry is settled. Passes repeat until *pending == 0: an interval that stays in the sketch across n chunks
 * takes n passes (one or two on ordinary sequence; the host bounds the count). A writer publishes wpos before it
 * clears `inherited`. */
WFB_KERNEL(mm_stitch_ends_pass_kernel, const MmChunk* chunks, int nchunks, int s, MmEndEnt* endstate, const int* endcount,
           MmCounters* counters, int* pending) {
  WFB_KERNEL_PROLOGUE
  const long long total = (long long)nchunks * s;
  for (long long e = (long long)bid * WFB_NT + WFB_TID; e < total; e += (long long)nblocks * WFB_NT) {
    const int c = (int)(e / s), j = (int)(e - (long long)c * s);
    if (j >= endcount[c] || chunks[c].first_of_seq) continue;
    MmEndEnt* cur = endstate + (long long)c * s + j;
    if (!((volatile MmEndEnt*)cur)->inherited) continue;
    const MmEndEnt* prev = endstate + (long long)(c - 1) * s;
    const int np = endcount[c - 1];
    const uint64_t h = cur->hash;
    int lo = 0, hi = np;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (prev[mid].hash < h) lo = mid + 1; else hi = mid; }
    if (lo < np && prev[lo].hash == h) {
      if (((volatile const MmEndEnt*)&prev[lo])->inherited) { wfb_atomic_add(pending, 1); continue; } /* not settled yet */
#ifndef WFB_EMU
      __threadfence();
#endif
      cur->wpos = ((volatile const MmEndEnt*)&prev[lo])->wpos;
#ifndef WFB_EMU
      __threadfence();
#endif
      ((volatile MmEndEnt*)cur)->inherited = 0;
    } else {
      atomicAdd_compat(&counters->stitch_miss, 1ULL);
      ((volatile MmEndEnt*)cur)->inherited = 0; /* settled with its local start, like the serial walk leaves it */
    }
  }
}

/* Stitch 2: records that inherited their start take it from the previous chunk's (resolved) end state. */
WFB_KERNEL(mm_stitch_records_kernel, MmRecord* recs, long long nrec, const MmChunk* chunks, int s, const MmEndEnt* endstate,
           const int* endcount, MmCounters* counters, const int* chunk_flag, long long nrec_filtered) {
  WFB_KERNEL_PROLOGUE
  for (long long r = (long long)bid * WFB_NT + WFB_TID; r < nrec; r += (long long)nblocks * WFB_NT) {
    if (!recs[r].inherited) continue;
    const int c = recs[r].chunk;
    if (r < nrec_filtered && chunk_flag[c]) continue; /* written by a filtered run that gave up: dropped by the post pass */
    if (chunks[c].first_of_seq) continue;
    const MmEndEnt* prev = endstate + (long long)(c - 1) * s;
    const int np = endcount[c - 1];
    const uint64_t h = recs[r].hash;
    int lo = 0, hi = np;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (prev[mid].hash < h) lo = mid + 1; else hi = mid; }
    if (lo < np && prev[lo].hash == h) { recs[r].wpos = prev[lo].wpos; recs[r].inherited = 0; }
    else atomicAdd_compat(&counters->stitch_miss, 1ULL);
  }
}
