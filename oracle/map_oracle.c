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

/* =================================================================================================
 * addMinmers (commonFunc.hpp:439-708): the reference's streaming window sketch, restated with
 * explicit containers so that every history-dependent quirk is preserved:
 *   Q            std::deque<(hash,strand,pos)>   -> ring buffer
 *   sortedWindow std::map<hash,(MinmerInfo,deque<KmerInfo>)> -> array sorted by hash, occurrence
 *                lists as singly linked lists in a node pool
 *   heapWindow   std::vector<KmerInfo> + push_heap/pop_heap with (hash,pos) ascending at the front
 *                -> binary min-heap (only front() and size() are observable, so any heap works)
 * chunk_begin/chunk_end/warm model the GPU decomposition (see orc_add_minmers_chunked).
 * ================================================================================================= */
typedef struct { uint64_t hash; int64_t pos; int16_t strand; } orc_kmer_t;
typedef struct { orc_kmer_t k; int next; } orc_node_t;
typedef struct { uint64_t hash; orc_minmer_t mi; int head, tail, count; } orc_went_t;

typedef struct {
  orc_kmer_t* a; int n, cap;
} orc_heap_t;

static int kmer_less(const orc_kmer_t* x, const orc_kmer_t* y) {
  return x->hash < y->hash || (x->hash == y->hash && x->pos < y->pos);
}
static void heap_push(orc_heap_t* h, orc_kmer_t v) {
  if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->a = (orc_kmer_t*)realloc(h->a, (size_t)h->cap * sizeof(orc_kmer_t)); }
  int i = h->n++;
  while (i > 0) {
    const int p = (i - 1) / 2;
    if (!kmer_less(&v, &h->a[p])) break;
    h->a[i] = h->a[p];
    i = p;
  }
  h->a[i] = v;
}
static void heap_sift_down(orc_heap_t* h, int i) {
  const orc_kmer_t v = h->a[i];
  for (;;) {
    int c = 2 * i + 1;
    if (c >= h->n) break;
    if (c + 1 < h->n && kmer_less(&h->a[c + 1], &h->a[c])) ++c;
    if (!kmer_less(&h->a[c], &v)) break;
    h->a[i] = h->a[c];
    i = c;
  }
  h->a[i] = v;
}
static void heap_pop(orc_heap_t* h) {
  h->a[0] = h->a[--h->n];
  if (h->n > 0) heap_sift_down(h, 0);
}

typedef struct {
  orc_minmer_t* v; int64_t n, cap;
} orc_mivec_t;
static void mivec_push(orc_mivec_t* o, orc_minmer_t m) {
  if (o->n == o->cap) { o->cap = o->cap ? o->cap * 2 : 4096; o->v = (orc_minmer_t*)realloc(o->v, (size_t)o->cap * sizeof(orc_minmer_t)); }
  o->v[o->n++] = m;
}

typedef struct {
  orc_node_t* nodes; int ncap, nfree; /* free list head */
} orc_pool_t;
static int pool_alloc(orc_pool_t* p) {
  if (p->nfree < 0) {
    const int old = p->ncap;
    p->ncap = old ? old * 2 : 4096;
    p->nodes = (orc_node_t*)realloc(p->nodes, (size_t)p->ncap * sizeof(orc_node_t));
    for (int i = old; i < p->ncap; ++i) p->nodes[i].next = (i + 1 < p->ncap) ? i + 1 : -1;
    p->nfree = old;
  }
  const int i = p->nfree;
  p->nfree = p->nodes[i].next;
  return i;
}
static void pool_free(orc_pool_t* p, int i) { p->nodes[i].next = p->nfree; p->nfree = i; }

static void went_push_back(orc_pool_t* p, orc_went_t* e, orc_kmer_t k) {
  const int n = pool_alloc(p);
  p->nodes[n].k = k; p->nodes[n].next = -1;
  if (e->tail >= 0) p->nodes[e->tail].next = n; else e->head = n;
  e->tail = n; e->count++;
}
static void went_pop_front(orc_pool_t* p, orc_went_t* e) {
  const int n = e->head;
  e->head = p->nodes[n].next;
  if (e->head < 0) e->tail = -1;
  e->count--;
  pool_free(p, n);
}
static void went_clear(orc_pool_t* p, orc_went_t* e) { while (e->head >= 0) went_pop_front(p, e); }

static int cmp_mi_pos(const void* a, const void* b) {
  const orc_minmer_t* x = (const orc_minmer_t*)a; const orc_minmer_t* y = (const orc_minmer_t*)b;
  if (x->wpos != y->wpos) return x->wpos < y->wpos ? -1 : 1;
  if (x->wpos_end != y->wpos_end) return x->wpos_end < y->wpos_end ? -1 : 1;
  return 0;
}

/* The stream part of addMinmers (:477-658) over positions [0, len-k]. Raw (pre-post-pass) records are
 * appended to raw. */
/* [run_begin, run_end): k-mer start positions processed (a sub-range models one GPU chunk incl. its
 * warm-up); records are kept only when emitted at a step >= keep_from; final_flush = the :646-658 tail. */
static void add_minmers_stream_range(const char* seq, int64_t len, int k, int w, int s, int32_t seqId, orc_mivec_t* raw,
                                     int64_t run_begin, int64_t run_end, int64_t keep_from, int final_flush, orc_mivec_t* open_out);
static void add_minmers_stream(const char* seq, int64_t len, int k, int w, int s, int32_t seqId, orc_mivec_t* raw) {
  add_minmers_stream_range(seq, len, k, w, s, seqId, raw, 0, len - k + 1, 0, 1, 0);
}
static void add_minmers_stream_range(const char* seq, int64_t len, int k, int w, int s, int32_t seqId, orc_mivec_t* raw,
                                     int64_t run_begin, int64_t run_end, int64_t keep_from, int final_flush, orc_mivec_t* open_out) {
  orc_kmer_t* Q = (orc_kmer_t*)malloc((size_t)(w + 2) * sizeof(orc_kmer_t));
  int qh = 0, qn = 0; const int qcap = w + 2;
  orc_went_t* W = (orc_went_t*)calloc((size_t)s + 2, sizeof(orc_went_t));
  int wn = 0;
  orc_heap_t H = {0, 0, 0};
  orc_pool_t P = {0, 0, -1};
  char* rc = (char*)malloc((size_t)k);
  int ambig = 0;
  if (run_begin > 0) { /* the counter a run from position 0 would hold here (armed by N at index >= k-1) */
    for (int64_t j = run_begin + k - 2; j >= run_begin && j >= k - 1; --j)
      if (seq[j] == 'N') { ambig = (int)(j - run_begin + 1); break; }
  }
  const int64_t raw_n0 = raw->n;
  int64_t keep_mark = -1; /* raw->n when step keep_from starts */
  for (int64_t i = run_begin; i < run_end; ++i) {
    if (i == keep_from) keep_mark = raw->n;
    const int64_t win = i + k - w; /* currentWindowId */
    if (H.n > 2 * w) { /* :485-495 */
      int m = 0;
      for (int j = 0; j < H.n; ++j) if (!(H.a[j].pos < win)) H.a[m++] = H.a[j];
      H.n = m;
      for (int j = H.n / 2 - 1; j >= 0; --j) heap_sift_down(&H, j);
    }
    revcomp(seq + i, rc, k);
    const uint64_t hf = orc_kmer_hash(seq + i, k), hb = orc_kmer_hash(rc, k);
    const uint64_t cur = hf < hb ? hf : hb;
    const int16_t cur_strand = hf < hb ? 1 : -1;
    /* leaving k-mer (:517-551) */
    if (qn > 0 && Q[qh].pos < win) {
      const orc_kmer_t lv = Q[qh];
      if (wn > 0 && lv.hash <= W[wn - 1].hash) {
        int lo = 0, hi = wn; /* find */
        while (lo < hi) { const int mid = (lo + hi) / 2; if (W[mid].hash < lv.hash) lo = mid + 1; else hi = mid; }
        if (lo < wn && W[lo].hash == lv.hash) {
          orc_went_t* e = &W[lo];
          if (e->count == 1) {
            e->mi.wpos_end = win;
            mivec_push(raw, e->mi);
            went_clear(&P, e);
            memmove(&W[lo], &W[lo + 1], (size_t)(wn - lo - 1) * sizeof(orc_went_t));
            --wn;
          } else {
            if (e->mi.strand - lv.strand == 0 || e->mi.strand == 0) {
              e->mi.wpos_end = win;
              mivec_push(raw, e->mi);
              e->mi.wpos = win;
              e->mi.wpos_end = -1;
            }
            e->mi.strand = (int16_t)(e->mi.strand - lv.strand);
            went_pop_front(&P, e);
          }
        } /* else: the reference dereferences end() here (undefined); unreachable with a consistent state */
      }
      qh = (qh + 1) % qcap; --qn;
    }
    if (seq[i + k - 1] == 'N') ambig = k;
    if (hb != hf && ambig == 0) {
      orc_kmer_t kk; kk.hash = cur; kk.pos = i; kk.strand = cur_strand;
      Q[(qh + qn) % qcap] = kk; ++qn;
      int lo = 0, hi = wn;
      while (lo < hi) { const int mid = (lo + hi) / 2; if (W[mid].hash < cur) lo = mid + 1; else hi = mid; }
      if (lo < wn && W[lo].hash == cur) {
        orc_went_t* e = &W[lo];
        went_push_back(&P, e, kk);
        if (e->mi.strand + cur_strand == 0 || e->mi.strand == 0) {
          e->mi.wpos_end = win;
          mivec_push(raw, e->mi);
          e->mi.wpos = win;
          e->mi.wpos_end = -1;
        }
        e->mi.strand = (int16_t)(e->mi.strand + cur_strand);
      } else {
        heap_push(&H, kk);
      }
    }
    if (ambig > 0) --ambig;
    if (win >= run_begin) { /* :593-643 (win >= 0 in a run from the sequence start) */
      while (H.n > 0 && H.a[0].pos < win) heap_pop(&H);
      if (wn > 0 && H.n > 0 && wn == s && H.a[0].hash < W[wn - 1].hash) {
        orc_went_t* e = &W[wn - 1];
        e->mi.wpos_end = win;
        mivec_push(raw, e->mi);
        for (int n = e->head; n >= 0; n = P.nodes[n].next)
          if (P.nodes[n].k.pos > win) heap_push(&H, P.nodes[n].k);
        went_clear(&P, e);
        --wn;
      }
      while (H.n > 0 && wn < s) {
        if (H.a[0].pos < win) heap_pop(&H);
        /* if that emptied the heap the reference reads front() of an empty vector; in practice it sees the
         * element just popped, which still sits in the vector's storage (heap_pop leaves a[0] in place) */
        const orc_kmer_t nk = H.a[0];
        int lo = 0, hi = wn;
        while (lo < hi) { const int mid = (lo + hi) / 2; if (W[mid].hash < nk.hash) lo = mid + 1; else hi = mid; }
        if (!(lo < wn && W[lo].hash == nk.hash)) {
          memmove(&W[lo + 1], &W[lo], (size_t)(wn - lo) * sizeof(orc_went_t));
          ++wn;
          W[lo].head = W[lo].tail = -1; W[lo].count = 0;
        } /* else: operator[] on an existing key assigns .first only and keeps the occurrence list */
        W[lo].hash = nk.hash;
        W[lo].mi.hash = nk.hash; W[lo].mi.wpos = win; W[lo].mi.wpos_end = -1; W[lo].mi.seqId = seqId; W[lo].mi.strand = 0; W[lo].mi.pad_ = 0;
        while (H.n > 0 && H.a[0].hash == nk.hash) {
          went_push_back(&P, &W[lo], H.a[0]);
          W[lo].mi.strand = (int16_t)(W[lo].mi.strand + H.a[0].strand);
          heap_pop(&H);
        }
      }
    }
  }
  if (keep_mark < 0) keep_mark = raw->n;
  if (keep_mark > raw_n0) { /* drop what the warm-up emitted */
    memmove(raw->v + raw_n0, raw->v + keep_mark, (size_t)(raw->n - keep_mark) * sizeof(orc_minmer_t));
    raw->n -= keep_mark - raw_n0;
  }
  if (open_out) for (int j = 0; j < wn; ++j) mivec_push(open_out, W[j].mi);
  /* :646-658 */
  if (final_flush) for (int j = 0; j < wn && j < s; ++j) {
    if (W[j].mi.wpos != -1) {
      W[j].mi.wpos_end = len - k + 1;
      mivec_push(raw, W[j].mi);
    }
  }
  for (int j = 0; j < wn; ++j) went_clear(&P, &W[j]);
  free(Q); free(W); free(H.a); free(P.nodes); free(rc);
}

/* ---- the reference's final std::sort (:696), restated ----
 * The comparator looks at (wpos, wpos_end) only and std::sort is not stable, so the order of records with equal keys is whatever the
 * library's algorithm leaves — unspecified by the standard, but a pure function of the input order for a given library. The reference
 * is built with GNU libstdc++, whose std::sort is: introsort (median-of-three of first+1 / middle / last-1 moved to first, Hoare-style
 * unguarded partition, recursion on the right part, depth limit 2*floor(log2 n) with a heapsort fall-back) down to runs of <= 16
 * elements, then one insertion sort over the whole range (guarded for the first 16 elements, unguarded after). That order decides which
 * of several minmers opened and closed by the same windows the L2 stage sees first (mappingCore.hpp:352-384 evaluates the sketch after
 * every single insertion), i.e. it is visible in the mappings of targets barely longer than one window. Restated here step by step so
 * that the oracle's order IS the reference's; pinned by tests/test_map_oracle_cpu.py against the compiled reference (exact order). */
static int mi_less(const orc_minmer_t* l, const orc_minmer_t* r) {
  return l->wpos < r->wpos || (l->wpos == r->wpos && l->wpos_end < r->wpos_end);
}
static void mi_swap(orc_minmer_t* a, orc_minmer_t* b) { const orc_minmer_t t = *a; *a = *b; *b = t; }
static void gss_push_heap(orc_minmer_t* first, int64_t hole, int64_t top, orc_minmer_t value) {
  int64_t parent = (hole - 1) / 2;
  while (hole > top && mi_less(&first[parent], &value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void gss_adjust_heap(orc_minmer_t* first, int64_t hole, int64_t len, orc_minmer_t value) {
  const int64_t top = hole;
  int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (mi_less(&first[child], &first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  gss_push_heap(first, hole, top, value);
}
static void gss_heapsort(orc_minmer_t* first, int64_t len) { /* partial_sort(first, last, last): make_heap, then sort_heap */
  if (len >= 2) {
    int64_t parent = (len - 2) / 2;
    for (;;) {
      gss_adjust_heap(first, parent, len, first[parent]);
      if (parent == 0) break;
      parent--;
    }
  }
  for (int64_t last = len; last > 1;) {
    --last;
    const orc_minmer_t value = first[last];
    first[last] = first[0];
    gss_adjust_heap(first, 0, last, value);
  }
}
static void gss_unguarded_linear_insert(orc_minmer_t* last) {
  const orc_minmer_t val = *last;
  orc_minmer_t* next = last - 1;
  while (mi_less(&val, next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}
static void gss_insertion_sort(orc_minmer_t* first, orc_minmer_t* last) {
  if (first == last) return;
  for (orc_minmer_t* i = first + 1; i != last; ++i) {
    if (mi_less(i, first)) {
      const orc_minmer_t val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(orc_minmer_t));
      *first = val;
    } else gss_unguarded_linear_insert(i);
  }
}
static void gss_introsort_loop(orc_minmer_t* first, orc_minmer_t* last, int64_t depth_limit) {
  while (last - first > 16) {
    if (depth_limit == 0) { gss_heapsort(first, last - first); return; }
    --depth_limit;
    /* median of (first + 1, mid, last - 1) to *first */
    orc_minmer_t *a = first + 1, *b = first + (last - first) / 2, *c = last - 1;
    if (mi_less(a, b)) {
      if (mi_less(b, c)) mi_swap(first, b);
      else if (mi_less(a, c)) mi_swap(first, c);
      else mi_swap(first, a);
    } else if (mi_less(a, c)) mi_swap(first, a);
    else if (mi_less(b, c)) mi_swap(first, c);
    else mi_swap(first, b);
    /* unguarded partition of [first + 1, last) around *first */
    orc_minmer_t *lo = first + 1, *hi = last;
    for (;;) {
      while (mi_less(lo, first)) ++lo;
      --hi;
      while (mi_less(first, hi)) --hi;
      if (!(lo < hi)) break;
      mi_swap(lo, hi);
      ++lo;
    }
    gss_introsort_loop(lo, last, depth_limit);
    last = lo;
  }
}
static void gnu_std_sort(orc_minmer_t* first, int64_t n) {
  if (n <= 0) return;
  int64_t lg = 0;
  for (int64_t t = n; t > 1; t >>= 1) ++lg;
  gss_introsort_loop(first, first + n, 2 * lg);
  if (n > 16) {
    gss_insertion_sort(first, first + 16);
    for (orc_minmer_t* i = first + 16; i != first + n; ++i) gss_unguarded_linear_insert(i);
  } else gss_insertion_sort(first, first + n);
}

/* Post passes (:660-706) in the reference's own order of operations: drop degenerate records, strand sign, records longer than w are
 * REMOVED and their pieces appended after all the others (:670-694), std::sort by (wpos, wpos_end) (above), std::unique on (wpos, hash). */
static int64_t add_minmers_post(orc_mivec_t* raw, int w, orc_minmer_t* out, int64_t cap) {
  orc_mivec_t v = {0, 0, 0}, pieces = {0, 0, 0};
  for (int64_t i = 0; i < raw->n; ++i) {
    orc_minmer_t m = raw->v[i];
    if (m.wpos < 0 || m.wpos_end < 0 || m.wpos == m.wpos_end) continue;
    m.strand = m.strand < 0 ? -1 : 1;
    if (m.wpos_end > m.wpos + w) {
      const int nch = (int)ceilf((float)(m.wpos_end - m.wpos) / (float)w);
      for (int c = 0; c < nch; ++c) {
        orc_minmer_t p = m;
        p.wpos = m.wpos + (int64_t)c * w;
        p.wpos_end = (m.wpos + (int64_t)c * w + w < m.wpos_end) ? m.wpos + (int64_t)c * w + w : m.wpos_end;
        mivec_push(&pieces, p);
      }
    } else {
      mivec_push(&v, m);
    }
  }
  for (int64_t i = 0; i < pieces.n; ++i) mivec_push(&v, pieces.v[i]);
  free(pieces.v);
  gnu_std_sort(v.v, v.n);
  int64_t n = 0;
  orc_minmer_t prev;
  memset(&prev, 0, sizeof(prev));
  for (int64_t i = 0; i < v.n; ++i) { /* std::unique compares with the last KEPT element */
    if (n > 0 && v.v[i].wpos == prev.wpos && v.v[i].hash == prev.hash) continue;
    prev = v.v[i];
    if (n < cap) out[n] = v.v[i];
    ++n;
  }
  free(v.v);
  return n;
}

int64_t orc_add_minmers(const char* seq, int64_t len, int k, int w, int s, int32_t seqId, orc_minmer_t* out, int64_t cap) {
  orc_mivec_t raw = {0, 0, 0};
  add_minmers_stream(seq, len, k, w, s, seqId, &raw);
  const int64_t n = add_minmers_post(&raw, w, out, cap);
  free(raw.v);
  return n;
}

/* TEST PROBE for the GPU decomposition: raw (pre-post-pass) records of the stream, either from one run
 * over the whole sequence (chunk <= 0) or from independent chunks of `chunk` positions each preceded by
 * `warm` warm-up positions. Records carry the state machine's own wpos (not stitched). */
int64_t orc_add_minmers_raw(const char* seq, int64_t len, int k, int w, int s, int32_t seqId, int64_t chunk, int64_t warm,
                            orc_minmer_t* out, int64_t cap) {
  orc_mivec_t raw = {0, 0, 0};
  const int64_t npos = len - k + 1;
  if (chunk <= 0) add_minmers_stream_range(seq, len, k, w, s, seqId, &raw, 0, npos, 0, 1, 0);
  else
    for (int64_t cb = 0; cb < npos; cb += chunk) {
      const int64_t ce = cb + chunk < npos ? cb + chunk : npos;
      const int64_t rb = cb - warm > 0 ? cb - warm : 0;
      add_minmers_stream_range(seq, len, k, w, s, seqId, &raw, rb, ce, cb, ce == npos, 0);
    }
  for (int64_t i = 0; i < raw.n && i < cap; ++i) out[i] = raw.v[i];
  const int64_t n = raw.n;
  free(raw.v);
  return n;
}

/* =================================================================================================
 * Index build (Sketch::build, winSketch.hpp:266-429) over the concatenated per-sequence minmers.
 * ================================================================================================= */
static int cmp_u64(const void* a, const void* b) {
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : (x > y);
}
typedef struct { uint64_t hash; int64_t idx; } orc_hidx_t;
static int cmp_hidx(const void* a, const void* b) {
  const orc_hidx_t* x = (const orc_hidx_t*)a; const orc_hidx_t* y = (const orc_hidx_t*)b;
  if (x->hash != y->hash) return x->hash < y->hash ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* count_threshold of winSketch.hpp:298-349 given all hash frequencies (any order). */
uint64_t orc_count_threshold(const uint64_t* freqs, int64_t nuniq, uint64_t total_windows, double max_kmer_freq) {
  const uint64_t min_occ = 10;
  uint64_t thr;
  if (max_kmer_freq <= 1.0) { const uint64_t t = (uint64_t)(total_windows * max_kmer_freq); thr = t > min_occ ? t : min_occ; }
  else { const uint64_t t = (uint64_t)max_kmer_freq; thr = t > min_occ ? t : min_occ; }
  uint64_t wpos = 0, wuniq = 0;
  for (int64_t i = 0; i < nuniq; ++i) if (freqs[i] > thr && freqs[i] > min_occ) { wuniq++; wpos += freqs[i]; }
  if (wpos > total_windows / 2 || wuniq > nuniq * 0.7) {
    uint64_t* all = (uint64_t*)malloc((size_t)nuniq * sizeof(uint64_t));
    memcpy(all, freqs, (size_t)nuniq * sizeof(uint64_t));
    qsort(all, (size_t)nuniq, sizeof(uint64_t), cmp_u64);
    size_t keep = (size_t)(nuniq * 0.999);
    if ((int64_t)keep >= nuniq) keep = (size_t)nuniq - 1;
    if (all[keep] > thr) thr = all[keep];
    free(all);
  }
  return thr;
}

/* Builds the kept minmerIndex and the interval points grouped by hash (ascending hash; inside a hash the
 * reference's push order). partition_of_seq[seqId] = index of the worker thread whose contiguous range of
 * sequences holds seqId (winSketch.hpp:271-277,358-362): abutting intervals merge only inside a partition.
 * Outputs (caller-allocated, sizes n / 2n): kept[], points[], and per unique kept hash uhash/ustart/ucount.
 * Returns the number of kept minmers; *npoints, *nuniq, *threshold are set. */
int64_t orc_index_build(const orc_minmer_t* mi, int64_t n, const int32_t* partition_of_seq, double max_kmer_freq,
                        orc_minmer_t* kept, orc_ipoint_t* points, int64_t* npoints, uint64_t* uhash, int64_t* ustart,
                        int64_t* ucount, int64_t* nuniq_out, uint64_t* threshold) {
  orc_hidx_t* hs = (orc_hidx_t*)malloc((size_t)(n ? n : 1) * sizeof(orc_hidx_t));
  for (int64_t i = 0; i < n; ++i) { hs[i].hash = mi[i].hash; hs[i].idx = i; }
  qsort(hs, (size_t)n, sizeof(orc_hidx_t), cmp_hidx);
  /* frequencies */
  uint64_t* freqs = (uint64_t*)malloc((size_t)(n ? n : 1) * sizeof(uint64_t));
  int64_t nu = 0;
  for (int64_t i = 0; i < n;) { int64_t j = i; while (j < n && hs[j].hash == hs[i].hash) ++j; freqs[nu++] = (uint64_t)(j - i); i = j; }
  const uint64_t thr = orc_count_threshold(freqs, nu, (uint64_t)n, max_kmer_freq);
  *threshold = thr;
  /* kept flags in index order */
  char* keepf = (char*)calloc((size_t)(n ? n : 1), 1);
  { int64_t u = 0; for (int64_t i = 0; i < n;) { int64_t j = i; while (j < n && hs[j].hash == hs[i].hash) ++j;
      const int kp = !(freqs[u] > thr && freqs[u] > 10); for (int64_t t = i; t < j; ++t) keepf[hs[t].idx] = (char)kp; ++u; i = j; } }
  int64_t nk = 0;
  for (int64_t i = 0; i < n; ++i) if (keepf[i]) kept[nk++] = mi[i];
  /* postings, :379-387 */
  int64_t np = 0, nuq = 0;
  for (int64_t i = 0; i < n;) {
    int64_t j = i;
    while (j < n && hs[j].hash == hs[i].hash) ++j;
    if (keepf[hs[i].idx]) {
      uhash[nuq] = hs[i].hash; ustart[nuq] = np;
      int cur_part = -1; int64_t list_begin = np;
      for (int64_t t = i; t < j; ++t) {
        const orc_minmer_t* m = &mi[hs[t].idx];
        const int part = partition_of_seq ? partition_of_seq[m->seqId] : 0;
        if (part != cur_part) { cur_part = part; list_begin = np; } /* a new thread-local list starts */
        if (np == list_begin || points[np - 1].pos != m->wpos) {
          points[np].pos = m->wpos; points[np].hash = m->hash; points[np].seqId = m->seqId; points[np].side = 1; ++np;
          points[np].pos = m->wpos_end; points[np].hash = m->hash; points[np].seqId = m->seqId; points[np].side = -1; ++np;
        } else {
          points[np - 1].pos = m->wpos_end;
        }
      }
      ucount[nuq] = np - ustart[nuq];
      ++nuq;
    }
    i = j;
  }
  *npoints = np; *nuniq_out = nuq;
  free(hs); free(freqs); free(keepf);
  return nk;
}

/* =================================================================================================
 * L1: getSeedIntervalPoints (mappingCore.hpp:81-131) + computeL1CandidateRegions (:136-301) as driven by
 * Map::doL1Mapping (computeMap.hpp:945-983), for fragments of length == windowLength (windowLen = 0),
 * stage1_topANI_filter = stage2_full_scan = true (parse_args.hpp:701-702).
 * ================================================================================================= */
static int cmp_ipoint(const void* a, const void* b) {
  const orc_ipoint_t* x = (const orc_ipoint_t*)a; const orc_ipoint_t* y = (const orc_ipoint_t*)b;
  if (x->seqId != y->seqId) return x->seqId < y->seqId ? -1 : 1;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return x->side < y->side ? -1 : (x->side > y->side);
}

static void l1_regions(const orc_ipoint_t* ip, int64_t n, int minimumHits, int q_sketch, int param_sketch, const int* cutoffs,
                       int ncut, int window_len_param, orc_l1_locus_t* out, int* nout, int cap) {
  if (n == 0) return;
  int overlap = 0, best = 0;
  int64_t tr = 0, ld = 0;
  while (ld < n) { /* pass 1, :160-187 */
    while (tr < n && ((ip[tr].seqId == ip[ld].seqId && ip[tr].pos <= ip[ld].pos) || ip[tr].seqId < ip[ld].seqId)) {
      if (ip[tr].side == -1) overlap--;
      tr++;
    }
    const int64_t cur = ip[ld].pos;
    while (ld < n && ip[ld].pos == cur) { if (ip[ld].side == 1) overlap++; ld++; }
    if (overlap > best) best = overlap;
  }
  if (best < minimumHits) return;
  {
    const double div = param_sketch / 1000.0 > 1.0 ? param_sketch / 1000.0 : 1.0; /* skch::fixed::ss_table_max */
    int idx = (int)((best < q_sketch ? best : q_sketch) / div);
    if (idx >= ncut) idx = ncut - 1;
    if (cutoffs[idx] > minimumHits) minimumHits = cutoffs[idx];
  }
  /* pass 2, :203-284 */
  int in_cand = 0;
  orc_l1_locus_t cur_out; memset(&cur_out, 0, sizeof(cur_out));
  orc_l1_locus_t* local = (orc_l1_locus_t*)malloc((size_t)(n + 1) * sizeof(orc_l1_locus_t));
  int nlocal = 0;
  tr = 0; ld = 0; overlap = 0;
  int prevOverlap = 0;
  int32_t prev_seq = 0; int64_t prev_pos = 0; /* SeqCoord prevPos (uninitialised in the reference until first change) */
  int32_t cur_seq = ip[0].seqId; int64_t cur_pos = ip[0].pos;
  while (ld < n) {
    prevOverlap = overlap;
    while (tr < n && ((ip[tr].seqId == ip[ld].seqId && ip[tr].pos <= ip[ld].pos) || ip[tr].seqId < ip[ld].seqId)) {
      if (ip[tr].side == -1) overlap--;
      tr++;
    }
    if (ip[ld].pos != cur_pos) { prev_seq = cur_seq; prev_pos = cur_pos; cur_seq = ip[ld].seqId; cur_pos = ip[ld].pos; }
    while (ld < n && ip[ld].pos == cur_pos) { if (ip[ld].side == 1) overlap++; ld++; }
    if (prevOverlap >= minimumHits) {
      if (cur_out.seqId != prev_seq && in_cand) { local[nlocal++] = cur_out; memset(&cur_out, 0, sizeof(cur_out)); in_cand = 0; }
      if (!in_cand) {
        cur_out.rangeStartPos = prev_pos; cur_out.rangeEndPos = prev_pos; cur_out.seqId = prev_seq; cur_out.intersectionSize = prevOverlap;
        in_cand = 1;
      } else { /* stage2_full_scan */
        if (prevOverlap > cur_out.intersectionSize) cur_out.intersectionSize = prevOverlap;
        cur_out.rangeEndPos = prev_pos;
      }
    } else {
      if (in_cand) { local[nlocal++] = cur_out; memset(&cur_out, 0, sizeof(cur_out)); }
      in_cand = 0;
    }
  }
  if (in_cand) local[nlocal++] = cur_out;
  for (int i = 0; i < nlocal; ++i) { /* :287-300 */
    if (*nout == 0 || local[i].seqId != out[*nout - 1].seqId || local[i].rangeStartPos > out[*nout - 1].rangeEndPos + window_len_param) {
      if (*nout < cap) out[*nout] = local[i];
      (*nout)++;
    } else {
      out[*nout - 1].rangeEndPos = local[i].rangeEndPos;
      if (local[i].intersectionSize > out[*nout - 1].intersectionSize) out[*nout - 1].intersectionSize = local[i].intersectionSize;
    }
  }
  free(local);
}

int orc_l1_fragment(const uint64_t* uhash, const int64_t* ustart, const int64_t* ucount, int64_t nuniq, const orc_ipoint_t* points,
                    const uint64_t* q_hashes, int q_n, int32_t q_seq_id, int q_group, const int32_t* ref_group, int skip_self,
                    int skip_prefix, int lower_triangular, int minimum_hits, int param_sketch_size, int window_len,
                    const int* cutoffs, int ncut, orc_l1_locus_t* out, int cap) {
  int64_t total = 0;
  for (int i = 0; i < q_n; ++i) {
    int64_t lo = 0, hi = nuniq;
    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (uhash[mid] < q_hashes[i]) lo = mid + 1; else hi = mid; }
    if (lo < nuniq && uhash[lo] == q_hashes[i]) total += ucount[lo];
  }
  orc_ipoint_t* ip = (orc_ipoint_t*)malloc((size_t)(total + 1) * sizeof(orc_ipoint_t));
  int64_t n = 0;
  for (int i = 0; i < q_n; ++i) {
    int64_t lo = 0, hi = nuniq;
    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (uhash[mid] < q_hashes[i]) lo = mid + 1; else hi = mid; }
    if (!(lo < nuniq && uhash[lo] == q_hashes[i])) continue;
    for (int64_t t = 0; t < ucount[lo]; ++t) {
      const orc_ipoint_t p = points[ustart[lo] + t];
      const int tg = ref_group[p.seqId];
      int skip = 0;
      if (skip_self && q_group == tg) skip = 1;
      if (skip_prefix && q_group == tg) skip = 1;
      if (lower_triangular && q_seq_id <= p.seqId) skip = 1;
      if (!skip) ip[n++] = p;
    }
  }
  qsort(ip, (size_t)n, sizeof(orc_ipoint_t), cmp_ipoint); /* = the heap merge order up to ties of equal keys */
  int nout = 0;
  int64_t b = 0;
  while (b < n) { /* per PanSN group slice, computeMap.hpp:964-982 */
    int64_t e = n;
    if (skip_prefix) { e = b; const int g = ref_group[ip[b].seqId]; while (e < n && ref_group[ip[e].seqId] == g) ++e; }
    l1_regions(ip + b, e - b, minimum_hits, q_n, param_sketch_size, cutoffs, ncut, window_len, out, &nout, cap);
    b = e;
  }
  free(ip);
  return nout;
}

/* =================================================================================================
 * L2: SlideMapper (slidingMap.hpp:28-215) + computeL2MappedRegions (mappingCore.hpp:306-442) for fragments
 * of length == windowLength (windowLen = 0, so the hash_to_freq multiplicity map never gates anything), and
 * the part of Map::doL2Mapping / mapSingleQueryFrag that turns L1 loci into fragment mappings
 * (computeMap.hpp:895-921, 989-1061).
 * ================================================================================================= */
typedef struct { uint64_t hash_val; int q_strand; int strand_vote; unsigned num_before_inc; int active; } orc_slot_t;
typedef struct {
  orc_slot_t* v; int size;     /* slidingWindowMinhashes: size = sketchSize + 1, slot 0 is the zero dummy */
  int pivot; long pivRank;     /* pivot as an index into v */
  int sketchSize;
  int sharedSketchElements, strand_votes, intersectionSize;
  int64_t dup_inserts;         /* test probe: a matching insert into an already active slot */
} orc_slidemap_t;

static void slidemap_init(orc_slidemap_t* m, const orc_minmer_t* q, int q_n) { /* ctor + init(), :84-125 */
  m->size = q_n + 1;
  m->v = (orc_slot_t*)calloc((size_t)m->size, sizeof(orc_slot_t));
  for (int i = 0; i < q_n; ++i) { m->v[i + 1].hash_val = q[i].hash; m->v[i + 1].q_strand = q[i].strand; m->v[i + 1].num_before_inc = 1; }
  m->pivot = m->size - 1; m->pivRank = m->size - 1; m->sketchSize = q_n;
  m->sharedSketchElements = 0; m->strand_votes = 0; m->intersectionSize = 0; m->dup_inserts = 0;
}
static int slidemap_lower_bound(const orc_slidemap_t* m, uint64_t h) { /* over [1, size) */
  int lo = 1, hi = m->size;
  while (lo < hi) { const int mid = (lo + hi) / 2; if (m->v[mid].hash_val < h) lo = mid + 1; else hi = mid; }
  return lo;
}
static void slidemap_insert(orc_slidemap_t* m, uint64_t hash, int strand) { /* :129-170 */
  const int loc = slidemap_lower_bound(m, hash);
  if (loc == m->size) return;
  orc_slot_t* s = &m->v[loc];
  if (s->hash_val == hash) {
    if (s->active) m->dup_inserts++;
    s->active = 1;
    s->strand_vote = (int16_t)(s->strand_vote + s->q_strand * strand);
    m->intersectionSize++;
    if (s->hash_val <= m->v[m->pivot].hash_val) { m->sharedSketchElements++; m->strand_votes += s->strand_vote; }
  } else {
    s->num_before_inc++;
    if (s->hash_val <= m->v[m->pivot].hash_val) m->pivRank++;
    if (m->pivRank > m->sketchSize) {
      m->sharedSketchElements -= m->v[m->pivot].active;
      m->strand_votes -= m->v[m->pivot].strand_vote;
      m->pivRank -= m->v[m->pivot].num_before_inc;
      m->pivot--;
    }
  }
}
static void slidemap_delete(orc_slidemap_t* m, uint64_t hash) { /* :176-214 */
  const int loc = slidemap_lower_bound(m, hash);
  if (loc == m->size) return;
  orc_slot_t* s = &m->v[loc];
  if (s->hash_val == hash) {
    if (s->hash_val <= m->v[m->pivot].hash_val) { m->sharedSketchElements--; m->strand_votes -= s->strand_vote; }
    s->active = 0; s->strand_vote = 0; m->intersectionSize--;
  } else {
    s->num_before_inc--;
    if (s->hash_val <= m->v[m->pivot].hash_val) m->pivRank--;
    if (m->pivot + 1 != m->size && m->pivRank + m->v[m->pivot + 1].num_before_inc <= (unsigned long)m->sketchSize) {
      m->pivot++;
      m->sharedSketchElements += m->v[m->pivot].active;
      m->strand_votes += m->v[m->pivot].strand_vote;
      m->pivRank += m->v[m->pivot].num_before_inc;
    }
  }
}

/* min-heap on wpos_end (std::push_heap / pop_heap with heap_cmp, mappingCore.hpp:320). The order in which equal
 * wpos_end leave is immaterial: every delete of one window step happens before anything is read. */
typedef struct { int64_t wpos_end; uint64_t hash; } orc_hent_t;
typedef struct { orc_hent_t* a; int n, cap; } orc_heap2_t;
static void heap2_push(orc_heap2_t* h, orc_hent_t e) {
  if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 64; h->a = (orc_hent_t*)realloc(h->a, (size_t)h->cap * sizeof(orc_hent_t)); }
  int i = h->n++;
  while (i > 0) { const int p = (i - 1) / 2; if (h->a[p].wpos_end <= e.wpos_end) break; h->a[i] = h->a[p]; i = p; }
  h->a[i] = e;
}
static void heap2_pop(orc_heap2_t* h) {
  const orc_hent_t e = h->a[--h->n];
  int i = 0;
  for (;;) {
    int c = 2 * i + 1;
    if (c >= h->n) break;
    if (c + 1 < h->n && h->a[c + 1].wpos_end < h->a[c].wpos_end) ++c;
    if (h->a[c].wpos_end >= e.wpos_end) break;
    h->a[i] = h->a[c]; i = c;
  }
  if (h->n > 0) h->a[i] = e;
}

static void l2_close_candidate(orc_l2_locus_t* out, int* nout, int cap, orc_l2_locus_t* cur, int32_t seqId, int strand_votes, int w) {
  cur->meanOptimalPos = (cur->optimalStart + cur->optimalEnd) / 2;
  cur->seqId = seqId;
  cur->strand = strand_votes >= 0 ? 1 : -1;
  if (*nout == 0 || out[(*nout - 1) < cap ? (*nout - 1) : cap - 1].optimalEnd + w < cur->optimalStart) {
    if (*nout < cap) out[*nout] = *cur;
    (*nout)++;
  } else {
    orc_l2_locus_t* b = &out[*nout - 1];
    b->optimalEnd = cur->optimalEnd;
    b->meanOptimalPos = (b->optimalStart + b->optimalEnd) / 2;
  }
}

/* index[] = Sketch::minmerIndex (kept minmers, sorted by (seqId, wpos)); q[] = Q.minmerTableQuery (ascending hash),
 * q_n = Q.sketchSize. Returns the number of L2_mapLocus_t of this L1 locus (l2_vec_out); out must hold them all for
 * the merge rule to be exact (cap >= (rangeEnd-rangeStart)/w + 2 is always enough). */
int orc_l2_locus(const orc_minmer_t* index, int64_t n_index, const orc_minmer_t* q, int q_n, int window_len_param, int32_t seqId,
                 int64_t rangeStartPos, int64_t rangeEndPos, orc_l2_locus_t* out, int cap, int* best_intersection, int64_t* dup_inserts) {
  const int w = window_len_param;
  int64_t lo = 0, hi = n_index;
  { const int64_t key = rangeStartPos - w - 1; /* std::lower_bound with MinmerInfo::operator< = (seqId, wpos), :318-319 */
    while (lo < hi) { const int64_t mid = (lo + hi) / 2;
      if (index[mid].seqId < seqId || (index[mid].seqId == seqId && index[mid].wpos < key)) lo = mid + 1; else hi = mid; } }
  int64_t it = lo;
  orc_slidemap_t sm; slidemap_init(&sm, q, q_n);
  orc_heap2_t hp = {0, 0, 0};
  int bestSketchSize = 1, bestIntersectionSize = 0, in_candidate = 0, nout = 0;
  orc_l2_locus_t l2; memset(&l2, 0, sizeof(l2));
  while (it < n_index && index[it].seqId == seqId && index[it].wpos < rangeStartPos) { /* set up the window, :339-355 */
    if (index[it].wpos_end > rangeStartPos) {
      const orc_hent_t e = {index[it].wpos_end, index[it].hash};
      heap2_push(&hp, e);
      slidemap_insert(&sm, index[it].hash, index[it].strand);
    }
    it++;
  }
  while (it < n_index && index[it].seqId == seqId && index[it].wpos <= rangeEndPos) { /* :358-423 (windowLen = 0) */
    const int prev_strand_votes = sm.strand_votes;
    while (hp.n > 0 && hp.a[0].wpos_end <= index[it].wpos) { slidemap_delete(&sm, hp.a[0].hash); heap2_pop(&hp); }
    slidemap_insert(&sm, index[it].hash, index[it].strand);
    { const orc_hent_t e = {index[it].wpos_end, index[it].hash}; heap2_push(&hp, e); }
    if (sm.intersectionSize > bestIntersectionSize) bestIntersectionSize = sm.intersectionSize;
    if (sm.sharedSketchElements > bestSketchSize) {
      nout = 0; /* l2_vec_out.clear() */
      in_candidate = 1;
      bestSketchSize = sm.sharedSketchElements;
      l2.sharedSketchSize = sm.sharedSketchElements;
      l2.optimalStart = index[it].wpos;
      l2.optimalEnd = index[it].wpos;
    } else if (sm.sharedSketchElements == bestSketchSize) {
      if (!in_candidate) { l2.sharedSketchSize = sm.sharedSketchElements; l2.optimalStart = index[it].wpos; }
      in_candidate = 1;
      l2.optimalEnd = index[it].wpos;
    } else {
      if (in_candidate) {
        l2_close_candidate(out, &nout, cap, &l2, index[it].seqId, prev_strand_votes, w);
        memset(&l2, 0, sizeof(l2));
      }
      in_candidate = 0;
    }
    it++;
  }
  if (in_candidate) l2_close_candidate(out, &nout, cap, &l2, index[it - 1].seqId, sm.strand_votes, w);
  if (best_intersection) *best_intersection = bestIntersectionSize;
  if (dup_inserts) *dup_inserts += sm.dup_inserts;
  free(sm.v); free(hp.a);
  return nout;
}

/* Stat::j2md / md2j (map_stats.hpp:56-80) with the reference's float / double mix. */
float orc_j2md(float j, int k) {
  if (j == 0) return 1.0f;
  if (j == 1) return 0.0f;
  const float mash_dist = (float)(1 - pow((double)(2 * j / (1 + j)), 1.0 / k));
  return mash_dist;
}
float orc_md2j(float d, int k) {
  const float sim = 1 - d;
  const float jaccard = (float)(pow((double)sim, (double)k) / (2 - pow((double)sim, (double)k)));
  return jaccard;
}
/* 1 when an L1 locus passes the stage-1 top-ANI test of doL2Mapping (computeMap.hpp:999-1012) */
int orc_stage1_pass(double hg_numerator, float ani_diff, int kmer_size, int q_sketch_size, int intersection_size) {
  const double jaccardSimilarity = hg_numerator / q_sketch_size;
  const double mash_dist = orc_j2md((float)jaccardSimilarity, kmer_size);
  const double cutoff_ani = fmax(0.0, (1 - mash_dist) - ani_diff);
  const double cutoff_j = orc_md2j((float)(1 - cutoff_ani), kmer_size);
  const double candidateJaccard = (double)intersection_size / q_sketch_size;
  return !(candidateJaccard < cutoff_j);
}

static int cmp_l2map(const void* a, const void* b) {
  const orc_l2_mapping_t* x = (const orc_l2_mapping_t*)a; const orc_l2_mapping_t* y = (const orc_l2_mapping_t*)b;
  if (x->refSeqId != y->refSeqId) return x->refSeqId < y->refSeqId ? -1 : 1;
  if (x->refStartPos != y->refStartPos) return x->refStartPos < y->refStartPos ? -1 : 1;
  /* the reference's std::sort compares (refSeqId, refStartPos) only; ties are put in a fixed order here */
  if (x->optimalStart != y->optimalStart) return x->optimalStart < y->optimalStart ? -1 : 1;
  if (x->conservedSketches != y->conservedSketches) return x->conservedSketches < y->conservedSketches ? -1 : 1;
  return 0;
}

/* mapSingleQueryFrag's L2 half (computeMap.hpp:895-921) over the L1 loci of one fragment: per PanSN group slice the
 * stage-1 filter keeps the loci whose intersectionSize passes (the heap pops the largest first and stops at the
 * first failure, so the kept SET is every passing locus), computeL2MappedRegions per kept locus, the identity test
 * on sharedSketchSize (min_shared = smallest passing value for this Q.sketchSize, 0 = keep all), and the final sort
 * by (refSeqId, refStartPos). stage1: 0 = off. */
int orc_l2_fragment(const orc_minmer_t* index, int64_t n_index, const orc_minmer_t* q, int q_n, float kmer_complexity, int kmer_size,
                    int window_len_param, const orc_l1_locus_t* loci, int n_loci, int stage1, double hg_numerator, float ani_diff,
                    int min_shared, orc_l2_mapping_t* out, int cap) {
  int n = 0;
  const int tcap = 4096;
  orc_l2_locus_t* tmp = (orc_l2_locus_t*)malloc((size_t)tcap * sizeof(orc_l2_locus_t));
  for (int i = 0; i < n_loci; ++i) {
    if (stage1 && !orc_stage1_pass(hg_numerator, ani_diff, kmer_size, q_n, loci[i].intersectionSize)) continue;
    const int m = orc_l2_locus(index, n_index, q, q_n, window_len_param, loci[i].seqId, loci[i].rangeStartPos, loci[i].rangeEndPos, tmp, tcap, 0, 0);
    for (int j = 0; j < m && j < tcap; ++j) {
      if (tmp[j].sharedSketchSize < min_shared) continue;
      const float mash_dist = orc_j2md((float)(1.0 * tmp[j].sharedSketchSize / q_n), kmer_size);
      orc_l2_mapping_t r;
      memset(&r, 0, sizeof(r));
      r.refSeqId = tmp[j].seqId; r.refStartPos = tmp[j].meanOptimalPos; r.optimalStart = tmp[j].optimalStart; r.optimalEnd = tmp[j].optimalEnd;
      r.conservedSketches = tmp[j].sharedSketchSize; r.strand = tmp[j].strand; r.nucIdentity = 1 - mash_dist; r.kmerComplexity = kmer_complexity;
      if (n < cap) out[n] = r;
      ++n;
    }
  }
  free(tmp);
  qsort(out, (size_t)(n < cap ? n : cap), sizeof(orc_l2_mapping_t), cmp_l2map);
  return n;
}

/* ---------------------------------------------------------------------------------------------------
 * ANI auto-identity sketch: the per-sequence part of skch::Stat::estimate_identity_for_groups
 * (src/map/include/map_stats.hpp:563-613) stated literally — upper-case on the fly, `ambig_kmer_count` state machine
 * (initial scan of the first k bases sets it to k whatever the position of the bad base), canonical Murmur3 hash,
 * StreamingMinHash::add_unsafe (streamingMinHash.hpp:89-99: a max-heap of `ssize` hashes, duplicates kept, a new hash
 * replaces the top only when strictly smaller). heap[] is the caller's heap (may already hold hashes: merging a group
 * is the same add loop, map_stats.hpp:617-637). Returns the new heap size.
 * ------------------------------------------------------------------------------------------------- */
static void ani_heap_add(uint64_t* heap, int* n, int ssize, uint64_t h) {
  if (*n < ssize) { /* push + sift up */
    int i = (*n)++;
    heap[i] = h;
    while (i > 0 && heap[(i - 1) / 2] < heap[i]) { uint64_t t = heap[i]; heap[i] = heap[(i - 1) / 2]; heap[(i - 1) / 2] = t; i = (i - 1) / 2; }
  } else if (h < heap[0]) { /* replace the maximum + sift down */
    int i = 0;
    heap[0] = h;
    for (;;) {
      int l = 2 * i + 1, r = l + 1, m = i;
      if (l < *n && heap[l] > heap[m]) m = l;
      if (r < *n && heap[r] > heap[m]) m = r;
      if (m == i) break;
      uint64_t t = heap[i]; heap[i] = heap[m]; heap[m] = t;
      i = m;
    }
  }
}

static char ani_upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
static int ani_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

int orc_ani_add_sequence(const char* seq, int64_t len, int k, int ssize, uint64_t* heap, int heap_n) {
  char fwd[64], rev[64];
  int ambig = 0;
  for (int j = 0; j < k && j < len; j++)
    if (!ani_acgt(ani_upper(seq[j]))) { ambig = k; break; }
  for (int64_t i = 0; i <= len - k; i++) {
    if (!ani_acgt(ani_upper(seq[i + k - 1]))) ambig = k;
    if (ambig == 0) {
      for (int j = 0; j < k; j++) fwd[j] = ani_upper(seq[i + j]);
      revcomp(fwd, rev, k);
      const uint64_t hf = orc_kmer_hash(fwd, k), hb = orc_kmer_hash(rev, k);
      if (hf != hb) ani_heap_add(heap, &heap_n, ssize, hf < hb ? hf : hb);
    }
    if (ambig > 0) ambig--;
  }
  return heap_n;
}

/* StreamingMinHash::add_unsafe for one hash (group merge) */
int orc_ani_add_hash(uint64_t h, int ssize, uint64_t* heap, int heap_n) {
  ani_heap_add(heap, &heap_n, ssize, h);
  return heap_n;
}
