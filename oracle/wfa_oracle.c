/*
 * wfa_oracle.c — plain-C restatement of the reference's gap-affine-2p WFA / biWFA
 * (TEST INFRASTRUCTURE ONLY; see oracle.h). Sequential, one alignment at a time, written for
 * clarity: every wavefront is an explicit [lo,hi] range + offsets array, out-of-range reads
 * return OFFSET_NULL (what the reference achieves with wavefront_compute_init_ends,
 * wavefront_compute.c:498-575), and the reverse aligner indexes the original sequences
 * back-to-front instead of copying reversed buffers (wavefront_sequences.c:83-101,274-309).
 *
 * Reference anchors (all under /root/reference/deps/WFA2-lib/wavefront/):
 *   compute      wavefront_compute_affine2p.c:45-106,334-368; wavefront_compute.c:40-86,306-352,
 *                409-494,579-632
 *   extend       wavefront_extend.c:86-211; wavefront_extend_kernels.c:68-152
 *   termination  wavefront_termination.c:37-114
 *   init         wavefront_aligner.c:252-420
 *   biWFA        wavefront_bialign.c:52-54,159-189,508-571,828-955,974-1221,1266-1293
 *   unialign     wavefront_unialign.c:147-273
 *   backtrace    wavefront_backtrace.c:49-219,320-529
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define OFFSET_NULL (INT32_MIN / 2) /* wavefront_offset.h:44 */
#define MAXI(a, b) ((a) > (b) ? (a) : (b))
#define MINI(a, b) ((a) < (b) ? (a) : (b))

enum { C_M = 0, C_I1 = 1, C_I2 = 2, C_D1 = 3, C_D2 = 4 };

enum { ST_OK = 0, ST_END_REACHED = 1, ST_END_UNREACHABLE = -1, ST_UNATTAINABLE = -3 };

/* wavefront_bialign.c:52-54 */
#define FALLBACK_MIN_SCORE 250
#define FALLBACK_MIN_LENGTH 100
#define RECOVERY_MIN_SCORE 500

typedef struct {
  int exists; /* reference: slot pointer != NULL */
  int lo, hi; /* lo > hi  <=>  reference ->null   */
  int kmin;   /* off[k - kmin]                    */
  int cap;
  int32_t* off;
} wf_t;

typedef struct {
  /* sequence view */
  const char* P;
  const char* T;
  int pb, pe, tb, te, rev;
  int plen, tlen;
  orc_penalties_t pen;
  int scope; /* max_score_scope, wavefront_components.c:101-112 */
  int cbegin, cend;
  int modular;
  int nslots;
  wf_t* wf[5];
  int num_null_steps;
  int status, status_score;
  int end_k, end_off;
  int endsfree, pbf, pef, tbf, tef, term_group; /* alignment form (wavefront_aligner.c:252-310) */
  orc_wfa_counters_t* cnt;
} wfa_t;

static const wf_t WF_NULL = {0, 1, -1, 0, 0, 0}; /* wavefront_init_null: lo=1, hi=-1 */

static inline char pch(const wfa_t* a, int v) { return a->rev ? a->P[a->pe - 1 - v] : a->P[a->pb + v]; }
static inline char tch(const wfa_t* a, int h) { return a->rev ? a->T[a->te - 1 - h] : a->T[a->tb + h]; }

static void wfa_alloc(wfa_t* a, int modular, int nslots) {
  memset(a, 0, sizeof(*a));
  a->modular = modular;
  a->nslots = nslots;
  for (int c = 0; c < 5; ++c) a->wf[c] = (wf_t*)calloc((size_t)nslots, sizeof(wf_t));
}
static void wfa_free(wfa_t* a) {
  for (int c = 0; c < 5; ++c) {
    for (int i = 0; i < a->nslots; ++i) free(a->wf[c][i].off);
    free(a->wf[c]);
  }
}
static void wfa_grow(wfa_t* a, int need) {
  if (need <= a->nslots) return;
  int n = a->nslots;
  while (n < need) n *= 2;
  for (int c = 0; c < 5; ++c) {
    a->wf[c] = (wf_t*)realloc(a->wf[c], (size_t)n * sizeof(wf_t));
    memset(a->wf[c] + a->nslots, 0, (size_t)(n - a->nslots) * sizeof(wf_t));
  }
  a->nslots = n;
}
static inline int slot_of(const wfa_t* a, int score) { return a->modular ? score % a->scope : score; }

static void wf_set_range(wf_t* w, int lo, int hi) {
  const int n = hi - lo + 1;
  if (n > w->cap) {
    free(w->off);
    w->cap = n + 16;
    w->off = (int32_t*)malloc((size_t)w->cap * sizeof(int32_t));
  }
  w->exists = 1;
  w->lo = lo;
  w->hi = hi;
  w->kmin = lo;
}
static inline int32_t wf_get(const wf_t* w, int k) {
  return (k < w->lo || k > w->hi) ? OFFSET_NULL : w->off[k - w->kmin];
}

/* wavefront_compute_get_*wavefront, wavefront_compute.c:266-305 */
static const wf_t* fetch(const wfa_t* a, int comp, int score) {
  if (score < 0) return &WF_NULL;
  const wf_t* w = &a->wf[comp][slot_of(a, score)];
  if (!w->exists || w->lo > w->hi) return &WF_NULL;
  return w;
}

/* wavefront_compute_trim_ends, wavefront_compute.c:579-613 */
static void trim_ends(const wfa_t* a, wf_t* w) {
  int k;
  const int lo = w->lo;
  for (k = w->hi; k >= lo; --k) {
    const int32_t off = w->off[k - w->kmin];
    const uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
    if (h <= (uint32_t)a->tlen && v <= (uint32_t)a->plen) break;
  }
  w->hi = k;
  const int hi = w->hi;
  for (k = w->lo; k <= hi; ++k) {
    const int32_t off = w->off[k - w->kmin];
    const uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
    if (h <= (uint32_t)a->tlen && v <= (uint32_t)a->plen) break;
  }
  w->lo = k;
}

/* wavefront_aligner_init_wf, wavefront_aligner.c:252-383 (end-to-end only) */
static void wfa_init(wfa_t* a, const char* P, int pb, int pe, const char* T, int tb, int te, int rev,
                     int cbegin, int cend) {
  a->P = P; a->T = T; a->pb = pb; a->pe = pe; a->tb = tb; a->te = te; a->rev = rev;
  a->plen = pe - pb;
  a->tlen = te - tb;
  a->cbegin = cbegin;
  a->cend = cend;
  a->num_null_steps = 0;
  a->status = ST_OK;
  a->status_score = 0;
  a->end_k = INT_MAX;
  a->end_off = OFFSET_NULL;
  for (int c = 0; c < 5; ++c)
    for (int i = 0; i < a->nslots; ++i) a->wf[c][i].exists = 0;
  wf_t* w0 = &a->wf[cbegin][0];
  wf_set_range(w0, 0, 0);
  w0->off[0] = 0;
}

/* wavefront_compute_affine2p, wavefront_compute_affine2p.c:334-368 */
static void wfa_compute(wfa_t* a, int score) {
  const orc_penalties_t* p = &a->pen;
  const wf_t* m_misms = fetch(a, C_M, score - p->x);
  const wf_t* m_open1 = fetch(a, C_M, score - p->o1 - p->e1);
  const wf_t* m_open2 = fetch(a, C_M, score - p->o2 - p->e2);
  const wf_t* i1_ext = fetch(a, C_I1, score - p->e1);
  const wf_t* i2_ext = fetch(a, C_I2, score - p->e2);
  const wf_t* d1_ext = fetch(a, C_D1, score - p->e1);
  const wf_t* d2_ext = fetch(a, C_D2, score - p->e2);
  if (!a->modular) wfa_grow(a, score + 1);
  const int slot = slot_of(a, score);
  wf_t* out[5];
  for (int c = 0; c < 5; ++c) out[c] = &a->wf[c][slot];
  /* re-fetch after a possible realloc in wfa_grow */
  m_misms = fetch(a, C_M, score - p->x);
  m_open1 = fetch(a, C_M, score - p->o1 - p->e1);
  m_open2 = fetch(a, C_M, score - p->o2 - p->e2);
  i1_ext = fetch(a, C_I1, score - p->e1);
  i2_ext = fetch(a, C_I2, score - p->e2);
  d1_ext = fetch(a, C_D1, score - p->e1);
  d2_ext = fetch(a, C_D2, score - p->e2);
  const int n_m = (m_misms == &WF_NULL), n_o1 = (m_open1 == &WF_NULL), n_o2 = (m_open2 == &WF_NULL);
  const int n_i1 = (i1_ext == &WF_NULL), n_i2 = (i2_ext == &WF_NULL);
  const int n_d1 = (d1_ext == &WF_NULL), n_d2 = (d2_ext == &WF_NULL);
  if (n_m && n_o1 && n_o2 && n_i1 && n_i2 && n_d1 && n_d2) {
    a->num_null_steps++;
    for (int c = 0; c < 5; ++c) out[c]->exists = 0; /* wavefront_compute_allocate_output_null */
    return;
  }
  a->num_null_steps = 0;
  /* wavefront_compute_limits_input, wavefront_compute.c:40-86 (null inputs contribute lo=1,hi=-1) */
  int lo = m_misms->lo, hi = m_misms->hi;
  if (lo > m_open1->lo - 1) lo = m_open1->lo - 1;
  if (hi < m_open1->hi + 1) hi = m_open1->hi + 1;
  if (lo > i1_ext->lo + 1) lo = i1_ext->lo + 1;
  if (hi < i1_ext->hi + 1) hi = i1_ext->hi + 1;
  if (lo > d1_ext->lo - 1) lo = d1_ext->lo - 1;
  if (hi < d1_ext->hi - 1) hi = d1_ext->hi - 1;
  if (lo > m_open2->lo - 1) lo = m_open2->lo - 1;
  if (hi < m_open2->hi + 1) hi = m_open2->hi + 1;
  if (lo > i2_ext->lo + 1) lo = i2_ext->lo + 1;
  if (hi < i2_ext->hi + 1) hi = i2_ext->hi + 1;
  if (lo > d2_ext->lo - 1) lo = d2_ext->lo - 1;
  if (hi < d2_ext->hi - 1) hi = d2_ext->hi - 1;
  /* wavefront_compute_allocate_output, wavefront_compute.c:409-494.
   * The inputs may alias the output slot only in modular mode for score-scope == slot, which
   * cannot happen (all input scores are in (score-scope, score)). */
  const int ex_i1 = !n_o1 || !n_i1, ex_d1 = !n_o1 || !n_d1;
  const int ex_i2 = !n_o2 || !n_i2, ex_d2 = !n_o2 || !n_d2;
  wf_set_range(out[C_M], lo, hi);
  if (ex_i1) wf_set_range(out[C_I1], lo, hi); else out[C_I1]->exists = 0;
  if (ex_i2) wf_set_range(out[C_I2], lo, hi); else out[C_I2]->exists = 0;
  if (ex_d1) wf_set_range(out[C_D1], lo, hi); else out[C_D1]->exists = 0;
  if (ex_d2) wf_set_range(out[C_D2], lo, hi); else out[C_D2]->exists = 0;
  /* wavefront_compute_affine2p_idm, wavefront_compute_affine2p.c:45-106 */
  for (int k = lo; k <= hi; ++k) {
    const int32_t ins1 = MAXI(wf_get(m_open1, k - 1), wf_get(i1_ext, k - 1)) + 1;
    const int32_t ins2 = MAXI(wf_get(m_open2, k - 1), wf_get(i2_ext, k - 1)) + 1;
    const int32_t ins = MAXI(ins1, ins2);
    const int32_t del1 = MAXI(wf_get(m_open1, k + 1), wf_get(d1_ext, k + 1));
    const int32_t del2 = MAXI(wf_get(m_open2, k + 1), wf_get(d2_ext, k + 1));
    const int32_t del = MAXI(del1, del2);
    const int32_t misms = wf_get(m_misms, k) + 1;
    int32_t mx = MAXI(del, MAXI(misms, ins));
    const uint32_t h = (uint32_t)mx, v = (uint32_t)(mx - k);
    if (h > (uint32_t)a->tlen) mx = OFFSET_NULL;
    if (v > (uint32_t)a->plen) mx = OFFSET_NULL;
    out[C_M]->off[k - lo] = mx;
    if (ex_i1) out[C_I1]->off[k - lo] = ins1;
    if (ex_i2) out[C_I2]->off[k - lo] = ins2;
    if (ex_d1) out[C_D1]->off[k - lo] = del1;
    if (ex_d2) out[C_D2]->off[k - lo] = del2;
  }
  if (a->cnt) a->cnt->cells += hi - lo + 1;
  /* wavefront_compute_process_ends, wavefront_compute.c:614-632 */
  for (int c = 0; c < 5; ++c)
    if (out[c]->exists) trim_ends(a, out[c]);
}

/* wavefront_termination_end2end, wavefront_termination.c:37-114 */
static int termination_end2end(wfa_t* a, int score) {
  const int alignment_k = a->tlen - a->plen;
  const int alignment_offset = a->tlen;
  const wf_t* w = &a->wf[a->cend][slot_of(a, score)];
  if (!w->exists || w->lo > alignment_k || alignment_k > w->hi) return 0;
  if (w->off[alignment_k - w->kmin] < alignment_offset) return 0;
  a->end_k = alignment_k;
  a->end_off = alignment_offset;
  return 1;
}

/* wavefront_extend_end2end(_max), wavefront_extend.c:86-211 + kernels :101-152.
 * Returns 1 when the alignment is finished (status set). */
static int wfa_extend(wfa_t* a, int score, int* max_ak) {
  if (max_ak) *max_ak = 0;
  wf_t* m = &a->wf[C_M][slot_of(a, score)];
  if (!m->exists) {
    if (a->num_null_steps > a->scope) {
      a->status = ST_END_UNREACHABLE;
      a->status_score = score;
      return 1;
    }
    return 0;
  }
  int mx = 0;
  int64_t matched = 0;
  for (int k = m->lo; k <= m->hi; ++k) {
    int32_t off = m->off[k - m->kmin];
    if (off == OFFSET_NULL) continue;
    int v = off - k, h = off;
    while (v < a->plen && h < a->tlen && pch(a, v) == tch(a, h)) { ++v; ++h; ++off; ++matched; }
    m->off[k - m->kmin] = off;
    const int ak = 2 * off - k;
    if (mx < ak) mx = ak;
  }
  if (a->cnt) a->cnt->extend_matches += matched;
  if (a->endsfree) return 0; /* handled by wfa_extend_endsfree */
  if (termination_end2end(a, score)) {
    a->status = ST_END_REACHED;
    a->status_score = score;
    return 1;
  }
  if (max_ak) *max_ak = mx;
  return 0;
}

/* ---- breakpoint detection ------------------------------------------------------------------ */

typedef struct {
  int score, score_forward, score_reverse;
  int k_forward, k_reverse, offset_forward, offset_reverse;
  int component;
} breakpoint_t;

/* wavefront_bialign_breakpoint_indel2indel (:508-571) and _m2m (:828-872); is_m selects m2m. */
static void breakpoint_scan(const wfa_t* a0, int bp_forward, int score_0, int score_1, const wf_t* w0,
                            const wf_t* w1, int component, int is_m, breakpoint_t* bp) {
  const int tlen = a0->tlen, plen = a0->plen;
  const orc_penalties_t* p = &a0->pen;
  const int gap_open = is_m ? 0 : ((component == C_I1 || component == C_D1) ? p->o1 : p->o2);
  const int lo_0 = w0->lo, hi_0 = w0->hi;
  const int lo_1 = (tlen - plen) - w1->hi, hi_1 = (tlen - plen) - w1->lo;
  if (hi_1 < lo_0 || hi_0 < lo_1) return;
  const int min_hi = MINI(hi_0, hi_1), max_lo = MAXI(lo_0, lo_1);
  if (!is_m && score_0 + score_1 - gap_open >= bp->score) return;
  if (a0->cnt && min_hi >= max_lo) a0->cnt->overlap_tests += min_hi - max_lo + 1;
  for (int k_0 = max_lo; k_0 <= min_hi; ++k_0) {
    const int k_1 = (tlen - plen) - k_0;
    const int32_t off_0 = w0->off[k_0 - w0->kmin];
    const int32_t off_1 = w1->off[k_1 - w1->kmin];
    if (off_0 + off_1 >= tlen) {
      if (!is_m) {
        /* out-of-bounds check on the forward-side coordinates (:540-556) */
        const int kk = bp_forward ? k_0 : k_1;
        const int oo = bp_forward ? off_0 : off_1;
        const int v = oo - kk, h = oo;
        if (v > plen || h > tlen) continue;
      }
      if (bp_forward) {
        bp->score_forward = score_0; bp->score_reverse = score_1;
        bp->k_forward = k_0; bp->k_reverse = k_1;
        bp->offset_forward = off_0; bp->offset_reverse = off_1;
      } else {
        bp->score_forward = score_1; bp->score_reverse = score_0;
        bp->k_forward = k_1; bp->k_reverse = k_0;
        bp->offset_forward = off_1; bp->offset_reverse = off_0;
      }
      bp->score = score_0 + score_1 - gap_open;
      bp->component = component;
      return;
    }
  }
}

/* wavefront_bialign_overlap, wavefront_bialign.c:877-955 */
static void bialign_overlap(const wfa_t* a0, const wfa_t* a1, int score_0, int score_1, int bp_forward,
                            breakpoint_t* bp) {
  const int scope = a0->scope;
  const orc_penalties_t* p = &a0->pen;
  const int s0 = score_0 % scope;
  const wf_t* m0 = &a0->wf[C_M][s0];
  if (!m0->exists) return;
  const wf_t* d1_0 = &a0->wf[C_D1][s0];
  const wf_t* i1_0 = &a0->wf[C_I1][s0];
  const wf_t* d2_0 = &a0->wf[C_D2][s0];
  const wf_t* i2_0 = &a0->wf[C_I2][s0];
  for (int i = 0; i < scope; ++i) {
    const int score_i = score_1 - i;
    if (score_i < 0) break;
    const int si = score_i % scope;
    if (score_0 + score_i - p->o2 >= bp->score) continue;
    const wf_t* d2_1 = &a1->wf[C_D2][si];
    if (d2_0->exists && d2_1->exists) breakpoint_scan(a0, bp_forward, score_0, score_i, d2_0, d2_1, C_D2, 0, bp);
    const wf_t* i2_1 = &a1->wf[C_I2][si];
    if (i2_0->exists && i2_1->exists) breakpoint_scan(a0, bp_forward, score_0, score_i, i2_0, i2_1, C_I2, 0, bp);
    if (score_0 + score_i - p->o1 >= bp->score) continue;
    const wf_t* d1_1 = &a1->wf[C_D1][si];
    if (d1_0->exists && d1_1->exists) breakpoint_scan(a0, bp_forward, score_0, score_i, d1_0, d1_1, C_D1, 0, bp);
    const wf_t* i1_1 = &a1->wf[C_I1][si];
    if (i1_0->exists && i1_1->exists) breakpoint_scan(a0, bp_forward, score_0, score_i, i1_0, i1_1, C_I1, 0, bp);
    if (score_0 + score_i >= bp->score) continue;
    const wf_t* m1 = &a1->wf[C_M][si];
    if (m1->exists) breakpoint_scan(a0, bp_forward, score_0, score_i, m0, m1, C_M, 1, bp);
  }
}

/* wavefront_bialign_find_breakpoint, wavefront_bialign.c:974-1082 */
static int find_breakpoint(wfa_t* f, wfa_t* r, breakpoint_t* bp) {
  const int max_antidiagonal = f->plen + f->tlen - 1;
  int score_forward = 0, score_reverse = 0, forward_max_ak = 0, reverse_max_ak = 0;
  bp->score = INT_MAX;
  if (wfa_extend(f, 0, &forward_max_ak)) return f->status;
  if (wfa_extend(r, 0, &reverse_max_ak)) return r->status;
  int max_ak = 0, last_wf_forward = 0;
  for (;;) {
    if (forward_max_ak + reverse_max_ak >= max_antidiagonal) break;
    ++score_forward;
    wfa_compute(f, score_forward);
    int quit = wfa_extend(f, score_forward, &max_ak);
    if (forward_max_ak < max_ak) forward_max_ak = max_ak;
    last_wf_forward = 1;
    if (quit) return f->status;
    if (forward_max_ak + reverse_max_ak >= max_antidiagonal) break;
    ++score_reverse;
    wfa_compute(r, score_reverse);
    quit = wfa_extend(r, score_reverse, &max_ak);
    if (reverse_max_ak < max_ak) reverse_max_ak = max_ak;
    last_wf_forward = 0;
    if (quit) return r->status;
  }
  const int scope = f->scope;
  const int gap_opening = MAXI(f->pen.o1, f->pen.o2);
  for (;;) {
    if (last_wf_forward) {
      const int min_score_reverse = (score_reverse > scope - 1) ? score_reverse - (scope - 1) : 0;
      if (score_forward + min_score_reverse - gap_opening >= bp->score) break;
      bialign_overlap(f, r, score_forward, score_reverse, 1, bp);
      ++score_reverse;
      wfa_compute(r, score_reverse);
      if (wfa_extend(r, score_reverse, NULL)) return r->status;
    }
    const int min_score_forward = (score_forward > scope - 1) ? score_forward - (scope - 1) : 0;
    if (min_score_forward + score_reverse - gap_opening >= bp->score) break;
    bialign_overlap(r, f, score_reverse, score_forward, 0, bp);
    ++score_forward;
    wfa_compute(f, score_forward);
    if (wfa_extend(f, score_forward, NULL)) return f->status;
    last_wf_forward = 1;
  }
  if (f->cnt) f->cnt->score_steps += score_forward + score_reverse;
  return ST_OK;
}

/* wavefront_termination_endsfree, wavefront_termination.c:115-160 */
static int termination_endsfree(const wfa_t* a, int k, int32_t offset) {
  const int h = offset, v = offset - k;
  if (h >= a->tlen && a->plen - v <= a->pef) return 1;
  if (v >= a->plen && a->tlen - h <= a->tef) return 1;
  return 0;
}

/* wavefront_extend_endsfree (wavefront_extend.c:259-293). All offsets end up fully extended; WHICH
 * terminating cell ends the alignment depends on the build of the reference:
 *   scalar (wavefront_extend_kernels.c:166-193): the first in ascending k;
 *   AVX2 / AVX-512 (wavefront_extend_kernels_avx.c:296-400, 592-691), G = 8 / 16 lanes: first the
 *   n % G lowest diagonals in order, then - group by group - the cells whose first 4 bases matched, then
 *   a final ascending pass over all cells.  term_group = 1, 8 or 16 selects the rule. */
static int wfa_extend_endsfree(wfa_t* a, int score) {
  wf_t* m = &a->wf[C_M][slot_of(a, score)];
  if (!m->exists) {
    if (a->num_null_steps > a->scope) { a->status = ST_END_UNREACHABLE; a->status_score = score; return 1; }
    return 0;
  }
  const int G = a->term_group, n = m->hi - m->lo + 1;
  const int peel = (G > 1) ? (n < G ? n : n % G) : n;
  int best_class = 3, best_j = 0;
  for (int k = m->lo; k <= m->hi; ++k) {
    int32_t off = m->off[k - m->kmin];
    if (off < 0) continue;
    int v = off - k, h = off, run = 0;
    while (v < a->plen && h < a->tlen && pch(a, v) == tch(a, h)) { ++v; ++h; ++off; ++run; }
    m->off[k - m->kmin] = off;
    if (a->cnt) a->cnt->extend_matches += run;
    if (termination_endsfree(a, k, off)) {
      const int j = k - m->lo;
      const int cls = (G <= 1 || j < peel) ? 0 : (run >= 4 ? 1 : 2);
      if (cls < best_class) { best_class = cls; best_j = j; }
    }
  }
  if (best_class < 3) {
    a->end_k = m->lo + best_j;
    a->end_off = m->off[best_j + m->lo - m->kmin];
    a->status = ST_END_REACHED;
    a->status_score = score;
    return 1;
  }
  return 0;
}

/* wavefront_aligner_init_wf_m with ends-free begin (wavefront_aligner.c:252-310), match == 0 */
static void wfa_init_endsfree(wfa_t* a, const char* P, int plen, const char* T, int tlen, int pbf, int pef, int tbf, int tef,
                              int term_group) {
  wfa_init(a, P, 0, plen, T, 0, tlen, 0, C_M, C_M);
  a->endsfree = 1; a->pbf = pbf; a->pef = pef; a->tbf = tbf; a->tef = tef; a->term_group = term_group;
  wf_t* w0 = &a->wf[C_M][0];
  wf_set_range(w0, -pbf, tbf);
  for (int k = -pbf; k <= tbf; ++k) w0->off[k + pbf] = k > 0 ? k : 0; /* (h,0) -> offset h ; (0,v) -> offset 0 */
}

/* ---- base case: unidirectional WFA + backtrace --------------------------------------------- */

typedef struct {
  char* ops;
  int cap, len;
} opsbuf_t;

static void ops_push(opsbuf_t* o, const char* src, int n) {
  if (o->len + n > o->cap) { o->len = o->cap + 1; return; } /* overflow marker */
  memcpy(o->ops + o->len, src, (size_t)n);
  o->len += n;
}
static void ops_fill(opsbuf_t* o, char c, int n) {
  if (o->len + n > o->cap) { o->len = o->cap + 1; return; }
  memset(o->ops + o->len, c, (size_t)n);
  o->len += n;
}

/* backtrace type codes, wavefront_backtrace.c:49-59 */
enum { BT_M = 9, BT_D2_EXT = 8, BT_D2_OPEN = 7, BT_D1_EXT = 6, BT_D1_OPEN = 5, BT_I2_EXT = 4, BT_I2_OPEN = 3,
       BT_I1_EXT = 2, BT_I1_OPEN = 1 };
#define BT_SET(off, type) ((((int64_t)(off)) << 4) | (type))

/* wavefront_backtrace_{misms,ins*,del*}, wavefront_backtrace.c:64-219 (full-memory aligner) */
static int64_t bt_src(const wfa_t* a, int comp, int score, int k, int dk, int plus, int type) {
  if (score < 0) return OFFSET_NULL;
  if (score >= a->nslots) return OFFSET_NULL;
  const wf_t* w = &a->wf[comp][score];
  if (w->exists && w->lo <= k + dk && k + dk <= w->hi) return BT_SET(w->off[k + dk - w->kmin] + plus, type);
  return OFFSET_NULL;
}

/* wavefront_backtrace_affine, wavefront_backtrace.c:320-529. Writes ops right-to-left into tmp. */
static int backtrace_affine(const wfa_t* a, int alignment_score, int alignment_k, int alignment_offset,
                            char* tmp, int tmpcap, int* begin_out) {
  const orc_penalties_t* p = &a->pen;
  int pos = tmpcap - 1; /* next write position (begin_offset) */
  int matrix_type = a->cend;
  int score = alignment_score, k = alignment_k;
  int h = alignment_offset, v = alignment_offset - alignment_k;
  int offset = alignment_offset;
  if (a->cend == C_M) {
    for (int i = a->plen - v; i > 0; --i) tmp[pos--] = 'D';
    for (int i = a->tlen - h; i > 0; --i) tmp[pos--] = 'I';
  }
  while (v > 0 && h > 0 && score > 0) {
    const int mismatch = score - p->x;
    const int gap_open1 = score - p->o1 - p->e1, gap_open2 = score - p->o2 - p->e2;
    const int gap_extend1 = score - p->e1, gap_extend2 = score - p->e2;
    int64_t max_all;
    switch (matrix_type) {
      case C_M: {
        const int64_t misms = bt_src(a, C_M, mismatch, k, 0, 1, BT_M);
        const int64_t max_ins1 = MAXI(bt_src(a, C_M, gap_open1, k, -1, 1, BT_I1_OPEN), bt_src(a, C_I1, gap_extend1, k, -1, 1, BT_I1_EXT));
        const int64_t max_del1 = MAXI(bt_src(a, C_M, gap_open1, k, +1, 0, BT_D1_OPEN), bt_src(a, C_D1, gap_extend1, k, +1, 0, BT_D1_EXT));
        const int64_t max_ins2 = MAXI(bt_src(a, C_M, gap_open2, k, -1, 1, BT_I2_OPEN), bt_src(a, C_I2, gap_extend2, k, -1, 1, BT_I2_EXT));
        const int64_t max_del2 = MAXI(bt_src(a, C_M, gap_open2, k, +1, 0, BT_D2_OPEN), bt_src(a, C_D2, gap_extend2, k, +1, 0, BT_D2_EXT));
        const int64_t max_ins = MAXI(max_ins1, max_ins2), max_del = MAXI(max_del1, max_del2);
        max_all = MAXI(misms, MAXI(max_ins, max_del));
        break;
      }
      case C_I1: max_all = MAXI(bt_src(a, C_M, gap_open1, k, -1, 1, BT_I1_OPEN), bt_src(a, C_I1, gap_extend1, k, -1, 1, BT_I1_EXT)); break;
      case C_I2: max_all = MAXI(bt_src(a, C_M, gap_open2, k, -1, 1, BT_I2_OPEN), bt_src(a, C_I2, gap_extend2, k, -1, 1, BT_I2_EXT)); break;
      case C_D1: max_all = MAXI(bt_src(a, C_M, gap_open1, k, +1, 0, BT_D1_OPEN), bt_src(a, C_D1, gap_extend1, k, +1, 0, BT_D1_EXT)); break;
      default:   max_all = MAXI(bt_src(a, C_M, gap_open2, k, +1, 0, BT_D2_OPEN), bt_src(a, C_D2, gap_extend2, k, +1, 0, BT_D2_EXT)); break;
    }
    if (max_all < 0) break;
    if (matrix_type == C_M) {
      const int max_offset = (int)(max_all >> 4);
      const int num_matches = offset - max_offset;
      for (int i = 0; i < num_matches; ++i) tmp[pos--] = 'M';
      offset = max_offset;
      v = offset - k; h = offset;
      if (v <= 0 || h <= 0) break;
    }
    const int bt = (int)(max_all & 0xF);
    switch (bt) {
      case BT_M: score = mismatch; matrix_type = C_M; break;
      case BT_I1_OPEN: score = gap_open1; matrix_type = C_M; break;
      case BT_I1_EXT: score = gap_extend1; matrix_type = C_I1; break;
      case BT_I2_OPEN: score = gap_open2; matrix_type = C_M; break;
      case BT_I2_EXT: score = gap_extend2; matrix_type = C_I2; break;
      case BT_D1_OPEN: score = gap_open1; matrix_type = C_M; break;
      case BT_D1_EXT: score = gap_extend1; matrix_type = C_D1; break;
      case BT_D2_OPEN: score = gap_open2; matrix_type = C_M; break;
      case BT_D2_EXT: score = gap_extend2; matrix_type = C_D2; break;
      default: return -1;
    }
    if (bt == BT_M) { tmp[pos--] = 'X'; --offset; }
    else if (bt <= BT_I2_EXT) { tmp[pos--] = 'I'; --k; --offset; }
    else { tmp[pos--] = 'D'; ++k; }
    v = offset - k; h = offset;
  }
  if (matrix_type == C_M) {
    if (v > 0 && h > 0) {
      const int num_matches = MINI(v, h);
      for (int i = 0; i < num_matches; ++i) tmp[pos--] = 'M';
      v -= num_matches; h -= num_matches;
    }
    while (v > 0) { tmp[pos--] = 'D'; --v; }
    while (h > 0) { tmp[pos--] = 'I'; --h; }
  } else {
    if (v != 0 || h != 0 || score != 0) return -1; /* reference exits here (:519-524) */
  }
  *begin_out = pos + 1;
  return 0;
}

/* wavefront_bialign_base (:159-189) = wavefront_unialign (:242-273) + terminate (:147-237) */
static int base_align(wfa_t* b, const char* P, int pb, int pe, const char* T, int tb, int te, int cbegin, int cend,
                      opsbuf_t* out, int* score_out) {
  wfa_init(b, P, pb, pe, T, tb, te, 0, cbegin, cend);
  int score = 0;
  for (;;) {
    if (wfa_extend(b, score, NULL)) break;
    ++score;
    wfa_compute(b, score);
  }
  if (b->status != ST_END_REACHED) return ST_UNATTAINABLE;
  const int cap = 2 * (b->plen + b->tlen) + 8;
  char* tmp = (char*)malloc((size_t)cap);
  int begin = 0;
  const int rc = backtrace_affine(b, score, b->end_k, b->end_off, tmp, cap, &begin);
  if (rc == 0) ops_push(out, tmp + begin, cap - begin);
  free(tmp);
  if (score_out) *score_out = score;
  if (b->cnt) b->cnt->score_steps += score;
  return rc == 0 ? ST_OK : ST_UNATTAINABLE;
}

/* ---- biWFA recursion ------------------------------------------------------------------------ */

typedef struct {
  wfa_t f, r, b;
  const char* P;
  const char* T;
  opsbuf_t out;
  int total_score;
} bictx_t;

/* wavefront_bialign_alignment, wavefront_bialign.c:1144-1221 */
static int bialign_rec(bictx_t* c, int pb, int pe, int tb, int te, int cbegin, int cend, int score_remaining,
                       int level) {
  const int plen = pe - pb, tlen = te - tb;
  if (tlen == 0) { ops_fill(&c->out, 'D', plen); return ST_OK; }
  if (plen == 0) { ops_fill(&c->out, 'I', tlen); return ST_OK; }
  if (score_remaining <= FALLBACK_MIN_SCORE) {
    int sc = 0;
    return base_align(&c->b, c->P, pb, pe, c->T, tb, te, cbegin, cend, &c->out, &sc);
  }
  wfa_init(&c->f, c->P, pb, pe, c->T, tb, te, 0, cbegin, cend);
  wfa_init(&c->r, c->P, pb, pe, c->T, tb, te, 1, cend, cbegin); /* wavefront_bialign_init :140-143 */
  breakpoint_t bp;
  const int st = find_breakpoint(&c->f, &c->r, &bp);
  if (st != ST_OK) {
    /* wavefront_bialign_find_breakpoint_exception, :1083-1110 */
    if (st == ST_END_REACHED) {
      const int score_reached = (c->f.status == ST_END_REACHED) ? c->f.status_score : c->r.status_score;
      if (score_reached <= RECOVERY_MIN_SCORE) {
        int sc = 0;
        return base_align(&c->b, c->P, pb, pe, c->T, tb, te, cbegin, cend, &c->out, &sc);
      }
      return ST_END_UNREACHABLE;
    }
    return st;
  }
  const int bh = bp.offset_forward, bv = bp.offset_forward - bp.k_forward;
  int rc = bialign_rec(c, pb, pb + bv, tb, tb + bh, cbegin, bp.component, bp.score_forward, level + 1);
  if (rc != ST_OK) return rc;
  rc = bialign_rec(c, pb + bv, pe, tb + bh, te, bp.component, cend, bp.score_reverse, level + 1);
  if (rc != ST_OK) return rc;
  if (level == 0) c->total_score = bp.score;
  return ST_OK;
}

static int scope_of(const orc_penalties_t* p) { return MAXI(MAXI(p->o2 + p->e2, p->o1 + p->e1), p->x) + 1; }

int orc_cigar_score(const char* ops, int n, const orc_penalties_t* p) {
  int score = 0, i = 0;
  while (i < n) {
    int j = i;
    while (j < n && ops[j] == ops[i]) ++j;
    const int len = j - i;
    if (ops[i] == 'X') score += p->x * len;
    else if (ops[i] == 'I' || ops[i] == 'D') score += MINI(p->o1 + p->e1 * len, p->o2 + p->e2 * len);
    i = j;
  }
  return score;
}

int orc_cigar_check(const char* pattern, int plen, const char* text, int tlen, const char* ops, int n) {
  int v = 0, h = 0;
  for (int i = 0; i < n; ++i) {
    switch (ops[i]) {
      case 'M': if (v >= plen || h >= tlen || pattern[v] != text[h]) return 0; ++v; ++h; break;
      case 'X': if (v >= plen || h >= tlen || pattern[v] == text[h]) return 0; ++v; ++h; break;
      case 'I': if (h >= tlen) return 0; ++h; break;
      case 'D': if (v >= plen) return 0; ++v; break;
      default: return 0;
    }
  }
  return v == plen && h == tlen;
}

int orc_biwfa_align(const char* pattern, int plen, const char* text, int tlen, const orc_penalties_t* pen,
                    char* ops_out, int ops_cap, int* ops_len, int* score, orc_wfa_counters_t* counters) {
  bictx_t c;
  memset(&c, 0, sizeof(c));
  const int scope = scope_of(pen);
  wfa_alloc(&c.f, 1, scope);
  wfa_alloc(&c.r, 1, scope);
  wfa_alloc(&c.b, 0, 64);
  c.f.pen = c.r.pen = c.b.pen = *pen;
  c.f.scope = c.r.scope = c.b.scope = scope;
  c.f.cnt = c.r.cnt = c.b.cnt = counters;
  c.P = pattern;
  c.T = text;
  c.out.ops = ops_out;
  c.out.cap = ops_cap;
  c.out.len = 0;
  /* wavefront_bialign, :1266-1293 */
  const int min_length = MAXI(plen, tlen) <= FALLBACK_MIN_LENGTH;
  int rc = bialign_rec(&c, 0, plen, 0, tlen, C_M, C_M, min_length ? 0 : INT_MAX, 0);
  if (rc == ST_OK && c.out.len > c.out.cap) rc = -1000;
  if (rc == ST_OK) {
    *ops_len = c.out.len;
    if (score) *score = -orc_cigar_score(ops_out, c.out.len, pen);
  } else {
    *ops_len = 0;
    if (score) *score = 0;
    if (rc > 0) rc = ST_UNATTAINABLE;
  }
  wfa_free(&c.f);
  wfa_free(&c.r);
  wfa_free(&c.b);
  return rc;
}

int orc_wfa_align(const char* pattern, int plen, const char* text, int tlen, const orc_penalties_t* pen,
                  char* ops_out, int ops_cap, int* ops_len, int* score, orc_wfa_counters_t* counters) {
  wfa_t b;
  wfa_alloc(&b, 0, 64);
  b.pen = *pen;
  b.scope = scope_of(pen);
  b.cnt = counters;
  opsbuf_t out = {ops_out, ops_cap, 0};
  int sc = 0;
  int rc = base_align(&b, pattern, 0, plen, text, 0, tlen, C_M, C_M, &out, &sc);
  if (rc == ST_OK && out.len > out.cap) rc = -1000;
  *ops_len = rc == ST_OK ? out.len : 0;
  if (score) *score = -sc;
  wfa_free(&b);
  return rc;
}

/* Ends-free unidirectional WFA (wflign.cpp:280-305 head patch, :368-397 tail patch use it with
 * MemoryMed; `high` and `med` produce identical CIGARs). term_group: see wfa_extend_endsfree. */
int orc_wfa_endsfree(const char* pattern, int plen, int pbf, int pef, const char* text, int tlen, int tbf, int tef,
                     const orc_penalties_t* pen, int term_group, char* ops_out, int ops_cap, int* ops_len, int* score) {
  wfa_t b;
  wfa_alloc(&b, 0, 64);
  b.pen = *pen;
  b.scope = scope_of(pen);
  b.cnt = 0;
  wfa_init_endsfree(&b, pattern, plen, text, tlen, pbf, pef, tbf, tef, term_group);
  int sc = 0;
  for (;;) {
    if (wfa_extend_endsfree(&b, sc)) break;
    ++sc;
    wfa_compute(&b, sc);
  }
  int rc = ST_UNATTAINABLE;
  *ops_len = 0;
  if (b.status == ST_END_REACHED) {
    const int cap = 2 * (plen + tlen) + 8;
    char* tmp = (char*)malloc((size_t)cap);
    int begin = 0;
    if (backtrace_affine(&b, sc, b.end_k, b.end_off, tmp, cap, &begin) == 0 && cap - begin <= ops_cap) {
      memcpy(ops_out, tmp + begin, (size_t)(cap - begin));
      *ops_len = cap - begin;
      rc = ST_OK;
    }
    free(tmp);
  }
  if (score) *score = -sc;
  wfa_free(&b);
  return rc;
}
