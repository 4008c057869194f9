// wfa_kernels.h — gap-affine-2p biWFA on the GPU (sm_100a), batched over mapping records.
//
// What it computes (bit-identical operation strings to the reference):
//   wavefront_bialign            deps/WFA2-lib/wavefront/wavefront_bialign.c:1266-1293
//   wavefront_bialign_alignment  :1144-1221   (recursion -> level-synchronous task queues here)
//   ..._find_breakpoint          :974-1082    (wfb_break_kernel: one CTA per sub-problem)
//   ..._overlap / breakpoint_*   :877-955, 508-571, 828-872
//   wavefront_compute_affine2p   wavefront_compute_affine2p.c:45-106,334-368 + wavefront_compute.c
//   wavefront_extend_end2end     wavefront_extend.c:86-211, wavefront_extend_kernels.c:68-152
//   wavefront_bialign_base       wavefront_bialign.c:159-189 = unialign + backtrace_affine
//                                (wfb_base_kernel: one small CTA per sub-problem)
//
// B200 mapping (not a translation of the CPU code):
//   * the reference's recursive divide & conquer becomes two device-resident task queues
//     (breakpoint tasks, base tasks) drained level by level; children are appended with atomics;
//   * one CTA owns one sub-problem; diagonals are spread over the CTA's threads (coalesced 32-bit
//     offsets), compute + extend + trim are FUSED in one pass with ONE __syncthreads per score;
//   * wavefront metadata ([lo,hi], existence) lives in shared memory as a ring of scope+1 scores,
//     trimmed ends are found with warp redux + shared-memory atomics instead of serial scans;
//   * the forward and reverse aligners read the sequences from a forward and a pre-reversed copy so
//     that extension is always an ascending, word-wise (4 bases / iteration) compare;
//   * operation strings are written straight to their final anti-diagonal slot (v+h) of the pair's
//     output buffer by whichever task produces them, then compacted — no serial concatenation.
#pragma once
#include "wfb_rt.h"

#ifndef WFB_UNROLL
#define WFB_UNROLL 1
#endif
#ifndef WFB_OVL_UNROLL
#define WFB_OVL_UNROLL 4
#endif
#ifndef WFB_PF_NEXT
#define WFB_PF_NEXT 1
#endif
#ifndef WFB_FASTPATH
#define WFB_FASTPATH 1
#endif
#ifndef WFB_DUAL_PHASE1
#define WFB_DUAL_PHASE1 1
#endif
#ifndef WFB_PREFETCH_DIST
#define WFB_PREFETCH_DIST 1
#endif
#ifndef WFB_STEP_INLINE
#ifdef WFB_STEP_NOINLINE
#define WFB_STEP_INLINE WFB_DEV_NOINLINE
#else
#define WFB_STEP_INLINE WFB_DEV
#endif
#endif

#ifndef WFB_EXT_BATCH
#define WFB_EXT_BATCH 1 /* the first 16 bases of the four extensions of a group compared together (packed sequences only) */
#endif
#ifndef WFB_NARROW_SCALAR
#define WFB_NARROW_SCALAR 4 /* profiles/r01_sweep14_narrow_scalar.log: 0 -> 14.65, 1 -> 14.87, 2 -> 15.21, 4 -> 15.28, 8 -> 15.16, 16 -> 14.49 Mbp/s */
#endif
#define WFB_OFFSET_NULL (INT32_MIN / 2) /* wavefront_offset.h:44 */
#define WFB_RMAX 40                     /* max ring slots = max_score_scope + 1 */
#define WFB_FALLBACK_MIN_SCORE 250      /* wavefront_bialign.c:52 */
#define WFB_FALLBACK_MIN_LENGTH 100     /* :53 */
#define WFB_RECOVERY_MIN_SCORE 500      /* :54 */

enum { WFB_M = 0, WFB_I1 = 1, WFB_I2 = 2, WFB_D1 = 3, WFB_D2 = 4 };
enum { WFB_ST_OK = 0, WFB_ST_END_REACHED = 1, WFB_ST_END_UNREACHABLE = 2 };
/* per-pair status codes written by the kernels (0 = fine) */
enum { WFB_PAIR_OK = 0, WFB_PAIR_UNATTAINABLE = -3, WFB_PAIR_BASE_SCORE_CAP = -4, WFB_PAIR_QUEUE_OVERFLOW = -5,
       WFB_PAIR_BACKTRACE = -6 };

struct WfbPen {
  int x, o1, e1, o2, e2;
  int scope; /* max_score_scope = max(x, o1+e1, o2+e2) + 1, wavefront_components.c:101-112 */
  int R;     /* ring slots = scope + 1 */
};

struct WfbTask { /* one sub-problem [pb,pe) x [tb,te) of a pair */
  int pair;
  int pb, pe, tb, te;
  int cbegin, cend;
  int score_remaining;
};

struct WfbPairDesc {
  long long p_off, t_off;       /* forward copies inside the sequence buffer            */
  long long prev_off, trev_off; /* reversed copies                                      */
  long long ops_off;            /* first byte of this pair's (plen+tlen) op slots       */
  int plen, tlen;
};

struct WfbQueue {
  WfbTask* tasks;
  int* count;
  int cap;
};

struct WfbCounters {
  unsigned long long cells, extend_matches, overlap_tests, score_steps, break_tasks, base_tasks;
  unsigned long long base_cells, base_extend_matches, base_score_steps; /* base kernel's share */
};

struct WfbTaskLog { /* optional per-task timeline (WFB_DEBUG_TASKS), one entry per breakpoint task */
  long long t0, t1; /* globaltimer ns */
  int smid, steps, score_f, score_r, plen, tlen, status, pad_;
};

WFB_DEV long long wfb_globaltimer() {
#ifndef WFB_EMU
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#else
  return 0;
#endif
}
WFB_DEV int wfb_smid() {
#ifndef WFB_EMU
  int s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  return s;
#else
  return 0;
#endif
}

struct WfbRing { /* shared-memory wavefront metadata of the last R scores */
  int lo[WFB_RMAX][5];
  int hi[WFB_RMAX][5];
  int boff[WFB_RMAX][5]; /* element offset so that cell(k) = basep[boff + k] */
  int mak[WFB_RMAX][5];  /* max anti-diagonal 2*offset-k over the computed cells (overlap pruning) */
  int cw[WFB_RMAX];      /* diagonals computed by the step (before trimming): the C counter of SURVEY 8(d) */
  unsigned char ex[WFB_RMAX][5];
};

template <bool B> struct WfbBool { static constexpr bool value = B; };

struct WfbIn {
  int off; /* cell(k) = basep[off + k] */
  int lo, hi;
};

struct WfbBreakpoint {
  int score, score_forward, score_reverse;
  int k_forward, k_reverse, offset_forward, offset_reverse;
  int component;
};

/* wavefront_compute_get_*wavefront (wavefront_compute.c:266-305): null => lo=1, hi=-1.
 * slot = score % R is derived by the caller from the current slot (one subtraction + wrap: a runtime modulo costs
 * ~40 dependent instructions, and a score step needs seven of these); valid = (score >= 0). Branch-free. */
WFB_DEV WfbIn wfb_fetch_slot(const WfbRing& r, int comp, bool valid, int slot) {
  const int lo = r.lo[slot][comp], hi = r.hi[slot][comp], off = r.boff[slot][comp];
  const bool ok = valid && r.ex[slot][comp] && lo <= hi;
  WfbIn w;
  w.off = ok ? off : 0;
  w.lo = ok ? lo : 1;
  w.hi = ok ? hi : -1;
  return w;
}
/* (cur - d) mod R for 0 <= d < R, cur in [0, R) */
WFB_DEV int wfb_slot_back(int cur, int d, int R) {
  const int s = cur - d;
  return s < 0 ? s + R : s;
}
WFB_DEV int32_t wfb_get(const int32_t* basep, const WfbIn& w, int k) {
  return (k >= w.lo && k <= w.hi) ? wfb_ld_row1(basep + w.off + k) : WFB_OFFSET_NULL;
}

/* Length of the common prefix of p[0..] and t[0..], capped at limit. Reads up to 23 bytes past
 * the cap (sequence buffers are padded). wavefront_extend_kernels.c:68-92 does the same 8 bytes at a
 * time on the CPU. */
WFB_DEV int wfb_match_run_words(const uint8_t* p, const uint8_t* t, int limit) {
  if (limit <= 0) return 0;
  const uintptr_t pa = (uintptr_t)p, ta = (uintptr_t)t;
  const uint32_t* pw = (const uint32_t*)(pa & ~(uintptr_t)3);
  const uint32_t* tw = (const uint32_t*)(ta & ~(uintptr_t)3);
  const unsigned ps = (unsigned)(pa & 3) * 8u, ts = (unsigned)(ta & 3) * 8u;
  uint32_t p0 = wfb_ldg32(pw), t0 = wfb_ldg32(tw);
  int n = 0;
  while (n < limit) {
    /* four words of each sequence per trip (16 bases), the eight loads are independent: a long exact run costs one
     * memory round trip per 16 bases */
    const uint32_t p1 = wfb_ldg32(pw + 1), t1 = wfb_ldg32(tw + 1);
    const uint32_t p2 = wfb_ldg32(pw + 2), t2 = wfb_ldg32(tw + 2);
    const uint32_t p3 = wfb_ldg32(pw + 3), t3 = wfb_ldg32(tw + 3);
    const uint32_t p4 = wfb_ldg32(pw + 4), t4 = wfb_ldg32(tw + 4);
    const uint32_t x1 = __funnelshift_r(p0, p1, ps) ^ __funnelshift_r(t0, t1, ts);
    if (x1) { n += (__ffs((int)x1) - 1) >> 3; break; }
    const uint32_t x2 = __funnelshift_r(p1, p2, ps) ^ __funnelshift_r(t1, t2, ts);
    if (x2) { n += 4 + ((__ffs((int)x2) - 1) >> 3); break; }
    const uint32_t x3 = __funnelshift_r(p2, p3, ps) ^ __funnelshift_r(t2, t3, ts);
    if (x3) { n += 8 + ((__ffs((int)x3) - 1) >> 3); break; }
    const uint32_t x4 = __funnelshift_r(p3, p4, ps) ^ __funnelshift_r(t3, t4, ts);
    if (x4) { n += 12 + ((__ffs((int)x4) - 1) >> 3); break; }
    n += 16;
    p0 = p4;
    t0 = t4;
    pw += 4;
    tw += 4;
  }
  return n < limit ? n : limit;
}

WFB_DEV int wfb_match_run(const uint8_t* p, const uint8_t* t, int limit) {
  if (limit <= 0) return 0;
  if (wfb_ldg8(p) != wfb_ldg8(t)) return 0; /* 3 of 4 cells of an unrelated diagonal stop right here */
  return wfb_match_run_words(p, t, limit);
}

/* ---- 2-bit packed sequence windows in shared memory -----------------------------------------------------------
 * A breakpoint task keeps its four sequence slices (pattern / text, forward / reversed) packed at 2 bits per base in
 * shared memory: the extension of a cell — the latency chain of every score step (first byte from L2 / HBM, then 16 bases per
 * round trip) — becomes shared-memory reads of 16 bases per word pair. Only equality matters, so any injective code works:
 * (c >> 1) & 3 maps A, C, T, G to 0, 1, 2, 3. Pairs holding any other byte (N after the reference's masking) keep the byte-wise
 * global path (wfb_reverse_kernel flags them). S[w] holds bases 16w .. 16w+15 of the 16-byte-aligned chunk the slice starts in. */
WFB_DEV uint32_t wfb_pk16(const uint32_t* S, int i) {
  const int w = i >> 4;
  return __funnelshift_r(S[w], S[w + 1], (unsigned)(i & 15) * 2u);
}
WFB_DEV int wfb_match_run_packed(const uint32_t* P, int pi, const uint32_t* T, int ti, int limit) {
  int n = 0;
  while (n < limit) {
    const uint32_t x = wfb_pk16(P, pi + n) ^ wfb_pk16(T, ti + n);
    if (x) { n += (__ffs((int)x) - 1) >> 1; break; }
    n += 16;
  }
  return n < limit ? n : limit;
}
/* pack nwords x 16 bases starting at the 16-byte-aligned address g (all threads; the caller syncs) */
WFB_DEV void wfb_pack_seq(const uint8_t* g, uint32_t* S, int nwords) {
#ifndef WFB_EMU
  for (int j = WFB_TID; j < nwords; j += WFB_NT) {
    const uint4 v = __ldg((const uint4*)(g + 16 * (size_t)j));
    uint32_t y0 = (v.x >> 1) & 0x03030303u, y1 = (v.y >> 1) & 0x03030303u, y2 = (v.z >> 1) & 0x03030303u, y3 = (v.w >> 1) & 0x03030303u;
    y0 = (y0 | (y0 >> 6) | (y0 >> 12) | (y0 >> 18)) & 0xffu;
    y1 = (y1 | (y1 >> 6) | (y1 >> 12) | (y1 >> 18)) & 0xffu;
    y2 = (y2 | (y2 >> 6) | (y2 >> 12) | (y2 >> 18)) & 0xffu;
    y3 = (y3 | (y3 >> 6) | (y3 >> 12) | (y3 >> 18)) & 0xffu;
    S[j] = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
  }
#else
  (void)g; (void)S; (void)nwords;
#endif
}

WFB_DEV void wfb_prefetch_row(const void* p) {
#ifndef WFB_EMU
#if (WFB_PF_NEXT & 3) == 2
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
#else
  (void)p;
#endif
}
WFB_DEV void wfb_prefetch_l2(const void* p) {
#ifndef WFB_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

/* The 128-bit row loads of a 4-diagonal group fetch one lane no cell reads (I1 / I2: the diagonal k0 + 3, D1 / D2: k0). ptxas hands the
 * registers of those dead lanes to the scalar halo loads that follow in the same load block, and a load whose destination register is still
 * owed by an earlier load has to wait for it (write-after-write on the scoreboard): the thirteen loads of a group, meant to be in flight
 * together, became two memory round trips — one scalar LDG held 9 % of the kernel's stall samples on scerevisiae8
 * (profiles/r02_persist_c3s8_summary.txt). A real use of the four dead lanes after the block keeps their registers allocated until
 * every load has been issued. Offsets and nulls only have bits 0-19 and 30-31 set, so the test is never true; if it were, TAK (an upper
 * bound used to prune the overlap scan) would only grow, which is safe. */
#ifndef WFB_EMU
#define WFB_KEEP_DEAD_LANES(VI1, VI2, VD1, VD2, TAK)                                  \
  {                                                                                   \
    const int dead__ = ((VI1).w ^ (VI2).w) | ((VD1).x ^ (VD2).x);                     \
    if (dead__ == 0x5a5a5a5b) (TAK) = max((TAK), dead__);                             \
  }
#else
#define WFB_KEEP_DEAD_LANES(VI1, VI2, VD1, VD2, TAK)
#endif

WFB_DEV bool wfb_inbounds(int32_t off, int k, int plen, int tlen) {
  return (uint32_t)off <= (uint32_t)tlen && (uint32_t)(off - k) <= (uint32_t)plen;
}

/* Optional phase timers (-DWFB_PHASE_TIMERS, tuning builds only): per wavefront-width bucket b in {<=128, <=1024,
 * <=4096, >4096}: [4b+0] steps, [4b+1] SM cycles entry -> end of the step's work, [4b+2] cycles of thread 0's own cells,
 * [4b+3] diagonals. [16] = cycles in overlap scans, [17] = overlap calls. [20],[21] = cycles / count of base-case tasks,
 * [22],[23] = of their single-thread backtraces, [24],[25] = of breakpoint tasks, [26] = cycles CTAs waited for a task.
 * Thread 0 sits at the lo edge of the wavefront. */
#if defined(WFB_PHASE_TIMERS) && !defined(WFB_EMU)
__device__ unsigned long long g_wfb_phase[32];
#define WFB_PT_CLOCK() ((long long)clock64())
#define WFB_PT_ADD(i, v) atomicAdd(&g_wfb_phase[(i)], (unsigned long long)(v))
#else
#define WFB_PT_CLOCK() 0ll
#define WFB_PT_ADD(i, v) ((void)0)
#endif

/* Per-thread accumulators that live in registers for the whole task. */
struct WfbAcc {
  unsigned long long cells;   /* uniform, counted by thread 0 */
  unsigned long long overlap; /* uniform, counted by thread 0 */
  unsigned long long matches; /* per thread */
  unsigned long long steps;   /* uniform */
};

/*
 * One score step of one direction: compute the five component wavefronts of `score` from the ring,
 * extend M along matches, trim the ends, detect end-of-alignment. Exactly one barrier.
 * Returns WFB_ST_OK / WFB_ST_END_REACHED / WFB_ST_END_UNREACHABLE (uniform over the CTA).
 *   alloc(slot, lo, hi, ob[5]) assigns storage for the outputs (uniform).
 *   red_maxak[3] : shared-memory reduction slots, rotated by score % 3.
 */
template <class Alloc>
WFB_STEP_INLINE void wfb_step_work(WfbRing& ring, int32_t* basep, const WfbPen& pen, int score, const uint8_t* pseq,
                     const uint8_t* tseq, int plen, int tlen, int cend, Alloc& alloc, int* red_maxak,
                     int* red_end, WfbAcc& acc, const int tid, const int nt, const int slot /* score % R */,
                     const int par /* score % 3 */) {
  const int R = pen.R, nslot = slot + 1 == R ? 0 : slot + 1;
  const int npar = par == 2 ? 0 : par + 1;
  (void)basep;
  /* every distance below is <= max_score_scope = R - 1 */
  const int d_x = pen.x, d_o1 = pen.o1 + pen.e1, d_o2 = pen.o2 + pen.e2;
  const int s_e1 = wfb_slot_back(slot, pen.e1, R), s_e2 = wfb_slot_back(slot, pen.e2, R);
  const WfbIn m_misms = wfb_fetch_slot(ring, WFB_M, score >= d_x, wfb_slot_back(slot, d_x, R));
  const WfbIn m_open1 = wfb_fetch_slot(ring, WFB_M, score >= d_o1, wfb_slot_back(slot, d_o1, R));
  const WfbIn m_open2 = wfb_fetch_slot(ring, WFB_M, score >= d_o2, wfb_slot_back(slot, d_o2, R));
  const WfbIn i1_ext = wfb_fetch_slot(ring, WFB_I1, score >= pen.e1, s_e1);
  const WfbIn i2_ext = wfb_fetch_slot(ring, WFB_I2, score >= pen.e2, s_e2);
  const WfbIn d1_ext = wfb_fetch_slot(ring, WFB_D1, score >= pen.e1, s_e1);
  const WfbIn d2_ext = wfb_fetch_slot(ring, WFB_D2, score >= pen.e2, s_e2);
  const bool n_m = m_misms.lo > m_misms.hi, n_o1 = m_open1.lo > m_open1.hi, n_o2 = m_open2.lo > m_open2.hi;
  const bool n_i1 = i1_ext.lo > i1_ext.hi, n_i2 = i2_ext.lo > i2_ext.hi;
  const bool n_d1 = d1_ext.lo > d1_ext.hi, n_d2 = d2_ext.lo > d2_ext.hi;
#if defined(WFB_PHASE_TIMERS) && !defined(WFB_EMU)
  const long long pt0 = WFB_PT_CLOCK();
#endif
  if (n_m && n_o1 && n_o2 && n_i1 && n_i2 && n_d1 && n_d2) {
    /* wavefront_compute_affine2p.c:341-351 + wavefront_extend.c:95-103 */
    if (tid == 0) {
      for (int c = 0; c < 5; ++c) {
        ring.ex[slot][c] = 0;
        ring.lo[nslot][c] = INT_MAX;
        ring.hi[nslot][c] = INT_MIN;
        ring.mak[nslot][c] = INT_MIN;
      }
      ring.cw[slot] = 0;
      red_maxak[npar] = 0;
    }
    return;
  }
  /* wavefront_compute_limits_input, wavefront_compute.c:40-86 */
  int lo = m_misms.lo, hi = m_misms.hi;
  lo = min(lo, m_open1.lo - 1); hi = max(hi, m_open1.hi + 1);
  lo = min(lo, i1_ext.lo + 1);  hi = max(hi, i1_ext.hi + 1);
  lo = min(lo, d1_ext.lo - 1);  hi = max(hi, d1_ext.hi - 1);
  lo = min(lo, m_open2.lo - 1); hi = max(hi, m_open2.hi + 1);
  lo = min(lo, i2_ext.lo + 1);  hi = max(hi, i2_ext.hi + 1);
  lo = min(lo, d2_ext.lo - 1);  hi = max(hi, d2_ext.hi - 1);
  /* wavefront_compute_allocate_output, wavefront_compute.c:447-493 */
  const bool ex_i1 = !n_o1 || !n_i1, ex_d1 = !n_o1 || !n_d1;
  const bool ex_i2 = !n_o2 || !n_i2, ex_d2 = !n_o2 || !n_d2;
  int ob[5];
  alloc(slot, lo, hi, ob);
  if (tid == 0) {
    ring.ex[slot][WFB_M] = 1;
    ring.ex[slot][WFB_I1] = ex_i1;
    ring.ex[slot][WFB_I2] = ex_i2;
    ring.ex[slot][WFB_D1] = ex_d1;
    ring.ex[slot][WFB_D2] = ex_d2;
    for (int c = 0; c < 5; ++c) {
      ring.boff[slot][c] = ob[c];
      ring.lo[nslot][c] = INT_MAX;
      ring.hi[nslot][c] = INT_MIN;
      ring.mak[nslot][c] = INT_MIN;
    }
    red_maxak[npar] = 0;
    ring.cw[slot] = hi - lo + 1;
  }
  /* per-thread trim / anti-diagonal accumulators */
  int tlo[5], thi[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) { tlo[c] = INT_MAX; thi[c] = INT_MIN; }
  int tmax = 0;        /* max anti-diagonal of extended in-bounds M cells (wavefront_extend_*_max)        */
  int tak = INT_MIN;   /* max anti-diagonal over EVERY computed value of every component (overlap pruning;
                          raw values incl. out-of-bounds ones: a superset of what the overlap scan sees) */
  const int ak_end = tlen - plen; /* the diagonal wavefront_termination_end2end looks at */

  /* One wavefront cell (wavefront_compute_affine2p_idm, wavefront_compute_affine2p.c:71-105) fused with
   * its extension (wavefront_extend_kernels.c:125-152) and trim bookkeeping (wavefront_compute.c:579-613). */
#define WFB_CELL_A(K, O1M, O1P, O2M, O2P, I1V, I2V, D1V, D2V, MMV, OUT_M, OUT_I1, OUT_I2, OUT_D1, OUT_D2)    \
  {                                                                                                       \
    const int k_ = (K);                                                                                   \
    const int32_t ins1 = max((O1M), (I1V)) + 1;                                                           \
    const int32_t ins2 = max((O2M), (I2V)) + 1;                                                           \
    const int32_t del1 = max((O1P), (D1V));                                                               \
    const int32_t del2 = max((O2P), (D2V));                                                               \
    int32_t mx = max(max(del1, del2), max((MMV) + 1, max(ins1, ins2)));                                   \
    if (mx >= 0) tak = max(tak, 2 * mx - k_); /* negative (null-ish) values can never be part of a hit */ \
    if (!wfb_inbounds(mx, k_, plen, tlen)) mx = WFB_OFFSET_NULL;                                          \
    /* an I value keeps the v of its in-bounds source and a D value keeps the h, null-ish values are hugely      \
     * negative: ONE unsigned compare decides what wavefront_compute_trim_ends (:594-596) tests with two */      \
    if (ex_i1 && (uint32_t)ins1 <= (uint32_t)tlen) { tlo[WFB_I1] = min(tlo[WFB_I1], k_); thi[WFB_I1] = max(thi[WFB_I1], k_); } \
    if (ex_i2 && (uint32_t)ins2 <= (uint32_t)tlen) { tlo[WFB_I2] = min(tlo[WFB_I2], k_); thi[WFB_I2] = max(thi[WFB_I2], k_); } \
    if (ex_d1 && (uint32_t)(del1 - k_) <= (uint32_t)plen) { tlo[WFB_D1] = min(tlo[WFB_D1], k_); thi[WFB_D1] = max(thi[WFB_D1], k_); } \
    if (ex_d2 && (uint32_t)(del2 - k_) <= (uint32_t)plen) { tlo[WFB_D2] = min(tlo[WFB_D2], k_); thi[WFB_D2] = max(thi[WFB_D2], k_); } \
    (OUT_M) = mx; (OUT_I1) = ins1; (OUT_I2) = ins2; (OUT_D1) = del1; (OUT_D2) = del2;                       \
  }
  /* bookkeeping of one in-bounds M cell after its extension by RUN matches (wavefront_extend_kernels.c:125-152) */
#define WFB_CELL_X(K, OUT_M, RUN)                                                                         \
  {                                                                                                       \
    const int k_ = (K);                                                                                   \
    const int wfb_x_run_ = (RUN);                                                                         \
    (OUT_M) += wfb_x_run_;                                                                                \
    if (alloc.runflag) alloc.runflag[k_ + alloc.runbias] = wfb_x_run_ >= 4 ? 1 : 0;                       \
    acc.matches += (unsigned)wfb_x_run_;                                                                  \
    tmax = max(tmax, 2 * (OUT_M) - k_);                                                                   \
    tlo[WFB_M] = min(tlo[WFB_M], k_);                                                                     \
    thi[WFB_M] = max(thi[WFB_M], k_);                                                                     \
  }
  /* the fused form used on the scalar paths */
#define WFB_CELL(K, O1M, O1P, O2M, O2P, I1V, I2V, D1V, D2V, MMV, OUT_M, OUT_I1, OUT_I2, OUT_D1, OUT_D2)      \
  {                                                                                                       \
    WFB_CELL_A(K, O1M, O1P, O2M, O2P, I1V, I2V, D1V, D2V, MMV, OUT_M, OUT_I1, OUT_I2, OUT_D1, OUT_D2)       \
    if ((OUT_M) >= 0) {                                                                                   \
      const int v__ = (OUT_M) - (K);                                                                      \
      const int lim__ = min(plen - v__, tlen - (OUT_M));                                                  \
      const int run__ = alloc.sp ? wfb_match_run_packed(alloc.sp, alloc.spo + v__, alloc.st, alloc.sto + (OUT_M), lim__) \
                                 : wfb_match_run(pseq + v__, tseq + (OUT_M), lim__);                       \
      WFB_CELL_X(K, OUT_M, run__)                                                                         \
    }                                                                                                     \
  }

  /* The four cells of a group with their extensions' first 16 bases compared together: a cell's extension is a chain of dependent
   * shared-memory reads and bit tricks, 3 of 4 cells stop inside their first word, and four chains issued back to back fill the issue
   * slots that one chain leaves empty (the step is bound by dependent-issue latency with 4 warps per scheduler, not by bandwidth).
   * Only with the packed sequences in shared memory (alloc.sp); a cell whose first 16 bases all match continues in the word loop.
   * Same results as four WFB_CELLs: wfb_match_run_packed's first trip is exactly this. */
#define WFB_EXT_FIRST(U, KK)                                                                              \
    const bool ok##U = rm[U] >= 0;                                                                        \
    const int v##U = ok##U ? rm[U] - (KK) : 0, h##U = ok##U ? rm[U] : 0;                                  \
    const int lim##U = min(plen - v##U, tlen - h##U);                                                     \
    const uint32_t x##U = wfb_pk16(alloc.sp, alloc.spo + v##U) ^ wfb_pk16(alloc.st, alloc.sto + h##U);
#define WFB_EXT_REST(U, KK)                                                                               \
    if (ok##U) {                                                                                          \
      int run__ = x##U ? (__ffs((int)x##U) - 1) >> 1 : 16;                                                \
      if (x##U == 0 && lim##U > 16)                                                                       \
        run__ += wfb_match_run_packed(alloc.sp, alloc.spo + v##U + 16, alloc.st, alloc.sto + h##U + 16, lim##U - 16); \
      run__ = min(run__, lim##U);                                                                         \
      WFB_CELL_X(KK, rm[U], run__)                                                                        \
    }
  /* hand the end component's offset on the final diagonal to the termination test (wavefront_termination.c:37-114) */
#define WFB_END_HANDOFF(K, VM, VI1, VI2, VD1, VD2)                                                          \
  if ((K) == ak_end && cend >= 0)                                                                          \
    red_end[par] = cend == WFB_M ? (VM) : cend == WFB_I1 ? (VI1) : cend == WFB_I2 ? (VI2) : cend == WFB_D1 ? (VD1) : (VD2);

  const int kalign = alloc.kalign;
  /* A step lasts as long as its busiest thread's chain of dependent cells (each with its own extension). The 4-diagonal groups
   * give ceil(width / 4) threads four cells each and leave the others idle; below 4 diagonals per thread the one-diagonal path
   * spreads the same cells over ALL threads (ceil(width / nt) cells each), which is the shorter critical path until the 128-bit
   * row loads of the group path win (profiles/r01_phase_timers_v6.log: steps of width <= 1024 are 28 % of the CTA cycles, a
   * step of width <= 128 costs 9.5 k cycles of which 8 k are one thread's group; A/B in profiles/r01_sweep14_narrow_scalar.log). */
  const bool vec_ok = kalign >= 0 && !(WFB_NARROW_SCALAR > 0 && (hi - lo + 1) <= WFB_NARROW_SCALAR * nt);
  if (vec_ok) {
    /* groups of 4 diagonals whose cells are 16-byte aligned in every row: 128-bit loads / stores, range
     * checks once per group; ragged groups at the ends of any input take the scalar path */
    int safe_lo = lo, safe_hi = hi; /* diagonals k with [k-1, k+4] inside every non-null input */
    if (!n_m)  { safe_lo = max(safe_lo, m_misms.lo + 1); safe_hi = min(safe_hi, m_misms.hi - 4); }
    if (!n_o1) { safe_lo = max(safe_lo, m_open1.lo + 1); safe_hi = min(safe_hi, m_open1.hi - 4); }
    if (!n_o2) { safe_lo = max(safe_lo, m_open2.lo + 1); safe_hi = min(safe_hi, m_open2.hi - 4); }
    if (!n_i1) { safe_lo = max(safe_lo, i1_ext.lo + 1);  safe_hi = min(safe_hi, i1_ext.hi - 4); }
    if (!n_i2) { safe_lo = max(safe_lo, i2_ext.lo + 1);  safe_hi = min(safe_hi, i2_ext.hi - 4); }
    if (!n_d1) { safe_lo = max(safe_lo, d1_ext.lo + 1);  safe_hi = min(safe_hi, d1_ext.hi - 4); }
    if (!n_d2) { safe_lo = max(safe_lo, d2_ext.lo + 1);  safe_hi = min(safe_hi, d2_ext.hi - 4); }
    safe_hi = min(safe_hi, hi - 3);
    const int kfirst = lo - ((lo + kalign) & 3); /* first group start (<= lo) */
#if WFB_PF_NEXT & 4
    if (Alloc::kFixedRows) {
      /* the first pass of the NEXT score step reads rows that already exist (all but the ones this step writes):
       * bring this thread's part of them towards the SM now */
      const int kp = kfirst + 4 * tid;
      if (kp <= hi) {
        if (d_x > 1)  wfb_prefetch_row(basep + alloc.row(wfb_slot_back(nslot, d_x, R), WFB_M) + kp);
        if (d_o1 > 1) wfb_prefetch_row(basep + alloc.row(wfb_slot_back(nslot, d_o1, R), WFB_M) + kp);
        if (d_o2 > 1) wfb_prefetch_row(basep + alloc.row(wfb_slot_back(nslot, d_o2, R), WFB_M) + kp);
        if (pen.e1 > 1) {
          wfb_prefetch_row(basep + alloc.row(wfb_slot_back(nslot, pen.e1, R), WFB_I1) + kp);
          wfb_prefetch_row(basep + alloc.row(wfb_slot_back(nslot, pen.e1, R), WFB_D1) + kp);
        }
      }
    }
#endif
    /* steady state: all seven inputs exist, so no load is predicated and no null needs materialising; the general
     * body is kept for the first scores of a task. ALLIN is a compile-time constant of each instantiation. */
    const bool o_n_m = n_m, o_n_o1 = n_o1, o_n_o2 = n_o2, o_n_i1 = n_i1, o_n_i2 = n_i2, o_n_d1 = n_d1, o_n_d2 = n_d2;
    const bool o_ex_i1 = ex_i1, o_ex_i2 = ex_i2, o_ex_d1 = ex_d1, o_ex_d2 = ex_d2;
    auto group_loop = [&](auto allin_tag) {
      constexpr bool ALLIN = decltype(allin_tag)::value;
      const bool n_m = ALLIN ? false : o_n_m, n_o1 = ALLIN ? false : o_n_o1, n_o2 = ALLIN ? false : o_n_o2;
      const bool n_i1 = ALLIN ? false : o_n_i1, n_i2 = ALLIN ? false : o_n_i2, n_d1 = ALLIN ? false : o_n_d1, n_d2 = ALLIN ? false : o_n_d2;
      const bool ex_i1 = ALLIN ? true : o_ex_i1, ex_i2 = ALLIN ? true : o_ex_i2, ex_d1 = ALLIN ? true : o_ex_d1, ex_d2 = ALLIN ? true : o_ex_d2;
    for (int k0 = kfirst + 4 * tid; k0 <= hi; k0 += 4 * nt) {
        int32_t rm[4], ri1[4], ri2[4], rd1[4], rd2[4];
#if WFB_PF_NEXT & 3
        { /* the rows this thread reads in its NEXT pass come from L2 / HBM: start them now, no registers needed */
          const int kn = k0 + 4 * nt;
          if (kn >= safe_lo && kn <= safe_hi) {
            if (!n_o1) wfb_prefetch_row(basep + m_open1.off + kn);
            if (!n_o2) wfb_prefetch_row(basep + m_open2.off + kn);
            if (!n_i1) wfb_prefetch_row(basep + i1_ext.off + kn);
            if (!n_i2) wfb_prefetch_row(basep + i2_ext.off + kn);
            if (!n_d1) wfb_prefetch_row(basep + d1_ext.off + kn);
            if (!n_d2) wfb_prefetch_row(basep + d2_ext.off + kn);
            if (!n_m)  wfb_prefetch_row(basep + m_misms.off + kn);
          }
        }
#endif
        const bool safe = k0 >= safe_lo && k0 <= safe_hi;
        const int4 NUL4 = make_int4(WFB_OFFSET_NULL, WFB_OFFSET_NULL, WFB_OFFSET_NULL, WFB_OFFSET_NULL);
        int4 vo1 = NUL4, vo2 = NUL4, vi1 = NUL4, vi2 = NUL4, vd1 = NUL4, vd2 = NUL4, vmm = NUL4;
        int32_t so1m = WFB_OFFSET_NULL, so1p = WFB_OFFSET_NULL, so2m = WFB_OFFSET_NULL, so2p = WFB_OFFSET_NULL;
        int32_t si1 = WFB_OFFSET_NULL, si2 = WFB_OFFSET_NULL, sd1 = WFB_OFFSET_NULL, sd2 = WFB_OFFSET_NULL;
        if (safe) {
          if (!n_o1) { const int32_t* p = basep + m_open1.off + k0; vo1 = wfb_ld_row4(p); so1m = wfb_ld_row1(p - 1); so1p = wfb_ld_row1(p + 4); }
          if (!n_o2) { const int32_t* p = basep + m_open2.off + k0; vo2 = wfb_ld_row4(p); so2m = wfb_ld_row1(p - 1); so2p = wfb_ld_row1(p + 4); }
          if (!n_i1) { const int32_t* p = basep + i1_ext.off + k0; vi1 = wfb_ld_row4(p); si1 = wfb_ld_row1(p - 1); }
          if (!n_i2) { const int32_t* p = basep + i2_ext.off + k0; vi2 = wfb_ld_row4(p); si2 = wfb_ld_row1(p - 1); }
          if (!n_d1) { const int32_t* p = basep + d1_ext.off + k0; vd1 = wfb_ld_row4(p); sd1 = wfb_ld_row1(p + 4); }
          if (!n_d2) { const int32_t* p = basep + d2_ext.off + k0; vd2 = wfb_ld_row4(p); sd2 = wfb_ld_row1(p + 4); }
          if (!n_m)  { vmm = wfb_ld_row4(basep + m_misms.off + k0); }
        } else {
          /* ragged group at an end of some input: the same 36 values by range-checked element loads, all independent
           * (one memory round trip for the group instead of one per cell). Diagonals outside [lo,hi] see only nulls
           * and produce nothing (no valid offset, no trim update); their stores are skipped below. */
          vo1 = make_int4(wfb_get(basep, m_open1, k0), wfb_get(basep, m_open1, k0 + 1), wfb_get(basep, m_open1, k0 + 2), wfb_get(basep, m_open1, k0 + 3));
          so1m = wfb_get(basep, m_open1, k0 - 1); so1p = wfb_get(basep, m_open1, k0 + 4);
          vo2 = make_int4(wfb_get(basep, m_open2, k0), wfb_get(basep, m_open2, k0 + 1), wfb_get(basep, m_open2, k0 + 2), wfb_get(basep, m_open2, k0 + 3));
          so2m = wfb_get(basep, m_open2, k0 - 1); so2p = wfb_get(basep, m_open2, k0 + 4);
          vi1 = make_int4(wfb_get(basep, i1_ext, k0), wfb_get(basep, i1_ext, k0 + 1), wfb_get(basep, i1_ext, k0 + 2), WFB_OFFSET_NULL);
          si1 = wfb_get(basep, i1_ext, k0 - 1);
          vi2 = make_int4(wfb_get(basep, i2_ext, k0), wfb_get(basep, i2_ext, k0 + 1), wfb_get(basep, i2_ext, k0 + 2), WFB_OFFSET_NULL);
          si2 = wfb_get(basep, i2_ext, k0 - 1);
          vd1 = make_int4(WFB_OFFSET_NULL, wfb_get(basep, d1_ext, k0 + 1), wfb_get(basep, d1_ext, k0 + 2), wfb_get(basep, d1_ext, k0 + 3));
          sd1 = wfb_get(basep, d1_ext, k0 + 4);
          vd2 = make_int4(WFB_OFFSET_NULL, wfb_get(basep, d2_ext, k0 + 1), wfb_get(basep, d2_ext, k0 + 2), wfb_get(basep, d2_ext, k0 + 3));
          sd2 = wfb_get(basep, d2_ext, k0 + 4);
          vmm = make_int4(wfb_get(basep, m_misms, k0), wfb_get(basep, m_misms, k0 + 1), wfb_get(basep, m_misms, k0 + 2), wfb_get(basep, m_misms, k0 + 3));
        }
        WFB_KEEP_DEAD_LANES(vi1, vi2, vd1, vd2, tak)
#if WFB_EXT_BATCH && !defined(WFB_EMU)
        if (alloc.sp) { /* uniform per task */
          WFB_CELL_A(k0 + 0, so1m,  vo1.y, so2m,  vo2.y, si1,   si2,   vd1.y, vd2.y, vmm.x, rm[0], ri1[0], ri2[0], rd1[0], rd2[0])
          WFB_CELL_A(k0 + 1, vo1.x, vo1.z, vo2.x, vo2.z, vi1.x, vi2.x, vd1.z, vd2.z, vmm.y, rm[1], ri1[1], ri2[1], rd1[1], rd2[1])
          WFB_CELL_A(k0 + 2, vo1.y, vo1.w, vo2.y, vo2.w, vi1.y, vi2.y, vd1.w, vd2.w, vmm.z, rm[2], ri1[2], ri2[2], rd1[2], rd2[2])
          WFB_CELL_A(k0 + 3, vo1.z, so1p,  vo2.z, so2p,  vi1.z, vi2.z, sd1,   sd2,   vmm.w, rm[3], ri1[3], ri2[3], rd1[3], rd2[3])
          WFB_EXT_FIRST(0, k0 + 0) WFB_EXT_FIRST(1, k0 + 1) WFB_EXT_FIRST(2, k0 + 2) WFB_EXT_FIRST(3, k0 + 3)
          WFB_EXT_REST(0, k0 + 0) WFB_EXT_REST(1, k0 + 1) WFB_EXT_REST(2, k0 + 2) WFB_EXT_REST(3, k0 + 3)
        } else
#endif
        {
        WFB_CELL(k0 + 0, so1m,  vo1.y, so2m,  vo2.y, si1,   si2,   vd1.y, vd2.y, vmm.x, rm[0], ri1[0], ri2[0], rd1[0], rd2[0])
        WFB_CELL(k0 + 1, vo1.x, vo1.z, vo2.x, vo2.z, vi1.x, vi2.x, vd1.z, vd2.z, vmm.y, rm[1], ri1[1], ri2[1], rd1[1], rd2[1])
        WFB_CELL(k0 + 2, vo1.y, vo1.w, vo2.y, vo2.w, vi1.y, vi2.y, vd1.w, vd2.w, vmm.z, rm[2], ri1[2], ri2[2], rd1[2], rd2[2])
        WFB_CELL(k0 + 3, vo1.z, so1p,  vo2.z, so2p,  vi1.z, vi2.z, sd1,   sd2,   vmm.w, rm[3], ri1[3], ri2[3], rd1[3], rd2[3])
        }
        if ((unsigned)(ak_end - k0) < 4u) { /* constant indices keep the arrays in registers */
          WFB_END_HANDOFF(k0 + 0, rm[0], ri1[0], ri2[0], rd1[0], rd2[0])
          WFB_END_HANDOFF(k0 + 1, rm[1], ri1[1], ri2[1], rd1[1], rd2[1])
          WFB_END_HANDOFF(k0 + 2, rm[2], ri1[2], ri2[2], rd1[2], rd2[2])
          WFB_END_HANDOFF(k0 + 3, rm[3], ri1[3], ri2[3], rd1[3], rd2[3])
        }
        if (k0 >= lo && k0 + 3 <= hi) {
          *(int4*)(basep + ob[WFB_M] + k0) = make_int4(rm[0], rm[1], rm[2], rm[3]);
          if (ex_i1) *(int4*)(basep + ob[WFB_I1] + k0) = make_int4(ri1[0], ri1[1], ri1[2], ri1[3]);
          if (ex_i2) *(int4*)(basep + ob[WFB_I2] + k0) = make_int4(ri2[0], ri2[1], ri2[2], ri2[3]);
          if (ex_d1) *(int4*)(basep + ob[WFB_D1] + k0) = make_int4(rd1[0], rd1[1], rd1[2], rd1[3]);
          if (ex_d2) *(int4*)(basep + ob[WFB_D2] + k0) = make_int4(rd2[0], rd2[1], rd2[2], rd2[3]);
        } else {
#define WFB_STORE1(U)                                                                  \
          if (k0 + (U) >= lo && k0 + (U) <= hi) {                                        \
            basep[ob[WFB_M] + k0 + (U)] = rm[U];                                         \
            if (ex_i1) basep[ob[WFB_I1] + k0 + (U)] = ri1[U];                            \
            if (ex_i2) basep[ob[WFB_I2] + k0 + (U)] = ri2[U];                            \
            if (ex_d1) basep[ob[WFB_D1] + k0 + (U)] = rd1[U];                            \
            if (ex_d2) basep[ob[WFB_D2] + k0 + (U)] = rd2[U];                            \
          }
          WFB_STORE1(0) WFB_STORE1(1) WFB_STORE1(2) WFB_STORE1(3)
#undef WFB_STORE1
        }
      }
    };
#if WFB_FASTPATH
    if (!(n_m || n_o1 || n_o2 || n_i1 || n_i2 || n_d1 || n_d2)) group_loop(WfbBool<true>{});
    else
#endif
      group_loop(WfbBool<false>{});
  } else {
    for (int k = lo + tid; k <= hi; k += nt) {
      int32_t rm, ri1, ri2, rd1, rd2;
      WFB_CELL(k, wfb_get(basep, m_open1, k - 1), wfb_get(basep, m_open1, k + 1), wfb_get(basep, m_open2, k - 1),
               wfb_get(basep, m_open2, k + 1), wfb_get(basep, i1_ext, k - 1), wfb_get(basep, i2_ext, k - 1),
               wfb_get(basep, d1_ext, k + 1), wfb_get(basep, d2_ext, k + 1), wfb_get(basep, m_misms, k), rm, ri1, ri2, rd1, rd2)
      WFB_END_HANDOFF(k, rm, ri1, ri2, rd1, rd2)
      basep[ob[WFB_M] + k] = rm;
      if (ex_i1) basep[ob[WFB_I1] + k] = ri1;
      if (ex_i2) basep[ob[WFB_I2] + k] = ri2;
      if (ex_d1) basep[ob[WFB_D1] + k] = rd1;
      if (ex_d2) basep[ob[WFB_D2] + k] = rd2;
    }
  }
#if defined(WFB_PHASE_TIMERS) && !defined(WFB_EMU)
  const long long pt1 = WFB_PT_CLOCK();
#endif
  /* trimmed [lo,hi] of each component = min / max diagonal holding an in-bounds offset */
  {
    const int lane = wfb_lane();
    const bool exs[5] = {true, ex_i1, ex_i2, ex_d1, ex_d2};
    const int vak = wfb_warp_max(tak);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      if (!exs[c]) continue;
      int v = wfb_warp_min(tlo[c]); if (lane == 0 && v != INT_MAX) wfb_smem_min(&ring.lo[slot][c], v);
      v = wfb_warp_max(thi[c]);     if (lane == 0 && v != INT_MIN) wfb_smem_max(&ring.hi[slot][c], v);
      if (lane == 0 && vak != INT_MIN) wfb_smem_max(&ring.mak[slot][c], vak);
    }
    const int v = wfb_warp_max(tmax);
    if (lane == 0) {
      if (v > 0) wfb_smem_max(&red_maxak[par], v);
      wfb_smem_max(&ring.mak[slot][WFB_M], v); /* extended M cells can exceed the raw bound */
    }
  }
#if defined(WFB_PHASE_TIMERS) && !defined(WFB_EMU)
  if (tid == 0) {
    const int wd = hi - lo + 1, b = wd <= 128 ? 0 : wd <= 1024 ? 1 : wd <= 4096 ? 2 : 3;
    WFB_PT_ADD(4 * b + 0, 1); WFB_PT_ADD(4 * b + 1, WFB_PT_CLOCK() - pt0); WFB_PT_ADD(4 * b + 2, pt1 - pt0); WFB_PT_ADD(4 * b + 3, wd);
  }
#endif
}

/* Second half of a score step, AFTER the barrier that follows wfb_step_work: every thread of the CTA derives the
 * (uniform) outcome from shared memory. wavefront_termination_end2end, wavefront_termination.c:37-114. */
WFB_DEV int wfb_step_finish(const WfbRing& ring, const WfbPen& pen, int score, int plen, int tlen, int cend, int& num_null,
                            const int* red_maxak, const int* red_end, int& max_ak_out, WfbAcc& acc, const int slot /* score % R */,
                            const int par /* score % 3 */, int* width_out = nullptr) {
  (void)score;
  acc.steps += 1;
  max_ak_out = 0;
  if (width_out) *width_out = ring.cw[slot];
  if (!ring.ex[slot][WFB_M]) { /* the all-null step */
    num_null++;
    return (num_null > pen.scope) ? WFB_ST_END_UNREACHABLE : WFB_ST_OK;
  }
  num_null = 0;
  acc.cells += (unsigned long long)ring.cw[slot];
  max_ak_out = red_maxak[par];
  const int ak_end = tlen - plen;
  if (cend >= 0 && ring.ex[slot][cend] && ring.lo[slot][cend] <= ak_end && ak_end <= ring.hi[slot][cend]) {
    /* ak_end lies inside the trimmed range => it was computed in this step => red_end[par] is fresh */
    if (red_end[par] >= tlen) return WFB_ST_END_REACHED;
  }
  return WFB_ST_OK;
}

/* The whole CTA on one direction: work, ONE barrier, outcome. */
template <class Alloc>
WFB_STEP_INLINE int wfb_step(WfbRing& ring, int32_t* basep, const WfbPen& pen, int score, const uint8_t* pseq,
                     const uint8_t* tseq, int plen, int tlen, int cend, int& num_null, Alloc& alloc, int* red_maxak,
                     int* red_end, int& max_ak_out, WfbAcc& acc) {
  const int slot = score % pen.R, par = score % 3;
  wfb_step_work(ring, basep, pen, score, pseq, tseq, plen, tlen, cend, alloc, red_maxak, red_end, acc, WFB_TID, WFB_NT, slot, par);
  WFB_SYNC();
  return wfb_step_finish(ring, pen, score, plen, tlen, cend, num_null, red_maxak, red_end, max_ak_out, acc, slot, par);
}

/* ------------------------------------------------------------------------------------------------
 * Breakpoint overlap scan: wavefront_bialign_overlap (:877-955) with breakpoint_indel2indel
 * (:508-571) / breakpoint_m2m (:828-872). All candidate (score_i, component) pairs are scanned in
 * parallel for their first satisfying diagonal; thread 0 then replays the reference's sequential
 * "first strictly better candidate wins" rule over those results. Two barriers.
 * `found` (shared, scope*5 ints) must be all INT_MAX on entry and is restored on exit.
 * ---------------------------------------------------------------------------------------------- */
WFB_DEV int wfb_gap_of(const WfbPen& pen, int comp) {
  return comp == WFB_M ? 0 : ((comp == WFB_I1 || comp == WFB_D1) ? pen.o1 : pen.o2);
}

/* Shared scratch of the overlap scan. */
struct WfbOverlapShared {
  int found[WFB_RMAX * 5]; /* first satisfying k0 per candidate q = i*5+j; INT_MAX = none (kept clean between calls) */
  int cand[WFB_RMAX * 5];  /* compact list of candidate q's that survive the cheap tests */
  int ncand;
  unsigned hitmask[(WFB_RMAX * 5 + 31) / 32]; /* bit q set <=> found[q] valid */
};

/* Block maxima of one wavefront's rows (all threads, no barrier inside): bm[(slot * 5 + c) * NB + b] = max offset over the row
 * positions p = k + kshift in block b (64 positions) that lie in the component's trimmed range. See wfb_overlap. */
#define WFB_BM_SHIFT 6
WFB_DEV void wfb_row_bmax(const WfbRing& r, int slot, const int32_t* base, int32_t* bm, int NB, int kshift) {
#ifndef WFB_EMU
  const int nwarps = WFB_NT >> 5, warp_id = WFB_TID >> 5, lane = WFB_TID & 31;
  for (int c = 0; c < 5; ++c) {
    if (!r.ex[slot][c]) continue;
    const int lo = r.lo[slot][c], hi = r.hi[slot][c];
    if (lo > hi) continue;
    const int32_t* p = base + r.boff[slot][c];
    int32_t* out = bm + (slot * 5 + c) * NB;
    const int bfirst = (lo + kshift) >> WFB_BM_SHIFT, blast = (hi + kshift) >> WFB_BM_SHIFT;
    for (int b = bfirst + 2 * warp_id; b <= blast; b += 2 * nwarps) { /* two blocks per trip: four independent loads per lane */
      const int ka = (b << WFB_BM_SHIFT) - kshift + lane;
      const int32_t v0 = (ka >= lo && ka <= hi) ? p[ka] : WFB_OFFSET_NULL;
      const int32_t v1 = (ka + 32 >= lo && ka + 32 <= hi) ? p[ka + 32] : WFB_OFFSET_NULL;
      const int32_t v2 = (b + 1 <= blast && ka + 64 >= lo && ka + 64 <= hi) ? p[ka + 64] : WFB_OFFSET_NULL;
      const int32_t v3 = (b + 1 <= blast && ka + 96 >= lo && ka + 96 <= hi) ? p[ka + 96] : WFB_OFFSET_NULL;
      const int m0 = wfb_warp_max(max(v0, v1)), m1 = wfb_warp_max(max(v2, v3));
      if (lane == 0) { out[b] = m0; if (b + 1 <= blast) out[b + 1] = m1; }
    }
  }
#else
  (void)r; (void)slot; (void)base; (void)bm; (void)NB; (void)kshift;
#endif
}

/* bm0 / bm1: block maxima of the rows of r0's / r1's direction (nullptr = none kept: every diagonal is tested). With them a
 * candidate's diagonal range is first walked block by block: off0[k0] + off1[kinv - k0] >= tlen needs
 * max(block of r0) + max(the one or two blocks of r1 that face it) >= tlen — a necessary condition, so skipping a block never loses a
 * hit, and blocks are visited in ascending k0, so the first hit found is still the reference's. While the two directions are
 * far from meeting (all of phase 2 but its last steps) almost every block is skipped: the reference's hottest function
 * (wavefront_bialign_breakpoint_indel2indel, 28 % of its alignment time) turns into a walk over 1/64 of the data. */
WFB_STEP_INLINE void wfb_overlap(const WfbRing& r0, const int32_t* base0, const WfbRing& r1, const int32_t* base1,
                         const WfbPen& pen, int score_0, int score_1, bool bp_forward, int plen, int tlen,
                         WfbBreakpoint* bp, WfbOverlapShared* os, WfbAcc& acc, int32_t* bm0 = nullptr, const int32_t* bm1 = nullptr,
                         int NB = 0, int kshift = 0) {
  const int R = pen.R, s0 = score_0 % R;
  if (!r0.ex[s0][WFB_M]) return; /* uniform */
  if (bm0) wfb_row_bmax(r0, s0, base0, bm0, NB, kshift); /* the newest wavefront's maxima: used now, and by the next `scope` calls of the other direction */
  const int s1cur = score_1 % R; /* slots of score_1 - i follow by subtraction (i < scope < R) */
#if defined(WFB_PHASE_TIMERS) && !defined(WFB_EMU)
  const long long pto = WFB_PT_CLOCK();
  struct PtScope { long long t0; WFB_DEV_MEMBER ~PtScope() { if (WFB_TID == 0) { WFB_PT_ADD(16, WFB_PT_CLOCK() - t0); WFB_PT_ADD(17, 1); } } } pt_scope{pto};
#endif
  const int best0 = bp->score;
  const int kinv = tlen - plen;
  const int npairs = pen.scope * 5;
  /* 1) one thread per candidate (score_i, component): O(1) tests in parallel instead of a serial loop.
   *    q = i*5 + j with j indexing the reference's scan order D2, I2, D1, I1, M (:911-953). */
  for (int q = WFB_TID; q < npairs; q += WFB_NT) {
    const int i = q / 5, j = q - i * 5;
    const int c = j == 0 ? WFB_D2 : j == 1 ? WFB_I2 : j == 2 ? WFB_D1 : j == 3 ? WFB_I1 : WFB_M;
    const int score_i = score_1 - i;
    bool ok = score_i >= 0;
    if (ok) {
      const int si = wfb_slot_back(s1cur, i, R);
      ok = (score_0 + score_i - wfb_gap_of(pen, c) < best0) && r0.ex[s0][c] && r1.ex[si][c];
      /* off0[k0] + off1[kinv-k0] >= tlen  <=>  ak0 + ak1 >= plen + tlen (ak = 2*off - k): no diagonal can
       * satisfy it unless the two wavefronts' maximal anti-diagonals do */
      ok = ok && ((long long)r0.mak[s0][c] + (long long)r1.mak[si][c] >= (long long)plen + tlen);
      if (ok) {
        const int lo_0 = r0.lo[s0][c], hi_0 = r0.hi[s0][c];
        const int lo1r = r1.lo[si][c], hi1r = r1.hi[si][c];
        ok = lo_0 <= hi_0 && lo1r <= hi1r && min(hi_0, kinv - lo1r) >= max(lo_0, kinv - hi1r);
      }
    }
    if (ok) os->cand[wfb_atomic_add(&os->ncand, 1)] = q;
  }
  WFB_SYNC();
  const int ncand = os->ncand;
  if (ncand == 0) return; /* uniform: nothing can overlap yet (ncand stays 0 for the next call) */
  /* 2) candidates are dealt round-robin to the warps; a warp scans its pair's diagonal range 32 at a
   *    time and stops at the first chunk holding a hit (ballot + ffs = lowest satisfying k0, exactly the
   *    reference's ascending scan :530-569, :841-870) */
#ifndef WFB_EMU
  const int nwarps = WFB_NT >> 5, warp_id = WFB_TID >> 5, lane = WFB_TID & 31;
#else
  const int nwarps = 1, warp_id = 0, lane = 0;
#endif
  for (int ci = warp_id; ci < ncand; ci += nwarps) {
    const int q = os->cand[ci];
    const int i = q / 5, j = q - i * 5;
    const int c = j == 0 ? WFB_D2 : j == 1 ? WFB_I2 : j == 2 ? WFB_D1 : j == 3 ? WFB_I1 : WFB_M;
    const int si = wfb_slot_back(s1cur, i, R);
    const int lo_0 = r0.lo[s0][c], hi_0 = r0.hi[s0][c];
    const int lo_1 = kinv - r1.hi[si][c], hi_1 = kinv - r1.lo[si][c];
    const int max_lo = max(lo_0, lo_1), min_hi = min(hi_0, hi_1);
    const int32_t* p0 = base0 + r0.boff[s0][c];
    const int32_t* p1 = base1 + r1.boff[si][c];
    int kfound = INT_MAX;
#ifndef WFB_EMU
    unsigned long long tested = 0; /* diagonals (and block bounds) actually examined */
    if (bm0) {
      const int32_t* m0row = bm0 + (s0 * 5 + c) * NB;
      const int32_t* m1row = bm1 + (si * 5 + c) * NB;
      const int bfirst = (max_lo + kshift) >> WFB_BM_SHIFT, blast = (min_hi + kshift) >> WFB_BM_SHIFT;
      for (int bb = bfirst; bb <= blast && kfound == INT_MAX; bb += 32) {
        const int b = bb + lane;
        bool flag = false;
        if (b <= blast) {
          const int k0a = max((b << WFB_BM_SHIFT) - kshift, max_lo), k0b = min((b << WFB_BM_SHIFT) + 63 - kshift, min_hi);
          const int p1lo = kinv - k0b + kshift, p1hi = kinv - k0a + kshift; /* row positions of r1 facing [k0a, k0b] */
          const int m1 = max(m1row[p1lo >> WFB_BM_SHIFT], m1row[p1hi >> WFB_BM_SHIFT]);
          flag = m0row[b] + m1 >= tlen;
        }
        unsigned bal = __ballot_sync(0xffffffffu, flag);
        tested += (unsigned)min(32, blast - bb + 1);
        while (bal && kfound == INT_MAX) { /* exact test of the flagged blocks, lowest first */
          const int bx = bb + __ffs((int)bal) - 1;
          bal &= bal - 1;
          const int k0a = max((bx << WFB_BM_SHIFT) - kshift, max_lo), k0b = min((bx << WFB_BM_SHIFT) + 63 - kshift, min_hi);
          tested += (unsigned)(k0b - k0a + 1);
          int32_t o0[2], o1[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k0 = k0a + 32 * u + lane;
            const bool in = k0 <= k0b;
            o0[u] = in ? p0[k0] : WFB_OFFSET_NULL;
            o1[u] = in ? p1[kinv - k0] : WFB_OFFSET_NULL;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k0 = k0a + 32 * u + lane;
            bool hit = o0[u] + o1[u] >= tlen;
            if (hit && c != WFB_M) {
              const int kk = bp_forward ? k0 : kinv - k0;
              const int32_t oo = bp_forward ? o0[u] : o1[u];
              if (oo - kk > plen || oo > tlen) hit = false;
            }
            const unsigned hb = __ballot_sync(0xffffffffu, hit);
            if (hb && kfound == INT_MAX) kfound = k0a + 32 * u + __ffs((int)hb) - 1;
          }
        }
      }
    } else {
    tested = (unsigned long long)(min_hi - max_lo + 1);
    /* WFB_OVL_UNROLL chunks of 32 diagonals per trip: their loads are independent and in flight together (the scan
     * is a chain of memory round trips otherwise); the ballots are then examined in ascending order, so the first
     * satisfying diagonal is still the lowest one */
    for (int kb = max_lo; kb <= min_hi && kfound == INT_MAX; kb += 32 * WFB_OVL_UNROLL) {
      int32_t o0[WFB_OVL_UNROLL], o1[WFB_OVL_UNROLL];
#pragma unroll
      for (int u = 0; u < WFB_OVL_UNROLL; ++u) {
        const int k0 = kb + 32 * u + lane;
        const bool in = k0 <= min_hi;
        o0[u] = in ? p0[k0] : WFB_OFFSET_NULL;
        o1[u] = in ? p1[kinv - k0] : WFB_OFFSET_NULL;
      }
#pragma unroll
      for (int u = 0; u < WFB_OVL_UNROLL; ++u) {
        const int k0 = kb + 32 * u + lane;
        bool hit = o0[u] + o1[u] >= tlen; /* two nulls: -2^30 - 2^30 stays negative */
        if (hit && c != WFB_M) {
          const int kk = bp_forward ? k0 : kinv - k0;
          const int32_t oo = bp_forward ? o0[u] : o1[u];
          if (oo - kk > plen || oo > tlen) hit = false; /* out-of-bounds coordinates: keep scanning */
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal && kfound == INT_MAX) kfound = kb + 32 * u + __ffs((int)bal) - 1;
      }
    }
    }
#else
    const unsigned long long tested = (unsigned long long)(min_hi - max_lo + 1);
    for (int k0 = max_lo; k0 <= min_hi; ++k0) { /* one diagonal per trip in the single-thread emulation */
      const int k1 = kinv - k0;
      const int32_t o0 = p0[k0], o1 = p1[k1];
      if (o0 + o1 >= tlen) {
        bool hit = true;
        if (c != WFB_M) {
          const int kk = bp_forward ? k0 : k1;
          const int32_t oo = bp_forward ? o0 : o1;
          if (oo - kk > plen || oo > tlen) hit = false;
        }
        if (hit) { kfound = k0; break; }
      }
    }
#endif
    if (lane == 0) {
      if (kfound != INT_MAX) {
        os->found[q] = kfound;
#ifndef WFB_EMU
        atomicOr(&os->hitmask[q >> 5], 1u << (q & 31));
#else
        os->hitmask[q >> 5] |= 1u << (q & 31);
#endif
      }
      acc.overlap += tested;
    }
  }
  WFB_SYNC();
  /* 3) thread 0 replays the reference's sequential rule over the hits in scan order: a candidate
   *    replaces the breakpoint only if it is strictly better than the best so far (:527,:564,:912-947;
   *    with 0 <= o1 <= o2, checked on the host, the per-class `continue`s equal per-candidate tests) */
  if (WFB_TID == 0) {
    for (int w = 0; w < (npairs + 31) / 32; ++w) {
      unsigned m = os->hitmask[w];
      os->hitmask[w] = 0;
      while (m) {
#ifndef WFB_EMU
        const int b = __ffs((int)m) - 1;
#else
        const int b = __builtin_ctz(m);
#endif
        m &= m - 1;
        const int q = w * 32 + b;
        const int i = q / 5, j = q - i * 5;
        const int c = j == 0 ? WFB_D2 : j == 1 ? WFB_I2 : j == 2 ? WFB_D1 : j == 3 ? WFB_I1 : WFB_M;
        const int score_i = score_1 - i, si = wfb_slot_back(s1cur, i, R);
        const int cand = score_0 + score_i - wfb_gap_of(pen, c);
        const int k0 = os->found[q];
        os->found[q] = INT_MAX;
        if (cand >= bp->score) continue;
        const int k1 = kinv - k0;
        const int32_t o0 = (base0 + r0.boff[s0][c])[k0];
        const int32_t o1 = (base1 + r1.boff[si][c])[k1];
        if (bp_forward) {
          bp->score_forward = score_0; bp->score_reverse = score_i;
          bp->k_forward = k0; bp->k_reverse = k1;
          bp->offset_forward = o0; bp->offset_reverse = o1;
        } else {
          bp->score_forward = score_i; bp->score_reverse = score_0;
          bp->k_forward = k1; bp->k_reverse = k0;
          bp->offset_forward = o1; bp->offset_reverse = o0;
        }
        bp->score = cand;
        bp->component = c;
      }
    }
    os->ncand = 0;
  }
  WFB_SYNC();
}

/* ------------------------------------------------------------------------------------------------
 * Task plumbing
 * ---------------------------------------------------------------------------------------------- */
/* Where new sub-problems go: the next level's two queues (level-synchronous driver) or ONE persistent
 * queue drained by the same running kernel (wfb_persist_kernel). */
struct WfbPQueue {
  WfbTask* tasks;
  int* ready;       /* ready[i] = 1 once tasks[i] is published */
  int* head;        /* next slot to claim   */
  int* tail;        /* next slot to publish */
  int* outstanding; /* tasks published and not yet finished */
  int* error;       /* != 0: capacity overflow / watchdog */
  int cap;
};
struct WfbSink {
  WfbQueue q_break, q_base;
  WfbPQueue pq;
  int persistent;
};

WFB_DEV void wfb_push(const WfbQueue& q, const WfbTask& t, int* pair_status) {
  const int idx = wfb_atomic_add(q.count, 1);
  if (idx < q.cap) q.tasks[idx] = t;
  else pair_status[t.pair] = WFB_PAIR_QUEUE_OVERFLOW;
}
WFB_DEV void wfb_ppush(const WfbPQueue& q, const WfbTask& t, int* pair_status) {
  wfb_atomic_add(q.outstanding, 1); /* before the parent retires */
  const int idx = wfb_atomic_add(q.tail, 1);
  if (idx >= q.cap) {
    pair_status[t.pair] = WFB_PAIR_QUEUE_OVERFLOW;
    *q.error = 1;
    wfb_atomic_add(q.outstanding, -1);
    return;
  }
  q.tasks[idx] = t;
#ifndef WFB_EMU
  __threadfence();
  atomicExch(&q.ready[idx], 1);
#else
  q.ready[idx] = 1;
#endif
}
/* thread 0 only */
WFB_DEV void wfb_sink_push(const WfbSink& s, const WfbTask& t, int* pair_status) {
  if (s.persistent) wfb_ppush(s.pq, t, pair_status);
  else if (t.score_remaining <= WFB_FALLBACK_MIN_SCORE) wfb_push(s.q_base, t, pair_status);
  else wfb_push(s.q_break, t, pair_status);
}

/* Fill n ops of kind `op` starting at DP cell (v,h) (absolute pair coordinates); all threads. */
WFB_DEV void wfb_fill_ops(char* ops, int v, int h, char op, int n) {
  const int stride = (op == 'M' || op == 'X') ? 2 : 1;
  for (int j = WFB_TID; j < n; j += WFB_NT) ops[v + h + j * stride] = op;
}

/* Dispatch one child sub-problem (wavefront_bialign_alignment :1159-1170): trivial cases are
 * written immediately (all threads), the rest is queued by thread 0. */
WFB_DEV void wfb_dispatch_child(const WfbTask& c, char* ops, const WfbSink& sink, int* pair_status) {
  const int plen = c.pe - c.pb, tlen = c.te - c.tb;
  if (tlen == 0) {
    wfb_fill_ops(ops, c.pb, c.tb, 'D', plen);
  } else if (plen == 0) {
    wfb_fill_ops(ops, c.pb, c.tb, 'I', tlen);
  } else if (WFB_TID == 0) {
    wfb_sink_push(sink, c, pair_status);
  }
}

struct WfbBreakShared {
  WfbRing ring[2];
  WfbBreakpoint bp;
  WfbOverlapShared os;
  int red_maxak[2][3];
  int red_end[2][3];
  int task_idx;
};

struct WfbAllocFixed { /* breakpoint kernel: every (slot, component) has a fixed row of W ints */
  int dirbase; /* dir * R * 5 * W + kshift */
  int W;
  int kalign;  /* (k + kalign) % 4 == 0  <=>  cell(k) is 16-byte aligned, the same for every row (W % 8 == 0) */
  const uint32_t* sp; /* 2-bit packed pattern / text of this direction in shared memory (nullptr: read the bytes from global) */
  const uint32_t* st;
  int spo, sto;       /* base index of pattern[0] / text[0] inside sp / st */
  static constexpr unsigned char* runflag = nullptr;
  static const int runbias = 0;
  static const bool kFixedRows = true; /* the row of (slot, component) is known before the step that fills it */
  WFB_DEV_MEMBER int row(int slot, int c) const { return dirbase + (slot * 5 + c) * W; }
  WFB_DEV_MEMBER void operator()(int slot, int lo, int hi, int ob[5]) const {
    (void)lo; (void)hi;
    for (int c = 0; c < 5; ++c) ob[c] = dirbase + (slot * 5 + c) * W;
  }
};


/* ------------------------------------------------------------------------------------------------
 * Team mode: SEVERAL CTAs on ONE score step of a wide wavefront.
 *
 * Why: a record whose alignment runs through tens of kilobases of non-homologous sequence has a score of several 10^4 and
 * wavefronts of several 10^4 diagonals; its breakpoint task is one long chain of wide score steps that a single CTA needs ~10 s
 * for (measured on scerevisiae8: six such records hold the kernel for 5 s after every other record has finished). Nothing in a
 * step depends on another diagonal of the same step, so idle CTAs of the persistent grid lend a hand:
 *   * the task's owner CTA posts how many helpers it could use (slot.want) on a board in global memory;
 *   * a CTA waiting for its next task attaches to an owner (slot.members) and from then on serves its steps;
 *   * per step the owner computes the uniform part (inputs, range, output rows) exactly as wfb_step_work does, publishes it
 *     as a WfbStepDesc, and everybody — owner included — claims chunks of WFB_TEAM_CHUNK diagonals with a CAS ticket;
 *     a chunk is the same 4-diagonal-group pass as in wfb_step_work, reading / writing the owner's rows in global memory;
 *   * trim ranges / anti-diagonal maxima are reduced with global atomics into the slot, `done` counts finished chunks; the owner
 *     waits for done == total, folds the reductions into its shared-memory ring and finishes the step as usual.
 * Every cell is computed by exactly one thread with exactly the arithmetic of wfb_step_work, so the results do not depend on who
 * helped. Two buffers (epoch parity) keep a helper that is still reading step e out of the way of step e + 1; the ticket
 * carries the epoch, so a stale helper can never claim a chunk of a later step. Visibility: writers fence before counting a chunk
 * done, readers fence after claiming a ticket (gpu-scope fences also invalidate the SM's L1, B300_MICROARCH.md "L1D flush trigger").
 * ---------------------------------------------------------------------------------------------- */
#ifndef WFB_TEAM_CHUNK
#define WFB_TEAM_CHUNK 1024 /* diagonals per ticket: one 4-diagonal group per thread of a 256-thread CTA */
#endif
#ifndef WFB_TEAM_MIN_WIDTH
#define WFB_TEAM_MIN_WIDTH 4096 /* wavefront width (both directions together in phase 1) from which help is worth asking for */
#endif
#ifndef WFB_TEAM_MAX_HELPERS
#define WFB_TEAM_MAX_HELPERS 47
#endif

struct WfbStepDesc { /* the uniform part of one score step of one direction */
  WfbIn in[7];        /* m_misms, m_open1, m_open2, i1_ext, i2_ext, d1_ext, d2_ext */
  int ob[5];
  int lo, hi, safe_lo, safe_hi, kfirst;
  int plen, tlen, cend, nchunks;
  unsigned flags;     /* bit i (0..6): input i is null; bits 8..11: I1, I2, D1, D2 exist */
  unsigned pad_;
  const uint8_t* pseq;
  const uint8_t* tseq;
};
struct WfbTeamRed { int lo[5], hi[5], ak, tmax, end, pad_[3]; };
struct WfbTeamSlot { /* one per CTA of the persistent grid, in global memory */
  unsigned epoch;     /* last published step */
  int want;           /* helpers the owner could use now (0 = none) */
  int members;        /* helpers attached */
  int quit;           /* the owner's task is over: helpers leave */
  int32_t* basep;     /* the owner's wavefront workspace */
  unsigned long long next[2]; /* (epoch << 32) | next chunk, per parity */
  int done[2], total[2], ndir[2];
  WfbStepDesc d[2][2]; /* [parity][direction] */
  WfbTeamRed red[2][2];
};

#ifndef WFB_EMU
WFB_DEV int wfb_ld_vol(const int* p) { return *(const volatile int*)p; }
WFB_DEV unsigned wfb_ld_volu(const unsigned* p) { return *(const volatile unsigned*)p; }

/* the uniform prologue of wfb_step_work: inputs, range, output rows, ring bookkeeping (thread 0). Returns false for the all-null step. */
WFB_DEV bool wfb_team_prologue(WfbRing& ring, const WfbPen& pen, int score, const uint8_t* pseq, const uint8_t* tseq, int plen, int tlen, int cend,
                               const WfbAllocFixed& alloc, int* red_maxak, const int slot, const int par, WfbStepDesc& D) {
  const int R = pen.R, nslot = slot + 1 == R ? 0 : slot + 1;
  const int npar = par == 2 ? 0 : par + 1;
  const int d_x = pen.x, d_o1 = pen.o1 + pen.e1, d_o2 = pen.o2 + pen.e2;
  const int s_e1 = wfb_slot_back(slot, pen.e1, R), s_e2 = wfb_slot_back(slot, pen.e2, R);
  D.in[0] = wfb_fetch_slot(ring, WFB_M, score >= d_x, wfb_slot_back(slot, d_x, R));
  D.in[1] = wfb_fetch_slot(ring, WFB_M, score >= d_o1, wfb_slot_back(slot, d_o1, R));
  D.in[2] = wfb_fetch_slot(ring, WFB_M, score >= d_o2, wfb_slot_back(slot, d_o2, R));
  D.in[3] = wfb_fetch_slot(ring, WFB_I1, score >= pen.e1, s_e1);
  D.in[4] = wfb_fetch_slot(ring, WFB_I2, score >= pen.e2, s_e2);
  D.in[5] = wfb_fetch_slot(ring, WFB_D1, score >= pen.e1, s_e1);
  D.in[6] = wfb_fetch_slot(ring, WFB_D2, score >= pen.e2, s_e2);
  unsigned flags = 0;
#pragma unroll
  for (int i = 0; i < 7; ++i) if (D.in[i].lo > D.in[i].hi) flags |= 1u << i;
  D.plen = plen; D.tlen = tlen; D.cend = cend; D.pseq = pseq; D.tseq = tseq; D.pad_ = 0;
  if ((flags & 0x7fu) == 0x7fu) { /* wavefront_compute_affine2p.c:341-351 + wavefront_extend.c:95-103 */
    if (WFB_TID == 0) {
      for (int c = 0; c < 5; ++c) {
        ring.ex[slot][c] = 0;
        ring.lo[nslot][c] = INT_MAX;
        ring.hi[nslot][c] = INT_MIN;
        ring.mak[nslot][c] = INT_MIN;
      }
      ring.cw[slot] = 0;
      red_maxak[npar] = 0;
    }
    D.flags = flags; D.nchunks = 0; D.lo = 1; D.hi = -1; D.safe_lo = 1; D.safe_hi = -1; D.kfirst = 0;
    for (int c = 0; c < 5; ++c) D.ob[c] = 0;
    return false;
  }
  const bool n_o1 = flags & 2u, n_o2 = flags & 4u, n_i1 = flags & 8u, n_i2 = flags & 16u, n_d1 = flags & 32u, n_d2 = flags & 64u;
  int lo = D.in[0].lo, hi = D.in[0].hi;
  lo = min(lo, D.in[1].lo - 1); hi = max(hi, D.in[1].hi + 1);
  lo = min(lo, D.in[3].lo + 1); hi = max(hi, D.in[3].hi + 1);
  lo = min(lo, D.in[5].lo - 1); hi = max(hi, D.in[5].hi - 1);
  lo = min(lo, D.in[2].lo - 1); hi = max(hi, D.in[2].hi + 1);
  lo = min(lo, D.in[4].lo + 1); hi = max(hi, D.in[4].hi + 1);
  lo = min(lo, D.in[6].lo - 1); hi = max(hi, D.in[6].hi - 1);
  const bool ex_i1 = !n_o1 || !n_i1, ex_d1 = !n_o1 || !n_d1, ex_i2 = !n_o2 || !n_i2, ex_d2 = !n_o2 || !n_d2;
  if (ex_i1) flags |= 1u << 8;
  if (ex_i2) flags |= 1u << 9;
  if (ex_d1) flags |= 1u << 10;
  if (ex_d2) flags |= 1u << 11;
  alloc(slot, lo, hi, D.ob);
  if (WFB_TID == 0) {
    ring.ex[slot][WFB_M] = 1;
    ring.ex[slot][WFB_I1] = ex_i1;
    ring.ex[slot][WFB_I2] = ex_i2;
    ring.ex[slot][WFB_D1] = ex_d1;
    ring.ex[slot][WFB_D2] = ex_d2;
    for (int c = 0; c < 5; ++c) {
      ring.boff[slot][c] = D.ob[c];
      ring.lo[nslot][c] = INT_MAX;
      ring.hi[nslot][c] = INT_MIN;
      ring.mak[nslot][c] = INT_MIN;
    }
    red_maxak[npar] = 0;
    ring.cw[slot] = hi - lo + 1;
  }
  int safe_lo = lo, safe_hi = hi;
#pragma unroll
  for (int i = 0; i < 7; ++i)
    if (!(flags & (1u << i))) { safe_lo = max(safe_lo, D.in[i].lo + 1); safe_hi = min(safe_hi, D.in[i].hi - 4); }
  safe_hi = min(safe_hi, hi - 3);
  D.lo = lo; D.hi = hi; D.safe_lo = safe_lo; D.safe_hi = safe_hi;
  D.kfirst = lo - ((lo + alloc.kalign) & 3);
  D.nchunks = (hi - D.kfirst) / WFB_TEAM_CHUNK + 1;
  D.flags = flags;
  return true;
}

/* One chunk of one published step: the 4-diagonal-group pass of wfb_step_work over diagonals
 * [kfirst + chunk * WFB_TEAM_CHUNK, + WFB_TEAM_CHUNK). Accumulates the caller's trim / anti-diagonal registers. */
WFB_DEV void wfb_team_chunk(const WfbStepDesc& D, int32_t* basep, int chunk, int* red_end_global, WfbAcc& acc, int tlo[5], int thi[5], int& tak, int& tmax) {
  const WfbIn m_misms = D.in[0], m_open1 = D.in[1], m_open2 = D.in[2], i1_ext = D.in[3], i2_ext = D.in[4], d1_ext = D.in[5], d2_ext = D.in[6];
  const bool n_m = D.flags & 1u, n_o1 = D.flags & 2u, n_o2 = D.flags & 4u, n_i1 = D.flags & 8u, n_i2 = D.flags & 16u, n_d1 = D.flags & 32u, n_d2 = D.flags & 64u;
  const bool ex_i1 = D.flags & (1u << 8), ex_i2 = D.flags & (1u << 9), ex_d1 = D.flags & (1u << 10), ex_d2 = D.flags & (1u << 11);
  const int plen = D.plen, tlen = D.tlen, cend = D.cend, lo = D.lo, hi = D.hi, safe_lo = D.safe_lo, safe_hi = D.safe_hi;
  const uint8_t* const pseq = D.pseq;
  const uint8_t* const tseq = D.tseq;
  const int ak_end = tlen - plen;
  int* const red_end = red_end_global; /* WFB_END_HANDOFF writes red_end[par] */
  const int par = 0;
  struct { unsigned char* runflag; int runbias; const uint32_t* sp; const uint32_t* st; int spo, sto; } alloc = {nullptr, 0, nullptr, nullptr, 0, 0}; /* helpers read the owner's sequences from global memory */
  const int* const ob = D.ob;
  const int kbeg = D.kfirst + chunk * WFB_TEAM_CHUNK;
  const int kend = min(hi, kbeg + WFB_TEAM_CHUNK - 1);
  for (int k0 = kbeg + 4 * WFB_TID; k0 <= kend; k0 += 4 * WFB_NT) {
    int32_t rm[4], ri1[4], ri2[4], rd1[4], rd2[4];
    const bool safe = k0 >= safe_lo && k0 <= safe_hi;
    const int4 NUL4 = make_int4(WFB_OFFSET_NULL, WFB_OFFSET_NULL, WFB_OFFSET_NULL, WFB_OFFSET_NULL);
    int4 vo1 = NUL4, vo2 = NUL4, vi1 = NUL4, vi2 = NUL4, vd1 = NUL4, vd2 = NUL4, vmm = NUL4;
    int32_t so1m = WFB_OFFSET_NULL, so1p = WFB_OFFSET_NULL, so2m = WFB_OFFSET_NULL, so2p = WFB_OFFSET_NULL;
    int32_t si1 = WFB_OFFSET_NULL, si2 = WFB_OFFSET_NULL, sd1 = WFB_OFFSET_NULL, sd2 = WFB_OFFSET_NULL;
    if (safe) {
      if (!n_o1) { const int32_t* p = basep + m_open1.off + k0; vo1 = wfb_ld_row4(p); so1m = wfb_ld_row1(p - 1); so1p = wfb_ld_row1(p + 4); }
      if (!n_o2) { const int32_t* p = basep + m_open2.off + k0; vo2 = wfb_ld_row4(p); so2m = wfb_ld_row1(p - 1); so2p = wfb_ld_row1(p + 4); }
      if (!n_i1) { const int32_t* p = basep + i1_ext.off + k0; vi1 = wfb_ld_row4(p); si1 = wfb_ld_row1(p - 1); }
      if (!n_i2) { const int32_t* p = basep + i2_ext.off + k0; vi2 = wfb_ld_row4(p); si2 = wfb_ld_row1(p - 1); }
      if (!n_d1) { const int32_t* p = basep + d1_ext.off + k0; vd1 = wfb_ld_row4(p); sd1 = wfb_ld_row1(p + 4); }
      if (!n_d2) { const int32_t* p = basep + d2_ext.off + k0; vd2 = wfb_ld_row4(p); sd2 = wfb_ld_row1(p + 4); }
      if (!n_m)  { vmm = wfb_ld_row4(basep + m_misms.off + k0); }
    } else {
      vo1 = make_int4(wfb_get(basep, m_open1, k0), wfb_get(basep, m_open1, k0 + 1), wfb_get(basep, m_open1, k0 + 2), wfb_get(basep, m_open1, k0 + 3));
      so1m = wfb_get(basep, m_open1, k0 - 1); so1p = wfb_get(basep, m_open1, k0 + 4);
      vo2 = make_int4(wfb_get(basep, m_open2, k0), wfb_get(basep, m_open2, k0 + 1), wfb_get(basep, m_open2, k0 + 2), wfb_get(basep, m_open2, k0 + 3));
      so2m = wfb_get(basep, m_open2, k0 - 1); so2p = wfb_get(basep, m_open2, k0 + 4);
      vi1 = make_int4(wfb_get(basep, i1_ext, k0), wfb_get(basep, i1_ext, k0 + 1), wfb_get(basep, i1_ext, k0 + 2), WFB_OFFSET_NULL);
      si1 = wfb_get(basep, i1_ext, k0 - 1);
      vi2 = make_int4(wfb_get(basep, i2_ext, k0), wfb_get(basep, i2_ext, k0 + 1), wfb_get(basep, i2_ext, k0 + 2), WFB_OFFSET_NULL);
      si2 = wfb_get(basep, i2_ext, k0 - 1);
      vd1 = make_int4(WFB_OFFSET_NULL, wfb_get(basep, d1_ext, k0 + 1), wfb_get(basep, d1_ext, k0 + 2), wfb_get(basep, d1_ext, k0 + 3));
      sd1 = wfb_get(basep, d1_ext, k0 + 4);
      vd2 = make_int4(WFB_OFFSET_NULL, wfb_get(basep, d2_ext, k0 + 1), wfb_get(basep, d2_ext, k0 + 2), wfb_get(basep, d2_ext, k0 + 3));
      sd2 = wfb_get(basep, d2_ext, k0 + 4);
      vmm = make_int4(wfb_get(basep, m_misms, k0), wfb_get(basep, m_misms, k0 + 1), wfb_get(basep, m_misms, k0 + 2), wfb_get(basep, m_misms, k0 + 3));
    }
    WFB_KEEP_DEAD_LANES(vi1, vi2, vd1, vd2, tak)
    WFB_CELL(k0 + 0, so1m,  vo1.y, so2m,  vo2.y, si1,   si2,   vd1.y, vd2.y, vmm.x, rm[0], ri1[0], ri2[0], rd1[0], rd2[0])
    WFB_CELL(k0 + 1, vo1.x, vo1.z, vo2.x, vo2.z, vi1.x, vi2.x, vd1.z, vd2.z, vmm.y, rm[1], ri1[1], ri2[1], rd1[1], rd2[1])
    WFB_CELL(k0 + 2, vo1.y, vo1.w, vo2.y, vo2.w, vi1.y, vi2.y, vd1.w, vd2.w, vmm.z, rm[2], ri1[2], ri2[2], rd1[2], rd2[2])
    WFB_CELL(k0 + 3, vo1.z, so1p,  vo2.z, so2p,  vi1.z, vi2.z, sd1,   sd2,   vmm.w, rm[3], ri1[3], ri2[3], rd1[3], rd2[3])
    if ((unsigned)(ak_end - k0) < 4u) {
      WFB_END_HANDOFF(k0 + 0, rm[0], ri1[0], ri2[0], rd1[0], rd2[0])
      WFB_END_HANDOFF(k0 + 1, rm[1], ri1[1], ri2[1], rd1[1], rd2[1])
      WFB_END_HANDOFF(k0 + 2, rm[2], ri1[2], ri2[2], rd1[2], rd2[2])
      WFB_END_HANDOFF(k0 + 3, rm[3], ri1[3], ri2[3], rd1[3], rd2[3])
    }
    if (k0 >= lo && k0 + 3 <= hi) {
      *(int4*)(basep + ob[WFB_M] + k0) = make_int4(rm[0], rm[1], rm[2], rm[3]);
      if (ex_i1) *(int4*)(basep + ob[WFB_I1] + k0) = make_int4(ri1[0], ri1[1], ri1[2], ri1[3]);
      if (ex_i2) *(int4*)(basep + ob[WFB_I2] + k0) = make_int4(ri2[0], ri2[1], ri2[2], ri2[3]);
      if (ex_d1) *(int4*)(basep + ob[WFB_D1] + k0) = make_int4(rd1[0], rd1[1], rd1[2], rd1[3]);
      if (ex_d2) *(int4*)(basep + ob[WFB_D2] + k0) = make_int4(rd2[0], rd2[1], rd2[2], rd2[3]);
    } else {
#define WFB_STORE1(U)                                                                  \
      if (k0 + (U) >= lo && k0 + (U) <= hi) {                                            \
        basep[ob[WFB_M] + k0 + (U)] = rm[U];                                             \
        if (ex_i1) basep[ob[WFB_I1] + k0 + (U)] = ri1[U];                                \
        if (ex_i2) basep[ob[WFB_I2] + k0 + (U)] = ri2[U];                                \
        if (ex_d1) basep[ob[WFB_D1] + k0 + (U)] = rd1[U];                                \
        if (ex_d2) basep[ob[WFB_D2] + k0 + (U)] = rd2[U];                                \
      }
      WFB_STORE1(0) WFB_STORE1(1) WFB_STORE1(2) WFB_STORE1(3)
#undef WFB_STORE1
    }
  }
}

struct WfbTeamShared {
  int ticket; unsigned epoch; int flag; int entry; /* owner side: chunk ticket, mirror of the slot's epoch, a broadcast flag, board entry (-1 = none) */
  unsigned hepoch, hlast; int hcode;              /* helper side */
};

/* A CTA that has nothing to do serves the steps of owner `ts` until that task ends, its own next task (queue slot `myslot`) is published,
 * or the launch is over. Thread 0 has already incremented ts->members. */
struct WfbPQueue;
WFB_DEV_NOINLINE void wfb_team_work(WfbTeamSlot* ts, unsigned epoch, WfbTeamShared& tsh, WfbAcc& acc);
WFB_DEV_NOINLINE void wfb_team_help(WfbTeamSlot* ts, WfbTeamShared& tsh, int* my_ready, int* outstanding, int* error, WfbAcc& acc) {
  if (WFB_TID == 0) tsh.hlast = wfb_ld_volu(&ts->epoch) - 1u; /* the step in progress (if any) is served too */
  for (;;) {
    WFB_SYNC();
    if (WFB_TID == 0) {
      int code = 0;
      unsigned e = 0, spins = 0;
      for (;;) {
        e = wfb_ld_volu(&ts->epoch);
        if (e != tsh.hlast) { code = 1; break; }
        if (wfb_ld_vol(&ts->quit) || wfb_ld_vol(&ts->want) == 0) break;
        if (my_ready && wfb_ld_vol(my_ready) != 0) break;
        if (wfb_ld_vol(outstanding) <= 0 || wfb_ld_vol(error) != 0) break;
        __nanosleep(32);
        if (++spins > (1u << 27)) break;
      }
      tsh.hepoch = e; tsh.hcode = code;
      if (code) tsh.hlast = e;
    }
    WFB_SYNC();
    if (!tsh.hcode) break;
    wfb_team_work(ts, tsh.hepoch, tsh, acc);
  }
  WFB_SYNC();
  if (WFB_TID == 0) atomicSub(&ts->members, 1);
}

/* Claim and compute chunks of the step published under `epoch` until none is left; fold this CTA's reductions into the slot and count its
 * chunks done. All threads of the CTA (owner or helper). */
WFB_DEV_NOINLINE void wfb_team_work(WfbTeamSlot* ts, unsigned epoch, WfbTeamShared& tsh, WfbAcc& acc) {
  const int p = (int)(epoch & 1u);
  int mine = 0, cur_dir = -1;
  int tlo[5], thi[5], tak = INT_MIN, tmax = 0;
  auto flush = [&](int dir) {
    const unsigned flags = *(const volatile unsigned*)&ts->d[p][dir].flags;
    const bool exs[5] = {true, ((flags >> 8) & 1u) != 0, ((flags >> 9) & 1u) != 0, ((flags >> 10) & 1u) != 0, ((flags >> 11) & 1u) != 0};
    WfbTeamRed* r = &ts->red[p][dir];
    const int lane = wfb_lane();
    const int vak = wfb_warp_max(tak);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      if (!exs[c]) continue;
      int v = wfb_warp_min(tlo[c]); if (lane == 0 && v != INT_MAX) atomicMin(&r->lo[c], v);
      v = wfb_warp_max(thi[c]);     if (lane == 0 && v != INT_MIN) atomicMax(&r->hi[c], v);
    }
    if (lane == 0 && vak != INT_MIN) atomicMax(&r->ak, vak);
    const int v = wfb_warp_max(tmax);
    if (lane == 0 && v > 0) atomicMax(&r->tmax, v);
  };
  for (;;) {
    WFB_SYNC();
    if (WFB_TID == 0) {
      int t = -1;
      for (;;) {
        const unsigned long long old = *(const volatile unsigned long long*)&ts->next[p];
        if ((unsigned)(old >> 32) != epoch) break;                                   /* not (or no longer) this step */
        if ((int)(unsigned)old >= wfb_ld_vol(&ts->total[p])) break;                  /* every chunk is taken */
        if (atomicCAS(&ts->next[p], old, old + 1) == old) { t = (int)(unsigned)old; break; }
      }
      tsh.ticket = t;
    }
    WFB_SYNC();
    const int t = tsh.ticket;
    if (t < 0) break;
    __threadfence(); /* acquire: the rows earlier steps wrote (on any SM), the descriptor */
    const int n0 = wfb_ld_vol(&ts->d[p][0].nchunks);
    const int dir = t < n0 ? 0 : 1;
    if (dir != cur_dir) {
      if (cur_dir >= 0) flush(cur_dir);
#pragma unroll
      for (int c = 0; c < 5; ++c) { tlo[c] = INT_MAX; thi[c] = INT_MIN; }
      tak = INT_MIN; tmax = 0;
      cur_dir = dir;
    }
    WfbStepDesc D;
    { /* the descriptor through L2 (another SM wrote it) */
      const int* src = (const int*)&ts->d[p][dir];
      int* dst = (int*)&D;
#pragma unroll
      for (int i = 0; i < (int)(sizeof(WfbStepDesc) / sizeof(int)); ++i) dst[i] = __ldcg(src + i);
    }
    int32_t* const basep = *(int32_t* const volatile*)&ts->basep;
    wfb_team_chunk(D, basep, dir == 0 ? t : t - n0, &ts->red[p][dir].end, acc, tlo, thi, tak, tmax);
    ++mine;
  }
  if (cur_dir >= 0) flush(cur_dir);
  __threadfence(); /* release: this thread's rows and reductions before the chunk count */
  WFB_SYNC();
  if (WFB_TID == 0 && mine) atomicAdd(&ts->done[p], mine);
}

/* Owner: publish one step (one or two directions), take part, wait for the team, fold the reductions into the shared-memory rings.
 * On return every thread may call wfb_step_finish for the published directions. Returns false on a watchdog / error abort. */
struct WfbTeamDir {
  WfbRing* ring; int score; const uint8_t* pseq; const uint8_t* tseq; int cend; WfbAllocFixed alloc; int* red_maxak; int* red_end; int slot, par;
};
WFB_DEV_NOINLINE bool wfb_team_step(WfbTeamSlot* ts, WfbTeamShared& tsh, int32_t* ws, const WfbPen& pen, WfbTeamDir* dirs, int ndir, int plen, int tlen, WfbAcc& acc, int* error) {
  WfbStepDesc D[2];
  bool live[2] = {false, false};
  for (int d = 0; d < ndir; ++d)
    live[d] = wfb_team_prologue(*dirs[d].ring, pen, dirs[d].score, dirs[d].pseq, dirs[d].tseq, plen, tlen, dirs[d].cend, dirs[d].alloc, dirs[d].red_maxak,
                                dirs[d].slot, dirs[d].par, D[d]);
  if (ndir == 1) { D[1] = D[0]; D[1].nchunks = 0; }
  const int total = D[0].nchunks + (ndir == 2 ? D[1].nchunks : 0);
  const unsigned epoch = tsh.epoch + 1; /* uniform: every thread read tsh.epoch after the caller's barrier */
  const int p = (int)(epoch & 1u);
  WFB_SYNC(); /* everybody has read tsh.epoch */
  if (WFB_TID == 0) {
    tsh.epoch = epoch;
    ts->basep = ws;
    ts->total[p] = total; ts->done[p] = 0; ts->ndir[p] = ndir;
    for (int d = 0; d < 2; ++d) {
      ts->d[p][d] = D[d];
      WfbTeamRed r;
      for (int c = 0; c < 5; ++c) { r.lo[c] = INT_MAX; r.hi[c] = INT_MIN; }
      r.ak = INT_MIN; r.tmax = 0; r.end = INT_MIN; r.pad_[0] = r.pad_[1] = r.pad_[2] = 0;
      ts->red[p][d] = r;
    }
    __threadfence();
    atomicExch(&ts->next[p], (unsigned long long)epoch << 32);
    __threadfence();
    *(volatile unsigned*)&ts->epoch = epoch;
  }
  if (total > 0) wfb_team_work(ts, epoch, tsh, acc);
  bool ok = true;
  if (WFB_TID == 0) {
    unsigned spins = 0;
    while (wfb_ld_vol(&ts->done[p]) < total) {
      if (++spins > (1u << 28)) { *error = 3; ok = false; break; }
    }
    __threadfence();
    /* both directions' reductions in ONE memory round trip (32 independent L2 loads), not one per value: the fold sits on the chain of
     * every team step */
    WfbTeamRed rr[2];
    {
      const int* src = (const int*)&ts->red[p][0];
      int* dst = (int*)&rr[0];
#pragma unroll
      for (int i = 0; i < (int)(2 * sizeof(WfbTeamRed) / sizeof(int)); ++i) dst[i] = __ldcg(src + i);
    }
    for (int d = 0; d < ndir; ++d) {
      if (!live[d]) continue;
      WfbRing& ring = *dirs[d].ring;
      const int slot = dirs[d].slot;
      const WfbTeamRed* r = &rr[d];
      const bool exs[5] = {true, (bool)((D[d].flags >> 8) & 1u), (bool)((D[d].flags >> 9) & 1u), (bool)((D[d].flags >> 10) & 1u), (bool)((D[d].flags >> 11) & 1u)};
      const int ak = r->ak, tmax = r->tmax;
      for (int c = 0; c < 5; ++c) {
        if (!exs[c]) continue;
        const int lo = r->lo[c], hi = r->hi[c];
        if (lo != INT_MAX && lo < ring.lo[slot][c]) ring.lo[slot][c] = lo;
        if (hi != INT_MIN && hi > ring.hi[slot][c]) ring.hi[slot][c] = hi;
        if (ak != INT_MIN && ak > ring.mak[slot][c]) ring.mak[slot][c] = ak;
      }
      if (tmax > 0 && tmax > dirs[d].red_maxak[dirs[d].par]) dirs[d].red_maxak[dirs[d].par] = tmax;
      if (tmax > ring.mak[slot][WFB_M]) ring.mak[slot][WFB_M] = tmax;
      dirs[d].red_end[dirs[d].par] = r->end;
    }
    tsh.flag = ok ? 1 : 0;
  }
  WFB_SYNC();
  return tsh.flag != 0;
}
#endif /* !WFB_EMU */

#undef WFB_EXT_FIRST
#undef WFB_EXT_REST
#undef WFB_CELL
#undef WFB_CELL_A
#undef WFB_CELL_X
#undef WFB_END_HANDOFF

WFB_DEV void wfb_ring_reset(WfbRing& r, int R) {
  for (int i = WFB_TID; i < R * 5; i += WFB_NT) {
    r.lo[i / 5][i % 5] = INT_MAX;
    r.hi[i / 5][i % 5] = INT_MIN;
    r.ex[i / 5][i % 5] = 0;
    r.boff[i / 5][i % 5] = 0;
    r.mak[i / 5][i % 5] = INT_MIN;
  }
}

/* Score-0 wavefront + its extension (wavefront_aligner_init_wf, wavefront_aligner.c:314-383;
 * first extend of wavefront_bialign_find_breakpoint :1003-1006 / wavefront_unialign :253).
 * Thread 0 only; caller syncs. Returns via *st / *max_ak (shared). */
WFB_DEV void wfb_init_score0(WfbRing& ring, int32_t* basep, int boff0, int cbegin, int cend, const uint8_t* pseq,
                             const uint8_t* tseq, int plen, int tlen, int* st, int* max_ak, WfbAcc& acc) {
  ring.ex[0][cbegin] = 1;
  ring.lo[0][cbegin] = 0;
  ring.hi[0][cbegin] = 0;
  ring.boff[0][cbegin] = boff0;
  int32_t off = 0;
  *st = WFB_ST_OK;
  *max_ak = 0;
  if (cbegin == WFB_M) {
    const int run = wfb_match_run(pseq, tseq, min(plen, tlen));
    off = run;
    acc.matches += (unsigned)run;
    *max_ak = 2 * off;
    /* termination (:37-114): needs the end component's wavefront at score 0 */
    if (cend == WFB_M && tlen - plen == 0 && off >= tlen) *st = WFB_ST_END_REACHED;
  }
  basep[boff0 + 0] = off;
  ring.mak[0][cbegin] = 2 * off;
}

/*
 * Breakpoint kernel: one CTA per sub-problem. Restates wavefront_bialign_find_breakpoint
 * (wavefront_bialign.c:974-1082) + the dispatch of both halves (:1188-1212) + the exception path
 * (:1083-1110).
 */
#ifndef WFB_BREAK_MAXTHREADS
#define WFB_BREAK_MAXTHREADS 256
#endif
#ifndef WFB_BREAK_MINBLOCKS
#define WFB_BREAK_MINBLOCKS 2
#endif

struct WfbDirState { /* what one direction of a breakpoint task needs for a score step: read from shared memory by the warps that serve it */
  WfbAllocFixed alloc;
  const uint8_t* pseq;
  const uint8_t* tseq;
  int cend;
  int pad_;
};
struct WfbBreakCtaShared {
  WfbBreakShared sh;
  int sh_st[2];
  int sh_ak[2];
  WfbDirState dir[2]; /* [0] forward, [1] reverse; written once per task */
#ifndef WFB_EMU
  WfbTeamShared team; /* team.epoch mirrors this CTA's slot epoch for the whole launch */
#endif
};

/* One breakpoint sub-problem on one CTA: wavefront_bialign_find_breakpoint (wavefront_bialign.c:974-1082) +
 * the dispatch of both halves (:1188-1212) + the exception path (:1083-1110). */
#define WFB_TEAM_LIST 32
struct WfbTeamCtx { /* team mode of the persistent kernel (nullptr slots = off) */
  WfbTeamSlot* slots; /* one per CTA */
  int* list;          /* WFB_TEAM_LIST entries: (owner CTA + 1) of the owners that currently want help, 0 = free */
  int* error;
  int self;           /* this CTA's slot */
};

#ifndef WFB_TASK_W
#define WFB_TASK_W 1 /* rows of a task are laid out with the task's own stride (plen + tlen + 8) instead of the batch's largest */
#endif
WFB_DEV void wfb_break_task(WfbBreakCtaShared& S, const WfbTask t, int ti, const WfbPairDesc* pairs, const uint8_t* seq, int32_t* ws, int W_batch,
                            const WfbPen& pen, const WfbSink& sink, char* ops_all, int* pair_status, WfbAcc& acc, WfbTaskLog* tasklog,
                            const WfbTeamCtx* team = nullptr, uint32_t* seq_smem = nullptr, int seq_smem_words = 0, const int* pair_flags = nullptr) {
#if WFB_TASK_W
  /* a sub-problem of 3 kb x 3 kb touches 270 rows: with the batch's stride (2 x 50 kb) they are 400 KB apart — one DRAM page / TLB entry
   * each; packed with the task's own stride the whole working set is contiguous */
  const int W = min(W_batch, (((t.pe - t.pb) + (t.te - t.tb) + 8 + 7) >> 3) << 3);
#else
  const int W = W_batch;
#endif
  WfbBreakShared& sh = S.sh;
  int* const sh_st = S.sh_st;
  int* const sh_ak = S.sh_ak;
  const WfbPairDesc pd = pairs[t.pair];
  char* const ops = ops_all + pd.ops_off;
  const int plen = t.pe - t.pb, tlen = t.te - t.tb;
  const long long tl_t0 = tasklog ? wfb_globaltimer() : 0;
  const unsigned long long tl_s0 = acc.steps;
  /* trivial cases, wavefront_bialign.c:1160-1165 */
  if (tlen == 0) { wfb_fill_ops(ops, t.pb, t.tb, 'D', plen); return; }
  if (plen == 0) { wfb_fill_ops(ops, t.pb, t.tb, 'I', tlen); return; }
  /* sequence views: forward reads the forward copies, reverse the reversed copies
   * (wavefront_sequences.c:288-295) */
  const uint8_t* pf = seq + pd.p_off + t.pb;
  const uint8_t* tf = seq + pd.t_off + t.tb;
  const uint8_t* pr = seq + pd.prev_off + (pd.plen - t.pe);
  const uint8_t* tr = seq + pd.trev_off + (pd.tlen - t.te);
  const int R = pen.R;
  const int kshift = plen + 1;
  WfbAllocFixed af, ar;
  af.W = ar.W = W;
  af.kalign = ar.kalign = kshift & 3; /* dirbase = (multiple of 8) + kshift */
  af.dirbase = kshift;
  ar.dirbase = R * 5 * W + kshift;
  af.sp = af.st = ar.sp = ar.st = nullptr;
  af.spo = af.sto = ar.spo = ar.sto = 0;
  if (seq_smem && pair_flags && pair_flags[t.pair] == 0) { /* uniform */
    /* the four slices start at arbitrary bases of the pair's sequences: pack from the 16-byte chunk each one starts in */
    const long long gpf = pd.p_off + t.pb, gtf = pd.t_off + t.tb, gpr = pd.prev_off + (pd.plen - t.pe), gtr = pd.trev_off + (pd.tlen - t.te);
    const int opf = (int)(gpf & 15), otf = (int)(gtf & 15), opr = (int)(gpr & 15), otr = (int)(gtr & 15);
    const int wpf = ((opf + plen + 15) >> 4) + 2, wtf = ((otf + tlen + 15) >> 4) + 2, wpr = ((opr + plen + 15) >> 4) + 2, wtr = ((otr + tlen + 15) >> 4) + 2;
    if (wpf + wtf + wpr + wtr <= seq_smem_words) {
      uint32_t* s0 = seq_smem; uint32_t* s1 = s0 + wpf; uint32_t* s2 = s1 + wtf; uint32_t* s3 = s2 + wpr;
      WFB_SYNC(); /* the previous task's steps are done with the buffer */
      wfb_pack_seq(seq + (gpf - opf), s0, wpf - 1);
      wfb_pack_seq(seq + (gtf - otf), s1, wtf - 1);
      wfb_pack_seq(seq + (gpr - opr), s2, wpr - 1);
      wfb_pack_seq(seq + (gtr - otr), s3, wtr - 1);
      af.sp = s0; af.spo = opf; af.st = s1; af.sto = otf;
      ar.sp = s2; ar.spo = opr; ar.st = s3; ar.sto = otr;
      /* the barrier after the ring reset below orders the packing before the first step */
    }
  }
  /* Both directions' constants live in shared memory: a thread serves ONE direction per step (which one changes as the warps are
   * re-dealt), and holding the other direction's pointers and row bases in registers too pushed the step's working set into local
   * memory (spill reloads in the cell loop were 7 % of the stall samples, profiles/r02_persist_c3s8_summary.txt). */
  if (WFB_TID == 0) {
    S.dir[0].alloc = af; S.dir[0].pseq = pf; S.dir[0].tseq = tf; S.dir[0].cend = t.cend; S.dir[0].pad_ = 0;
    S.dir[1].alloc = ar; S.dir[1].pseq = pr; S.dir[1].tseq = tr; S.dir[1].cend = t.cbegin; S.dir[1].pad_ = 0;
  }
  wfb_ring_reset(sh.ring[0], R);
  wfb_ring_reset(sh.ring[1], R);
  for (int i = WFB_TID; i < WFB_RMAX * 5; i += WFB_NT) sh.os.found[i] = INT_MAX;
  if (WFB_TID < (WFB_RMAX * 5 + 31) / 32) sh.os.hitmask[WFB_TID] = 0;
  if (WFB_TID == 0) {
    for (int d = 0; d < 2; ++d) for (int j = 0; j < 3; ++j) sh.red_maxak[d][j] = 0;
    sh.bp.score = INT_MAX;
    sh.os.ncand = 0;
  }
  WFB_SYNC();
  if (WFB_TID == 0) {
    /* wavefront_bialign_init :114-143: reverse aligner swaps begin/end components */
    int ob[5];
    af(0, 0, 0, ob);
    wfb_init_score0(sh.ring[0], ws, ob[t.cbegin], t.cbegin, t.cend, pf, tf, plen, tlen, &sh_st[0], &sh_ak[0], acc);
    ar(0, 0, 0, ob);
    wfb_init_score0(sh.ring[1], ws, ob[t.cend], t.cend, t.cbegin, pr, tr, plen, tlen, &sh_st[1], &sh_ak[1], acc);
  }
  WFB_SYNC();
  int status = WFB_ST_OK; /* != OK => a direction reached the end (or is unreachable) */
  int score_reached = 0;
  int score_forward = 0, score_reverse = 0;
  int forward_max_ak = sh_ak[0], reverse_max_ak = sh_ak[1];
  int null_f = 0, null_r = 0;
  if (sh_st[0] != WFB_ST_OK) { status = sh_st[0]; score_reached = 0; }
  else if (sh_st[1] != WFB_ST_OK) { status = sh_st[1]; score_reached = 0; }
  const int max_antidiagonal = plen + tlen - 1;
  bool last_wf_forward = false;
  int max_ak = 0;
  /* phase 1 (:1010-1043): alternate until the furthest points of both directions may collide.
   * The forward step to score_forward+1 and the reverse step to score_reverse+1 do not depend on each other, so the
   * two halves of the CTA compute them at the same time and share ONE barrier (a score step is a chain of dependent
   * latencies; narrow wavefronts leave most warps idle). The reference's order is kept exactly: the forward outcome is
   * examined first, and when it ends the phase the reverse step is not accepted — nothing of it is visible afterwards
   * (its ring slot and the slot it resets lie outside the scope window because R = scope + 2, and re-running the step
   * later rewrites the same values). */
#ifndef WFB_EMU
  /* Team mode (see above): once the wavefronts are wide, post how many helpers could be used; a step is run by the team whenever
   * at least one helper is attached, by this CTA alone otherwise — the results are the same either way. */
  WfbTeamSlot* const tslot = (team && team->slots) ? team->slots + team->self : nullptr;
  bool team_posted = false;
  auto team_ready = [&](int width) -> bool { /* uniform; one barrier */
    if (!tslot || width < WFB_TEAM_MIN_WIDTH) return false;
    if (WFB_TID == 0) {
      int desired = width / WFB_TEAM_CHUNK;
      desired = desired > WFB_TEAM_MAX_HELPERS ? WFB_TEAM_MAX_HELPERS : desired;
      if (wfb_ld_vol(&tslot->want) != desired) {
        *(volatile int*)&tslot->want = desired;
        if (S.team.entry < 0) { /* put this CTA on the board (when it is full the task simply gets no help) */
          __threadfence();
          for (int i = 0; i < WFB_TEAM_LIST; ++i)
            if (atomicCAS(&team->list[i], 0, team->self + 1) == 0) { S.team.entry = i; break; }
        }
      }
      S.team.flag = wfb_ld_vol(&tslot->members) > 0 ? 1 : 0;
    }
    team_posted = true;
    WFB_SYNC();
    const bool r = S.team.flag != 0;
    WFB_SYNC(); /* S.team.flag is reused by the step */
    return r;
  };
  auto team_close = [&]() {
    if (!team_posted) return;
    WFB_SYNC();
    if (WFB_TID == 0 && wfb_ld_vol(&tslot->want) != 0) {
      if (S.team.entry >= 0) { atomicExch(&team->list[S.team.entry], 0); S.team.entry = -1; }
      *(volatile int*)&tslot->want = 0;
      *(volatile int*)&tslot->quit = 1;
      __threadfence();
      unsigned spins = 0;
      while (wfb_ld_vol(&tslot->members) > 0) { __nanosleep(64); if (++spins > (1u << 26)) { *team->error = 4; break; } }
      *(volatile int*)&tslot->quit = 0;
      __threadfence();
    }
    WFB_SYNC();
  };
  /* one direction's step by the team; d = 0 forward, 1 reverse. Ends with a barrier (like wfb_step_work + WFB_SYNC). */
  auto team_one = [&](int d, int score, int slot_i, int par_i) -> bool {
    WfbTeamDir td;
    td.ring = &sh.ring[d]; td.score = score; td.pseq = S.dir[d].pseq; td.tseq = S.dir[d].tseq; td.cend = S.dir[d].cend; td.alloc = S.dir[d].alloc;
    td.red_maxak = sh.red_maxak[d]; td.red_end = sh.red_end[d]; td.slot = slot_i; td.par = par_i;
    WfbAcc ta; ta.cells = ta.overlap = ta.matches = ta.steps = 0; /* a separate object: acc must not escape to the out-of-line call */
    const bool ok = wfb_team_step(tslot, S.team, ws, pen, &td, 1, plen, tlen, ta, team->error);
    acc.matches += ta.matches;
    return ok;
  };
#endif
#if WFB_DUAL_PHASE1
  int width_f = 1, width_r = 1;
  /* ring slot / reduction parity of the NEXT forward and reverse score, advanced by hand (no runtime modulo per step) */
  int nslot_f = (score_forward + 1) % R, npar_f = (score_forward + 1) % 3;
  int nslot_r = (score_reverse + 1) % R, npar_r = (score_reverse + 1) % 3;
  while (status == WFB_ST_OK) {
    if (forward_max_ak + reverse_max_ak >= max_antidiagonal) break;
    const unsigned long long m_before = acc.matches;
#ifndef WFB_EMU
    bool rev_half = WFB_TID >= (WFB_NT >> 1);
    if (team_ready(width_f + width_r)) {
      WfbTeamDir td[2];
      td[0].ring = &sh.ring[0]; td[0].score = score_forward + 1; td[0].pseq = S.dir[0].pseq; td[0].tseq = S.dir[0].tseq; td[0].cend = S.dir[0].cend; td[0].alloc = S.dir[0].alloc;
      td[0].red_maxak = sh.red_maxak[0]; td[0].red_end = sh.red_end[0]; td[0].slot = nslot_f; td[0].par = npar_f;
      td[1].ring = &sh.ring[1]; td[1].score = score_reverse + 1; td[1].pseq = S.dir[1].pseq; td[1].tseq = S.dir[1].tseq; td[1].cend = S.dir[1].cend; td[1].alloc = S.dir[1].alloc;
      td[1].red_maxak = sh.red_maxak[1]; td[1].red_end = sh.red_end[1]; td[1].slot = nslot_r; td[1].par = npar_r;
      WfbAcc ta; ta.cells = ta.overlap = ta.matches = ta.steps = 0;
      const bool ok = wfb_team_step(tslot, S.team, ws, pen, td, 2, plen, tlen, ta, team->error);
      acc.matches += ta.matches;
      if (!ok) { status = WFB_ST_END_UNREACHABLE; score_reached = INT_MAX; break; }
      rev_half = false; /* the statistics of a dropped speculative step stay counted in team mode */
    } else {
    /* warps are dealt to the two directions in proportion to the widths of their last wavefronts */
    const int nwarps = WFB_NT >> 5;
    int fw = nwarps >> 1; /* eighths of the CTA for the forward direction, without a division */
    if (3 * width_f > 5 * width_r) fw = (nwarps * 5) >> 3;
    if (width_f > 3 * width_r) fw = (nwarps * 3) >> 2;
    if (3 * width_r > 5 * width_f) fw = (nwarps * 3) >> 3;
    if (width_r > 3 * width_f) fw = nwarps >> 2;
    fw = fw < 1 ? 1 : (fw > nwarps - 1 ? nwarps - 1 : fw);
    const int fnt = fw << 5;
    rev_half = WFB_TID >= fnt;
    { /* one call site (one copy of the step body in the instruction cache); the direction is a per-warp choice */
      const int d = rev_half ? 1 : 0;
      const WfbDirState& ds = S.dir[d];
      const WfbAllocFixed ad = ds.alloc;
      wfb_step_work(sh.ring[d], ws, pen, (rev_half ? score_reverse : score_forward) + 1, ds.pseq, ds.tseq, plen, tlen,
                    ds.cend, ad, sh.red_maxak[d], sh.red_end[d], acc, rev_half ? WFB_TID - fnt : WFB_TID,
                    rev_half ? WFB_NT - fnt : fnt, rev_half ? nslot_r : nslot_f, rev_half ? npar_r : npar_f);
    }
    }
#else
    const bool rev_half = true; /* the single emulated thread plays both halves, one after the other */
    wfb_step_work(sh.ring[0], ws, pen, score_forward + 1, pf, tf, plen, tlen, t.cend, af, sh.red_maxak[0], sh.red_end[0], acc, 0, 1, nslot_f, npar_f);
    const unsigned long long m_mid = acc.matches;
    wfb_step_work(sh.ring[1], ws, pen, score_reverse + 1, pr, tr, plen, tlen, t.cbegin, ar, sh.red_maxak[1], sh.red_end[1], acc, 0, 1, nslot_r, npar_r);
#endif
    WFB_SYNC();
    ++score_forward;
    int st = wfb_step_finish(sh.ring[0], pen, score_forward, plen, tlen, t.cend, null_f, sh.red_maxak[0], sh.red_end[0], max_ak, acc, nslot_f, npar_f, &width_f);
    nslot_f = nslot_f + 1 == R ? 0 : nslot_f + 1; npar_f = npar_f == 2 ? 0 : npar_f + 1;
    if (forward_max_ak < max_ak) forward_max_ak = max_ak;
    last_wf_forward = true;
    const bool stop_after_forward = (st != WFB_ST_OK) || (forward_max_ak + reverse_max_ak >= max_antidiagonal);
    if (stop_after_forward) {
      /* the reverse step was speculative: forget the matches it counted (cells / steps are only counted on acceptance) */
#ifndef WFB_EMU
      if (rev_half) acc.matches = m_before;
#else
      acc.matches = m_mid;
#endif
      if (st != WFB_ST_OK) { status = st; score_reached = score_forward; }
      break;
    }
    (void)rev_half; (void)m_before;
    ++score_reverse;
    st = wfb_step_finish(sh.ring[1], pen, score_reverse, plen, tlen, t.cbegin, null_r, sh.red_maxak[1], sh.red_end[1], max_ak, acc, nslot_r, npar_r, &width_r);
    nslot_r = nslot_r + 1 == R ? 0 : nslot_r + 1; npar_r = npar_r == 2 ? 0 : npar_r + 1;
    if (reverse_max_ak < max_ak) reverse_max_ak = max_ak;
    last_wf_forward = false;
    if (st != WFB_ST_OK) { status = st; score_reached = score_reverse; break; }
  }
#else
  while (status == WFB_ST_OK) {
    if (forward_max_ak + reverse_max_ak >= max_antidiagonal) break;
    ++score_forward;
    int st = wfb_step(sh.ring[0], ws, pen, score_forward, pf, tf, plen, tlen, t.cend, null_f, af, sh.red_maxak[0], sh.red_end[0], max_ak, acc);
    if (forward_max_ak < max_ak) forward_max_ak = max_ak;
    last_wf_forward = true;
    if (st != WFB_ST_OK) { status = st; score_reached = score_forward; break; }
    if (forward_max_ak + reverse_max_ak >= max_antidiagonal) break;
    ++score_reverse;
    st = wfb_step(sh.ring[1], ws, pen, score_reverse, pr, tr, plen, tlen, t.cbegin, null_r, ar, sh.red_maxak[1], sh.red_end[1], max_ak, acc);
    if (reverse_max_ak < max_ak) reverse_max_ak = max_ak;
    last_wf_forward = false;
    if (st != WFB_ST_OK) { status = st; score_reached = score_reverse; break; }
  }
#endif
  /* phase 2 (:1045-1079): advance while scanning for overlaps */
  const int gap_opening = max(pen.o1, pen.o2);
  /* block maxima of the wavefront rows for the overlap scan (see wfb_overlap): kept behind the rows in the CTA's workspace */
  const int NB = (W + 63) >> WFB_BM_SHIFT;
  int32_t* const bmf = ws + 2 * R * 5 * W;
  int32_t* const bmr = bmf + R * 5 * NB;
  if (status == WFB_ST_OK) { /* uniform */
    for (int i = 0; i < pen.scope; ++i) {
      if (score_forward - i >= 0) wfb_row_bmax(sh.ring[0], (score_forward - i) % R, ws, bmf, NB, kshift);
      if (score_reverse - i >= 0) wfb_row_bmax(sh.ring[1], (score_reverse - i) % R, ws, bmr, NB, kshift);
    }
    WFB_SYNC();
  }
  while (status == WFB_ST_OK) {
    if (last_wf_forward) {
      const int min_score_reverse = (score_reverse > pen.scope - 1) ? score_reverse - (pen.scope - 1) : 0;
      if (score_forward + min_score_reverse - gap_opening >= sh.bp.score) break;
      wfb_overlap(sh.ring[0], ws, sh.ring[1], ws, pen, score_forward, score_reverse, true, plen, tlen, &sh.bp, &sh.os, acc, bmf, bmr, NB, kshift);
      ++score_reverse;
      int st;
#ifndef WFB_EMU
      if (team_ready(sh.ring[1].cw[(score_reverse - 1) % R])) {
        const int sl = score_reverse % R, pa = score_reverse % 3;
        if (!team_one(1, score_reverse, sl, pa)) { status = WFB_ST_END_UNREACHABLE; score_reached = INT_MAX; break; }
        st = wfb_step_finish(sh.ring[1], pen, score_reverse, plen, tlen, t.cbegin, null_r, sh.red_maxak[1], sh.red_end[1], max_ak, acc, sl, pa);
      } else
#endif
      {
        WfbAllocFixed a1 = S.dir[1].alloc;
        st = wfb_step(sh.ring[1], ws, pen, score_reverse, S.dir[1].pseq, S.dir[1].tseq, plen, tlen, t.cbegin, null_r, a1, sh.red_maxak[1], sh.red_end[1], max_ak, acc);
      }
      if (st != WFB_ST_OK) { status = st; score_reached = score_reverse; break; }
    }
    const int min_score_forward = (score_forward > pen.scope - 1) ? score_forward - (pen.scope - 1) : 0;
    if (min_score_forward + score_reverse - gap_opening >= sh.bp.score) break;
    wfb_overlap(sh.ring[1], ws, sh.ring[0], ws, pen, score_reverse, score_forward, false, plen, tlen, &sh.bp, &sh.os, acc, bmr, bmf, NB, kshift);
    ++score_forward;
    int st;
#ifndef WFB_EMU
    if (team_ready(sh.ring[0].cw[(score_forward - 1) % R])) {
      const int sl = score_forward % R, pa = score_forward % 3;
      if (!team_one(0, score_forward, sl, pa)) { status = WFB_ST_END_UNREACHABLE; score_reached = INT_MAX; break; }
      st = wfb_step_finish(sh.ring[0], pen, score_forward, plen, tlen, t.cend, null_f, sh.red_maxak[0], sh.red_end[0], max_ak, acc, sl, pa);
    } else
#endif
    {
      WfbAllocFixed a0 = S.dir[0].alloc;
      st = wfb_step(sh.ring[0], ws, pen, score_forward, S.dir[0].pseq, S.dir[0].tseq, plen, tlen, t.cend, null_f, a0, sh.red_maxak[0], sh.red_end[0], max_ak, acc);
    }
    if (st != WFB_ST_OK) { status = st; score_reached = score_forward; break; }
    last_wf_forward = true;
  }
#ifndef WFB_EMU
  team_close();
#endif
  if (tasklog && WFB_TID == 0) {
    WfbTaskLog tl;
    tl.t0 = tl_t0; tl.t1 = wfb_globaltimer(); tl.smid = wfb_smid(); tl.steps = (int)(acc.steps - tl_s0);
    tl.score_f = score_forward; tl.score_r = score_reverse; tl.plen = plen; tl.tlen = tlen; tl.status = status; tl.pad_ = 0;
    tasklog[ti] = tl;
  }
  if (status != WFB_ST_OK) {
    /* wavefront_bialign_find_breakpoint_exception :1083-1110 */
    if (WFB_TID == 0) {
      if (status == WFB_ST_END_REACHED && score_reached <= WFB_RECOVERY_MIN_SCORE) {
        WfbTask c = t;
        c.score_remaining = 0;
        wfb_sink_push(sink, c, pair_status);
      } else {
        pair_status[t.pair] = WFB_PAIR_UNATTAINABLE;
      }
    }
    return;
  }
  /* both halves, :1188-1212 */
  const WfbBreakpoint bp = sh.bp;
  const int bh = bp.offset_forward, bv = bp.offset_forward - bp.k_forward;
  WfbTask c0, c1;
  c0.pair = t.pair; c0.pb = t.pb; c0.pe = t.pb + bv; c0.tb = t.tb; c0.te = t.tb + bh;
  c0.cbegin = t.cbegin; c0.cend = bp.component; c0.score_remaining = bp.score_forward;
  c1.pair = t.pair; c1.pb = t.pb + bv; c1.pe = t.pe; c1.tb = t.tb + bh; c1.te = t.te;
  c1.cbegin = bp.component; c1.cend = t.cend; c1.score_remaining = bp.score_reverse;
  wfb_dispatch_child(c0, ops, sink, pair_status);
  wfb_dispatch_child(c1, ops, sink, pair_status);
}

WFB_KERNEL_LB(wfb_break_kernel, WFB_BREAK_MAXTHREADS, WFB_BREAK_MINBLOCKS, const WfbTask* tasks, int ntasks, int* task_counter, const WfbPairDesc* pairs,
           const uint8_t* seq, int32_t* ws_all, long long ws_stride /* ints per CTA */, int W, WfbPen pen,
           WfbQueue q_break, WfbQueue q_base, char* ops_all, int* pair_status, WfbCounters* counters, WfbTaskLog* tasklog) {
  WFB_KERNEL_PROLOGUE
  WFB_SHARED WfbBreakCtaShared S;
  int32_t* const ws = ws_all + (long long)bid * ws_stride;
  WfbSink sink;
  sink.q_break = q_break; sink.q_base = q_base; sink.persistent = 0;
  sink.pq.tasks = nullptr; sink.pq.ready = nullptr; sink.pq.head = nullptr; sink.pq.tail = nullptr; sink.pq.outstanding = nullptr;
  sink.pq.error = nullptr; sink.pq.cap = 0;
  WfbAcc acc;
  acc.cells = acc.overlap = acc.matches = acc.steps = 0;
  unsigned long long ntask_done = 0;
  for (;;) {
    WFB_SYNC();
    if (WFB_TID == 0) S.sh.task_idx = wfb_atomic_add(task_counter, 1);
    WFB_SYNC();
    const int ti = S.sh.task_idx;
    if (ti >= ntasks) break;
    ntask_done++;
    wfb_break_task(S, tasks[ti], ti, pairs, seq, ws, W, pen, sink, ops_all, pair_status, acc, tasklog);
  }
  /* counters */
  {
    unsigned long long m = acc.matches;
#ifndef WFB_EMU
    for (int o = 16; o > 0; o >>= 1) m += __shfl_down_sync(0xffffffffu, m, o);
#endif
    if (wfb_lane() == 0 && m) wfb_atomic_add64(&counters->extend_matches, m);
    if (wfb_lane() == 0 && acc.overlap) wfb_atomic_add64(&counters->overlap_tests, acc.overlap);
    if (WFB_TID == 0) {
      wfb_atomic_add64(&counters->cells, acc.cells);

      wfb_atomic_add64(&counters->score_steps, acc.steps);
      wfb_atomic_add64(&counters->break_tasks, ntask_done);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Base kernel: unidirectional full-memory WFA + backtrace on small-score sub-problems.
 * wavefront_bialign_base (:159-189) = wavefront_unialign (wavefront_unialign.c:242-273) +
 * wavefront_backtrace_affine (wavefront_backtrace.c:320-529).
 * Per-CTA global scratch:  arena (offsets of every score), a metadata log per score for the
 * backtrace, and a run list.
 * ---------------------------------------------------------------------------------------------- */
struct WfbBaseMeta { /* one per (score, component) */
  int lo, hi, boff, ex;
};

struct WfbRun {
  int idx;   /* first op slot (v+h, absolute) */
  int count;
  int op;
};

struct WfbAllocBump {
  static const int kalign = -1; /* rows are packed back to back: no common alignment, scalar path only */
  static constexpr const uint32_t* sp = nullptr; /* sequences are read from global memory */
  static constexpr const uint32_t* st = nullptr;
  static const int spo = 0, sto = 0;
  unsigned char* runflag; /* ends-free only: runflag[k + runbias] = 1 when M(k) matched >= 4 bases this step; else NULL */
  int runbias;
  int bump; /* next free int in the arena */
  static const bool kFixedRows = false;
  WFB_DEV_MEMBER int row(int slot, int c) const { (void)slot; (void)c; return 0; }
  WFB_DEV_MEMBER void operator()(int slot, int lo, int hi, int ob[5]) {
    (void)slot;
    const int n = hi - lo + 1;
    for (int c = 0; c < 5; ++c) ob[c] = bump + c * n - lo;
    bump += 5 * n;
  }
};

#define WFB_BT_SET(off, type) ((((long long)(off)) << 4) | (type))

/* wavefront_backtrace_{misms,ins*,del*} (wavefront_backtrace.c:64-219) over the metadata log */
WFB_DEV long long wfb_bt_src(const WfbBaseMeta* log, const int32_t* arena, int nscores, int comp, int score, int k, int dk,
                             int plus, int type) {
  if (score < 0 || score >= nscores) return WFB_OFFSET_NULL;
  const WfbBaseMeta m = log[score * 5 + comp];
  if (m.ex && m.lo <= k + dk && k + dk <= m.hi) return WFB_BT_SET(arena[m.boff + k + dk] + plus, type);
  return WFB_OFFSET_NULL;
}
WFB_DEV long long wfb_max64(long long a, long long b) { return a > b ? a : b; }

WFB_DEV void wfb_emit(WfbRun* runs, int& nruns, int maxruns, int& err, int op, int count, int idx) {
  if (count <= 0) return;
  if (nruns > 0 && runs[nruns - 1].op == op && (op == 'M' || op == 'X' ? idx + 2 * count : idx + count) == runs[nruns - 1].idx) {
    runs[nruns - 1].idx = idx;
    runs[nruns - 1].count += count;
    return;
  }
  if (nruns >= maxruns) { err = 1; return; }
  runs[nruns].idx = idx;
  runs[nruns].count = count;
  runs[nruns].op = op;
  ++nruns;
}

/* Thread 0: wavefront_backtrace_affine (wavefront_backtrace.c:320-529). Emits runs (right to left);
 * `base_idx` = pb + tb turns local (v,h) into the pair's op-slot index. */
WFB_DEV int wfb_backtrace(const WfbBaseMeta* log, const int32_t* arena, int nscores, const WfbPen& pen, int cbegin, int cend,
                          int plen, int tlen, int alignment_score, int end_k, int end_off, int base_idx, WfbRun* runs, int maxruns,
                          int* nruns_out) {
  enum { BT_M = 9, BT_D2_EXT = 8, BT_D2_OPEN = 7, BT_D1_EXT = 6, BT_D1_OPEN = 5, BT_I2_EXT = 4, BT_I2_OPEN = 3,
         BT_I1_EXT = 2, BT_I1_OPEN = 1 };
  (void)cbegin;
  int nruns = 0, err = 0;
  int matrix_type = cend;
  int score = alignment_score;
  int k = end_k;
  int offset = end_off;
  int v = end_off - end_k, h = end_off;
  if (cend == WFB_M) { /* ends-free tails (:347-356), written right to left: the D's end the transcript, the I's precede them */
    if (v < plen) wfb_emit(runs, nruns, maxruns, err, 'D', plen - v, base_idx + v + tlen);
    if (h < tlen) wfb_emit(runs, nruns, maxruns, err, 'I', tlen - h, base_idx + v + h);
  }
  while (v > 0 && h > 0 && score > 0) {
    const int mismatch = score - pen.x;
    const int gap_open1 = score - pen.o1 - pen.e1, gap_open2 = score - pen.o2 - pen.e2;
    const int gap_extend1 = score - pen.e1, gap_extend2 = score - pen.e2;
    long long max_all;
    if (matrix_type == WFB_M) {
      const long long misms = wfb_bt_src(log, arena, nscores, WFB_M, mismatch, k, 0, 1, BT_M);
      const long long max_ins1 = wfb_max64(wfb_bt_src(log, arena, nscores, WFB_M, gap_open1, k, -1, 1, BT_I1_OPEN),
                                           wfb_bt_src(log, arena, nscores, WFB_I1, gap_extend1, k, -1, 1, BT_I1_EXT));
      const long long max_del1 = wfb_max64(wfb_bt_src(log, arena, nscores, WFB_M, gap_open1, k, +1, 0, BT_D1_OPEN),
                                           wfb_bt_src(log, arena, nscores, WFB_D1, gap_extend1, k, +1, 0, BT_D1_EXT));
      const long long max_ins2 = wfb_max64(wfb_bt_src(log, arena, nscores, WFB_M, gap_open2, k, -1, 1, BT_I2_OPEN),
                                           wfb_bt_src(log, arena, nscores, WFB_I2, gap_extend2, k, -1, 1, BT_I2_EXT));
      const long long max_del2 = wfb_max64(wfb_bt_src(log, arena, nscores, WFB_M, gap_open2, k, +1, 0, BT_D2_OPEN),
                                           wfb_bt_src(log, arena, nscores, WFB_D2, gap_extend2, k, +1, 0, BT_D2_EXT));
      max_all = wfb_max64(misms, wfb_max64(wfb_max64(max_ins1, max_ins2), wfb_max64(max_del1, max_del2)));
    } else if (matrix_type == WFB_I1) {
      max_all = wfb_max64(wfb_bt_src(log, arena, nscores, WFB_M, gap_open1, k, -1, 1, BT_I1_OPEN),
                          wfb_bt_src(log, arena, nscores, WFB_I1, gap_extend1, k, -1, 1, BT_I1_EXT));
    } else if (matrix_type == WFB_I2) {
      max_all = wfb_max64(wfb_bt_src(log, arena, nscores, WFB_M, gap_open2, k, -1, 1, BT_I2_OPEN),
                          wfb_bt_src(log, arena, nscores, WFB_I2, gap_extend2, k, -1, 1, BT_I2_EXT));
    } else if (matrix_type == WFB_D1) {
      max_all = wfb_max64(wfb_bt_src(log, arena, nscores, WFB_M, gap_open1, k, +1, 0, BT_D1_OPEN),
                          wfb_bt_src(log, arena, nscores, WFB_D1, gap_extend1, k, +1, 0, BT_D1_EXT));
    } else {
      max_all = wfb_max64(wfb_bt_src(log, arena, nscores, WFB_M, gap_open2, k, +1, 0, BT_D2_OPEN),
                          wfb_bt_src(log, arena, nscores, WFB_D2, gap_extend2, k, +1, 0, BT_D2_EXT));
    }
    if (max_all < 0) break;
    if (matrix_type == WFB_M) {
      const int max_offset = (int)(max_all >> 4);
      const int num_matches = offset - max_offset;
      wfb_emit(runs, nruns, maxruns, err, 'M', num_matches, base_idx + (v - num_matches) + (h - num_matches));
      offset = max_offset;
      v = offset - k;
      h = offset;
      if (v <= 0 || h <= 0) break;
    }
    const int bt = (int)(max_all & 0xF);
    switch (bt) {
      case BT_M: score = mismatch; matrix_type = WFB_M; break;
      case BT_I1_OPEN: score = gap_open1; matrix_type = WFB_M; break;
      case BT_I1_EXT: score = gap_extend1; matrix_type = WFB_I1; break;
      case BT_I2_OPEN: score = gap_open2; matrix_type = WFB_M; break;
      case BT_I2_EXT: score = gap_extend2; matrix_type = WFB_I2; break;
      case BT_D1_OPEN: score = gap_open1; matrix_type = WFB_M; break;
      case BT_D1_EXT: score = gap_extend1; matrix_type = WFB_D1; break;
      case BT_D2_OPEN: score = gap_open2; matrix_type = WFB_M; break;
      case BT_D2_EXT: score = gap_extend2; matrix_type = WFB_D2; break;
      default: *nruns_out = 0; return 1;
    }
    if (bt == BT_M) { wfb_emit(runs, nruns, maxruns, err, 'X', 1, base_idx + (v - 1) + (h - 1)); --offset; }
    else if (bt <= BT_I2_EXT) { wfb_emit(runs, nruns, maxruns, err, 'I', 1, base_idx + v + (h - 1)); --k; --offset; }
    else { wfb_emit(runs, nruns, maxruns, err, 'D', 1, base_idx + (v - 1) + h); ++k; }
    v = offset - k;
    h = offset;
  }
  if (matrix_type == WFB_M) {
    if (v > 0 && h > 0) {
      const int num_matches = min(v, h);
      wfb_emit(runs, nruns, maxruns, err, 'M', num_matches, base_idx + (v - num_matches) + (h - num_matches));
      v -= num_matches;
      h -= num_matches;
    }
    if (v > 0) { wfb_emit(runs, nruns, maxruns, err, 'D', v, base_idx + 0 + h); v = 0; }
    if (h > 0) { wfb_emit(runs, nruns, maxruns, err, 'I', h, base_idx + 0 + 0); h = 0; }
  } else {
    if (v != 0 || h != 0 || score != 0) err = 1; /* the reference aborts here (:519-524) */
  }
  *nruns_out = nruns;
  return err;
}

struct WfbBaseShared {
  WfbRing ring;
  int red_maxak[3];
  int red_end[3];
  int task_idx;
  int st0, ak0;
  int nruns, bt_err;
};

/* One base sub-problem on one CTA: wavefront_bialign_base (wavefront_bialign.c:159-189). */
WFB_DEV void wfb_base_task(WfbBaseShared& sh, const WfbTask t, const WfbPairDesc* pairs, const uint8_t* seq, int32_t* arena,
                           long long arena_stride, WfbBaseMeta* log, int score_cap, WfbRun* runs, int maxruns, const WfbPen& pen,
                           char* ops_all, int* pair_status, WfbAcc& acc) {
  const WfbPairDesc pd = pairs[t.pair];
  char* const ops = ops_all + pd.ops_off;
  const int plen = t.pe - t.pb, tlen = t.te - t.tb;
  if (tlen == 0) { wfb_fill_ops(ops, t.pb, t.tb, 'D', plen); return; }
  if (plen == 0) { wfb_fill_ops(ops, t.pb, t.tb, 'I', tlen); return; }
  const uint8_t* pf = seq + pd.p_off + t.pb;
  const uint8_t* tf = seq + pd.t_off + t.tb;
  const int R = pen.R;
  wfb_ring_reset(sh.ring, R);
  if (WFB_TID == 0) { sh.red_maxak[0] = sh.red_maxak[1] = sh.red_maxak[2] = 0; }
  WFB_SYNC();
  WfbAllocBump ab;
  ab.runflag = nullptr; ab.runbias = 0;
  ab.bump = 1; /* cell 0 = the score-0 wavefront */
  if (WFB_TID == 0) {
    wfb_init_score0(sh.ring, arena, 0, t.cbegin, t.cend, pf, tf, plen, tlen, &sh.st0, &sh.ak0, acc);
    for (int c = 0; c < 5; ++c) {
      WfbBaseMeta m;
      m.lo = 0; m.hi = 0; m.boff = 0; m.ex = (c == t.cbegin);
      log[c] = m;
    }
  }
  WFB_SYNC();
  int status = sh.st0;
  int score = 0, num_null = 0, max_ak = 0;
  /* wavefront_unialign, wavefront_unialign.c:251-270 */
  while (status == WFB_ST_OK) {
    ++score;
    if (score > score_cap || (long long)ab.bump + 5LL * (2 * score + 3) > arena_stride) { status = -1; break; }
    status = wfb_step(sh.ring, arena, pen, score, pf, tf, plen, tlen, t.cend, num_null, ab, sh.red_maxak, sh.red_end, max_ak, acc);
    if (WFB_TID == 0) {
      const int slot = score % R;
      for (int c = 0; c < 5; ++c) {
        WfbBaseMeta m;
        m.lo = sh.ring.lo[slot][c]; m.hi = sh.ring.hi[slot][c]; m.boff = sh.ring.boff[slot][c]; m.ex = sh.ring.ex[slot][c];
        log[score * 5 + c] = m;
      }
    }
  }
  if (status != WFB_ST_END_REACHED) {
    if (WFB_TID == 0) pair_status[t.pair] = (status == -1) ? WFB_PAIR_BASE_SCORE_CAP : WFB_PAIR_UNATTAINABLE;
    return;
  }
  if (WFB_TID == 0) {
    int nr = 0;
    const long long pt_bt0 = WFB_PT_CLOCK();
    sh.bt_err = wfb_backtrace(log, arena, score + 1, pen, t.cbegin, t.cend, plen, tlen, score, tlen - plen, tlen, t.pb + t.tb, runs, maxruns, &nr);
    WFB_PT_ADD(22, WFB_PT_CLOCK() - pt_bt0); WFB_PT_ADD(23, 1);
    (void)pt_bt0;
    sh.nruns = nr;
  }
  WFB_SYNC();
  if (sh.bt_err) {
    if (WFB_TID == 0) pair_status[t.pair] = WFB_PAIR_BACKTRACE;
    return;
  }
  const int nr = sh.nruns;
  for (int r = 0; r < nr; ++r) {
    const WfbRun ru = runs[r];
    const int stride = (ru.op == 'M' || ru.op == 'X') ? 2 : 1;
    for (int j = WFB_TID; j < ru.count; j += WFB_NT) ops[ru.idx + j * stride] = (char)ru.op;
  }
}

WFB_KERNEL(wfb_base_kernel, const WfbTask* tasks, int ntasks, int* task_counter, const WfbPairDesc* pairs,
           const uint8_t* seq, int32_t* arena_all, long long arena_stride /* ints per CTA */, WfbBaseMeta* log_all,
           int score_cap, WfbRun* runs_all, int maxruns, WfbPen pen, char* ops_all, int* pair_status,
           WfbCounters* counters) {
  WFB_KERNEL_PROLOGUE
  WFB_SHARED WfbBaseShared sh;
  int32_t* const arena = arena_all + (long long)bid * arena_stride;
  WfbBaseMeta* const log = log_all + (long long)bid * (score_cap + 1) * 5;
  WfbRun* const runs = runs_all + (long long)bid * maxruns;
  WfbAcc acc;
  acc.cells = acc.overlap = acc.matches = acc.steps = 0;
  unsigned long long ntask_done = 0;
  for (;;) {
    WFB_SYNC();
    if (WFB_TID == 0) sh.task_idx = wfb_atomic_add(task_counter, 1);
    WFB_SYNC();
    const int ti = sh.task_idx;
    if (ti >= ntasks) break;
    ntask_done++;
    wfb_base_task(sh, tasks[ti], pairs, seq, arena, arena_stride, log, score_cap, runs, maxruns, pen, ops_all, pair_status, acc);
  }
  {
    unsigned long long m = acc.matches;
#ifndef WFB_EMU
    for (int o = 16; o > 0; o >>= 1) m += __shfl_down_sync(0xffffffffu, m, o);
#endif
    if (wfb_lane() == 0 && m) wfb_atomic_add64(&counters->base_extend_matches, m);
    if (WFB_TID == 0) {
      wfb_atomic_add64(&counters->base_cells, acc.cells);
      wfb_atomic_add64(&counters->base_score_steps, acc.steps);
      wfb_atomic_add64(&counters->base_tasks, ntask_done);
    }
  }
}


/* ------------------------------------------------------------------------------------------------
 * Persistent driver: ONE launch drains the whole recursion tree of a batch. Resident CTAs claim slots of a
 * global queue in order, wait for the slot to be published, run the task (breakpoint or base, decided by
 * score_remaining like wavefront_bialign_alignment :1166), publish its children into later slots, and exit
 * when no task is outstanding. Removes the per-level launch barrier of the level-synchronous driver: a CTA
 * that finishes early immediately continues with sub-problems of other alignments.
 * Requires every CTA of the grid to be resident (the host sizes the grid from the occupancy query).
 * ---------------------------------------------------------------------------------------------- */
struct WfbPersistShared {
  WfbBreakCtaShared brk;
  WfbBaseShared base;
  int slot;
  int help_owner;
};

WFB_KERNEL_LB(wfb_persist_kernel, WFB_BREAK_MAXTHREADS, WFB_BREAK_MINBLOCKS, WfbPQueue q, const WfbPairDesc* pairs, const uint8_t* seq,
              int32_t* ws_all, long long ws_stride, int W, int32_t* arena_all, long long arena_stride, WfbBaseMeta* log_all, int score_cap,
              WfbRun* runs_all, int maxruns, WfbPen pen, char* ops_all, int* pair_status, WfbCounters* counters,
              long long* cta_log /* optional (WFB_TRACE): per CTA {ns busy in tasks, exit time, start time, tasks} */,
              WfbTeamSlot* team_slots /* optional: team mode (one zeroed slot per CTA) */, int* team_list /* WFB_TEAM_LIST zeroed ints */,
              const int* pair_flags /* optional: 0 = the pair is pure ACGT */, int seq_smem_words /* dynamic shared memory, in words */) {
  WFB_KERNEL_PROLOGUE
  WFB_SHARED WfbPersistShared S;
#ifndef WFB_EMU
  extern __shared__ uint32_t wfb_seq_smem[];
  uint32_t* const seq_smem = seq_smem_words > 0 ? wfb_seq_smem : nullptr;
#endif
#ifndef WFB_EMU
  WfbTeamCtx tctx;
  tctx.slots = team_slots; tctx.list = team_list; tctx.error = q.error; tctx.self = bid;
  if (WFB_TID == 0) { S.brk.team.epoch = 0; S.brk.team.entry = -1; S.brk.team.flag = 0; }
  int held_slot = -1; /* thread 0: a queue slot claimed before an excursion as a helper */
  long long log_help = 0;
#endif
  const long long log_t0 = cta_log ? wfb_globaltimer() : 0;
  long long log_busy = 0;
  int32_t* const ws = ws_all + (long long)bid * ws_stride;
  int32_t* const arena = arena_all + (long long)bid * arena_stride;
  WfbBaseMeta* const log = log_all + (long long)bid * (score_cap + 1) * 5;
  WfbRun* const runs = runs_all + (long long)bid * maxruns;
  WfbSink sink;
  sink.persistent = 1;
  sink.pq = q;
  sink.q_break.tasks = nullptr; sink.q_break.count = nullptr; sink.q_break.cap = 0;
  sink.q_base = sink.q_break;
  WfbAcc acc, acc_base;
  acc.cells = acc.overlap = acc.matches = acc.steps = 0;
  acc_base = acc;
  unsigned long long n_break = 0, n_base = 0;
  for (;;) {
    WFB_SYNC();
    if (WFB_TID == 0) {
      const long long pt_idle0 = WFB_PT_CLOCK();
      (void)pt_idle0;
#ifndef WFB_EMU
      int slot = held_slot >= 0 ? held_slot : wfb_atomic_add(q.head, 1);
      held_slot = -1;
#else
      int slot = wfb_atomic_add(q.head, 1);
#endif
      if (slot >= q.cap) {
        slot = -1;
      } else {
#ifndef WFB_EMU
        unsigned spins = 0;
        while (atomicAdd(&q.ready[slot], 0) == 0) {
          if (atomicAdd(q.outstanding, 0) <= 0 || atomicAdd(q.error, 0) != 0) { slot = -1; break; }
          if (team_slots && (spins & 3u) == 0) { /* nothing to do: does a wide task want help? (every 4th poll: an idle CTA's instructions
                                                    compete with the working CTA of the same SM) */
            int owner = -1;
            for (int i = 0; i < WFB_TEAM_LIST && owner < 0; ++i) {
              const int o = wfb_ld_vol(&team_list[(i + bid) & (WFB_TEAM_LIST - 1)]) - 1;
              if (o < 0 || o == bid) continue;
              WfbTeamSlot* ts = team_slots + o;
              const int w = wfb_ld_vol(&ts->want);
              if (w <= 0 || wfb_ld_vol(&ts->quit) || wfb_ld_vol(&ts->members) >= w) continue;
              if (atomicAdd(&ts->members, 1) < w) owner = o;
              else atomicSub(&ts->members, 1);
            }
            if (owner >= 0) { held_slot = slot; S.help_owner = owner; slot = -2; break; }
          }
          __nanosleep(256);
          if (++spins > (1u << 26)) { *q.error = 2; slot = -1; break; } /* watchdog (~20 s) */
        }
        __threadfence();
#else
        if (q.ready[slot] == 0) slot = -1; /* serial emulation: an unpublished slot means the tree is exhausted */
#endif
      }
      S.slot = slot;
      WFB_PT_ADD(26, WFB_PT_CLOCK() - pt_idle0);
    }
    WFB_SYNC();
    const int slot = S.slot;
#ifndef WFB_EMU
    if (slot == -2) { /* lend a hand to a wide task until our own next task is published */
      const long long h0 = cta_log ? wfb_globaltimer() : 0;
      int* my_ready = nullptr;
      if (WFB_TID == 0) my_ready = &q.ready[held_slot];
      WfbAcc ta; ta.cells = ta.overlap = ta.matches = ta.steps = 0;
      wfb_team_help(team_slots + S.help_owner, S.brk.team, my_ready, q.outstanding, q.error, ta);
      acc.matches += ta.matches;
      if (cta_log) log_help += wfb_globaltimer() - h0;
      continue;
    }
#endif
    if (slot < 0) break;
    WfbTask t;
#ifndef WFB_EMU
    { /* read through L2: the task was published by another SM */
      const int* src = (const int*)&q.tasks[slot];
      int* dst = (int*)&t;
#pragma unroll
      for (int i = 0; i < (int)(sizeof(WfbTask) / sizeof(int)); ++i) dst[i] = __ldcg(src + i);
    }
#else
    t = q.tasks[slot];
#endif
    const long long pt_task0 = WFB_PT_CLOCK();
    const long long log_task0 = cta_log ? wfb_globaltimer() : 0;
    if (t.score_remaining <= WFB_FALLBACK_MIN_SCORE && t.pe > t.pb && t.te > t.tb) {
      n_base++;
      wfb_base_task(S.base, t, pairs, seq, arena, arena_stride, log, score_cap, runs, maxruns, pen, ops_all, pair_status, acc_base);
      if (WFB_TID == 0) { WFB_PT_ADD(20, WFB_PT_CLOCK() - pt_task0); WFB_PT_ADD(21, 1); }
    } else {
      n_break++;
#ifndef WFB_EMU
      wfb_break_task(S.brk, t, slot, pairs, seq, ws, W, pen, sink, ops_all, pair_status, acc, nullptr, team_slots ? &tctx : nullptr, seq_smem, seq_smem_words,
                     pair_flags);
#else
      wfb_break_task(S.brk, t, slot, pairs, seq, ws, W, pen, sink, ops_all, pair_status, acc, nullptr);
#endif
      if (WFB_TID == 0) { WFB_PT_ADD(24, WFB_PT_CLOCK() - pt_task0); WFB_PT_ADD(25, 1); }
    }
    (void)pt_task0;
    WFB_SYNC();
    if (cta_log) {
      const long long now = wfb_globaltimer();
      log_busy += now - log_task0;
      if (WFB_TID == 0) { /* per pair: CTA time spent on its tasks, time its last task ended (after the per-CTA block) */
        unsigned long long* pl = (unsigned long long*)cta_log + 4 * nblocks + 2 * (long long)t.pair;
#ifndef WFB_EMU
        atomicAdd(pl, (unsigned long long)(now - log_task0));
        atomicMax(pl + 1, (unsigned long long)now);
#else
        (void)pl;
#endif
      }
    }
    if (WFB_TID == 0) {
#ifndef WFB_EMU
      __threadfence();
#endif
      wfb_atomic_add(q.outstanding, -1);
    }
  }
  if (cta_log && WFB_TID == 0) {
    cta_log[4 * bid + 0] = log_busy; cta_log[4 * bid + 1] = wfb_globaltimer(); cta_log[4 * bid + 2] = log_t0;
#ifndef WFB_EMU
    cta_log[4 * bid + 3] = log_help;
#else
    cta_log[4 * bid + 3] = 0;
#endif
  }
  {
    unsigned long long m = acc.matches, mb = acc_base.matches;
#ifndef WFB_EMU
    for (int o = 16; o > 0; o >>= 1) { m += __shfl_down_sync(0xffffffffu, m, o); mb += __shfl_down_sync(0xffffffffu, mb, o); }
#endif
    if (wfb_lane() == 0 && m) wfb_atomic_add64(&counters->extend_matches, m);
    if (wfb_lane() == 0 && mb) wfb_atomic_add64(&counters->base_extend_matches, mb);
    if (wfb_lane() == 0 && acc.overlap) wfb_atomic_add64(&counters->overlap_tests, acc.overlap);
    if (WFB_TID == 0) {
      wfb_atomic_add64(&counters->cells, acc.cells);
      wfb_atomic_add64(&counters->score_steps, acc.steps);
      wfb_atomic_add64(&counters->break_tasks, n_break);
      wfb_atomic_add64(&counters->base_cells, acc_base.cells);
      wfb_atomic_add64(&counters->base_score_steps, acc_base.steps);
      wfb_atomic_add64(&counters->base_tasks, n_base);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Ends-free kernel: wfmash's head / tail patch alignments (wflign.cpp:280-305, 368-397):
 * WFAlignerGapAffine2Pieces(..., Alignment, MemoryMed).alignEndsFree(). `med` (piggy-backed backtrace)
 * and `high` memory produce the same transcript, so this is wavefront_unialign with the ends-free
 * initial wavefront (wavefront_aligner.c:252-310), wavefront_extend_endsfree's termination
 * (wavefront_extend.c:259-293, wavefront_termination.c:115-160) and wavefront_backtrace_affine.
 * term_group selects which terminating cell wins (1 = scalar build, 8 = AVX2, 16 = AVX-512 of the
 * reference: wavefront_extend_kernels.c:166-193 vs wavefront_extend_kernels_avx.c:296-400,592-691).
 * ---------------------------------------------------------------------------------------------- */
struct WfbEndsFree {
  int pbf, pef, tbf, tef;
};

WFB_DEV bool wfb_term_endsfree(int k, int off, int plen, int tlen, int pef, int tef) {
  const int h = off, v = off - k;
  if (h >= tlen && plen - v <= pef) return true;
  if (v >= plen && tlen - h <= tef) return true;
  return false;
}

struct WfbEfShared {
  WfbRing ring;
  int red_maxak[3];
  int red_end[3];
  int task_idx;
  int term_key; /* (class << 26 | j), INT_MAX = none */
  int nruns, bt_err;
};

/* search the trimmed M row of `slot` for the terminating cell the reference would pick; all threads */
WFB_DEV void wfb_ef_search(WfbEfShared& sh, const int32_t* arena, const unsigned char* runflag, int runbias, int slot, int plen, int tlen,
                           const WfbEndsFree& ef, int G) {
  const int lo = sh.ring.lo[slot][WFB_M], hi = sh.ring.hi[slot][WFB_M];
  if (!sh.ring.ex[slot][WFB_M] || lo > hi) return;
  const int n = hi - lo + 1;
  const int peel = (G > 1) ? (n < G ? n : n % G) : n;
  const int32_t* m = arena + sh.ring.boff[slot][WFB_M];
  int best = INT_MAX;
  for (int kk = lo + WFB_TID; kk <= hi; kk += WFB_NT) {
    const int32_t off = m[kk];
    if (off < 0) continue;
    if (!wfb_term_endsfree(kk, off, plen, tlen, ef.pef, ef.tef)) continue;
    const int j = kk - lo;
    const int cls = (G <= 1 || j < peel) ? 0 : (runflag[kk + runbias] ? 1 : 2);
    best = min(best, (cls << 26) | j);
  }
  best = wfb_warp_min(best);
  if (wfb_lane() == 0 && best != INT_MAX) wfb_smem_min(&sh.term_key, best);
}

WFB_KERNEL(wfb_endsfree_kernel, const WfbTask* tasks, const WfbEndsFree* efs, int ntasks, int* task_counter, const WfbPairDesc* pairs,
           const uint8_t* seq, int32_t* arena_all, long long arena_stride, WfbBaseMeta* log_all, int score_cap, WfbRun* runs_all,
           int maxruns, unsigned char* runflag_all, int runflag_stride, int term_group, WfbPen pen, char* ops_all, int* pair_status) {
  WFB_KERNEL_PROLOGUE
  WFB_SHARED WfbEfShared sh;
  int32_t* const arena = arena_all + (long long)bid * arena_stride;
  WfbBaseMeta* const log = log_all + (long long)bid * (score_cap + 1) * 5;
  WfbRun* const runs = runs_all + (long long)bid * maxruns;
  unsigned char* const runflag = runflag_all + (long long)bid * runflag_stride;
  WfbAcc acc;
  acc.cells = acc.overlap = acc.matches = acc.steps = 0;
  for (;;) {
    WFB_SYNC();
    if (WFB_TID == 0) sh.task_idx = wfb_atomic_add(task_counter, 1);
    WFB_SYNC();
    const int ti = sh.task_idx;
    if (ti >= ntasks) break;
    const WfbTask t = tasks[ti];
    const WfbEndsFree ef = efs[ti];
    const WfbPairDesc pd = pairs[t.pair];
    char* const ops = ops_all + pd.ops_off;
    const int plen = t.pe - t.pb, tlen = t.te - t.tb;
    const uint8_t* pf = seq + pd.p_off + t.pb;
    const uint8_t* tf = seq + pd.t_off + t.tb;
    const int R = pen.R;
    const int runbias = plen + 2;
    if (plen + tlen + 8 > runflag_stride || (long long)(ef.pbf + ef.tbf + 1) + 16 > arena_stride) {
      if (WFB_TID == 0) pair_status[t.pair] = WFB_PAIR_BASE_SCORE_CAP;
      continue;
    }
    wfb_ring_reset(sh.ring, R);
    if (WFB_TID == 0) { sh.red_maxak[0] = sh.red_maxak[1] = sh.red_maxak[2] = 0; sh.term_key = INT_MAX; }
    WFB_SYNC();
    /* score 0: M on diagonals [-pbf, tbf]: (h,0) -> offset h, (0,v) -> offset 0, each extended */
    const int lo0 = -ef.pbf, hi0 = ef.tbf, n0 = hi0 - lo0 + 1;
    for (int kk = lo0 + WFB_TID; kk <= hi0; kk += WFB_NT) {
      int off = kk > 0 ? kk : 0;
      const int v = off - kk;
      const int run = wfb_match_run(pf + v, tf + off, min(plen - v, tlen - off));
      off += run;
      arena[kk - lo0] = off;
      runflag[kk + runbias] = run >= 4 ? 1 : 0;
    }
    if (WFB_TID == 0) {
      sh.ring.ex[0][WFB_M] = 1; sh.ring.lo[0][WFB_M] = lo0; sh.ring.hi[0][WFB_M] = hi0; sh.ring.boff[0][WFB_M] = -lo0;
      for (int c = 0; c < 5; ++c) { WfbBaseMeta m; m.lo = lo0; m.hi = hi0; m.boff = -lo0; m.ex = (c == WFB_M); log[c] = m; }
    }
    WFB_SYNC();
    wfb_ef_search(sh, arena, runflag, runbias, 0, plen, tlen, ef, term_group);
    WFB_SYNC();
    WfbAllocBump ab;
    ab.runflag = runflag; ab.runbias = runbias;
    ab.bump = n0;
    int status = (sh.term_key != INT_MAX) ? WFB_ST_END_REACHED : WFB_ST_OK;
    int score = 0, num_null = 0, max_ak = 0;
    while (status == WFB_ST_OK) { /* wavefront_unialign.c:251-270 with wavefront_extend_endsfree */
      ++score;
      if (score > score_cap || (long long)ab.bump + 5LL * (plen + tlen + 8) > arena_stride) { status = -1; break; }
      status = wfb_step(sh.ring, arena, pen, score, pf, tf, plen, tlen, -1, num_null, ab, sh.red_maxak, sh.red_end, max_ak, acc);
      const int slot = score % R;
      if (WFB_TID == 0)
        for (int c = 0; c < 5; ++c) {
          WfbBaseMeta m;
          m.lo = sh.ring.lo[slot][c]; m.hi = sh.ring.hi[slot][c]; m.boff = sh.ring.boff[slot][c]; m.ex = sh.ring.ex[slot][c];
          log[score * 5 + c] = m;
        }
      if (status != WFB_ST_OK) break; /* END_UNREACHABLE */
      wfb_ef_search(sh, arena, runflag, runbias, slot, plen, tlen, ef, term_group);
      WFB_SYNC();
      if (sh.term_key != INT_MAX) status = WFB_ST_END_REACHED;
    }
    if (status != WFB_ST_END_REACHED) {
      if (WFB_TID == 0) pair_status[t.pair] = (status == -1) ? WFB_PAIR_BASE_SCORE_CAP : WFB_PAIR_UNATTAINABLE;
      continue;
    }
    if (WFB_TID == 0) {
      const int slot = score % R;
      const int end_k = sh.ring.lo[slot][WFB_M] + (sh.term_key & ((1 << 26) - 1));
      const int end_off = arena[sh.ring.boff[slot][WFB_M] + end_k];
      int nr = 0;
      sh.bt_err = wfb_backtrace(log, arena, score + 1, pen, WFB_M, WFB_M, plen, tlen, score, end_k, end_off, t.pb + t.tb, runs, maxruns, &nr);
      sh.nruns = nr;
    }
    WFB_SYNC();
    if (sh.bt_err) {
      if (WFB_TID == 0) pair_status[t.pair] = WFB_PAIR_BACKTRACE;
      continue;
    }
    const int nr = sh.nruns;
    for (int r = 0; r < nr; ++r) {
      const WfbRun ru = runs[r];
      const int stride = (ru.op == 'M' || ru.op == 'X') ? 2 : 1;
      for (int j = WFB_TID; j < ru.count; j += WFB_NT) ops[ru.idx + j * stride] = (char)ru.op;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Sequence staging: reversed copies (wavefront_sequences.c:83-101 does this per aligner on the CPU).
 * ---------------------------------------------------------------------------------------------- */
/* pair_flags (optional): set to 1 for pairs holding a byte other than A, C, G, T (they cannot use the 2-bit shared-memory windows) */
WFB_KERNEL(wfb_reverse_kernel, const WfbPairDesc* pairs, int npairs, uint8_t* seq, int* pair_flags) {
  WFB_KERNEL_PROLOGUE
  for (int i = bid; i < npairs; i += nblocks) {
    const WfbPairDesc pd = pairs[i];
    bool other = false;
    for (int j = WFB_TID; j < pd.plen; j += WFB_NT) {
      const uint8_t c = seq[pd.p_off + pd.plen - 1 - j];
      seq[pd.prev_off + j] = c;
      other |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
    }
    for (int j = WFB_TID; j < pd.tlen; j += WFB_NT) {
      const uint8_t c = seq[pd.t_off + pd.tlen - 1 - j];
      seq[pd.trev_off + j] = c;
      other |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
    }
    if (pair_flags && other) pair_flags[i] = 1; /* benign race: every writer stores 1 */
  }
}

/* Compaction of a pair's op slots (zero = hole) into a dense operation string, in place is not
 * possible in parallel, so dense output goes to ops_out + ops_off. One CTA per pair, chunked scan. */
WFB_KERNEL(wfb_compact_kernel, const WfbPairDesc* pairs, int npairs, const char* ops_slots, char* ops_out, int* ops_len) {
  WFB_KERNEL_PROLOGUE
  WFB_SHARED int sh_warp[32];
  WFB_SHARED int sh_base;
  for (int i = bid; i < npairs; i += nblocks) {
    const WfbPairDesc pd = pairs[i];
    const int n = pd.plen + pd.tlen;
    const char* src = ops_slots + pd.ops_off;
    char* dst = ops_out + pd.ops_off;
    WFB_SYNC();
    if (WFB_TID == 0) sh_base = 0;
    WFB_SYNC();
    for (int start = 0; start < n; start += WFB_NT * 8) {
      /* each thread owns 8 consecutive slots */
      const int b = start + WFB_TID * 8;
      char loc[8];
      int cnt = 0;
      for (int j = 0; j < 8; ++j) {
        const char ch = (b + j < n) ? src[b + j] : 0;
        if (ch) loc[cnt++] = ch;
      }
      int incl = cnt;
#ifndef WFB_EMU
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
        sh_warp[lane] = w; /* inclusive scan of warp totals */
      }
      WFB_SYNC();
      const int woff = wid ? sh_warp[wid - 1] : 0;
      const int total = sh_warp[((WFB_NT + 31) >> 5) - 1];
#else
      const int woff = 0;
      const int total = incl;
      (void)sh_warp;
#endif
      const int pos = sh_base + woff + incl - cnt;
      for (int j = 0; j < cnt; ++j) dst[pos + j] = loc[j];
      WFB_SYNC();
      if (WFB_TID == 0) sh_base += total;
      WFB_SYNC();
    }
    if (WFB_TID == 0) ops_len[i] = sh_base;
  }
}
