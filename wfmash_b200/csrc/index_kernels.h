// index_kernels.h — GPU-resident reference minmer index + query-side L1 (sm_100a).
//
// Replaces, for fragments of length == windowLength (every fragment on the CLI path):
//   Sketch::build index part          src/map/include/winSketch.hpp:351-429  (postings; see index_host.cu)
//   MappingCore::getSeedIntervalPoints src/map/include/mappingCore.hpp:81-131
//   MappingCore::computeL1CandidateRegions :136-301, driven per PanSN group slice by
//   Map::doL1Mapping                   src/map/include/computeMap.hpp:945-983
//
// B200 mapping:
//   * minmerPosLookupIndex (hash -> vector<IntervalPoint>) becomes an open-addressing table of 32-slot
//     buckets (512 B, one coalesced warp read per probe; lanes compare 32 keys at once and a ballot picks
//     the hit / detects an empty slot) pointing into one CSR array of interval points pre-packed as sortable
//     64-bit keys (seqId | pos | side);
//   * one CTA per query fragment: sketch (sketch_kernels.h) -> warp-cooperative probes -> gather of the
//     posting lists into shared memory (global scratch for the rare oversized fragment) -> bitonic sort =
//     the reference's k-way heap merge order -> the two L1 sweeps.
#pragma once
#include "sketch_kernels.h"

#define IX_EMPTY 0xFFFFFFFFFFFFFFFFULL
#define IX_BUCKET 32

struct IxSlot {
  uint64_t key;
  uint32_t start, count;
};

/* interval point packed as a sort key: seqId (22 bits) | pos (40 bits) | side (1 bit: CLOSE=0 < OPEN=1) */
WFB_DEV uint64_t ix_pack(int seq_id, long long pos, int open) {
  return ((uint64_t)(uint32_t)seq_id << 41) | ((uint64_t)pos << 1) | (uint64_t)(open ? 1 : 0);
}
WFB_DEV int ix_seq(uint64_t k) { return (int)(k >> 41); }
WFB_DEV long long ix_pos(uint64_t k) { return (long long)((k >> 1) & ((1ULL << 40) - 1)); }
WFB_DEV int ix_open(uint64_t k) { return (int)(k & 1); }

WFB_DEV uint64_t ix_mix(uint64_t h) { /* minmer hashes are already Murmur3 outputs: fold high bits only */
  return h ^ (h >> 29);
}

/* insert one unique key (thread per key) */
WFB_KERNEL(ix_insert_kernel, IxSlot* table, long long nbuckets, const uint64_t* uhash, const uint32_t* ustart, const uint32_t* ucount,
           long long nuniq, int* fail) {
  WFB_KERNEL_PROLOGUE
  for (long long u = (long long)bid * WFB_NT + WFB_TID; u < nuniq; u += (long long)nblocks * WFB_NT) {
    const uint64_t h = uhash[u];
    if (h == IX_EMPTY) { *fail = 2; continue; } /* a hash equal to the empty marker (p = 2^-64) is handled on the host */
    long long b = (long long)(ix_mix(h) & (uint64_t)(nbuckets - 1));
    bool done = false;
    for (long long tries = 0; tries < nbuckets && !done; ++tries) {
      IxSlot* bk = table + b * IX_BUCKET;
      for (int j = 0; j < IX_BUCKET && !done; ++j) {
#ifndef WFB_EMU
        const unsigned long long old = atomicCAS((unsigned long long*)&bk[j].key, (unsigned long long)IX_EMPTY, (unsigned long long)h);
#else
        const uint64_t old = bk[j].key;
        if (old == IX_EMPTY) bk[j].key = h;
#endif
        if (old == IX_EMPTY) { bk[j].start = ustart[u]; bk[j].count = ucount[u]; done = true; }
      }
      b = (b + 1) & (nbuckets - 1);
    }
    if (!done) *fail = 1;
  }
}

/* warp-cooperative lookup: returns (start,count) through refs; all lanes get the same answer */
WFB_DEV bool ix_lookup_warp(const IxSlot* table, long long nbuckets, uint64_t h, uint32_t& start, uint32_t& count) {
  long long b = (long long)(ix_mix(h) & (uint64_t)(nbuckets - 1));
#ifndef WFB_EMU
  const int lane = wfb_lane();
  for (long long tries = 0; tries < nbuckets; ++tries) {
    const IxSlot sl = table[b * IX_BUCKET + lane];
    const unsigned hit = __ballot_sync(0xffffffffu, sl.key == h);
    if (hit) {
      const int src = __ffs((int)hit) - 1;
      start = __shfl_sync(0xffffffffu, sl.start, src);
      count = __shfl_sync(0xffffffffu, sl.count, src);
      return true;
    }
    if (__ballot_sync(0xffffffffu, sl.key == IX_EMPTY)) return false;
    b = (b + 1) & (nbuckets - 1);
  }
  return false;
#else
  for (long long tries = 0; tries < nbuckets; ++tries) {
    for (int j = 0; j < IX_BUCKET; ++j) {
      const IxSlot sl = table[b * IX_BUCKET + j];
      if (sl.key == h) { start = sl.start; count = sl.count; return true; }
      if (sl.key == IX_EMPTY) return false;
    }
    b = (b + 1) & (nbuckets - 1);
  }
  return false;
#endif
}

struct IxL1Params {
  int k, w, s;          /* kmerSize, windowLength, sketchSize (param.sketchSize) */
  int minimum_hits;     /* cached_minimum_hits, computeMap.hpp:160,959-961 */
  int skip_self, skip_prefix, lower_triangular;
  int ncut;             /* sketchCutoffs.size() */
  int smem_cap;         /* interval points that fit the shared-memory buffer (power of two) */
  int gcap;             /* ... the per-CTA global scratch (power of two) */
  int max_loci;         /* per fragment */
  int par_sweep;        /* 1 = data-parallel L1 sweep (ix_l1_regions_par), 0 = thread 0 walks (kept for A/B runs: WFB_L1_SERIAL=1) */
  float complexity_threshold;
};

struct IxL1Locus { /* L1_candidateLocus_t, mappingCore.hpp:24-30 */
  int seqId, intersectionSize;
  long long rangeStartPos, rangeEndPos;
};

/* the two sweeps of computeL1CandidateRegions over sorted keys [0,n) (thread 0) */
WFB_DEV void ix_l1_regions(const uint64_t* ip, int n, int minimumHits, int q_sketch, const IxL1Params& P, const int* cutoffs,
                           IxL1Locus* out, int& nout, IxL1Locus* local, int local_cap, int& err) {
  if (n == 0) return;
  int overlap = 0, best = 0;
  int tr = 0, ld = 0;
  while (ld < n) { /* pass 1, :160-187 (windowLen == 0) */
    const uint64_t lead = ip[ld];
    /* trailing: same seq && pos <= lead.pos, or smaller seq  <=>  key with side forced to OPEN <= lead|1 */
    const uint64_t lim = lead | 1ULL;
    while (tr < n && ip[tr] <= lim) { if (!ix_open(ip[tr])) overlap--; tr++; }
    const long long cur = ix_pos(lead);
    while (ld < n && ix_pos(ip[ld]) == cur) { if (ix_open(ip[ld])) overlap++; ld++; }
    if (overlap > best) best = overlap;
  }
  if (best < minimumHits) return;
  {
    const double div = P.s / 1000.0 > 1.0 ? P.s / 1000.0 : 1.0; /* skch::fixed::ss_table_max, :193-198 */
    int idx = (int)((best < q_sketch ? best : q_sketch) / div);
    if (idx >= P.ncut) idx = P.ncut - 1;
    if (cutoffs[idx] > minimumHits) minimumHits = cutoffs[idx];
  }
  /* pass 2, :203-284 */
  bool in_cand = false;
  IxL1Locus cur_out;
  cur_out.seqId = 0; cur_out.intersectionSize = 0; cur_out.rangeStartPos = 0; cur_out.rangeEndPos = 0;
  int nlocal = 0;
  tr = 0; ld = 0; overlap = 0;
  int prev_seq = 0; long long prev_pos = 0;
  int cur_seq = ix_seq(ip[0]); long long cur_pos = ix_pos(ip[0]);
  while (ld < n) {
    const int prevOverlap = overlap;
    const uint64_t lim = ip[ld] | 1ULL;
    while (tr < n && ip[tr] <= lim) { if (!ix_open(ip[tr])) overlap--; tr++; }
    if (ix_pos(ip[ld]) != cur_pos) { prev_seq = cur_seq; prev_pos = cur_pos; cur_seq = ix_seq(ip[ld]); cur_pos = ix_pos(ip[ld]); }
    while (ld < n && ix_pos(ip[ld]) == cur_pos) { if (ix_open(ip[ld])) overlap++; ld++; }
    if (prevOverlap >= minimumHits) {
      if (cur_out.seqId != prev_seq && in_cand) {
        if (nlocal < local_cap) local[nlocal++] = cur_out; else err = 1;
        cur_out.seqId = 0; cur_out.intersectionSize = 0; cur_out.rangeStartPos = 0; cur_out.rangeEndPos = 0;
        in_cand = false;
      }
      if (!in_cand) {
        cur_out.rangeStartPos = prev_pos; cur_out.rangeEndPos = prev_pos; cur_out.seqId = prev_seq; cur_out.intersectionSize = prevOverlap;
        in_cand = true;
      } else { /* stage2_full_scan, :263-266 */
        if (prevOverlap > cur_out.intersectionSize) cur_out.intersectionSize = prevOverlap;
        cur_out.rangeEndPos = prev_pos;
      }
    } else {
      if (in_cand) {
        if (nlocal < local_cap) local[nlocal++] = cur_out; else err = 1;
        cur_out.seqId = 0; cur_out.intersectionSize = 0; cur_out.rangeStartPos = 0; cur_out.rangeEndPos = 0;
      }
      in_cand = false;
    }
  }
  if (in_cand) { if (nlocal < local_cap) local[nlocal++] = cur_out; else err = 1; }
  for (int i = 0; i < nlocal; ++i) { /* join, :287-300 */
    if (nout == 0 || local[i].seqId != out[nout - 1].seqId || local[i].rangeStartPos > out[nout - 1].rangeEndPos + P.w) {
      if (nout < P.max_loci) out[nout++] = local[i]; else err = 1;
    } else {
      out[nout - 1].rangeEndPos = local[i].rangeEndPos;
      if (local[i].intersectionSize > out[nout - 1].intersectionSize) out[nout - 1].intersectionSize = local[i].intersectionSize;
    }
  }
}

/* ------------------------------------------------------------------------------------------------------------------
 * The same two sweeps (+ the join, :287-300) for ALL group slices of a fragment at once, data-parallel over the CTA.
 * The serial walk above is a chain of ~3 n dependent shared-memory reads on one thread while the other warps wait at a
 * barrier (65 % of the L1 kernel's time, profiles/r01_ncu_map_kernels_summary.txt); here every quantity it carries is a
 * prefix sum, a segment id or a segmented maximum over the sorted points:
 *   slice    = maximal run of points whose sequences share a PanSN group (computeMap.hpp:964-982; one slice without -Y)
 *   step     = maximal run of consecutive points with equal pos (POSITION ONLY, like the reference's leading loop, :177
 *              / :238: a step may spill across a sequence boundary)
 *   spc run  = maximal run of points with equal (seqId, pos): where the trailing pointer stops (ip[tr] <= lead | 1)
 *   overlap after step j = opens before the step's end - closes before the end of the spc run its first point starts
 *   candidate run = consecutive steps (never a slice's last one: the reference looks at a step's count one iteration
 *              later) with overlap >= the slice's minimumHits on one sequence -> one local locus; loci closer than w on the
 *              same sequence join (the rule only looks at the previous local locus).
 * One 64-bit block scan carries five 12-bit counters (n <= IX_PAR_CAP = 2048 points): opens | closes | steps | spc runs |
 * slices. Larger fragments and fragments with more than IX_PAR_SLICES slices take the serial walk (same results). */
#define IX_PAR_CAP 2048
#define IX_PAR_SLICES 256
#define IX_F_OPEN(x) ((int)((x) & 0xFFF))
#define IX_F_CLOSE(x) ((int)(((x) >> 12) & 0xFFF))
#define IX_F_STEP(x) ((int)(((x) >> 24) & 0xFFF))
#define IX_F_SPC(x) ((int)(((x) >> 36) & 0xFFF))
#define IX_F_SLICE(x) ((int)(((x) >> 48) & 0xFFF))
#define IX_ID_ONE ((1ULL << 24) | (1ULL << 36) | (1ULL << 48))

struct IxParAux { /* views into the CTA's dynamic shared memory behind the IX_PAR_CAP keys */
  unsigned long long* pre;   /* [n + 1] per point: opens / closes BEFORE it, ids of its step / spc run / slice; [n] = totals */
  unsigned short* stepfirst; /* [nsteps + 1] first point of a step; [nsteps] = n */
  unsigned short* spcfirst;  /* [nspc + 1] */
  short* ov;                 /* [nsteps] overlap after the step */
};
#ifndef WFB_EMU
__host__ __device__
#endif
static inline size_t ix_par_aux_bytes() { return (size_t)(IX_PAR_CAP + 2) * 8 + (size_t)(IX_PAR_CAP + 8) * 2 * 3; }
WFB_DEV IxParAux ix_par_carve(unsigned char* p) {
  IxParAux a;
  a.pre = (unsigned long long*)p; p += (size_t)(IX_PAR_CAP + 2) * 8;
  a.stepfirst = (unsigned short*)p; p += (size_t)(IX_PAR_CAP + 8) * 2;
  a.spcfirst = (unsigned short*)p; p += (size_t)(IX_PAR_CAP + 8) * 2;
  a.ov = (short*)p;
  return a;
}

/* exclusive scan of one 64-bit value per thread over the CTA (two barriers); total = the CTA-wide sum */
WFB_DEV unsigned long long ix_cta_exscan(unsigned long long v, unsigned long long* sh, unsigned long long& total) {
#ifndef WFB_EMU
  const int lane = WFB_TID & 31, wid = WFB_TID >> 5, nw = (WFB_NT + 31) >> 5;
  unsigned long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[wid] = x;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
  for (int w = 0; w < nw; ++w) { const unsigned long long t = sh[w]; if (w < wid) base += t; tot += t; }
  __syncthreads();
  total = tot;
  return base + x - v;
#else
  (void)sh;
  total = v;
  return 0;
#endif
}

struct IxParShared { /* static shared memory of the parallel sweep */
  unsigned long long scan[33];
  int slbest[IX_PAR_SLICES], slmh[IX_PAR_SLICES], slstep[IX_PAR_SLICES]; /* per slice: best overlap, minimumHits, first step */
  int nout, err;
};

/* All threads of the CTA call this with the same arguments. buf[0..n) sorted keys in shared memory (n <= IX_PAR_CAP).
 * Returns false (uniform, nothing written) when the fragment has more slices than IX_PAR_SLICES; otherwise the joined
 * loci are in out[0 .. S.nout) and S.err tells whether a capacity was exceeded. local: scratch of local_cap loci. */
WFB_DEV bool ix_l1_regions_par(const uint64_t* buf, int n, int q_sketch, const IxL1Params& P, const int* cutoffs, const int* ref_group,
                               const IxParAux& A, IxParShared& S, IxL1Locus* out, IxL1Locus* local, int local_cap) {
  const int tid = WFB_TID, nt = WFB_NT;
  for (int i = tid; i < IX_PAR_SLICES; i += nt) S.slbest[i] = 0;
  if (tid == 0) { S.nout = 0; S.err = 0; }
  /* P1: flags -> one packed scan */
  unsigned long long carry = 0;
  for (int base = 0; base < n; base += nt) {
    const int i = base + tid;
    unsigned long long v = 0;
    bool sb = false, st = false, spc = false;
    if (i < n) {
      const uint64_t key = buf[i], prev = i > 0 ? buf[i - 1] : 0ULL;
      sb = i == 0 || (P.skip_prefix && ref_group[ix_seq(key)] != ref_group[ix_seq(prev)]);
      st = sb || ix_pos(key) != ix_pos(prev);
      spc = sb || (key >> 1) != (prev >> 1);
      v = (ix_open(key) ? 1ULL : (1ULL << 12)) | (st ? (1ULL << 24) : 0ULL) | (spc ? (1ULL << 36) : 0ULL) | (sb ? (1ULL << 48) : 0ULL);
    }
    unsigned long long total;
    const unsigned long long ex = ix_cta_exscan(v, S.scan, total) + carry;
    if (i < n) {
      const unsigned long long incl = ex + v; /* every id field of incl is >= 1: point 0 carries all three flags */
      const unsigned long long rec = (ex & 0xFFFFFFULL) | ((incl - IX_ID_ONE) & ~0xFFFFFFULL);
      A.pre[i] = rec;
      if (st) A.stepfirst[IX_F_STEP(rec)] = (unsigned short)i;
      if (spc) A.spcfirst[IX_F_SPC(rec)] = (unsigned short)i;
      if (sb && IX_F_SLICE(rec) < IX_PAR_SLICES) S.slstep[IX_F_SLICE(rec)] = IX_F_STEP(rec);
    }
    carry += total;
  }
  const int nsteps = IX_F_STEP(carry), nspc = IX_F_SPC(carry), nslices = IX_F_SLICE(carry);
  if (nslices > IX_PAR_SLICES) { WFB_SYNC(); return false; }
  if (tid == 0) { A.pre[n] = carry; A.stepfirst[nsteps] = (unsigned short)n; A.spcfirst[nspc] = (unsigned short)n; }
  WFB_SYNC();
  /* P2: overlap after every step (pass 1, :160-187) + the slice's best */
  for (int j = tid; j < nsteps; j += nt) {
    const int f = A.stepfirst[j], l = A.stepfirst[j + 1];
    const unsigned long long pf = A.pre[f];
    const int tr = A.spcfirst[IX_F_SPC(pf) + 1];
    const int sl = IX_F_SLICE(pf);
    const unsigned long long p0 = A.pre[A.stepfirst[S.slstep[sl]]];
    const int ov = (IX_F_OPEN(A.pre[l]) - IX_F_OPEN(p0)) - (IX_F_CLOSE(A.pre[tr]) - IX_F_CLOSE(p0));
    A.ov[j] = (short)ov;
    wfb_smem_max(&S.slbest[sl], ov);
  }
  WFB_SYNC();
  /* P3: minimumHits of every slice (:189-198) */
  for (int sl = tid; sl < nslices; sl += nt) {
    const int best = S.slbest[sl];
    int mh = INT_MAX; /* best < minimumHits: the slice returns early */
    if (best >= P.minimum_hits) {
      mh = P.minimum_hits;
      const double div = P.s / 1000.0 > 1.0 ? P.s / 1000.0 : 1.0;
      int idx = (int)((best < q_sketch ? best : q_sketch) / div);
      if (idx >= P.ncut) idx = P.ncut - 1;
      if (cutoffs[idx] > mh) mh = cutoffs[idx];
    }
    S.slmh[sl] = mh;
  }
  WFB_SYNC();
  /* P4: candidate runs (pass 2, :203-284) -> local[] in order */
#define IX_STEP_Q(J, SL) ((J) + 1 < nsteps && IX_F_SLICE(A.pre[A.stepfirst[(J) + 1]]) == (SL) && (int)A.ov[(J)] >= S.slmh[(SL)])
  int nlocal = 0;
  for (int base = 0; base < nsteps; base += nt) {
    const int j = base + tid;
    bool rs = false;
    int sl = 0, f = 0;
    if (j < nsteps) {
      f = A.stepfirst[j];
      sl = IX_F_SLICE(A.pre[f]);
      if (IX_STEP_Q(j, sl)) {
        rs = j == S.slstep[sl] || !((int)A.ov[j - 1] >= S.slmh[sl]) || ix_seq(buf[f]) != ix_seq(buf[A.stepfirst[j - 1]]);
      }
    }
    unsigned long long total;
    const int r = nlocal + (int)ix_cta_exscan(rs ? 1ULL : 0ULL, S.scan, total);
    if (rs) {
      IxL1Locus c;
      c.seqId = ix_seq(buf[f]); c.rangeStartPos = ix_pos(buf[f]); c.rangeEndPos = c.rangeStartPos; c.intersectionSize = A.ov[j];
      for (int t = j + 1; IX_STEP_Q(t, sl) && ix_seq(buf[A.stepfirst[t]]) == c.seqId; ++t) {
        c.rangeEndPos = ix_pos(buf[A.stepfirst[t]]);
        if ((int)A.ov[t] > c.intersectionSize) c.intersectionSize = A.ov[t];
      }
      if (r < local_cap) local[r] = c; else S.err = 1;
    }
    nlocal += (int)total;
  }
#undef IX_STEP_Q
  if (nlocal > local_cap) nlocal = local_cap; /* S.err is set */
#ifndef WFB_EMU
  __threadfence_block();
#endif
  WFB_SYNC();
  /* P5: join (:287-300) */
  int nout = 0;
  for (int base = 0; base < nlocal; base += nt) {
    const int r = base + tid;
    bool ng = false;
    IxL1Locus c;
    if (r < nlocal) {
      c = local[r];
      ng = r == 0 || c.seqId != local[r - 1].seqId || c.rangeStartPos > local[r - 1].rangeEndPos + P.w;
    }
    unsigned long long total;
    const int g = nout + (int)ix_cta_exscan(ng ? 1ULL : 0ULL, S.scan, total);
    if (ng) {
      for (int t = r + 1; t < nlocal; ++t) {
        const IxL1Locus d = local[t];
        if (d.seqId != local[t - 1].seqId || d.rangeStartPos > local[t - 1].rangeEndPos + P.w) break;
        c.rangeEndPos = d.rangeEndPos;
        if (d.intersectionSize > c.intersectionSize) c.intersectionSize = d.intersectionSize;
      }
      if (g < P.max_loci) out[g] = c; else S.err = 1;
    }
    nout += (int)total;
  }
  if (tid == 0) S.nout = nout < P.max_loci ? nout : P.max_loci;
#ifndef WFB_EMU
  __threadfence_block();
#endif
  WFB_SYNC();
  return true;
}

struct IxFragQuery { /* per-fragment query metadata (QueryMetaData, base_types.hpp:336-349) */
  int q_seq_id, q_group;
};

/* dynamic smem: [sketch buffers: npow2*12 + seq] ... reused afterwards as the interval-point buffer */
WFB_KERNEL(ix_l1_kernel, const uint8_t* seq_base, const wfb_frag_t* frags, const IxFragQuery* fq, int nfrags, int npow2_max,
           IxL1Params P, const IxSlot* table, long long nbuckets, const uint64_t* points, const int* ref_group, const int* cutoffs,
           wfb_minmer_t* q_out, int* q_count, float* q_complexity, unsigned long long* q_maxhash, uint64_t* gscratch_all, IxL1Locus* loci_tmp_all,
           IxL1Locus* loci_out, int* loci_frag, unsigned long long* loci_counter, long long loci_cap, long long* frag_loci_off, int* frag_loci_n,
           int* frag_status, const int* frag_list /* optional: only these fragments (the redo pass with larger per-fragment scratch) */, int n_list
#ifdef WFB_EMU
           , unsigned char* smem_emu
#endif
) {
  WFB_KERNEL_PROLOGUE
#ifndef WFB_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
#else
  unsigned char* smem = smem_emu;
#endif
  WFB_SHARED int sh_warp[32];
  WFB_SHARED uint32_t sh_start[512], sh_cnt[512], sh_off[513];
  WFB_SHARED int sh_total, sh_nloci, sh_err;
  WFB_SHARED unsigned long long sh_base;
  WFB_SHARED IxParShared sh_par;
  uint64_t* gscratch = gscratch_all + (long long)bid * P.gcap;
  IxL1Locus* ltmp = loci_tmp_all + (long long)bid * 2 * P.max_loci; /* [0,max) = out list, [max,2max) = local */
  const int n_todo = frag_list ? n_list : nfrags;
  for (int fi = bid; fi < n_todo; fi += nblocks) {
    const int f = frag_list ? frag_list[fi] : fi;
    WFB_SYNC();
    if (frag_list) { /* whatever the first pass left of this fragment (a clipped list) is dropped: the L2 kernel skips loci of fragment -1 */
      const long long o0 = frag_loci_off[f];
      const int on = frag_loci_n[f];
      for (int i = WFB_TID; i < on; i += WFB_NT) loci_frag[o0 + i] = -1;
      WFB_SYNC();
    }
    const wfb_frag_t fr = frags[f];
    wfb_minmer_t* qo = q_out + (size_t)f * P.s;
    const int qn = sk_sketch_block(smem, sh_warp, seq_base, fr, P.k, P.s, npow2_max, qo);
    WFB_SYNC(); /* qo[] visible, smem reusable */
    if (WFB_TID == 0) {
      q_count[f] = qn;
      sh_nloci = 0; sh_err = 0; sh_total = 0;
      float kc = 0.f;
      if (qn > 0) { /* mappingCore.hpp:72-74. The reference divides in x87 long double; this double-precision value is only used
                       for the kmerComplexityThreshold test below (0 on the CLI path). The value handed back to the caller is
                       recomputed on the host in long double from q_maxhash (index_host.cu ix_kmer_complexity). */
        const double max_hash_01 = (double)qo[qn - 1].hash / 18446744073709551615.0;
        kc = (float)(((double)qn / max_hash_01) / ((double)(fr.len - P.k + 1) * 2));
      }
      q_maxhash[f] = qn > 0 ? qo[qn - 1].hash : 0ULL;
      q_complexity[f] = kc;
      frag_loci_off[f] = 0; frag_loci_n[f] = 0; frag_status[f] = 0;
    }
    WFB_SYNC();
    if (qn == 0 || q_complexity[f] < P.complexity_threshold) continue; /* computeMap.hpp:951-953 */
    /* probes: one warp per query hash */
    {
#ifndef WFB_EMU
      const int nwarps = WFB_NT >> 5, warp_id = WFB_TID >> 5;
#else
      const int nwarps = 1, warp_id = 0;
#endif
      for (int i = warp_id; i < qn; i += nwarps) {
        uint32_t st = 0, ct = 0;
        const bool found = ix_lookup_warp(table, nbuckets, qo[i].hash, st, ct);
        if (wfb_lane() == 0) { sh_start[i] = st; sh_cnt[i] = found ? ct : 0; }
      }
    }
    WFB_SYNC();
    if (WFB_TID == 0) {
      uint32_t acc = 0;
      for (int i = 0; i < qn; ++i) { sh_off[i] = acc; acc += sh_cnt[i]; }
      sh_off[qn] = acc;
      sh_total = (int)acc;
    }
    WFB_SYNC();
    const int total = sh_total;
    if (total == 0) continue;
    int N = 1;
    while (N < total) N <<= 1;
    uint64_t* buf;
    if (N <= P.smem_cap) buf = (uint64_t*)smem;
    else if (N <= P.gcap) buf = gscratch;
    else { if (WFB_TID == 0) frag_status[f] = WFB_ECAP; continue; }
    /* gather + group filters of getSeedIntervalPoints (:110-119); dropped points become +inf keys */
    const IxFragQuery q = fq[f];
    for (int i = 0; i < qn; ++i) {
      const uint32_t st = sh_start[i], ct = sh_cnt[i], off = sh_off[i];
      for (uint32_t t = WFB_TID; t < ct; t += WFB_NT) {
        uint64_t key = points[st + t];
        const int sid = ix_seq(key);
        const int tg = ref_group[sid];
        bool skip = false;
        if (P.skip_self && q.q_group == tg) skip = true;
        if (P.skip_prefix && q.q_group == tg) skip = true;
        if (P.lower_triangular && q.q_seq_id <= sid) skip = true;
        buf[off + t] = skip ? IX_EMPTY : key;
      }
    }
    for (int i = total + WFB_TID; i < N; i += WFB_NT) buf[i] = IX_EMPTY;
    WFB_SYNC();
    for (int kk = 2; kk <= N; kk <<= 1) { /* bitonic sort = heap-merge order by (seqId, pos, side) */
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int t = WFB_TID; t < (N >> 1); t += WFB_NT) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int p = i | j;
          const bool up = (i & kk) == 0;
          const uint64_t a = buf[i], b = buf[p];
          if ((a > b) == up) { buf[i] = b; buf[p] = a; }
        }
        WFB_SYNC();
      }
    }
    /* points that survived the group filters (the dropped ones sorted to the end as +inf keys) */
    if (WFB_TID == 0) sh_total = 0;
    WFB_SYNC();
    for (int i = WFB_TID; i < total; i += WFB_NT)
      if (buf[i] != IX_EMPTY && (i + 1 == total || buf[i + 1] == IX_EMPTY)) sh_total = i + 1;
    WFB_SYNC();
    const int n = sh_total;
    bool swept = false;
    if (P.par_sweep && buf == (uint64_t*)smem && N <= IX_PAR_CAP) {
      const IxParAux A = ix_par_carve(smem + (size_t)IX_PAR_CAP * 8);
      swept = ix_l1_regions_par(buf, n, qn, P, cutoffs, ref_group, A, sh_par, ltmp, ltmp + P.max_loci, P.max_loci);
      if (swept && WFB_TID == 0) {
        sh_nloci = sh_par.nout; sh_err = sh_par.err;
        if (sh_par.nout > 0) sh_base = atomicAdd_compat(loci_counter, (unsigned long long)sh_par.nout);
      }
    }
    if (!swept && WFB_TID == 0) { /* oversized fragment / too many group slices: the serial walk */
      int nout = 0, err = 0;
      int b = 0;
      while (b < n) { /* per PanSN group slice, computeMap.hpp:964-982 */
        int e = n;
        if (P.skip_prefix) { e = b; const int g = ref_group[ix_seq(buf[b])]; while (e < n && ref_group[ix_seq(buf[e])] == g) ++e; }
        ix_l1_regions(buf + b, e - b, P.minimum_hits, qn, P, cutoffs, ltmp, nout, ltmp + P.max_loci, P.max_loci, err);
        b = e;
      }
      sh_nloci = nout; sh_err = err;
      if (nout > 0) sh_base = atomicAdd_compat(loci_counter, (unsigned long long)nout);
    }
    WFB_SYNC();
    const int nl = sh_nloci;
    if (nl > 0) {
      const unsigned long long base = sh_base;
      if ((long long)(base + nl) <= loci_cap) {
        for (int i = WFB_TID; i < nl; i += WFB_NT) { loci_out[base + i] = ltmp[i]; loci_frag[base + i] = f; }
        if (WFB_TID == 0) { frag_loci_off[f] = (long long)base; frag_loci_n[f] = nl; }
      } else if (WFB_TID == 0) frag_status[f] = WFB_ECAP;
    }
    if (WFB_TID == 0 && sh_err) frag_status[f] = WFB_ECAP;
  }
}

/* ---- index construction kernels (winSketch.hpp:266-429) ---- */
WFB_KERNEL(ix_hash_keys_kernel, const wfb_minmer_t* mi, long long n, unsigned long long* keys, int* idx) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) { keys[i] = mi[i].hash; idx[i] = (int)i; }
}
/* after the stable sort by hash: flags of hash-run heads */
WFB_KERNEL(ix_head_flags_kernel, const unsigned long long* skeys, long long n, int* head) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) head[i] = (i == 0 || skeys[i] != skeys[i - 1]) ? 1 : 0;
}
/* run id (inclusive scan of head - 1) -> frequency per run */
WFB_KERNEL(ix_run_freq_kernel, const int* head, const long long* runid_incl, long long n, unsigned long long* ufreq, const unsigned long long* skeys,
           unsigned long long* uhash_all) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    const long long r = runid_incl[i] - 1;
    atomicAdd_compat(&ufreq[r], 1ULL);
    if (head[i]) uhash_all[r] = skeys[i];
  }
}
/* keep flag per sorted position + per original index (freq > threshold && freq > min_occ are dropped, :374-377) */
WFB_KERNEL(ix_keep_kernel, const long long* runid_incl, const unsigned long long* ufreq, const int* sidx, long long n,
           unsigned long long threshold, int* keep_sorted, int* keep_orig) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    const unsigned long long fq = ufreq[runid_incl[i] - 1];
    const int kp = !(fq > threshold && fq > 10ULL);
    keep_sorted[i] = kp;
    keep_orig[sidx[i]] = kp;
  }
}
/* postings run starts in hash-sorted order (:379-387): position i starts a new OPEN/CLOSE pair unless it
 * abuts the previous kept minmer of the same hash inside the same worker partition */
WFB_KERNEL(ix_pair_start_kernel, const wfb_minmer_t* mi, const int* sidx, const unsigned long long* skeys, const int* keep_sorted,
           const int* part_of_seq, long long n, int* pstart) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    int st = 0;
    if (keep_sorted[i]) {
      st = 1;
      if (i > 0 && skeys[i - 1] == skeys[i]) { /* same hash => previous is kept too (same frequency) */
        const wfb_minmer_t a = mi[sidx[i - 1]], b = mi[sidx[i]];
        if (part_of_seq[a.seqId] == part_of_seq[b.seqId] && a.wpos_end == b.wpos) st = 0;
      }
    }
    pstart[i] = st;
  }
}
/* write packed points: pair p (= inclusive scan of pstart - 1) gets OPEN at 2p from its first minmer and CLOSE
 * at 2p+1 from its last (the seqId of a merged CLOSE stays that of the first minmer, :383-386) */
WFB_KERNEL(ix_points_kernel, const wfb_minmer_t* mi, const int* sidx, const int* keep_sorted, const int* pstart,
           const long long* pair_incl, long long n, uint64_t* points) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    if (!keep_sorted[i]) continue;
    const long long p = pair_incl[i] - 1;
    const wfb_minmer_t m = mi[sidx[i]];
    if (pstart[i]) points[2 * p] = ix_pack(m.seqId, m.wpos, 1);
    const bool last = (i + 1 >= n) || !keep_sorted[i + 1] || pstart[i + 1];
    if (last) {
      /* seqId of the pair = its first minmer's: walk back to the pair start (runs are short) */
      long long j = i;
      while (!pstart[j]) --j;
      points[2 * p + 1] = ix_pack(mi[sidx[j]].seqId, m.wpos_end, 0);
    }
  }
}
/* per kept unique hash: first point and number of points */
WFB_KERNEL(ix_uniq_kernel, const int* head, const int* keep_sorted, const int* pstart, const long long* pair_incl, const long long* ukept_incl,
           const unsigned long long* skeys, long long n, unsigned long long* uhash, uint32_t* ustart, uint32_t* ucount_pairs) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    if (!keep_sorted[i]) continue;
    const long long u = ukept_incl[i] - 1; /* index among kept unique hashes (scan of head&&keep) */
    if (head[i]) { uhash[u] = skeys[i]; ustart[u] = (uint32_t)(2 * (pair_incl[i] - 1)); }
    if (pstart[i]) wfb_atomic_add((int*)&ucount_pairs[u], 2);
  }
}
WFB_KERNEL(ix_and_kernel, const int* a, const int* b, long long n, int* out) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) out[i] = a[i] && b[i];
}
WFB_KERNEL(ix_compact_minmers_kernel, const wfb_minmer_t* mi, const int* keep_orig, const long long* off, long long n, wfb_minmer_t* out) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT)
    if (keep_orig[i]) out[off[i]] = mi[i];
}
WFB_KERNEL(ix_fill_kernel, unsigned long long* p, long long n, unsigned long long v) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) p[i] = v;
}
