// l2_kernels.h — L2 sliding-window Jaccard stage on the GPU-resident index (sm_100a).
//
// Replaces, for fragments of length == windowLength (windowLen == 0; every fragment on the CLI path):
//   SlideMapper                          src/map/include/slidingMap.hpp:28-215
//   MappingCore::computeL2MappedRegions  src/map/include/mappingCore.hpp:306-442
//   the locus selection of Map::doL2Mapping (stage-1 top-ANI test, identity test)
//                                        src/map/include/computeMap.hpp:989-1061
//
// B200 mapping: ONE WARP PER L1 LOCUS, persistent grid striding over the locus list the L1 kernel left in HBM
// (no host round trip between L1 and L2).
//   * std::lower_bound over minmerIndex (23 dependent probes for 7 M minmers) becomes a 33-ary search: 32 lanes probe
//     32 splitters per round, a ballot picks the sub-range: 5 rounds.
//   * minmerIndex is streamed 32 records (1 KB, 128-bit loads, fully coalesced) per trip into a per-warp staging
//     buffer in shared memory; every lane then reads the same record (broadcast).
//   * SlideMapper's sorted vector of query minmers lives in shared memory; its lower_bound is a rank count: lanes
//     compare the reference hash with their share of the <= s query hashes and a warp add gives the slot.
//   * the reference's min-heap of open reference intervals (only used to find the intervals that END before the
//     window start) becomes an unordered array in shared memory that all lanes scan at once (ballot = expired set,
//     ballot + popc = in-place compaction); a warp-wide minimum of the end positions skips the scan on the steps
//     in which nothing expires. Only intervals whose hash is <= the largest query hash are kept (the others never
//     touch SlideMapper's state, slidingMap.hpp:137-140).
//   * pivot / rank / vote bookkeeping is replayed exactly, as warp-uniform registers (every lane computes the same
//     scalars from the same shared-memory reads; lane 0 does the writes).
// The state after each window step is a function of the SET of open intervals, so the order in which the intervals
// expiring in one step are deleted (heap order in the reference, array order here) does not matter; the oracle test
// checks that no hash is ever open twice (the one situation where SlideMapper is not a pure function of that set).
#pragma once
#include "index_kernels.h"

#ifndef WFB_EMU
#define L2_W 32
WFB_DEV unsigned l2_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
WFB_DEV unsigned l2_lt_mask() { return (1u << (threadIdx.x & 31)) - 1u; }
WFB_DEV int l2_shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
WFB_DEV unsigned long long l2_shfl64(unsigned long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }
WFB_DEV long long l2_warp_min64(long long v) {
  for (int o = 16; o; o >>= 1) { const long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
  return v;
}
#define L2_SYNCW() __syncwarp()
#else
#define L2_W 1
WFB_DEV unsigned l2_ballot(bool p) { return p ? 1u : 0u; }
WFB_DEV unsigned l2_lt_mask() { return 0u; }
WFB_DEV int l2_shfl(int v, int) { return v; }
WFB_DEV unsigned long long l2_shfl64(unsigned long long v, int) { return v; }
WFB_DEV long long l2_warp_min64(long long v) { return v; }
#define L2_SYNCW() ((void)0)
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
#endif

#define L2_INF 0x7fffffffffffffffLL
#define L2_STAGE 32 /* index records staged per trip */

struct L2Params {
  int k, w, s;       /* kmerSize, windowLength, param.sketchSize */
  int vec_cap;       /* L2_mapLocus_t per locus the per-warp slab holds */
  int warps_per_cta;
};

struct L2Entry { /* L2_mapLocus_t, mappingCore.hpp:33-41 */
  long long optimalStart, optimalEnd;
  int seqId, sharedSketchSize, strand, pad_;
};

/* bytes of shared memory one warp needs */
#ifndef WFB_EMU
__host__ __device__
#endif
static inline size_t l2_warp_smem(int s) {
  const size_t S1 = (size_t)s + 1;
  size_t b = 0;
  b += sizeof(wfb_minmer_t) * L2_STAGE;   /* stage  */
  b += 8 * S1;                            /* qh     */
  b += 8 * (S1 + 1);                      /* a_end  */
  b += 4 * S1;                            /* nbi    */
  b += 2 * S1;                            /* vote   */
  b += 2 * (S1 + 1);                      /* a_loc  */
  b += 1 * S1;                            /* qstr   */
  b += 1 * S1;                            /* act    */
  return (b + 15) & ~(size_t)15;
}

struct L2Warp { /* per-warp shared-memory views + the warp-uniform SlideMapper scalars */
  wfb_minmer_t* stage;
  unsigned long long* qh;   /* [1..qn] */
  long long* a_end;         /* open intervals: end position */
  int* nbi;                 /* num_before_inc, [0..qn] */
  short* vote;              /* strand_vote */
  unsigned short* a_loc;    /* open intervals: slot | match << 15 */
  signed char* qstr;        /* q_strand */
  unsigned char* act;       /* active */
  int qn, pivot, pivRank, shared, votes, isect, na;
  int open_cap;             /* = param.sketchSize: a window sketch cannot hold more open intervals */
  long long min_end;
};

WFB_DEV void l2_carve(L2Warp& W, unsigned char* base, int s) {
  const size_t S1 = (size_t)s + 1;
  unsigned char* p = base;
  W.stage = (wfb_minmer_t*)p; p += sizeof(wfb_minmer_t) * L2_STAGE;
  W.qh = (unsigned long long*)p; p += 8 * S1;
  W.a_end = (long long*)p; p += 8 * (S1 + 1);
  W.nbi = (int*)p; p += 4 * S1;
  W.vote = (short*)p; p += 2 * S1;
  W.a_loc = (unsigned short*)p; p += 2 * (S1 + 1);
  W.qstr = (signed char*)p; p += S1;
  W.act = (unsigned char*)p;
}

/* SlideMapper::SlideMapper + init(), slidingMap.hpp:84-125 */
WFB_DEV void l2_init(L2Warp& W, const wfb_minmer_t* q, int qn, int lane) {
  L2_SYNCW();
  for (int j = lane; j <= qn; j += L2_W) {
    if (j == 0) { W.qh[0] = 0; W.nbi[0] = 0; W.vote[0] = 0; W.qstr[0] = 0; W.act[0] = 0; }
    else { W.qh[j] = q[j - 1].hash; W.nbi[j] = 1; W.vote[j] = 0; W.qstr[j] = (signed char)q[j - 1].strand; W.act[j] = 0; }
  }
  W.qn = qn; W.pivot = qn; W.pivRank = qn; W.shared = 0; W.votes = 0; W.isect = 0; W.na = 0; W.min_end = L2_INF;
  L2_SYNCW();
}

/* std::lower_bound over slots [1, qn]: 1 + #{query hashes < h}; qn + 1 = end() */
WFB_DEV int l2_slot(const L2Warp& W, unsigned long long h, int lane) {
  unsigned c = 0;
  for (int j = 1 + lane; j <= W.qn; j += L2_W) c += (W.qh[j] < h) ? 1u : 0u;
  return 1 + (int)wfb_warp_add(c);
}

/* insert_minmer, slidingMap.hpp:129-170. Returns slot | match << 15, or 0 when the hash is beyond the last slot. */
WFB_DEV unsigned l2_insert(L2Warp& W, unsigned long long h, int strand, int lane) {
  const int loc = l2_slot(W, h, lane);
  if (loc > W.qn) return 0;
  const bool match = W.qh[loc] == h;
  const bool le_pivot = loc <= W.pivot; /* hash_val <= pivot->hash_val: the slots are strictly ascending */
  if (match) {
    const int v = (short)(W.vote[loc] + W.qstr[loc] * strand);
    W.isect++;
    if (le_pivot) { W.shared++; W.votes += v; }
    L2_SYNCW();
    if (lane == 0) { W.act[loc] = 1; W.vote[loc] = (short)v; }
    L2_SYNCW();
  } else {
    const int n1 = W.nbi[loc] + 1;
    if (le_pivot) W.pivRank++;
    if (W.pivRank > W.qn) {
      W.shared -= W.act[W.pivot];
      W.votes -= W.vote[W.pivot];
      W.pivRank -= (W.pivot == loc) ? n1 : W.nbi[W.pivot];
      W.pivot--;
    }
    L2_SYNCW();
    if (lane == 0) W.nbi[loc] = n1;
    L2_SYNCW();
  }
  return (unsigned)loc | (match ? 0x8000u : 0u);
}

/* delete_minmer, slidingMap.hpp:176-214, for an interval recorded by l2_insert */
WFB_DEV void l2_delete(L2Warp& W, unsigned rec, int lane) {
  const int loc = (int)(rec & 0x7fffu);
  const bool le_pivot = loc <= W.pivot;
  if (rec & 0x8000u) {
    if (le_pivot) { W.shared--; W.votes -= W.vote[loc]; }
    W.isect--;
    L2_SYNCW();
    if (lane == 0) { W.act[loc] = 0; W.vote[loc] = 0; }
    L2_SYNCW();
  } else {
    const int n1 = W.nbi[loc] - 1;
    if (le_pivot) W.pivRank--;
    if (W.pivot + 1 <= W.qn) {
      const int nn = (W.pivot + 1 == loc) ? n1 : W.nbi[W.pivot + 1];
      if (W.pivRank + nn <= W.qn) {
        W.pivot++;
        W.shared += W.act[W.pivot];
        W.votes += W.vote[W.pivot];
        W.pivRank += nn;
      }
    }
    L2_SYNCW();
    if (lane == 0) W.nbi[loc] = n1;
    L2_SYNCW();
  }
}

/* delete every open interval with end <= wpos (the reference's heap loop, mappingCore.hpp:363-373) and compact */
WFB_DEV void l2_expire(L2Warp& W, long long wpos, int lane) {
  if (wpos < W.min_end) return;
  int kept = 0;
  long long mn = L2_INF;
  const int na = W.na;
  for (int base = 0; base < na; base += L2_W) {
    const int j = base + lane;
    const bool valid = j < na;
    const long long e = valid ? W.a_end[j] : L2_INF;
    const unsigned rec = valid ? W.a_loc[j] : 0u;
    const bool expired = valid && e <= wpos;
    unsigned mexp = l2_ballot(expired);
    const unsigned mkeep = l2_ballot(valid && !expired);
    while (mexp) {
      const int src = __ffs((int)mexp) - 1;
      mexp &= mexp - 1;
      l2_delete(W, (unsigned)l2_shfl((int)rec, src), lane);
    }
    L2_SYNCW();
    if (valid && !expired) {
      const int d = kept + __popc(mkeep & l2_lt_mask());
      W.a_end[d] = e;
      W.a_loc[d] = (unsigned short)rec;
      mn = e < mn ? e : mn;
    }
    kept += __popc(mkeep);
    L2_SYNCW();
  }
  W.na = kept;
  W.min_end = l2_warp_min64(mn);
}

WFB_DEV void l2_open(L2Warp& W, unsigned rec, long long wpos_end, int lane, int& err) {
  if (!rec) return; /* beyond the last query hash: never touches the state */
  if (W.na > W.open_cap) { err = 1; return; } /* more open intervals than a window sketch can hold: inconsistent index */
  if (lane == 0) { W.a_end[W.na] = wpos_end; W.a_loc[W.na] = (unsigned short)rec; }
  W.na++;
  if (wpos_end < W.min_end) W.min_end = wpos_end;
  L2_SYNCW();
}

/* first index i with (seqId_i, wpos_i) >= (seq_id, key): 33-ary search, all lanes return the same value */
WFB_DEV long long l2_lower_bound(const wfb_minmer_t* index, long long n, int seq_id, long long key, int lane) {
  long long lo = 0, hi = n;
  while (hi > lo) {
    const long long step = (hi - lo) / (L2_W + 1) + 1;
    const long long idx = lo + step * (lane + 1) - 1;
    bool less = false;
    if (idx < hi) {
      const int sid = index[idx].seqId;
      less = sid < seq_id || (sid == seq_id && index[idx].wpos < key);
    }
    const int c = __popc(l2_ballot(less));
    const long long nlo = lo + step * c;
    long long nhi = lo + step * (c + 1) - 1;
    if (nhi > hi) nhi = hi;
    lo = nlo; hi = nhi;
  }
  return lo;
}

/* push_back-or-merge of a closed candidate (mappingCore.hpp:404-416, 427-440). `back` is l2_vec_out.back() held in
 * registers; it is flushed to the slab when another entry is pushed. */
WFB_DEV void l2_close(L2Entry& back, int& nvec, L2Entry* slab, int vec_cap, const L2Entry& cur, int seq_id, int strand_votes, int w, int lane, int& err) {
  if (nvec == 0 || back.optimalEnd + w < cur.optimalStart) {
    if (nvec > 0) {
      if (nvec - 1 < vec_cap) { if (lane == 0) slab[nvec - 1] = back; } else err = 1;
    }
    back = cur;
    back.seqId = seq_id;
    back.strand = strand_votes >= 0 ? 1 : -1;
    nvec++;
  } else {
    back.optimalEnd = cur.optimalEnd;
  }
}

struct L2Counters { unsigned long long loci, steps, skipped_stage1, out_entries; };

/* One warp per L1 locus. loci / locus_frag / *n_loci_ptr are what ix_l1_kernel left in device memory. */
WFB_KERNEL(l2_kernel, const wfb_minmer_t* index, long long n_index, const IxL1Locus* loci, const int* locus_frag,
           const unsigned long long* n_loci_ptr, long long loci_cap, const wfb_minmer_t* q_all, const int* q_count, L2Params P,
           const int* stage1_min_hits, const int* l2_min_shared, L2Entry* slab_all, wfb_l2_mapping_t* out,
           unsigned long long* out_counter, long long out_cap, int* frag_status, L2Counters* counters
#ifdef WFB_EMU
           , unsigned char* smem_emu
#endif
) {
  WFB_KERNEL_PROLOGUE
#ifndef WFB_EMU
  extern __shared__ __align__(16) unsigned char l2_smem_raw[];
  unsigned char* smem = l2_smem_raw;
  const int lane = wfb_lane();
  const int warp_in_cta = WFB_TID >> 5;
#else
  unsigned char* smem = smem_emu;
  const int lane = 0;
  const int warp_in_cta = 0;
#endif
  L2Warp W;
  l2_carve(W, smem + (size_t)warp_in_cta * l2_warp_smem(P.s), P.s);
  W.open_cap = P.s;
  const long long gwarp = (long long)bid * P.warps_per_cta + warp_in_cta;
  const long long nwarps = (long long)nblocks * P.warps_per_cta;
  L2Entry* slab = slab_all + gwarp * P.vec_cap;
  long long n_loci = (long long)*n_loci_ptr;
  if (n_loci > loci_cap) n_loci = loci_cap;
  unsigned long long c_loci = 0, c_steps = 0, c_skip = 0, c_out = 0;
  for (long long li = gwarp; li < n_loci; li += nwarps) {
    const IxL1Locus L = loci[li];
    const int f = locus_frag[li];
    if (f < 0) continue; /* a clipped list that ix_l1_kernel's redo pass replaced */
    const int qn = q_count[f];
    if (qn <= 0) continue;
    /* stage-1 top-ANI test (computeMap.hpp:999-1012): the heap pops the best locus first and stops at the first one
     * below the cut-off, i.e. every locus at or above the cut-off is processed */
    if (stage1_min_hits && L.intersectionSize < stage1_min_hits[qn]) { c_skip++; continue; }
    c_loci++;
    l2_init(W, q_all + (size_t)f * P.s, qn, lane);
    int err = 0;
    long long it = l2_lower_bound(index, n_index, L.seqId, L.rangeStartPos - P.w - 1, lane);
    int bestSketchSize = 1, nvec = 0;
    bool in_candidate = false, done = false;
    L2Entry cur, back;
    cur.optimalStart = 0; cur.optimalEnd = 0; cur.seqId = 0; cur.sharedSketchSize = 0; cur.strand = 0; cur.pad_ = 0;
    back = cur;
    int last_seq = L.seqId;
    while (!done && it < n_index) {
      const long long left = n_index - it;
      const int cnt = left < L2_STAGE ? (int)left : L2_STAGE;
      L2_SYNCW();
      for (int j = lane; j < cnt; j += L2_W) {
#ifndef WFB_EMU
        const uint4* src = (const uint4*)(index + it + j);
        uint4* dst = (uint4*)(W.stage + j);
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        dst[0] = a; dst[1] = b;
#else
        W.stage[j] = index[it + j];
#endif
      }
      L2_SYNCW();
      for (int j = 0; j < cnt; ++j) {
        const wfb_minmer_t m = W.stage[j];
        if (m.seqId != L.seqId) { done = true; break; }
        if (m.wpos < L.rangeStartPos) { /* set up the window, mappingCore.hpp:339-355 */
          if (m.wpos_end > L.rangeStartPos) {
            const unsigned rec = l2_insert(W, m.hash, m.strand, lane);
            l2_open(W, rec, m.wpos_end, lane, err);
          }
          c_steps++;
          continue;
        }
        if (m.wpos > L.rangeEndPos) { done = true; break; }
        c_steps++;
        /* one window step, mappingCore.hpp:358-423 with windowLen == 0 */
        const int prev_strand_votes = W.votes;
        l2_expire(W, m.wpos, lane);
        const unsigned rec = l2_insert(W, m.hash, m.strand, lane);
        l2_open(W, rec, m.wpos_end, lane, err);
        last_seq = m.seqId;
        if (W.shared > bestSketchSize) {
          nvec = 0; /* l2_vec_out.clear() */
          in_candidate = true;
          bestSketchSize = W.shared;
          cur.sharedSketchSize = W.shared;
          cur.optimalStart = m.wpos;
          cur.optimalEnd = m.wpos;
        } else if (W.shared == bestSketchSize) {
          if (!in_candidate) { cur.sharedSketchSize = W.shared; cur.optimalStart = m.wpos; }
          in_candidate = true;
          cur.optimalEnd = m.wpos;
        } else {
          if (in_candidate) {
            l2_close(back, nvec, slab, P.vec_cap, cur, m.seqId, prev_strand_votes, P.w, lane, err);
            cur.optimalStart = 0; cur.optimalEnd = 0; cur.sharedSketchSize = 0;
          }
          in_candidate = false;
        }
      }
      it += cnt;
    }
    if (in_candidate) l2_close(back, nvec, slab, P.vec_cap, cur, last_seq, W.votes, P.w, lane, err);
    if (nvec > 0) {
      if (nvec - 1 < P.vec_cap) { if (lane == 0) slab[nvec - 1] = back; } else err = 1;
    }
    L2_SYNCW();
#ifndef WFB_EMU
    __threadfence_block();
#endif
    if (err) { if (lane == 0) frag_status[f] = WFB_ECAP; continue; }
    /* identity test on sharedSketchSize (computeMap.hpp:1018-1024, table from the host) + append */
    /* every entry of l2_vec_out carries bestSketchSize (the vector is cleared whenever the best grows) */
    const int npass = (!l2_min_shared || bestSketchSize >= l2_min_shared[qn]) ? nvec : 0;
    if (npass == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd_compat(out_counter, (unsigned long long)npass);
    base = l2_shfl64(base, 0);
    c_out += (unsigned long long)npass;
    if ((long long)(base + npass) > out_cap) continue; /* the host sees the counter and reports WFB_ECAP */
    for (int i = lane; i < nvec; i += L2_W) {
      const L2Entry e = (i == nvec - 1) ? back : slab[i];
      wfb_l2_mapping_t r;
      r.frag = f; r.refSeqId = e.seqId;
      r.optimalStart = e.optimalStart; r.optimalEnd = e.optimalEnd;
      r.refStartPos = (e.optimalStart + e.optimalEnd) / 2; /* meanOptimalPos */
      r.conservedSketches = e.sharedSketchSize; r.strand = e.strand;
      r.nucIdentity = 0.f; r.kmerComplexity = 0.f; /* filled by the host part (libm pow, computeMap.hpp:1018-1019) */
      out[base + i] = r;
    }
  }
  if (lane == 0 && counters) {
    atomicAdd_compat(&counters->loci, c_loci);
    atomicAdd_compat(&counters->steps, c_steps);
    atomicAdd_compat(&counters->skipped_stage1, c_skip);
    atomicAdd_compat(&counters->out_entries, c_out);
  }
}

/* sort keys of the mappings: (frag, refSeqId, refStartPos) = mapSingleQueryFrag's final std::sort per fragment
 * (computeMap.hpp:919-920) with the fragments of the batch kept apart */
WFB_KERNEL(l2_keys_kernel, const wfb_l2_mapping_t* m, long long n, unsigned long long* key_pos, unsigned int* key_frag, unsigned int* idx) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    key_pos[i] = ix_pack(m[i].refSeqId, m[i].refStartPos, 0);
    key_frag[i] = (unsigned int)m[i].frag;
    idx[i] = (unsigned int)i;
  }
}
WFB_KERNEL(l2_gather_u32_kernel, const unsigned int* src, const unsigned int* idx, long long n, unsigned int* dst) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) dst[i] = src[idx[i]];
}
WFB_KERNEL(l2_permute_kernel, const wfb_l2_mapping_t* src, const unsigned int* idx, long long n, wfb_l2_mapping_t* dst) {
  WFB_KERNEL_PROLOGUE
  for (long long i = (long long)bid * WFB_NT + WFB_TID; i < n; i += (long long)nblocks * WFB_NT) {
    dst[i] = src[idx[i]];
  }
}
