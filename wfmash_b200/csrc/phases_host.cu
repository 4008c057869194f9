// phases_host.cu — the mapping phase and the alignment phase of `wfmash` as one C-ABI call each over sequences in host
// memory: host C++ orchestration of the library's own entry points, in the order src/interface/main.cpp,
// skch::Map (src/map/include/computeMap.hpp:300-860) and align::Aligner (src/align/include/computeAlignments.hpp:318-720)
// run them. Nothing here computes a mapping or an alignment itself.
#include "wfmash_b200.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */

void wfb_trace_mark_(const char* tag); /* wfa_host.cu */

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* skch::SequenceIdManager (sequenceIds.hpp:284-441): ids in first-seen order, targets before queries; group = the name up to
 * its last delimiter (the whole name without one), numbered from 1 in sorted-name order */
struct Ids {
  std::vector<std::string> names;
  std::vector<int64_t> lens;
  std::unordered_map<std::string, int32_t> id_of;
  std::vector<int32_t> group;
  void add(const wfb_seq_t* s, int32_t n) {
    for (int32_t i = 0; i < n; ++i) {
      const std::string nm(s[i].name);
      if (id_of.count(nm)) continue;
      id_of[nm] = (int32_t)names.size();
      names.push_back(nm);
      lens.push_back(s[i].len);
    }
  }
  void build_groups(int delim) {
    std::vector<std::pair<std::string, int32_t>> sorted;
    for (size_t i = 0; i < names.size(); ++i) sorted.emplace_back(names[i], (int32_t)i);
    std::sort(sorted.begin(), sorted.end());
    std::unordered_map<std::string, int32_t> keys;
    group.assign(names.size(), 0);
    int32_t next = 0;
    for (auto& [nm, idx] : sorted) {
      std::string key = nm;
      if (delim > 0) { const size_t pos = nm.rfind((char)delim); if (pos != std::string::npos) key = nm.substr(0, pos); }
      auto it = keys.find(key);
      if (it == keys.end()) it = keys.emplace(key, ++next).first;
      group[(size_t)idx] = it->second;
    }
  }
};

char* to_c_text(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  if (!p) return nullptr;
  memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  return p;
}

}  // namespace

extern "C" void wfb_free_text(char* text) { free(text); }

/* The order in which a ONE-worker taskflow executor (`wfmash -t 1`) runs the fragment tasks of every query = the order in which their
 * results reach Map::filterSubsetMappings, which decides the ch:Z: tags (computeMap.hpp:532-640). n_frag[q] = fragment tasks of the
 * q-th entry of querySequenceNames (0 for sequences shorter than the segment length: they still get a query task). Restates the
 * executor's scheduling (src/common/taskflow/core/executor.hpp:1271-1310,1452-1493; tsq.hpp:436-484 with
 * TF_DEFAULT_BOUNDED_TASK_QUEUE_LOG_SIZE 8 => 255 usable slots): scheduled nodes go to the worker's bounded LIFO queue, overflow to the
 * unbounded FIFO free list; a worker waiting in Subflow::join keeps popping its own queue (which may hold OTHER queries' tasks: they
 * run nested inside the join) and takes from the free list only when its queue is empty. order[q] = fragment indices of query q. */
static void one_thread_fragment_order(const std::vector<int64_t>& n_frag, std::vector<std::vector<int32_t>>& order) {
  const size_t kQueueCapacity = 255;
  struct Node { int32_t q, i; };
  order.assign(n_frag.size(), {});
  std::vector<int64_t> remaining(n_frag);
  std::vector<Node> stack, fifo;
  std::vector<int32_t> waiting;
  size_t head = 0;
  auto schedule = [&](Node n) { if (stack.size() < kQueueCapacity) stack.push_back(n); else fifo.push_back(n); };
  for (size_t q = 0; q < n_frag.size(); ++q) schedule(Node{(int32_t)q, -1});
  for (;;) {
    if (!waiting.empty() && remaining[(size_t)waiting.back()] == 0) { waiting.pop_back(); continue; }
    Node n;
    if (!stack.empty()) { n = stack.back(); stack.pop_back(); }
    else if (head < fifo.size()) n = fifo[head++];
    else break;
    if (n.i < 0) { /* a query task: emplaces its fragments, then joins */
      for (int64_t j = 0; j < n_frag[(size_t)n.q]; ++j) schedule(Node{n.q, (int32_t)j});
      waiting.push_back(n.q);
    } else {
      order[(size_t)n.q].push_back(n.i);
      --remaining[(size_t)n.q];
    }
  }
}

#ifdef WFB_EMU
/* TEST-ONLY (never in the product library): the index build and the mapping kernels are not part of the host emulation, so a test
 * injects the fragment mappings (the oracle's) and wfb_map_phase runs everything AROUND the device calls for real: ids, groups,
 * fragments, run-level constants, fragment order, boundary check, chain merge + filters, PAF text. */
static const wfb_l2_mapping_t* g_emu_l2 = nullptr;
static const int64_t* g_emu_off = nullptr;
static int64_t g_emu_nfrag = -1;
extern "C" void wfb_emu_inject_l2(const wfb_l2_mapping_t* mappings, const int64_t* frag_map_offset, int64_t n_frags) {
  g_emu_l2 = mappings; g_emu_off = frag_map_offset; g_emu_nfrag = n_frags;
}
#endif

extern "C" int wfb_map_phase(int device, const wfb_map_phase_params_t* params, const wfb_seq_t* targets, int32_t n_targets, const wfb_seq_t* queries,
                             int32_t n_queries, char** paf, int64_t* paf_len, wfb_map_phase_stats_t* stats) {
  return wfb_map_phase_subset(device, params, targets, n_targets, queries, n_queries, nullptr, paf, paf_len, stats);
}

extern "C" int wfb_map_phase_subset(int device, const wfb_map_phase_params_t* params, const wfb_seq_t* targets, int32_t n_targets, const wfb_seq_t* queries,
                                    int32_t n_queries, const uint8_t* query_select, char** paf, int64_t* paf_len, wfb_map_phase_stats_t* stats) {
  if (params && query_select && params->filter.filter_mode == WFB_FILTER_ONETOONE) {
    wfb_set_last_error_("the one-to-one filter needs every query in one call");
    return WFB_EINVAL;
  }
  if (!params || !paf || !paf_len || n_targets <= 0 || n_queries < 0 || !targets || (n_queries > 0 && !queries) || params->kmer_size < 1 ||
      params->window_length <= 0) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  const double t_begin = now_s();
  wfb_trace_mark_("map_phase: begin");
  *paf = nullptr; *paf_len = 0;
  wfb_map_phase_params_t P = *params;
  const int k = P.kmer_size;
  const int64_t w = P.window_length;
  Ids ids;
  ids.add(targets, n_targets);
  ids.add(queries, n_queries);
  ids.build_groups(P.skip_prefix ? P.prefix_delim : 0);
  int rc = WFB_OK;
  int32_t stale_absorbed = 0;
  double ani_seconds = 0, ani_kernel_ms = 0, index_kernel_ms = 0;

  if (!(P.percentage_identity > 0)) {
    const double t_ani = now_s(); /* main.cpp:75-134: ANI auto-identity, then the sketch size follows the estimate */
    std::vector<int32_t> gid(ids.group.begin(), ids.group.end());
    std::sort(gid.begin(), gid.end());
    gid.erase(std::unique(gid.begin(), gid.end()), gid.end());
    std::map<int32_t, int32_t> dense;
    for (size_t i = 0; i < gid.size(); ++i) dense[gid[i]] = (int32_t)i;
    const int32_t ng = (int32_t)gid.size();
    auto role = [&](const wfb_seq_t* s, int32_t n, std::vector<uint64_t>& sk, std::vector<int32_t>& cnt, std::vector<int32_t>& present_gid) -> int {
      std::vector<const char*> ptr((size_t)n); std::vector<int64_t> len((size_t)n); std::vector<int32_t> g((size_t)n);
      std::vector<char> has((size_t)ng, 0);
      for (int32_t i = 0; i < n; ++i) { ptr[(size_t)i] = s[i].seq; len[(size_t)i] = s[i].len; g[(size_t)i] = dense[ids.group[(size_t)ids.id_of[s[i].name]]]; has[(size_t)g[(size_t)i]] = 1; }
      std::vector<uint64_t> all((size_t)ng * 4096); std::vector<int32_t> c((size_t)ng);
      wfb_ani_stats_t as; memset(&as, 0, sizeof as);
      const int r = wfb_ani_group_sketches(device, ptr.data(), len.data(), g.data(), n, ng, 21, 4096, all.data(), c.data(), &as);
      if (r != WFB_OK) return r;
      ani_kernel_ms += as.hash_kernel_ms + as.sort_kernel_ms;
      for (int32_t j = 0; j < ng; ++j)
        if (has[(size_t)j]) { sk.insert(sk.end(), all.begin() + (size_t)j * 4096, all.begin() + (size_t)(j + 1) * 4096); cnt.push_back(c[(size_t)j]); present_gid.push_back(gid[(size_t)j]); }
      return WFB_OK;
    };
    std::vector<uint64_t> qs, ts; std::vector<int32_t> qc, tc, qg, tg;
    if ((rc = role(queries, n_queries, qs, qc, qg)) != WFB_OK) return rc;
    if ((rc = role(targets, n_targets, ts, tc, tg)) != WFB_OK) return rc;
    P.percentage_identity = (float)wfb_ani_estimate_identity(qs.data(), qc.data(), qg.data(), (int32_t)qg.size(), ts.data(), tc.data(), tg.data(),
                                                            (int32_t)tg.size(), 4096, 21, P.ani_percentile, P.ani_adjustment, nullptr);
    if (params->sketch_size <= 0) P.sketch_size = (int32_t)std::min<int64_t>(wfb_sketch_size(P.percentage_identity, w, k), w);
    ani_seconds = now_s() - t_ani;
    wfb_trace_mark_("map_phase: ids + ANI auto-identity");
  }
  if (P.sketch_size <= 0) P.sketch_size = wfb_sketch_size(P.percentage_identity, w, k);
  const int s = P.sketch_size;
  P.filter.window_length = w; P.filter.percentage_identity = P.percentage_identity; P.filter.skip_prefix = P.skip_prefix;

  /* index over the targets (build_index task, computeMap.hpp:472-484) */
  std::vector<const char*> tptr((size_t)n_targets); std::vector<int64_t> tlen((size_t)n_targets); std::vector<int32_t> tid((size_t)n_targets);
  for (int32_t i = 0; i < n_targets; ++i) { tptr[(size_t)i] = targets[i].seq; tlen[(size_t)i] = targets[i].len; tid[(size_t)i] = ids.id_of[targets[i].name]; }
  wfb_index_params_t ip; memset(&ip, 0, sizeof ip);
  ip.kmer_size = k; ip.window_size = (int32_t)w; ip.sketch_size = s; ip.index_threads = std::max(1, P.index_threads); ip.max_kmer_freq = P.max_kmer_freq;
  const double t_ix = now_s();
#ifndef WFB_EMU
  wfb_index_stats_t ixs; memset(&ixs, 0, sizeof ixs);
  wfb_index_t* ix = wfb_index_build(device, &ip, tptr.data(), tlen.data(), tid.data(), n_targets, &ixs);
  stale_absorbed = (int32_t)std::min<uint64_t>(ixs.minmer.stale_absorbed, 0x7fffffffu);
  index_kernel_ms = ixs.minmer.total_kernel_ms + ixs.index_kernel_ms;
  if (!ix) return WFB_ECUDA; /* message already set */
#else
  wfb_index_t* ix = nullptr;
  if (g_emu_nfrag < 0) { wfb_set_last_error_("the mapping kernels are not part of the host emulation (see wfb_emu_inject_l2)"); return WFB_ENODEV; }
#endif
  const double index_seconds = now_s() - t_ix;
  wfb_trace_mark_("map_phase: index build");

  /* run-level constants (computeMap.hpp:150-160, 224-226, 999-1024) */
  const int32_t min_hits = std::max(P.minimum_hits, wfb_estimate_minimum_hits_relaxed(s, k, P.percentage_identity, 0.95f));
  std::vector<int32_t> cutoffs((size_t)std::min(s, 1000) + 1), s1((size_t)s + 1), shared((size_t)s + 1);
  wfb_sketch_cutoffs(s, k, P.ani_diff, P.ani_diff_conf, P.stage1_top_ani_filter, cutoffs.data(), (int32_t)cutoffs.size());
  if (P.stage1_top_ani_filter) wfb_stage1_min_hits(P.hg_numerator, P.ani_diff, k, s, s1.data());
  if (P.keep_low_pct_id) wfb_l2_min_shared_relaxed(P.percentage_identity, k, s, 0.95f, shared.data());
  else wfb_l2_min_shared(P.percentage_identity, k, s, shared.data());

  /* query fragments (computeMap.hpp:560-630): floor(len / w) pieces + one overlapping tail piece; shorter queries map nowhere */
  std::vector<int32_t> mapped;
  std::vector<int64_t> base;
  int64_t blob_bytes = 0;
  for (int32_t q = 0; q < n_queries; ++q)
    if (queries[q].len >= w && (!query_select || query_select[q])) { mapped.push_back(q); base.push_back(blob_bytes); blob_bytes += queries[q].len; }
  std::string blob;
  blob.reserve((size_t)blob_bytes);
  for (int32_t q : mapped) blob.append(queries[q].seq, (size_t)queries[q].len);
  std::vector<wfb_frag_t> frags; std::vector<wfb_frag_query_t> fq; std::vector<int32_t> frag_index; std::vector<int64_t> q_frag(1, 0);
  for (size_t m = 0; m < mapped.size(); ++m) {
    const wfb_seq_t& Q = queries[mapped[m]];
    const int32_t qid = ids.id_of[Q.name];
    const int64_t nfull = Q.len / w;
    for (int64_t i = 0; i < nfull; ++i) { frags.push_back(wfb_frag_t{base[m] + i * w, (int32_t)w, qid}); fq.push_back(wfb_frag_query_t{qid, ids.group[(size_t)qid]}); frag_index.push_back((int32_t)i); }
    if (Q.len % w != 0) { frags.push_back(wfb_frag_t{base[m] + Q.len - w, (int32_t)w, qid}); fq.push_back(wfb_frag_query_t{qid, ids.group[(size_t)qid]}); frag_index.push_back((int32_t)nfull); }
    q_frag.push_back((int64_t)frags.size());
  }
  wfb_trace_mark_("map_phase: query blob + fragments");
  std::string text;
  int64_t n_l2 = 0, n_out = 0;
  double map_ms = 0, filter_seconds = 0;
  if (!frags.empty()) {
    const int32_t nf = (int32_t)frags.size();
    wfb_l1_params_t lp; memset(&lp, 0, sizeof lp);
    lp.minimum_hits = min_hits; lp.sketch_cutoffs = cutoffs.data(); lp.n_cutoffs = (int32_t)cutoffs.size(); lp.ref_group = ids.group.data();
    lp.n_ref_group = (int32_t)ids.group.size(); lp.skip_self = P.skip_self; lp.skip_prefix = P.skip_prefix; lp.lower_triangular = P.lower_triangular;
    wfb_l2_params_t l2p; memset(&l2p, 0, sizeof l2p);
    if (P.stage1_top_ani_filter) { l2p.stage1_min_hits = s1.data(); l2p.n_stage1_min_hits = (int32_t)s1.size(); }
    l2p.l2_min_shared = shared.data(); l2p.n_l2_min_shared = (int32_t)shared.size();
    std::vector<wfb_l2_mapping_t> maps; std::vector<int64_t> moff((size_t)nf + 1); std::vector<int32_t> fst((size_t)nf);
    int64_t cap = 32 * (int64_t)nf + 1024;
    for (int attempt = 0; attempt < 6; ++attempt) { /* grow the mapping buffer when the batch needs more */
      maps.resize((size_t)cap);
      wfb_map_out_t mo; memset(&mo, 0, sizeof mo);
      mo.mappings = maps.data(); mo.mappings_cap = cap; mo.frag_map_offset = moff.data(); mo.frag_status = fst.data();
#ifndef WFB_EMU
      rc = wfb_map_fragments_batch(ix, &lp, &l2p, blob.data(), (int64_t)blob.size(), frags.data(), fq.data(), nf, &mo);
#else
      if (g_emu_nfrag != nf) { wfb_set_last_error_("injected fragment count differs from the phase's fragments"); return WFB_EINVAL; }
      if (g_emu_off[nf] > cap) rc = WFB_ECAP;
      else {
        memcpy(maps.data(), g_emu_l2, sizeof(wfb_l2_mapping_t) * (size_t)g_emu_off[nf]);
        memcpy(moff.data(), g_emu_off, sizeof(int64_t) * ((size_t)nf + 1));
        mo.n_mappings = g_emu_off[nf];
        rc = WFB_OK;
      }
#endif
      if (rc == WFB_ECAP) { cap *= 4; continue; }
      if (rc == WFB_OK) { n_l2 = mo.n_mappings; map_ms = mo.l1_kernel_ms + mo.l2_kernel_ms + mo.sort_kernel_ms; }
      break;
    }
    if (rc == WFB_OK)
      for (int32_t f = 0; f < nf; ++f)
        if (fst[(size_t)f] != 0) { wfb_set_last_error_("a fragment exceeded an internal capacity of the mapping kernels"); rc = WFB_ECAP; break; }
    if (rc != WFB_OK) { if (ix) wfb_index_free(ix); return rc; }
    wfb_trace_mark_("map_phase: L1 + L2 (wfb_map_fragments_batch)");
    const double t_f = now_s();
    /* per query: MappingResult construction + boundary check, then the chain / filter stage for the whole batch */
    std::vector<wfb_mapping_t> all((size_t)n_l2 + 1); /* + 1: never a null pointer, also when no fragment mapped anywhere */
    std::vector<int64_t> q_off(1, 0), q_len;
    /* The order in which the fragments' results reach the chain merge decides the ch:Z: tags (chain ids rank the smallest ORIGINAL
     * index of each chain, mappingFilter.hpp:401-404,498-520). The reference appends them as its fragment tasks finish
     * (computeMap.hpp:590-597), i.e. in a schedule-dependent order; its only reproducible schedule is the one-thread run
     * (one_thread_fragment_order above). That order is used here, so the text equals `wfmash -m -t 1`. */
    std::vector<int64_t> tasks((size_t)n_queries, 0); /* every query of the RUN has a task in the reference's executor, also the ones another rank maps */
    for (int32_t q = 0; q < n_queries; ++q) tasks[(size_t)q] = queries[q].len >= w ? queries[q].len / w + (queries[q].len % w != 0 ? 1 : 0) : 0;
    std::vector<std::vector<int32_t>> frag_order;
    one_thread_fragment_order(tasks, frag_order);
    std::vector<wfb_l2_mapping_t> ordered;
    for (size_t m = 0; m < mapped.size(); ++m) {
      const int64_t a = moff[(size_t)q_frag[m]], b = moff[(size_t)q_frag[m + 1]];
      ordered.clear();
      for (int32_t fo : frag_order[(size_t)mapped[m]]) {
        const int64_t f = q_frag[m] + fo;
        for (int64_t i = moff[(size_t)f]; i < moff[(size_t)f + 1]; ++i) ordered.push_back(maps[(size_t)i]);
      }
      if ((rc = wfb_l2_to_query_mappings(ordered.data(), b - a, frag_index.data(), w, queries[mapped[m]].len, ids.lens.data(), all.data() + a)) != WFB_OK) break;
      q_off.push_back(b);
      q_len.push_back(queries[mapped[m]].len);
    }
    std::vector<wfb_mapping_t> out((size_t)n_l2 + 16); std::vector<wfb_chain_info_t> chain((size_t)n_l2 + 16); std::vector<int64_t> oo(mapped.size() + 1, 0);
    if (rc == WFB_OK)
      rc = wfb_filter_mappings_batch(&P.filter, all.data(), q_off.data(), q_len.data(), (int32_t)mapped.size(), ids.group.data(), ids.lens.data(), out.data(),
                                     chain.data(), (int64_t)out.size(), oo.data(), 0);
    bool default_chain = false;
    if (rc == WFB_OK && P.filter.filter_mode == WFB_FILTER_ONETOONE) { /* computeMap.hpp:788-850 */
      std::vector<wfb_mapping_t> kept(4 * (size_t)oo.back() + 16); std::vector<int32_t> owner(kept.size());
      const int64_t n = wfb_one_to_one_filter(&P.filter, out.data(), oo.data(), (int32_t)mapped.size(), ids.group.data(), ids.lens.data(), kept.data(), owner.data(),
                                              (int64_t)kept.size());
      if (n < 0) rc = (int)n;
      else {
        out.assign(kept.begin(), kept.begin() + n);
        std::fill(oo.begin(), oo.end(), 0);
        for (int64_t i = 0; i < n; ++i) oo[(size_t)owner[(size_t)i] + 1]++;
        for (size_t m = 0; m < mapped.size(); ++m) oo[m + 1] += oo[m];
        default_chain = true;
      }
    }
    if (rc != WFB_OK) { if (ix) wfb_index_free(ix); return rc; }
    std::vector<const char*> names(ids.names.size());
    for (size_t i = 0; i < names.size(); ++i) names[i] = ids.names[i].c_str();
    std::vector<char> buf(1 << 16);
    for (size_t m = 0; m < mapped.size(); ++m) {
      const int64_t n = oo[m + 1] - oo[m];
      int64_t need = 0;
      int64_t got = wfb_mapping_paf_format(&P.filter, out.data() + oo[m], default_chain ? nullptr : chain.data() + oo[m], n, queries[mapped[m]].name, queries[mapped[m]].len,
                                           names.data(), ids.lens.data(), buf.data(), (int64_t)buf.size(), &need);
      if (got == WFB_ECAP) {
        buf.resize((size_t)need + 64);
        got = wfb_mapping_paf_format(&P.filter, out.data() + oo[m], default_chain ? nullptr : chain.data() + oo[m], n, queries[mapped[m]].name, queries[mapped[m]].len,
                                     names.data(), ids.lens.data(), buf.data(), (int64_t)buf.size(), &need);
      }
      if (got < 0) { if (ix) wfb_index_free(ix); return (int)got; }
      text.append(buf.data(), (size_t)got);
    }
    n_out = oo.back();
    filter_seconds = now_s() - t_f;
  }
  wfb_trace_mark_("map_phase: chain + filters + text");
  if (ix) wfb_index_free(ix);
  *paf = to_c_text(text);
  wfb_trace_mark_("map_phase: index freed, text copied");
  if (!*paf) { wfb_set_last_error_("out of host memory"); return WFB_ENOMEM; }
  *paf_len = (int64_t)text.size();
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->fragments = (int64_t)frags.size(); stats->l2_mappings = n_l2; stats->mappings = n_out; stats->sketch_size = s; stats->minimum_hits = min_hits;
    stats->percentage_identity = P.percentage_identity; stats->index_seconds = index_seconds; stats->map_kernel_ms = map_ms; stats->filter_seconds = filter_seconds;
    stats->total_seconds = now_s() - t_begin; stats->stale_absorbed = stale_absorbed; stats->ani_seconds = ani_seconds;
    stats->index_kernel_ms = index_kernel_ms; stats->ani_kernel_ms = ani_kernel_ms;
  }
  return WFB_OK;
}

extern "C" int wfb_align_phase(wfb_aligner_t* aligner, const wfb_align_phase_params_t* params, const char* mapping_paf, int64_t mapping_paf_len,
                               const wfb_seq_t* targets, int32_t n_targets, const wfb_seq_t* queries, int32_t n_queries, char** out, int64_t* out_len,
                               wfb_align_phase_stats_t* stats) {
  if (!aligner || !params || !out || !out_len || mapping_paf_len < 0 || (mapping_paf_len > 0 && !mapping_paf) || n_targets < 0 || n_queries < 0 ||
      (n_targets > 0 && !targets) || (n_queries > 0 && !queries)) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  const double t_begin = now_s();
  wfb_trace_mark_("align_phase: begin");
  *out = nullptr; *out_len = 0;
  std::unordered_map<std::string, const wfb_seq_t*> tmap, qmap;
  for (int32_t i = 0; i < n_targets; ++i) tmap.emplace(targets[i].name, &targets[i]);
  for (int32_t i = 0; i < n_queries; ++i) qmap.emplace(queries[i].name, &queries[i]);
  /* makeUpperCaseAndValidDNA (commonFunc.hpp:132-142) and reverseComplement (commonFunc.hpp:74-83) as byte tables */
  unsigned char clean[256], comp[256];
  for (int c = 0; c < 256; ++c) { clean[c] = 'N'; comp[c] = (unsigned char)c; }
  for (const char* p = "ACGT"; *p; ++p) { clean[(unsigned char)*p] = (unsigned char)*p; clean[(unsigned char)(*p + 32)] = (unsigned char)*p; }
  comp['A'] = 'T'; comp['C'] = 'G'; comp['G'] = 'C'; comp['T'] = 'A';

  struct Rec { std::string qname, tname, query, target; wfb_mapping_row_t row; const wfb_seq_t *qs, *ts; int64_t q0, q1, t0, t1; };
  std::vector<Rec> recs;
  int64_t skipped = 0, slice_bytes = 0;
  uint64_t aligned_bp = 0;
  for (int64_t a = 0; a < mapping_paf_len;) {
    const char* nl = (const char*)memchr(mapping_paf + a, '\n', (size_t)(mapping_paf_len - a));
    const int64_t b = nl ? (int64_t)(nl - mapping_paf) : mapping_paf_len;
    if (b > a) {
      Rec r;
      if (wfb_mapping_paf_parse(mapping_paf + a, b - a, params->target_padding, params->query_padding, params->wflign_max_len_minor, &r.row) != WFB_OK) ++skipped;
      else {
        r.qname.assign(mapping_paf + a + r.row.q_name_off, (size_t)r.row.q_name_len);
        r.tname.assign(mapping_paf + a + r.row.r_name_off, (size_t)r.row.r_name_len);
        auto qi = qmap.find(r.qname), ti = tmap.find(r.tname);
        if (qi == qmap.end() || ti == tmap.end()) ++skipped; /* "sequence not found": the reference reports and drops the record */
        else {
          /* faidx clamps a fetch to the sequence (src/common/faigz.h:432-438); an empty fetch drops the record */
          const int64_t q0 = std::min(r.row.q_start, qi->second->len), q1 = std::min(r.row.q_end, qi->second->len);
          const int64_t t0 = std::min(r.row.r_start, ti->second->len), t1 = std::min(r.row.r_end, ti->second->len);
          if (q1 <= q0 || t1 <= t0) ++skipped;
          else {
            r.qs = qi->second; r.ts = ti->second; r.q0 = q0; r.q1 = q1; r.t0 = t0; r.t1 = t1;
            slice_bytes += (q1 - q0) + (t1 - t0);
            aligned_bp += (uint64_t)(r.row.q_end - r.row.q_start);
            recs.push_back(std::move(r));
          }
        }
      }
    }
    a = b + 1;
  }
  wfb_trace_mark_("align_phase: rows parsed");
  /* the records' slices (upper-cased, N-masked, query strand-corrected), over the host cores: records are independent */
  {
    const int64_t nrec = (int64_t)recs.size();
    int nt = (int)std::min<int64_t>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u), nrec / 4);
#ifndef WFB_HOST_PAR_MIN_BYTES
#define WFB_HOST_PAR_MIN_BYTES (1 << 20)
#endif
    if (slice_bytes < (int64_t)WFB_HOST_PAR_MIN_BYTES) nt = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
      for (int64_t i; (i = next.fetch_add(1)) < nrec;) {
        Rec& r = recs[(size_t)i];
        const int64_t tn = r.t1 - r.t0, qn = r.q1 - r.q0;
        r.target.resize((size_t)tn);
        for (int64_t j = 0; j < tn; ++j) r.target[(size_t)j] = (char)clean[(unsigned char)r.ts->seq[r.t0 + j]];
        r.query.resize((size_t)qn);
        if (r.row.strand == 1) for (int64_t j = 0; j < qn; ++j) r.query[(size_t)j] = (char)clean[(unsigned char)r.qs->seq[r.q0 + j]];
        else for (int64_t j = 0; j < qn; ++j) r.query[(size_t)j] = (char)comp[clean[(unsigned char)r.qs->seq[r.q1 - 1 - j]]];
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
  }
  wfb_trace_mark_("align_phase: slices");
  int64_t written = 0;
  double kernel_ms = 0;
  /* One launch per batch, and every launch ends with a tail (few CTAs finishing the costliest records alone): as few batches as the
   * memory allows. A 50 kb record needs ~0.45 MB of device memory (4 sequence copies + 2 x (plen + tlen) operation slots). */
  const int32_t batch = params->batch_records > 0 ? params->batch_records : 32768;
  /* The records are independent (computeAlignments.hpp:398-435) and their cost is wildly uneven (~ score^2: a 50 kb record at 20 %
   * divergence costs a thousand times one at 1 %). Batches are formed in order of decreasing expected cost — expected edits from the
   * mapping's identity estimate — so that records of similar cost share a launch: the heavy ones keep all CTAs busy together
   * instead of each batch ending with its one heavy record. The text is re-assembled in row order below. */
  std::vector<uint32_t> order(recs.size());
  for (size_t i = 0; i < recs.size(); ++i) order[i] = (uint32_t)i;
  {
    std::vector<float> key(recs.size());
    for (size_t i = 0; i < recs.size(); ++i) {
      const float id = recs[i].row.mashmap_estimated_identity;
      const float d = (id > 0.f && id <= 1.f) ? std::max(1.f - id, 0.002f) : 0.05f;
      key[i] = d * (float)std::max(recs[i].query.size(), recs[i].target.size());
    }
    if (!(getenv("WFB_COST_SORT") && atoi(getenv("WFB_COST_SORT")) == 0))
      std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return key[x] > key[y]; });
  }
  std::vector<std::string> line(recs.size());
  wfb_align_stats_t sum; memset(&sum, 0, sizeof sum);
  int64_t n_batches = 0;
  std::vector<wfb_record_t> arr;
  std::unique_ptr<char[]> buf; /* not a vector: resize() would zero-fill hundreds of MB on one thread */
  size_t buf_cap = 0;
  for (size_t b0 = 0; b0 < recs.size(); b0 += (size_t)batch) {
    const size_t b1 = std::min(recs.size(), b0 + (size_t)batch);
    arr.assign(b1 - b0, wfb_record_t());
    size_t bytes = 4096;
    for (size_t i = b0; i < b1; ++i) {
      const Rec& r = recs[order[i]];
      wfb_record_t& w = arr[i - b0];
      memset(&w, 0, sizeof w);
      w.query_name = r.qname.c_str(); w.query = r.query.data(); w.query_total_length = (uint64_t)r.qs->len; w.query_offset = (uint64_t)r.row.q_start;
      w.query_length = (uint64_t)r.query.size(); w.query_is_rev = r.row.strand != 1; w.chain_id = (int32_t)r.row.chain_id;
      w.target_name = r.tname.c_str(); w.target = r.target.data(); w.target_total_length = (uint64_t)r.ts->len; w.target_offset = (uint64_t)r.row.r_start;
      w.target_length = (uint64_t)r.target.size(); w.chain_length = (int32_t)r.row.chain_length; w.chain_pos = (int32_t)r.row.chain_pos;
      w.mashmap_estimated_identity = r.row.mashmap_estimated_identity;
      bytes += (r.query.size() + r.target.size()) * (params->output.sam_format ? 2 : 1) / (params->output.sam_format ? 1 : 4) + 1024;
    }
    std::vector<int64_t> off(b1 - b0 + 1); std::vector<int32_t> st(b1 - b0);
    int64_t len = 0;
    wfb_align_stats_t as; memset(&as, 0, sizeof as);
    if (bytes > buf_cap) { buf.reset(new (std::nothrow) char[bytes]); buf_cap = buf ? bytes : 0; }
    if (!buf) { wfb_set_last_error_("out of host memory"); return WFB_ENOMEM; }
    int rc = wfb_biwfa_paf_batch(aligner, arr.data(), (int32_t)(b1 - b0), &params->output, buf.get(), (int64_t)buf_cap, &len, off.data(), st.data(), &as);
    if (rc == WFB_ECAP && len > (int64_t)buf_cap) {
      buf.reset(new (std::nothrow) char[(size_t)len + 64]); buf_cap = buf ? (size_t)len + 64 : 0;
      if (!buf) { wfb_set_last_error_("out of host memory"); return WFB_ENOMEM; }
      rc = wfb_biwfa_paf_batch(aligner, arr.data(), (int32_t)(b1 - b0), &params->output, buf.get(), (int64_t)buf_cap, &len, off.data(), st.data(), &as);
    }
    if (rc != WFB_OK) return rc;
    wfb_trace_mark_("align_phase: batch aligned");
    /* Aligner::processMappingRecord re-emits every line that carries a cg:Z: field as its whitespace-separated fields joined by
     * single tabs (computeAlignments.hpp:486-516): the trailing tab do_biwfa_alignment writes before the newline disappears.
     * Lines without a CIGAR field (SAM records) pass through unchanged. */
    auto reemit = [&](size_t i) {
      std::string& text = line[order[i]];
      for (int64_t a = off[i - b0]; a < off[i - b0 + 1];) {
        const char* nl = (const char*)memchr(buf.get() + a, '\n', (size_t)(off[i - b0 + 1] - a));
        const int64_t b = nl ? (int64_t)(nl - buf.get()) : off[i - b0 + 1];
        if (b > a) {
          std::vector<std::pair<int64_t, int64_t>> fields;
          bool has_cigar = false;
          for (int64_t p = a; p < b;) {
            while (p < b && isspace((unsigned char)buf[(size_t)p])) ++p;
            if (p >= b) break;
            const int64_t f0 = p;
            while (p < b && !isspace((unsigned char)buf[(size_t)p])) ++p;
            fields.emplace_back(f0, p);
            if (p - f0 >= 5 && memcmp(buf.get() + f0, "cg:Z:", 5) == 0) has_cigar = true;
          }
          if (has_cigar) {
            for (size_t f = 0; f < fields.size(); ++f) {
              if (f) text.push_back('\t');
              text.append(buf.get() + fields[f].first, (size_t)(fields[f].second - fields[f].first));
            }
            text.push_back('\n');
          } else {
            text.append(buf.get() + a, (size_t)(b - a));
            text.push_back('\n');
          }
        }
        a = b + 1;
      }
    };
    { /* records are independent, every line lands in its own slot: over the host cores */
      int nt = (int)std::min<size_t>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u), (b1 - b0) / 4);
      if (len < (int64_t)WFB_HOST_PAR_MIN_BYTES) nt = 1;
      std::atomic<size_t> next(b0);
      auto worker = [&]() {
        for (;;) {
          const size_t lo = next.fetch_add(64);
          if (lo >= b1) return;
          const size_t hi = std::min(b1, lo + 64);
          for (size_t i = lo; i < hi; ++i) reemit(i);
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(worker);
      worker();
      for (auto& t : th) t.join();
    }
    for (size_t i = 0; i < st.size(); ++i) written += st[i] == WFB_REC_WRITTEN;
    kernel_ms += as.kernel_ms;
    sum.break_kernel_ms += as.break_kernel_ms; sum.patch_kernel_ms += as.patch_kernel_ms; sum.cells += as.cells; sum.base_cells += as.base_cells;
    sum.extend_matches += as.extend_matches; sum.base_extend_matches += as.base_extend_matches; sum.overlap_tests += as.overlap_tests;
    sum.score_steps += as.score_steps; sum.base_score_steps += as.base_score_steps; sum.h2d_bytes += as.h2d_bytes; sum.d2h_bytes += as.d2h_bytes;
    sum.patch_cap_kept_main += as.patch_cap_kept_main; sum.main_device_cap += as.main_device_cap;
    ++n_batches;
  }
  wfb_trace_mark_("align_phase: lines re-emitted");
  std::string text;
  {
    size_t total = 0;
    for (const std::string& l : line) total += l.size();
    text.reserve(total);
    for (const std::string& l : line) text += l;
  }
  *out = to_c_text(text);
  wfb_trace_mark_("align_phase: text assembled");
  if (!*out) { wfb_set_last_error_("out of host memory"); return WFB_ENOMEM; }
  *out_len = (int64_t)text.size();
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->records = (int64_t)recs.size(); stats->written = written; stats->skipped_lines = skipped; stats->aligned_bp = aligned_bp; stats->kernel_ms = kernel_ms;
    stats->persist_kernel_ms = sum.break_kernel_ms; stats->patch_kernel_ms = sum.patch_kernel_ms; stats->batches = n_batches; stats->cells = sum.cells;
    stats->base_cells = sum.base_cells; stats->extend_matches = sum.extend_matches; stats->base_extend_matches = sum.base_extend_matches;
    stats->overlap_tests = sum.overlap_tests; stats->score_steps = sum.score_steps; stats->base_score_steps = sum.base_score_steps;
    stats->h2d_bytes = sum.h2d_bytes; stats->d2h_bytes = sum.d2h_bytes;
    stats->patch_cap_kept_main = sum.patch_cap_kept_main; stats->main_device_cap = sum.main_device_cap;
    stats->total_seconds = now_s() - t_begin;
  }
  return WFB_OK;
}
