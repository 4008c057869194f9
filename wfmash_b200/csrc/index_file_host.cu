// index_file_host.cu — the reference's binary index file (second half): Sketch::writeIndex / readIndex and
// their helpers (src/map/include/winSketch.hpp:554-979), SequenceIdManager::exportIdMapping / importIdMapping
// (src/map/include/sequenceIds.hpp:102-200). Pure host I/O over the wfb_index_export layout; wfb_index_import
// (index_host.cu) turns what was read into a device index.
#include "wfmash_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <string>
#include <vector>

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */

namespace {
const uint64_t IXF_MAGIC = 0xDEADBEEFCAFEBABEull;

struct RefPoint { /* skch::IntervalPoint (base_types.hpp:63-76), 24 bytes */
  int64_t pos;
  uint64_t hash;
  int32_t seqId;
  int8_t side; /* OPEN = 1, CLOSE = -1 (base_types.hpp:95-99) */
  int8_t pad_[3];
};
static_assert(sizeof(RefPoint) == 24, "IntervalPoint layout");
static_assert(sizeof(wfb_minmer_t) == 32, "MinmerInfo layout");

std::vector<std::string> split_lines(const char* s) {
  std::vector<std::string> v;
  if (!s) return v;
  const char* a = s;
  for (const char* p = s;; ++p)
    if (*p == '\n' || *p == 0) {
      if (p > a) v.emplace_back(a, (size_t)(p - a));
      if (*p == 0) break;
      a = p + 1;
    }
  return v;
}
bool put(FILE* f, const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
bool get(FILE* f, void* p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }
}  // namespace

extern "C" int wfb_index_file_write(const char* path, int32_t append, const wfb_index_file_header_t* h, const wfb_index_view_t* ix) {
  if (!path || !h || !ix || ix->n_minmers < 0 || ix->n_uniq < 0 || ix->n_points < 0 || (ix->n_minmers > 0 && !ix->minmers) ||
      (ix->n_uniq > 0 && (!ix->uhash || !ix->ustart || !ix->ucount)) || (ix->n_points > 0 && !ix->points)) {
    wfb_set_last_error_("bad argument");
    return WFB_EINVAL;
  }
  const std::vector<std::string> targets = split_lines(h->target_names), ids = split_lines(h->id_names);
  if ((int32_t)targets.size() != h->n_targets || (int32_t)ids.size() != h->n_ids || (h->n_ids > 0 && !h->id_values)) {
    wfb_set_last_error_("header name lists do not match their counts");
    return WFB_EINVAL;
  }
  FILE* f = fopen(path, append ? "ab" : "wb");
  if (!f) { wfb_set_last_error_(std::string("cannot open index file for writing: ") + path); return WFB_EINVAL; }
  bool ok = true;
  /* writeSubIndexHeader, winSketch.hpp:637-659 */
  ok = ok && put(f, &IXF_MAGIC, 8) && put(f, &h->batch_idx, 8) && put(f, &h->total_batches, 8) && put(f, &h->index_by_size, 8);
  const uint64_t nt = targets.size();
  ok = ok && put(f, &nt, 8);
  for (const auto& n : targets) { const uint64_t l = n.size(); ok = ok && put(f, &l, 8) && put(f, n.data(), n.size()); }
  /* exportIdMapping, sequenceIds.hpp:102-115 */
  const uint64_t nm = ids.size();
  ok = ok && put(f, &nm, 8);
  for (size_t i = 0; i < ids.size(); ++i) { const uint64_t l = ids[i].size(); ok = ok && put(f, &l, 8) && put(f, ids[i].data(), ids[i].size()) && put(f, &h->id_values[i], 4); }
  ok = ok && put(f, &h->next_id, 4);
  /* writeParameters, :604-610 */
  ok = ok && put(f, &h->window_length, 8) && put(f, &h->sketch_size, 4) && put(f, &h->kmer_size, 4);
  /* writeSketchBinary, :569-574 */
  const uint64_t n = (uint64_t)ix->n_minmers;
  ok = ok && put(f, &n, 8) && put(f, ix->minmers, sizeof(wfb_minmer_t) * (size_t)n);
  /* writePosListBinary, :579-593 */
  const uint64_t nk = (uint64_t)ix->n_uniq;
  ok = ok && put(f, &nk, 8);
  std::vector<RefPoint> buf;
  for (int64_t i = 0; ok && i < ix->n_uniq; ++i) {
    const uint64_t cnt = ix->ucount[i];
    buf.assign((size_t)cnt, RefPoint());
    for (uint64_t j = 0; j < cnt; ++j) {
      const uint64_t p = ix->points[(size_t)ix->ustart[i] + j];
      RefPoint& r = buf[(size_t)j];
      memset(&r, 0, sizeof r);
      r.pos = (int64_t)((p >> 1) & ((1ull << 40) - 1)); r.hash = ix->uhash[i]; r.seqId = (int32_t)(p >> 41); r.side = (p & 1) ? 1 : -1;
    }
    ok = ok && put(f, &ix->uhash[i], 8) && put(f, &cnt, 8) && put(f, buf.data(), sizeof(RefPoint) * (size_t)cnt);
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) { wfb_set_last_error_(std::string("short write on index file: ") + path); return WFB_EINVAL; }
  return WFB_OK;
}

extern "C" void wfb_index_file_release(wfb_index_file_header_t* h, wfb_index_view_t* ix) {
  if (h) { free(h->target_names); free(h->id_names); free(h->id_values); h->target_names = h->id_names = nullptr; h->id_values = nullptr; }
  if (ix) { free(ix->minmers); free(ix->uhash); free(ix->ustart); free(ix->ucount); free(ix->points); memset(ix, 0, sizeof *ix); }
}

static int index_file_read_impl(const char* path, int64_t* offset, wfb_index_file_header_t* h, wfb_index_view_t* ix);

/* Sizes in the file are not trusted: every count is checked against the bytes the file still holds before anything is allocated, and
 * an allocation failure returns WFB_ENOMEM instead of unwinding through the C boundary. */
extern "C" int wfb_index_file_read(const char* path, int64_t* offset, wfb_index_file_header_t* h, wfb_index_view_t* ix) {
  try {
    return index_file_read_impl(path, offset, h, ix);
  } catch (const std::bad_alloc&) {
    if (h && ix) wfb_index_file_release(h, ix);
    wfb_set_last_error_("out of host memory while reading the index file");
    return WFB_ENOMEM;
  } catch (...) {
    if (h && ix) wfb_index_file_release(h, ix);
    wfb_set_last_error_("unexpected failure while reading the index file");
    return WFB_EINVAL;
  }
}

static int index_file_read_impl(const char* path, int64_t* offset, wfb_index_file_header_t* h, wfb_index_view_t* ix) {
  if (!path || !offset || !h || !ix || *offset < 0) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  memset(h, 0, sizeof *h);
  memset(ix, 0, sizeof *ix);
  FILE* f = fopen(path, "rb");
  if (!f) { wfb_set_last_error_(std::string("cannot open index file: ") + path); return WFB_EINVAL; }
  auto fail = [&](const char* why) { fclose(f); wfb_index_file_release(h, ix); wfb_set_last_error_(std::string(why) + ": " + path); return WFB_EINVAL; };
  if (fseek(f, 0, SEEK_END) != 0) return fail("cannot seek in index file");
  const uint64_t file_bytes = (uint64_t)ftell(f);
  auto left = [&]() -> uint64_t { const long at = ftell(f); return at < 0 || (uint64_t)at > file_bytes ? 0 : file_bytes - (uint64_t)at; };
  if (fseek(f, (long)*offset, SEEK_SET) != 0) return fail("cannot seek in index file");
  uint64_t magic = 0, nt = 0, nm = 0;
  if (!get(f, &magic, 8) || magic != IXF_MAGIC) return fail("invalid magic number in index file"); /* readSubIndexHeader, :869-890 */
  if (!get(f, &h->batch_idx, 8) || !get(f, &h->total_batches, 8) || h->total_batches < 1 || h->total_batches > 1000 || h->batch_idx >= h->total_batches)
    return fail("invalid batch information in index file");
  if (!get(f, &h->index_by_size, 8) || !get(f, &nt, 8) || nt > 1000000) return fail("invalid number of sequences in index file");
  std::string names;
  for (uint64_t i = 0; i < nt; ++i) {
    uint64_t l = 0;
    if (!get(f, &l, 8) || l > 10000) return fail("invalid sequence name length in index file");
    std::string s((size_t)l, '\0');
    if (!get(f, &s[0], (size_t)l)) return fail("truncated index file");
    names += s; names += '\n';
  }
  h->n_targets = (int32_t)nt;
  h->target_names = strdup(names.c_str());
  if (!h->target_names) { fclose(f); wfb_index_file_release(h, ix); wfb_set_last_error_("out of host memory"); return WFB_ENOMEM; }
  if (!get(f, &nm, 8) || nm > 1000000) return fail("invalid mapping size in index file"); /* importIdMapping, sequenceIds.hpp:140-146 */
  names.clear();
  h->id_values = (int32_t*)malloc(4 * (size_t)(nm ? nm : 1));
  if (!h->id_values) { fclose(f); wfb_index_file_release(h, ix); wfb_set_last_error_("out of host memory"); return WFB_ENOMEM; }
  for (uint64_t i = 0; i < nm; ++i) {
    uint64_t l = 0;
    if (!get(f, &l, 8) || l > 10000) return fail("invalid sequence name length in index file");
    std::string s((size_t)l, '\0');
    if (!get(f, &s[0], (size_t)l) || !get(f, &h->id_values[i], 4)) return fail("truncated index file");
    names += s; names += '\n';
  }
  h->n_ids = (int32_t)nm;
  h->id_names = strdup(names.c_str());
  if (!h->id_names) { fclose(f); wfb_index_file_release(h, ix); wfb_set_last_error_("out of host memory"); return WFB_ENOMEM; }
  if (!get(f, &h->next_id, 4)) return fail("truncated index file");
  if (!get(f, &h->window_length, 8) || !get(f, &h->sketch_size, 4) || !get(f, &h->kmer_size, 4)) return fail("truncated index file"); /* readParameters */
  uint64_t n = 0, nk = 0;
  if (!get(f, &n, 8) || n > left() / sizeof(wfb_minmer_t)) return fail("truncated index file"); /* readSketchBinary, :682-688 */
  ix->n_minmers = (int64_t)n;
  ix->minmers = (wfb_minmer_t*)malloc(sizeof(wfb_minmer_t) * (size_t)(n ? n : 1));
  if (!ix->minmers || !get(f, ix->minmers, sizeof(wfb_minmer_t) * (size_t)n)) return fail("truncated index file");
  for (uint64_t i = 0; i < n; ++i) ix->minmers[i].pad_ = 0; /* the reference writes its struct padding as it happens to be */
  if (!get(f, &nk, 8) || nk > left() / 16) return fail("truncated index file"); /* readPosListBinary, :693-708: >= 16 bytes per hash */
  /* the file lists the hashes in the writer's hash-map order: collect, then order them by hash (the export layout) */
  std::vector<uint64_t> keys((size_t)nk), cnts((size_t)nk), starts((size_t)nk);
  std::vector<uint64_t> pts;
  std::vector<RefPoint> buf;
  for (uint64_t i = 0; i < nk; ++i) {
    uint64_t cnt = 0;
    if (!get(f, &keys[(size_t)i], 8) || !get(f, &cnt, 8) || cnt > left() / sizeof(RefPoint)) return fail("truncated index file");
    /* postings offsets are 32-bit (ustart / ucount), a point packs seqId into 22 bits and pos into 40 */
    if (pts.size() + cnt >= (1ull << 32)) return fail("too many interval points for 32-bit postings offsets in index file");
    buf.resize((size_t)cnt);
    if (!get(f, buf.data(), sizeof(RefPoint) * (size_t)cnt)) return fail("truncated index file");
    starts[(size_t)i] = pts.size(); cnts[(size_t)i] = cnt;
    for (const RefPoint& r : buf) {
      if ((uint64_t)(uint32_t)r.seqId >= (1ull << 22) || (uint64_t)r.pos >= (1ull << 40)) return fail("sequence id / position out of range in index file");
      pts.push_back(((uint64_t)(uint32_t)r.seqId << 41) | ((uint64_t)r.pos << 1) | (r.side == 1 ? 1u : 0u));
    }
  }
  *offset = (int64_t)ftell(f);
  fclose(f);
  std::vector<size_t> order((size_t)nk);
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
  ix->n_uniq = (int64_t)nk; ix->n_points = (int64_t)pts.size();
  ix->uhash = (uint64_t*)malloc(8 * (size_t)(nk ? nk : 1)); ix->ustart = (uint32_t*)malloc(4 * (size_t)(nk ? nk : 1)); ix->ucount = (uint32_t*)malloc(4 * (size_t)(nk ? nk : 1));
  ix->points = (uint64_t*)malloc(8 * (pts.size() ? pts.size() : 1));
  if (!ix->uhash || !ix->ustart || !ix->ucount || !ix->points) { wfb_index_file_release(h, ix); wfb_set_last_error_("out of host memory"); return WFB_ENOMEM; }
  uint64_t o = 0;
  for (size_t j = 0; j < order.size(); ++j) {
    const size_t i = order[j];
    ix->uhash[j] = keys[i]; ix->ustart[j] = (uint32_t)o; ix->ucount[j] = (uint32_t)cnts[i];
    memcpy(ix->points + o, pts.data() + starts[i], 8 * (size_t)cnts[i]);
    o += cnts[i];
  }
  return WFB_OK;
}
