/*
 * TEST INFRASTRUCTURE ONLY. Compiles the reference's UNMODIFIED src/map/include/winSketch.hpp (skch::Sketch: per-sequence
 * addMinmers through its thread pool, hash frequencies, the frequency cut-off with its safety re-threshold, the
 * minmerPosLookupIndex / minmerIndex build, winSketch.hpp:175-457) and its UNMODIFIED sequenceIds.hpp against
 * oracle/shims/htslib/faidx.h (htslib is absent: the shim reads an uncompressed FASTA + .fai) and the silent
 * progress-meter shim, so that SURVEY 8 row a4 (index build) is pinned by the real code instead of by reading it.
 * The driver writes the caller's sequences to a FASTA + .fai, constructs SequenceIdManager and Sketch exactly like
 * Map::Map / the build_index task do (computeMap.hpp:150-175,472-484) and flattens the public members.
 */
#include <cassert>
#include <climits>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "map/include/winSketch.hpp"

static_assert(sizeof(skch::MinmerInfo) == 32, "MinmerInfo layout");

struct ref_point { uint64_t hash; int64_t pos; int32_t seqId; int32_t side; };

extern "C" {
struct ref_sketch {
  skch::SequenceIdManager* ids;
  skch::Sketch* sk;
  std::vector<std::string> targets;
};

static void write_fasta(const char* fasta_path, const char* const* names, const char* const* seqs, const int64_t* lens, int32_t n) {
  std::ofstream fa(fasta_path), fai(std::string(fasta_path) + ".fai");
  int64_t off = 0;
  for (int32_t i = 0; i < n; ++i) {
    const std::string head = std::string(">") + names[i] + "\n";
    fa << head;
    off += (int64_t)head.size();
    fa.write(seqs[i], lens[i]);
    fa << "\n";
    fai << names[i] << "\t" << lens[i] << "\t" << off << "\t" << (lens[i] > 0 ? lens[i] : 1) << "\t" << (lens[i] > 0 ? lens[i] : 1) + 1 << "\n";
    off += lens[i] + 1;
  }
}

/* names[i] / seqs[i] / lens[i]: target sequences in file order (ids 0..n-1). Returns NULL on failure. */
void* ref_sketch_build(const char* fasta_path, const char* const* names, const char* const* seqs, const int64_t* lens, int32_t n, int32_t kmer_size,
                       int64_t window_length, int32_t sketch_size, int32_t threads, double max_kmer_freq, const char* prefix_delim) {
  write_fasta(fasta_path, names, seqs, lens, n);
  skch::Parameters p;
  p.kmerSize = kmer_size; p.windowLength = window_length; p.sketchSize = sketch_size; p.threads = threads; p.max_kmer_freq = max_kmer_freq;
  p.refSequences = {fasta_path}; p.querySequences = {fasta_path}; p.alphabetSize = 4; p.use_progress_bar = false; p.hgNumerator = 1.0;
  p.percentageIdentity = 0.9f; p.world_minimizers = false; p.use_spaced_seeds = false;
  p.use_streaming_minhash = false; /* the CLI sets it from a flag that defaults to off (parse_args.hpp:177): windowed addMinmers */
  ref_sketch* h = new ref_sketch;
  h->ids = new skch::SequenceIdManager({fasta_path}, {fasta_path}, {}, {}, std::string(prefix_delim ? prefix_delim : ""));
  h->targets.assign(names, names + n);
  h->sk = new skch::Sketch(p, *h->ids, h->targets);
  return h;
}

/* Sketch::writeIndex (winSketch.hpp:616-635): the `-W` file of this index */
void ref_sketch_write_index(void* hv, const char* index_path) {
  ref_sketch* h = (ref_sketch*)hv;
  h->sk->writeIndex(h->targets, index_path, false, 0, 1);
}

/* Sketch(param, idManager, targets, &indexStream) = Sketch::readIndex (winSketch.hpp:840-866): the `-I` path. The FASTA + .fai
 * are only needed by the SequenceIdManager. */
void* ref_sketch_read_index(const char* fasta_path, const char* index_path, const char* const* names, const char* const* seqs, const int64_t* lens,
                            int32_t n, int32_t kmer_size, int64_t window_length, int32_t sketch_size, const char* prefix_delim) {
  write_fasta(fasta_path, names, seqs, lens, n);
  skch::Parameters p;
  p.kmerSize = kmer_size; p.windowLength = window_length; p.sketchSize = sketch_size; p.threads = 1; p.max_kmer_freq = 0.0002;
  p.refSequences = {fasta_path}; p.alphabetSize = 4; p.use_progress_bar = false; p.hgNumerator = 1.0; p.prefix_delim = prefix_delim && prefix_delim[0] ? prefix_delim[0] : '\0';
  p.percentageIdentity = 0.9f; p.use_streaming_minhash = false;
  ref_sketch* h = new ref_sketch;
  h->ids = new skch::SequenceIdManager({}, {fasta_path}, {}, {}, std::string(prefix_delim ? prefix_delim : ""));
  h->targets.assign(names, names + n);
  std::ifstream in(index_path, std::ios::binary);
  if (!in) { delete h->ids; delete h; return nullptr; }
  h->sk = new skch::Sketch(p, *h->ids, h->targets, &in);
  return h;
}

void ref_sketch_sizes(void* hv, int64_t* n_minmers, int64_t* n_hashes, int64_t* n_points) {
  ref_sketch* h = (ref_sketch*)hv;
  *n_minmers = (int64_t)h->sk->minmerIndex.size();
  *n_hashes = (int64_t)h->sk->minmerPosLookupIndex.size();
  int64_t p = 0;
  for (auto& kv : h->sk->minmerPosLookupIndex) p += (int64_t)kv.second.size();
  *n_points = p;
}

/* minmers[n_minmers] in the reference's order; points grouped by ascending hash (the map itself is unordered), each
 * group in the reference's order; hash_start[n_hashes + 1] */
void ref_sketch_export(void* hv, skch::MinmerInfo* minmers, uint64_t* hashes, int64_t* hash_start, ref_point* points) {
  ref_sketch* h = (ref_sketch*)hv;
  std::copy(h->sk->minmerIndex.begin(), h->sk->minmerIndex.end(), minmers);
  std::vector<uint64_t> keys;
  for (auto& kv : h->sk->minmerPosLookupIndex) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  int64_t o = 0;
  for (size_t i = 0; i < keys.size(); ++i) {
    hashes[i] = keys[i];
    hash_start[i] = o;
    for (const auto& ip : h->sk->minmerPosLookupIndex[keys[i]]) points[o++] = ref_point{ip.hash, ip.pos, ip.seqId, (int32_t)ip.side};
  }
  hash_start[keys.size()] = o;
}

int32_t ref_sketch_group(void* hv, int32_t seq_id) { return ((ref_sketch*)hv)->ids->getRefGroup(seq_id); }

void ref_sketch_free(void* hv) {
  ref_sketch* h = (ref_sketch*)hv;
  delete h->sk; delete h->ids; delete h;
}
}
