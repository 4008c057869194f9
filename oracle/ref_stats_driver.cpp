/*
 * TEST INFRASTRUCTURE ONLY. Compiles the reference's UNMODIFIED src/map/include/map_stats.hpp — Stat::j2md / md2j /
 * md_lower_bound / estimateMinimumHitsRelaxed (:56-180) and Stat::estimate_identity_for_groups (:325-822, the ANI
 * auto-identity of the default CLI run) — with the silent progress-meter shim, oracle/shims/common/faigz.h
 * (FASTA access; htslib absent) and oracle/shims/gsl (GSL absent: the two cdf functions it needs are defined below from
 * their textbook definitions — the reference's LOGIC around them is what this library pins, GSL's last digits are not).
 */
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "map/include/map_stats.hpp"

static double ln_choose(double n, double m) { return lgamma(n + 1) - lgamma(m + 1) - lgamma(n - m + 1); }
extern "C" double gsl_cdf_binomial_Q(unsigned int k, double p, unsigned int n) {
  if (k >= n || p <= 0) return 0.0;
  if (p >= 1) return 1.0;
  long double q = 0;
  for (unsigned i = k + 1; i <= n; ++i) q += expl((long double)ln_choose(n, i) + i * logl(p) + (n - i) * log1pl(-p));
  return (double)(q > 1 ? 1 : q);
}
extern "C" double gsl_ran_hypergeometric_pdf(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t) {
  if (t > n1 + n2) t = n1 + n2;
  if (k > n1 || k > t) return 0;
  if (t > n2 && k + n2 < t) return 0;
  return exp(ln_choose(n1, k) + ln_choose(n2, t - k) - ln_choose(n1 + n2, t));
}
extern "C" double gsl_cdf_hypergeometric_P(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t) {
  double p = 0;
  for (unsigned i = 0; i <= k; ++i) p += gsl_ran_hypergeometric_pdf(i, n1, n2, t);
  return p > 1 ? 1 : p;
}

static void write_fasta(const char* path, const char* const* names, const char* const* seqs, const int64_t* lens, int32_t n) {
  std::ofstream fa(path), fai(std::string(path) + ".fai");
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

extern "C" {
int32_t ref_estimate_minimum_hits_relaxed(int32_t s, int32_t k, float pid, float ci) { return skch::Stat::estimateMinimumHitsRelaxed(s, k, pid, ci); }
float ref_md_lower_bound(float d, int32_t s, int32_t k, float ci) { return skch::Stat::md_lower_bound(d, s, k, ci); }
float ref_j2md(float j, int32_t k) { return skch::Stat::j2md(j, k); }
float ref_md2j(float d, int32_t k) { return skch::Stat::md2j(d, k); }

/* Stat::estimate_identity_for_groups exactly as main.cpp:75-104 calls it. The query / target sequences are written to two
 * FASTA files (the same file twice when same_file). Returns the adjusted identity the CLI would adopt. */
double ref_estimate_identity(const char* dir, const char* const* q_names, const char* const* q_seqs, const int64_t* q_lens, int32_t nq,
                             const char* const* t_names, const char* const* t_seqs, const int64_t* t_lens, int32_t nt, int32_t same_file,
                             const char* prefix_delim, int32_t ani_percentile, float ani_adjustment, int32_t threads) {
  const std::string qf = std::string(dir) + "/q.fa", tf = same_file ? qf : std::string(dir) + "/t.fa";
  write_fasta(qf.c_str(), q_names, q_seqs, q_lens, nq);
  if (!same_file) write_fasta(tf.c_str(), t_names, t_seqs, t_lens, nt);
  skch::Parameters p;
  p.querySequences = {qf}; p.refSequences = {tf}; p.threads = threads; p.use_progress_bar = false;
  p.ani_percentile = ani_percentile; p.ani_adjustment = ani_adjustment; p.prefix_delim = prefix_delim && prefix_delim[0] ? prefix_delim[0] : '\0';
  skch::SequenceIdManager ids(p.querySequences, p.refSequences, {}, {}, std::string(prefix_delim ? prefix_delim : ""));
  return skch::Stat::estimate_identity_for_groups(p, ids);
}
}
