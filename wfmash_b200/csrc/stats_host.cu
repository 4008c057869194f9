// stats_host.cu — the run-level constants the mapping kernels take as inputs (SURVEY §8 a6 / f1 "FP-derived once per run"):
//   * sketch size from identity / window / k                    (src/interface/parse_args.hpp:642-644)
//   * minimum L1 hits  = Stat::estimateMinimumHitsRelaxed       (src/map/include/map_stats.hpp:56-180)
//   * sketchCutoffs    = Map::setProbs                          (src/map/include/computeMap.hpp:234-293)
//   * the L2 identity test with keep_low_pct_id (the CLI default; computeMap.hpp:1016-1024) as a table over
//     (Q.sketchSize, sharedSketchSize), like wfb_l2_min_shared does for the plain test.
// The reference takes three distribution functions from GNU GSL (a third-party dependency that is not vendored in the
// reference tree and not installed here): gsl_cdf_binomial_Q, gsl_ran_hypergeometric_pdf, gsl_cdf_hypergeometric_P.
// They are restated from their definitions with log-gamma sums in double precision; every consumer below only compares
// them with a threshold, so the integer outputs are insensitive to last-digit differences. PARITY UNPINNED at this
// boundary (no reference test holds these values); tests/test_stats_cpu.py cross-checks the three functions against
// scipy.stats and the derived integers against an independent numpy restatement.
#include "wfmash_b200.h"

#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */

namespace {

inline float st_j2md(float j, int k) { /* Stat::j2md, map_stats.hpp:56-66 */
  if (j == 0) return 1.0f;
  if (j == 1) return 0.0f;
  const float mash_dist = 1 - std::pow(2 * j / (1 + j), 1.0 / k);
  return mash_dist;
}
inline float st_md2j(float d, int k) { /* Stat::md2j, map_stats.hpp:74-79 */
  const float sim = 1 - d;
  const float jaccard = std::pow(sim, k) / (2 - std::pow(sim, k));
  return jaccard;
}

inline double ln_choose(double n, double m) { return lgamma(n + 1) - lgamma(m + 1) - lgamma(n - m + 1); }

/* gsl_cdf_binomial_Q(k, p, n) = P(X > k), X ~ Binomial(n, p) */
double binomial_Q(unsigned k, double p, unsigned n) {
  if (k >= n) return 0.0;
  if (p <= 0.0) return 0.0;
  if (p >= 1.0) return 1.0;
  const double lp = log(p), lq = log1p(-p);
  /* sum the shorter tail */
  const double mean = n * p;
  if ((double)k + 1 > mean) {
    double q = 0;
    for (unsigned i = n; i > k; --i) q += exp(ln_choose(n, i) + i * lp + (n - i) * lq);
    return std::min(q, 1.0);
  }
  double c = 0;
  for (unsigned i = 0; i <= k; ++i) c += exp(ln_choose(n, i) + i * lp + (n - i) * lq);
  return std::max(0.0, 1.0 - c);
}

/* gsl_ran_hypergeometric_pdf(k, n1, n2, t): k successes in t draws without replacement from n1 good + n2 bad */
double hypergeometric_pdf(unsigned k, unsigned n1, unsigned n2, unsigned t) {
  if (t > n1 + n2) t = n1 + n2;
  if (k > n1 || k > t) return 0.0;
  if (t > n2 && k + n2 < t) return 0.0;
  return exp(ln_choose(n1, k) + ln_choose(n2, t - k) - ln_choose(n1 + n2, t));
}

/* gsl_cdf_hypergeometric_P(k, n1, n2, t) = P(X <= k) */
double hypergeometric_P(unsigned k, unsigned n1, unsigned n2, unsigned t) {
  if (t > n1 + n2) t = n1 + n2;
  if (k >= n1 || k >= t) return 1.0;
  double p = 0;
  for (unsigned i = 0; i <= k; ++i) p += hypergeometric_pdf(i, n1, n2, t);
  return std::min(p, 1.0);
}

float md_lower_bound(float d, int s, int k, float ci) { /* Stat::md_lower_bound, map_stats.hpp:93-126 (GSL branch) */
  const float q2 = (1.0 - ci) / 2;
  int x = std::max(int(ceil(s * st_md2j(d, k))), 1);
  while (x <= s) {
    const double cdf_complement = binomial_Q((unsigned)(x - 1), st_md2j(d, k), (unsigned)s);
    if (cdf_complement < q2) { x--; break; }
    x++;
  }
  const float jaccard = float(x) / s;
  return st_j2md(jaccard, k);
}

int estimate_minimum_hits(int s, int k, float perc_identity) { /* map_stats.hpp:135-147 */
  const float mash_dist = 1.0 - perc_identity;
  const float jaccard = st_md2j(mash_dist, k);
  return (int)ceil(1.0 * s * jaccard);
}

}  // namespace

extern "C" int32_t wfb_sketch_size(float percentage_identity, int64_t window_length, int32_t kmer_size) {
  const double md = 1 - percentage_identity; /* parse_args.hpp:642-644: float identity, double arithmetic, truncation to int */
  const double dens = 0.02 * (1 + (md / 0.1));
  return (int32_t)(dens * (window_length - kmer_size));
}

extern "C" int32_t wfb_estimate_minimum_hits_relaxed(int32_t sketch_size, int32_t kmer_size, float percentage_identity, float confidence_interval) {
  if (sketch_size < 1 || kmer_size < 1) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  const int first = estimate_minimum_hits(sketch_size, kmer_size, percentage_identity); /* map_stats.hpp:159-180 */
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    const float jaccard = 1.0 * i / sketch_size;
    const float d = st_j2md(jaccard, kmer_size);
    const float d_lower = md_lower_bound(d, sketch_size, kmer_size, confidence_interval);
    const float id_upper = 1.0 - d_lower;
    if (id_upper >= percentage_identity) relaxed = i;
    else break;
  }
  return relaxed;
}

extern "C" int wfb_sketch_cutoffs(int32_t sketch_size, int32_t kmer_size, float ani_diff, float ani_diff_conf, int32_t stage1_top_ani_filter, int32_t* out,
                                  int32_t out_len) {
  const int ss = (int)std::min<double>(sketch_size, 1000.0); /* skch::fixed::ss_table_max */
  if (!out || sketch_size < 1 || kmer_size < 1 || out_len < ss + 1) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  for (int i = 0; i <= ss; ++i) out[i] = 1; /* Map's constructor, computeMap.hpp:150 */
  if (!stage1_top_ani_filter) return WFB_OK; /* setProbs only runs with the stage-1 filter, computeMap.hpp:224-226 */
  const float deltaANI = ani_diff;
  const float min_p = 1 - ani_diff_conf;
  std::vector<std::vector<double>> prob((size_t)ss + 1, std::vector<double>((size_t)ss + 1, 0.0));
  for (int ci = 0; ci <= ss; ci++)
    for (int y = 0; y <= ci; y++) prob[(size_t)ci][(size_t)y] = hypergeometric_pdf((unsigned)y, (unsigned)ss, (unsigned)(ss - ci), (unsigned)ci);
  auto dist_diff = [&](int cmax, int ci) { /* computeMap.hpp:253-270 */
    double above = 0;
    for (double ymax = 0; ymax <= cmax; ymax++) {
      const double pymax = prob[(size_t)cmax][(size_t)ymax];
      const double yi_cutoff = deltaANI == 0 ? ymax : std::floor(st_md2j(st_j2md(ymax / ss, kmer_size) + deltaANI, kmer_size) * ss);
      double pi_acc = (yi_cutoff - 1) >= 0 ? hypergeometric_P((unsigned)(yi_cutoff - 1), (unsigned)ss, (unsigned)(ss - ci), (unsigned)ci) : 0;
      pi_acc = 1 - pi_acc;
      above += pymax * pi_acc;
      if (above > min_p) return true;
    }
    return above > min_p;
  };
  std::vector<int> range((size_t)ss + 1);
  std::iota(range.begin(), range.end(), 0);
  for (int cmax = 1; cmax <= ss; cmax++) { /* the reference's binary search, with its comparator that ignores the probe value */
    const int ci = (int)std::distance(range.begin(), std::upper_bound(range.begin(), range.begin() + ss, false,
                                                                      [&](bool, int c) { return dist_diff(cmax, c); }));
    out[cmax] = ci == 0 ? 1 : ci;
  }
  return WFB_OK;
}

extern "C" int wfb_l2_min_shared_relaxed(float percentage_identity, int32_t kmer_size, int32_t sketch_size, float confidence_interval, int32_t* out) {
  if (!out || sketch_size < 1 || kmer_size < 1) { wfb_set_last_error_("bad argument"); return WFB_EINVAL; }
  out[0] = 0;
  for (int qs = 1; qs <= sketch_size; ++qs) { /* computeMap.hpp:1016-1024 with keep_low_pct_id == true, for Q.sketchSize == qs */
    int v = 0;
    for (; v <= qs; ++v) {
      const float mash_dist = st_j2md(1.0 * v / qs, kmer_size);
      const float nucIdentity = (1 - mash_dist);
      const float upper = 1 - md_lower_bound(mash_dist, qs, kmer_size, confidence_interval);
      if (upper >= percentage_identity || nucIdentity >= percentage_identity) break;
    }
    out[qs] = v; /* qs + 1 = nothing passes */
  }
  return WFB_OK;
}

/* test hooks for the three GSL restatements (cross-checked against scipy.stats) */
extern "C" double wfb_stat_binomial_Q(uint32_t k, double p, uint32_t n) { return binomial_Q(k, p, n); }
extern "C" double wfb_stat_hypergeometric_pdf(uint32_t k, uint32_t n1, uint32_t n2, uint32_t t) { return hypergeometric_pdf(k, n1, n2, t); }
extern "C" double wfb_stat_hypergeometric_P(uint32_t k, uint32_t n1, uint32_t n2, uint32_t t) { return hypergeometric_P(k, n1, n2, t); }
