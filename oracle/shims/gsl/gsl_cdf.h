/* TEST INFRASTRUCTURE ONLY — GNU GSL is not installed and not vendored in the reference tree. Declarations of the two
 * cdf functions the reference's map_stats.hpp / computeMap.hpp call; oracle/ref_stats_driver.cpp defines them from the
 * distributions' definitions (log-gamma sums), independently of the product's restatement in stats_host.cu. */
#ifndef WFB_SHIM_GSL_CDF_H
#define WFB_SHIM_GSL_CDF_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_cdf_binomial_Q(unsigned int k, double p, unsigned int n);
double gsl_cdf_hypergeometric_P(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t);
#ifdef __cplusplus
}
#endif
#endif
