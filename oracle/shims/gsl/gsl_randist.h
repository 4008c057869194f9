/* TEST INFRASTRUCTURE ONLY — see gsl_cdf.h in this directory. */
#ifndef WFB_SHIM_GSL_RANDIST_H
#define WFB_SHIM_GSL_RANDIST_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_ran_hypergeometric_pdf(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t);
#ifdef __cplusplus
}
#endif
#endif
