/* oracle/chain.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * Scalar pieces of the Metropolis-Hastings step (reference src/mcmc_eq.c:137-229, 885-895,
 * 1038-1039, 1070-1071, 1114-1117), pinned against oracle/_ref by tests/test_oracle_pin.py. */
#ifndef ORACLE_CHAIN_H
#define ORACLE_CHAIN_H
#ifdef __cplusplus
extern "C" {
#endif
float ch_nexp(float v);                                                  /* src/mcmc_eq.c:137-142 */
/* 0 = valid, 1 = invalid; src/mcmc_eq.c:180-229 */
int ch_model_valid(int dim, const float *z, const float *vp, const float *vpvs, float dz, float zmin, float zmax,
                   float inv_control);
/* sum_c mf_c/sigma_c^2 with the reference's float/double mix; mf and noise indexed 2*class+phase */
double ch_misfit(const float *mf, const float *noise);                   /* src/mcmc_eq.c:885-888 */
double ch_rms(const float *mf, int sum_of_picks);                        /* src/mcmc_eq.c:889 */
float ch_alpha(double log_fac, double new_ll, double old_ll);            /* src/mcmc_eq.c:895,1048 */
double ch_logfac_birth(float sdevvp, float vpmin, float vpmax, float vp_new, float vp_parent, float sdevvpvs,
                       float vpvsmin, float vpvsmax, float vpvs_new, float vpvs_parent); /* :1038-1039 */
double ch_logfac_death(float sdevvp, float vpmin, float vpmax, float vp_dead, float vp_nb, float sdevvpvs,
                       float vpvsmin, float vpvsmax, float vpvs_dead, float vpvs_nb);    /* :1070-1071 */
double ch_logfac_noise(const int *n_class, const float *noise_old, const float *noise_new); /* :1114-1117 */
#ifdef __cplusplus
}
#endif
#endif
