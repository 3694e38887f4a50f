// chain.cuh -- pieces of the Metropolis-Hastings step shared between translation units.
#pragma once
#include <math.h>
#include <stdint.h>

#include "state.h"

namespace mq {

// sum_c mf_c / sigma_c^2 with the reference's float/double mix (src/mcmc_eq.c:885-888):
// each mf/sigma/sigma is a float expression, the first two are added in float, the running
// sum is double.  mf and noise are indexed 2*class + phase.
__host__ __device__ inline double chain_misfit(const float* mf, const float* s)
{
    double m = (double)(mf[0] / s[0] / s[0] + mf[1] / s[1] / s[1]);
    for (int c = 1; c < 4; c++) {
        m = m + (double)(mf[2 * c] / s[2 * c] / s[2 * c]);
        m = m + (double)(mf[2 * c + 1] / s[2 * c + 1] / s[2 * c + 1]);
    }
    return m;
}

// sqrt((mfp0+mfp1+mfp2+mfp3+mfs0+mfs1+mfs2+mfs3)/sum_of_picks), src/mcmc_eq.c:889
__host__ __device__ inline double chain_rms(const float* mf, int sum_of_picks)
{
    const float s = ((((((mf[0] + mf[2]) + mf[4]) + mf[6]) + mf[1]) + mf[3]) + mf[5]) + mf[7];
    return sqrt((double)(s / (float)sum_of_picks));
}

// reference nexp (src/mcmc_eq.c:137-142) and alpha12 = min(1, nexp(log_fac + new_ll - old_ll))
__host__ __device__ inline float chain_alpha(double log_fac, double new_ll, double old_ll)
{
    const float v = (float)(log_fac + new_ll - old_ll);
    const double cap = 81.81508377308622;   // log(FLT_MAX / 1000.0)
    const float e = (float)exp(((double)v < cap) ? (double)v : cap);
    return (1.0 < (double)e) ? 1.0f : e;
}

void sampler_destroy(Handle* h);
void comm_destroy(Handle* h);
int forward_current_device(Handle* h, int calct);
// device error words (capi.cu): synchronous check / copy left for a later call / report of a completed copy
int check_device_errors(Handle* h);
int flags_enqueue(Handle* h);
int flags_poll(Handle* h, bool wait);

}  // namespace mq
