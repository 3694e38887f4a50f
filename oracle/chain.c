/* oracle/chain.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  See chain.h. */
#include "chain.h"

#include <float.h>
#include <math.h>

#define CH_PI 3.141592653 /* src/mc.h:50 */
#define CH_MAXDIM 1000    /* MD, src/mc.h:49 */

float ch_nexp(float v)
{
    float h;
    if (v < log(FLT_MAX / 1000.0)) h = exp(v);
    else h = exp(log(FLT_MAX / 1000.0));
    return h;
}

int ch_model_valid(int dim, const float *z, const float *vp, const float *vpvs, float dz, float zmin, float zmax,
                   float inv_control)
{
    float zz[CH_MAXDIM], bd[CH_MAXDIM], p[CH_MAXDIM], s[CH_MAXDIM];
    float thin = FLT_MAX, tmp, th;
    int i, swapped, lvz = 0;
    if (dim == 1) return 0;
    for (i = 0; i < dim; i++) { zz[i] = z[i]; p[i] = vp[i]; s[i] = p[i] / vpvs[i]; }
    do { /* bubble sort by depth, carrying the velocities */
        swapped = 0;
        for (i = 1; i < dim; i++)
            if (zz[i - 1] > zz[i]) {
                tmp = zz[i - 1]; zz[i - 1] = zz[i]; zz[i] = tmp;
                tmp = p[i - 1]; p[i - 1] = p[i]; p[i] = tmp;
                tmp = s[i - 1]; s[i - 1] = s[i]; s[i] = tmp;
                swapped = 1;
            }
    } while (swapped);
    for (i = 0; i < dim - 1; i++) bd[i] = (zz[i] + zz[i + 1]) / 2.0;
    bd[dim - 1] = zmax;
    for (i = 0; i < dim; i++) {
        th = (i == 0) ? bd[0] - zmin : bd[i] - bd[i - 1];
        if (th < thin) thin = th;
    }
    for (i = 0; i < dim - 1; i++) if (p[i] > p[i + 1]) lvz++;
    for (i = 0; i < dim - 1; i++) if (s[i] > s[i + 1]) lvz++;
    if (thin < (sqrt(inv_control * inv_control) * dz)) return 1;
    if (inv_control < 0 && lvz > 0) return 1;
    return 0;
}

double ch_misfit(const float *mf, const float *n)
{
    double m;
    m = mf[0] / n[0] / n[0] + mf[1] / n[1] / n[1];
    m = m + mf[2] / n[2] / n[2] + mf[3] / n[3] / n[3];
    m = m + mf[4] / n[4] / n[4] + mf[5] / n[5] / n[5];
    m = m + mf[6] / n[6] / n[6] + mf[7] / n[7] / n[7];
    return m;
}

double ch_rms(const float *mf, int sum_of_picks)
{   /* reference order: mfp0+mfp1+mfp2+mfp3+mfs0+mfs1+mfs2+mfs3 */
    return sqrt((mf[0] + mf[2] + mf[4] + mf[6] + mf[1] + mf[3] + mf[5] + mf[7]) / sum_of_picks);
}

float ch_alpha(double log_fac, double new_ll, double old_ll)
{
    const float e = ch_nexp(log_fac + new_ll - old_ll);
    float a = (1.0 < e) ? 1.0 : e;
    return a;
}

double ch_logfac_birth(float sdevvp, float vpmin, float vpmax, float vp_new, float vp_parent, float sdevvpvs,
                       float vpvsmin, float vpvsmax, float vpvs_new, float vpvs_parent)
{
    double lf = log(sdevvp * sqrt(2.0 * CH_PI) / (vpmax - vpmin)) +
                (vp_new - vp_parent) * (vp_new - vp_parent) / 2.0 / sdevvp / sdevvp;
    if (sdevvpvs != 0)
        lf = lf + log(sdevvpvs * sqrt(2.0 * CH_PI) / (vpvsmax - vpvsmin)) +
             (vpvs_new - vpvs_parent) * (vpvs_new - vpvs_parent) / 2.0 / sdevvpvs / sdevvpvs;
    return lf;
}

double ch_logfac_death(float sdevvp, float vpmin, float vpmax, float vp_dead, float vp_nb, float sdevvpvs,
                       float vpvsmin, float vpvsmax, float vpvs_dead, float vpvs_nb)
{
    double lf = log((vpmax - vpmin) / sdevvp / sqrt(2.0 * CH_PI)) -
                (vp_dead - vp_nb) * (vp_dead - vp_nb) / 2.0 / sdevvp / sdevvp;
    if (sdevvpvs != 0)
        lf = lf + log((vpvsmax - vpvsmin) / sdevvpvs / sqrt(2.0 * CH_PI)) -
             (vpvs_dead - vpvs_nb) * (vpvs_dead - vpvs_nb) / 2.0 / sdevvpvs / sdevvpvs;
    return lf;
}

double ch_logfac_noise(const int *nc, const float *o, const float *n)
{
    double lf;
    lf = nc[0] * log(o[0] / n[0]) + nc[1] * log(o[1] / n[1]);
    lf = lf + nc[2] * log(o[2] / n[2]) + nc[3] * log(o[3] / n[3]);
    lf = lf + nc[4] * log(o[4] / n[4]) + nc[5] * log(o[5] / n[5]);
    lf = lf + nc[6] * log(o[6] / n[6]) + nc[7] * log(o[7] / n[7]);
    return lf;
}
