/* oracle/fwd_model.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  See fwd_model.h.
 * Every expression keeps the reference's operand types (float arithmetic, double only
 * where the reference promotes) so that results are bit-identical; build with
 * -ffp-contract=off. */
#include "fwd_model.h"
#include "pl_eikonal.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>

int fm_nxmod(const fm_grid *g) { return (int)sqrt(g->nx * g->nx + g->ny * g->ny); }

int fm_find_in_cell(const float *z, int dim, float zq)
{
    int i, j = 0;
    float best = FLT_MAX;
    for (i = 0; i < dim; i++)
        if ((z[i] - zq) * (z[i] - zq) <= best) { /* ties -> highest index */
            best = (z[i] - zq) * (z[i] - zq);
            j = i;
        }
    return j;
}

int fm_find_neighbor_cell(const float *z, int dim, int n)
{
    int i, j = 0;
    float best = FLT_MAX;
    for (i = 0; i < dim; i++)
        if (i != n && (z[i] - z[n]) * (z[i] - z[n]) <= best) {
            best = (z[i] - z[n]) * (z[i] - z[n]);
            j = i;
        }
    return j;
}

float fm_dst(float x1, float x2, float y1, float y2)
{
    return (sqrt(((x1 - x2) * (x1 - x2)) + ((y1 - y2) * (y1 - y2))));
}

void fm_receiver(const fm_grid *g, float z, int *layer, float *w1, float *w2)
{
    *layer = (int)((z - g->z0) / g->h);
    *w2 = -(*layer * g->h + g->z0 - z) / g->h;
    *w1 = 1.0 - *w2;
}

void fm_rasterise(const fm_grid *g, int dim, const float *z, const float *vp, const float *vpvs, int ps, float *slow)
{
    int iz;
    for (iz = 0; iz < g->nz; iz++) {
        const float zq = g->z0 + (float)iz * g->h;
        const int k = fm_find_in_cell(z, dim, zq);
        const float p = vp[k];
        const float s = p / vpvs[k];
        slow[iz] = g->h / ((ps == 1) ? p : s);
    }
}

/* TRIA = 1 (config line 29): linear interpolation between the depth-sorted nuclei, src/misfit.c:217-250.  The sort is a
 * stable one (the reference's bubble sort swaps on strict > only); the segment index of a depth that no segment holds
 * is the one of the previous depth node (the reference's k is not reset inside the depth loop, :233). */
void fm_rasterise_tria(const fm_grid *g, int dim, const float *z, const float *vp, const float *vpvs, int ps, float *slow)
{
    float *zs = (float *)malloc(3 * (size_t)(dim > 0 ? dim : 1) * sizeof(float));
    float *ps_ = zs + dim, *rs = ps_ + dim;
    int i, j, iz, k = 0;
    for (i = 0; i < dim; i++) {           /* insertion sort, stable */
        const float kz = z[i], kp = vp[i], kr = vpvs[i];
        for (j = i - 1; j >= 0 && zs[j] > kz; j--) { zs[j + 1] = zs[j]; ps_[j + 1] = ps_[j]; rs[j + 1] = rs[j]; }
        zs[j + 1] = kz; ps_[j + 1] = kp; rs[j + 1] = kr;
    }
    for (iz = 0; iz < g->nz; iz++) {
        const float zq = g->z0 + (float)iz * g->h;
        float a, b, v;
        for (i = 0; i < dim - 1; i++)
            if (zq >= zs[i] && zq < zs[i + 1]) k = i;
        if (ps == 1) {
            a = (ps_[k + 1] - ps_[k]) / (zs[k + 1] - zs[k]);
            b = ps_[k] - a * zs[k];
        } else {
            a = (ps_[k + 1] / rs[k + 1] - ps_[k] / rs[k]) / (zs[k + 1] - zs[k]);
            b = ps_[k] / rs[k] - a * zs[k];
        }
        v = a * zq + b;
        slow[iz] = g->h / v;
    }
    free(zs);
}

int fm_build_table(const fm_grid *g, const float *slow, float *ttt)
{
    const int nxmod = fm_nxmod(g), nz = g->nz;
    float *hs = (float *)malloc((size_t)nxmod * nz * sizeof(float));
    float *t = (float *)malloc((size_t)nxmod * nz * sizeof(float));
    int ix, iz, j, worst = 0;
    if (!hs || !t) { free(hs); free(t); return PL_ERR_ALLOC; }
    for (ix = 0; ix < nxmod; ix++)
        for (iz = 0; iz < nz; iz++) hs[(size_t)ix * nz + iz] = slow[iz];
    for (iz = 0; iz < nz; iz++) {
        /* the reference ignores time_2d's return code (src/misfit.c:278); keep the worst one */
        const int rc = pl_time_2d(hs, t, nxmod, nz, 0.0f, (float)iz, 0.001f, 0);
        if (rc < worst) worst = rc;
        for (j = 0; j < nz; j++)
            for (ix = 0; ix < nxmod; ix++) ttt[((size_t)j * nz + iz) * nxmod + ix] = t[(size_t)ix * nz + j];
    }
    free(hs);
    free(t);
    return worst;
}

float fm_traveltime(const float *tl, const fm_grid *g, float dist, float z)
{
    const int nxmod = fm_nxmod(g);
    const float h = g->h, z0 = g->z0;
    int m1, iz1, m2, iz2;
    float x1, y1, x2, y2, v, v1, v2, v3, v4, x, y;
    m1 = (int)(dist / h);
    iz1 = (int)((z - z0) / h);
    x = dist;
    y = z - z0;
    if (m1 >= nxmod - 1 || iz1 >= g->nz - 1) return (1e30);
    m2 = m1 + 1;
    iz2 = iz1 + 1;
    x1 = (float)m1 * h;
    y1 = (float)iz1 * h;
    x2 = (float)m2 * h;
    y2 = (float)iz2 * h;
    v1 = tl[(size_t)iz1 * nxmod + m1];
    v2 = tl[(size_t)iz1 * nxmod + m2];
    v3 = tl[(size_t)iz2 * nxmod + m1];
    v4 = tl[(size_t)iz2 * nxmod + m2];
    v = 1.0 / (x2 - x1) / (y2 - y1) *
        (v1 * (x2 - x) * (y2 - y) + v2 * (x - x1) * (y2 - y) + v3 * (x2 - x) * (y - y1) + v4 * (x - x1) * (y - y1));
    return (v);
}

int fm_misfit(const fm_grid *g, const fm_picks *p, const float *eq, const float *pres, const float *sres,
              const float *tttp, const float *ttts, int eikonal, int dim, const float *z, const float *vp,
              const float *vpvs, float *mf, float *origin, float *resid, float *tpred)
{
    const int nxmod = fm_nxmod(g);
    const size_t lstride = (size_t)g->nz * nxmod;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int e, j, c;
    float *diff = (float *)malloc((size_t)(p->n_picks > 0 ? p->n_picks : 1) * sizeof(float));
    if (!diff) return PL_ERR_ALLOC;
    for (e = 0; e < p->n_events; e++) {
        const int b = p->ev_off[e], end = p->ev_off[e + 1], np = p->n_p[e];
        const float ex = eq[3 * e], ey = eq[3 * e + 1], ez = eq[3 * e + 2];
        float sum = 0;
        for (j = b; j < end; j++) {
            const int isS = (j - b) >= np;
            const float *ttt = isS ? ttts : tttp;
            const float dist = fm_dst(p->x[j], ex, p->y[j], ey);
            float tt = 0, w1, w2, corr;
            int layer;
            fm_receiver(g, p->z[j], &layer, &w1, &w2);
            if (eikonal == 0) {
                const int k0 = fm_find_in_cell(z, dim, 0.0);
                if (!isS) tt = sqrt(dist * dist + ez * ez) / vp[k0];
                else tt = sqrt(dist * dist + ez * ez) / (vp[k0] / vpvs[k0]);
            } else {
                tt = fm_traveltime(ttt + (size_t)layer * lstride, g, dist, ez) * w1 +
                     fm_traveltime(ttt + (size_t)(layer + 1) * lstride, g, dist, ez) * w2;
            }
            corr = isS ? sres[p->st_id[j]] : pres[p->st_id[j]];
            if (corr < -1000) { free(diff); return -5; }
            tt += corr;
            diff[j] = tt - p->t[j];
            if (tpred) tpred[j] = tt;
        }
        /* P sum first, then S, as src/misfit.c:101-119 (same order as the storage order) */
        for (j = b; j < end; j++) sum = sum + diff[j];
        sum = sum / (end - b);
        origin[e] = -sum;
        for (j = b; j < end; j++) diff[j] = diff[j] - sum;
        /* class by class, P before S inside a class (src/misfit.c:146-153) */
        for (c = 0; c < 4; c++) {
            for (j = b; j < b + np; j++) if (p->cls[j] == c) acc[2 * c] = acc[2 * c] + diff[j] * diff[j];
            for (j = b + np; j < end; j++) if (p->cls[j] == c) acc[2 * c + 1] = acc[2 * c + 1] + diff[j] * diff[j];
        }
        if (resid) for (j = b; j < end; j++) resid[j] = diff[j];
    }
    for (c = 0; c < 8; c++) mf[c] = acc[c];
    free(diff);
    return 0;
}
