/* oracle/fwd_model.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * CPU restatement of mcmc_eq's forward model on flat arrays: Voronoi rasterisation,
 * travel-time table build, bilinear table lookup, residuals / origin time / class sums.
 * Reference: src/misfit.c:45-293, src/interpol.c:43-83, src/mod_grd.c:72-110,
 * src/mcmc_eq.c:503-517 (receiver weights), :1303-1306 (dst).
 * Pinned bit-for-bit against oracle/_ref by tests/test_oracle_pin.py.
 */
#ifndef ORACLE_FWD_MODEL_H
#define ORACLE_FWD_MODEL_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float h; int nx, ny, nz; float x0, y0, z0; } fm_grid; /* struct GRDHEAD, src/mc.h:91-100 */

/* Flattened struct DATA / struct OBS (src/mc.h:102-134): picks of event e are
 * [ev_off[e], ev_off[e+1]), the first n_p[e] of them P, the rest S, file order. */
typedef struct {
    int n_events, n_picks;
    const int *ev_off, *n_p, *st_id, *cls;
    const float *x, *y, *z, *t;
} fm_picks;

int fm_nxmod(const fm_grid *g);                                        /* src/mcmc_eq.c:520 */
int fm_find_in_cell(const float *z, int dim, float zq);                /* src/mod_grd.c:93-110 */
int fm_find_neighbor_cell(const float *z, int dim, int n);             /* src/mod_grd.c:72-90 */
float fm_dst(float x1, float x2, float y1, float y2);                  /* src/mcmc_eq.c:1303-1306 */
/* receiver layer and elevation weights of a station at depth z, src/mcmc_eq.c:507-509 */
void fm_receiver(const fm_grid *g, float z, int *layer, float *w1, float *w2);
/* slow[iz] = h / v(z0 + iz*h), ps = 1 (P) or 2 (S); src/misfit.c:205-214, 256-266 */
void fm_rasterise(const fm_grid *g, int dim, const float *z, const float *vp, const float *vpvs, int ps, float *slow);
void fm_rasterise_tria(const fm_grid *g, int dim, const float *z, const float *vp, const float *vpvs, int ps, float *slow);
/* ttt[(j*nz + iz)*nxmod + i] = time at distance node i, receiver row j, source row iz; src/misfit.c:270-289 */
int fm_build_table(const fm_grid *g, const float *slow, float *ttt);
/* bilinear lookup in one receiver layer (ttt_layer = &ttt[j*nz*nxmod]); src/interpol.c:43-83 */
float fm_traveltime(const float *ttt_layer, const fm_grid *g, float dist, float z);
/* residual loop of cal_fit_newx (src/misfit.c:83-159).  mf[2*class+phase] (phase 0 = P),
 * origin[n_events]; resid/tpred ([n_picks], may be NULL) are the de-meaned residual and the
 * predicted time incl. station correction.  eikonal = 0 selects the straight-ray branch
 * (then dim/z/vp/vpvs are used and tttp/ttts ignored).  Returns 0, or -5 when a pick points
 * to a station correction < -1000 (the reference exit(0)s there). */
int fm_misfit(const fm_grid *g, const fm_picks *p, const float *eq, const float *pres, const float *sres,
              const float *tttp, const float *ttts, int eikonal, int dim, const float *z, const float *vp,
              const float *vpvs, float *mf, float *origin, float *resid, float *tpred);

#ifdef __cplusplus
}
#endif
#endif
