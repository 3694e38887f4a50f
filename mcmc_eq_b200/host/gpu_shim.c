/* gpu_shim.c -> libmcmceq_shim.so: the reference's FUNCTION seam on top of libmcmceq_b200.so.
 *
 * Exports the four functions the reference's hot path is made of, with the reference's own signatures, so that the
 * unmodified chain driver (src/mcmc_eq.c), fw (src/fw.c) and fw_mod (src/fw_mod.c) can be linked against the GPU
 * library instead of their CPU code:
 *
 *   time_2d          src/fdtimes.h:6-7, src/time_2d.c:301    one eikonal solve                 -> mq_time_2d
 *   setup_table_new  src/misfit.c:165                         table of one phase                -> mq_get_table
 *   cal_fit_newx     src/misfit.c:45                          tables (by calct) + misfit        -> mq_forward_host
 *   traveltimet      src/interpol.c:43                        one bilinear lookup               -> mq_traveltimet
 *
 * Two ways to use it (oracle/Makefile builds both from the reference sources where they lie, as test binaries):
 *   (a) replace time_2d.o only:   gcc -Isrc src/mcmc_eq.c src/mod_grd.c -lmcmceq_shim            (misfit.c, interpol.c stay)
 *   (b) replace all four: the reference textually includes "interpol.c" and "misfit.c" (src/mcmc_eq.c:77-78); compiling
 *       it through a link in a directory that holds two empty files of those names leaves the four symbols undefined,
 *       and this library supplies them.
 *
 * State the reference keeps in its callers and this shim has to honour:
 *   * the travel-time tables are caller-owned host arrays (float ***, nz x nz rows of nxmod floats, src/mcmc_eq.c:525-528)
 *     that the chain driver backs up and restores by copying (src/mcmc_eq.c:856,1161,1171).  In mode (b) the tables
 *     live on the device; cal_fit_newx writes a generation stamp into element [0][0][0] of the host array whenever it
 *     rebuilds a table, the driver's copies carry the stamp along, and the next call recognises from it which of the two
 *     device versions (current / previous) the host array stands for -- the device-side twin of the backup and restore.
 *     setup_table_new fills the whole host table with real values (mq_get_table) for callers that read it themselves.
 *   * TRIA, a global of the main program (src/mc.h:59): read through a weak reference.
 *   * errors: the reference prints and exit(0)s (src/misfit.c:93,185); so does this shim, the C ABI underneath returns codes.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mcmceq_b200.h"
#include "ref_abi.h"

extern int TRIA __attribute__((weak));

static mq_handle* g_h;              /* one chain, created at the first cal_fit_newx call */
static mq_picks g_pk;
static int g_ne, g_np, g_ns;
static struct GRDHEAD g_gh;
static int g_eikonal = -1;
static float g_gen = 1.0f;          /* generation counter of table rebuilds (exact in a float up to 2^24) */
static float g_dev[2], g_old[2];    /* stamp of the table in device buffer 0 (current) / 1 (previous), per phase */
static int g_device;

static void die(const char* what)
{
    fprintf(stderr, "%s: %s\n", what, mq_last_error());
    exit(0);                        /* the reference's error convention, e.g. src/misfit.c:93 */
}

static int same_grid(const struct GRDHEAD* a, const struct GRDHEAD* b)
{
    return a->nx == b->nx && a->ny == b->ny && a->nz == b->nz && a->h == b->h && a->x0 == b->x0 && a->y0 == b->y0 && a->z0 == b->z0;
}

static void fill_config(mq_config* c, const struct GRDHEAD* gh, int eikonal)
{
    memset(c, 0, sizeof *c);
    c->grid.h = gh->h; c->grid.nx = gh->nx; c->grid.ny = gh->ny; c->grid.nz = gh->nz;
    c->grid.x0 = gh->x0; c->grid.y0 = gh->y0; c->grid.z0 = gh->z0;
    c->max_dim = REF_MD;
    c->eikonal = eikonal;
    c->tria = (&TRIA != NULL) ? TRIA : 0;
    c->deci = 1 << 30;
    strcpy(c->dstring_start, "Q"); strcpy(c->dstring_main, "Q");
}

/* struct DATA[ne] -> flattened picks (P picks first, then S picks, file order: the order src/misfit.c:87-119 sums in) */
static void build_handle(const struct Model* m, const struct DATA* d, int ne, const struct GRDHEAD* gh, int eikonal)
{
    int e, j, np = 0, k = 0;
    int32_t *ev_off, *n_p, *st_id, *cls;
    float *x, *y, *z, *t;
    double* reftime;
    mq_config cfg;
    if (g_h) { mq_destroy(g_h); g_h = NULL; }
    for (e = 0; e < ne; e++) np += d[e].nobs_p + d[e].nobs_s;
    ev_off = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ne + 1)); n_p = (int32_t*)malloc(sizeof(int32_t) * (size_t)ne);
    st_id = (int32_t*)malloc(sizeof(int32_t) * (size_t)np); cls = (int32_t*)malloc(sizeof(int32_t) * (size_t)np);
    x = (float*)malloc(sizeof(float) * (size_t)np); y = (float*)malloc(sizeof(float) * (size_t)np);
    z = (float*)malloc(sizeof(float) * (size_t)np); t = (float*)malloc(sizeof(float) * (size_t)np);
    reftime = (double*)malloc(sizeof(double) * (size_t)ne);
    if (!ev_off || !n_p || !st_id || !cls || !x || !y || !z || !t || !reftime) { fprintf(stderr, "Memory allocation error\n"); exit(0); }
    for (e = 0; e < ne; e++) {
        ev_off[e] = k;
        n_p[e] = d[e].nobs_p;
        reftime[e] = d[e].reftime;
        for (j = 0; j < d[e].nobs_p + d[e].nobs_s; j++, k++) {
            const struct OBS* o = j < d[e].nobs_p ? &d[e].p_picks[j] : &d[e].s_picks[j - d[e].nobs_p];
            st_id[k] = o->st_id; cls[k] = o->cl; x[k] = o->x; y[k] = o->y; z[k] = o->z; t[k] = o->t;
        }
    }
    ev_off[ne] = k;
    memset(&g_pk, 0, sizeof g_pk);
    g_pk.n_events = ne; g_pk.n_picks = np; g_pk.n_stations = (int)m->nos;
    g_pk.ev_off = ev_off; g_pk.n_p = n_p; g_pk.st_id = st_id; g_pk.x = x; g_pk.y = y; g_pk.z = z; g_pk.t = t; g_pk.cls = cls;
    g_pk.reftime = reftime; g_pk.fix = NULL;
    fill_config(&cfg, gh, eikonal);
    if (getenv("MCMCEQ_DEVICE")) g_device = atoi(getenv("MCMCEQ_DEVICE"));
    if (mq_create(&cfg, &g_pk, 1, g_device, 1, &g_h) != MQ_OK) die("mq_create");
    g_ne = ne; g_np = np; g_ns = (int)m->nos; g_gh = *gh; g_eikonal = eikonal;
    g_dev[0] = g_dev[1] = g_old[0] = g_old[1] = 0.f;
    /* the pick arrays are copied to the device by mq_create */
    free(ev_off); free(n_p); free(st_id); free(cls); free(x); free(y); free(z); free(t); free(reftime);
}

static void model_view(struct Model* m, int32_t* dim32, mq_models* v)
{
    *dim32 = (int32_t)m->dimension;
    memset(v, 0, sizeof *v);
    v->n_chains = 1; v->max_dim = REF_MD; v->n_events = g_ne; v->n_stations = g_ns;
    v->dim = dim32; v->z = m->z; v->vp = m->vp; v->vpvs = m->vpvs;
    v->eq = &m->eq[0].x;                   /* struct QUAKE is three floats: [noq][3] */
    v->pres = m->pres; v->sres = m->sres; v->noise = m->noise; v->origin = m->origin;
}

/* ---- time_2d ------------------------------------------------------------------------------------------------ */
int time_2d(float* hs, float* t, int nx, int ny, float xs, float ys, float eps_init, int messages)
{
    const int rc = mq_time_2d(hs, t, nx, ny, xs, ys, eps_init, messages);
    if (rc == MQ_OK) return 0;
    fprintf(stderr, "time_2d (GPU): %s\n", mq_last_error());
    return rc == MQ_ERR_UNSUPPORTED ? -1 : -2;      /* the reference returns negative codes and goes on (src/time_2d.c:263-271) */
}

/* ---- traveltimet -------------------------------------------------------------------------------------------- */
float traveltimet(float** ttt, int nx, int ny, int nz, float h, float dist, float z, float z0)
{
    float v = 1e30f;
    if (mq_traveltimet(ttt, nx, ny, nz, h, dist, z, z0, &v, g_device) != MQ_OK) die("traveltimet");
    return v;
}

/* ---- setup_table_new ---------------------------------------------------------------------------------------- */
void setup_table_new(struct Model* m, float*** ttt, struct GRDHEAD gh, int ps)
{
    static mq_handle* th;           /* a handle of its own: one dummy pick, the table does not depend on the picks */
    static struct GRDHEAD tgh;
    const int nxmod = (int)sqrt((double)(gh.nx * gh.nx + gh.ny * gh.ny));
    int32_t dim32;
    mq_models v;
    float* flat;
    int j, k;
    if (!th || !same_grid(&tgh, &gh)) {
        const int32_t ev_off[2] = {0, 1}, n_p[1] = {1}, st_id[1] = {0}, cls[1] = {0};
        const float x[1] = {gh.x0}, y[1] = {gh.y0}, z[1] = {gh.z0}, t[1] = {0.f};
        mq_picks pk;
        mq_config cfg;
        if (th) mq_destroy(th);
        memset(&pk, 0, sizeof pk);
        pk.n_events = 1; pk.n_picks = 1; pk.n_stations = (int)(m->nos > 0 ? m->nos : 1);
        pk.ev_off = ev_off; pk.n_p = n_p; pk.st_id = st_id; pk.x = x; pk.y = y; pk.z = z; pk.t = t; pk.cls = cls;
        fill_config(&cfg, &gh, 1);
        if (mq_create(&cfg, &pk, 1, g_device, 1, &th) != MQ_OK) die("setup_table_new");
        tgh = gh;
    }
    dim32 = (int32_t)m->dimension;
    memset(&v, 0, sizeof v);
    v.n_chains = 1; v.max_dim = REF_MD; v.n_events = 1; v.n_stations = (int)(m->nos > 0 ? m->nos : 1);
    v.dim = &dim32; v.z = m->z; v.vp = m->vp; v.vpvs = m->vpvs; v.eq = &m->eq[0].x; v.pres = m->pres; v.sres = m->sres;
    v.noise = m->noise; v.origin = NULL;
    if (mq_set_models(th, &v) != MQ_OK) die("setup_table_new");
    flat = (float*)malloc(sizeof(float) * (size_t)gh.nz * gh.nz * nxmod);
    if (!flat) { fprintf(stderr, "Memory allocation error\n"); exit(0); }      /* src/misfit.c:185 */
    if (mq_get_table(th, 0, ps, flat) != MQ_OK) die("setup_table_new");
    for (j = 0; j < gh.nz; j++)
        for (k = 0; k < gh.nz; k++) memcpy(ttt[j][k], flat + ((size_t)j * gh.nz + k) * nxmod, sizeof(float) * (size_t)nxmod);
    free(flat);
}

/* ---- cal_fit_newx ------------------------------------------------------------------------------------------- */
/* which device version does the host table stand for?  (see the file comment) */
static void select_version(float*** ttt, int ph, int rebuilt)
{
    const int mask = 1 << ph;
    if (rebuilt) {
        /* The table the caller holds becomes the previous version.  It is device buffer 0 unless the caller has just
         * restored its backup after a rejection: then it already sits in buffer 1 and buffer 0 (the rejected table) is
         * simply overwritten. */
        const float cur = ttt[0][0][0];
        if (g_dev[ph] != 0.f && cur == g_dev[ph]) {
            if (mq_tables_save_phases(g_h, mask) != MQ_OK) die("cal_fit_newx");
            g_old[ph] = g_dev[ph];
        }
        g_gen += 1.0f;
        g_dev[ph] = g_gen;
        ttt[0][0][0] = g_gen;
        return;
    }
    {
        const float want = ttt[0][0][0];
        if (want == g_dev[ph]) return;
        if (want == g_old[ph] && want != 0.f) {          /* the driver restored its backup (src/mcmc_eq.c:1171) */
            if (mq_tables_restore_phases(g_h, mask) != MQ_OK) die("cal_fit_newx");
            g_dev[ph] = want;
            return;
        }
        fprintf(stderr, "cal_fit_newx (GPU): the %c table passed in is not one this library built\n", ph ? 'S' : 'P');
        exit(0);
    }
}

float cal_fit_newx(struct Model* m, struct DATA* d, int ne, float*** tttp, float*** ttts, struct GRDHEAD gh, int calct, float* mfp0,
                   float* mfs0, float* mfp1, float* mfs1, float* mfp2, float* mfs2, float* mfp3, float* mfs3, int flag, int eikonal,
                   int out)
{
    float mf[8];
    int32_t dim32;
    mq_models v;
    int rc;
    *mfp0 = *mfs0 = *mfp1 = *mfs1 = *mfp2 = *mfs2 = *mfp3 = *mfs3 = 0.f;
    if (flag == 1) return 1.0f;                                            /* prior only, src/misfit.c:61 */
    if (!g_h || ne != g_ne || (int)m->nos != g_ns || !same_grid(&g_gh, &gh) || eikonal != g_eikonal) build_handle(m, d, ne, &gh, eikonal);
    if (eikonal == 1) {
        select_version(tttp, 0, (calct & 1) != 0);
        select_version(ttts, 1, (calct & 2) != 0);
    }
    model_view(m, &dim32, &v);
    rc = mq_forward_host(g_h, &v, eikonal == 1 ? calct : 0, mf, m->origin);
    if (rc == MQ_ERR_STATCOR) { fprintf(stderr, "ERROR points to invalid station correction\n"); exit(0); }   /* src/misfit.c:93 */
    if (rc != MQ_OK) die("cal_fit_newx");
    *mfp0 = mf[0]; *mfs0 = mf[1]; *mfp1 = mf[2]; *mfs1 = mf[3]; *mfp2 = mf[4]; *mfs2 = mf[5]; *mfp3 = mf[6]; *mfs3 = mf[7];
    if (out == 1) {                                                        /* per-pick lines, src/misfit.c:130-143 */
        float* resid = (float*)malloc(sizeof(float) * (size_t)g_np);
        float* tpred = (float*)malloc(sizeof(float) * (size_t)g_np);
        int e, j, k = 0;
        if (!resid || !tpred || mq_get_predictions(g_h, 0, resid, tpred) != MQ_OK) die("cal_fit_newx");
        for (e = 0; e < ne; e++) {
            fprintf(stdout, "EVENT %d  %lf %f %f %f %f\n", e, d[e].reftime, m->eq[e].x, m->eq[e].y, m->eq[e].z, m->origin[e]);
            for (j = 0; j < d[e].nobs_p + d[e].nobs_s; j++, k++) {
                const struct OBS* o = j < d[e].nobs_p ? &d[e].p_picks[j] : &d[e].s_picks[j - d[e].nobs_p];
                const float dx = o->x - m->eq[e].x, dy = o->y - m->eq[e].y;
                fprintf(stdout, "%f %f %f %f %f %f %c\n", resid[k], (float)sqrt(dx * dx + dy * dy), m->eq[e].z, m->origin[e], o->t, tpred[k],
                        j < d[e].nobs_p ? 'P' : 'S');
            }
        }
        free(resid); free(tpred);
    }
    return 1.0f;
}
