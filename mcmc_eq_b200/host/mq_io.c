/* mq_io.c -- readers/writers for the reference's file formats.  See mq_io.h. */
#include "mq_io.h"

#include <ctype.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

static char g_ioerr[512];
const char* mqio_last_error(void) { return g_ioerr; }
static int fail(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ioerr, sizeof g_ioerr, fmt, ap);
    va_end(ap);
    return MQ_ERR_ARG;
}

/* One config line: everything up to the newline, blank-padded to at least 80 characters so
 * that a short or missing line scans as "no value" (reference read_single_line, src/mod_grd.c:51-63). */
static void config_line(FILE* fp, char* buf, size_t cap)
{
    size_t i = 0;
    int c;
    while ((c = getc(fp)) != EOF && c != '\n')
        if (i + 1 < cap) buf[i++] = (char)c;
    while (i < 80 && i + 1 < cap) buf[i++] = ' ';
    buf[i] = '\0';
}

int mqio_read_config(const char* path, mq_config* c)
{
    FILE* fp = fopen(path, "r");
    char line[1024], dummy[1000];
    int idummy;
    if (!fp) return fail("could not open config file %s", path);
    memset(c, 0, sizeof *c);
#define L(...) do { config_line(fp, line, sizeof line); sscanf(line, __VA_ARGS__); } while (0)
    L("%f ", &c->grid.h);                      /*  1 forward dx          */
    L("%d ", &c->grid.nx);                     /*  2 */
    L("%d ", &c->grid.ny);                     /*  3 */
    L("%d ", &c->grid.nz);                     /*  4 */
    L("%f ", &c->grid.x0);                     /*  5 */
    L("%f ", &c->grid.y0);                     /*  6 */
    L("%f ", &c->grid.z0);                     /*  7 */
    L("%d ", &c->max_dim);                     /*  8 */
    L("%f ", &c->vpmin);                       /*  9 */
    L("%f ", &c->vpmax);                       /* 10 */
    L("%f ", &c->vpvsmin);                     /* 11 */
    L("%f ", &c->vpvsmax);                     /* 12 */
    L("%f ", &c->noise_min);                   /* 13 */
    L("%f ", &c->noise_max);                   /* 14 */
    L("%f ", &c->residual_min);                /* 15 */
    L("%f ", &c->residual_max);                /* 16 */
    L("%f ", &c->sdevx);                       /* 17 */
    L("%f ", &c->sdevy);                       /* 18 */
    L("%f ", &c->sdevz);                       /* 19 */
    L("%f ", &c->sdevvp);                      /* 20 */
    L("%f ", &c->sdevvpvs);                    /* 21 */
    L("%f ", &c->sdevn);                       /* 22 */
    L("%f %f ", &c->sdevxs, &c->epi_search);   /* 23 */
    L("%f ", &c->sdevys);                      /* 24 */
    L("%f ", &c->sdevzs);                      /* 25 */
    L("%f ", &c->sdevresidual);                /* 26 */
    L("%f ", &c->inv_control);                 /* 27 */
    L("%d %d %f %f ", &c->reference_station, &c->scor_flag, &c->ref_statcor_P, &c->ref_statcor_S); /* 28 */
    L("%d ", &c->tria);                        /* 29 */
    L("%d %d ", &c->j_max_start, &c->j_max_main); /* 30 */
    L("%d ", &c->deci);                        /* 31 */
    L("%d %d", &c->true_random, &c->eikonal);  /* 32 */
    L("%63s %63s", c->dstring_start, c->dstring_main); /* 33 */
    L("%d %15s ", &c->aflag, c->inp_model_switch);     /* 34 */
    L("%d %999s %d ", &idummy, dummy, &idummy);        /* 35 topo line, unused by the reference too */
    L("%f %f %f ", &c->start_vp, &c->sdev_start_vp, &c->start_vp_grad); /* 36 */
    L("%f %f ", &c->start_vpvs, &c->sdev_start_vpvs);  /* 37 */
    L("%d %d ", &c->start_cell_number, &c->sdev_start_cell_number); /* 38 */
    L("%f ", &c->start_noise);                 /* 39 */
    L("%f %f ", &c->start_delay, &c->sdev_start_delay); /* 40 */
    L("%f %f ", &c->r_start_eqh, &c->r_start_eqv);      /* 41 */
#undef L
    fclose(fp);
    if (c->inv_control == 0.0f) return fail("inv_control should be != 0!");   /* src/mcmc_eq.c:373 */
    if (c->grid.nx < 1 || c->grid.ny < 1 || c->grid.nz < 2 || !(c->grid.h > 0))
        return fail("bad grid in %s: h=%g nx=%d ny=%d nz=%d", path, c->grid.h, c->grid.nx, c->grid.ny, c->grid.nz);
    return MQ_OK;
}

void mqio_free_picks(mqio_picks* p)
{
    if (!p) return;
    free(p->ev_off); free(p->n_p); free(p->st_id); free(p->cls); free(p->eq_id);
    free(p->x); free(p->y); free(p->z); free(p->t); free(p->reftime); free(p->fix);
    memset(p, 0, sizeof *p);
}

typedef struct { int st_id, cls, isS; float x, y, z, t; } raw_pick;

int mqio_read_picks(const char* path, mqio_picks* out)
{
    FILE* f = fopen(path, "r");
    char buf[4096], name[64], ph[64];
    raw_pick* ev = NULL;      /* picks of the event being read */
    size_t ev_n = 0, ev_cap = 0;
    size_t cap_p = 0, cap_e = 0;
    int n_events = 0, n_picks = 0, have_event = 0, hdr_p = 0, hdr_s = 0, max_st = -1;
    int rc = MQ_OK;
    memset(out, 0, sizeof *out);
    if (!f) return fail("could not open data file %s", path);

#define GROW_E() do { if ((size_t)n_events + 2 > cap_e) { cap_e = cap_e ? cap_e * 2 : 256; \
        out->ev_off = realloc(out->ev_off, (cap_e + 1) * sizeof(int32_t)); out->n_p = realloc(out->n_p, cap_e * sizeof(int32_t)); \
        out->eq_id = realloc(out->eq_id, cap_e * sizeof(int32_t)); out->reftime = realloc(out->reftime, cap_e * sizeof(double)); \
        out->fix = realloc(out->fix, cap_e * 3 * sizeof(double)); } } while (0)
#define GROW_P(n) do { if ((size_t)n_picks + (n) > cap_p) { while ((size_t)n_picks + (n) > cap_p) cap_p = cap_p ? cap_p * 2 : 4096; \
        out->st_id = realloc(out->st_id, cap_p * sizeof(int32_t)); out->cls = realloc(out->cls, cap_p * sizeof(int32_t)); \
        out->x = realloc(out->x, cap_p * sizeof(float)); out->y = realloc(out->y, cap_p * sizeof(float)); \
        out->z = realloc(out->z, cap_p * sizeof(float)); out->t = realloc(out->t, cap_p * sizeof(float)); } } while (0)

    /* close the current event: P picks first, then S picks, each in file order */
#define FLUSH_EVENT() do { if (have_event) { size_t q_; int np_ = 0, ns_ = 0, pass_; \
        for (q_ = 0; q_ < ev_n; q_++) { if (ev[q_].isS) ns_++; else np_++; } \
        if (np_ != hdr_p || ns_ != hdr_s) { rc = fail("event %d: header announces %d P / %d S picks, file has %d / %d", \
                                                      n_events - 1, hdr_p, hdr_s, np_, ns_); goto done; } \
        GROW_P(ev_n); \
        for (pass_ = 0; pass_ < 2; pass_++) for (q_ = 0; q_ < ev_n; q_++) if (ev[q_].isS == pass_) { \
            out->st_id[n_picks] = ev[q_].st_id; out->cls[n_picks] = ev[q_].cls; out->x[n_picks] = ev[q_].x; \
            out->y[n_picks] = ev[q_].y; out->z[n_picks] = ev[q_].z; out->t[n_picks] = ev[q_].t; \
            out->n_class[2 * ev[q_].cls + pass_]++; n_picks++; } \
        out->n_p[n_events - 1] = np_; out->ev_off[n_events] = n_picks; ev_n = 0; } } while (0)

    while (fgets(buf, sizeof buf, f)) {
        const char* s = buf;
        while (*s && isspace((unsigned char)*s)) s++;
        if (!*s) continue; /* blank line */
        if (strchr(buf, '#')) {
            int id = 0, np = 0, ns = 0;
            double ref = 0, fx = -9999.0, fy = -9999.0, fz = -9999.0;
            FLUSH_EVENT();
            sscanf(buf, "%63s %d %d %d %lf %lf %lf %lf", name, &id, &np, &ns, &ref, &fx, &fy, &fz);
            GROW_E();
            if (n_events == 0) out->ev_off[0] = 0;
            out->eq_id[n_events] = id;
            out->reftime[n_events] = ref;
            out->fix[3 * n_events] = fx; out->fix[3 * n_events + 1] = fy; out->fix[3 * n_events + 2] = fz;
            hdr_p = np; hdr_s = ns;
            n_events++;
            have_event = 1;
        } else {
            int st = 0, cl = 0;
            float x = 0, y = 0, z = 0;
            double t = 0;
            if (!have_event) { rc = fail("%s: pick line before the first '#' header", path); goto done; }
            ph[0] = 0;
            if (sscanf(buf, "%63s %d %63s %f %f %f %lf %d", name, &st, ph, &x, &y, &z, &t, &cl) < 8) {
                rc = fail("%s: malformed pick line: %.60s", path, buf); goto done;
            }
            if (cl > 3 || cl < 0) { rc = fail("pick class to large! (class %d)", cl); goto done; }
            if (st < 0) { rc = fail("negative station id %d", st); goto done; }
            if (ev_n == ev_cap) { ev_cap = ev_cap ? ev_cap * 2 : 256; ev = realloc(ev, ev_cap * sizeof *ev); }
            ev[ev_n].st_id = st; ev[ev_n].cls = cl; ev[ev_n].isS = (strchr(ph, 'P') == NULL);
            ev[ev_n].x = x; ev[ev_n].y = y; ev[ev_n].z = z; ev[ev_n].t = (float)t;
            ev_n++;
            if (st > max_st) max_st = st;
        }
    }
    FLUSH_EVENT();
    if (n_events == 0) { rc = fail("%s: no events", path); goto done; }
done:
    fclose(f);
    free(ev);
    if (rc != MQ_OK) { mqio_free_picks(out); return rc; }
    out->view.n_events = n_events;
    out->view.n_picks = n_picks;
    out->view.n_stations = max_st + 1;   /* src/mcmc_eq.c:447-450 */
    out->view.ev_off = out->ev_off; out->view.n_p = out->n_p; out->view.st_id = out->st_id;
    out->view.x = out->x; out->view.y = out->y; out->view.z = out->z; out->view.t = out->t;
    out->view.cls = out->cls; out->view.reftime = out->reftime; out->view.fix = out->fix;
    return MQ_OK;
#undef GROW_E
#undef GROW_P
#undef FLUSH_EVENT
}

int mqio_check_picks(const mq_config* c, const mqio_picks* p, FILE* log)
{
    const mq_grid* g = &c->grid;
    const float xmin = g->x0, xmax = g->x0 + (g->nx - 1) * g->h;
    const float ymin = g->y0, ymax = g->y0 + (g->ny - 1) * g->h;
    const float zmin = g->z0, zmax = g->z0 + (g->nz - 1) * g->h;
    const int n = p->view.n_picks, nos = p->view.n_stations;
    int i, k;
    char* seen = calloc((size_t)(nos > 0 ? nos : 1), 1);
    for (i = 0; i < n; i++) seen[p->st_id[i]] = 1;
    for (k = 0; k < nos; k++)
        if (!seen[k] && log) fprintf(log, "WARNING station %d missing in pick file\n", k);
    free(seen);
    for (i = 0; i < p->view.n_events; i++)
        if (p->eq_id[i] != i && log) fprintf(log, "WARNING quakes not correctly sorted in pick file around EQ_ID %d \n", i);
    if (c->reference_station >= nos) return fail("reference station not in data file");
    for (i = 0; i < n; i++) {
        if (p->x[i] < xmin || p->x[i] > xmax) return fail("station x position outside search boundaries %f", p->x[i]);
        if (p->y[i] < ymin || p->y[i] > ymax) return fail("station y position outside search boundaries %f", p->y[i]);
        if (p->z[i] < zmin || p->z[i] > zmax) return fail("station z position outside search boundaries %f", p->z[i]);
    }
    return MQ_OK;
}

void mqio_write_record(FILE* f, const char* tag, const char* code, long number, long dim, double rms,
                       const float* noise, const float* z, const float* vp, const float* vpvs, int n_events,
                       const float* eq, const double* reftime, const float* origin, int n_stations,
                       const float* pres, const float* sres)
{
    int i;
    const float r = (float)rms;   /* print_model_raw takes rms as float (src/mcmc_eq.c:234) */
    /* noise order on the line: p0 p1 p2 p3 s0 s1 s2 s3 (src/mcmc_eq.c:237) */
    fprintf(f, "%3s %2s %8ld %3ld %f %f %f %f %f %f %f %f %f", tag, code, number, dim, r, noise[0], noise[2], noise[4],
            noise[6], noise[1], noise[3], noise[5], noise[7]);
    for (i = 0; i < dim; i++) fprintf(f, " %f %f %f", z[i], vp[i], vpvs[i]);
    fprintf(f, "\n");
    for (i = 0; i < n_events; i++)
        fprintf(f, "EQ  %2s %8ld %d %f %f %f %f %lf %f\n", code, number, i, r, eq[3 * i], eq[3 * i + 1], eq[3 * i + 2],
                reftime[i], origin[i]);
    for (i = 0; i < n_stations; i++) fprintf(f, "RES %2s %8ld %d %f %f %f\n", code, number, i, r, pres[i], sres[i]);
    fflush(f);
}

void mqio_write_counts(FILE* f, const int64_t* c)
{
    static const char* names[8] = {"noise   ", "P-vel   ", "Vp/Vs   ", "quake   ", "resid   ", "move    ", "birth   ", "death   "};
    int i;
    fprintf(f, "cnt RMS tested   %8ld\n", (long)c[0]);
    for (i = 0; i < 8; i++) fprintf(f, "cnt %s a/r %8ld %8ld\n", names[i], (long)c[1 + 2 * i], (long)c[2 + 2 * i]);
}
