/* mcmc_eq_main.c -- the reference's command line on top of the B200 library (plain C host code, C ABI only).
 *
 *     mcmc_eq <config_eqx.dat> <out> <picks> [-n chains] [-d device] [-s seed] [-q]
 *
 * The first three arguments are the reference's (src/mcmc_eq.c:332-338; it ignores anything after them).
 * With one chain (the default) <out> is written exactly like the reference writes it: a "sta ST" record, a
 * "mod X." record every deci-th accepted model, the "bat BF" record and nine "cnt" lines
 * (src/mcmc_eq.c:763,1163,1196-1207; format of print_model_raw :234-248).  With -n N the N chains that the
 * reference would run as N processes (run/srun_mcmc_eq.sh:13,35) run concurrently on one GPU and each writes
 * its own file: <out> is a printf pattern with one integer conversion ("rjx-%03d.out"); without a conversion
 * the chain number is inserted before the extension ("rjx.out" -> "rjx-001.out", the naming scriptsV2/dispe.sh:29
 * globs for).  Chain numbers start at 1 like the SLURM array index.
 *
 * Seeds: config line 32 first value > 0 seeds chain k with value + k (chain 0 == the configured seed);
 * otherwise /dev/urandom (src/mcmc_eq.c:250-265,393-394).  -s overrides.  The random streams are this
 * library's counter-based generator, not libc rand(): chains are statistically, not bitwise, the reference's.
 *
 * Output path: decimated records wait in a device-side ring (-r slots per chain, default 4); after every chunk of
 * iterations the main thread starts an asynchronous drain (mq_drain_begin: pack kernel + device-to-host copy on a second
 * stream) and goes on stepping, a writer thread waits for the batch and formats the print_model_raw text
 * (src/mcmc_eq.c:234-248) into the per-chain files.  The reference writes and flushes each record from inside the
 * sampling loop (:1163).
 *
 * Unlike the reference (which exit(0)s on every error) the exit status is non-zero on failure.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/mcmceq_b200.h"
#include "mq_io.h"

typedef struct {
    int n_chains;
    FILE** out;
    const mqio_picks* pk;
    long* written;
} sink_t;

static const char* kind_tag(int kind) { return kind == MQ_REC_MODEL ? "mod" : kind == MQ_REC_BEST ? "bat" : "sta"; }

static int write_record(void* user, const mq_record* r)
{
    sink_t* s = (sink_t*)user;
    char code[3];
    if (r->chain < 0 || r->chain >= s->n_chains) return 1;
    if (r->kind == MQ_REC_MODEL) { code[0] = r->code; code[1] = '.'; }     /* "Q." "P." ... src/mcmc_eq.c:870-1096 */
    else if (r->kind == MQ_REC_BEST) { code[0] = 'B'; code[1] = 'F'; }     /* :1196 */
    else { code[0] = 'S'; code[1] = 'T'; }                                  /* :763  */
    code[2] = 0;
    mqio_write_record(s->out[r->chain], kind_tag(r->kind), code, (long)r->number, (long)r->dim, r->rms, r->noise, r->z,
                      r->vp, r->vpvs, s->pk->view.n_events, r->eq, s->pk->reftime, r->origin, s->pk->view.n_stations,
                      r->pres, r->sres);
    s->written[r->chain]++;
    return 0;
}

/* Start model from model.dat in the working directory (aflag == 3, src/mcmc_eq.c:381,636-731): an analyse_eq result file
 * whose STAN lines give the velocity model (depth, Vp = column 7, Vp/Vs = column 9 of the line), EQ lines the hypocentres,
 * RES lines the station corrections and the NOISE line the eight sigmas; the letters of config line 34 (V Q R N) select
 * which of them replace the random start values -- in every chain, as every reference process would read the same file. */
static int apply_model_dat(const mq_config* cfg, mq_models* m, FILE* log)
{
    FILE* f = fopen("model.dat", "r");
    char line[4096], tag[64];
    const char* sw = cfg->inp_model_switch;
    int c, k, nv = 0, nq = 0;
    float zz[1000], vv[1000], rr[1000];
    if (!f) { fprintf(stderr, "could not open model file\n"); return 1; }
    while (fgets(line, sizeof line, f)) {
        float a[16];
        int di;
        if (sscanf(line, "%63s", tag) != 1 || tag[0] == '#') continue;
        if (!strcmp(tag, "STAN") && strchr(sw, 'V')) {
            if (sscanf(line, "%*s %f %f %f %f %f %f %f %f", &a[0], &a[1], &a[2], &a[3], &a[4], &a[5], &a[6], &a[7]) < 8) continue;
            if (nv >= m->max_dim || nv >= 1000) { fprintf(stderr, "model larger than reserved space, increase 'max # of cells/layers' in config file\n"); fclose(f); return 1; }
            zz[nv] = a[0]; vv[nv] = a[5]; rr[nv] = a[7]; nv++;
        } else if (!strcmp(tag, "EQ") && strchr(sw, 'Q')) {
            if (sscanf(line, "%*s %d %f %f %f", &di, &a[0], &a[1], &a[2]) < 4 || di < 0 || di >= m->n_events) continue;
            for (c = 0; c < m->n_chains; c++)
                for (k = 0; k < 3; k++) m->eq[((size_t)c * m->n_events + di) * 3 + k] = a[k];
            nq++;
        } else if (!strcmp(tag, "RES") && strchr(sw, 'R')) {
            if (sscanf(line, "%*s %d %f %f", &di, &a[0], &a[1]) < 3 || di < 0 || di >= m->n_stations) continue;
            for (c = 0; c < m->n_chains; c++) { m->pres[(size_t)c * m->n_stations + di] = a[0]; m->sres[(size_t)c * m->n_stations + di] = a[1]; }
        } else if (!strcmp(tag, "NOISE") && strchr(sw, 'N')) {
            if (sscanf(line, "%*s %f %f %f %f %f %f %f %f", &a[0], &a[1], &a[2], &a[3], &a[4], &a[5], &a[6], &a[7]) < 8) continue;
            for (c = 0; c < m->n_chains; c++)
                for (k = 0; k < 4; k++) { m->noise[8 * (size_t)c + 2 * k] = a[k]; m->noise[8 * (size_t)c + 2 * k + 1] = a[4 + k]; }   /* file: p0..p3 s0..s3 */
        }
    }
    fclose(f);
    if (strchr(sw, 'V')) {
        if (nv < 1) { fprintf(stderr, "model.dat holds no STAN line\n"); return 1; }
        for (c = 0; c < m->n_chains; c++) {
            m->dim[c] = nv;
            for (k = 0; k < nv; k++) {
                m->z[(size_t)c * m->max_dim + k] = zz[k]; m->vp[(size_t)c * m->max_dim + k] = vv[k]; m->vpvs[(size_t)c * m->max_dim + k] = rr[k];
            }
        }
    }
    if (strchr(sw, 'Q') && nq != m->n_events) { fprintf(stderr, "number of quakes does not fit pick file!\n"); return 1; }
    if (log) fprintf(log, "start model from model.dat (%s): %d layers, %d quakes\n", sw, nv, nq);
    return 0;
}

/* ---- writer thread: takes begun batches from a two-deep queue, waits for their records, writes them ----------------- */
typedef struct {
    pthread_t thread;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    mq_batch* q[2];
    int n_q, stop, failed;
    long lost, records;
    sink_t* sink;
} writer_t;

static void* writer_main(void* arg)
{
    writer_t* w = (writer_t*)arg;
    for (;;) {
        mq_batch* b;
        int n = 0, lost = 0;
        pthread_mutex_lock(&w->mu);
        while (w->n_q == 0 && !w->stop) pthread_cond_wait(&w->cv, &w->mu);
        if (w->n_q == 0) { pthread_mutex_unlock(&w->mu); return NULL; }
        b = w->q[0];
        pthread_mutex_unlock(&w->mu);
        if (mq_batch_wait(b, &n, &lost) != MQ_OK || mq_batch_deliver(b, write_record, w->sink) != MQ_OK) w->failed = 1;
        mq_batch_release(b);
        pthread_mutex_lock(&w->mu);
        w->q[0] = w->q[1]; w->n_q--;
        w->lost += lost; w->records += n;
        pthread_cond_broadcast(&w->cv);
        pthread_mutex_unlock(&w->mu);
    }
}

static void writer_wait_slot(writer_t* w)             /* until one of the handle's two batches is free again */
{
    pthread_mutex_lock(&w->mu);
    while (w->n_q == 2) pthread_cond_wait(&w->cv, &w->mu);
    pthread_mutex_unlock(&w->mu);
}

static void writer_push(writer_t* w, mq_batch* b)
{
    pthread_mutex_lock(&w->mu);
    w->q[w->n_q++] = b;
    pthread_cond_broadcast(&w->cv);
    pthread_mutex_unlock(&w->mu);
}

static void writer_idle(writer_t* w)                  /* until everything pushed so far is on disk */
{
    pthread_mutex_lock(&w->mu);
    while (w->n_q > 0) pthread_cond_wait(&w->cv, &w->mu);
    pthread_mutex_unlock(&w->mu);
}

static unsigned long urandom_seed(void)
{
    unsigned long v = (unsigned long)time(NULL);
    FILE* f = fopen("/dev/urandom", "rb");
    if (f) { if (fread(&v, sizeof v, 1, f) != 1) v = (unsigned long)time(NULL); fclose(f); }
    return v;
}

static void out_name(char* dst, size_t cap, const char* pattern, int n_chains, int chain1)
{
    const char* pct = strchr(pattern, '%');
    if (n_chains == 1 && !pct) { snprintf(dst, cap, "%s", pattern); return; }
    if (pct) { snprintf(dst, cap, pattern, chain1); return; }
    {
        const char* dot = strrchr(pattern, '.');
        const char* slash = strrchr(pattern, '/');
        if (dot && (!slash || dot > slash)) snprintf(dst, cap, "%.*s-%03d%s", (int)(dot - pattern), pattern, chain1, dot);
        else snprintf(dst, cap, "%s-%03d", pattern, chain1);
    }
}

#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); rc = 1; goto done; } while (0)
#define MQ(call) do { int rc_ = (call); if (rc_ != MQ_OK) FAIL("%s failed (%d): %s", #call, rc_, mq_last_error()); } while (0)

int main(int argc, char** argv)
{
    int rc = 0, n_chains = 1, device = 0, quiet = 0, i, c, from_file = 0, ring = 0, writer_on = 0;
    writer_t wr;
    long seed_arg = -1;
    mq_config cfg;
    mqio_picks pk;
    mq_handle* h = NULL;
    sink_t sink;
    int64_t* counts = NULL;
    double *ll = NULL, *rms = NULL;
    long iters_done = 0;
    const char* env;
    clock_t t0;

    memset(&pk, 0, sizeof pk);
    memset(&sink, 0, sizeof sink);
    memset(&wr, 0, sizeof wr);
    if (argc < 4) {
        fprintf(stderr, "usage: %s config_eqx.dat outfile picks [-n chains] [-d device] [-s seed] [-r ring slots] [-q]\n", argv[0]);
        return 2;
    }
    if ((env = getenv("MCMCEQ_CHAINS")) != NULL) n_chains = atoi(env);
    if ((env = getenv("MCMCEQ_DEVICE")) != NULL) device = atoi(env);
    for (i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "-n") && i + 1 < argc) n_chains = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-d") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) seed_arg = atol(argv[++i]);
        else if (!strcmp(argv[i], "-r") && i + 1 < argc) ring = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-q")) quiet = 1;
        /* anything else is ignored, as the reference ignores extra arguments */
    }
    if (n_chains < 1) n_chains = 1;

    if (mqio_read_config(argv[1], &cfg) != MQ_OK) FAIL("%s", mqio_last_error());
    if (mqio_read_picks(argv[3], &pk) != MQ_OK) FAIL("%s", mqio_last_error());
    if (mqio_check_picks(&cfg, &pk, quiet ? NULL : stderr) != MQ_OK) FAIL("%s", mqio_last_error());
    if (cfg.aflag == 3) { from_file = 1; cfg.aflag = 0; }      /* src/mcmc_eq.c:381 */

    {
        unsigned long seed = seed_arg >= 0 ? (unsigned long)seed_arg : (cfg.true_random > 0 ? (unsigned long)cfg.true_random : urandom_seed());
        if (!quiet) fprintf(stderr, "%d chain(s) on device %d, seed %lu, %d events, %d stations, %d picks\n", n_chains, device, seed,
                            pk.view.n_events, pk.view.n_stations, pk.view.n_picks);
        MQ(mq_create(&cfg, &pk.view, n_chains, device, (uint64_t)seed, &h));
        if (ring > 0) MQ(mq_set_ring(h, ring));
    }

    sink.n_chains = n_chains;
    sink.pk = &pk;
    sink.out = (FILE**)calloc((size_t)n_chains, sizeof(FILE*));
    sink.written = (long*)calloc((size_t)n_chains, sizeof(long));
    counts = (int64_t*)calloc((size_t)n_chains * 20, sizeof(int64_t));
    ll = (double*)calloc((size_t)n_chains, sizeof(double));
    rms = (double*)calloc((size_t)n_chains, sizeof(double));
    if (!sink.out || !sink.written || !counts || !ll || !rms) FAIL("out of memory");
    for (c = 0; c < n_chains; c++) {
        char name[4096];
        out_name(name, sizeof name, argv[2], n_chains, c + 1);
        sink.out[c] = fopen(name, "w");
        if (!sink.out[c]) FAIL("Could not open: %s", name);
    }

    /* start models + first forward (src/mcmc_eq.c:559-765) */
    t0 = clock();
    MQ(mq_init_chains(h));
    if (from_file) {
        /* the random start values stay for whatever model.dat does not supply (src/mcmc_eq.c:559-630 run before :636) */
        mq_models m;
        const size_t n = (size_t)n_chains, md = (size_t)(cfg.max_dim < 1000 ? cfg.max_dim : 1000), ne = (size_t)pk.view.n_events,
                     ns = (size_t)pk.view.n_stations;
        int bad;
        memset(&m, 0, sizeof m);
        m.n_chains = n_chains; m.max_dim = (int)md; m.n_events = (int)ne; m.n_stations = (int)ns;
        m.dim = (int32_t*)calloc(n, sizeof(int32_t));
        m.z = (float*)calloc(n * md, sizeof(float)); m.vp = (float*)calloc(n * md, sizeof(float)); m.vpvs = (float*)calloc(n * md, sizeof(float));
        m.eq = (float*)calloc(n * ne * 3, sizeof(float)); m.pres = (float*)calloc(n * ns, sizeof(float)); m.sres = (float*)calloc(n * ns, sizeof(float));
        m.noise = (float*)calloc(n * 8, sizeof(float)); m.origin = (float*)calloc(n * ne, sizeof(float));
        bad = !m.dim || !m.z || !m.vp || !m.vpvs || !m.eq || !m.pres || !m.sres || !m.noise || !m.origin;
        if (!bad) bad = mq_get_models(h, &m) != MQ_OK || apply_model_dat(&cfg, &m, quiet ? NULL : stderr) || mq_set_models(h, &m) != MQ_OK ||
                        mq_step(h, 0, NULL) != MQ_OK;      /* a step of zero iterations scores the new start models */
        free(m.dim); free(m.z); free(m.vp); free(m.vpvs); free(m.eq); free(m.pres); free(m.sres); free(m.noise); free(m.origin);
        if (bad) FAIL("start from model.dat failed: %s", mq_last_error());
    }
    MQ(mq_get_stats(h, counts, ll, rms));
    if (!quiet) {
        const long ms = (long)((clock() - t0) * 1000 / CLOCKS_PER_SEC);
        fprintf(stderr, "Time taken %ld seconds %ld milliseconds\n", ms / 1000, ms % 1000);
        for (c = 0; c < n_chains && c < 10; c++)
            fprintf(stderr, "Start model found with loglikelihood %f RMS=%f\n", ll[c], rms[c]);
    }
    MQ(mq_snapshot_all(h, 0, write_record, &sink));

    /* main loop (src/mcmc_eq.c:845-1192): the chain ends after j_max_start + j_max_main ACCEPTED models.  A chain's
     * ring holds `ring` decimated records (at most one per deci iterations), so a chunk of ring/2 * deci iterations
     * between two drains loses nothing even with one drain still in flight. */
    {
        const long target = (long)cfg.j_max_start + (long)cfg.j_max_main;
        const int slots = ring > 0 ? ring : 4;
        long chunk_l = (cfg.deci > 0 ? (long)cfg.deci : 1000L) * (slots > 1 ? slots / 2 : 1);
        int chunk;
        if (chunk_l > 2000) chunk_l = 2000;
        if (cfg.deci > 0 && chunk_l < cfg.deci && cfg.deci <= 2000) chunk_l = cfg.deci;
        chunk = (int)chunk_l;
        wr.sink = &sink;
        pthread_mutex_init(&wr.mu, NULL);
        pthread_cond_init(&wr.cv, NULL);
        if (pthread_create(&wr.thread, NULL, writer_main, &wr) != 0) FAIL("could not start the writer thread");
        writer_on = 1;
        for (;;) {
            int running = 0;
            mq_batch* b = NULL;
            MQ(mq_step(h, chunk, NULL));
            iters_done += chunk;
            writer_wait_slot(&wr);
            MQ(mq_drain_begin(h, &b));        /* returns at once; the next chunk runs next to the copy */
            writer_push(&wr, b);
            if (wr.failed) FAIL("writing records failed: %s", mq_last_error());
            MQ(mq_get_stats(h, counts, ll, rms));
            for (c = 0; c < n_chains; c++)
                if (counts[20 * c + 17] < target) running++;
            if (!quiet) {
                double a = 0, r = 0;
                for (c = 0; c < n_chains; c++) { a += (double)counts[20 * c + 17]; r += (double)counts[20 * c + 18]; }
                fprintf(stderr, "Test  %8ld  chains running %d/%d  RMS(chain 1)=%16.10f [s] %5.1f accepted\n", iters_done, running,
                        n_chains, rms[0], 100.0 * a / (a + r > 0 ? a + r : 1));
            }
            if (!running) break;
        }
    }

    /* best model and diagnostics (src/mcmc_eq.c:1196-1207) */
    writer_idle(&wr);
    if (wr.lost) fprintf(stderr, "warning: %ld decimated record(s) were dropped because a chain's ring was full (-r)\n", wr.lost);
    MQ(mq_snapshot_all(h, 1, write_record, &sink));
    for (c = 0; c < n_chains; c++) mqio_write_counts(sink.out[c], counts + 20 * c);

done:
    if (writer_on) {
        pthread_mutex_lock(&wr.mu);
        wr.stop = 1;
        pthread_cond_broadcast(&wr.cv);
        pthread_mutex_unlock(&wr.mu);
        pthread_join(wr.thread, NULL);
    }
    if (sink.out)
        for (c = 0; c < n_chains; c++)
            if (sink.out[c]) fclose(sink.out[c]);
    free(sink.out); free(sink.written); free(counts); free(ll); free(rms);
    if (h) mq_destroy(h);
    mqio_free_picks(&pk);
    return rc;
}
