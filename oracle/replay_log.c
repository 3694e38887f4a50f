/* oracle/replay_log.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, chain level).
 *
 * Runs the UNMODIFIED reference chain driver -- oracle/_ref/libmcmceq_ref.so, i.e. src/mcmc_eq.c compiled from where
 * it lies with main renamed to ref_mcmc_eq_main (oracle/Makefile) -- and records its proposal stream without touching
 * a line of it: this executable defines cal_fit_newx, copy_model and rand itself, the dynamic linker binds the
 * library's calls to them (they go through the PLT), and each wrapper forwards to the real function (RTLD_NEXT).
 *
 *   cal_fit_newx(m, ..., calct, ...)   one call per evaluated proposal (src/mcmc_eq.c:884,930,953,975,1004,1041,
 *                                      1082,1119) and one for the start model (:739): the proposed Model, calct and
 *                                      the eight class sums the reference computed are recorded;
 *   rand()                             the first draw after a proposal's cal_fit_newx is the uniform deviate of the
 *                                      accept test, rand_eq() at src/mcmc_eq.c:1141;
 *   copy_model(dest, src)              src == the proposed Model before the next proposal means "accepted"
 *                                      (src/mcmc_eq.c:1156).
 *
 * usage: replay_log <log.bin> <config> <out> <picks>      (the last three are the reference's own arguments)
 *
 * log.bin: int32 magic 0x4d435251, int32 noq, int32 nos, then per record
 *   int32 calct, dim, accepted; float u; float mf[8]; float noise[8]   (mf, noise indexed 2*class+phase)
 *   float z[dim], vp[dim], vpvs[dim]; float eq[3*noq]; float pres[nos], sres[nos]; float origin[noq]
 * tools/make_golden.py turns it into tests/golden/replay_example2.npz.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mc.h" /* the reference's own header, included from REF_SRC (never copied) */

int ref_mcmc_eq_main(int argc, char** argv);

typedef float (*calfit_fn)(struct Model*, struct DATA*, int, float***, float***, struct GRDHEAD, int, float*, float*, float*,
                           float*, float*, float*, float*, float*, int, int, int);
typedef void (*copy_fn)(struct Model*, struct Model*);
typedef int (*rand_fn)(void);

static FILE* g_log;
static struct Model* g_model;   /* Model of the pending record */
static struct Model g_snap;     /* its state right after cal_fit_newx returned */
static int g_pending, g_want_u, g_calct, g_accepted, g_header;
static float g_u, g_mf[8];

static void flush_record(void)
{
    int32_t head[3];
    int i;
    if (!g_pending || !g_log) return;
    if (!g_header) {
        int32_t h[3] = {0x4d435251, (int32_t)g_snap.noq, (int32_t)g_snap.nos};
        fwrite(h, sizeof h, 1, g_log);
        g_header = 1;
    }
    head[0] = g_calct; head[1] = (int32_t)g_snap.dimension; head[2] = g_accepted;
    fwrite(head, sizeof head, 1, g_log);
    fwrite(&g_u, sizeof(float), 1, g_log);
    fwrite(g_mf, sizeof(float), 8, g_log);
    fwrite(&g_snap.p_noise0, sizeof(float), 8, g_log);   /* p0 s0 p1 s1 p2 s2 p3 s3: contiguous in struct Model */
    fwrite(g_snap.z, sizeof(float), (size_t)g_snap.dimension, g_log);
    fwrite(g_snap.vp, sizeof(float), (size_t)g_snap.dimension, g_log);
    fwrite(g_snap.vpvs, sizeof(float), (size_t)g_snap.dimension, g_log);
    for (i = 0; i < g_snap.noq; i++) fwrite(&g_snap.eq[i], sizeof(float), 3, g_log);
    fwrite(g_snap.pres, sizeof(float), (size_t)g_snap.nos, g_log);
    fwrite(g_snap.sres, sizeof(float), (size_t)g_snap.nos, g_log);
    fwrite(g_snap.origin, sizeof(float), (size_t)g_snap.noq, g_log);
    g_pending = 0;
}

static void at_exit(void)
{
    flush_record();
    if (g_log) fclose(g_log);
    g_log = NULL;
}

float cal_fit_newx(struct Model* m, struct DATA* d, int ne, float*** tttp, float*** ttts, struct GRDHEAD gh, int calct,
                   float* mfp0, float* mfs0, float* mfp1, float* mfs1, float* mfp2, float* mfs2, float* mfp3, float* mfs3,
                   int flag, int eikonal, int out)
{
    static calfit_fn real;
    float r;
    if (!real) real = (calfit_fn)dlsym(RTLD_NEXT, "cal_fit_newx");
    flush_record();
    r = real(m, d, ne, tttp, ttts, gh, calct, mfp0, mfs0, mfp1, mfs1, mfp2, mfs2, mfp3, mfs3, flag, eikonal, out);
    g_model = m;
    memcpy(&g_snap, m, sizeof g_snap);
    g_calct = calct;
    g_mf[0] = *mfp0; g_mf[1] = *mfs0; g_mf[2] = *mfp1; g_mf[3] = *mfs1;
    g_mf[4] = *mfp2; g_mf[5] = *mfs2; g_mf[6] = *mfp3; g_mf[7] = *mfs3;
    g_accepted = g_header ? 0 : 1;   /* the first call scores the start model */
    g_u = -1.f;
    g_want_u = 1;
    g_pending = 1;
    return r;
}

void copy_model(struct Model* dest, struct Model* src)
{
    static copy_fn real;
    if (!real) real = (copy_fn)dlsym(RTLD_NEXT, "copy_model");
    if (g_pending && g_header && src == g_model && !g_want_u) g_accepted = 1;
    real(dest, src);
}

int rand(void)
{
    static rand_fn real;
    int v;
    if (!real) real = (rand_fn)dlsym(RTLD_NEXT, "rand");
    v = real();
    if (g_want_u && g_header) { g_u = (float)v / RAND_MAX; g_want_u = 0; }   /* rand_eq(), src/mcmc_eq.c:168-172 */
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s log.bin config out picks\n", argv[0]); return 2; }
    g_log = fopen(argv[1], "wb");
    if (!g_log) { perror(argv[1]); return 1; }
    atexit(at_exit);
    return ref_mcmc_eq_main(argc - 1, argv + 1);   /* ends in exit(0), src/mcmc_eq.c:1210 */
}
