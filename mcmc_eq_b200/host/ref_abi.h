/* ref_abi.h -- binary layout of the types that cross the reference's function seam, declared here
 * so that libmcmceq_shim.so can be linked with the UNMODIFIED reference sources in place of time_2d.o / misfit.c /
 * interpol.c.  The layouts are the reference's own (src/mc.h:49-134): field order, types and array bounds are an ABI
 * and cannot differ; the names of the tags are the reference's because its translation units spell them.
 * Nothing here is used by libmcmceq_b200.so itself.
 */
#ifndef MCMCEQ_REF_ABI_H
#define MCMCEQ_REF_ABI_H

#define REF_MD 1000        /* MD, src/mc.h:49        */
#define REF_MAX_OBS 1000   /* MAX_OBS, src/mc.h:52   */
#define REF_MAX_STAT 1000  /* MAX_STAT, src/mc.h:53  */
#define REF_MAX_NOQ 3500   /* MAX_NOQ, src/mc.h:54   */

struct QUAKE { float x, y, z; };                     /* src/mc.h:61-66 */

struct Model {                                       /* src/mc.h:68-89 */
    long number, dimension, noq, nos;
    float pres[REF_MAX_STAT], sres[REF_MAX_STAT];
    float origin[REF_MAX_NOQ];
    float noise[8];                                  /* p_noise0, s_noise0, p_noise1, ... s_noise3: index 2*class + phase */
    float z[REF_MD], vp[REF_MD], vpvs[REF_MD];
    struct QUAKE eq[REF_MAX_NOQ];
};

struct GRDHEAD { int nx, ny, nz; float h, x0, y0, z0; };   /* src/mc.h:91-100 */

struct OBS {                                         /* src/mc.h:102-113 */
    int st_id;
    float x, y, z, t;
    int cl, layer;
    float w1, w2;
};

struct DATA {                                        /* src/mc.h:115-134 */
    int eq_id;
    double reftime, xfix, yfix, zfix;
    int nobs_p, nobs_s;
    int nobs_class[8];                               /* nobs_p0, nobs_s0, ... nobs_p3, nobs_s3 */
    struct OBS p_picks[REF_MAX_OBS], s_picks[REF_MAX_OBS];
};

#endif
