/* mq_io.h -- host-side readers/writers for the reference's file formats (C, no CUDA).
 *
 *   config_eqx.dat : 41 positional lines, parsed like src/mcmc_eq.c:345-388
 *   pick file      : "# id nP nS reftime [xfix yfix zfix]" headers followed by
 *                    "name st_id P|S x y z t class" rows, src/mcmc_eq.c:1217-1300
 *   chain output   : print_model_raw, src/mcmc_eq.c:234-248, and the cnt lines :1199-1207
 */
#ifndef MQ_IO_H
#define MQ_IO_H

#include <stdio.h>
#include "../../include/mcmceq_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Owning version of mq_picks (arrays are malloc'ed; release with mqio_free_picks). */
typedef struct mqio_picks {
    mq_picks view;        /* const pointers into the arrays below */
    int32_t *ev_off, *n_p, *st_id, *cls, *eq_id;
    float *x, *y, *z, *t;
    double *reftime, *fix;
    int32_t n_class[MQ_MAX_CLASSES]; /* picks per class, index 2*class+phase (n_ppicks0, n_spicks0, ...) */
} mqio_picks;

/* Returns MQ_OK or MQ_ERR_ARG (message via mqio_last_error()). */
int mqio_read_config(const char* path, mq_config* cfg);
int mqio_read_picks(const char* path, mqio_picks* out);
void mqio_free_picks(mqio_picks* p);
const char* mqio_last_error(void);

/* Checks the reference performs after loading the picks (src/mcmc_eq.c:447-500): station ids,
 * reference station, stations inside the model box.  Warnings go to `log` (may be NULL). */
int mqio_check_picks(const mq_config* cfg, const mqio_picks* p, FILE* log);

/* One record in the format of print_model_raw (src/mcmc_eq.c:234-248): tag is "sta"/"mod"/"bat",
 * code the two-letter decision ("ST", "Q.", "BF" ...).  noise is indexed 2*class+phase. */
void mqio_write_record(FILE* f, const char* tag, const char* code, long number, long dim, double rms,
                       const float* noise, const float* z, const float* vp, const float* vpvs, int n_events,
                       const float* eq, const double* reftime, const float* origin, int n_stations,
                       const float* pres, const float* sres);
/* The nine cnt lines (src/mcmc_eq.c:1199-1207); counts as returned by mq_get_stats for one chain. */
void mqio_write_counts(FILE* f, const int64_t* counts);

#ifdef __cplusplus
}
#endif
#endif
