/* Minimal GSL-API shim: gsl_rng.  The generator behind it is a pluggable
 * draw source (oracle/draw_source.h): MT19937 (GSL's documented default),
 * the structured Philox stream of the B200 build, or a recorded tape. */
#ifndef SHIM_GSL_RNG_H
#define SHIM_GSL_RNG_H
#include <stddef.h>

typedef struct { const char *name; } gsl_rng_type;
typedef struct shim_rng gsl_rng;

extern const gsl_rng_type *gsl_rng_default;
extern unsigned long int gsl_rng_default_seed;

const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
double gsl_rng_uniform_pos(const gsl_rng *r);
unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n);
#endif
