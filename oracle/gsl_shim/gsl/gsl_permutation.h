/* Minimal GSL-API shim: gsl_permutation and gsl_permute_vector_int. */
#ifndef SHIM_GSL_PERMUTATION_H
#define SHIM_GSL_PERMUTATION_H
#include <stddef.h>
#include "gsl_vector.h"

typedef struct { size_t size; size_t *data; } gsl_permutation;

gsl_permutation *gsl_permutation_alloc(size_t n);
gsl_permutation *gsl_permutation_calloc(size_t n); /* identity */
void gsl_permutation_free(gsl_permutation *p);
int gsl_permutation_inverse(gsl_permutation *inv, const gsl_permutation *p); /* inv[p[i]] = i */
int gsl_permutation_valid(const gsl_permutation *p);                        /* 0 == valid */
int gsl_permute_vector_int(const gsl_permutation *p, gsl_vector_int *v);    /* v'[i] = v[p[i]] */

#ifdef HAVE_INLINE
static inline size_t gsl_permutation_get(const gsl_permutation *p, const size_t i) { return p->data[i]; }
#else
size_t gsl_permutation_get(const gsl_permutation *p, const size_t i);
#endif
#endif
