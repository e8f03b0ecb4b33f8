/* Minimal GSL-API shim: gsl_matrix_int (row-major, tda == size2). */
#ifndef SHIM_GSL_MATRIX_H
#define SHIM_GSL_MATRIX_H
#include <stddef.h>
#include "gsl_vector.h"

typedef struct { size_t size1; size_t size2; size_t tda; int *data; int owner; } gsl_matrix_int;

gsl_matrix_int *gsl_matrix_int_calloc(size_t n1, size_t n2);
void gsl_matrix_int_free(gsl_matrix_int *m);
int gsl_matrix_int_get_col(gsl_vector_int *v, const gsl_matrix_int *m, const size_t j);
_gsl_vector_int_view gsl_matrix_int_row(gsl_matrix_int *m, const size_t i);

#ifdef HAVE_INLINE
static inline int gsl_matrix_int_get(const gsl_matrix_int *m, const size_t i, const size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_int_set(gsl_matrix_int *m, const size_t i, const size_t j, int x) { m->data[i * m->tda + j] = x; }
#else
int gsl_matrix_int_get(const gsl_matrix_int *m, const size_t i, const size_t j);
void gsl_matrix_int_set(gsl_matrix_int *m, const size_t i, const size_t j, int x);
#endif
#endif
