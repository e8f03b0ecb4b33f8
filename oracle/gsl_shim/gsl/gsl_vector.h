/* Minimal GSL-API shim: gsl_vector (double) and gsl_vector_int.
 * Accessors are out-of-line with range checks (like a stock libgsl build
 * without HAVE_INLINE) unless compiled with -DHAVE_INLINE, GSL's own opt-in. */
#ifndef SHIM_GSL_VECTOR_H
#define SHIM_GSL_VECTOR_H
#include <stddef.h>

typedef struct { size_t size; size_t stride; double *data; int owner; } gsl_vector;
typedef struct { size_t size; size_t stride; int *data; int owner; } gsl_vector_int;
typedef struct { gsl_vector_int vector; } _gsl_vector_int_view;
typedef _gsl_vector_int_view gsl_vector_int_view;

gsl_vector *gsl_vector_alloc(size_t n);
gsl_vector *gsl_vector_calloc(size_t n);
void gsl_vector_free(gsl_vector *v);
gsl_vector_int *gsl_vector_int_alloc(size_t n);
gsl_vector_int *gsl_vector_int_calloc(size_t n);
void gsl_vector_int_free(gsl_vector_int *v);
int gsl_vector_int_reverse(gsl_vector_int *v);
void shim_range_error(const char *what, size_t i, size_t n);

#ifdef HAVE_INLINE
static inline double gsl_vector_get(const gsl_vector *v, const size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector *v, const size_t i, double x) { v->data[i * v->stride] = x; }
static inline int gsl_vector_int_get(const gsl_vector_int *v, const size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_int_set(gsl_vector_int *v, const size_t i, int x) { v->data[i * v->stride] = x; }
#else
double gsl_vector_get(const gsl_vector *v, const size_t i);
void gsl_vector_set(gsl_vector *v, const size_t i, double x);
int gsl_vector_int_get(const gsl_vector_int *v, const size_t i);
void gsl_vector_int_set(gsl_vector_int *v, const size_t i, int x);
#endif
#endif
