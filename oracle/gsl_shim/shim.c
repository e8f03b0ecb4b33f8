/*
 * shim.c -- the part of the GSL API that the reference sampler
 * (/root/reference/C_Implementation/mcmc.c) links against, written from the
 * GSL documentation because GSL is not installed in this image.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/README.md).  Containers are plain; the RNG
 * entry points forward to oracle/draw_source.h so a run can be driven by
 * MT19937, by the structured Philox stream, or by a recorded tape, and can
 * record the tape that the B200 replay mode consumes.
 *
 * Environment:
 *   GSL_RNG_SEED=<n>       MT19937 seed (GSL's own variable; echoed to stderr)
 *   SER_RNG=philox         use the structured Philox stream instead
 *   SER_SEED=<n> SER_CHAIN=<n>   Philox key (seed, chain id)
 *   SER_TAPE_IN=<file>     replay raw little-endian doubles from <file>
 *   SER_TAPE_OUT=<file>    record the tape; written by gsl_rng_free()
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gsl/gsl_math.h"
#include "gsl/gsl_matrix.h"
#include "gsl/gsl_permutation.h"
#include "gsl/gsl_randist.h"
#include "gsl/gsl_rng.h"
#include "gsl/gsl_vector.h"

#include "../draw_source.h"

/* ------------------------------------------------------------------ errors */
void shim_range_error(const char *what, size_t i, size_t n)
{
  fprintf(stderr, "gsl shim: %s index %zu out of range [0,%zu)\n", what, i, n);
  abort();
}

static void *xmalloc(size_t n)
{
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "gsl shim: out of memory\n"); abort(); }
  return p;
}

/* ----------------------------------------------------------------- vectors */
gsl_vector *gsl_vector_alloc(size_t n)
{
  gsl_vector *v = (gsl_vector *)xmalloc(sizeof(*v));
  v->size = n; v->stride = 1; v->owner = 1;
  v->data = (double *)xmalloc(n * sizeof(double));
  return v;
}
gsl_vector *gsl_vector_calloc(size_t n)
{
  gsl_vector *v = gsl_vector_alloc(n);
  memset(v->data, 0, n * sizeof(double));
  return v;
}
void gsl_vector_free(gsl_vector *v) { if (v) { if (v->owner) free(v->data); free(v); } }

gsl_vector_int *gsl_vector_int_alloc(size_t n)
{
  gsl_vector_int *v = (gsl_vector_int *)xmalloc(sizeof(*v));
  v->size = n; v->stride = 1; v->owner = 1;
  v->data = (int *)xmalloc(n * sizeof(int));
  return v;
}
gsl_vector_int *gsl_vector_int_calloc(size_t n)
{
  gsl_vector_int *v = gsl_vector_int_alloc(n);
  memset(v->data, 0, n * sizeof(int));
  return v;
}
void gsl_vector_int_free(gsl_vector_int *v) { if (v) { if (v->owner) free(v->data); free(v); } }

int gsl_vector_int_reverse(gsl_vector_int *v)
{
  size_t i, n = v->size, s = v->stride;
  for (i = 0; i < n / 2; i++) {
    int t = v->data[i * s];
    v->data[i * s] = v->data[(n - 1 - i) * s];
    v->data[(n - 1 - i) * s] = t;
  }
  return 0;
}

#ifndef HAVE_INLINE
double gsl_vector_get(const gsl_vector *v, const size_t i)
{
  if (i >= v->size) shim_range_error("gsl_vector_get", i, v->size);
  return v->data[i * v->stride];
}
void gsl_vector_set(gsl_vector *v, const size_t i, double x)
{
  if (i >= v->size) shim_range_error("gsl_vector_set", i, v->size);
  v->data[i * v->stride] = x;
}
int gsl_vector_int_get(const gsl_vector_int *v, const size_t i)
{
  if (i >= v->size) shim_range_error("gsl_vector_int_get", i, v->size);
  return v->data[i * v->stride];
}
void gsl_vector_int_set(gsl_vector_int *v, const size_t i, int x)
{
  if (i >= v->size) shim_range_error("gsl_vector_int_set", i, v->size);
  v->data[i * v->stride] = x;
}
#endif

/* ------------------------------------------------------------------ matrix */
gsl_matrix_int *gsl_matrix_int_calloc(size_t n1, size_t n2)
{
  gsl_matrix_int *m = (gsl_matrix_int *)xmalloc(sizeof(*m));
  m->size1 = n1; m->size2 = n2; m->tda = n2; m->owner = 1;
  m->data = (int *)xmalloc(n1 * n2 * sizeof(int));
  memset(m->data, 0, n1 * n2 * sizeof(int));
  return m;
}
void gsl_matrix_int_free(gsl_matrix_int *m) { if (m) { if (m->owner) free(m->data); free(m); } }

int gsl_matrix_int_get_col(gsl_vector_int *v, const gsl_matrix_int *m, const size_t j)
{
  size_t i;
  if (j >= m->size2) shim_range_error("gsl_matrix_int_get_col", j, m->size2);
  if (v->size != m->size1) shim_range_error("gsl_matrix_int_get_col(len)", v->size, m->size1);
  for (i = 0; i < m->size1; i++) v->data[i * v->stride] = m->data[i * m->tda + j];
  return 0;
}

_gsl_vector_int_view gsl_matrix_int_row(gsl_matrix_int *m, const size_t i)
{
  _gsl_vector_int_view view;
  if (i >= m->size1) shim_range_error("gsl_matrix_int_row", i, m->size1);
  view.vector.size = m->size2;
  view.vector.stride = 1;
  view.vector.data = m->data + i * m->tda;
  view.vector.owner = 0;
  return view;
}

#ifndef HAVE_INLINE
int gsl_matrix_int_get(const gsl_matrix_int *m, const size_t i, const size_t j)
{
  if (i >= m->size1) shim_range_error("gsl_matrix_int_get(row)", i, m->size1);
  if (j >= m->size2) shim_range_error("gsl_matrix_int_get(col)", j, m->size2);
  return m->data[i * m->tda + j];
}
void gsl_matrix_int_set(gsl_matrix_int *m, const size_t i, const size_t j, int x)
{
  if (i >= m->size1) shim_range_error("gsl_matrix_int_set(row)", i, m->size1);
  if (j >= m->size2) shim_range_error("gsl_matrix_int_set(col)", j, m->size2);
  m->data[i * m->tda + j] = x;
}
#endif

/* ------------------------------------------------------------ permutations */
gsl_permutation *gsl_permutation_alloc(size_t n)
{
  gsl_permutation *p = (gsl_permutation *)xmalloc(sizeof(*p));
  p->size = n;
  p->data = (size_t *)xmalloc(n * sizeof(size_t));
  return p;
}
gsl_permutation *gsl_permutation_calloc(size_t n)
{
  size_t i;
  gsl_permutation *p = gsl_permutation_alloc(n);
  for (i = 0; i < n; i++) p->data[i] = i;
  return p;
}
void gsl_permutation_free(gsl_permutation *p) { if (p) { free(p->data); free(p); } }

int gsl_permutation_inverse(gsl_permutation *inv, const gsl_permutation *p)
{
  size_t i;
  if (inv->size != p->size) shim_range_error("gsl_permutation_inverse(len)", inv->size, p->size);
  for (i = 0; i < p->size; i++) inv->data[p->data[i]] = i;
  return 0;
}

int gsl_permutation_valid(const gsl_permutation *p)
{
  size_t i, j, n = p->size;
  for (i = 0; i < n; i++) {
    if (p->data[i] >= n) return -1;
    for (j = 0; j < i; j++)
      if (p->data[i] == p->data[j]) return -1;
  }
  return 0; /* GSL_SUCCESS */
}

int gsl_permute_vector_int(const gsl_permutation *p, gsl_vector_int *v)
{
  size_t i, n = v->size;
  int *tmp;
  if (p->size != n) shim_range_error("gsl_permute_vector_int(len)", p->size, n);
  tmp = (int *)xmalloc(n * sizeof(int));
  for (i = 0; i < n; i++) tmp[i] = v->data[p->data[i] * v->stride];
  for (i = 0; i < n; i++) v->data[i * v->stride] = tmp[i];
  free(tmp);
  return 0;
}

#ifndef HAVE_INLINE
size_t gsl_permutation_get(const gsl_permutation *p, const size_t i)
{
  if (i >= p->size) shim_range_error("gsl_permutation_get", i, p->size);
  return p->data[i];
}
#endif

/* --------------------------------------------------------------------- rng */
struct shim_rng { draw_source src; char *tape_out; double *tape_in_buf; };

static const gsl_rng_type shim_mt19937 = { "mt19937" };
const gsl_rng_type *gsl_rng_default = &shim_mt19937;
unsigned long int gsl_rng_default_seed = 0;

static struct shim_rng *g_last_rng = NULL;

const gsl_rng_type *gsl_rng_env_setup(void)
{
  const char *p = getenv("GSL_RNG_SEED");
  gsl_rng_default = &shim_mt19937;
  if (p) {
    gsl_rng_default_seed = strtoul(p, 0, 0);
    fprintf(stderr, "GSL_RNG_SEED=%lu\n", gsl_rng_default_seed);
  }
  return gsl_rng_default;
}

static double *read_tape_file(const char *path, size_t *len)
{
  FILE *f = fopen(path, "rb");
  long sz;
  double *buf;
  if (!f) { fprintf(stderr, "gsl shim: cannot open tape %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf = (double *)xmalloc((size_t)sz);
  if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "gsl shim: short read on %s\n", path); exit(2); }
  fclose(f);
  *len = (size_t)sz / sizeof(double);
  return buf;
}

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
  struct shim_rng *r = (struct shim_rng *)xmalloc(sizeof(*r));
  const char *kind = getenv("SER_RNG"), *tin = getenv("SER_TAPE_IN"), *tout = getenv("SER_TAPE_OUT");
  (void)T;
  r->tape_out = NULL;
  r->tape_in_buf = NULL;
  if (tin && *tin) {
    size_t len;
    r->tape_in_buf = read_tape_file(tin, &len);
    ds_init_tape(&r->src, r->tape_in_buf, len);
  } else if (kind && strcmp(kind, "philox") == 0) {
    const char *s = getenv("SER_SEED"), *c = getenv("SER_CHAIN");
    ds_init_philox(&r->src, s ? (uint32_t)strtoul(s, 0, 0) : 0u, c ? (uint32_t)strtoul(c, 0, 0) : 0u);
  } else {
    ds_init_mt(&r->src, gsl_rng_default_seed);
  }
  if (tout && *tout) {
    r->tape_out = strdup(tout);
    r->src.recording = 1;
  }
  g_last_rng = r;
  return r;
}

/* harness access: the draw source behind the most recently allocated rng */
draw_source *shim_source(void) { return g_last_rng ? &g_last_rng->src : NULL; }
void shim_set_recording(int on) { if (g_last_rng) g_last_rng->src.recording = on; }

void gsl_rng_free(gsl_rng *r)
{
  if (!r) return;
  if (r->tape_out) {
    FILE *f = fopen(r->tape_out, "wb");
    if (!f || fwrite(r->src.rec, sizeof(double), r->src.rec_n, f) != r->src.rec_n) {
      fprintf(stderr, "gsl shim: cannot write tape %s\n", r->tape_out);
      exit(2);
    }
    fclose(f);
    free(r->tape_out);
  }
  free(r->src.rec);
  free(r->tape_in_buf);
  if (g_last_rng == r) g_last_rng = NULL;
  free(r);
}

double gsl_rng_uniform(const gsl_rng *r) { return ds_uniform(&((struct shim_rng *)r)->src); }
double gsl_rng_uniform_pos(const gsl_rng *r) { return ds_uniform_pos(&((struct shim_rng *)r)->src); }
unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n)
{
  return ds_uniform_int(&((struct shim_rng *)r)->src, n);
}

/* ----------------------------------------------------------------- randist */
double gsl_ran_beta(const gsl_rng *r, const double a, const double b)
{
  return ds_beta(&((struct shim_rng *)r)->src, a, b, NULL, NULL);
}

/* GSL: Fisher-Yates from the top, one uniform_int(i+1) per i = n-1 .. 1 */
void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t n, size_t size)
{
  size_t i;
  char *b = (char *)base, *tmp = (char *)xmalloc(size);
  for (i = n - 1; i > 0; i--) {
    size_t j = gsl_rng_uniform_int(r, i + 1);
    if (i != j) {
      memcpy(tmp, b + i * size, size);
      memcpy(b + i * size, b + j * size, size);
      memcpy(b + j * size, tmp, size);
    }
  }
  free(tmp);
}

/* GSL: sequential selection sampling, one uniform per inspected source item */
void *gsl_ran_choose(const gsl_rng *r, void *dest, size_t k, void *src, size_t n, size_t size)
{
  size_t i, j = 0;
  if (k > n) { fprintf(stderr, "gsl shim: gsl_ran_choose k > n\n"); abort(); }
  for (i = 0; i < n && j < k; i++) {
    if ((double)(n - i) * gsl_rng_uniform(r) < (double)(k - j)) {
      memcpy((char *)dest + size * j, (char *)src + size * i, size);
      j++;
    }
  }
  return dest;
}
