/*
 * draw_source.h -- pluggable random-draw source for the ORACLE side.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path includes this file.
 * It sits behind (a) the GSL-API shim that the unmodified reference mcmc.c is
 * linked against (oracle/gsl_shim/shim.c) and (b) the CPU restatement
 * (oracle/seriation_oracle.c), so both see the same draws.
 *
 * Kinds
 *   DS_MT      MT19937 the way GSL documents gsl_rng_mt19937 (default type,
 *              seed 0 -> 4357, uniform = get/2^32, uniform_int by scaled
 *              rejection).  GSL itself is absent here, so this is "as
 *              documented", NOT verified against libgsl: parity is pinned at
 *              the tape level instead.
 *   DS_PHILOX  the structured Philox4x32-10 stream the B200 build uses in
 *              free-running mode (csrc/ser_detmath.h).  The position inside a
 *              sweep is tracked by a small state machine driven by the *kind*
 *              of call, which the sampler's fixed program order makes
 *              unambiguous (scalar c/d only).
 *   DS_TAPE    replays a recorded tape.
 *
 * Tape grammar (one flat double[] per chain, program order; the contract of
 * the B200 replay mode):
 *   uniform()        1 slot : u in [0,1)
 *   uniform_pos()    1 slot : u in (0,1)
 *   uniform_int(n)   1 slot : (k+0.5)/n, so (unsigned long)(slot*n) == k
 *   beta(a,b)        3 slots: y, log(y), log(1-exp(log(y)))  -- the last two
 *                    evaluated with the host libm, i.e. the very bits the
 *                    reference computes at mcmc.c:760 and :847-848.
 */
#ifndef ORACLE_DRAW_SOURCE_H
#define ORACLE_DRAW_SOURCE_H

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../seriation-in-paleontological-data-using-mcmc_b200/csrc/ser_detmath.h"

enum { DS_MT = 0, DS_PHILOX = 1, DS_TAPE = 2 };
enum { DS_ST_INIT = 0, DS_ST_C = 1, DS_ST_D = 2, DS_ST_AB = 3, DS_ST_PI = 4 };

typedef struct draw_source {
  int kind;
  /* MT19937 */
  uint32_t mt[624];
  int mti;
  /* structured Philox */
  uint32_t seed, chain, sweep, block, idx;
  int state;
  int manycd, ntaxa, nbeta; /* manycd: M Betas for c then M for d per sweep; nbeta counts them */
  /* tape in */
  const double *tape;
  size_t tape_len, cur;
  /* tape out */
  double *rec;
  size_t rec_n, rec_cap;
  int recording;
  /* call statistics */
  unsigned long long n_uniform, n_pos, n_int, n_beta;
} draw_source;

static inline void ds_rec_push(draw_source *s, double v)
{
  if (!s->recording) return;
  if (s->rec_n == s->rec_cap) {
    s->rec_cap = s->rec_cap ? s->rec_cap * 2 : (1u << 16);
    s->rec = (double *)realloc(s->rec, s->rec_cap * sizeof(double));
    if (!s->rec) { fprintf(stderr, "draw_source: out of memory\n"); exit(2); }
  }
  s->rec[s->rec_n++] = v;
}

static inline void ds_mt_seed(draw_source *s, unsigned long seed)
{
  if (seed == 0) seed = 4357; /* GSL: "the default seed" */
  s->mt[0] = (uint32_t)(seed & 0xffffffffUL);
  for (int i = 1; i < 624; i++)
    s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->mti = 624;
}

static inline uint32_t ds_mt_get(draw_source *s)
{
  if (s->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624; kk++) {
      uint32_t y = (s->mt[kk] & 0x80000000u) | (s->mt[(kk + 1) % 624] & 0x7fffffffu);
      s->mt[kk] = s->mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->mti = 0;
  }
  uint32_t k = s->mt[s->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680u;
  k ^= (k << 15) & 0xefc60000u;
  k ^= (k >> 18);
  return k;
}

static inline void ds_init_mt(draw_source *s, unsigned long seed)
{
  memset(s, 0, sizeof(*s));
  s->kind = DS_MT;
  ds_mt_seed(s, seed);
}

static inline void ds_init_philox(draw_source *s, uint32_t seed, uint32_t chain)
{
  memset(s, 0, sizeof(*s)); /* scalar c/d; ds_set_manycd() switches to per-taxon Betas */
  s->kind = DS_PHILOX;
  s->seed = seed;
  s->chain = chain;
  s->sweep = SER_SWEEP_INIT;
  s->block = SER_BLK_INIT;
  s->state = DS_ST_INIT;
}

static inline void ds_set_manycd(draw_source *s, int ntaxa) { s->manycd = 1; s->ntaxa = ntaxa; }

static inline void ds_init_tape(draw_source *s, const double *tape, size_t len)
{
  memset(s, 0, sizeof(*s));
  s->kind = DS_TAPE;
  s->tape = tape;
  s->tape_len = len;
}

static inline double ds_tape_next(draw_source *s)
{
  if (s->cur >= s->tape_len) {
    fprintf(stderr, "draw_source: tape exhausted at slot %zu\n", s->cur);
    exit(3);
  }
  return s->tape[s->cur++];
}

/* raw uniform in [0,1) from the native generator (not recorded) */
static inline double ds_raw_uniform(draw_source *s, int is_plain_uniform)
{
  if (s->kind == DS_MT) return ds_mt_get(s) / 4294967296.0;
  /* DS_PHILOX: advance the block state machine on the call kind */
  if (s->state == DS_ST_INIT) {
    /* stays in the init block */
  } else if (is_plain_uniform) {
    if (s->state == DS_ST_D) { s->state = DS_ST_AB; s->block = SER_BLK_AB; s->idx = 0; }
    else if (s->state != DS_ST_AB) { fprintf(stderr, "draw_source: unexpected uniform() in state %d\n", s->state); exit(4); }
  } else {
    if (s->state == DS_ST_AB) { s->state = DS_ST_PI; s->block = SER_BLK_PI; s->idx = 0; }
    else if (s->state != DS_ST_PI) { fprintf(stderr, "draw_source: unexpected int/pos draw in state %d\n", s->state); exit(4); }
  }
  return ser_stream_uniform(s->seed, s->chain, s->sweep, s->block, s->idx++);
}

static inline double ds_uniform(draw_source *s)
{
  double u;
  s->n_uniform++;
  if (s->kind == DS_TAPE) u = ds_tape_next(s);
  else u = ds_raw_uniform(s, 1);
  ds_rec_push(s, u);
  return u;
}

static inline double ds_uniform_pos(draw_source *s)
{
  double u;
  s->n_pos++;
  if (s->kind == DS_TAPE) u = ds_tape_next(s);
  else if (s->kind == DS_MT) { do { u = ds_raw_uniform(s, 0); } while (u == 0.0); }
  else u = ser_pos(ds_raw_uniform(s, 0));
  ds_rec_push(s, u);
  return u;
}

static inline unsigned long ds_uniform_int(draw_source *s, unsigned long n)
{
  unsigned long k;
  s->n_int++;
  if (s->kind == DS_TAPE) {
    double u = ds_tape_next(s);
    k = (unsigned long)(u * (double)n);
    ds_rec_push(s, u);
    return k;
  }
  if (s->kind == DS_MT) { /* GSL: scale = range / n, reject k >= n */
    unsigned long scale = 0xffffffffUL / n;
    do { k = ds_mt_get(s) / scale; } while (k >= n);
  } else {
    k = (unsigned long)(ds_raw_uniform(s, 0) * (double)n);
  }
  ds_rec_push(s, ((double)k + 0.5) / (double)n);
  return k;
}

/* Gamma(shape>=1) from sequential MT draws (libm math; MT kind only). */
static inline double ds_mt_gamma(draw_source *s, double shape)
{
  const double d = shape - 1.0 / 3.0, c = (1.0 / 3.0) / sqrt(d);
  for (;;) {
    double x, v;
    do {
      double a, b, r2;
      do {
        a = 2.0 * (ds_mt_get(s) / 4294967296.0) - 1.0;
        b = 2.0 * (ds_mt_get(s) / 4294967296.0) - 1.0;
        r2 = a * a + b * b;
      } while (r2 >= 1.0 || r2 == 0.0);
      x = a * sqrt(-2.0 * log(r2) / r2);
      v = 1.0 + c * x;
    } while (v <= 0.0);
    v = v * v * v;
    double u;
    do { u = ds_mt_get(s) / 4294967296.0; } while (u == 0.0);
    if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
    if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
  }
}

/*
 * Beta(a, b) with a, b >= 1 (the sampler always passes 1 + count).
 * Returns the variate; *logy / *log1m receive the libm companions that the
 * tape carries.  In DS_PHILOX kind the variate is produced by the shared
 * bit-reproducible sampler so that the GPU free-running mode yields the same y.
 */
static inline double ds_beta(draw_source *s, double a, double b, double *logy, double *log1m)
{
  double y, ly, l1;
  s->n_beta++;
  if (s->kind == DS_TAPE) {
    y = ds_tape_next(s);
    ly = ds_tape_next(s);
    l1 = ds_tape_next(s);
  } else {
    if (s->kind == DS_MT) {
      double g1 = ds_mt_gamma(s, a), g2 = ds_mt_gamma(s, b);
      y = g1 / (g1 + g2);
    } else if (s->manycd) {
      /* per-taxon c then per-taxon d: Beta number k of the sweep belongs to taxon k % M */
      uint32_t blk;
      if (s->state != DS_ST_C && s->state != DS_ST_D) { /* first Beta of a new sweep */
        s->sweep = (s->state == DS_ST_INIT) ? 0u : s->sweep + 1u;
        s->state = DS_ST_C;
        s->nbeta = 0;
      }
      blk = SER_BLK_MANYCD + 4u * (uint32_t)(s->nbeta % s->ntaxa) + (s->nbeta >= s->ntaxa ? 2u : 0u);
      if (++s->nbeta >= s->ntaxa) s->state = DS_ST_D;
      {
        double g1 = ser_gamma_ge1(a, s->seed, s->chain, s->sweep, blk);
        double g2 = ser_gamma_ge1(b, s->seed, s->chain, s->sweep, blk + 1u);
        y = ser_beta_from_gammas(g1, g2);
      }
    } else {
      uint32_t blk;
      if (s->state == DS_ST_C) { s->state = DS_ST_D; blk = SER_BLK_D_GAMMA_A; }
      else { /* first Beta of a new sweep */
        s->sweep = (s->state == DS_ST_INIT) ? 0u : s->sweep + 1u;
        s->state = DS_ST_C;
        blk = SER_BLK_C_GAMMA_A;
      }
      {
        double g1 = ser_gamma_ge1(a, s->seed, s->chain, s->sweep, blk);
        double g2 = ser_gamma_ge1(b, s->seed, s->chain, s->sweep, blk + 1u);
        y = ser_beta_from_gammas(g1, g2);
      }
    }
    ly = (y > 0.0) ? log(y) : -INFINITY;
    l1 = log(1. - exp(ly));
  }
  ds_rec_push(s, y);
  ds_rec_push(s, ly);
  ds_rec_push(s, l1);
  if (logy) *logy = ly;
  if (log1m) *log1m = l1;
  return y;
}

#endif /* ORACLE_DRAW_SOURCE_H */
