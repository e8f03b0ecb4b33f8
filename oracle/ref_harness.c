/*
 * ref_harness.c -- driver around the UNMODIFIED reference sampler.
 *
 * TEST INFRASTRUCTURE ONLY.  Linked with /root/reference/C_Implementation/
 * mcmc.c (compiled where it lies with -Dmain=ref_main) and the GSL-API shim.
 * Everything here goes through the reference's public prototypes (mcmc.h:47-73).
 *
 *   ref_mcmc cli <chain_idx>                      the reference's own main()
 *        (stdin = dataset, writes Chains/chain_XX/, honours GSL_RNG_SEED)
 *   ref_mcmc trace <dataset> <burn_calls> <sample_calls> <dump_file> [step] [manycd]
 *        mcmc_init / readmodel / randomize, then <burn>+<sample> calls of
 *        mcmc_sample(); the full model state is appended to <dump_file> after
 *        randomize and after every call.  With "step" the 10-sweep body of
 *        mcmc_sample (mcmc.c:225-244) is unrolled here so the state is dumped
 *        after every sub-sampler call (debug granularity).  SER_TAPE_OUT
 *        records the draw tape (see draw_source.h).
 *   ref_mcmc bench <dataset> <sample_calls>       times mcmc_sample() only
 *
 * Dump file: int32 header {0x5345524d (0x5345524e with manycd), N, M, nh}; then records
 *   int32 kind, int32 ret, int64 tape_slots,
 *   int32 a[M], b[M], pi[N], rpi[N], t0[M], f0[M], t1[M], f1[M], tot[4],
 *   double c, d, loglik            (c[0], d[0]); with manycd followed by double c[M], d[M]
 * kind: 0 after randomize, 1 after mcmc_sample, 10 samplec, 11 sampled,
 *       12 sampleab, 13 pi2(swap), 14 pi1, 15 pi2(0), 16 pi3
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <gsl/gsl_matrix.h>
#include <gsl/gsl_permutation.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_vector.h>

#include "mcmc.h" /* the reference's own header, found via -I */

#include "draw_source.h"

extern int ref_main(int argc, char *argv[]);
extern draw_source *shim_source(void);
static int g_manycd = 0;

static void put_i32(FILE *f, int32_t v) { fwrite(&v, 4, 1, f); }
static void put_i64(FILE *f, int64_t v) { fwrite(&v, 8, 1, f); }
static void put_f64(FILE *f, double v) { fwrite(&v, 8, 1, f); }

static void dump_state(FILE *f, const mcmc_model *x, int kind, int ret)
{
  int i;
  draw_source *s = shim_source();
  put_i32(f, kind);
  put_i32(f, ret);
  put_i64(f, s ? (int64_t)(s->n_uniform + s->n_pos + s->n_int + 3 * s->n_beta) : -1);
  for (i = 0; i < x->M; i++) put_i32(f, gsl_vector_int_get(x->a, i));
  for (i = 0; i < x->M; i++) put_i32(f, gsl_vector_int_get(x->b, i));
  for (i = 0; i < x->N; i++) put_i32(f, (int32_t)gsl_permutation_get(x->pi, i));
  for (i = 0; i < x->N; i++) put_i32(f, (int32_t)gsl_permutation_get(x->rpi, i));
  for (i = 0; i < x->M; i++) put_i32(f, gsl_vector_int_get(x->t0, i));
  for (i = 0; i < x->M; i++) put_i32(f, gsl_vector_int_get(x->f0, i));
  for (i = 0; i < x->M; i++) put_i32(f, gsl_vector_int_get(x->t1, i));
  for (i = 0; i < x->M; i++) put_i32(f, gsl_vector_int_get(x->f1, i));
  put_i32(f, x->t0a); put_i32(f, x->f0a); put_i32(f, x->t1a); put_i32(f, x->f1a);
  put_f64(f, gsl_vector_get(x->c, 0));
  put_f64(f, gsl_vector_get(x->d, 0));
  put_f64(f, x->loglik);
  if (g_manycd) {
    for (i = 0; i < x->M; i++) put_f64(f, gsl_vector_get(x->c, i));
    for (i = 0; i < x->M; i++) put_f64(f, gsl_vector_get(x->d, i));
  }
}

/* mcmc_consistent (mcmc.c:1078-1080) replaces x->loglik by a fresh recount as a side effect;
 * keep the incrementally maintained value so that the check does not perturb the trace. */
static int check_consistent(mcmc_model *x)
{
  double keep = x->loglik;
  int rc = mcmc_consistent(x);
  x->loglik = keep;
  return rc;
}

static FILE *open_or_die(const char *path, const char *mode)
{
  FILE *f = fopen(path, mode);
  if (!f) { fprintf(stderr, "ref_mcmc: cannot open %s\n", path); exit(2); }
  return f;
}

static int do_trace(int argc, char **argv)
{
  mcmc_model x;
  FILE *fin, *fout;
  int burn, samp, call, step = 0, i, j;
  gsl_vector_int *p;
  if (argc < 6) return 64;
  burn = atoi(argv[3]);
  samp = atoi(argv[4]);
  for (i = 6; i < argc; i++) {
    if (strcmp(argv[i], "step") == 0) step = 1;
    if (strcmp(argv[i], "manycd") == 0) g_manycd = 1;
  }
  fin = open_or_die(argv[2], "r");
  fout = open_or_die(argv[5], "wb");

  mcmc_init();
  mcmc_readmodel(&x, fin, g_manycd);
  fclose(fin);
  if (g_manycd && shim_source()) ds_set_manycd(shim_source(), x.M);
  mcmc_randomize(&x);
  if (check_consistent(&x)) { fprintf(stderr, "ref_mcmc: inconsistent after randomize\n"); return 1; }

  put_i32(fout, g_manycd ? 0x5345524e : 0x5345524d); put_i32(fout, x.N); put_i32(fout, x.M); put_i32(fout, x.nh);
  dump_state(fout, &x, 0, 0);

  for (call = 0; call < burn + samp; call++) {
    if (!step) {
      int ret = mcmc_sample(&x);
      dump_state(fout, &x, 1, ret);
    } else {
      /* same call order as the body of mcmc_sample, mcmc.c:225-244 */
      p = gsl_vector_int_alloc(x.N);
      for (i = 0; i < 10; i++) {
        dump_state(fout, &x, 10, mcmc_samplec(&x));
        dump_state(fout, &x, 11, mcmc_sampled(&x));
        dump_state(fout, &x, 12, mcmc_sampleab(&x));
        dump_state(fout, &x, 13, mcmc_samplepi2(&x, 1));
        for (j = 0; j < 5; j++) {
          dump_state(fout, &x, 14, mcmc_samplepi1(&x));
          dump_state(fout, &x, 15, mcmc_samplepi2(&x, 0));
          dump_state(fout, &x, 16, mcmc_samplepi3(&x, p));
        }
      }
      gsl_vector_int_free(p);
      dump_state(fout, &x, 1, 0);
    }
    if (check_consistent(&x)) { fprintf(stderr, "ref_mcmc: inconsistent after call %d\n", call); return 1; }
  }
  fclose(fout);
  mcmc_freemodel(&x);
  mcmc_free(); /* gsl_rng_free writes SER_TAPE_OUT */
  return 0;
}

static int do_bench(int argc, char **argv)
{
  mcmc_model x;
  FILE *fin;
  int calls, i;
  struct timespec t0, t1;
  double dt;
  if (argc < 4) return 64;
  calls = atoi(argv[3]);
  fin = open_or_die(argv[2], "r");
  mcmc_init();
  mcmc_readmodel(&x, fin, 0);
  fclose(fin);
  mcmc_randomize(&x);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (i = 0; i < calls; i++) mcmc_sample(&x);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  if (mcmc_consistent(&x)) return 1;
  printf("{\"sweeps\": %d, \"seconds\": %.6f, \"sweeps_per_s\": %.3f, \"loglik\": %.9f}\n",
         calls * 10, dt, calls * 10 / dt, x.loglik);
  mcmc_freemodel(&x);
  mcmc_free();
  return 0;
}

int main(int argc, char **argv)
{
  if (argc >= 2 && strcmp(argv[1], "cli") == 0) {
    /* hand argv[2..] to the reference main as its argv[1..] */
    argv[1] = argv[0];
    return ref_main(argc - 1, argv + 1);
  }
  if (argc >= 2 && strcmp(argv[1], "trace") == 0) {
    int rc = do_trace(argc, argv);
    if (rc != 64) return rc;
  } else if (argc >= 2 && strcmp(argv[1], "bench") == 0) {
    int rc = do_bench(argc, argv);
    if (rc != 64) return rc;
  }
  fprintf(stderr,
          "usage: %s cli <chain_idx> < dataset.txt\n"
          "       %s trace <dataset> <burn_calls> <sample_calls> <dump_file> [step]\n"
          "       %s bench <dataset> <sample_calls>\n",
          argv[0], argv[0], argv[0]);
  return 64;
}
