/*
 * seriation_oracle.h -- CPU restatement of the reference sampler's hot path.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load liboracle.so.
 * The product path (libseriation_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py drives this restatement
 * with tapes recorded from the unmodified reference (oracle/_ref/ref_mcmc) and
 * requires bit-identical a, b, pi, rpi, counts, c, d and loglik after every
 * mcmc_sample() call and after every sub-sampler call; committed fixtures under
 * tests/golden/ hold the same comparison for boxes without /root/reference.
 * The per-taxon c, d variant (manycd = 1) is pinned the same way (ref_mcmc trace ... manycd,
 * tests/golden/ref_g10s10_manycd.npz).  Unpinned: GSL's MT19937 stream itself (no GSL here).
 */
#ifndef SERIATION_ORACLE_H
#define SERIATION_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;

/* X: N*M row-major 0/1 bytes (row = site, column = taxon); hard: N flags. */
orc_model *orc_create(int N, int M, const uint8_t *X, const uint8_t *hard);
void orc_free(orc_model *x);

/* draw source (see oracle/draw_source.h) */
void orc_source_mt(orc_model *x, unsigned long seed);
void orc_source_philox(orc_model *x, uint32_t seed, uint32_t chain);
void orc_source_tape(orc_model *x, const double *tape, size_t len); /* borrowed pointer */
void orc_record(orc_model *x, int on);
size_t orc_tape_len(const orc_model *x);              /* recorded slots */
void orc_tape_copy(const orc_model *x, double *out);  /* copy recorded tape */
long long orc_tape_slots(const orc_model *x);         /* slots consumed so far */

/* 0: libm log/exp exactly like the reference (default); 1: ser_detmath.h */
void orc_set_detmath(orc_model *x, int on);
/* per-taxon c, d (mcmc_readmodel's manycd, mcmc.c:363, :777-785, :807-815); set before the draw source */
void orc_set_manycd(orc_model *x, int on);
void orc_get_cd(const orc_model *x, double *c, double *d);

void orc_randomize(orc_model *x);     /* mcmc.c:477-578 */
int orc_samplec(orc_model *x);        /* mcmc.c:768-795 */
int orc_sampled(orc_model *x);        /* mcmc.c:798-825 */
int orc_sampleab(orc_model *x);       /* mcmc.c:918-996 */
int orc_samplepi1(orc_model *x);      /* mcmc.c:1127-1308 */
int orc_samplepi2(orc_model *x, int swap); /* mcmc.c:1311-1486 */
int orc_samplepi3(orc_model *x);      /* mcmc.c:1489-1682 */
int orc_sweep(orc_model *x);          /* one iteration of mcmc.c:225-244 */
int orc_sample(orc_model *x);         /* mcmc.c:214-258: 10 sweeps */
int orc_consistent(orc_model *x);     /* mcmc.c:999-1094; 0 == consistent */
void orc_recount(orc_model *x);       /* mcmc.c:651-708 + 625-648 */

/* state export: any pointer may be NULL */
void orc_get_state(const orc_model *x, int32_t *a, int32_t *b, int32_t *pi, int32_t *rpi,
                   int32_t *t0, int32_t *f0, int32_t *t1, int32_t *f1, int32_t tot[4],
                   double cdl[3]);
/* overwrite the chain state (pi, a, b, c, d) and recount; for unit tests */
void orc_set_state(orc_model *x, const int32_t *a, const int32_t *b, const int32_t *pi, double c,
                   double d);

/*
 * Run `burn_calls` + `sample_calls` mcmc_sample() calls.  After every sampling
 * call the state is appended to the caller's arrays (row-major, one row per
 * sample): a,b [samples][M], pi [samples][N], cdl [samples][3] = c,d,loglik,
 * counts [samples][4] = t0a,f0a,t1a,f1a.  sums[3] = sum(-loglik), sum(exp c),
 * sum(exp d) exactly as compute_exp_data (mcmc.c:53-58) accumulates them.
 */
void orc_run(orc_model *x, int burn_calls, int sample_calls, int32_t *a, int32_t *b, int32_t *pi,
             double *cdl, int32_t *counts, double sums[3]);

/* decision-margin audit (smallest distance of a draw from a decision boundary) */
void orc_margins(const orc_model *x, double *min_pick, double *min_accept,
                 long long *n_degenerate, long long *n_proposals);

#ifdef __cplusplus
}
#endif
#endif
