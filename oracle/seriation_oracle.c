/*
 * seriation_oracle.c -- CPU restatement of the seriation sampler's hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see seriation_oracle.h).  Written from scratch on
 * plain arrays; every function cites the reference lines it restates
 * (/root/reference/C_Implementation/mcmc.c).  Floating-point expressions keep
 * the reference's operand order and are compiled with -ffp-contract=off, so
 * with the same draw tape the state is bit-identical to the reference's
 * (pinned by tests/test_oracle_vs_ref.py and tests/golden/).
 */
#include "seriation_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "draw_source.h"

/* mcmc.h:25-30 */
#define ORC_LOGEPSILON (-32.236191301916641) /* log(1e-14) */
#define ORC_MINC (-6.9077552789821368)       /* log(.001)  */
#define ORC_MAXC (-2.3025850929940455)       /* log(.1)    */
#define ORC_MIND (-1.6094379124341003)       /* log(.2)    */
#define ORC_MAXD (-0.22314355131420971)      /* log(.8)    */

struct orc_model {
  int N, M, nh;
  uint8_t *X; /* N*M, row = site */
  uint8_t *h; /* hard-site flags, file order */
  int *a, *b; /* taxon m alive at positions a[m] <= pos < b[m] */
  int *pi;    /* pi[site] = position */
  int *rpi;   /* rpi[position] = site */
  int *t0, *f0, *t1, *f1;
  int t0a, f0a, t1a, f1a;
  double *c, *d; /* log P(false 1), log P(false 0) per taxon; all equal unless manycd (mcmc.h:39-40) */
  int manycd;
  double loglik;
  draw_source src;
  int detmath;
  /* scratch */
  uint8_t *v;
  double *q;
  int *dt0, *df0, *dt1, *df1, *p;
  /* audit */
  double min_pick, min_accept;
  long long n_degenerate, n_proposals, n_tape_mismatch;
};

void orc_initab(orc_model *x);

static void *xcalloc(size_t n, size_t sz)
{
  void *p = calloc(n ? n : 1, sz);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); exit(2); }
  return p;
}

static double orc_log(const orc_model *x, double v) { return x->detmath ? ser_log(v) : log(v); }
/* log(1 - e^v): the "true" side of a (c or d) log-probability */
static double orc_log1mexp(const orc_model *x, double v)
{
  return x->detmath ? ser_log(1. - ser_exp(v)) : log(1. - exp(v));
}

static void invert(const int *p, int *inv, int n)
{
  for (int i = 0; i < n; i++) inv[p[i]] = i;
}

orc_model *orc_create(int N, int M, const uint8_t *X, const uint8_t *hard)
{
  orc_model *x = (orc_model *)xcalloc(1, sizeof(*x));
  x->N = N; x->M = M;
  x->X = (uint8_t *)xcalloc((size_t)N * M, 1);
  x->h = (uint8_t *)xcalloc(N, 1);
  for (size_t i = 0; i < (size_t)N * M; i++) x->X[i] = X[i] ? 1 : 0;
  for (int n = 0; n < N; n++) { x->h[n] = hard && hard[n] ? 1 : 0; x->nh += x->h[n]; }
  x->a = (int *)xcalloc(M, sizeof(int)); x->b = (int *)xcalloc(M, sizeof(int));
  x->pi = (int *)xcalloc(N, sizeof(int)); x->rpi = (int *)xcalloc(N, sizeof(int));
  x->t0 = (int *)xcalloc(M, sizeof(int)); x->f0 = (int *)xcalloc(M, sizeof(int));
  x->t1 = (int *)xcalloc(M, sizeof(int)); x->f1 = (int *)xcalloc(M, sizeof(int));
  x->v = (uint8_t *)xcalloc(N, 1);
  x->q = (double *)xcalloc(N + 1, sizeof(double));
  x->dt0 = (int *)xcalloc(N + 1, sizeof(int)); x->df0 = (int *)xcalloc(N + 1, sizeof(int));
  x->dt1 = (int *)xcalloc(N + 1, sizeof(int)); x->df1 = (int *)xcalloc(N + 1, sizeof(int));
  x->p = (int *)xcalloc(N, sizeof(int));
  x->c = (double *)xcalloc(M, sizeof(double)); x->d = (double *)xcalloc(M, sizeof(double));
  ds_init_mt(&x->src, 0);
  x->min_pick = x->min_accept = INFINITY;
  /* mcmc_readmodel, mcmc.c:405-433: identity order, a/b from the data,
   * c = log .01, d = log .3, counts, likelihood */
  for (int n = 0; n < N; n++) x->pi[n] = x->rpi[n] = n;
  for (int m = 0; m < M; m++) { x->c[m] = log(.01); x->d[m] = log(.3); }
  orc_initab(x);
  orc_recount(x);
  return x;
}

void orc_free(orc_model *x)
{
  if (!x) return;
  free(x->X); free(x->h); free(x->a); free(x->b); free(x->pi); free(x->rpi);
  free(x->t0); free(x->f0); free(x->t1); free(x->f1);
  free(x->v); free(x->q); free(x->dt0); free(x->df0); free(x->dt1); free(x->df1); free(x->p); free(x->c); free(x->d);
  free(x->src.rec);
  free(x);
}

void orc_source_mt(orc_model *x, unsigned long seed) { free(x->src.rec); ds_init_mt(&x->src, seed); }
void orc_source_philox(orc_model *x, uint32_t seed, uint32_t chain)
{
  free(x->src.rec);
  ds_init_philox(&x->src, seed, chain);
  if (x->manycd) ds_set_manycd(&x->src, x->M);
}
void orc_source_tape(orc_model *x, const double *tape, size_t len) { free(x->src.rec); ds_init_tape(&x->src, tape, len); }
void orc_record(orc_model *x, int on) { x->src.recording = on; }
size_t orc_tape_len(const orc_model *x) { return x->src.rec_n; }
void orc_tape_copy(const orc_model *x, double *out) { memcpy(out, x->src.rec, x->src.rec_n * sizeof(double)); }
long long orc_tape_slots(const orc_model *x)
{
  return (long long)(x->src.n_uniform + x->src.n_pos + x->src.n_int + 3 * x->src.n_beta);
}
void orc_set_detmath(orc_model *x, int on) { x->detmath = on; }
/* per-taxon c, d (mcmc_readmodel's manycd argument, mcmc.c:363); call before choosing the draw source */
void orc_set_manycd(orc_model *x, int on)
{
  x->manycd = on;
  if (on && x->src.kind == DS_PHILOX) ds_set_manycd(&x->src, x->M);
}

/* mcmc_initab, mcmc.c:440-474: a = first position holding a 1, b = last + 1;
 * an all-zero column spans everything. */
void orc_initab(orc_model *x)
{
  for (int m = 0; m < x->M; m++) {
    int first = -1, last = -1;
    for (int pos = 0; pos < x->N; pos++)
      if (x->X[(size_t)x->rpi[pos] * x->M + m]) { if (first < 0) first = pos; last = pos; }
    if (first < 0) { x->a[m] = 0; x->b[m] = x->N; }
    else { x->a[m] = first; x->b[m] = last + 1; }
  }
}

/* mcmc_logl, mcmc.c:625-648 (operand order of the 4-term sum kept) */
static double orc_logl(const orc_model *x)
{
  double loglik = 0.;
  for (int m = 0; m < x->M; m++) {
    const double c = x->c[m], d = x->d[m];
    loglik += x->t0[m] * orc_log1mexp(x, c) + x->f0[m] * d + x->t1[m] * orc_log1mexp(x, d) + x->f1[m] * c;
  }
  return loglik;
}

/* mcmc_count01, mcmc.c:651-708 */
static void orc_count01(orc_model *x)
{
  x->t0a = x->f0a = x->t1a = x->f1a = 0;
  for (int m = 0; m < x->M; m++) {
    int t0 = 0, f0 = 0, t1 = 0, f1 = 0;
    for (int n = 0; n < x->N; n++) {
      const int alive = x->a[m] <= x->pi[n] && x->pi[n] < x->b[m];
      const int one = x->X[(size_t)n * x->M + m];
      if (alive) { if (one) t1++; else f0++; }
      else { if (one) f1++; else t0++; }
    }
    x->t0[m] = t0; x->f0[m] = f0; x->t1[m] = t1; x->f1[m] = f1;
    x->t0a += t0; x->f0a += f0; x->t1a += t1; x->f1a += f1;
  }
}

void orc_recount(orc_model *x)
{
  orc_count01(x);
  x->loglik = orc_logl(x);
}

/* mcmc_randomize, mcmc.c:477-578, with gsl_ran_choose / gsl_ran_shuffle
 * spelled out in draws (oracle/gsl_shim/shim.c has the same two loops). */
void orc_randomize(orc_model *x)
{
  const int N = x->N, nh = x->nh;
  if (nh == 0) { /* :486-494 -- pi shuffled, a/b deliberately NOT re-initialised */
    for (int i = N - 1; i > 0; i--) {
      int j = (int)ds_uniform_int(&x->src, (unsigned long)i + 1);
      int t = x->pi[i]; x->pi[i] = x->pi[j]; x->pi[j] = t;
    }
    invert(x->pi, x->rpi, N);
    orc_recount(x);
    return;
  }
  if (nh == N) return; /* :495-498 */

  int *rest = (int *)xcalloc(N, sizeof(int)), *chosen = (int *)xcalloc(nh, sizeof(int));
  /* :519 choose nh of the N positions, in increasing order */
  int j = 0;
  for (int i = 0; i < N && j < nh; i++)
    if ((double)(N - i) * ds_uniform(&x->src) < (double)(nh - j)) chosen[j++] = i;
  /* :528-538 the remaining positions */
  int k = 0;
  j = 0;
  for (int i = 0; i < N; i++) {
    if (j < nh && i == chosen[j]) j++;
    else rest[k++] = i;
  }
  /* :548 shuffle them */
  for (int i = N - nh - 1; i > 0; i--) {
    int r = (int)ds_uniform_int(&x->src, (unsigned long)i + 1);
    int t = rest[i]; rest[i] = rest[r]; rest[r] = t;
  }
  /* :557-563 hard sites take the chosen positions in file order */
  j = k = 0;
  for (int i = 0; i < N; i++) x->pi[i] = x->h[i] ? chosen[j++] : rest[k++];
  invert(x->pi, x->rpi, N);
  free(rest); free(chosen);
  orc_initab(x);
  orc_recount(x);
}

/* mcmc_samplebeta, mcmc.c:751-765: a direct conditional draw; an out-of-range
 * value leaves the old one in place (no redraw). */
static void orc_samplebeta(orc_model *x, double *val, double a, double b, double low, double high)
{
  double tape_ly, tape_l1;
  double y = ds_beta(&x->src, 1. + a, 1. + b, &tape_ly, &tape_l1);
  if (y > 0.) {
    double ly = orc_log(x, y);
    if (!x->detmath && x->src.kind == DS_TAPE) { /* the tape's libm companions must be ours */
      if (memcmp(&ly, &tape_ly, 8) != 0) x->n_tape_mismatch++;
      double l1 = log(1. - exp(ly));
      if (memcmp(&l1, &tape_l1, 8) != 0) x->n_tape_mismatch++;
    }
    if (low <= ly && ly <= high) *val = ly;
  }
}

int orc_samplec(orc_model *x) /* mcmc.c:768-795 */
{
  if (x->manycd) { /* :777-785 one Beta per taxon from its own counts */
    for (int m = 0; m < x->M; m++) orc_samplebeta(x, &x->c[m], x->f1[m], x->t0[m], ORC_MINC, ORC_MAXC);
    return x->M;
  }
  double y = x->c[0];
  orc_samplebeta(x, &y, x->f1a, x->t0a, ORC_MINC, ORC_MAXC);
  for (int m = 0; m < x->M; m++) x->c[m] = y;
  return 1;
}

int orc_sampled(orc_model *x) /* mcmc.c:798-825 */
{
  if (x->manycd) {
    for (int m = 0; m < x->M; m++) orc_samplebeta(x, &x->d[m], x->f0[m], x->t1[m], ORC_MIND, ORC_MAXD);
    return x->M;
  }
  double y = x->d[0];
  orc_samplebeta(x, &y, x->f0a, x->t1a, ORC_MIND, ORC_MAXD);
  for (int m = 0; m < x->M; m++) x->d[m] = y;
  return 1;
}

/* mcmc_logtop + mcmc_randompick, mcmc.c:711-748 and :901-915 */
static int orc_softmax_pick(orc_model *x, double *q, int n)
{
  double z = q[0], sum = 0.;
  for (int i = 1; i < n; i++) if (q[i] > z) z = q[i];
  for (int i = 0; i < n; i++) {
    double e = q[i] - z;
    double y = exp(ORC_LOGEPSILON > e ? ORC_LOGEPSILON : e);
    q[i] = y;
    sum += y;
  }
  for (int i = 0; i < n; i++) q[i] = q[i] / sum;
  int i = 0;
  double r = ds_uniform(&x->src) - q[0], prev = INFINITY;
  while (r > 0. && i < n - 1) { prev = r; r -= q[++i]; }
  { /* audit: distance of the draw from the two neighbouring CDF steps */
    double m = fabs(r) < prev ? fabs(r) : prev;
    if (i < n - 1 && m < x->min_pick) x->min_pick = m;
  }
  return i;
}

/*
 * mcmc_auxa, mcmc.c:828-898.  `v` is the taxon's column in position order,
 * candidates are 0..bound; counts are expressed as changes relative to the
 * current boundary *cur (running prefix, both directions).
 */
static void orc_gibbs_boundary(orc_model *x, const uint8_t *v, int bound, double c, double d,
                               int *cur, int *t0, int *f0, int *t1, int *f1)
{
  int *dt0 = x->dt0, *df0 = x->df0, *dt1 = x->dt1, *df1 = x->df1;
  const int a = *cur;
  dt0[a] = df0[a] = dt1[a] = df1[a] = 0;
  const double cc = orc_log1mexp(x, c), dd = orc_log1mexp(x, d);
  for (int i = a - 1; i >= 0; i--) { /* boundary moves down: cell i becomes alive */
    const int one = v[i];
    dt0[i] = dt0[i + 1] - !one; df0[i] = df0[i + 1] + !one;
    dt1[i] = dt1[i + 1] + one;  df1[i] = df1[i + 1] - one;
  }
  for (int i = a + 1; i <= bound; i++) { /* boundary moves up: cell i-1 dies */
    const int one = v[i - 1];
    dt0[i] = dt0[i - 1] + !one; df0[i] = df0[i - 1] - !one;
    dt1[i] = dt1[i - 1] - one;  df1[i] = df1[i - 1] + one;
  }
  for (int i = 0; i <= bound; i++)
    x->q[i] = dt0[i] * cc + df0[i] * d + dt1[i] * dd + df1[i] * c;
  const int pick = orc_softmax_pick(x, x->q, bound + 1);
  *cur = pick;
  *t0 += dt0[pick]; *f0 += df0[pick]; *t1 += dt1[pick]; *f1 += df1[pick];
}

/* mcmc_sampleab, mcmc.c:918-996 */
int orc_sampleab(orc_model *x)
{
  const int N = x->N;
  int changed = 0;
  for (int m = 0; m < x->M; m++) {
    for (int pos = 0; pos < N; pos++) x->v[pos] = x->X[(size_t)x->rpi[pos] * x->M + m];
    int t = x->a[m], t0 = x->t0[m], f0 = x->f0[m], t1 = x->t1[m], f1 = x->f1[m];
    orc_gibbs_boundary(x, x->v, x->b[m], x->c[m], x->d[m], &t, &t0, &f0, &t1, &f1);
    if (t != x->a[m]) { x->a[m] = t; changed++; }
    for (int i = 0; i < N / 2; i++) { uint8_t s = x->v[i]; x->v[i] = x->v[N - 1 - i]; x->v[N - 1 - i] = s; }
    t = N - x->b[m];
    orc_gibbs_boundary(x, x->v, N - x->a[m], x->c[m], x->d[m], &t, &t0, &f0, &t1, &f1);
    x->t0[m] = t0; x->f0[m] = f0; x->t1[m] = t1; x->f1[m] = f1;
    if (t != N - x->b[m]) { x->b[m] = N - t; changed++; }
  }
  x->t0a = x->f0a = x->t1a = x->f1a = 0;
  for (int m = 0; m < x->M; m++) {
    x->t0a += x->t0[m]; x->f0a += x->f0[m]; x->t1a += x->t1[m]; x->f1a += x->f1[m];
  }
  x->loglik = orc_logl(x);
  return changed;
}

/* mcmc_ininterval, mcmc.c:1097-1124 (lo <= hi at every call site) */
static int in_window(int v, int lo, int hi, int inc_lo, int inc_hi)
{
  return (inc_lo ? lo <= v : lo < v) && (inc_hi ? v <= hi : v < hi);
}

/* the 4-term per-taxon likelihood change, operand order of mcmc.c:1214 */
static double orc_term(const orc_model *x, int m, int dt0, int df0, int dt1, int df1)
{
  const double c = x->c[m], d = x->d[m];
  return dt0 * orc_log1mexp(x, c) + df0 * d + dt1 * orc_log1mexp(x, d) + df1 * c;
}

/* common MH tail: mcmc.c:1261, :1441, :1636 */
static int orc_accept(orc_model *x, double delta, long sdt0, long sdt1, int any)
{
  x->n_proposals++;
  if (sdt0 == 0 && sdt1 == 0 && any && delta != 0.) x->n_degenerate++;
  if (delta >= 0.) return 1;
  double lu = log(ds_uniform_pos(&x->src));
  double m = fabs(delta - lu);
  if (m < x->min_accept) x->min_accept = m;
  return delta > lu;
}

static void orc_after_move(orc_model *x, double delta)
{
  invert(x->rpi, x->pi, x->N);
  x->loglik += delta;
  orc_count01(x); /* the reference recounts from scratch: mcmc.c:1303,1480,1676 */
}

/* mcmc_samplepi1, mcmc.c:1127-1308: move the site at position i to position j */
int orc_samplepi1(orc_model *x)
{
  const int N = x->N, M = x->M;
  int i = (int)ds_uniform_int(&x->src, N);
  int j = (int)ds_uniform_int(&x->src, N - 1);
  if (j >= i) j++;
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  if (x->h[x->rpi[i]]) { /* :1160-1169 a hard site may not pass another hard site */
    int cnt = 0;
    for (int n = lo; n <= hi; n++) { cnt += x->h[x->rpi[n]]; if (cnt > 1) return 0; }
  }
  const uint8_t *row = x->X + (size_t)x->rpi[i] * M;
  double delta = 0.;
  long s0 = 0, s1 = 0; int any = 0;
  for (int m = 0; m < M; m++) {
    int dt0 = 0, df0 = 0, dt1 = 0, df1 = 0;
    const int a = x->a[m], b = x->b[m];
    int gain, lose; /* does the moved site become alive / dead for taxon m */
    if (i < j) { /* :1177-1216 */
      const int ain = lo < a && a <= hi + 1, bin = lo < b && b <= hi + 1;
      gain = ain && !bin; lose = !ain && bin;
    } else { /* :1217-1256 */
      const int ain = lo <= a && a <= hi, bin = lo <= b && b <= hi;
      gain = !ain && bin; lose = ain && !bin;
    }
    if (gain) { if (row[m]) { dt1++; df1--; } else { dt0--; df0++; } }
    else if (lose) { if (row[m]) { dt1--; df1++; } else { dt0++; df0--; } }
    delta += orc_term(x, m, dt0, df0, dt1, df1);
    s0 += dt0; s1 += dt1; any |= (dt0 | dt1) != 0;
  }
  if (!orc_accept(x, delta, s0, s1, any)) return 0;
  if (i < j) { /* :1266-1281 */
    for (int m = 0; m < M; m++) {
      if (lo < x->a[m] && x->a[m] <= hi + 1) x->a[m]--;
      if (lo < x->b[m] && x->b[m] <= hi + 1) x->b[m]--;
    }
    const int t = x->rpi[i];
    for (int n = i; n < j; n++) x->rpi[n] = x->rpi[n + 1];
    x->rpi[j] = t;
  } else { /* :1282-1297 */
    for (int m = 0; m < M; m++) {
      if (lo <= x->a[m] && x->a[m] <= hi) x->a[m]++;
      if (lo <= x->b[m] && x->b[m] <= hi) x->b[m]++;
    }
    const int t = x->rpi[i];
    for (int n = i; n > j; n--) x->rpi[n] = x->rpi[n - 1];
    x->rpi[j] = t;
  }
  orc_after_move(x, delta);
  return 1;
}

/* mirror rule shared by pi2 and pi3: mcmc.c:1446-1465 / :1576-1595 / :1641-1660 */
static void mirror_ab(int a, int b, int ain, int bin, int s, int *na, int *nb)
{
  *na = a; *nb = b;
  if (ain && !bin) *na = s - a;
  else if (!ain && bin) *nb = s - b;
  else if (ain && bin) { *nb = s - a; *na = s - b; }
}

/* mcmc_samplepi2, mcmc.c:1311-1486: reverse positions [i, j] */
int orc_samplepi2(orc_model *x, int swap)
{
  const int N = x->N, M = x->M;
  int i, j;
  if (!swap) { /* :1323-1337 ordered distinct pair */
    i = (int)ds_uniform_int(&x->src, N);
    j = (int)ds_uniform_int(&x->src, N - 1);
    if (j >= i) j++;
    else { int t = i; i = j; j = t; }
  } else { /* :1338-1342 */
    i = (int)ds_uniform_int(&x->src, N - 1);
    j = i + 1;
  }
  { /* :1348-1354 forbidden iff two or more hard sites inside */
    int cnt = 0;
    for (int n = i; n <= j; n++) { cnt += x->h[x->rpi[n]]; if (cnt > 1) return 0; }
  }
  const int inc1 = (int)ds_uniform_int(&x->src, 2), inc2 = (int)ds_uniform_int(&x->src, 2);
  double delta = 0.;
  long s0 = 0, s1 = 0; int any = 0;
  for (int m = 0; m < M; m++) { /* :1368-1436 */
    int dt0 = 0, df0 = 0, dt1 = 0, df1 = 0;
    const int a = x->a[m], b = x->b[m];
    const int ain = in_window(a, i, j + 1, inc1, inc2), bin = in_window(b, i, j + 1, inc1, inc2);
    if (ain != bin) {
      /* split point: cells [i,split) and [split,j] swap roles */
      const int split = ain ? a : b;
      for (int n = i; n <= j; n++) {
        const int one = x->X[(size_t)x->rpi[n] * M + m];
        const int gains = ain ? (n < split) : (n >= split);
        if (gains) { if (one) { dt1++; df1--; } else { dt0--; df0++; } }
        else { if (one) { dt1--; df1++; } else { dt0++; df0--; } }
      }
    }
    delta += orc_term(x, m, dt0, df0, dt1, df1);
    s0 += dt0; s1 += dt1; any |= (dt0 | dt1) != 0;
  }
  if (!orc_accept(x, delta, s0, s1, any)) return 0;
  for (int m = 0; m < M; m++) {
    const int a = x->a[m], b = x->b[m];
    const int ain = in_window(a, i, j + 1, inc1, inc2), bin = in_window(b, i, j + 1, inc1, inc2);
    mirror_ab(a, b, ain, bin, i + j + 1, &x->a[m], &x->b[m]);
  }
  for (int lo = i, hi = j; lo < hi; lo++, hi--) { int t = x->rpi[lo]; x->rpi[lo] = x->rpi[hi]; x->rpi[hi] = t; }
  orc_after_move(x, delta);
  return 1;
}

/* mcmc_samplepi3, mcmc.c:1489-1682: reverse only the non-hard sites of a window */
int orc_samplepi3(orc_model *x)
{
  const int N = x->N, M = x->M, free_sites = N - x->nh;
  if (free_sites < 2) return 0; /* :1502-1503, no draws */
  int r1 = (int)ds_uniform_int(&x->src, free_sites);
  int r2 = (int)ds_uniform_int(&x->src, free_sites - 1);
  int i, j;
  if (r1 <= r2) { i = r1; j = r2 + 1; } else { i = r2; j = r1; }
  { /* :1518-1533 ranks among non-hard sites -> positions */
    int n = 0;
    while (n <= i) { if (x->h[x->rpi[n]]) { i++; j++; } n++; }
    while (n <= j) { if (x->h[x->rpi[n]]) j++; n++; }
  }
  int *p = x->p;
  for (int lo = i, hi = j; lo <= hi;) { /* :1534-1555 involution: hard fixed, the rest mirrored */
    if (x->h[x->rpi[lo]]) { p[lo] = lo; lo++; }
    else if (x->h[x->rpi[hi]]) { p[hi] = hi; hi--; }
    else { p[lo] = hi; p[hi] = lo; lo++; hi--; }
  }
  const int inc1 = (int)ds_uniform_int(&x->src, 2), inc2 = (int)ds_uniform_int(&x->src, 2);
  double delta = 0.;
  long s0 = 0, s1 = 0; int any = 0;
  for (int m = 0; m < M; m++) { /* :1569-1631 */
    int dt0 = 0, df0 = 0, dt1 = 0, df1 = 0, na, nb;
    const int a = x->a[m], b = x->b[m];
    const int ain = in_window(a, i, j + 1, inc1, inc2), bin = in_window(b, i, j + 1, inc1, inc2);
    mirror_ab(a, b, ain, bin, i + j + 1, &na, &nb);
    for (int n = i; n <= j; n++) {
      const int dest = p[n];
      const int was = a <= n && n < b, is = na <= dest && dest < nb;
      if (was == is) continue;
      const int one = x->X[(size_t)x->rpi[n] * M + m];
      if (was) { if (one) { dt1--; df1++; } else { df0--; dt0++; } }
      else { if (one) { dt1++; df1--; } else { df0++; dt0--; } }
    }
    delta += orc_term(x, m, dt0, df0, dt1, df1);
    s0 += dt0; s1 += dt1; any |= (dt0 | dt1) != 0;
  }
  if (!orc_accept(x, delta, s0, s1, any)) return 0;
  for (int m = 0; m < M; m++) {
    const int a = x->a[m], b = x->b[m];
    const int ain = in_window(a, i, j + 1, inc1, inc2), bin = in_window(b, i, j + 1, inc1, inc2);
    mirror_ab(a, b, ain, bin, i + j + 1, &x->a[m], &x->b[m]);
  }
  for (int n = i; n <= j; n++) p[n] = x->rpi[p[n]]; /* :1664-1670 */
  for (int n = i; n <= j; n++) x->rpi[n] = p[n];
  orc_after_move(x, delta);
  return 1;
}

/* one iteration of the loop body mcmc.c:225-244 */
int orc_sweep(orc_model *x)
{
  int acc = 0;
  acc += orc_samplec(x);
  acc += orc_sampled(x);
  acc += orc_sampleab(x);
  acc += orc_samplepi2(x, 1);
  for (int j = 0; j < 5; j++) {
    acc += orc_samplepi1(x);
    acc += orc_samplepi2(x, 0);
    acc += orc_samplepi3(x);
  }
  return acc;
}

int orc_sample(orc_model *x) /* mcmc.c:214-258 */
{
  int acc = 0;
  for (int i = 0; i < 10; i++) acc += orc_sweep(x);
  return acc;
}

/* mcmc_consistent, mcmc.c:999-1094 */
int orc_consistent(orc_model *x)
{
  int bad = 0;
  for (int m = 0; m < x->M; m++)
    if (!(0 <= x->a[m] && x->a[m] <= x->b[m] && x->b[m] <= x->N)) bad |= 1;
  for (int n = 0; n < x->N; n++) {
    if (x->pi[n] < 0 || x->pi[n] >= x->N) { bad |= 2; continue; }
    if (x->rpi[x->pi[n]] != n) bad |= 2;
  }
  int last = -1, cnt = 0;
  for (int n = 0; n < x->N; n++)
    if (x->h[n]) { cnt++; if (last >= 0 && x->pi[n] < last) bad |= 4; last = x->pi[n]; }
  if (cnt != x->nh) bad |= 4;
  const int t0 = x->t0a, f0 = x->f0a, t1 = x->t1a, f1 = x->f1a;
  const double ll = x->loglik;
  orc_recount(x);
  if (t0 != x->t0a || f0 != x->f0a || t1 != x->t1a || f1 != x->f1a || fabs(ll - x->loglik) > 1e-8) bad |= 8;
  x->loglik = ll; /* keep the incrementally maintained value: it is what gets saved */
  return bad;
}

void orc_get_state(const orc_model *x, int32_t *a, int32_t *b, int32_t *pi, int32_t *rpi,
                   int32_t *t0, int32_t *f0, int32_t *t1, int32_t *f1, int32_t tot[4],
                   double cdl[3])
{
  for (int m = 0; m < x->M; m++) {
    if (a) a[m] = x->a[m];
    if (b) b[m] = x->b[m];
    if (t0) t0[m] = x->t0[m];
    if (f0) f0[m] = x->f0[m];
    if (t1) t1[m] = x->t1[m];
    if (f1) f1[m] = x->f1[m];
  }
  for (int n = 0; n < x->N; n++) {
    if (pi) pi[n] = x->pi[n];
    if (rpi) rpi[n] = x->rpi[n];
  }
  if (tot) { tot[0] = x->t0a; tot[1] = x->f0a; tot[2] = x->t1a; tot[3] = x->f1a; }
  if (cdl) { cdl[0] = x->c[0]; cdl[1] = x->d[0]; cdl[2] = x->loglik; }
}

void orc_set_state(orc_model *x, const int32_t *a, const int32_t *b, const int32_t *pi, double c,
                   double d)
{
  for (int m = 0; m < x->M; m++) { x->a[m] = a[m]; x->b[m] = b[m]; }
  for (int n = 0; n < x->N; n++) x->pi[n] = pi[n];
  invert(x->pi, x->rpi, x->N);
  for (int m = 0; m < x->M; m++) { x->c[m] = c; x->d[m] = d; }
  orc_recount(x);
}

void orc_run(orc_model *x, int burn_calls, int sample_calls, int32_t *a, int32_t *b, int32_t *pi,
             double *cdl, int32_t *counts, double sums[3])
{
  double s_ll = 0., s_c = 0., s_d = 0.;
  for (int i = 0; i < burn_calls; i++) orc_sample(x);
  for (int s = 0; s < sample_calls; s++) {
    orc_sample(x);
    int32_t tot[4];
    double v[3];
    orc_get_state(x, a ? a + (size_t)s * x->M : NULL, b ? b + (size_t)s * x->M : NULL,
                  pi ? pi + (size_t)s * x->N : NULL, NULL, NULL, NULL, NULL, NULL, tot, v);
    if (cdl) memcpy(cdl + (size_t)s * 3, v, sizeof(v));
    if (counts) memcpy(counts + (size_t)s * 4, tot, sizeof(tot));
    /* compute_exp_data, mcmc.c:53-58 */
    s_ll += -(x->loglik);
    s_c += exp(x->c[0]);
    s_d += exp(x->d[0]);
  }
  if (sums) { sums[0] = s_ll; sums[1] = s_c; sums[2] = s_d; }
}

void orc_margins(const orc_model *x, double *min_pick, double *min_accept,
                 long long *n_degenerate, long long *n_proposals)
{
  if (min_pick) *min_pick = x->min_pick;
  if (min_accept) *min_accept = x->min_accept;
  if (n_degenerate) *n_degenerate = x->n_degenerate;
  if (n_proposals) *n_proposals = x->n_proposals;
}

long long orc_tape_mismatches(const orc_model *x) { return x->n_tape_mismatch; }

/* per-taxon c, d (any pointer may be NULL) */
void orc_get_cd(const orc_model *x, double *c, double *d)
{
  for (int m = 0; m < x->M; m++) { if (c) c[m] = x->c[m]; if (d) d[m] = x->d[m]; }
}
