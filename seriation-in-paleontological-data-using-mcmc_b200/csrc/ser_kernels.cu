/*
 * ser_kernels.cu -- sm_100a kernels of the seriation sweep and the run object behind the C ABI.
 *
 * Mapping: one CTA per chain, one thread per taxon (+ one thread owning the hard-site mask).
 * The chain's occurrence matrix lives in shared memory as position-ordered bit columns
 * V[word][column]; a/b of a taxon live in its thread's registers.  A sweep is
 *   stage draws -> c,d -> a/b Gibbs (per thread) -> 16 pi proposals (per-thread integer deltas,
 *   REDUX + one __syncthreads, redundant uniform decision, per-thread column update on accept).
 * No tensor cores, no global traffic inside a sweep except the draw tape (replay) and the
 * thinned samples.  See DESIGN.md for the layout and the roofline that bounds each phase.
 *
 * Reference: /root/reference/C_Implementation/mcmc.c (line numbers next to each device function).
 */
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>

#include "ser_chain_core.h"
#include "ser_internal.h"

#define SER_MINC (-6.9077552789821368)
#define SER_MAXC (-2.3025850929940455)
#define SER_MIND (-1.6094379124341003)
#define SER_MAXD (-0.22314355131420971)

#define SER_PI_DRAWS 72 /* >= 4 + 5*13 = 69 draws a sweep's pi part can consume */
#define SER_MAX_WARPS 32
#define SER_MAX_GROUPS 8 /* column groups the item weights of a Gibbs step are evaluated in */

/* ------------------------------------------------------------------ per-chain global state */
struct ChainScalars {
  double c, cc, d, dd; /* log P(false 1), log(1-e^c), log P(false 0), log(1-e^d) */
  double loglik;
  double sum_negll, sum_ec, sum_ed; /* compute_exp_data, mcmc.c:53-58 */
  long long cursor;                 /* replay: tape slots consumed */
  long long counters[8];            /* c, d, ab changed, pi1, pi2(0), pi2(swap), pi3, sweeps */
  int t0a, f0a, t1a, f1a;
  unsigned int sweep; /* free-running: sweep index = Philox counter word */
  int n_samples;
  int flags; /* bit0 tape exhausted, bit1.. consistency failures */
  int pad;
};

struct KParams {
  int N, M, W, C, nh, Mw, Npad, Mpad;
  const uint32_t *Xs;  /* [N][Mw] site-major bits */
  const uint8_t *hard; /* [N] file order */
  const int *ones;     /* [M] ones per column */
  const uint16_t *order;    /* [M] column -> taxon (columns are sorted by ones, descending) */
  const int *off;           /* [M+1] first item of each column; a column has ones+1 items */
  const uint32_t *item_col; /* [I] item -> (column << 16) | index of the item inside its column */
  int I;                    /* ones_total + M */
  /* large-shape path (ser_sweep_kernel_big): per-CTA-slot scratch in global memory */
  int Cs;                   /* column stride of the scratch bit matrix (>= M+1) */
  uint32_t *gV;             /* [slot][W][Cs] */
  uint16_t *gpre;           /* [slot][W+1][Cs] */
  const int *bgrp;          /* large-shape column groups: [g] = {first column, first item}, big_ng + 1 entries */
  int big_ng, big_icap, big_gcap;
  int n_chains;
  uint16_t *ab;        /* [chain][2][Mpad] */
  uint16_t *rpi;       /* [chain][Npad] */
  ChainScalars *scal;  /* [chain] */
  int mode, chain_offset;
  unsigned int seed;
  const double *tape;
  const unsigned long long *tape_off;
  int n_calls, sweeps_per_call, sampling;
  int store, max_samples;
  uint16_t *samp_a, *samp_b, *samp_pi;
  double *samp_cdl;
  double c0, cc0, d0, dd0, eps;
  long long ones_total;
  /* per-taxon c, d (manycd = 1, mcmc.c:777-785, :807-815) */
  int manycd;
  double *cd4;         /* [chain][4][Mpad]: c, log(1-e^c), d, log(1-e^d) per column */
  double *samp_cd_all; /* [chain][sample][2][M]: c, d per taxon (SER_STORE_FULL) */
  /* the item weights of a Gibbs step are evaluated group by group of columns through a buffer of
   * Ival doubles: a smaller buffer = more resident chains per SM */
  int n_groups, Ival;
  int grp_c[SER_MAX_GROUPS + 1], grp_e[SER_MAX_GROUPS + 1];
};

/* ------------------------------------------------------------------ shared-memory carve-up */
struct Smem {
  double *draws_pi; /* SER_PI_DRAWS */
  double *logdraw;  /* SER_PI_DRAWS: log of each staged draw (only read where a draw is a U+) */
  double *draws_cd; /* 8 */
  double *terms;    /* C */
  double *H;        /* N + 2: geometric partial sums of the current sweep (ser_h_entry) */
  uint32_t *V;      /* W*C */
  int *red;         /* 2 * SER_MAX_WARPS * 4 */
  double *val;      /* I+1: item weights of the running Gibbs step */
  double *lmax;     /* C: per-column maximum log-weight of the running step */
  uint16_t *pos;    /* I+1: ascending positions of the ones of every column (postings) */
  uint16_t *st4;    /* 4*C: per-column step geometry: cur, bound, ocur, kb */
  uint16_t *ones16; /* C: ones per column (static; keeps the dense item loop free of dependent global loads) */
  uint16_t *pre;    /* (W+1)*C: pre[w][col] = ones of the column in words < w */
  uint16_t *hp;     /* N+1: hard positions, ascending */
  double *wcol;     /* manycd only: 4*C per-column weights A, g, 1/g, 1/(1-e^-g) for the dense item phase */
  double *redd;     /* manycd only: 2*32 doubles of reduction scratch */
  uint16_t *rpi, *tmp16, *perm16; /* N each */
};

__host__ __device__ inline size_t smem_layout(Smem *s, unsigned char *base, int N, int W, int C, int I, int manycd = 0, int Ival = -1)
{
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  size_t o_ld = take(sizeof(double) * SER_PI_DRAWS);
  size_t o_dp = take(sizeof(double) * SER_PI_DRAWS), o_dc = take(sizeof(double) * 8), o_t = take(sizeof(double) * C);
  size_t o_H = take(sizeof(double) * (N + 2));
  size_t o_val = take(sizeof(double) * ((Ival < 0 ? I : Ival) + 1)), o_lm = take(sizeof(double) * C);
  size_t o_wc = take(manycd ? sizeof(double) * 4 * C : 0), o_rd = take(manycd ? sizeof(double) * 2 * SER_MAX_WARPS : 0);
  size_t o_pos = take(sizeof(uint16_t) * (I + 1)), o_st = take(sizeof(uint16_t) * 4 * C), o_on = take(sizeof(uint16_t) * C);
  size_t o_v = take(sizeof(uint32_t) * (size_t)W * C), o_r = take(sizeof(int) * 2 * SER_MAX_WARPS * 4);
  size_t o_h = take(sizeof(uint16_t) * (size_t)(W + 1) * C), o_hp = take(sizeof(uint16_t) * (N + 1));
  size_t o_p = take(sizeof(uint16_t) * N), o_q = take(sizeof(uint16_t) * N), o_m = take(sizeof(uint16_t) * N);
  if (s) {
    s->logdraw = (double *)(base + o_ld);
    s->draws_pi = (double *)(base + o_dp); s->draws_cd = (double *)(base + o_dc); s->terms = (double *)(base + o_t);
    s->H = (double *)(base + o_H);
    s->val = (double *)(base + o_val); s->lmax = (double *)(base + o_lm);
    s->wcol = (double *)(base + o_wc); s->redd = (double *)(base + o_rd);
    s->pos = (uint16_t *)(base + o_pos); s->st4 = (uint16_t *)(base + o_st); s->ones16 = (uint16_t *)(base + o_on);
    s->V = (uint32_t *)(base + o_v); s->red = (int *)(base + o_r); s->pre = (uint16_t *)(base + o_h); s->hp = (uint16_t *)(base + o_hp);
    s->rpi = (uint16_t *)(base + o_p); s->tmp16 = (uint16_t *)(base + o_q); s->perm16 = (uint16_t *)(base + o_m);
  }
  return off;
}

/* ------------------------------------------------------------------ block helpers */
/* sum of three ints over the CTA; every thread gets the totals.  One __syncthreads; `buf`
 * alternates between calls so a warp that runs ahead never overwrites live partials. */
__device__ __forceinline__ void block_sum3(int v0, int v1, int v2, int *red, int &buf, int *o0, int *o1, int *o2)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v0 = __reduce_add_sync(0xffffffffu, v0);
  v1 = __reduce_add_sync(0xffffffffu, v1);
  v2 = __reduce_add_sync(0xffffffffu, v2);
  int *r = red + buf * (SER_MAX_WARPS * 4);
  if (lane == 0) { r[warp * 4 + 0] = v0; r[warp * 4 + 1] = v1; r[warp * 4 + 2] = v2; }
  __syncthreads();
  /* second level: lane w picks up warp w's partials, one more REDUX per value */
  int s0 = 0, s1 = 0, s2 = 0;
  if (lane < nwarp) { s0 = r[lane * 4 + 0]; s1 = r[lane * 4 + 1]; s2 = r[lane * 4 + 2]; }
  s0 = __reduce_add_sync(0xffffffffu, s0);
  s1 = __reduce_add_sync(0xffffffffu, s1);
  s2 = __reduce_add_sync(0xffffffffu, s2);
  buf ^= 1;
  *o0 = s0; *o1 = s1; *o2 = s2;
}

/* position-ordered columns from the site-major data and rpi; column M = hard mask */
__device__ void build_columns(const KParams &p, const Smem &sm)
{
  const int tid = threadIdx.x, C = p.C;
  const int mw = tid >> 5, mb = tid & 31;
  for (int w = 0; w < p.W; w++) {
    uint32_t word = 0;
    const int pend = min(32 * w + 32, p.N);
    if (tid < p.M) {
      for (int pos = 32 * w; pos < pend; pos++)
        word |= ((p.Xs[(size_t)sm.rpi[pos] * p.Mw + mw] >> mb) & 1u) << (pos & 31);
    } else if (tid == p.M) {
      for (int pos = 32 * w; pos < pend; pos++) word |= (uint32_t)(p.hard[sm.rpi[pos]] != 0) << (pos & 31);
    }
    sm.V[w * C + tid] = word;
  }
  ser_col_build_pre(sm.V + tid, sm.pre + tid, C, p.W);
}

/* sorted hard positions from the hard-mask column (its owner thread, tid == M) */
__device__ void rebuild_hard(const KParams &p, const Smem &sm) { ser_hard_list(sm.V + p.M, p.C, p.W, sm.hp); }

__device__ __forceinline__ void set_weights(SerWeights &wt, double c, double cc, double d, double dd)
{
  ser_set_weights(&wt, c, cc, d, dd);
}

/* totals and log-likelihood from the block-reduced alive-ones / lifespan sums (mcmc.c:977-986) */
__device__ __forceinline__ void totals_from(const KParams &p, const SerWeights &wt, int T1, int LEN, int *t0a, int *f0a,
                                            int *t1a, int *f1a, double *loglik)
{
  const int f1 = (int)p.ones_total - T1, f0 = LEN - T1, t0 = p.N * p.M - LEN - f1;
  *t1a = T1; *f1a = f1; *f0a = f0; *t0a = t0;
  *loglik = SER_ADD(SER_ADD(SER_ADD(SER_MUL((double)t0, wt.cc), SER_MUL((double)f0, wt.d)), SER_MUL((double)T1, wt.dd)),
                    SER_MUL((double)f1, wt.c));
}

/* ------------------------------------------------------------------ V-free helpers
 * The init / export / check kernels do not need the bit columns: a thread walks its taxa's cells in
 * position order straight from the site-major matrix.  They work for every shape. */
__device__ __forceinline__ int cell(const KParams &p, const uint16_t *rpi, int pos, int c)
{
  return (p.Xs[(size_t)rpi[pos] * p.Mw + (c >> 5)] >> (c & 31)) & 1u;
}
__device__ int taxon_count(const KParams &p, const uint16_t *rpi, int c, int lo, int hi)
{
  int n = 0;
  for (int pos = lo; pos < hi; pos++) n += cell(p, rpi, pos, c);
  return n;
}
/* mcmc_initab, mcmc.c:440-474 */
__device__ void taxon_init_ab(const KParams &p, const uint16_t *rpi, int c, int *a, int *b)
{
  int first = -1, last = -1;
  for (int pos = 0; pos < p.N; pos++)
    if (cell(p, rpi, pos, c)) { if (first < 0) first = pos; last = pos; }
  if (first < 0) { *a = 0; *b = p.N; } else { *a = first; *b = last + 1; }
}

struct AuxSmem { /* init / export / check kernels */
  int *red;
  uint16_t *rpi, *tmp16;
};
__host__ __device__ inline size_t aux_layout(AuxSmem *s, unsigned char *base, int N)
{
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  size_t o_r = take(sizeof(int) * 2 * SER_MAX_WARPS * 4), o_p = take(sizeof(uint16_t) * N), o_t = take(sizeof(uint16_t) * N);
  if (s) { s->red = (int *)(base + o_r); s->rpi = (uint16_t *)(base + o_p); s->tmp16 = (uint16_t *)(base + o_t); }
  return off;
}

/* ------------------------------------------------------------------ init kernel */
/* mcmc_readmodel's initial state + mcmc_randomize (mcmc.c:405-433, :477-578) */
__global__ void ser_init_kernel(KParams p)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AuxSmem sm;
  const size_t used = aux_layout(&sm, smem_raw, p.N);
  /* behind the common layout: 2N staged draws, pi / rest / chosen as u16 */
  double *stage = (double *)(smem_raw + used);
  uint16_t *pi16 = (uint16_t *)(stage + 2 * p.N);
  uint16_t *rest16 = pi16 + p.N, *chosen16 = rest16 + p.N;

  const int chain = blockIdx.x, tid = threadIdx.x, N = p.N, M = p.M, C = blockDim.x, nh = p.nh;
  const unsigned int gchain = (unsigned int)(p.chain_offset + chain);
  const double *tape = nullptr;
  long long tape_len = 0;
  if (p.mode == SER_MODE_REPLAY) {
    tape = p.tape + p.tape_off[chain];
    tape_len = (long long)(p.tape_off[chain + 1] - p.tape_off[chain]);
  }
  for (int t = tid; t < 2 * N; t += C) {
    if (p.mode == SER_MODE_REPLAY) stage[t] = (t < tape_len) ? tape[t] : 0.0;
    else stage[t] = ser_stream_uniform(p.seed, gchain, SER_SWEEP_INIT, SER_BLK_INIT, (uint32_t)t);
  }
  for (int n = tid; n < N; n += C) sm.rpi[n] = (uint16_t)n;
  __syncthreads();

  uint16_t *ab = p.ab + (size_t)chain * 2 * p.Mpad;
  if (nh == 0) { /* identity-order a/b are kept although pi is shuffled (mcmc.c:486-494) */
    for (int c = tid; c < M; c += C) {
      int a, b;
      taxon_init_ab(p, sm.rpi, c, &a, &b);
      ab[c] = (uint16_t)a; ab[p.Mpad + c] = (uint16_t)b;
    }
    __syncthreads();
  }

  __shared__ int s_used;
  if (tid == 0) {
    int used_draws = 0;
    for (int n = 0; n < N; n++) pi16[n] = (uint16_t)n;
    if (nh == 0) {
      for (int i = N - 1; i > 0; i--) {
        const int j = ser_draw_int(stage[used_draws++], i + 1);
        const uint16_t t = pi16[i]; pi16[i] = pi16[j]; pi16[j] = t;
      }
    } else if (nh < N) {
      int j = 0;
      for (int i = 0; i < N && j < nh; i++)
        if (SER_MUL((double)(N - i), stage[used_draws++]) < (double)(nh - j)) chosen16[j++] = (uint16_t)i;
      int k = 0;
      j = 0;
      for (int i = 0; i < N; i++) {
        if (j < nh && i == chosen16[j]) j++;
        else rest16[k++] = (uint16_t)i;
      }
      for (int i = N - nh - 1; i > 0; i--) {
        const int r = ser_draw_int(stage[used_draws++], i + 1);
        const uint16_t t = rest16[i]; rest16[i] = rest16[r]; rest16[r] = t;
      }
      j = k = 0;
      for (int i = 0; i < N; i++) pi16[i] = p.hard[i] ? chosen16[j++] : rest16[k++];
    }
    s_used = used_draws;
  }
  __syncthreads();
  for (int n = tid; n < N; n += C) sm.rpi[pi16[n]] = (uint16_t)n;
  __syncthreads();

  SerWeights wt;
  wt.eps = p.eps;
  set_weights(wt, p.c0, p.cc0, p.d0, p.dd0);
  int t1 = 0, len = 0;
  for (int c = tid; c < M; c += C) {
    int a, b;
    if (nh != 0) { taxon_init_ab(p, sm.rpi, c, &a, &b); ab[c] = (uint16_t)a; ab[p.Mpad + c] = (uint16_t)b; }
    else { a = ab[c]; b = ab[p.Mpad + c]; }
    t1 += taxon_count(p, sm.rpi, c, a, b);
    len += b - a;
  }
  int buf = 0, T1, LEN, dummy;
  block_sum3(t1, len, 0, sm.red, buf, &T1, &LEN, &dummy);
  int t0a, f0a, t1a, f1a;
  double loglik;
  totals_from(p, wt, T1, LEN, &t0a, &f0a, &t1a, &f1a, &loglik);

  for (int n = tid; n < N; n += C) p.rpi[(size_t)chain * p.Npad + n] = sm.rpi[n];
  if (p.manycd)
    for (int c = tid; c < M; c += C) {
      double *cd = p.cd4 + (size_t)chain * 4 * p.Mpad + c;
      cd[0] = p.c0; cd[p.Mpad] = p.cc0; cd[2 * p.Mpad] = p.d0; cd[3 * p.Mpad] = p.dd0;
    }
  if (tid == 0) {
    ChainScalars sc;
    memset(&sc, 0, sizeof(sc));
    sc.c = p.c0; sc.cc = p.cc0; sc.d = p.d0; sc.dd = p.dd0;
    sc.loglik = loglik;
    sc.t0a = t0a; sc.f0a = f0a; sc.t1a = t1a; sc.f1a = f1a;
    sc.cursor = s_used;
    sc.flags = (p.mode == SER_MODE_REPLAY && s_used > tape_len) ? 1 : 0;
    p.scal[chain] = sc;
  }
}

/* ------------------------------------------------------------------ the sweep kernel */
struct PropState { /* thread-uniform bookkeeping of the pi part */
  int k;           /* next slot of draws_pi */
  int buf;         /* reduction double-buffer index */
};

/* sum of the M per-taxon terms in taxon order (the reference's own order of additions).  One warp walks
 * the dependent chain and publishes the result; the others wait at the barrier instead of issuing the same
 * M additions (the kernel is issue-bound and shares the SM with other chains).  terms[] is published. */
__device__ __forceinline__ double sequential_term_sum(const Smem &sm, int M)
{
  if (threadIdx.x < 32) {
    double acc = 0.0;
    for (int m = 0; m < M; m++) acc = SER_ADD(acc, sm.terms[m]);
    if (threadIdx.x == 0) sm.draws_cd[7] = acc;
  }
  __syncthreads();
  return sm.draws_cd[7];
}

/* block sum of three ints and one double behind one barrier (per-taxon c, d) */
__device__ __forceinline__ void block_sum3d(int v0, int v1, int v2, double x, int *red, double *redd, int &buf, int *o0, int *o1,
                                            int *o2, double *ox)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v0 = __reduce_add_sync(0xffffffffu, v0);
  v1 = __reduce_add_sync(0xffffffffu, v1);
  v2 = __reduce_add_sync(0xffffffffu, v2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  int *r = red + buf * (SER_MAX_WARPS * 4);
  double *rd = redd + buf * SER_MAX_WARPS;
  if (lane == 0) { r[warp * 4 + 0] = v0; r[warp * 4 + 1] = v1; r[warp * 4 + 2] = v2; rd[warp] = x; }
  __syncthreads();
  int s0 = 0, s1 = 0, s2 = 0;
  double sx = 0.0;
  if (lane < nwarp) { s0 = r[lane * 4 + 0]; s1 = r[lane * 4 + 1]; s2 = r[lane * 4 + 2]; sx = rd[lane]; }
  s0 = __reduce_add_sync(0xffffffffu, s0);
  s1 = __reduce_add_sync(0xffffffffu, s1);
  s2 = __reduce_add_sync(0xffffffffu, s2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);
  buf ^= 1;
  *o0 = s0; *o1 = s1; *o2 = s2; *ox = sx;
}

/* MH tail shared by the three proposals (mcmc.c:1261/:1441/:1636): block-reduce the integer deltas, form
 * delta, accept.  Every thread computes the same decision.
 * Scalar c, d: delta follows from the integer totals; if they cancel while single taxa changed, the
 * reference's sequential float sum (mcmc.c:1214/1435/1630) may leave a residual whose SIGN decides whether
 * a draw is consumed, so that sum is re-created exactly.
 * Per-taxon c, d (MANY): delta is a float sum over taxa -- reduced in parallel (good to ~1e-12), re-done in the
 * reference's order when the sign could be ambiguous.  On a sampled sweep every accepted delta is the
 * reference's own sum, so the saved log-likelihood carries its bits. */
template <bool MANY>
__device__ __forceinline__ bool mh_decide(const KParams &p, const Smem &sm, const SerWeights &wt, PropState &ps, int taxon,
                                          bool is_taxon, int dt0, int dt1, bool exact, int *D0, int *D1, double *delta_out)
{
  int nz;
  double delta;
  bool seq = false;
  auto reference_sum = [&]() { /* per-taxon terms in the reference's operand order, added in taxon order */
    double acc = 0.0;
    __syncthreads(); /* terms[] may still be read from an earlier call */
    if (is_taxon) sm.terms[taxon] = ser_term(wt, dt0, dt1);
    __syncthreads();
    acc = sequential_term_sum(sm, p.M);
    return acc;
  };
  if constexpr (MANY) {
    block_sum3d(dt0, dt1, (dt0 | dt1) != 0, is_taxon ? ser_term(wt, dt0, dt1) : 0.0, sm.red, sm.redd, ps.buf, D0, D1, &nz, &delta);
    if (!nz) delta = 0.0;
    else if (fabs(delta) < 1e-7) { delta = reference_sum(); seq = true; }
  } else {
    block_sum3(dt0, dt1, (dt0 | dt1) != 0, sm.red, ps.buf, D0, D1, &nz);
    if (*D0 == 0 && *D1 == 0) {
      delta = 0.0;
      if (nz) { delta = reference_sum(); seq = true; }
    } else {
      delta = ser_term(wt, *D0, *D1);
    }
  }
  bool accept = delta >= 0.0;
  if (!accept) accept = delta > sm.logdraw[ps.k++];
  if (accept && exact && !seq && nz) delta = reference_sum();
  *delta_out = delta;
  return accept;
}

/* MAXT = largest block the instantiation is launched with: the small-block instantiation may use
 * more registers per thread (shared memory, not registers, limits residency there).
 * MANY = per-taxon c, d (manycd = 1, mcmc.c:777-785, :807-815): same choreography; the Beta draws, weights
 * and likelihood terms are per thread, the geometric run sums are evaluated on the fly (no shared table)
 * and delta / loglik are float sums over taxa (mh_decide). */
template <int MAXT, int MINB, bool MANY>
__global__ void __launch_bounds__(MAXT, MINB) ser_sweep_kernel(KParams p)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem sm;
  smem_layout(&sm, smem_raw, p.N, p.W, p.C, p.I, MANY ? 1 : 0, p.Ival);

  const int chain = blockIdx.x, tid = threadIdx.x, N = p.N, M = p.M, C = p.C, W = p.W;
  const unsigned int gchain = (unsigned int)(p.chain_offset + chain);
  uint32_t *col = sm.V + tid;
  uint16_t *pre = sm.pre + tid;
  const bool is_taxon = tid < M, is_col = tid <= M;

  /* ---- load chain state */
  ChainScalars sc = p.scal[chain];
  for (int n = tid; n < N; n += C) sm.rpi[n] = p.rpi[(size_t)chain * p.Npad + n];
  int a = 0, b = 0, taxon = 0, off_c = 0, ones_c = 0;
  double c = p.c0, cc = p.cc0, d = p.d0, dd = p.dd0; /* MANY: this taxon's c, log(1-e^c), d, log(1-e^d) */
  if (is_taxon) {
    a = p.ab[(size_t)chain * 2 * p.Mpad + tid];
    b = p.ab[(size_t)chain * 2 * p.Mpad + p.Mpad + tid];
    taxon = p.order[tid]; /* the taxon this column holds: indexes the tape, the samples, terms[] */
    off_c = p.off[tid];
    ones_c = p.ones[tid];
    sm.ones16[tid] = (uint16_t)ones_c;
    if constexpr (MANY) {
      const double *cd = p.cd4 + (size_t)chain * 4 * p.Mpad + tid;
      c = cd[0]; cc = cd[p.Mpad]; d = cd[2 * p.Mpad]; dd = cd[3 * p.Mpad];
    }
  }
  __syncthreads();
  build_columns(p, sm);
  __syncthreads();
  if (tid == M) rebuild_hard(p, sm);
  __syncthreads();

  const double *tape = nullptr;
  long long tape_len = 0;
  if (p.mode == SER_MODE_REPLAY) {
    tape = p.tape + p.tape_off[chain];
    tape_len = (long long)(p.tape_off[chain + 1] - p.tape_off[chain]);
  }

  SerWeights wt;
  if constexpr (MANY) {
    ser_set_weights_own(&wt, c, cc, d, dd, N);
  } else {
    wt.H = sm.H;
    wt.hmax = 0;
    set_weights(wt, sc.c, sc.cc, sc.d, sc.dd);
  }
  wt.eps = p.eps;
  SerHard hd;
  hd.hcol = sm.V + M; hd.hpre = sm.pre + M; hd.hp = sm.hp; hd.C = C; hd.W = W; hd.N = N; hd.nh = p.nh;
  PropState ps;
  ps.k = 0; ps.buf = 0;

  for (int call = 0; call < p.n_calls && !(sc.flags & 1); call++) {
    for (int s = 0; s < p.sweeps_per_call; s++) {
      __syncthreads(); /* every thread is done reading the previous sweep's staged draws */
      double ua = 0.0, ub = 0.0;
      if constexpr (MANY) {
        /* ================= draws: M Betas for c, M for d, 2M uniforms, then the pi draws ================= */
        double yc = 0.0, lyc = 0.0, l1c = 0.0, yd = 0.0, lyd = 0.0, l1d = 0.0;
        if (p.mode == SER_MODE_REPLAY) {
          const long long need = sc.cursor + 8 * (long long)M;
          if (need > tape_len) { sc.flags |= 1; break; }
          for (int t = tid; t < SER_PI_DRAWS; t += C) {
            const long long idx = need + t;
            const double u = idx < tape_len ? tape[idx] : 0.5;
            sm.draws_pi[t] = u;
            sm.logdraw[t] = log(u);
          }
          if (is_taxon) {
            const double *tc = tape + sc.cursor + 3 * taxon, *td = tape + sc.cursor + 3 * (long long)M + 3 * taxon;
            yc = tc[0]; lyc = tc[1]; l1c = tc[2];
            yd = td[0]; lyd = td[1]; l1d = td[2];
            ua = tape[sc.cursor + 6 * (long long)M + 2 * taxon]; ub = tape[sc.cursor + 6 * (long long)M + 2 * taxon + 1];
          }
        } else {
          for (int t = tid; t < SER_PI_DRAWS; t += C) {
            const double u = ser_stream_uniform(p.seed, gchain, sc.sweep, SER_BLK_PI, (uint32_t)t);
            sm.draws_pi[t] = u;
            sm.logdraw[t] = log(ser_pos(u));
          }
          if (is_taxon) { /* Beta(1+f1_m, 1+t0_m) and Beta(1+f0_m, 1+t1_m) from the taxon's own counts */
            const int t1 = ser_col_popc(col, pre, C, a, b), len = b - a;
            const int f1 = ones_c - t1, f0 = len - t1, t0 = N - len - f1;
            const uint32_t blk = SER_BLK_MANYCD + 4u * (uint32_t)taxon;
            yc = ser_beta_from_gammas(ser_gamma_ge1(1.0 + (double)f1, p.seed, gchain, sc.sweep, blk),
                                      ser_gamma_ge1(1.0 + (double)t0, p.seed, gchain, sc.sweep, blk + 1u));
            yd = ser_beta_from_gammas(ser_gamma_ge1(1.0 + (double)f0, p.seed, gchain, sc.sweep, blk + 2u),
                                      ser_gamma_ge1(1.0 + (double)t1, p.seed, gchain, sc.sweep, blk + 3u));
            if (yc > 0.0) { lyc = ser_log(yc); l1c = ser_log(SER_SUB(1.0, ser_exp(lyc))); }
            if (yd > 0.0) { lyd = ser_log(yd); l1d = ser_log(SER_SUB(1.0, ser_exp(lyd))); }
            uint32_t o[4];
            ser_philox4x32_10((uint32_t)taxon, SER_BLK_AB, sc.sweep, 0u, p.seed, gchain, o);
            ua = ser_u53(o[0], o[1]); ub = ser_u53(o[2], o[3]);
          }
        }
        /* ================= c_m, d_m (mcmc_samplebeta per taxon) ================= */
        if (is_taxon) {
          if (yc > 0.0 && SER_MINC <= lyc && lyc <= SER_MAXC) { c = lyc; cc = l1c; }
          if (yd > 0.0 && SER_MIND <= lyd && lyd <= SER_MAXD) { d = lyd; dd = l1d; }
          ser_set_weights_own(&wt, c, cc, d, dd, N);
          sm.wcol[4 * tid + 0] = wt.A; sm.wcol[4 * tid + 1] = wt.g; sm.wcol[4 * tid + 2] = wt.inv_g; sm.wcol[4 * tid + 3] = wt.hs;
          if (taxon == 0) { sm.draws_cd[0] = c; sm.draws_cd[1] = d; }
        }
        sc.counters[0] += M; sc.counters[1] += M;

      } else {
        /* ================= stage this sweep's draws ================= */
        if (p.mode == SER_MODE_REPLAY) {
          const long long need = sc.cursor + 6 + 2 * (long long)M;
          if (need > tape_len) { sc.flags |= 1; break; } /* uniform across the CTA */
          if (tid < 6) sm.draws_cd[tid] = tape[sc.cursor + tid];
          for (int t = tid; t < SER_PI_DRAWS; t += C) {
            const long long idx = need + t;
            const double u = idx < tape_len ? tape[idx] : 0.5;
            sm.draws_pi[t] = u;
            sm.logdraw[t] = log(u);
          }
          if (is_taxon) { ua = tape[sc.cursor + 6 + 2 * taxon]; ub = tape[sc.cursor + 7 + 2 * taxon]; }
        } else {
          if (tid < 4) { /* Beta(1+f1a,1+t0a) and Beta(1+f0a,1+t1a) as Gamma ratios (mcmc.c:790, :820) */
            const int cnt = tid == 0 ? sc.f1a : tid == 1 ? sc.t0a : tid == 2 ? sc.f0a : sc.t1a;
            const double g = ser_gamma_ge1(1.0 + (double)cnt, p.seed, gchain, sc.sweep, (uint32_t)tid);
            const double go = __shfl_xor_sync(0xfu, g, 1);
            if (tid == 0 || tid == 2) {
              const double y = ser_beta_from_gammas(g, go);
              double val = tid == 0 ? sc.c : sc.d, l1m = tid == 0 ? sc.cc : sc.dd;
              const double lo = tid == 0 ? SER_MINC : SER_MIND, hi = tid == 0 ? SER_MAXC : SER_MAXD;
              if (y > 0.0) { /* mcmc_samplebeta, mcmc.c:751-765 */
                const double ly = ser_log(y);
                if (lo <= ly && ly <= hi) { val = ly; l1m = ser_log(SER_SUB(1.0, ser_exp(ly))); }
              }
              sm.draws_cd[tid] = val; sm.draws_cd[tid + 1] = l1m;
            }
          }
          for (int t = tid; t < SER_PI_DRAWS; t += C) {
            const double u = ser_stream_uniform(p.seed, gchain, sc.sweep, SER_BLK_PI, (uint32_t)t);
            sm.draws_pi[t] = u;
            sm.logdraw[t] = log(ser_pos(u));
          }
          if (is_taxon) {
            uint32_t o[4];
            ser_philox4x32_10((uint32_t)taxon, SER_BLK_AB, sc.sweep, 0u, p.seed, gchain, o);
            ua = ser_u53(o[0], o[1]); ub = ser_u53(o[2], o[3]);
          }
        }
        __syncthreads();

        /* ================= c and d (mcmc_samplec / mcmc_sampled) ================= */
        if (p.mode == SER_MODE_REPLAY) {
          const double yc = sm.draws_cd[0], lyc = sm.draws_cd[1], l1c = sm.draws_cd[2];
          const double yd = sm.draws_cd[3], lyd = sm.draws_cd[4], l1d = sm.draws_cd[5];
          if (yc > 0.0 && SER_MINC <= lyc && lyc <= SER_MAXC) { sc.c = lyc; sc.cc = l1c; }
          if (yd > 0.0 && SER_MIND <= lyd && lyd <= SER_MAXD) { sc.d = lyd; sc.dd = l1d; }
        } else {
          sc.c = sm.draws_cd[0]; sc.cc = sm.draws_cd[1]; sc.d = sm.draws_cd[2]; sc.dd = sm.draws_cd[3];
        }
        set_weights(wt, sc.c, sc.cc, sc.d, sc.dd);
        sc.counters[0]++; sc.counters[1]++;
        /* geometric partial sums for this sweep's g (shared by all taxa: c, d are scalar) */
        wt.hmax = ser_hmax(wt.g, N);
        for (int m = tid; m <= wt.hmax; m += C) sm.H[m] = ser_h_entry(wt.g, m);
        __syncthreads();

      }

      /* ================= a/b Gibbs (mcmc_sampleab, mcmc.c:918-996) =================
       * item formulation (ser_chain_core.h): postings of the column, then for the a-step and
       * the b-step: per-column maximum (own thread), item weights (dense over the CTA),
       * per-column scan + inverse CDF (own thread). */
      if (is_taxon) ser_expand_ones(col, C, W, sm.pos + off_c);
      int changed = 0;
#pragma unroll 1
      for (int step = 0; step < 2; step++) {
        SerStep st;
        double lmax = 0.0;
        if (is_taxon) {
          st = step == 0 ? ser_step_a(col, pre, C, W, N, a, b) : ser_step_b(col, pre, C, W, N, a, b);
          lmax = ser_step_lmax(wt, st, sm.pos + off_c);
          sm.lmax[tid] = lmax;
          sm.st4[4 * tid + 0] = (uint16_t)st.cur; sm.st4[4 * tid + 1] = (uint16_t)st.bound;
          sm.st4[4 * tid + 2] = (uint16_t)st.ocur; sm.st4[4 * tid + 3] = (uint16_t)st.kb;
        }
        __syncthreads();
#pragma unroll 1
        for (int g = 0; g < p.n_groups; g++) { /* columns grp_c[g]..grp_c[g+1] = items grp_e[g]..grp_e[g+1] */
          const int e0 = p.grp_e[g], e1 = p.grp_e[g + 1];
          if (g) __syncthreads(); /* the previous group's scans are done with val */
          uint32_t ck_next = e0 + tid < e1 ? p.item_col[e0 + tid] : 0u; /* item -> column map, fetched one iteration ahead */
          for (int e = e0 + tid; e < e1; e += C) {
            const uint32_t ck = ck_next;
            if (e + C < e1) ck_next = p.item_col[e + C];
            const int c = (int)(ck >> 16), kk = (int)(ck & 0xffffu);
            const uint2 g4 = *reinterpret_cast<const uint2 *>(sm.st4 + 4 * c); /* cur, bound | ocur, kb */
            SerStep it;
            it.cur = (int)(g4.x & 0xffffu); it.bound = (int)(g4.x >> 16); it.ocur = (int)(g4.y & 0xffffu); it.kb = (int)(g4.y >> 16);
            if (kk <= it.kb) {
              it.nones = sm.ones16[c]; it.N = N; it.rev = step;
              if constexpr (MANY) { /* the column's own weights; geometric sums on the fly */
                SerWeights w;
                w.A = sm.wcol[4 * c + 0]; w.g = sm.wcol[4 * c + 1]; w.inv_g = sm.wcol[4 * c + 2]; w.hs = sm.wcol[4 * c + 3];
                w.eps = p.eps; w.H = nullptr; w.hmax = N + 1;
                sm.val[e - e0] = ser_item_weight<0>(w, it, sm.pos + (e - kk), kk, sm.lmax[c]);
              } else {
                sm.val[e - e0] = ser_item_weight<1>(wt, it, sm.pos + (e - kk), kk, sm.lmax[c]);
              }
            }
          }
          __syncthreads();
          if (is_taxon && tid >= p.grp_c[g] && tid < p.grp_c[g + 1]) {
            const int pick = ser_step_pick<MANY ? 0 : 1>(wt, st, sm.pos + off_c, sm.val + (off_c - e0), lmax, step == 0 ? ua : ub);
            if (step == 0) { changed += pick != a; a = pick; }
            else { changed += (N - pick) != b; b = N - pick; }
          }
        }
      }
      /* the log-likelihood is only ever observed after the last sweep of a sampling call
       * (mcmc_save_chain); there it is formed with the reference's own sequential sums */
      const bool exact = p.sampling && s == p.sweeps_per_call - 1;
      if constexpr (MANY) {
        {
          int t1 = 0, len = 0, T1, LEN, CH;
          double term = 0.0, ll;
          if (is_taxon) {
            t1 = ser_col_popc(col, pre, C, a, b); len = b - a;
            const int f1 = ones_c - t1, f0 = len - t1, t0 = N - len - f1;
            term = SER_ADD(SER_ADD(SER_ADD(SER_MUL((double)t0, wt.cc), SER_MUL((double)f0, wt.d)), SER_MUL((double)t1, wt.dd)),
                           SER_MUL((double)f1, wt.c));
          }
          block_sum3d(t1, len, changed, term, sm.red, sm.redd, ps.buf, &T1, &LEN, &CH, &ll);
          sc.t1a = T1; sc.f1a = (int)p.ones_total - T1; sc.f0a = LEN - T1; sc.t0a = N * M - LEN - sc.f1a;
          sc.loglik = ll;
          sc.counters[2] += CH;
          if (exact) { /* mcmc_logl's own order */
            if (is_taxon) sm.terms[taxon] = term;
            __syncthreads();
            sc.loglik = sequential_term_sum(sm, M);
          }
        }
      } else {
        int t1 = 0, len = 0;
        if (is_taxon) { t1 = ser_col_popc(col, pre, C, a, b); len = b - a; }
        {
          int T1, LEN, CH;
          block_sum3(t1, len, changed, sm.red, ps.buf, &T1, &LEN, &CH);
          totals_from(p, wt, T1, LEN, &sc.t0a, &sc.f0a, &sc.t1a, &sc.f1a, &sc.loglik);
          sc.counters[2] += CH;
          if (exact) { /* mcmc_logl, mcmc.c:625-648: sum over taxa of t0*cc + f0*d + t1*dd + f1*c */
            if (is_taxon) {
              const int f1 = ones_c - t1, f0 = len - t1, t0 = N - len - f1;
              sm.terms[taxon] = SER_ADD(SER_ADD(SER_ADD(SER_MUL((double)t0, wt.cc), SER_MUL((double)f0, wt.d)), SER_MUL((double)t1, wt.dd)),
                                        SER_MUL((double)f1, wt.c));
            }
            __syncthreads();
            sc.loglik = sequential_term_sum(sm, M);
          }
        }
      }

      /* ================= 16 proposals for pi (mcmc.c:237-243) ================= */
      ps.k = 0;
      for (int prop = 0; prop < 16; prop++) {
        /* order: pi2(swap), then 5 x (pi1, pi2(0), pi3) */
        const int kind = prop == 0 ? 3 : ((prop - 1) % 3); /* 0 pi1, 1 pi2(0), 2 pi3, 3 pi2(swap) */
        int dt0 = 0, dt1 = 0, D0, D1;
        double delta;
        if (kind == 0) { /* ---------------- mcmc_samplepi1, mcmc.c:1127-1308 */
          const int i = ser_draw_int(sm.draws_pi[ps.k], N);
          int j = ser_draw_int(sm.draws_pi[ps.k + 1], N - 1);
          ps.k += 2;
          if (j >= i) j++;
          const int lo = min(i, j), hi = max(i, j);
          const int nhw = ser_hard_count(hd, lo, hi); /* hard sites in the window */
          if (ser_is_hard(hd, i) && nhw > 1) continue;
          if (is_taxon) ser_pi1_delta(col, C, a, b, i, j, &dt0, &dt1);
          if (!mh_decide<MANY>(p, sm, wt, ps, taxon, is_taxon, dt0, dt1, exact, &D0, &D1, &delta)) continue;
          if (is_taxon) ser_pi1_apply_ab(&a, &b, i, j);
          if (is_col) ser_col_rotate(col, C, W, i, j, pre);
          for (int n = lo + tid; n <= hi; n += C)
            sm.tmp16[n] = sm.rpi[i < j ? (n < j ? n + 1 : i) : (n > j ? n - 1 : i)];
          __syncthreads();
          for (int n = lo + tid; n <= hi; n += C) sm.rpi[n] = sm.tmp16[n];
          if (tid == M && nhw) rebuild_hard(p, sm); /* the hard column only changed if the window holds a hard site */
          sc.counters[3]++;
        } else if (kind == 1 || kind == 3) { /* ---------------- mcmc_samplepi2, mcmc.c:1311-1486 */
          int i, j;
          if (kind == 1) {
            i = ser_draw_int(sm.draws_pi[ps.k], N);
            j = ser_draw_int(sm.draws_pi[ps.k + 1], N - 1);
            ps.k += 2;
            if (j >= i) j++;
            else { const int t = i; i = j; j = t; }
          } else {
            i = ser_draw_int(sm.draws_pi[ps.k], N - 1);
            ps.k += 1;
            j = i + 1;
          }
          const int nhw = ser_hard_count(hd, i, j);
          if (nhw > 1) continue;
          const int inc1 = ser_draw_int(sm.draws_pi[ps.k], 2), inc2 = ser_draw_int(sm.draws_pi[ps.k + 1], 2);
          ps.k += 2;
          if (is_taxon) ser_pi2_delta(col, pre, C, a, b, i, j, inc1, inc2, &dt0, &dt1);
          if (!mh_decide<MANY>(p, sm, wt, ps, taxon, is_taxon, dt0, dt1, exact, &D0, &D1, &delta)) continue;
          if (is_taxon) {
            const int ain = ser_in_window(a, i, j + 1, inc1, inc2), bin = ser_in_window(b, i, j + 1, inc1, inc2);
            ser_mirror_ab(a, b, ain, bin, i + j + 1, &a, &b);
          }
          if (is_col) ser_col_reverse(col, C, W, i, j, pre);
          for (int n = i + tid; 2 * n < i + j; n += C) { /* mirror the site order: disjoint pairs, no staging */
            const uint16_t t = sm.rpi[n];
            sm.rpi[n] = sm.rpi[i + j - n]; sm.rpi[i + j - n] = t;
          }
          if (tid == M && nhw) rebuild_hard(p, sm);
          sc.counters[kind == 1 ? 4 : 5]++;
        } else { /* ---------------- mcmc_samplepi3, mcmc.c:1489-1682 */
          const int nfree = N - p.nh;
          if (nfree < 2) continue;
          const int r1 = ser_draw_int(sm.draws_pi[ps.k], nfree), r2 = ser_draw_int(sm.draws_pi[ps.k + 1], nfree - 1);
          ps.k += 2;
          int ir, jr;
          if (r1 <= r2) { ir = r1; jr = r2 + 1; } else { ir = r2; jr = r1; }
          const SerPi3 g = ser_pi3_window(hd, ir, jr);
          const int inc1 = ser_draw_int(sm.draws_pi[ps.k], 2), inc2 = ser_draw_int(sm.draws_pi[ps.k + 1], 2);
          ps.k += 2;
          if (is_taxon) ser_pi3_delta(col, pre, C, hd, g, a, b, inc1, inc2, &dt0, &dt1);
          if (!mh_decide<MANY>(p, sm, wt, ps, taxon, is_taxon, dt0, dt1, exact, &D0, &D1, &delta)) continue;
          for (int n = g.i + tid; n <= g.j; n += C) sm.perm16[n] = (uint16_t)ser_pi3_perm(hd, g, n);
          __syncthreads();
          if (is_taxon) {
            const int ain = ser_in_window(a, g.i, g.j + 1, inc1, inc2), bin = ser_in_window(b, g.i, g.j + 1, inc1, inc2);
            ser_mirror_ab(a, b, ain, bin, g.i + g.j + 1, &a, &b);
            ser_col_permute(col, C, W, g.i, g.j, sm.perm16, pre);
          }
          for (int n = g.i + tid; n <= g.j; n += C) { /* the permutation is an involution: disjoint pairs */
            const int m2 = sm.perm16[n];
            if (m2 > n) { const uint16_t t = sm.rpi[n]; sm.rpi[n] = sm.rpi[m2]; sm.rpi[m2] = t; }
          }
          sc.counters[6]++;
        }
        /* accepted: fold the integer deltas into the totals (the reference recounts, mcmc.c:1303) */
        sc.t0a += D0; sc.f0a -= D0; sc.t1a += D1; sc.f1a -= D1;
        sc.loglik = SER_ADD(sc.loglik, delta);
        __syncthreads(); /* columns / hard mask / rpi visible before the next proposal */
      }

      if (p.mode == SER_MODE_REPLAY) sc.cursor += (MANY ? 8 * (long long)M : 6 + 2 * (long long)M) + ps.k;
      else sc.sweep++;
      sc.counters[7]++;
    }
    if (sc.flags & 1) break;

    /* ================= thinned sample (mcmc_save_chain + compute_exp_data) ================= */
    if constexpr (MANY) {
      if (p.sampling) {
        const int sidx = sc.n_samples;
        const double c_first = sm.draws_cd[0], d_first = sm.draws_cd[1]; /* taxon 0's c, d (compute_exp_data, mcmc.c:56-57) */
        if (sidx < p.max_samples) {
          const size_t row = (size_t)chain * p.max_samples + sidx;
          if (p.store >= SER_STORE_PI)
            for (int pos = tid; pos < N; pos += C) p.samp_pi[row * N + sm.rpi[pos]] = (uint16_t)pos;
          if (p.store >= SER_STORE_FULL) {
            if (is_taxon) {
              p.samp_a[row * M + taxon] = (uint16_t)a; p.samp_b[row * M + taxon] = (uint16_t)b;
              p.samp_cd_all[(row * 2 + 0) * M + taxon] = c; p.samp_cd_all[(row * 2 + 1) * M + taxon] = d;
            }
            if (tid == 0) { p.samp_cdl[row * 3 + 0] = c_first; p.samp_cdl[row * 3 + 1] = d_first; p.samp_cdl[row * 3 + 2] = sc.loglik; }
          }
        }
        sc.c = c_first; sc.d = d_first;
        sc.sum_negll = SER_ADD(sc.sum_negll, -sc.loglik);
        sc.sum_ec = SER_ADD(sc.sum_ec, exp(c_first));
        sc.sum_ed = SER_ADD(sc.sum_ed, exp(d_first));
        sc.n_samples++;
      }
    } else {
      if (p.sampling) {
        const int sidx = sc.n_samples;
        if (sidx < p.max_samples) {
          const size_t row = (size_t)chain * p.max_samples + sidx;
          if (p.store >= SER_STORE_PI)
            for (int pos = tid; pos < N; pos += C) p.samp_pi[row * N + sm.rpi[pos]] = (uint16_t)pos;
          if (p.store >= SER_STORE_FULL) {
            if (is_taxon) { p.samp_a[row * M + taxon] = (uint16_t)a; p.samp_b[row * M + taxon] = (uint16_t)b; }
            if (tid == 0) { p.samp_cdl[row * 3 + 0] = sc.c; p.samp_cdl[row * 3 + 1] = sc.d; p.samp_cdl[row * 3 + 2] = sc.loglik; }
          }
        }
        sc.sum_negll = SER_ADD(sc.sum_negll, -sc.loglik);
        sc.sum_ec = SER_ADD(sc.sum_ec, exp(sc.c));
        sc.sum_ed = SER_ADD(sc.sum_ed, exp(sc.d));
        sc.n_samples++;
      }
    }
  }

  /* ---- save chain state */
  if constexpr (MANY) {
    __syncthreads();
    if (is_taxon) {
      p.ab[(size_t)chain * 2 * p.Mpad + tid] = (uint16_t)a;
      p.ab[(size_t)chain * 2 * p.Mpad + p.Mpad + tid] = (uint16_t)b;
      double *cd = p.cd4 + (size_t)chain * 4 * p.Mpad + tid;
      cd[0] = c; cd[p.Mpad] = cc; cd[2 * p.Mpad] = d; cd[3 * p.Mpad] = dd;
      if (taxon == 0) { sm.draws_cd[0] = c; sm.draws_cd[1] = cc; sm.draws_cd[2] = d; sm.draws_cd[3] = dd; }
    }
    for (int n = tid; n < N; n += C) p.rpi[(size_t)chain * p.Npad + n] = sm.rpi[n];
    __syncthreads();
    if (tid == 0) { /* the scalar slots carry taxon 0's c, d */
      sc.c = sm.draws_cd[0]; sc.cc = sm.draws_cd[1]; sc.d = sm.draws_cd[2]; sc.dd = sm.draws_cd[3];
      p.scal[chain] = sc;
    }
  } else {
    __syncthreads();
    if (is_taxon) {
      p.ab[(size_t)chain * 2 * p.Mpad + tid] = (uint16_t)a;
      p.ab[(size_t)chain * 2 * p.Mpad + p.Mpad + tid] = (uint16_t)b;
    }
    for (int n = tid; n < N; n += C) p.rpi[(size_t)chain * p.Npad + n] = sm.rpi[n];
    if (tid == 0) p.scal[chain] = sc;
  }
}

/* phase timing of the large-shape kernel (debug builds: NVCC_EXTRA=-DSER_PHASE_TIMING): thread 0 of every
 * CTA adds the cycles between marks; ser_debug_phase_cycles() reads and clears the totals */
#ifdef SER_PHASE_TIMING
__device__ unsigned long long ser_phase_cycles[8];
#define PHASE_T0() long long ph_t = clock64()
#define PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long ph_n = clock64(); atomicAdd(&ser_phase_cycles[i], (unsigned long long)(ph_n - ph_t)); ph_t = ph_n; } } while (0)
extern "C" int ser_debug_phase_cycles(unsigned long long out[8])
{
  unsigned long long zero[8] = {0};
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(out, ser_phase_cycles, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(ser_phase_cycles, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#else
#define PHASE_T0() do { } while (0)
#define PHASE_MARK(i) do { } while (0)
#endif

/* ------------------------------------------------------------------ the sweep kernel, large shapes
 * Same algorithm and building blocks as ser_sweep_kernel, for matrices whose bit columns, prefix
 * tables and item buffers exceed shared memory (e.g. 1024 sites x 4096 taxa: 0.7 MB + 0.3 MB +
 * 3.2 MB per chain).  A CTA owns a slot of L2-resident global scratch and walks over chains
 * (persistent grid); every thread owns the columns tid, tid+C, ...; a/b live in shared memory. */
struct BigSmem {
  double *draws_pi, *logdraw, *draws_cd, *H;
  double *lmax; /* gcap: per column of the running group */
  double *val;  /* icap: item weights / cumulative weights of the running group */
  double *incl; /* gcap: inclusive chunk totals, one per (column, lane) unit of the running group */
  int *red;
  uint16_t *a16, *b16, *hp, *rpi, *tmp16, *perm16;
  uint16_t *st4; /* 4 * gcap */
  uint16_t *gones; /* gcap: ones of the running group's columns */
  int *goff;       /* gcap: first item of each column relative to the group's first item */
  uint16_t *pos; /* icap: postings of the running group's columns */
};
__host__ __device__ inline size_t big_layout(BigSmem *s, unsigned char *base, int N, int M, int icap, int gcap)
{
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  size_t o_dp = take(8 * SER_PI_DRAWS), o_ld = take(8 * SER_PI_DRAWS), o_dc = take(8 * 8), o_H = take(8 * (size_t)(N + 2));
  size_t o_go = take(4 * (size_t)gcap), o_gn = take(2 * (size_t)gcap), o_lm = take(8 * (size_t)gcap), o_in = take(8 * (size_t)gcap), o_val = take(8 * (size_t)icap), o_r = take(sizeof(int) * 2 * SER_MAX_WARPS * 4);
  size_t o_a = take(2 * (size_t)M), o_b = take(2 * (size_t)M), o_st = take(2 * 4 * (size_t)gcap), o_hp = take(2 * (size_t)(N + 1));
  size_t o_p = take(2 * (size_t)N), o_q = take(2 * (size_t)N), o_m = take(2 * (size_t)N), o_pos = take(2 * (size_t)icap);
  if (s) {
    s->val = (double *)(base + o_val); s->pos = (uint16_t *)(base + o_pos); s->incl = (double *)(base + o_in);
    s->goff = (int *)(base + o_go); s->gones = (uint16_t *)(base + o_gn);
    s->draws_pi = (double *)(base + o_dp); s->logdraw = (double *)(base + o_ld); s->draws_cd = (double *)(base + o_dc);
    s->H = (double *)(base + o_H); s->lmax = (double *)(base + o_lm); s->red = (int *)(base + o_r);
    s->a16 = (uint16_t *)(base + o_a); s->b16 = (uint16_t *)(base + o_b); s->st4 = (uint16_t *)(base + o_st);
    s->hp = (uint16_t *)(base + o_hp); s->rpi = (uint16_t *)(base + o_p); s->tmp16 = (uint16_t *)(base + o_q);
    s->perm16 = (uint16_t *)(base + o_m);
  }
  return off;
}

/* MH tail for the large-shape kernel: the thread's deltas are already summed over its columns;
 * the degenerate case re-evaluates the per-taxon deltas through `redo` (a lambda) */
template <typename Redo>
__device__ __forceinline__ bool mh_decide_big(const KParams &p, const BigSmem &sm, const SerWeights &wt, PropState &ps,
                                              double *terms, int dt0, int dt1, int nz, bool exact, int *D0, int *D1,
                                              double *delta_out, Redo redo)
{
  int NZ;
  block_sum3(dt0, dt1, nz, sm.red, ps.buf, D0, D1, &NZ);
  auto reference_sum = [&]() { /* see mh_decide */
    double acc = 0.0;
    __syncthreads();
    for (int c = threadIdx.x; c < p.M; c += blockDim.x) {
      int x0, x1;
      redo(c, &x0, &x1);
      terms[p.order[c]] = ser_term(wt, x0, x1);
    }
    __syncthreads();
    if (threadIdx.x < 32) { /* one warp walks the dependent chain, the others wait (see sequential_term_sum) */
      for (int m = 0; m < p.M; m++) acc = SER_ADD(acc, terms[m]);
      if (threadIdx.x == 0) sm.draws_cd[7] = acc;
    }
    __syncthreads();
    return sm.draws_cd[7];
  };
  double delta;
  bool seq = false;
  if (*D0 == 0 && *D1 == 0) {
    delta = 0.0;
    if (NZ) { delta = reference_sum(); seq = true; }
  } else {
    delta = ser_term(wt, *D0, *D1);
  }
  bool accept = delta >= 0.0;
  if (!accept) accept = delta > sm.logdraw[ps.k++];
  if (accept && exact && !seq && NZ) delta = reference_sum();
  *delta_out = delta;
  return accept;
}

__global__ void __launch_bounds__(1024, 1) ser_sweep_kernel_big(KParams p)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BigSmem sm;
  big_layout(&sm, smem_raw, p.N, p.M, p.big_icap, p.big_gcap);
  const int tid = threadIdx.x, N = p.N, M = p.M, C = blockDim.x, W = p.W, Cs = p.Cs;
  uint32_t *V = p.gV + (size_t)blockIdx.x * W * Cs;
  uint16_t *PRE = p.gpre + (size_t)blockIdx.x * (W + 1) * Cs;
  double *TERMS = sm.val; /* per-taxon terms of the exact sums: the item-weight buffer is idle outside the Gibbs phase (icap >= M) */

  for (int chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
    const unsigned int gchain = (unsigned int)(p.chain_offset + chain);
    __syncthreads(); /* previous chain's state fully saved before the scratch is reused */
    ChainScalars sc = p.scal[chain];
    for (int n = tid; n < N; n += C) sm.rpi[n] = p.rpi[(size_t)chain * p.Npad + n];
    for (int c = tid; c < M; c += C) {
      sm.a16[c] = p.ab[(size_t)chain * 2 * p.Mpad + c];
      sm.b16[c] = p.ab[(size_t)chain * 2 * p.Mpad + p.Mpad + c];
    }
    __syncthreads();
    /* position-ordered columns + prefix tables of the owned columns */
    for (int c = tid; c <= M; c += C) {
      for (int w = 0; w < W; w++) {
        uint32_t word = 0;
        const int pend = min(32 * w + 32, N);
        if (c < M) { for (int pos = 32 * w; pos < pend; pos++) word |= (uint32_t)cell(p, sm.rpi, pos, c) << (pos & 31); }
        else { for (int pos = 32 * w; pos < pend; pos++) word |= (uint32_t)(p.hard[sm.rpi[pos]] != 0) << (pos & 31); }
        V[w * Cs + c] = word;
      }
      ser_col_build_pre(V + c, PRE + c, Cs, W);
      if (c == M) ser_hard_list(V + M, Cs, W, sm.hp);
    }
    __syncthreads();

    const double *tape = nullptr;
    long long tape_len = 0;
    if (p.mode == SER_MODE_REPLAY) {
      tape = p.tape + p.tape_off[chain];
      tape_len = (long long)(p.tape_off[chain + 1] - p.tape_off[chain]);
    }
    SerWeights wt;
    wt.eps = p.eps; wt.H = sm.H; wt.hmax = 0;
    set_weights(wt, sc.c, sc.cc, sc.d, sc.dd);
    SerHard hd;
    hd.hcol = V + M; hd.hpre = PRE + M; hd.hp = sm.hp; hd.C = Cs; hd.W = W; hd.N = N; hd.nh = p.nh;
    PropState ps;
    ps.k = 0; ps.buf = 0;

    for (int call = 0; call < p.n_calls && !(sc.flags & 1); call++) {
      for (int s = 0; s < p.sweeps_per_call; s++) {
        /* ================= stage this sweep's draws ================= */
        __syncthreads();
        PHASE_T0();
        if (p.mode == SER_MODE_REPLAY) {
          const long long need = sc.cursor + 6 + 2 * (long long)M;
          if (need > tape_len) { sc.flags |= 1; break; }
          if (tid < 6) sm.draws_cd[tid] = tape[sc.cursor + tid];
          for (int t = tid; t < SER_PI_DRAWS; t += C) {
            const long long idx = need + t;
            const double u = idx < tape_len ? tape[idx] : 0.5;
            sm.draws_pi[t] = u; sm.logdraw[t] = log(u);
          }
        } else {
          if (tid < 4) {
            const int cnt = tid == 0 ? sc.f1a : tid == 1 ? sc.t0a : tid == 2 ? sc.f0a : sc.t1a;
            const double g = ser_gamma_ge1(1.0 + (double)cnt, p.seed, gchain, sc.sweep, (uint32_t)tid);
            const double go = __shfl_xor_sync(0xfu, g, 1);
            if (tid == 0 || tid == 2) {
              const double y = ser_beta_from_gammas(g, go);
              double val = tid == 0 ? sc.c : sc.d, l1m = tid == 0 ? sc.cc : sc.dd;
              const double lo = tid == 0 ? SER_MINC : SER_MIND, hi = tid == 0 ? SER_MAXC : SER_MAXD;
              if (y > 0.0) {
                const double ly = ser_log(y);
                if (lo <= ly && ly <= hi) { val = ly; l1m = ser_log(SER_SUB(1.0, ser_exp(ly))); }
              }
              sm.draws_cd[tid] = val; sm.draws_cd[tid + 1] = l1m;
            }
          }
          for (int t = tid; t < SER_PI_DRAWS; t += C) {
            const double u = ser_stream_uniform(p.seed, gchain, sc.sweep, SER_BLK_PI, (uint32_t)t);
            sm.draws_pi[t] = u; sm.logdraw[t] = log(ser_pos(u));
          }
        }
        __syncthreads();
        if (p.mode == SER_MODE_REPLAY) {
          const double yc = sm.draws_cd[0], lyc = sm.draws_cd[1], l1c = sm.draws_cd[2];
          const double yd = sm.draws_cd[3], lyd = sm.draws_cd[4], l1d = sm.draws_cd[5];
          if (yc > 0.0 && SER_MINC <= lyc && lyc <= SER_MAXC) { sc.c = lyc; sc.cc = l1c; }
          if (yd > 0.0 && SER_MIND <= lyd && lyd <= SER_MAXD) { sc.d = lyd; sc.dd = l1d; }
        } else {
          sc.c = sm.draws_cd[0]; sc.cc = sm.draws_cd[1]; sc.d = sm.draws_cd[2]; sc.dd = sm.draws_cd[3];
        }
        set_weights(wt, sc.c, sc.cc, sc.d, sc.dd);
        sc.counters[0]++; sc.counters[1]++;
        wt.hmax = ser_hmax(wt.g, N);
        for (int m = tid; m <= wt.hmax; m += C) sm.H[m] = ser_h_entry(wt.g, m);

        /* ================= a/b Gibbs, item formulation, one column group at a time =================
         * The group's postings and item weights live in shared memory (icap items), so the per-column
         * loops run at shared-memory latency.  A column is served by `lpc` adjacent lanes (1..8, as many as
         * the block can spare for the group), which split its loops: the maximum is a lane-strided partial
         * maximum + shuffle, the cumulative weights are a per-lane serial sum over a contiguous chunk + a
         * shuffle scan of the lane totals.  Per group: postings; then for the a-step and the b-step:
         * geometry + maximum, run weights (dense over the group's items), scan + inverse CDF. */
        int changed = 0;
#pragma unroll 1
        for (int g = 0; g < p.big_ng; g++) {
          const int c0 = p.bgrp[2 * g], e0 = p.bgrp[2 * g + 1], c1 = p.bgrp[2 * g + 2], e1 = p.bgrp[2 * g + 3], nc = c1 - c0;
          int lpc = 1, lsh = 0;
          while (lpc < 8 && nc * lpc * 2 <= C) { lpc <<= 1; lsh++; }
          const int units = nc << lsh, sub = tid & (lpc - 1);
          __syncthreads(); /* previous group is done with pos / val; first group: publishes H */
          PHASE_MARK(0);
          { /* postings: a unit = (run of wq words, column), column fastest so that a warp reads 32
             * consecutive columns of one word row; the prefix table gives the unit's first slot */
            const int wq = (W + lpc - 1) >> lsh;
            for (int u = tid; u < units; u += C) {
              const int qq = u / nc, cl = u - qq * nc, c = c0 + cl, w0 = qq * wq, w1 = min(W, w0 + wq);
              /* every load of the unit is issued before the first one is needed */
              uint32_t vv[8];
#pragma unroll
              for (int k = 0; k < 8; k++) vv[k] = w0 + k < w1 ? V[(w0 + k) * Cs + c] : 0u;
              const int first = w0 < w1 ? (int)PRE[w0 * Cs + c] : 0, off_c = p.off[c] - e0;
              if (qq == 0) { sm.goff[cl] = off_c; sm.gones[cl] = (uint16_t)p.ones[c]; } /* the group's column table */
              uint16_t *out = sm.pos + off_c + first;
              for (int wb = w0; wb < w1; wb += 8) {
                if (wb > w0) {
#pragma unroll
                  for (int k = 0; k < 8; k++) vv[k] = wb + k < w1 ? V[(wb + k) * Cs + c] : 0u;
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                  uint32_t v = vv[k];
                  while (v) { *out++ = (uint16_t)(32 * (wb + k) + SER_FFS(v) - 1); v &= v - 1u; }
                }
              }
            }
          }
          __syncthreads();
          PHASE_MARK(1);
#pragma unroll 1
          for (int step = 0; step < 2; step++) {
            for (int ub = 0; ub < units; ub += C) { /* warp-uniform trip count: the shuffles need every lane */
              const int u = ub + tid;
              const bool live = u < units;
              const int cl = live ? (u >> lsh) : 0, c = c0 + cl;
              const SerStep st = step == 0 ? ser_step_a(V + c, PRE + c, Cs, W, N, sm.a16[c], sm.b16[c])
                                           : ser_step_b(V + c, PRE + c, Cs, W, N, sm.a16[c], sm.b16[c]);
              const uint16_t *pos = sm.pos + sm.goff[cl];
              double lm = -1.0e300;
              if (live) { /* every item's log-weight stays in val for the dense pass */
                double *Lc = sm.val + sm.goff[cl];
                for (int kk = sub; kk <= st.kb; kk += lpc) {
                  int q, n;
                  const double L = ser_item_eval(wt, st, pos, kk, &q, &n);
                  Lc[kk] = L;
                  lm = ser_fmax(lm, L);
                }
              }
              for (int o = lpc >> 1; o > 0; o >>= 1) lm = ser_fmax(lm, __shfl_xor_sync(0xffffffffu, lm, o));
              if (live && sub == 0) {
                sm.lmax[cl] = lm;
                *reinterpret_cast<uint2 *>(sm.st4 + 4 * cl) =
                    make_uint2((uint32_t)st.cur | ((uint32_t)st.bound << 16), (uint32_t)st.ocur | ((uint32_t)st.kb << 16));
              }
            }
            __syncthreads();
            PHASE_MARK(2);
            uint32_t ck_next = e0 + tid < e1 ? p.item_col[e0 + tid] : 0u; /* fetched one iteration ahead */
            for (int e = e0 + tid; e < e1; e += C) {
              const uint32_t ck = ck_next;
              if (e + C < e1) ck_next = p.item_col[e + C];
              const int cl = (int)(ck >> 16) - c0, kk = (int)(ck & 0xffffu);
              const int kb = (int)sm.st4[4 * cl + 3];
              if (kk <= kb) { /* log-weight -> run weight, in place; the run length from the postings */
                const uint16_t *pos = sm.pos + (e - kk - e0);
                const int nones = (int)sm.gones[cl], bound = (int)sm.st4[4 * cl + 1];
                int q, qprev; /* ser_item_eval's q and qprev */
                if (step) { q = kk < kb ? N - 1 - (int)pos[nones - 1 - kk] : bound; qprev = kk > 0 ? N - 1 - (int)pos[nones - kk] : -1; }
                else { q = kk < kb ? (int)pos[kk] : bound; qprev = kk > 0 ? (int)pos[kk - 1] : -1; }
                sm.val[e - e0] = ser_item_weight_cached<1>(wt, sm.val[e - e0], q - qprev, sm.lmax[cl]);
              }
            }
            __syncthreads();
            PHASE_MARK(3);
            for (int ub = 0; ub < units; ub += C) { /* cumulative weights, chunk-relative: lane `sub` owns items [k0, k1) */
              const int u = ub + tid;
              const bool live = u < units;
              const int cl = live ? (u >> lsh) : 0, c = c0 + cl;
              const int kb = (int)sm.st4[4 * cl + 3];
              double *val = sm.val + sm.goff[cl];
              const int chunk = (kb + lpc) >> lsh, k0 = min(kb + 1, sub * chunk), k1 = min(kb + 1, k0 + chunk);
              double tot = 0.0;
              if (live) for (int kk = k0; kk < k1; kk++) { tot = SER_ADD(tot, val[kk]); val[kk] = tot; }
              /* inclusive totals of the chunks, added left to right so that incl[j] == incl[j-1] + (last
               * relative prefix of chunk j) bit for bit */
              double incl = tot;
              for (int j = 1; j < lpc; j++) {
                const double t = __shfl_sync(0xffffffffu, incl, (tid & ~(lpc - 1) & 31) + j - 1);
                if (sub == j) incl = SER_ADD(t, tot);
              }
              if (live) sm.incl[u] = incl;
            }
            __syncthreads();
            PHASE_MARK(7);
            for (int cl = tid; cl < nc; cl += C) { /* inverse CDF: chunk, item inside the chunk, candidate inside the run */
              const int c = c0 + cl;
              const uint2 g4 = *reinterpret_cast<const uint2 *>(sm.st4 + 4 * cl);
              SerStep st;
              st.cur = (int)(g4.x & 0xffffu); st.bound = (int)(g4.x >> 16); st.ocur = (int)(g4.y & 0xffffu); st.kb = (int)(g4.y >> 16);
              st.nones = sm.gones[cl]; st.N = N; st.rev = step;
              const uint16_t *pos = sm.pos + sm.goff[cl];
              const double *val = sm.val + sm.goff[cl], *incl = sm.incl + (cl << lsh);
              const int taxon = p.order[c];
              double uu;
              if (p.mode == SER_MODE_REPLAY) uu = tape[sc.cursor + 6 + 2 * taxon + step];
              else {
                uint32_t o[4];
                ser_philox4x32_10((uint32_t)taxon, SER_BLK_AB, sc.sweep, 0u, p.seed, gchain, o);
                uu = step == 0 ? ser_u53(o[0], o[1]) : ser_u53(o[2], o[3]);
              }
              const double target = SER_MUL(uu, incl[lpc - 1]);
              int j = 0;
              while (j < lpc - 1 && incl[j] < target) j++;
              const double base = j ? incl[j - 1] : 0.0;
              const int chunk = (st.kb + lpc) >> lsh, k0 = min(st.kb + 1, j * chunk), k1 = min(st.kb + 1, k0 + chunk);
              int lo = k0, hi = k1 - 1; /* first item of the chunk whose cumulative weight reaches the target */
              while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (SER_ADD(base, val[mid]) >= target) hi = mid; else lo = mid + 1;
              }
              int q, n;
              const double le = SER_SUB(ser_item_eval(wt, st, pos, lo, &q, &n), sm.lmax[cl]);
              const int pick = q - n + 1 + ser_run_pick<1>(wt, n, le, lo > k0 ? SER_ADD(base, val[lo - 1]) : base, target);
              if (step == 0) { changed += pick != sm.a16[c]; sm.a16[c] = (uint16_t)pick; }
              else { changed += (N - pick) != sm.b16[c]; sm.b16[c] = (uint16_t)(N - pick); }
            }
            __syncthreads(); /* the b-step's geometry is computed under a different column -> thread map */
            PHASE_MARK(4);
          }
        }
        __syncthreads();
        const bool exact = p.sampling && s == p.sweeps_per_call - 1;
        {
          int t1 = 0, len = 0, T1, LEN, CH;
          for (int c = tid; c < M; c += C) {
            const int t1c = ser_col_popc(V + c, PRE + c, Cs, sm.a16[c], sm.b16[c]), lenc = sm.b16[c] - sm.a16[c];
            t1 += t1c; len += lenc;
            if (exact) { /* mcmc_logl's per-taxon term, mcmc.c:643-644 */
              const int f1 = p.ones[c] - t1c, f0 = lenc - t1c, t0 = N - lenc - f1;
              TERMS[p.order[c]] = SER_ADD(SER_ADD(SER_ADD(SER_MUL((double)t0, wt.cc), SER_MUL((double)f0, wt.d)), SER_MUL((double)t1c, wt.dd)),
                                          SER_MUL((double)f1, wt.c));
            }
          }
          block_sum3(t1, len, changed, sm.red, ps.buf, &T1, &LEN, &CH);
          totals_from(p, wt, T1, LEN, &sc.t0a, &sc.f0a, &sc.t1a, &sc.f1a, &sc.loglik);
          sc.counters[2] += CH;
          if (exact) {
            if (tid < 32) {
              double acc = 0.0;
              for (int m = 0; m < M; m++) acc = SER_ADD(acc, TERMS[m]);
              if (tid == 0) sm.draws_cd[7] = acc;
            }
            __syncthreads();
            sc.loglik = sm.draws_cd[7];
          }
        }

        PHASE_MARK(5);
        /* ================= 16 proposals for pi ================= */
        ps.k = 0;
        for (int prop = 0; prop < 16; prop++) {
          const int kind = prop == 0 ? 3 : ((prop - 1) % 3);
          int dt0 = 0, dt1 = 0, nz = 0, D0, D1;
          double delta;
          if (kind == 0) { /* pi1 */
            const int i = ser_draw_int(sm.draws_pi[ps.k], N);
            int j = ser_draw_int(sm.draws_pi[ps.k + 1], N - 1);
            ps.k += 2;
            if (j >= i) j++;
            const int lo = min(i, j), hi = max(i, j);
            if (ser_is_hard(hd, i) && ser_hard_count(hd, lo, hi) > 1) continue;
            auto redo = [&](int c, int *x0, int *x1) { ser_pi1_delta(V + c, Cs, sm.a16[c], sm.b16[c], i, j, x0, x1); };
            for (int c = tid; c < M; c += C) { int x0, x1; redo(c, &x0, &x1); dt0 += x0; dt1 += x1; nz |= (x0 | x1) != 0; }
            if (!mh_decide_big(p, sm, wt, ps, TERMS, dt0, dt1, nz, exact, &D0, &D1, &delta, redo)) continue;
            for (int c = tid; c <= M; c += C) {
              if (c < M) { int a = sm.a16[c], b = sm.b16[c]; ser_pi1_apply_ab(&a, &b, i, j); sm.a16[c] = (uint16_t)a; sm.b16[c] = (uint16_t)b; }
              ser_col_rotate(V + c, Cs, W, i, j, PRE + c);
              if (c == M) ser_hard_list(V + M, Cs, W, sm.hp);
            }
            for (int n = lo + tid; n <= hi; n += C) sm.tmp16[n] = sm.rpi[i < j ? (n < j ? n + 1 : i) : (n > j ? n - 1 : i)];
            __syncthreads();
            for (int n = lo + tid; n <= hi; n += C) sm.rpi[n] = sm.tmp16[n];
            sc.counters[3]++;
          } else if (kind == 1 || kind == 3) { /* pi2 */
            int i, j;
            if (kind == 1) {
              i = ser_draw_int(sm.draws_pi[ps.k], N);
              j = ser_draw_int(sm.draws_pi[ps.k + 1], N - 1);
              ps.k += 2;
              if (j >= i) j++;
              else { const int t = i; i = j; j = t; }
            } else {
              i = ser_draw_int(sm.draws_pi[ps.k], N - 1);
              ps.k += 1;
              j = i + 1;
            }
            if (ser_hard_count(hd, i, j) > 1) continue;
            const int inc1 = ser_draw_int(sm.draws_pi[ps.k], 2), inc2 = ser_draw_int(sm.draws_pi[ps.k + 1], 2);
            ps.k += 2;
            auto redo = [&](int c, int *x0, int *x1) { ser_pi2_delta(V + c, PRE + c, Cs, sm.a16[c], sm.b16[c], i, j, inc1, inc2, x0, x1); };
            for (int c = tid; c < M; c += C) { int x0, x1; redo(c, &x0, &x1); dt0 += x0; dt1 += x1; nz |= (x0 | x1) != 0; }
            if (!mh_decide_big(p, sm, wt, ps, TERMS, dt0, dt1, nz, exact, &D0, &D1, &delta, redo)) continue;
            for (int c = tid; c <= M; c += C) {
              if (c < M) {
                int a = sm.a16[c], b = sm.b16[c];
                const int ain = ser_in_window(a, i, j + 1, inc1, inc2), bin = ser_in_window(b, i, j + 1, inc1, inc2);
                ser_mirror_ab(a, b, ain, bin, i + j + 1, &a, &b);
                sm.a16[c] = (uint16_t)a; sm.b16[c] = (uint16_t)b;
              }
              ser_col_reverse(V + c, Cs, W, i, j, PRE + c);
              if (c == M) ser_hard_list(V + M, Cs, W, sm.hp);
            }
            for (int n = i + tid; n <= j; n += C) sm.tmp16[n] = sm.rpi[i + j - n];
            __syncthreads();
            for (int n = i + tid; n <= j; n += C) sm.rpi[n] = sm.tmp16[n];
            sc.counters[kind == 1 ? 4 : 5]++;
          } else { /* pi3 */
            const int nfree = N - p.nh;
            if (nfree < 2) continue;
            const int r1 = ser_draw_int(sm.draws_pi[ps.k], nfree), r2 = ser_draw_int(sm.draws_pi[ps.k + 1], nfree - 1);
            ps.k += 2;
            int ir, jr;
            if (r1 <= r2) { ir = r1; jr = r2 + 1; } else { ir = r2; jr = r1; }
            const SerPi3 g = ser_pi3_window(hd, ir, jr);
            const int inc1 = ser_draw_int(sm.draws_pi[ps.k], 2), inc2 = ser_draw_int(sm.draws_pi[ps.k + 1], 2);
            ps.k += 2;
            auto redo = [&](int c, int *x0, int *x1) { ser_pi3_delta(V + c, PRE + c, Cs, hd, g, sm.a16[c], sm.b16[c], inc1, inc2, x0, x1); };
            for (int c = tid; c < M; c += C) { int x0, x1; redo(c, &x0, &x1); dt0 += x0; dt1 += x1; nz |= (x0 | x1) != 0; }
            if (!mh_decide_big(p, sm, wt, ps, TERMS, dt0, dt1, nz, exact, &D0, &D1, &delta, redo)) continue;
            for (int n = g.i + tid; n <= g.j; n += C) sm.perm16[n] = (uint16_t)ser_pi3_perm(hd, g, n);
            __syncthreads();
            for (int c = tid; c < M; c += C) {
              int a = sm.a16[c], b = sm.b16[c];
              const int ain = ser_in_window(a, g.i, g.j + 1, inc1, inc2), bin = ser_in_window(b, g.i, g.j + 1, inc1, inc2);
              ser_mirror_ab(a, b, ain, bin, g.i + g.j + 1, &a, &b);
              sm.a16[c] = (uint16_t)a; sm.b16[c] = (uint16_t)b;
              ser_col_permute(V + c, Cs, W, g.i, g.j, sm.perm16, PRE + c);
            }
            for (int n = g.i + tid; n <= g.j; n += C) sm.tmp16[n] = sm.rpi[sm.perm16[n]];
            __syncthreads();
            for (int n = g.i + tid; n <= g.j; n += C) sm.rpi[n] = sm.tmp16[n];
            sc.counters[6]++;
          }
          sc.t0a += D0; sc.f0a -= D0; sc.t1a += D1; sc.f1a -= D1;
          sc.loglik = SER_ADD(sc.loglik, delta);
          __syncthreads();
        }

        if (p.mode == SER_MODE_REPLAY) sc.cursor += 6 + 2 * (long long)M + ps.k;
        else sc.sweep++;
        sc.counters[7]++;
        PHASE_MARK(6);
      }
      if (sc.flags & 1) break;

      if (p.sampling) {
        const int sidx = sc.n_samples;
        if (sidx < p.max_samples) {
          const size_t row = (size_t)chain * p.max_samples + sidx;
          if (p.store >= SER_STORE_PI)
            for (int pos = tid; pos < N; pos += C) p.samp_pi[row * N + sm.rpi[pos]] = (uint16_t)pos;
          if (p.store >= SER_STORE_FULL) {
            for (int c = tid; c < M; c += C) { p.samp_a[row * M + p.order[c]] = sm.a16[c]; p.samp_b[row * M + p.order[c]] = sm.b16[c]; }
            if (tid == 0) { p.samp_cdl[row * 3 + 0] = sc.c; p.samp_cdl[row * 3 + 1] = sc.d; p.samp_cdl[row * 3 + 2] = sc.loglik; }
          }
        }
        sc.sum_negll = SER_ADD(sc.sum_negll, -sc.loglik);
        sc.sum_ec = SER_ADD(sc.sum_ec, exp(sc.c));
        sc.sum_ed = SER_ADD(sc.sum_ed, exp(sc.d));
        sc.n_samples++;
      }
    }

    __syncthreads();
    for (int c = tid; c < M; c += C) {
      p.ab[(size_t)chain * 2 * p.Mpad + c] = sm.a16[c];
      p.ab[(size_t)chain * 2 * p.Mpad + p.Mpad + c] = sm.b16[c];
    }
    for (int n = tid; n < N; n += C) p.rpi[(size_t)chain * p.Npad + n] = sm.rpi[n];
    if (tid == 0) p.scal[chain] = sc;
  }
}

/* ------------------------------------------------------------------ export / check kernels */
/* int32 view of one chain's state incl. the derived per-taxon counts (mcmc_count01) */
__global__ void ser_export_kernel(KParams p, int chain, int *out_a, int *out_b, int *out_pi, int *out_rpi, int *out_cnt)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AuxSmem sm;
  aux_layout(&sm, smem_raw, p.N);
  const int tid = threadIdx.x, C = blockDim.x;
  for (int n = tid; n < p.N; n += C) sm.rpi[n] = p.rpi[(size_t)chain * p.Npad + n];
  __syncthreads();
  for (int n = tid; n < p.N; n += C) { out_rpi[n] = sm.rpi[n]; out_pi[sm.rpi[n]] = n; }
  for (int c = tid; c < p.M; c += C) {
    const int a = p.ab[(size_t)chain * 2 * p.Mpad + c], b = p.ab[(size_t)chain * 2 * p.Mpad + p.Mpad + c];
    const int t1 = taxon_count(p, sm.rpi, c, a, b), ones = p.ones[c];
    const int tx = p.order[c];
    out_a[tx] = a; out_b[tx] = b;
    out_cnt[tx] = p.N - (b - a) - (ones - t1); out_cnt[p.M + tx] = (b - a) - t1;
    out_cnt[2 * p.M + tx] = t1; out_cnt[3 * p.M + tx] = ones - t1;
  }
}

/* mcmc_consistent (mcmc.c:999-1094) for every chain; flags |= 2 a/b range, 4 permutation,
 * 8 hard-site order, 16 totals / log-likelihood */
__global__ void ser_check_kernel(KParams p, int *bad_count)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AuxSmem sm;
  aux_layout(&sm, smem_raw, p.N);
  const int chain = blockIdx.x, tid = threadIdx.x, C = blockDim.x, N = p.N, M = p.M;
  __shared__ int s_flags;
  if (tid == 0) s_flags = 0;
  for (int n = tid; n < N; n += C) { sm.rpi[n] = p.rpi[(size_t)chain * p.Npad + n]; sm.tmp16[n] = 0xffff; }
  __syncthreads();
  for (int n = tid; n < N; n += C) {
    const int site = sm.rpi[n];
    if (site >= N) { atomicOr(&s_flags, 4); sm.rpi[n] = 0; }
    else sm.tmp16[site] = (uint16_t)n; /* pi */
  }
  __syncthreads();
  for (int n = tid; n < N; n += C) if (sm.tmp16[n] == 0xffff) atomicOr(&s_flags, 4);
  if (tid == 0) { /* hard sites in increasing position in file order */
    int last = -1, cnt = 0;
    for (int n = 0; n < N; n++)
      if (p.hard[n]) { cnt++; if (last >= 0 && (int)sm.tmp16[n] < last) s_flags |= 8; last = sm.tmp16[n]; }
    if (cnt != p.nh) atomicOr(&s_flags, 8);
  }
  int t1 = 0, len = 0;
  double llp = 0.0; /* manycd: the log-likelihood is a sum of per-taxon terms */
  for (int c = tid; c < M; c += C) {
    const int a = p.ab[(size_t)chain * 2 * p.Mpad + c], b = p.ab[(size_t)chain * 2 * p.Mpad + p.Mpad + c];
    if (!(0 <= a && a <= b && b <= N)) atomicOr(&s_flags, 2);
    else {
      const int k1 = taxon_count(p, sm.rpi, c, a, b);
      t1 += k1; len += b - a;
      if (p.manycd) {
        const double *cd = p.cd4 + (size_t)chain * 4 * p.Mpad + c;
        const int f1 = p.ones[c] - k1, f0 = (b - a) - k1, t0 = N - (b - a) - f1;
        llp += (double)t0 * cd[p.Mpad] + (double)f0 * cd[2 * p.Mpad] + (double)k1 * cd[3 * p.Mpad] + (double)f1 * cd[0];
      }
    }
  }
  int buf = 0, T1, LEN, dummy;
  block_sum3(t1, len, 0, sm.red, buf, &T1, &LEN, &dummy);
  __shared__ double s_ll[32];
  if (p.manycd) {
    for (int o = 16; o > 0; o >>= 1) llp += __shfl_xor_sync(0xffffffffu, llp, o);
    if ((tid & 31) == 0) s_ll[tid >> 5] = llp;
    __syncthreads();
  }
  if (tid == 0) {
    const ChainScalars sc = p.scal[chain];
    SerWeights wt;
    set_weights(wt, sc.c, sc.cc, sc.d, sc.dd);
    int t0a, f0a, t1a, f1a;
    double ll;
    totals_from(p, wt, T1, LEN, &t0a, &f0a, &t1a, &f1a, &ll);
    if (p.manycd) { ll = 0.0; for (int w = 0; w < (C + 31) / 32; w++) ll += s_ll[w]; }
    /* the reference allows 1e-8 absolute (mcmc.c:1084); on large matrices |loglik| ~ 1e6 and the taxon-order
     * sum of a sampled sweep differs from this recount's closed form by more than that in the last bits */
    if (t0a != sc.t0a || f0a != sc.f0a || t1a != sc.t1a || f1a != sc.f1a || fabs(ll - sc.loglik) > 1e-8 + 1e-12 * fabs(ll)) s_flags |= 16;
    const int fl = s_flags | (sc.flags & 1);
    if (fl) atomicAdd(bad_count, 1);
    p.scal[chain].flags = (sc.flags & 1) | fl;
  }
}

/* ------------------------------------------------------------------ cross-chain kernels */
__global__ void ser_stats_kernel(const ChainScalars *scal, int n, double *out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = scal[i].n_samples > 0 ? scal[i].sum_negll / (double)scal[i].n_samples : 0.0;
}

__device__ double block_reduce_d(double v, double *sh, int op) /* 0 sum, 1 min */
{
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    v = op ? fmin(v, t) : v + t;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < nw; w++) r = op ? fmin(r, sh[w]) : r + sh[w];
  return r;
}

/* choose_chains (script.py:70-99) on one CTA: min, population sigma over all chains, the k
 * smallest inside (min-sigma, min+sigma), ids ascending */
__global__ void ser_select_kernel(const double *e, int n, int k, int *chosen, double *info)
{
  __shared__ double sh[32];
  __shared__ double s_best;
  __shared__ int s_besti;
  const int tid = threadIdx.x, nt = blockDim.x;
  double s = 0.0, mn = 1.0e300;
  for (int i = tid; i < n; i += nt) { s += e[i]; mn = fmin(mn, e[i]); }
  const double mean = block_reduce_d(s, sh, 0) / (double)n;
  mn = block_reduce_d(mn, sh, 1);
  double v = 0.0;
  for (int i = tid; i < n; i += nt) { const double d = e[i] - mean; v += d * d; }
  const double sigma = sqrt(block_reduce_d(v, sh, 0) / (double)n);
  const double lo = mn - sigma, hi = mn + sigma;
  /* k rounds of arg-min over the not-yet-taken candidates, ties by lower id */
  double last_v = -1.0e300;
  int last_i = -1, found = 0;
  for (int r = 0; r < k; r++) {
    double bv = 1.0e300;
    int bi = -1;
    for (int i = tid; i < n; i += nt) {
      const double x = e[i];
      if (!(x > lo && x < hi)) continue;
      if (x < last_v || (x == last_v && i <= last_i)) continue;
      if (x < bv || (x == bv && i < bi)) { bv = x; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    __syncthreads();
    if (tid == 0) { s_best = 1.0e300; s_besti = -1; }
    __syncthreads();
    for (int w = 0; w < (nt >> 5); w++) {
      if ((tid >> 5) == w && (tid & 31) == 0 && bi >= 0)
        if (s_besti < 0 || bv < s_best || (bv == s_best && bi < s_besti)) { s_best = bv; s_besti = bi; }
      __syncthreads();
    }
    if (s_besti < 0) break;
    last_v = s_best; last_i = s_besti;
    if (tid == 0) chosen[found] = s_besti;
    found++;
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) {
    for (int r = found; r < k; r++) chosen[r] = -1;
    for (int x = 1; x < found; x++) { /* ids ascending (script.py:98) */
      const int key = chosen[x];
      int y = x - 1;
      while (y >= 0 && chosen[y] > key) { chosen[y + 1] = chosen[y]; y--; }
      chosen[y + 1] = key;
    }
    info[0] = (double)found; info[1] = mn; info[2] = sigma;
  }
}

/* pair-order counts (script.py:178-189) for one chosen chain per blockIdx.z */
__global__ void ser_po_kernel(const uint16_t *samp_pi, int N, int max_samples, int n_samples, const int *chosen,
                              int chain_offset, int n_local, int *counts)
{
  const int g = chosen[blockIdx.z];
  if (g < chain_offset || g >= chain_offset + n_local) return;
  const int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || j >= N) return;
  const uint16_t *pi = samp_pi + (size_t)(g - chain_offset) * max_samples * N;
  int c = 0;
  for (int t = 0; t < n_samples; t++) c += pi[(size_t)t * N + i] < pi[(size_t)t * N + j];
  counts[((size_t)blockIdx.z * N + i) * N + j] = (i == j) ? -n_samples : c;
}

/* posterior sums over the stored samples of one chosen chain per block (script.py:129-152, :230-276) */
__global__ void ser_posterior_kernel(const uint16_t *samp_pi, const uint16_t *samp_a, const uint16_t *samp_b, int N, int M,
                                     int max_samples, int n_samples, const int *chosen, int chain_offset, int n_local,
                                     long long *corr_num, int *pi_sum, int *a_sum, int *b_sum)
{
  const int g = chosen[blockIdx.x];
  if (g < chain_offset || g >= chain_offset + n_local) return;
  const size_t base = (size_t)(g - chain_offset) * max_samples;
  long long s = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    int acc = 0;
    for (int t = 0; t < n_samples; t++) acc += samp_pi[(base + t) * N + i];
    pi_sum[(size_t)blockIdx.x * N + i] = acc;
    s += (long long)i * acc;
  }
  if (a_sum && samp_a)
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      int sa = 0, sb = 0;
      for (int t = 0; t < n_samples; t++) { sa += samp_a[(base + t) * M + m]; sb += samp_b[(base + t) * M + m]; }
      a_sum[(size_t)blockIdx.x * M + m] = sa;
      if (b_sum) b_sum[(size_t)blockIdx.x * M + m] = sb;
    }
  atomicAdd((unsigned long long *)&corr_num[blockIdx.x], (unsigned long long)s);
}

/* alive[c][j][m] = #{t : a_t(m) <= j <= b_t(m)} over the stored samples of chosen chain c
 * (script.py:321-329; closed at b, as the reference tests it).  One thread per taxon: +1 / -1
 * marks at a and b+1 in its own column of the slab, then a running sum down the positions. */
__global__ void ser_alive_kernel(const uint16_t *samp_a, const uint16_t *samp_b, int N, int M, int max_samples, int n_samples,
                                 const int *chosen, int chain_offset, int n_local, int *alive)
{
  const int g = chosen[blockIdx.x];
  if (g < chain_offset || g >= chain_offset + n_local) return;
  const int m = blockIdx.y * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const size_t base = (size_t)(g - chain_offset) * max_samples;
  int *col = alive + (size_t)blockIdx.x * N * M + m;
  for (int j = 0; j < N; j++) col[(size_t)j * M] = 0;
  for (int t = 0; t < n_samples; t++) {
    const int a = samp_a[(base + t) * M + m], b = samp_b[(base + t) * M + m];
    if (a < N) col[(size_t)a * M] += 1;
    if (b + 1 < N) col[(size_t)(b + 1) * M] -= 1;
  }
  int acc = 0;
  for (int j = 0; j < N; j++) { acc += col[(size_t)j * M]; col[(size_t)j * M] = acc; }
}

/* ------------------------------------------------------------------ micro-benchmarks */
__global__ void mb_fp64_kernel(double *out, int iters)
{
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void mb_lds_kernel(unsigned *out, int iters)
{
  __shared__ uint4 buf[1024];
  buf[threadIdx.x] = make_uint4(threadIdx.x, 1, 2, 3);
  __syncthreads();
  uint4 acc = make_uint4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint4 v = buf[(idx + u * 32) & 1023];
      acc.x += v.x; acc.y ^= v.y; acc.z += v.z; acc.w ^= v.w;
    }
    idx = (idx + acc.y) & 1023;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
__global__ void mb_popc_kernel(unsigned *out, int iters)
{
  unsigned x0 = threadIdx.x + 1, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int i = 0; i < iters; i++) {
    s0 += __popc(x0 ^ s3); s1 += __popc(x1 ^ s0); s2 += __popc(x2 ^ s1); s3 += __popc(x3 ^ s2);
    s0 += __popc(x0 + s2); s1 += __popc(x1 + s3); s2 += __popc(x2 + s0); s3 += __popc(x3 + s1);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

/* ================================================================== host side: the run object */
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ser_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return SER_E_CUDA;                                                                        \
    }                                                                                           \
  } while (0)

struct ser_run {
  ser_run_config cfg;
  KParams kp;
  int N, M, W, C, nh;
  uint8_t *h_hard;
  uint32_t *d_Xs;
  uint8_t *d_hard;
  int *d_ones, *d_off;
  uint16_t *d_order;
  uint32_t *d_item_col;
  uint16_t *d_ab, *d_rpi;
  ChainScalars *d_scal;
  double *d_tape;
  unsigned long long *d_tape_off;
  uint16_t *d_samp_a, *d_samp_b, *d_samp_pi;
  double *d_samp_cdl, *d_cd4, *d_samp_cd_all;
  size_t smem_many;
  int *d_scratch_i; /* export buffers: a,b,pi,rpi,cnt[4M] */
  int *d_bad;
  cudaStream_t stream;
  cudaEvent_t ev_start, ev_stop;
  int timing_open;
  double elapsed_ms;
  long long launches;
  size_t smem_sweep, smem_init, smem_small, smem_big;
  int Caux, big, big_threads, big_slots, variant, variant_many;
  uint32_t *d_gV;
  uint16_t *d_gpre;
  int *d_bgrp;
  int initialized, have_tapes;
};

/* Stream-ordered pool allocation: cudaMalloc/cudaFree are synchronous driver calls with erratic
 * latency (tens to hundreds of ms on shared hosts); the default memory pool with an unlimited
 * release threshold keeps freed blocks cached, so creating / destroying runs and the temporary
 * buffers of the cross-chain steps cost microseconds after the first use. */
static cudaError_t pool_alloc(void **ptr, size_t bytes, cudaStream_t stream)
{
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    configured[dev] = true;
  }
  return cudaMallocAsync(ptr, bytes ? bytes : 1, stream);
}
#define POOL_ALLOC(ptr, bytes) pool_alloc((void **)(ptr), (bytes), run->stream)

static int set_device(const ser_run *run) { CUDA_TRY(cudaSetDevice(run->cfg.device)); return SER_OK; }

static void mark_launch(ser_run *run)
{
  if (!run->timing_open) { cudaEventRecord(run->ev_start, run->stream); run->timing_open = 1; }
  run->launches++;
}

/* Column groups of the Gibbs step: the item weights of a step go through a buffer of Ival doubles,
 * one group of columns at a time.  Fewer, larger groups = fewer barriers; a smaller buffer = more
 * resident chains per SM (measured on B200: +15-20 % per extra resident CTA, -4 % per extra group).
 * Take the fewest groups that reach the best residency; SER_SWEEP_GROUPS forces a count. */
template <typename K>
static int choose_groups(KParams &kp, const std::vector<int> &off, int M, int N, int W, int C, int manycd, K kernel, size_t *smem_out)
{
  int best_occ = 0, force = 0;
  if (const char *v = getenv("SER_SWEEP_GROUPS")) force = std::max(1, std::min(SER_MAX_GROUPS, atoi(v)));
  for (int G = 1; G <= SER_MAX_GROUPS; G++) {
    if (force && G != force) continue;
    int gc[SER_MAX_GROUPS + 1], ge[SER_MAX_GROUPS + 1], ng = 0, ival = 0;
    gc[0] = 0; ge[0] = 0;
    for (int c = 0; c < M && ng < G; c++) /* close a group at the column where its share of the items is reached */
      if (c + 1 == M || off[c + 1] >= (long long)kp.I * (ng + 1) / G) {
        ng++; gc[ng] = c + 1; ge[ng] = off[c + 1];
        ival = std::max(ival, ge[ng] - ge[ng - 1]);
      }
    const size_t sz = smem_layout(nullptr, nullptr, N, W, C, kp.I, manycd, ival);
    if (sz > 227 * 1024) continue;
    int occ = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sz) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, C, sz) != cudaSuccess) {
      ser_set_error("ser_run_create: occupancy query failed: %s", cudaGetErrorString(cudaGetLastError()));
      return SER_E_CUDA;
    }
    if (occ > best_occ) {
      best_occ = occ;
      kp.n_groups = ng; kp.Ival = ival;
      memcpy(kp.grp_c, gc, sizeof(gc)); memcpy(kp.grp_e, ge, sizeof(ge));
      *smem_out = sz;
    }
  }
  if (!best_occ) { ser_set_error("ser_run_create: the chain state does not fit shared memory"); return SER_E_ARG; }
  return SER_OK;
}

extern "C" int ser_run_create(const ser_dataset *ds, const ser_run_config *cfg, ser_run **out)
{
  if (!ds || !cfg || !out) { ser_set_error("ser_run_create: null argument"); return SER_E_ARG; }
  const int N = ds->N, M = ds->M;
  if (N < 2 || N > SER_MAX_SITES) { ser_set_error("ser_run_create: N=%d outside [2,%d]", N, SER_MAX_SITES); return SER_E_ARG; }
  if (M < 1 || M > SER_MAX_TAXA) { ser_set_error("ser_run_create: M=%d outside [1,%d]", M, SER_MAX_TAXA); return SER_E_ARG; }
  if (cfg->n_chains < 1 || cfg->sweeps_per_call < 1) { ser_set_error("ser_run_create: n_chains and sweeps_per_call must be >= 1"); return SER_E_ARG; }
  if (cfg->mode != SER_MODE_FREE && cfg->mode != SER_MODE_REPLAY) { ser_set_error("ser_run_create: bad mode"); return SER_E_ARG; }
  if (cfg->manycd != 0 && cfg->manycd != 1) { ser_set_error("ser_run_create: manycd must be 0 or 1"); return SER_E_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    ser_set_error("ser_run_create: no CUDA device (this library has no CPU path)");
    return SER_E_CUDA;
  }
  ser_run *run = (ser_run *)calloc(1, sizeof(ser_run));
  run->cfg = *cfg;
  run->N = N; run->M = M; run->nh = ds->nh;
  run->W = N / 32 + 1;
  run->C = ((M + 1) + 31) / 32 * 32;
  run->Caux = std::min(1024, (M + 31) / 32 * 32);
  /* large-shape path when one thread per column does not fit a CTA or its shared memory;
   * SER_FORCE_BIG=<threads> forces it (tests run the whole parity suite through it) */
  run->big_threads = 1024;
  if (const char *fb = getenv("SER_FORCE_BIG")) { run->big = 1; if (atoi(fb) >= 32) run->big_threads = std::min(1024, atoi(fb) / 32 * 32); }
  if (run->C > 1024) { run->big = 1; run->C = 1024; }
  if (const char *bt = getenv("SER_BIG_THREADS")) { if (atoi(bt) >= 32) run->big_threads = std::min(1024, atoi(bt) / 32 * 32); }
  CUDA_TRY(cudaSetDevice(cfg->device));
  CUDA_TRY(cudaStreamCreateWithFlags(&run->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&run->ev_start));
  CUDA_TRY(cudaEventCreate(&run->ev_stop));

  KParams &kp = run->kp;
  memset(&kp, 0, sizeof(kp));
  kp.N = N; kp.M = M; kp.W = run->W; kp.C = run->C; kp.nh = ds->nh;
  kp.Mw = (M + 31) / 32; kp.Npad = (N + 7) / 8 * 8; kp.Mpad = (M + 7) / 8 * 8;
  kp.mode = cfg->mode; kp.chain_offset = cfg->chain_offset; kp.seed = cfg->seed;
  kp.sweeps_per_call = cfg->sweeps_per_call; kp.store = cfg->store; kp.max_samples = cfg->max_samples;
  /* initial c, d and the flooring constant with the HOST libm: the bits the reference gets */
  kp.c0 = log(.01); kp.d0 = log(.3);
  kp.cc0 = log(1. - exp(kp.c0)); kp.dd0 = log(1. - exp(kp.d0));
  kp.eps = exp(-32.236191301916641); /* exp(LOGEPSILON), mcmc.h:26 */

  /* Columns = taxa sorted by number of occurrences (descending), so that the threads of a warp
   * own taxa with similar item counts; `order` maps a column back to its taxon.  Site-major bit
   * matrix over columns, ones per column, and the static item tables of the Gibbs step. */
  std::vector<int> tones(M, 0);
  long long ones_total = 0;
  for (int n = 0; n < N; n++)
    for (int m = 0; m < M; m++)
      if (ds->X[(size_t)n * M + m]) { tones[m]++; ones_total++; }
  std::vector<uint16_t> order(M);
  for (int m = 0; m < M; m++) order[m] = (uint16_t)m;
  std::stable_sort(order.begin(), order.end(), [&](uint16_t x, uint16_t y) { return tones[x] > tones[y]; });
  std::vector<uint32_t> Xs((size_t)N * kp.Mw, 0u);
  std::vector<int> ones(M, 0), off(M + 1, 0);
  for (int c = 0; c < M; c++) {
    ones[c] = tones[order[c]];
    off[c + 1] = off[c] + ones[c] + 1;
    for (int n = 0; n < N; n++)
      if (ds->X[(size_t)n * M + order[c]]) Xs[(size_t)n * kp.Mw + (c >> 5)] |= 1u << (c & 31);
  }
  kp.I = off[M];
  std::vector<uint32_t> item_col(kp.I);
  for (int c = 0; c < M; c++)
    for (int e = off[c]; e < off[c + 1]; e++) item_col[e] = ((uint32_t)c << 16) | (uint32_t)(e - off[c]);
  kp.ones_total = ones_total;
  run->h_hard = (uint8_t *)malloc(N);
  memcpy(run->h_hard, ds->hard, N);

  const size_t nc = (size_t)cfg->n_chains;
  CUDA_TRY(POOL_ALLOC(&run->d_Xs, Xs.size() * 4));
  CUDA_TRY(POOL_ALLOC(&run->d_hard, N));
  CUDA_TRY(POOL_ALLOC(&run->d_ones, M * sizeof(int)));
  CUDA_TRY(POOL_ALLOC(&run->d_off, (M + 1) * sizeof(int)));
  CUDA_TRY(POOL_ALLOC(&run->d_order, M * sizeof(uint16_t)));
  CUDA_TRY(POOL_ALLOC(&run->d_item_col, (size_t)kp.I * sizeof(uint32_t)));
  CUDA_TRY(POOL_ALLOC(&run->d_ab, nc * 2 * kp.Mpad * sizeof(uint16_t)));
  CUDA_TRY(POOL_ALLOC(&run->d_rpi, nc * kp.Npad * sizeof(uint16_t)));
  CUDA_TRY(POOL_ALLOC(&run->d_scal, nc * sizeof(ChainScalars)));
  CUDA_TRY(POOL_ALLOC(&run->d_scratch_i, (size_t)(2 * N + 6 * M + 16) * sizeof(int)));
  CUDA_TRY(POOL_ALLOC(&run->d_bad, sizeof(int)));
  CUDA_TRY(cudaMemcpyAsync(run->d_Xs, Xs.data(), Xs.size() * 4, cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(run->d_hard, ds->hard, N, cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(run->d_ones, ones.data(), M * sizeof(int), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(run->d_off, off.data(), (M + 1) * sizeof(int), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(run->d_order, order.data(), M * sizeof(uint16_t), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(run->d_item_col, item_col.data(), (size_t)kp.I * sizeof(uint32_t), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemsetAsync(run->d_scal, 0, nc * sizeof(ChainScalars), run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  if (cfg->store >= SER_STORE_PI && cfg->max_samples > 0) {
    CUDA_TRY(POOL_ALLOC(&run->d_samp_pi, nc * cfg->max_samples * N * sizeof(uint16_t)));
  }
  if (cfg->store >= SER_STORE_FULL && cfg->max_samples > 0) {
    CUDA_TRY(POOL_ALLOC(&run->d_samp_a, nc * cfg->max_samples * M * sizeof(uint16_t)));
    CUDA_TRY(POOL_ALLOC(&run->d_samp_b, nc * cfg->max_samples * M * sizeof(uint16_t)));
    CUDA_TRY(POOL_ALLOC(&run->d_samp_cdl, nc * cfg->max_samples * 3 * sizeof(double)));
  }
  kp.Xs = run->d_Xs; kp.hard = run->d_hard; kp.ones = run->d_ones;
  kp.order = run->d_order; kp.off = run->d_off; kp.item_col = run->d_item_col;
  kp.ab = run->d_ab; kp.rpi = run->d_rpi; kp.scal = run->d_scal;
  kp.samp_a = run->d_samp_a; kp.samp_b = run->d_samp_b; kp.samp_pi = run->d_samp_pi; kp.samp_cdl = run->d_samp_cdl;
  if (cfg->manycd) { /* per-taxon c, d: state rows and, with the full store, per-sample rows */
    kp.manycd = 1;
    CUDA_TRY(POOL_ALLOC(&run->d_cd4, nc * 4 * kp.Mpad * sizeof(double)));
    if (cfg->store >= SER_STORE_FULL && cfg->max_samples > 0)
      CUDA_TRY(POOL_ALLOC(&run->d_samp_cd_all, nc * cfg->max_samples * 2 * M * sizeof(double)));
    kp.cd4 = run->d_cd4; kp.samp_cd_all = run->d_samp_cd_all;
  }

  run->smem_small = aux_layout(nullptr, nullptr, N);
  run->smem_init = run->smem_small + sizeof(double) * 2 * N + sizeof(uint16_t) * 3 * N + 64;
  CUDA_TRY(cudaFuncSetAttribute(ser_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)run->smem_init));
  if (cfg->manycd) {
    run->smem_many = smem_layout(nullptr, nullptr, N, run->W, run->C, run->kp.I, 1);
    /* the per-taxon kernel needs its 80 registers (2.7 M vs 2.5 M sweeps/s on g2s2 with 64): size the
     * groups for the instantiation that will run */
    if (!run->big && run->smem_many <= 227 * 1024 &&
        (run->C <= 384 ? choose_groups(kp, off, M, N, run->W, run->C, 1, ser_sweep_kernel<384, 2, true>, &run->smem_many)
                       : choose_groups(kp, off, M, N, run->W, run->C, 1, ser_sweep_kernel<1024, 1, true>, &run->smem_many)))
      return SER_E_ARG;
    if (run->big || run->smem_many > 227 * 1024) {
      ser_set_error("ser_run_create: manycd=1 needs one thread per taxon and %zu B of shared memory per chain (M <= 1023)", run->smem_many);
      return SER_E_ARG;
    }
    CUDA_TRY(cudaFuncSetAttribute(ser_sweep_kernel<1024, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)run->smem_many));
    CUDA_TRY(cudaFuncSetAttribute(ser_sweep_kernel<384, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)run->smem_many));
    int occ64 = 0, occ85 = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, ser_sweep_kernel<1024, 1, true>, run->C, run->smem_many));
    if (run->C <= 384) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ85, ser_sweep_kernel<384, 2, true>, run->C, run->smem_many));
    run->variant_many = (occ85 >= occ64 && occ85 > 0) ? 1 : 0;
  }
  if (!run->big) {
    run->smem_sweep = smem_layout(nullptr, nullptr, N, run->W, run->C, run->kp.I);
    if (run->smem_sweep > 227 * 1024) run->big = 1; /* columns + items do not fit: use the L2-resident variant */
    else {
      /* two register budgets: 64 regs (any block size) and 85 regs (blocks <= 384 threads, two of
       * them resident); take the one with more resident CTAs, the roomier one on a tie */
      if (!cfg->manycd && choose_groups(kp, off, M, N, run->W, run->C, 0, ser_sweep_kernel<1024, 1, false>, &run->smem_sweep)) return SER_E_ARG;
      CUDA_TRY(cudaFuncSetAttribute(ser_sweep_kernel<1024, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)run->smem_sweep));
      CUDA_TRY(cudaFuncSetAttribute(ser_sweep_kernel<384, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)run->smem_sweep));
      int occ64 = 0, occ85 = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, ser_sweep_kernel<1024, 1, false>, run->C, run->smem_sweep));
      if (run->C <= 384) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ85, ser_sweep_kernel<384, 2, false>, run->C, run->smem_sweep));
      run->variant = (occ85 >= occ64 && occ85 > 0) ? 1 : 0;
      if (const char *v = getenv("SER_SWEEP_VARIANT")) run->variant = (atoi(v) == 1 && run->C <= 384) ? 1 : 0;
    }
  }
  if (run->big) {
    /* column groups of the Gibbs phase: as many items as the shared-memory budget holds
     * (SER_BIG_SMEM_KB, default 220), at most 1024 columns, never splitting a column */
    int budget_kb = 220;
    if (const char *v = getenv("SER_BIG_SMEM_KB")) budget_kb = std::max(16, std::min(227, atoi(v)));
    const int gcap = std::max(std::min(M, 1024), run->big_threads); /* also bounds (columns x lanes per column) */
    const size_t fixed = big_layout(nullptr, nullptr, N, M, 0, gcap);
    long long icap = ((long long)budget_kb * 1024 - (long long)fixed - 64) / 10 / 32 * 32; /* val 8 + pos 2 bytes per item */
    icap = std::min<long long>(icap, (long long)(kp.I + 31) / 32 * 32);
    if (icap < std::max(N + 1, M)) { /* one whole column, and the M per-taxon terms of the exact sums */
      ser_set_error("ser_run_create: shape needs %zu B of shared memory per chain before any item", fixed);
      return SER_E_ARG;
    }
    std::vector<int> bgrp;
    {
      int c0 = 0;
      while (c0 < M) {
        int c1 = c0;
        while (c1 < M && c1 - c0 < gcap && off[c1 + 1] - off[c0] <= icap) c1++;
        bgrp.push_back(c0); bgrp.push_back(off[c0]);
        c0 = c1;
      }
      bgrp.push_back(M); bgrp.push_back(off[M]);
    }
    kp.big_ng = (int)bgrp.size() / 2 - 1; kp.big_icap = (int)icap; kp.big_gcap = gcap;
    CUDA_TRY(POOL_ALLOC(&run->d_bgrp, bgrp.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(run->d_bgrp, bgrp.data(), bgrp.size() * sizeof(int), cudaMemcpyHostToDevice, run->stream));
    CUDA_TRY(cudaStreamSynchronize(run->stream));
    kp.bgrp = run->d_bgrp;
    run->smem_big = big_layout(nullptr, nullptr, N, M, (int)icap, gcap);
    if (run->smem_big > 227 * 1024) { ser_set_error("ser_run_create: shape needs %zu B of shared memory per chain", run->smem_big); return SER_E_ARG; }
    CUDA_TRY(cudaFuncSetAttribute(ser_sweep_kernel_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)run->smem_big));
    int per_sm = 1, n_sm = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ser_sweep_kernel_big, run->big_threads, run->smem_big));
    CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
    run->big_slots = std::max(1, std::min(cfg->n_chains, per_sm * n_sm));
    kp.Cs = ((M + 1) + 31) / 32 * 32;
    const size_t sl = (size_t)run->big_slots;
    CUDA_TRY(POOL_ALLOC(&run->d_gV, sl * run->W * kp.Cs * sizeof(uint32_t)));
    CUDA_TRY(POOL_ALLOC(&run->d_gpre, sl * (run->W + 1) * kp.Cs * sizeof(uint16_t)));
    kp.gV = run->d_gV; kp.gpre = run->d_gpre;
  }
  kp.n_chains = cfg->n_chains;
  *out = run;
  return SER_OK;
}

extern "C" void ser_run_destroy(ser_run *run)
{
  if (!run) return;
  cudaSetDevice(run->cfg.device);
  void *bufs[] = {run->d_Xs, run->d_hard, run->d_ones, run->d_off, run->d_order, run->d_item_col, run->d_ab, run->d_rpi,
                  run->d_scal, run->d_tape, run->d_tape_off, run->d_samp_a, run->d_samp_b, run->d_samp_pi, run->d_samp_cdl,
                  run->d_scratch_i, run->d_bad, run->d_cd4, run->d_samp_cd_all, run->d_gV, run->d_gpre, run->d_bgrp};
  for (void *b : bufs) if (b) cudaFreeAsync(b, run->stream);
  cudaStreamSynchronize(run->stream);
  cudaEventDestroy(run->ev_start); cudaEventDestroy(run->ev_stop);
  cudaStreamDestroy(run->stream);
  free(run->h_hard);
  free(run);
}

extern "C" int ser_run_dims(const ser_run *run, int32_t *N, int32_t *M, int32_t *nh, int32_t *n_chains)
{
  if (!run) return SER_E_ARG;
  if (N) *N = run->N;
  if (M) *M = run->M;
  if (nh) *nh = run->nh;
  if (n_chains) *n_chains = run->cfg.n_chains;
  return SER_OK;
}
extern "C" const uint8_t *ser_run_hard_flags(const ser_run *run) { return run ? run->h_hard : nullptr; }
extern "C" int ser_run_is_manycd(const ser_run *run) { return run ? run->cfg.manycd : 0; }

extern "C" int ser_run_set_tapes(ser_run *run, const double *flat, const uint64_t *offsets)
{
  if (!run || !flat || !offsets) { ser_set_error("ser_run_set_tapes: null argument"); return SER_E_ARG; }
  if (run->cfg.mode != SER_MODE_REPLAY) { ser_set_error("ser_run_set_tapes: run is not in replay mode"); return SER_E_STATE; }
  if (set_device(run)) return SER_E_CUDA;
  const size_t nc = (size_t)run->cfg.n_chains, total = (size_t)offsets[nc];
  if (run->d_tape) cudaFreeAsync(run->d_tape, run->stream);
  if (run->d_tape_off) cudaFreeAsync(run->d_tape_off, run->stream);
  run->d_tape = nullptr; run->d_tape_off = nullptr;
  CUDA_TRY(POOL_ALLOC(&run->d_tape, (total ? total : 1) * sizeof(double)));
  CUDA_TRY(POOL_ALLOC(&run->d_tape_off, (nc + 1) * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemcpyAsync(run->d_tape, flat, total * sizeof(double), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(run->d_tape_off, offsets, (nc + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  run->kp.tape = run->d_tape; run->kp.tape_off = run->d_tape_off;
  run->have_tapes = 1;
  return SER_OK;
}

extern "C" int ser_run_init(ser_run *run)
{
  if (!run) return SER_E_ARG;
  if (run->cfg.mode == SER_MODE_REPLAY && !run->have_tapes) { ser_set_error("ser_run_init: replay mode needs ser_run_set_tapes first"); return SER_E_TAPE; }
  if (set_device(run)) return SER_E_CUDA;
  mark_launch(run);
  ser_init_kernel<<<run->cfg.n_chains, run->Caux, run->smem_init, run->stream>>>(run->kp);
  CUDA_TRY(cudaGetLastError());
  run->initialized = 1;
  return SER_OK;
}

extern "C" int ser_run_advance(ser_run *run, int32_t n_calls, int32_t sampling)
{
  if (!run) return SER_E_ARG;
  if (!run->initialized) { ser_set_error("ser_run_advance: call ser_run_init first"); return SER_E_STATE; }
  if (n_calls <= 0) return SER_OK;
  if (set_device(run)) return SER_E_CUDA;
  KParams kp = run->kp;
  kp.n_calls = n_calls; kp.sampling = sampling;
  mark_launch(run);
  if (run->cfg.manycd && run->variant_many) ser_sweep_kernel<384, 2, true><<<run->cfg.n_chains, run->C, run->smem_many, run->stream>>>(kp);
  else if (run->cfg.manycd) ser_sweep_kernel<1024, 1, true><<<run->cfg.n_chains, run->C, run->smem_many, run->stream>>>(kp);
  else if (run->big) ser_sweep_kernel_big<<<run->big_slots, run->big_threads, run->smem_big, run->stream>>>(kp);
  else if (run->variant == 1) ser_sweep_kernel<384, 2, false><<<run->cfg.n_chains, run->C, run->smem_sweep, run->stream>>>(kp);
  else ser_sweep_kernel<1024, 1, false><<<run->cfg.n_chains, run->C, run->smem_sweep, run->stream>>>(kp);
  CUDA_TRY(cudaGetLastError());
  return SER_OK;
}

extern "C" int ser_run_sync(ser_run *run)
{
  if (!run) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  return SER_OK;
}

extern "C" int ser_run_elapsed_ms(ser_run *run, double *ms, int32_t reset)
{
  if (!run || !ms) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  if (run->timing_open) {
    CUDA_TRY(cudaEventRecord(run->ev_stop, run->stream));
    CUDA_TRY(cudaEventSynchronize(run->ev_stop));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, run->ev_start, run->ev_stop));
    run->elapsed_ms += t;
    run->timing_open = 0;
  }
  *ms = run->elapsed_ms;
  if (reset) run->elapsed_ms = 0.0;
  return SER_OK;
}

extern "C" int ser_run_kernel_launches(const ser_run *run, int64_t *n)
{
  if (!run || !n) return SER_E_ARG;
  *n = run->launches;
  return SER_OK;
}

static int chain_ok(ser_run *run, int chain)
{
  if (!run) { ser_set_error("null run"); return 0; }
  if (chain < 0 || chain >= run->cfg.n_chains) { ser_set_error("chain %d out of range [0,%d)", chain, run->cfg.n_chains); return 0; }
  return 1;
}

extern "C" int ser_run_get_state(ser_run *run, int32_t chain, int32_t *a, int32_t *b, int32_t *pi, int32_t *rpi, int32_t *t0,
                                 int32_t *f0, int32_t *t1, int32_t *f1, int32_t tot[4], double cdl[3], int64_t *tape_slots)
{
  if (!chain_ok(run, chain)) return SER_E_ARG;
  if (!run->initialized) { ser_set_error("ser_run_get_state: run not initialised"); return SER_E_STATE; }
  if (set_device(run)) return SER_E_CUDA;
  const int N = run->N, M = run->M;
  int *d = run->d_scratch_i;
  int *d_a = d, *d_b = d + M, *d_pi = d + 2 * M, *d_rpi = d + 2 * M + N, *d_cnt = d + 2 * M + 2 * N;
  mark_launch(run);
  ser_export_kernel<<<1, run->Caux, run->smem_small, run->stream>>>(run->kp, chain, d_a, d_b, d_pi, d_rpi, d_cnt);
  CUDA_TRY(cudaGetLastError());
  std::vector<int> h(2 * N + 6 * M);
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(int), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal + chain, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  if (a) memcpy(a, h.data(), M * 4);
  if (b) memcpy(b, h.data() + M, M * 4);
  if (pi) memcpy(pi, h.data() + 2 * M, N * 4);
  if (rpi) memcpy(rpi, h.data() + 2 * M + N, N * 4);
  const int *cnt = h.data() + 2 * M + 2 * N;
  if (t0) memcpy(t0, cnt, M * 4);
  if (f0) memcpy(f0, cnt + M, M * 4);
  if (t1) memcpy(t1, cnt + 2 * M, M * 4);
  if (f1) memcpy(f1, cnt + 3 * M, M * 4);
  if (tot) { tot[0] = sc.t0a; tot[1] = sc.f0a; tot[2] = sc.t1a; tot[3] = sc.f1a; }
  if (cdl) { cdl[0] = sc.c; cdl[1] = sc.d; cdl[2] = sc.loglik; }
  if (tape_slots) *tape_slots = sc.cursor;
  if (sc.flags & 1) { ser_set_error("chain %d: replay tape exhausted", chain); return SER_E_TAPE; }
  return SER_OK;
}

extern "C" int ser_run_get_counters(ser_run *run, int32_t chain, int64_t out[8])
{
  if (!chain_ok(run, chain) || !out) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal + chain, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  for (int i = 0; i < 8; i++) out[i] = sc.counters[i];
  return SER_OK;
}

/* flags of one chain after ser_run_check: 1 tape exhausted, 2 a/b range, 4 permutation, 8 hard-site order, 16 totals / loglik */
extern "C" int ser_run_get_flags(ser_run *run, int32_t chain, int32_t *flags)
{
  if (!chain_ok(run, chain) || !flags) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal + chain, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  *flags = sc.flags;
  return SER_OK;
}

extern "C" int ser_run_check(ser_run *run, int32_t *n_bad)
{
  if (!run || !n_bad) return SER_E_ARG;
  if (!run->initialized) { ser_set_error("ser_run_check: run not initialised"); return SER_E_STATE; }
  if (set_device(run)) return SER_E_CUDA;
  CUDA_TRY(cudaMemsetAsync(run->d_bad, 0, sizeof(int), run->stream));
  mark_launch(run);
  ser_check_kernel<<<run->cfg.n_chains, run->Caux, run->smem_small, run->stream>>>(run->kp, run->d_bad);
  CUDA_TRY(cudaGetLastError());
  int bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, run->d_bad, sizeof(int), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  *n_bad = bad;
  if (bad) { ser_set_error("ser_run_check: %d inconsistent chain(s)", bad); return SER_E_CHECK; }
  return SER_OK;
}

extern "C" int ser_run_chain_sums(ser_run *run, int32_t chain, double sums[3], int32_t *n_samples)
{
  if (!chain_ok(run, chain)) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal + chain, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  if (sums) { sums[0] = sc.sum_negll; sums[1] = sc.sum_ec; sums[2] = sc.sum_ed; }
  if (n_samples) *n_samples = sc.n_samples;
  return SER_OK;
}

extern "C" int ser_run_chain_stats(ser_run *run, double *e_negloglik, double *e_c, double *e_d, int32_t *n_samples)
{
  if (!run) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  const int nc = run->cfg.n_chains;
  std::vector<ChainScalars> sc(nc);
  CUDA_TRY(cudaMemcpyAsync(sc.data(), run->d_scal, (size_t)nc * sizeof(ChainScalars), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  for (int i = 0; i < nc; i++) {
    const double n = sc[i].n_samples > 0 ? (double)sc[i].n_samples : 1.0;
    if (e_negloglik) e_negloglik[i] = sc[i].sum_negll / n;
    if (e_c) e_c[i] = sc[i].sum_ec / n;
    if (e_d) e_d[i] = sc[i].sum_ed / n;
  }
  if (n_samples) *n_samples = sc[0].n_samples;
  return SER_OK;
}

extern "C" int ser_run_chain_stats_device(ser_run *run, double *d_e_negloglik)
{
  if (!run || !d_e_negloglik) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  const int nc = run->cfg.n_chains;
  mark_launch(run);
  ser_stats_kernel<<<(nc + 255) / 256, 256, 0, run->stream>>>(run->d_scal, nc, d_e_negloglik);
  CUDA_TRY(cudaGetLastError());
  return SER_OK;
}

extern "C" int ser_run_fetch_samples(ser_run *run, int32_t chain, int32_t *a, int32_t *b, int32_t *pi, double *c, double *d,
                                     double *loglik, int32_t *n)
{
  if (!chain_ok(run, chain)) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal + chain, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  const int ns = sc.n_samples < run->cfg.max_samples ? sc.n_samples : run->cfg.max_samples;
  if (n) *n = ns;
  const int N = run->N, M = run->M;
  const size_t row0 = (size_t)chain * run->cfg.max_samples;
  if ((a || b || c || d || loglik) && run->cfg.store < SER_STORE_FULL) { ser_set_error("ser_run_fetch_samples: run was created without SER_STORE_FULL"); return SER_E_STATE; }
  if (pi && run->cfg.store < SER_STORE_PI) { ser_set_error("ser_run_fetch_samples: run was created without a sample store"); return SER_E_STATE; }
  std::vector<uint16_t> tmp;
  auto fetch16 = [&](const uint16_t *src, int width, int32_t *dst) -> int {
    tmp.resize((size_t)ns * width);
    if (cudaMemcpyAsync(tmp.data(), src + row0 * width, tmp.size() * 2, cudaMemcpyDeviceToHost, run->stream) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(run->stream) != cudaSuccess) return 1;
    for (size_t i = 0; i < tmp.size(); i++) dst[i] = tmp[i];
    return 0;
  };
  if (ns > 0) {
    if (a && fetch16(run->d_samp_a, M, a)) { ser_set_error("fetch a failed"); return SER_E_CUDA; }
    if (b && fetch16(run->d_samp_b, M, b)) { ser_set_error("fetch b failed"); return SER_E_CUDA; }
    if (pi && fetch16(run->d_samp_pi, N, pi)) { ser_set_error("fetch pi failed"); return SER_E_CUDA; }
    if (c || d || loglik) {
      std::vector<double> cdl((size_t)ns * 3);
      CUDA_TRY(cudaMemcpyAsync(cdl.data(), run->d_samp_cdl + row0 * 3, cdl.size() * 8, cudaMemcpyDeviceToHost, run->stream));
      CUDA_TRY(cudaStreamSynchronize(run->stream));
      for (int s = 0; s < ns; s++) {
        if (c) c[s] = cdl[3 * s];
        if (d) d[s] = cdl[3 * s + 1];
        if (loglik) loglik[s] = cdl[3 * s + 2];
      }
    }
  }
  return SER_OK;
}

/* manycd runs: per-taxon c, d of one chain (file order) */
extern "C" int ser_run_get_cd(ser_run *run, int32_t chain, double *c, double *d)
{
  if (!chain_ok(run, chain)) return SER_E_ARG;
  if (!run->cfg.manycd) { ser_set_error("ser_run_get_cd: run was created with manycd = 0"); return SER_E_STATE; }
  if (!run->initialized) { ser_set_error("ser_run_get_cd: run not initialised"); return SER_E_STATE; }
  if (set_device(run)) return SER_E_CUDA;
  const int M = run->M, Mpad = run->kp.Mpad;
  std::vector<double> h((size_t)4 * Mpad);
  std::vector<uint16_t> order(M);
  CUDA_TRY(cudaMemcpyAsync(h.data(), run->d_cd4 + (size_t)chain * 4 * Mpad, h.size() * 8, cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaMemcpyAsync(order.data(), run->d_order, M * sizeof(uint16_t), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  for (int col = 0; col < M; col++) {
    if (c) c[order[col]] = h[col];
    if (d) d[order[col]] = h[2 * Mpad + col];
  }
  return SER_OK;
}

/* manycd runs with SER_STORE_FULL: per-taxon c, d of every stored sample, [n][M] each */
extern "C" int ser_run_fetch_cd_samples(ser_run *run, int32_t chain, double *c, double *d, int32_t *n)
{
  if (!chain_ok(run, chain)) return SER_E_ARG;
  if (!run->cfg.manycd || !run->d_samp_cd_all) { ser_set_error("ser_run_fetch_cd_samples: needs manycd = 1 and SER_STORE_FULL"); return SER_E_STATE; }
  if (set_device(run)) return SER_E_CUDA;
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal + chain, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  const int ns = sc.n_samples < run->cfg.max_samples ? sc.n_samples : run->cfg.max_samples, M = run->M;
  if (n) *n = ns;
  if (ns > 0 && (c || d)) {
    std::vector<double> h((size_t)ns * 2 * M);
    CUDA_TRY(cudaMemcpyAsync(h.data(), run->d_samp_cd_all + (size_t)chain * run->cfg.max_samples * 2 * M, h.size() * 8,
                             cudaMemcpyDeviceToHost, run->stream));
    CUDA_TRY(cudaStreamSynchronize(run->stream));
    for (int s = 0; s < ns; s++) {
      if (c) memcpy(c + (size_t)s * M, h.data() + ((size_t)s * 2 + 0) * M, M * 8);
      if (d) memcpy(d + (size_t)s * M, h.data() + ((size_t)s * 2 + 1) * M, M * 8);
    }
  }
  return SER_OK;
}

extern "C" int ser_select_chains_device(const double *d_e, int32_t n, int32_t k, int32_t *d_chosen, double *d_info, int32_t device,
                                        void *stream)
{
  if (!d_e || !d_chosen || !d_info || n < 1 || k < 1) { ser_set_error("ser_select_chains_device: bad argument"); return SER_E_ARG; }
  CUDA_TRY(cudaSetDevice(device));
  ser_select_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_e, n, k, d_chosen, d_info);
  CUDA_TRY(cudaGetLastError());
  return SER_OK;
}

extern "C" int ser_run_po_counts_device(ser_run *run, const int32_t *d_chosen, int32_t k, int32_t *d_counts)
{
  if (!run || !d_chosen || !d_counts || k < 1) return SER_E_ARG;
  if (run->cfg.store < SER_STORE_PI) { ser_set_error("ser_run_po_counts: run has no pi sample store"); return SER_E_STATE; }
  if (set_device(run)) return SER_E_CUDA;
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  const int ns = sc.n_samples < run->cfg.max_samples ? sc.n_samples : run->cfg.max_samples;
  dim3 blk(32, 8), grd((run->N + 31) / 32, (run->N + 7) / 8, k);
  mark_launch(run);
  ser_po_kernel<<<grd, blk, 0, run->stream>>>(run->d_samp_pi, run->N, run->cfg.max_samples, ns, d_chosen, run->cfg.chain_offset,
                                              run->cfg.n_chains, d_counts);
  CUDA_TRY(cudaGetLastError());
  return SER_OK;
}

extern "C" int ser_run_po_counts(ser_run *run, const int32_t *chosen, int32_t k, int32_t *counts)
{
  if (!run || !chosen || !counts || k < 1) return SER_E_ARG;
  if (set_device(run)) return SER_E_CUDA;
  int *d_ch = nullptr, *d_cnt = nullptr;
  const size_t nn = (size_t)k * run->N * run->N;
  CUDA_TRY(POOL_ALLOC(&d_ch, k * sizeof(int)));
  CUDA_TRY(POOL_ALLOC(&d_cnt, nn * sizeof(int)));
  CUDA_TRY(cudaMemcpyAsync(d_ch, chosen, k * sizeof(int), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemsetAsync(d_cnt, 0, nn * sizeof(int), run->stream));
  int rc = ser_run_po_counts_device(run, d_ch, k, d_cnt);
  if (rc == SER_OK) {
    if (cudaMemcpyAsync(counts, d_cnt, nn * sizeof(int), cudaMemcpyDeviceToHost, run->stream) != cudaSuccess ||
        cudaStreamSynchronize(run->stream) != cudaSuccess) { ser_set_error("ser_run_po_counts: copy back failed"); rc = SER_E_CUDA; }
  }
  cudaFreeAsync(d_ch, run->stream); cudaFreeAsync(d_cnt, run->stream);
  return rc;
}

extern "C" int ser_run_posterior_sums(ser_run *run, const int32_t *chosen, int32_t k, int64_t *corr_num, int32_t *pi_sum,
                                      int32_t *a_sum, int32_t *b_sum, int32_t *n_samples)
{
  if (!run || !chosen || !corr_num || !pi_sum || k < 1) { ser_set_error("ser_run_posterior_sums: bad argument"); return SER_E_ARG; }
  if (run->cfg.store < SER_STORE_PI) { ser_set_error("ser_run_posterior_sums: run has no pi sample store"); return SER_E_STATE; }
  if ((a_sum || b_sum) && run->cfg.store < SER_STORE_FULL) { ser_set_error("ser_run_posterior_sums: a/b sums need SER_STORE_FULL"); return SER_E_STATE; }
  if (set_device(run)) return SER_E_CUDA;
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  const int ns = sc.n_samples < run->cfg.max_samples ? sc.n_samples : run->cfg.max_samples, N = run->N, M = run->M;
  int *d_ch = nullptr, *d_pi = nullptr, *d_a = nullptr, *d_b = nullptr;
  long long *d_corr = nullptr;
  CUDA_TRY(POOL_ALLOC(&d_ch, k * sizeof(int)));
  CUDA_TRY(POOL_ALLOC(&d_corr, k * sizeof(long long)));
  CUDA_TRY(POOL_ALLOC(&d_pi, (size_t)k * N * sizeof(int)));
  if (a_sum) { CUDA_TRY(POOL_ALLOC(&d_a, (size_t)k * M * sizeof(int))); CUDA_TRY(POOL_ALLOC(&d_b, (size_t)k * M * sizeof(int))); }
  CUDA_TRY(cudaMemcpyAsync(d_ch, chosen, k * sizeof(int), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(d_corr, corr_num, k * sizeof(long long), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(d_pi, pi_sum, (size_t)k * N * sizeof(int), cudaMemcpyHostToDevice, run->stream));
  if (a_sum) {
    CUDA_TRY(cudaMemcpyAsync(d_a, a_sum, (size_t)k * M * sizeof(int), cudaMemcpyHostToDevice, run->stream));
    if (b_sum) CUDA_TRY(cudaMemcpyAsync(d_b, b_sum, (size_t)k * M * sizeof(int), cudaMemcpyHostToDevice, run->stream));
  }
  mark_launch(run);
  ser_posterior_kernel<<<k, 256, 0, run->stream>>>(run->d_samp_pi, run->d_samp_a, run->d_samp_b, N, M, run->cfg.max_samples, ns, d_ch,
                                                   run->cfg.chain_offset, run->cfg.n_chains, d_corr, d_pi, d_a, b_sum ? d_b : nullptr);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(corr_num, d_corr, k * sizeof(long long), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaMemcpyAsync(pi_sum, d_pi, (size_t)k * N * sizeof(int), cudaMemcpyDeviceToHost, run->stream));
  if (a_sum) {
    CUDA_TRY(cudaMemcpyAsync(a_sum, d_a, (size_t)k * M * sizeof(int), cudaMemcpyDeviceToHost, run->stream));
    if (b_sum) CUDA_TRY(cudaMemcpyAsync(b_sum, d_b, (size_t)k * M * sizeof(int), cudaMemcpyDeviceToHost, run->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  cudaFreeAsync(d_ch, run->stream); cudaFreeAsync(d_corr, run->stream); cudaFreeAsync(d_pi, run->stream);
  if (d_a) cudaFreeAsync(d_a, run->stream);
  if (d_b) cudaFreeAsync(d_b, run->stream);
  if (n_samples) *n_samples = ns;
  return SER_OK;
}

extern "C" int ser_run_alive_counts(ser_run *run, const int32_t *chosen, int32_t k, int32_t *alive, int32_t *n_samples)
{
  if (!run || !chosen || !alive || k < 1) { ser_set_error("ser_run_alive_counts: bad argument"); return SER_E_ARG; }
  if (run->cfg.store < SER_STORE_FULL) { ser_set_error("ser_run_alive_counts: needs SER_STORE_FULL (a, b samples)"); return SER_E_STATE; }
  if (set_device(run)) return SER_E_CUDA;
  ChainScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, run->d_scal, sizeof(sc), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  const int ns = sc.n_samples < run->cfg.max_samples ? sc.n_samples : run->cfg.max_samples, N = run->N, M = run->M;
  const size_t cells = (size_t)k * N * M;
  int *d_ch = nullptr, *d_alive = nullptr;
  CUDA_TRY(POOL_ALLOC(&d_ch, k * sizeof(int)));
  CUDA_TRY(POOL_ALLOC(&d_alive, cells * sizeof(int)));
  CUDA_TRY(cudaMemcpyAsync(d_ch, chosen, k * sizeof(int), cudaMemcpyHostToDevice, run->stream));
  CUDA_TRY(cudaMemcpyAsync(d_alive, alive, cells * sizeof(int), cudaMemcpyHostToDevice, run->stream));
  mark_launch(run);
  ser_alive_kernel<<<dim3(k, (M + 127) / 128), 128, 0, run->stream>>>(run->d_samp_a, run->d_samp_b, N, M, run->cfg.max_samples, ns, d_ch,
                                                                     run->cfg.chain_offset, run->cfg.n_chains, d_alive);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(alive, d_alive, cells * sizeof(int), cudaMemcpyDeviceToHost, run->stream));
  CUDA_TRY(cudaStreamSynchronize(run->stream));
  cudaFreeAsync(d_ch, run->stream); cudaFreeAsync(d_alive, run->stream);
  if (n_samples) *n_samples = ns;
  return SER_OK;
}

extern "C" int ser_microbench(int32_t device, double out[3])
{
  if (!out) return SER_E_ARG;
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 2, threads = 1024, iters = 20000;
  void *buf = nullptr;
  CUDA_TRY(cudaMalloc(&buf, (size_t)blocks * threads * 8));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  float ms = 0.f;
  for (int rep = 0; rep < 2; rep++) { /* first pass warms up */
    CUDA_TRY(cudaEventRecord(e0));
    mb_fp64_kernel<<<blocks, threads>>>((double *)buf, iters);
    CUDA_TRY(cudaEventRecord(e1)); CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  }
  out[0] = (double)blocks * threads * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
  for (int rep = 0; rep < 2; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    mb_lds_kernel<<<blocks, threads>>>((unsigned *)buf, iters / 4);
    CUDA_TRY(cudaEventRecord(e1)); CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  }
  out[1] = (double)blocks * threads * (iters / 4) * 8.0 * 16.0 / (ms * 1e-3) / 1e9;
  for (int rep = 0; rep < 2; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    mb_popc_kernel<<<blocks, threads>>>((unsigned *)buf, iters);
    CUDA_TRY(cudaEventRecord(e1)); CUDA_TRY(cudaEventSynchronize(e1));
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  }
  out[2] = (double)blocks * threads * iters * 8.0 / (ms * 1e-3) / 1e9;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  return SER_OK;
}
