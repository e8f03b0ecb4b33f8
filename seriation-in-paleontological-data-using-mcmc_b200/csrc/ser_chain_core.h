/*
 * ser_chain_core.h -- per-taxon building blocks of the B200 sweep.
 *
 * One chain = one CTA; one taxon = one thread.  Each thread owns one *column*
 * of the chain's bit matrix V: V[w][col] (word-major, stride C words) holds the
 * taxon's occurrences in POSITION order, bit p = X[rpi[p]][m].  Column M is
 * the hard-site mask in position order.  Next to V lives the prefix table
 * pre[w][col] = number of ones in words < w (w = 0..W), so every range
 * popcount is two table reads + two POPCs, loop-free.  Everything the
 * reference evaluates cell by cell (mcmc.c:828-898, :1127-1682) becomes range
 * popcounts on that column.  The functions below are the per-thread pieces; the block-level
 * choreography (reductions, draws, accept) lives in ser_kernels.cu and, for
 * CPU-side testing of the same logic, in tests/emul/chain_emul.cpp.
 *
 * Reference (file:line under /root/reference/C_Implementation):
 *   ser_step_a/_b, ser_item_eval/_weight, ser_run_sum/_pick, ser_step_lmax/_pick
 *                       mcmc_auxa :828-898 + mcmc_logtop :711-748 + mcmc_randompick :901-915
 *   ser_pi1_delta       mcmc_samplepi1 :1171-1256
 *   ser_pi2_delta       mcmc_samplepi2 :1363-1436 (with mcmc_ininterval :1097-1124)
 *   ser_pi3_delta       mcmc_samplepi3 :1564-1631
 *   ser_mirror_ab       :1446-1465 / :1641-1660
 *   ser_col_*           the rpi rotations / reversals :1277-1297, :1469-1474, :1664-1670
 */
#ifndef SER_CHAIN_CORE_H
#define SER_CHAIN_CORE_H

#include "ser_detmath.h"

#ifndef __CUDACC__
#include <math.h>
#endif

#define SER_MAXW 64 /* words per column: N <= 2048 */

/* Debug builds (NVCC_EXTRA=-DSER_DEBUG, see profiles/r02/README.md): every index into shared memory / a column that is
 * computed from chain state is range-checked; a violation prints its location and traps the kernel.  Compiled out otherwise. */
#if defined(SER_DEBUG) && defined(__CUDA_ARCH__)
#define SER_CHECK(cond)                                                                                              \
  do {                                                                                                               \
    if (!(cond)) {                                                                                                   \
      printf("SER_DEBUG: (%s) failed at %s:%d, block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
      __trap();                                                                                                      \
    }                                                                                                                \
  } while (0)
#else
#define SER_CHECK(cond) do { } while (0)
#endif

/* ------------------------------------------------------------------ bit intrinsics */
#if defined(__CUDA_ARCH__)
#define SER_POPC(x) __popc((x))
#define SER_BREV(x) __brev((x))
#define SER_FFS(x) __ffs((int)(x))
SER_HD uint32_t ser_funnel_r(uint32_t lo, uint32_t hi, int sh) { return __funnelshift_r(lo, hi, sh); }
SER_HD double ser_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
SER_HD double ser_fmax(double a, double b) { return a > b ? a : b; } /* no NaNs here; fmax() costs twice the instructions */
#else
#define SER_POPC(x) __builtin_popcount((x))
#define SER_FFS(x) __builtin_ffs((int)(x))
SER_HD uint32_t SER_BREV(uint32_t v)
{
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
  return (v >> 16) | (v << 16);
}
SER_HD uint32_t ser_funnel_r(uint32_t lo, uint32_t hi, int sh)
{
  sh &= 31;
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
SER_HD double ser_fma(double a, double b, double c) { return fma(a, b, c); }
SER_HD double ser_fmax(double a, double b) { return a > b ? a : b; }
#endif

/* bits [0,n), n in [0,31] -- the hot-path variant (n = p & 31) */
SER_HD uint32_t ser_mask_lo(int n) { return (1u << n) - 1u; }

/* bits [0,n), n in [0,32] */
SER_HD uint32_t ser_mask_lt(int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

/* bits of word w (positions 32w .. 32w+31) that fall into [lo, hi) */
SER_HD uint32_t ser_range_mask(int w, int lo, int hi)
{
  int l = lo - 32 * w, h = hi - 32 * w;
  if (l < 0) l = 0;
  if (h > 32) h = 32;
  if (h <= l) return 0u;
  return ser_mask_lt(h) & ~ser_mask_lt(l);
}

/* ------------------------------------------------------------------ columns */
/* A column is addressed as col[w * C] (word w of the column, stride C words). */
SER_HD int ser_col_bit(const uint32_t *col, int C, int p) { return (col[(p >> 5) * C] >> (p & 31)) & 1u; }

/* The prefix table of a column: G = 0 (columns in shared memory) one entry per word, pre[w * C] = ones in words < w, w = 0..W.
 * G > 0 (the large-shape kernel's columns in global memory, where the table's bytes compete with the columns for the L2): one
 * entry per 2^G words, pre[k * C] = ones in words < k 2^G, k = 0..W >> G; the words between an entry and w are counted on the fly. */
template <int G = 0>
SER_HD int ser_pre_at(const uint32_t *col, const uint16_t *pre, int C, int w)
{
  int r = (int)pre[(w >> G) * C];
  if (G > 0) for (int x = (w >> G) << G; x < w; x++) r += SER_POPC(col[x * C]);
  return r;
}
template <int G = 0>
SER_HD void ser_pre_put(uint16_t *pre, int C, int w, int ones_below)
{
  if (G == 0 || (w & ((1 << G) - 1)) == 0) pre[(w >> G) * C] = (uint16_t)ones_below;
}

/* ones in bits [0, p), p in [0, N]: prefix table + one POPC */
template <int G = 0>
SER_HD int ser_rank1(const uint32_t *col, const uint16_t *pre, int C, int p)
{
  SER_CHECK(p >= 0 && p <= 32 * SER_MAXW);
  const int w = p >> 5;
  return ser_pre_at<G>(col, pre, C, w) + SER_POPC(col[w * C] & ser_mask_lo(p & 31));
}

/* popcount of bits [lo, hi) */
template <int G = 0>
SER_HD int ser_col_popc(const uint32_t *col, const uint16_t *pre, int C, int lo, int hi)
{
  return hi <= lo ? 0 : ser_rank1<G>(col, pre, C, hi) - ser_rank1<G>(col, pre, C, lo);
}

/* (re)build the table from the words */
template <int G = 0>
SER_HD void ser_col_build_pre(const uint32_t *col, uint16_t *pre, int C, int W)
{
  int acc = 0;
  pre[0] = 0;
  for (int w = 0; w < W; w++) { acc += SER_POPC(col[w * C]); ser_pre_put<G>(pre, C, w + 1, acc); }
}

/* after a permutation of bits inside words w0..w1 (the multiset of bits there is unchanged)
 * only the counts of the boundaries w0+1..w1 can differ */
template <int G = 0>
SER_HD void ser_col_fix_pre(const uint32_t *col, uint16_t *pre, int C, int w0, int w1)
{
  int acc = ser_pre_at<G>(col, pre, C, w0);
  for (int w = w0; w < w1; w++) { acc += SER_POPC(col[w * C]); ser_pre_put<G>(pre, C, w + 1, acc); }
}

/* The three column moves below take an optional prefix table: with `pre` the counts pre[w0+1..w1] are
 * rewritten from the new words as they are produced (what ser_col_fix_pre would do in a second pass
 * that re-reads the column). */
/* in-place: reverse bits [i, j] (new[p] = old[i+j-p]) */
template <int G = 0>
SER_HD void ser_col_reverse(uint32_t *col, int C, int W, int i, int j, uint16_t *pre = 0)
{
  uint32_t old[SER_MAXW];
  const int w0 = i >> 5, w1 = j >> 5;
  SER_CHECK(0 <= i && i <= j && w1 < W);
  if (i + 1 == j && w0 == w1) { /* adjacent swap inside one word: the common (swap) case */
    uint32_t v = col[w0 * C];
    const uint32_t x = ((v >> (i & 31)) ^ (v >> (j & 31))) & 1u;
    col[w0 * C] = v ^ ((x << (i & 31)) | (x << (j & 31)));
    return;
  }
  for (int w = w0; w <= w1; w++) old[w] = col[w * C];
  const int s = i + j;
  int acc = pre ? ser_pre_at<G>(col, pre, C, w0) : 0;
  for (int wn = w0; wn <= w1; wn++) {
    /* new bit p = old[s - p]: 32 old bits ending at s - 32wn, reversed */
    const int q = s - 32 * wn - 31;
    const int qw = q >> 5;
    const uint32_t lo = (qw >= w0 && qw <= w1) ? old[qw] : 0u;
    const uint32_t hi = (qw + 1 >= w0 && qw + 1 <= w1) ? old[qw + 1] : 0u;
    const uint32_t bits = SER_BREV(ser_funnel_r(lo, hi, q & 31));
    const uint32_t m = ser_range_mask(wn, i, j + 1);
    const uint32_t nw = (old[wn] & ~m) | (bits & m);
    col[wn * C] = nw;
    if (pre && wn < w1) { acc += SER_POPC(nw); ser_pre_put<G>(pre, C, wn + 1, acc); }
  }
}

/* in-place: move bit i to position j, shifting the bits in between by one (pi1).  No copy of the column: a word
 * only needs its old neighbour on the side the bits come from, and the words are rewritten in the order that
 * leaves that neighbour untouched (upwards for i < j, downwards for i > j). */
template <int G = 0>
SER_HD void ser_col_rotate(uint32_t *col, int C, int W, int i, int j, uint16_t *pre = 0)
{
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  const int w0 = lo >> 5, w1 = hi >> 5;
  SER_CHECK(lo >= 0 && w1 < W);
  const uint32_t moved = (col[(i >> 5) * C] >> (i & 31)) & 1u;
  if (i < j) { /* new[p] = old[p+1] for p in [i, j-1] */
    int acc = pre ? ser_pre_at<G>(col, pre, C, w0) : 0;
    uint32_t cur = col[w0 * C];
    for (int wn = w0; wn <= w1; wn++) {
      const uint32_t up = (wn + 1 <= w1) ? col[(wn + 1) * C] : 0u;
      const uint32_t m = ser_range_mask(wn, i, j);
      uint32_t nw = (cur & ~m) | (((cur >> 1) | (up << 31)) & m);
      if (wn == w1) nw = (nw & ~(1u << (j & 31))) | (moved << (j & 31));
      col[wn * C] = nw;
      if (pre && wn < w1) { acc += SER_POPC(nw); ser_pre_put<G>(pre, C, wn + 1, acc); }
      cur = up;
    }
  } else { /* new[p] = old[p-1] for p in [j+1, i] */
    uint32_t cur = col[w1 * C];
    for (int wn = w1; wn >= w0; wn--) {
      const uint32_t dn = (wn - 1 >= w0) ? col[(wn - 1) * C] : 0u;
      const uint32_t m = ser_range_mask(wn, j + 1, i + 1);
      uint32_t nw = (cur & ~m) | (((cur << 1) | (dn >> 31)) & m);
      if (wn == w0) nw = (nw & ~(1u << (j & 31))) | (moved << (j & 31));
      col[wn * C] = nw;
      cur = dn;
    }
    if (pre) ser_col_fix_pre<G>(col, pre, C, w0, w1);
  }
}

/* in-place: new[p] = old[perm[p]] for p in [i, j] (pi3; perm is an involution on the window) */
template <int G = 0>
SER_HD void ser_col_permute(uint32_t *col, int C, int W, int i, int j, const uint16_t *perm, uint16_t *pre = 0)
{
  uint32_t old[SER_MAXW];
  const int w0 = i >> 5, w1 = j >> 5;
  SER_CHECK(0 <= i && i <= j && w1 < W);
  for (int w = w0; w <= w1; w++) old[w] = col[w * C];
  int acc = pre ? ser_pre_at<G>(col, pre, C, w0) : 0;
  for (int wn = w0; wn <= w1; wn++) {
    uint32_t nw = old[wn];
    const int p0 = (32 * wn > i) ? 32 * wn : i, p1 = (32 * wn + 31 < j) ? 32 * wn + 31 : j;
    for (int p = p0; p <= p1; p++) {
      const int src = perm[p];
      SER_CHECK(src >= i && src <= j);
      const uint32_t bit = (old[src >> 5] >> (src & 31)) & 1u;
      nw = (nw & ~(1u << (p & 31))) | (bit << (p & 31));
    }
    col[wn * C] = nw;
    if (pre && wn < w1) { acc += SER_POPC(nw); ser_pre_put<G>(pre, C, wn + 1, acc); }
  }
}

/* ------------------------------------------------------------------ hard sites */
/* hcol = column M (hard mask in position order) and its prefix table */
struct SerHard {
  const uint32_t *hcol;
  const uint16_t *hpre; /* hpre[w * C] = #hard positions in words < w, w = 0..W */
  const uint16_t *hp;   /* the nh hard positions, ascending */
  int C, W, N, nh;
  /* optional lookup tables (the one-thread-per-column kernel keeps them in shared memory; rebuilt whenever a hard
   * site moves): rank_tab[p] = hard positions < p, p = 0..N; nonhard_tab[r] = position of the r-th non-hard site */
  const uint16_t *rank_tab, *nonhard_tab;
};

/* rebuild the sorted list of hard positions from the mask (owner thread, after a move) */
SER_HD void ser_hard_list(const uint32_t *hcol, int C, int W, uint16_t *hp)
{
  int k = 0;
  for (int w = 0; w < W; w++) {
    uint32_t v = hcol[w * C];
    while (v) { hp[k++] = (uint16_t)(32 * w + SER_FFS(v) - 1); v &= v - 1u; }
  }
}

/* number of hard positions < p, p in [0, N] */
SER_HD int ser_hard_rank(const SerHard &h, int p) { SER_CHECK(p >= 0 && p <= h.N); return h.rank_tab ? (int)h.rank_tab[p] : ser_rank1(h.hcol, h.hpre, h.C, p); }
SER_HD int ser_is_hard(const SerHard &h, int p) { return (h.hcol[(p >> 5) * h.C] >> (p & 31)) & 1u; }
/* number of hard positions in [lo, hi] */
SER_HD int ser_hard_count(const SerHard &h, int lo, int hi) { return ser_hard_rank(h, hi + 1) - ser_hard_rank(h, lo); }
/* position of the r-th (from 0) non-hard position; r < N - nh.  hp[k] - k = number of non-hard
 * positions before hard site k (non-decreasing), so the answer is r + #{k : hp[k] - k <= r}:
 * an upper-bound search over the sorted hard list. */
SER_HD int ser_select_nonhard(const SerHard &h, int r)
{
  SER_CHECK(r >= 0 && r < h.N - h.nh);
  if (h.nonhard_tab) return (int)h.nonhard_tab[r];
  int lo = 0, hi = h.nh; /* first k with hp[k] - k > r */
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int)h.hp[mid] - mid <= r) lo = mid + 1; else hi = mid;
  }
  return r + lo;
}

/* ------------------------------------------------------------------ likelihood weights */
#define SER_LOGEPSILON (-32.236191301916641) /* log(1e-14), mcmc.h:26 */

struct SerWeights {
  double c, cc, d, dd; /* log P(false 1), log(1-e^c), log P(false 0), log(1-e^d) */
  double w1, w0;       /* log-odds of an in-range cell: one -> dd - c (>0), zero -> d - cc (<0) */
  double g, A;         /* g = -w0: log-weight gained per zero passed; A = w0 - w1: change per one passed */
  double inv_g;
  double eps;          /* exp(LOGEPSILON) as the host libm evaluates it (mcmc.c:734) */
  const double *H;     /* H[m] = sum_{u<m} exp(-u g), m = 0..hmax: geometric partial sums (per sweep);
                          NULL with per-taxon c/d (manycd): evaluated on the fly from hs */
  double hs;           /* 1 / (1 - exp(-g)) */
  int hmax;
};

/* geometric partial sum H(m) = sum_{u<m} exp(-u g).  TAB: 1 = from the per-sweep table (scalar c, d),
 * 0 = evaluated on the fly (per-taxon c, d), 2 = decided at run time by w.H.  The hot loops pass the
 * constant so that the other variant's code (a whole exp()) is not even emitted. */
template <int TAB = 2>
SER_HD double ser_H(const SerWeights &w, int m)
{
  SER_CHECK(m >= 0 && (!(TAB == 1 || (TAB == 2 && w.H)) || m <= w.hmax));
  if (TAB == 1 || (TAB == 2 && w.H)) return w.H[m];
  return SER_MUL(SER_SUB(1.0, exp(-SER_MUL(w.g, (double)m))), w.hs);
}

/* number of entries (-1) the per-sweep table H needs for a given g and N.  With the reference's prior bounds
 * (c <= log .1, d <= log .8: mcmc.h:27-30) g = log(1 - e^c) - d >= log .9 - log .8 = 0.1178, so the table never needs
 * more than 277 entries; the small-shape kernel sizes its shared-memory copy with SER_HCAP. */
#define SER_HCAP 288
SER_HD int ser_hmax(double g, int N)
{
  int h = (int)(-SER_LOGEPSILON / g) + 3;
  return h > N + 1 ? N + 1 : h;
}
/* H[m] = (1 - q^m) / (1 - q), q = exp(-g); evaluated per entry so threads can fill it in parallel */
SER_HD double ser_h_entry(double g, int m) { return (1.0 - exp(-g * (double)m)) / (1.0 - exp(-g)); }

/* derived quantities of (c, cc, d, dd); H is filled separately (ser_h_entry) */
SER_HD void ser_set_weights(struct SerWeights *wt, double c, double cc, double d, double dd);

SER_HD void ser_set_weights(struct SerWeights *wt, double c, double cc, double d, double dd)
{
  wt->c = c; wt->cc = cc; wt->d = d; wt->dd = dd;
  wt->w1 = dd - c; wt->w0 = d - cc;
  wt->g = -wt->w0; wt->A = wt->w0 - wt->w1;
  wt->inv_g = 1.0 / wt->g;
}
/* per-taxon weights (manycd): no shared table, geometric sums on the fly */
SER_HD void ser_set_weights_own(struct SerWeights *wt, double c, double cc, double d, double dd, int N)
{
  ser_set_weights(wt, c, cc, d, dd);
  wt->H = 0;
  wt->hs = 1.0 / (1.0 - exp(-wt->g));
  wt->hmax = N + 1;
}

/* int -> double.  (A 2^52-bias trick on the fp64 add pipe was measured 3-5 % slower than the native
 * conversion once three chains are resident per SM: the kernel is issue-bound, and I2F is one
 * instruction against three.) */
SER_HD double ser_i2d(int k) { return (double)k; }

/* exp(x) for x <= ~0 (weights relative to the maximum); 0 below -708.  FMA Horner, ~1 ulp.
 * On the device the coefficients sit in the constant bank, so every FMA takes its constant as an
 * operand; as literals each costs two extra instructions to materialise (this is the hot loop). */
#if defined(__CUDACC__)
__constant__ double ser_expw_c[14] = {
#else
static const double ser_expw_c[14] = {
#endif
    1.4426950408889634074,       /* 0: log2(e) */
    -6.93147180369123816490e-01, /* 1: -ln2 hi */
    -1.90821492927058770002e-10, /* 2: -ln2 lo */
    1.6059043836821613e-10,      /* 3: 1/13! */
    2.08767569878681e-09,        /* 4: 1/12! */
    2.505210838544172e-08,       /* 5: 1/11! */
    2.755731922398589e-07,       /* 6: 1/10! */
    2.7557319223985893e-06,      /* 7: 1/9!  */
    2.48015873015873e-05,        /* 8: 1/8!  */
    1.984126984126984e-04,       /* 9: 1/7!  */
    1.388888888888889e-03,       /* 10: 1/6! */
    8.333333333333333e-03,       /* 11: 1/5! */
    4.1666666666666664e-02,      /* 12: 1/4! */
    1.6666666666666666e-01};     /* 13: 1/3! */

SER_HD double ser_exp_weight(double x)
{
  if (!(x > -708.0)) return 0.0;
  const double *k = ser_expw_c;
  const double t = ser_fma(x, k[0], 6755399441055744.0); /* round to nearest int */
  const double kd = t - 6755399441055744.0;
  double r = ser_fma(kd, k[1], x);
  r = ser_fma(kd, k[2], r);
  double p = k[3];
  p = ser_fma(p, r, k[4]);
  p = ser_fma(p, r, k[5]);
  p = ser_fma(p, r, k[6]);
  p = ser_fma(p, r, k[7]);
  p = ser_fma(p, r, k[8]);
  p = ser_fma(p, r, k[9]);
  p = ser_fma(p, r, k[10]);
  p = ser_fma(p, r, k[11]);
  p = ser_fma(p, r, k[12]);
  p = ser_fma(p, r, k[13]);
  p = ser_fma(p, r, 0.5);
  p = ser_fma(p, r, 1.0);
  p = ser_fma(p, r, 1.0);
  const int64_t kk = (int64_t)(int32_t)(uint32_t)ser_d2u(t); /* low word of the magic sum = k */
  return ser_u2d(ser_d2u(p) + ((uint64_t)kk << 52));
}

/*
 * One run of candidates inside which only zeros are passed: n candidates whose log-weights rise
 * by g per step and end at le (<= ~0, relative to the maximum).  The reference floors every
 * weight at exp(LOGEPSILON) (mcmc.c:734); m = the candidates that stay above the floor (they
 * are the LAST m of the run).  Returns the run's total weight.
 */
template <int TAB = 2>
SER_HD double ser_run_sum(const SerWeights &w, int n, double le, int *m_out, double *ye_out)
{
  if (le < SER_LOGEPSILON) { *m_out = 0; *ye_out = 0.0; return SER_MUL(ser_i2d(n), w.eps); }
  int m = (int)SER_MUL(SER_SUB(le, SER_LOGEPSILON), w.inv_g) + 1;
  if (m > n) m = n;
  if (m > w.hmax) m = w.hmax;
  const double ye = ser_exp_weight(le);
  *m_out = m; *ye_out = ye;
  return ser_fma(ye, ser_H<TAB>(w, m), SER_MUL(ser_i2d(n - m), w.eps));
}

/*
 * ---- Gibbs draw of one boundary, item formulation ---------------------------------------
 * (the a-step, or the b-step on the reversed column: mcmc_auxa + mcmc_logtop +
 * mcmc_randompick, mcmc.c:828-915)
 *
 * Logical string s (REV: s[q] = v[N-1-q]); candidates 0..bound; `cur` = current value;
 * weight(i) ~ exp(L(i)) floored at eps relative to the maximum, L(i) = dn1 A + di g with dn1 /
 * di the ones / cells between cur and i.  Passing a zero raises L by g, passing a one lowers it
 * by w1, so the candidates split into RUNS that end on a local maximum: the candidate just
 * below each one of s[0,bound) and the candidate `bound`.  One run = one ITEM; its weight is the
 * closed form ser_run_sum().  A column with K ones below `bound` has K+1 items, so the whole
 * step costs O(#ones) instead of O(#sites), and the items of all taxa of a chain are evaluated
 * densely across the CTA (ser_kernels.cu) -- only a short scan over a taxon's own items stays
 * with the taxon's thread.
 */
struct SerStep {
  int cur, bound; /* logical coordinates */
  int ocur;       /* ones of s below cur */
  int kb;         /* ones of s below bound = number of one-items; item kb is the bound item */
  int nones, N, rev;
};

/* ascending positions of the ones of a column */
SER_HD int ser_expand_ones(const uint32_t *col, int C, int W, uint16_t *out)
{
  int k = 0;
  for (int w = 0; w < W; w++) {
    uint32_t v = col[w * C];
    while (v) { out[k++] = (uint16_t)(32 * w + SER_FFS(v) - 1); v &= v - 1u; }
  }
  return k;
}

template <int G = 0>
SER_HD SerStep ser_step_a(const uint32_t *col, const uint16_t *pre, int C, int W, int N, int a, int b)
{
  SerStep st;
  st.cur = a; st.bound = b; st.ocur = ser_rank1<G>(col, pre, C, a); st.kb = ser_rank1<G>(col, pre, C, b);
  st.nones = ser_pre_at<G>(col, pre, C, W); st.N = N; st.rev = 0;
  return st;
}
/* b-step on the reversed column: boundary t = N - b, candidates 0..N-a */
template <int G = 0>
SER_HD SerStep ser_step_b(const uint32_t *col, const uint16_t *pre, int C, int W, int N, int a, int b)
{
  SerStep st;
  st.nones = ser_pre_at<G>(col, pre, C, W); st.N = N; st.rev = 1;
  st.cur = N - b; st.bound = N - a;
  st.ocur = st.nones - ser_rank1<G>(col, pre, C, b); st.kb = st.nones - ser_rank1<G>(col, pre, C, a);
  return st;
}

/* logical position of the kk-th logical one (kk < kb); pos[] = ascending physical positions */
SER_HD int ser_item_q(const SerStep &st, const uint16_t *pos, int kk)
{
  SER_CHECK(kk >= 0 && kk < st.nones);
  return st.rev ? st.N - 1 - (int)pos[st.nones - 1 - kk] : (int)pos[kk];
}

/* item kk (0..kb): last candidate q, run length n, log-weight of q relative to cur */
SER_HD double ser_item_eval(const SerWeights &wt, const SerStep &st, const uint16_t *pos, int kk, int *q_out, int *n_out)
{
  const int q = kk < st.kb ? ser_item_q(st, pos, kk) : st.bound;
  const int qprev = kk > 0 ? ser_item_q(st, pos, kk - 1) : -1;
  *q_out = q; *n_out = q - qprev;
  return ser_fma(ser_i2d(kk - st.ocur), wt.A, SER_MUL(ser_i2d(q - st.cur), wt.g));
}

/* maximum log-weight over the items (the reference's z, mcmc.c:727-730) */
SER_HD double ser_step_lmax(const SerWeights &wt, const SerStep &st, const uint16_t *pos)
{
  double lmax = -1.0e300;
  for (int kk = 0; kk <= st.kb; kk++) {
    int q, n;
    lmax = ser_fmax(lmax, ser_item_eval(wt, st, pos, kk, &q, &n));
  }
  return lmax;
}

/* the same pass, keeping every item's log-weight and run length for the dense pass (which then needs
 * neither the postings nor the step geometry): Lc[kk], nc[kk] for kk = 0..kb */
SER_HD double ser_step_lmax_cache(const SerWeights &wt, const SerStep &st, const uint16_t *pos, double *Lc, uint16_t *nc)
{
  double lmax = -1.0e300;
  for (int kk = 0; kk <= st.kb; kk++) {
    int q, n;
    const double L = ser_item_eval(wt, st, pos, kk, &q, &n);
    Lc[kk] = L; nc[kk] = (uint16_t)n;
    lmax = ser_fmax(lmax, L);
  }
  return lmax;
}

/* weight of an item from its cached log-weight and run length (bit for bit ser_item_weight) */
template <int TAB = 2>
SER_HD double ser_item_weight_cached(const SerWeights &wt, double L, int n, double lmax)
{
  int m; double ye;
  return ser_run_sum<TAB>(wt, n, SER_SUB(L, lmax), &m, &ye);
}

/* weight of item kk given the step's maximum */
template <int TAB = 2>
SER_HD double ser_item_weight(const SerWeights &wt, const SerStep &st, const uint16_t *pos, int kk, double lmax)
{
  int q, n, m; double ye;
  const double le = SER_SUB(ser_item_eval(wt, st, pos, kk, &q, &n), lmax);
  return ser_run_sum<TAB>(wt, n, le, &m, &ye);
}

/* inside one run: s = cumulative weight before it; returns t in [0,n): the first candidate of
 * the run whose cumulative weight reaches target (the last one if none does) */
template <int TAB = 2>
SER_HD int ser_run_pick(const SerWeights &wt, int n, double le, double s, double target)
{
  int m; double ye;
  ser_run_sum<TAB>(wt, n, le, &m, &ye);
  const int nf = n - m; /* nf floored candidates of weight eps each, then m geometric ones */
  const double sf = ser_fma(ser_i2d(nf), wt.eps, s);
  if (nf > 0 && (sf >= target || m == 0)) {
    const double td = SER_DIV(SER_SUB(target, s), wt.eps); /* ~ whole eps steps below the target */
    int t = !(td > 0.0) ? 0 : (td >= (double)(nf - 1) ? nf - 1 : (int)td);
    while (t > 0 && ser_fma(ser_i2d(t), wt.eps, s) >= target) t--;
    while (t < nf - 1 && ser_fma(ser_i2d(t + 1), wt.eps, s) < target) t++;
    return t;
  }
  /* cumulative weight through geometric candidate k (k = 0..m-1): sf + ye (H[m] - H[m-1-k]) */
  const double hm = ser_H<TAB>(wt, m);
  int lo = 0, hi = m - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ser_fma(ye, SER_SUB(hm, ser_H<TAB>(wt, m - 1 - mid)), sf) >= target) hi = mid; else lo = mid + 1;
  }
  return nf + lo;
}

/* val[0..kb] already holds the CUMULATIVE item weights: inverts the CDF (mcmc_randompick) and
 * returns the picked candidate (logical index) */
template <int TAB = 2>
SER_HD int ser_step_pick_scanned(const SerWeights &wt, const SerStep &st, const uint16_t *pos, const double *val, double lmax,
                                 double U)
{
  const double target = SER_MUL(U, val[st.kb]);
  int lo = 0, hi = st.kb; /* first item whose cumulative weight reaches the target */
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (val[mid] >= target) hi = mid; else lo = mid + 1;
  }
  int q, n;
  const double le = SER_SUB(ser_item_eval(wt, st, pos, lo, &q, &n), lmax);
  return q - n + 1 + ser_run_pick<TAB>(wt, n, le, lo ? val[lo - 1] : 0.0, target);
}

/* the taxon's own part: val[0..kb] holds the item weights; turns them into cumulative sums and picks */
template <int TAB = 2>
SER_HD int ser_step_pick(const SerWeights &wt, const SerStep &st, const uint16_t *pos, double *val, double lmax,
                         double U)
{
  double S = 0.0;
  for (int kk = 0; kk <= st.kb; kk++) { S = SER_ADD(S, val[kk]); val[kk] = S; }
  return ser_step_pick_scanned<TAB>(wt, st, pos, val, lmax, U);
}

/* ------------------------------------------------------------------ pi proposals */
/* mcmc_ininterval, mcmc.c:1097-1124 (lo <= hi at every call site) */
SER_HD int ser_in_window(int v, int lo, int hi, int inc_lo, int inc_hi)
{
  return (inc_lo ? lo <= v : lo < v) && (inc_hi ? v <= hi : v < hi);
}

SER_HD void ser_mirror_ab(int a, int b, int ain, int bin, int s, int *na, int *nb)
{
  *na = a; *nb = b;
  if (ain && !bin) *na = s - a;
  else if (!ain && bin) *nb = s - b;
  else if (ain && bin) { *nb = s - a; *na = s - b; }
}

/* pi1: site at position i moves to j.  Only the moved site's status can change. */
SER_HD void ser_pi1_delta(const uint32_t *col, int C, int a, int b, int i, int j, int *dt0, int *dt1)
{
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  int gain, lose;
  if (i < j) {
    const int ain = lo < a && a <= hi + 1, bin = lo < b && b <= hi + 1;
    gain = ain && !bin; lose = !ain && bin;
  } else {
    const int ain = lo <= a && a <= hi, bin = lo <= b && b <= hi;
    gain = !ain && bin; lose = ain && !bin;
  }
  *dt0 = 0; *dt1 = 0;
  if (gain | lose) {
    const int one = ser_col_bit(col, C, i);
    const int sgn = gain ? 1 : -1;
    if (one) *dt1 = sgn; else *dt0 = -sgn;
  }
}
SER_HD void ser_pi1_apply_ab(int *a, int *b, int i, int j)
{
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  if (i < j) {
    if (lo < *a && *a <= hi + 1) (*a)--;
    if (lo < *b && *b <= hi + 1) (*b)--;
  } else {
    if (lo <= *a && *a <= hi) (*a)++;
    if (lo <= *b && *b <= hi) (*b)++;
  }
}

/* pi2: positions [i, j] reversed */
template <int G = 0>
SER_HD void ser_pi2_delta(const uint32_t *col, const uint16_t *pre, int C, int a, int b, int i, int j, int inc1,
                          int inc2, int *dt0, int *dt1)
{
  const int ain = ser_in_window(a, i, j + 1, inc1, inc2), bin = ser_in_window(b, i, j + 1, inc1, inc2);
  *dt0 = 0; *dt1 = 0;
  if (ain == bin) return;
  const int split = ain ? a : b;
  const int r = ser_rank1<G>(col, pre, C, split);
  const int oL = r - ser_rank1<G>(col, pre, C, i), oR = ser_rank1<G>(col, pre, C, j + 1) - r;
  const int zL = (split - i) - oL, zR = (j + 1 - split) - oR;
  if (ain) { *dt1 = oL - oR; *dt0 = zR - zL; }  /* left part becomes alive, right part dies */
  else { *dt1 = oR - oL; *dt0 = zL - zR; }      /* left part dies, right part becomes alive */
}

/* window geometry of a pi3 proposal, shared by all taxa */
struct SerPi3 {
  int i, j;        /* window positions (both non-hard) */
  int irank;       /* global non-hard rank of position i */
  int K;           /* non-hard sites in the window */
  int hr_i;        /* hard rank of i */
};
SER_HD SerPi3 ser_pi3_window(const SerHard &h, int irank, int jrank)
{
  SerPi3 g;
  g.i = ser_select_nonhard(h, irank);
  g.j = ser_select_nonhard(h, jrank);
  g.irank = irank;
  g.K = jrank - irank + 1;
  g.hr_i = ser_hard_rank(h, g.i);
  return g;
}
/* non-hard positions in [i, x), x in [i, j+1] */
SER_HD int ser_pi3_wrank(const SerHard &h, const SerPi3 &g, int x) { return (x - g.i) - (ser_hard_rank(h, x) - g.hr_i); }
/* position of the window's rho-th non-hard site; rho == K -> j+1 */
SER_HD int ser_pi3_wpos(const SerHard &h, const SerPi3 &g, int rho) { return rho >= g.K ? g.j + 1 : ser_select_nonhard(h, g.irank + rho); }
/* where position n of the window goes (the involution p[] of mcmc.c:1534-1555) */
SER_HD int ser_pi3_perm(const SerHard &h, const SerPi3 &g, int n)
{
  if (ser_is_hard(h, n)) return n;
  return ser_select_nonhard(h, g.irank + g.K - 1 - ser_pi3_wrank(h, g, n));
}

/* pi3: only the non-hard sites of [i, j] are reversed; a/b mirror as in pi2 */
/* HB: the column's ones at the hard sites come from `hbits` (bit k = the column has a one at the k-th hard site;
 * hard sites keep their relative order, mcmc.c:1049-1072, so this is a constant of the column; needs nh <= 32)
 * instead of a walk over the hard positions inside the two ranges */
template <bool HB = false, int G = 0>
SER_HD void ser_pi3_delta(const uint32_t *col, const uint16_t *pre, int C, const SerHard &h, const SerPi3 &g, int a,
                          int b, int inc1, int inc2, int *dt0, int *dt1, uint32_t hbits = 0u)
{
  const int i = g.i, j = g.j;
  const int ain = ser_in_window(a, i, j + 1, inc1, inc2), bin = ser_in_window(b, i, j + 1, inc1, inc2);
  *dt0 = 0; *dt1 = 0;
  if (!ain && !bin) return; /* no boundary inside the window: every cell keeps its status */
  int na, nb;
  ser_mirror_ab(a, b, ain, bin, i + j + 1, &na, &nb);
  /* was alive: [a,b) cut to the window */
  const int alo = a > i ? a : i, ahi = b < j + 1 ? b : j + 1;
  /* is alive afterwards, in OLD position coordinates:
   *   hard sites stay put            -> hard positions in [na,nb)
   *   non-hard site of window rank r -> lands on rank K-1-r, alive iff that is in [R(na),R(nb)),
   *                                     i.e. non-hard positions in [lo2, hi2) */
  int nac = na < i ? i : (na > j + 1 ? j + 1 : na), nbc = nb < i ? i : (nb > j + 1 ? j + 1 : nb);
  if (nbc < nac) nbc = nac;
  const int hr_na = ser_hard_rank(h, nac), hr_nb = ser_hard_rank(h, nbc);
  const int lo2 = ser_pi3_wpos(h, g, g.K - ((nbc - i) - (hr_nb - g.hr_i)));
  const int hi2 = ser_pi3_wpos(h, g, g.K - ((nac - i) - (hr_na - g.hr_i)));
  const int hr_lo2 = ser_hard_rank(h, lo2), hr_hi2 = ser_hard_rank(h, hi2);
  const int nA = ahi > alo ? ahi - alo : 0, oA = ser_col_popc<G>(col, pre, C, alo, ahi);
  /* B = (hard in [nac,nbc)) + (non-hard in [lo2,hi2)) */
  const int nB = (hr_nb - hr_na) + ((hi2 > lo2 ? hi2 - lo2 : 0) - (hr_hi2 - hr_lo2));
  int oB = ser_col_popc<G>(col, pre, C, lo2, hi2);
  if (HB) { /* hard ranks in [x, y) -> bits x..y-1 of hbits */
    oB -= SER_POPC(hbits & ser_mask_lt(hr_hi2) & ~ser_mask_lt(hr_lo2));
    oB += SER_POPC(hbits & ser_mask_lt(hr_nb) & ~ser_mask_lt(hr_na));
  } else {
    for (int k = hr_lo2; k < hr_hi2; k++) oB -= ser_col_bit(col, C, h.hp[k]); /* hard ones inside [lo2,hi2) do not count */
    for (int k = hr_na; k < hr_nb; k++) oB += ser_col_bit(col, C, h.hp[k]);   /* hard ones in [nac,nbc) do */
  }
  *dt1 = oB - oA;
  *dt0 = (nA - oA) - (nB - oB); /* true zeros = dead zeros: gain what the alive zeros lose */
}

/* per-taxon likelihood term in the reference's operand order (mcmc.c:1214/1435/1630),
 * df0 = -dt0 and df1 = -dt1 always; never contracted into FMAs */
SER_HD double ser_term(const SerWeights &w, int dt0, int dt1)
{
  return SER_ADD(SER_ADD(SER_ADD(SER_MUL((double)dt0, w.cc), SER_MUL((double)(-dt0), w.d)),
                         SER_MUL((double)dt1, w.dd)),
                 SER_MUL((double)(-dt1), w.c));
}

/* per-taxon counts from the column: t1 = ones alive, the rest by difference (mcmc.c:651-708) */
SER_HD void ser_counts(const uint32_t *col, const uint16_t *pre, int C, int N, int a, int b, int ones, int *t0,
                       int *f0, int *t1, int *f1)
{
  const int o = ser_col_popc(col, pre, C, a, b);
  *t1 = o;
  *f1 = ones - o;
  *f0 = (b - a) - o;
  *t0 = N - (b - a) - (ones - o);
}

/* draws -> integers exactly as the tape grammar defines them */
SER_HD int ser_draw_int(double u, int n) { return (int)SER_MUL(u, (double)n); }

#endif /* SER_CHAIN_CORE_H */
