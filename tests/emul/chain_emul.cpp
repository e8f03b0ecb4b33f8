// chain_emul.cpp -- CPU emulation of the sweep kernel's choreography (TEST ONLY).
//
// Runs ONE chain with the very same per-taxon building blocks the CUDA kernel
// uses (csrc/ser_chain_core.h compiled for the host), "threads" being a plain
// loop over columns and block reductions being plain sums.  It exists so that
// the bit-level logic (range popcounts, rank/select over the hard mask, the
// pi3 mask construction, the Gibbs walk) can be checked against the oracle on
// a machine without a GPU.  It is not part of the product and not a fallback:
// nothing in the package loads it.
//
// Replay mode only (tape grammar of oracle/draw_source.h).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../seriation-in-paleontological-data-using-mcmc_b200/csrc/ser_chain_core.h"

namespace {

constexpr double MINC = -6.9077552789821368, MAXC = -2.3025850929940455;
constexpr double MIND = -1.6094379124341003, MAXD = -0.22314355131420971;

struct Chain {
  int N, M, W, C, nh;
  std::vector<uint8_t> X, hard;
  std::vector<uint32_t> V; // [W][C], column M = hard mask
  std::vector<int> a, b, ones;
  std::vector<uint16_t> pre, hp; // pre[(W+1)][C] prefix ones; hp = sorted hard positions
  std::vector<uint16_t> rpi;
  SerWeights wt;
  bool manycd = false;
  std::vector<SerWeights> wts; // per-taxon weights (manycd)
  const SerWeights &WT(int m) const { return manycd ? wts[m] : wt; }
  double loglik;
  int t0a, f0a, t1a, f1a;
  const double *tape;
  size_t tape_len, cur;
  long long n_degenerate;
  long long chunk_picks = 0, chunk_mismatch = 0, chunk_finder_bad = 0; // the warp batches' chunked passes beside the sequential ones

  uint32_t *col(int m) { return V.data() + m; }
  uint16_t *pcol(int m) { return pre.data() + m; }
  SerHard hardinfo() const { return SerHard{V.data() + M, pre.data() + M, hp.data(), C, W, N, nh}; }
  double next() {
    if (cur >= tape_len) { fprintf(stderr, "emul: tape exhausted\n"); exit(3); }
    return tape[cur++];
  }
  std::vector<double> H;
  void set_cd(double c, double cc, double d, double dd) {
    ser_set_weights(&wt, c, cc, d, dd);
    wt.hmax = ser_hmax(wt.g, N);
    H.resize(wt.hmax + 1);
    for (int m = 0; m <= wt.hmax; m++) H[m] = ser_h_entry(wt.g, m);
    wt.H = H.data();
  }
  void rebuild_hard() { ser_hard_list(V.data() + M, C, W, hp.data()); }
  /* the column moves maintain the prefix tables themselves; this re-derives them the slow way and
   * aborts on any difference (the pi3 permutation leaves the hard column alone) */
  void fix_pre(int p0, int p1) {
    const std::vector<uint16_t> kept(pre);
    for (int m = 0; m <= M; m++) ser_col_fix_pre(col(m), pcol(m), C, p0 >> 5, p1 >> 5);
    if (kept != pre) { fprintf(stderr, "emulator: fused prefix update differs from ser_col_fix_pre\n"); abort(); }
  }
  void build_columns() {
    std::fill(V.begin(), V.end(), 0u);
    for (int p = 0; p < N; p++) {
      const int site = rpi[p];
      for (int m = 0; m < M; m++)
        if (X[(size_t)site * M + m]) V[(p >> 5) * C + m] |= 1u << (p & 31);
      if (hard[site]) V[(p >> 5) * C + M] |= 1u << (p & 31);
    }
    for (int m = 0; m <= M; m++) ser_col_build_pre(col(m), pcol(m), C, W);
    rebuild_hard();
  }
  void initab() {
    for (int m = 0; m < M; m++) {
      int first = -1, last = -1;
      for (int w = 0; w < W; w++) {
        uint32_t v = V[w * C + m];
        if (v) { if (first < 0) first = 32 * w + SER_FFS(v) - 1; last = 32 * w + 31 - __builtin_clz(v); }
      }
      if (first < 0) { a[m] = 0; b[m] = N; } else { a[m] = first; b[m] = last + 1; }
    }
  }
  void totals() {
    long T1 = 0, LEN = 0, ONES = 0;
    for (int m = 0; m < M; m++) { T1 += ser_col_popc(col(m), pcol(m), C, a[m], b[m]); LEN += b[m] - a[m]; ONES += ones[m]; }
    t1a = (int)T1; f1a = (int)(ONES - T1); f0a = (int)(LEN - T1); t0a = (int)((long)N * M - LEN - f1a);
    loglik = t0a * wt.cc + f0a * wt.d + t1a * wt.dd + f1a * wt.c;
    if (manycd) { // mcmc_logl, mcmc.c:625-648, per-taxon coefficients
      loglik = 0.0;
      for (int m = 0; m < M; m++) {
        int t0, f0, t1, f1;
        ser_counts(col(m), pcol(m), C, N, a[m], b[m], ones[m], &t0, &f0, &t1, &f1);
        loglik += t0 * wts[m].cc + f0 * wts[m].d + t1 * wts[m].dd + f1 * wts[m].c;
      }
    }
  }
};

// one MH decision from the block-reduced integer deltas; emulates the reference's float sum
// only where its sign is not determined by the integers (DESIGN.md "degenerate proposals")
bool decide(Chain &ch, const std::vector<int> &dt0, const std::vector<int> &dt1, double *delta_out) {
  long D0 = 0, D1 = 0; bool any = false;
  for (int m = 0; m < ch.M; m++) { D0 += dt0[m]; D1 += dt1[m]; any |= (dt0[m] | dt1[m]) != 0; }
  double delta;
  if (ch.manycd) { // per-taxon coefficients: the reference's sequential sum is the definition
    delta = 0.0;
    for (int m = 0; m < ch.M; m++) delta = delta + ser_term(ch.wts[m], dt0[m], dt1[m]);
  } else if (D0 == 0 && D1 == 0) {
    delta = 0.0;
    if (any) { // sequential sum of the per-taxon terms, in taxon order, like mcmc.c:1214
      for (int m = 0; m < ch.M; m++) delta = delta + ser_term(ch.wt, dt0[m], dt1[m]);
      if (delta != 0.0) ch.n_degenerate++;
    }
  } else {
    delta = ser_term(ch.wt, (int)D0, (int)D1);
  }
  *delta_out = delta;
  if (delta >= 0.0) return true;
  return delta > std::log(ch.next());
}

void after_accept(Chain &ch, long D0, long D1, double delta) {
  ch.t0a += (int)D0; ch.f0a -= (int)D0; ch.t1a += (int)D1; ch.f1a -= (int)D1;
  ch.loglik += delta;
}

int pi1(Chain &ch) {
  const int N = ch.N, M = ch.M, C = ch.C, W = ch.W;
  int i = ser_draw_int(ch.next(), N), j = ser_draw_int(ch.next(), N - 1);
  if (j >= i) j++;
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  SerHard h = ch.hardinfo();
  if (ser_is_hard(h, i) && ser_hard_count(h, lo, hi) > 1) return 0;
  std::vector<int> dt0(M), dt1(M);
  for (int m = 0; m < M; m++) ser_pi1_delta(ch.col(m), C, ch.a[m], ch.b[m], i, j, &dt0[m], &dt1[m]);
  double delta;
  if (!decide(ch, dt0, dt1, &delta)) return 0;
  long D0 = 0, D1 = 0;
  for (int m = 0; m < M; m++) { D0 += dt0[m]; D1 += dt1[m]; }
  for (int m = 0; m < M; m++) ser_pi1_apply_ab(&ch.a[m], &ch.b[m], i, j);
  for (int m = 0; m <= M; m++) ser_col_rotate(ch.col(m), C, W, i, j, ch.pcol(m));
  const uint16_t t = ch.rpi[i];
  if (i < j) for (int n = i; n < j; n++) ch.rpi[n] = ch.rpi[n + 1];
  else for (int n = i; n > j; n--) ch.rpi[n] = ch.rpi[n - 1];
  ch.rpi[j] = t;
  ch.fix_pre(lo, hi);
  ch.rebuild_hard();
  after_accept(ch, D0, D1, delta);
  return 1;
}

int pi2(Chain &ch, int swap) {
  const int N = ch.N, M = ch.M, C = ch.C, W = ch.W;
  int i, j;
  if (!swap) {
    i = ser_draw_int(ch.next(), N); j = ser_draw_int(ch.next(), N - 1);
    if (j >= i) j++; else { int t = i; i = j; j = t; }
  } else { i = ser_draw_int(ch.next(), N - 1); j = i + 1; }
  SerHard h = ch.hardinfo();
  if (ser_hard_count(h, i, j) > 1) return 0;
  const int inc1 = ser_draw_int(ch.next(), 2), inc2 = ser_draw_int(ch.next(), 2);
  std::vector<int> dt0(M), dt1(M);
  for (int m = 0; m < M; m++) ser_pi2_delta(ch.col(m), ch.pcol(m), C, ch.a[m], ch.b[m], i, j, inc1, inc2, &dt0[m], &dt1[m]);
  double delta;
  if (!decide(ch, dt0, dt1, &delta)) return 0;
  long D0 = 0, D1 = 0;
  for (int m = 0; m < M; m++) { D0 += dt0[m]; D1 += dt1[m]; }
  for (int m = 0; m < M; m++) {
    const int ain = ser_in_window(ch.a[m], i, j + 1, inc1, inc2), bin = ser_in_window(ch.b[m], i, j + 1, inc1, inc2);
    ser_mirror_ab(ch.a[m], ch.b[m], ain, bin, i + j + 1, &ch.a[m], &ch.b[m]);
  }
  for (int m = 0; m <= M; m++) ser_col_reverse(ch.col(m), C, W, i, j, ch.pcol(m));
  for (int l = i, r = j; l < r; l++, r--) { uint16_t t = ch.rpi[l]; ch.rpi[l] = ch.rpi[r]; ch.rpi[r] = t; }
  ch.fix_pre(i, j);
  ch.rebuild_hard();
  after_accept(ch, D0, D1, delta);
  return 1;
}

int pi3(Chain &ch) {
  const int N = ch.N, M = ch.M, C = ch.C, W = ch.W, nfree = N - ch.nh;
  if (nfree < 2) return 0;
  const int r1 = ser_draw_int(ch.next(), nfree), r2 = ser_draw_int(ch.next(), nfree - 1);
  int ir, jr;
  if (r1 <= r2) { ir = r1; jr = r2 + 1; } else { ir = r2; jr = r1; }
  SerHard h = ch.hardinfo();
  const SerPi3 g = ser_pi3_window(h, ir, jr);
  const int inc1 = ser_draw_int(ch.next(), 2), inc2 = ser_draw_int(ch.next(), 2);
  std::vector<int> dt0(M), dt1(M);
  for (int m = 0; m < M; m++) ser_pi3_delta(ch.col(m), ch.pcol(m), C, h, g, ch.a[m], ch.b[m], inc1, inc2, &dt0[m], &dt1[m]);
  double delta;
  if (!decide(ch, dt0, dt1, &delta)) return 0;
  long D0 = 0, D1 = 0;
  for (int m = 0; m < M; m++) { D0 += dt0[m]; D1 += dt1[m]; }
  std::vector<uint16_t> perm(N);
  for (int n = g.i; n <= g.j; n++) perm[n] = (uint16_t)ser_pi3_perm(h, g, n);
  for (int m = 0; m < M; m++) {
    const int ain = ser_in_window(ch.a[m], g.i, g.j + 1, inc1, inc2), bin = ser_in_window(ch.b[m], g.i, g.j + 1, inc1, inc2);
    ser_mirror_ab(ch.a[m], ch.b[m], ain, bin, g.i + g.j + 1, &ch.a[m], &ch.b[m]);
  }
  for (int m = 0; m < M; m++) ser_col_permute(ch.col(m), C, W, g.i, g.j, perm.data(), ch.pcol(m));
  std::vector<uint16_t> tmp(ch.rpi);
  for (int n = g.i; n <= g.j; n++) ch.rpi[n] = tmp[perm[n]];
  ch.fix_pre(g.i, g.j);
  after_accept(ch, D0, D1, delta);
  return 1;
}

void sample_cd(Chain &ch) {
  if (ch.manycd) { // mcmc.c:777-785, :807-815: M Betas for c, then M for d (the draws do not depend on acceptance)
    for (int pass = 0; pass < 2; pass++)
      for (int m = 0; m < ch.M; m++) {
        const double y = ch.next(), ly = ch.next(), l1 = ch.next();
        SerWeights &w = ch.wts[m];
        double c = w.c, cc = w.cc, d = w.d, dd = w.dd;
        if (pass == 0) { if (y > 0. && MINC <= ly && ly <= MAXC) { c = ly; cc = l1; } }
        else { if (y > 0. && MIND <= ly && ly <= MAXD) { d = ly; dd = l1; } }
        const double eps = w.eps;
        ser_set_weights_own(&w, c, cc, d, dd, ch.N);
        w.eps = eps;
      }
    return;
  }
  double c = ch.wt.c, cc = ch.wt.cc, d = ch.wt.d, dd = ch.wt.dd;
  { const double y = ch.next(), ly = ch.next(), l1 = ch.next();
    if (y > 0. && MINC <= ly && ly <= MAXC) { c = ly; cc = l1; } }
  { const double y = ch.next(), ly = ch.next(), l1 = ch.next();
    if (y > 0. && MIND <= ly && ly <= MAXD) { d = ly; dd = l1; } }
  ch.set_cd(c, cc, d, dd);
}

// The large-shape kernel's warp batches (ser_sweep_kernel_big.cuh, gibbs_warp): a column's 2^lsh lanes own contiguous chunks of
// its items through all passes.  Emulated lane by lane: the chunk-fused log-weights / run weights must equal the dense ones bit
// for bit (same formula, same operands); the chunk-relative cumulative sums, the scan of the chunk totals (the shuffle steps in
// their order), the first chunk that reaches U x total, the search inside it and the closed-form pick give the column's new
// boundary, which is compared with the sequential inverse CDF (ser_step_pick).
int chunked_pick(Chain &ch, const SerWeights &w, const SerStep &st, const uint16_t *pos, const double *dense_w, double lmax, double U, int lsh) {
  const int lpc = 1 << lsh, kb = st.kb, chunk = (kb + lpc) >> lsh;
  std::vector<double> cum(kb + 1), incl(lpc), tot(lpc);
  std::vector<int> k0(lpc), k1(lpc);
  double lm_all = -1.0e300;
  std::vector<double> L(kb + 1);
  for (int sub = 0; sub < lpc; sub++) { // pass 1: log-weights of the chunk, partial maximum
    k0[sub] = std::min(kb + 1, sub * chunk); k1[sub] = std::min(kb + 1, k0[sub] + chunk);
    for (int kk = k0[sub]; kk < k1[sub]; kk++) {
      const int q = kk < kb ? ser_item_q(st, pos, kk) : st.bound;
      L[kk] = ser_fma(ser_i2d(kk - st.ocur), w.A, SER_MUL(ser_i2d(q - st.cur), w.g));
      lm_all = ser_fmax(lm_all, L[kk]);
    }
  }
  if (lm_all != lmax) { fprintf(stderr, "emulator: chunked maximum differs\n"); abort(); }
  for (int sub = 0; sub < lpc; sub++) { // pass 2: run weights into the chunk's cumulative sums
    int qprev = (k0[sub] > 0 && k0[sub] < k1[sub]) ? ser_item_q(st, pos, k0[sub] - 1) : -1;
    double t = 0.0;
    for (int kk = k0[sub]; kk < k1[sub]; kk++) {
      const int q = kk < kb ? ser_item_q(st, pos, kk) : st.bound;
      const double wk = ser_item_weight_cached(w, L[kk], q - qprev, lm_all);
      if (wk != dense_w[kk]) { fprintf(stderr, "emulator: chunk-fused item weight differs\n"); abort(); }
      t = SER_ADD(t, wk); cum[kk] = t; qprev = q;
    }
    tot[sub] = t; incl[sub] = t;
  }
  for (int o = 1; o < lpc; o <<= 1) { // the shuffle scan: every lane adds the value of the lane o below, all lanes at once
    const std::vector<double> old(incl);
    for (int sub = o; sub < lpc; sub++) incl[sub] = SER_ADD(old[sub], old[sub - o]);
  }
  const double total = incl[lpc - 1], target = SER_MUL(U, total);
  int finder = -1, finders = 0;
  for (int sub = 0; sub < lpc; sub++) {
    const double base = sub ? incl[sub - 1] : 0.0;
    if (k0[sub] < k1[sub] && incl[sub] >= target && (sub == 0 || base < target)) { finders++; if (finder < 0) finder = sub; }
  }
  if (finders != 1) { ch.chunk_finder_bad++; if (finder < 0) return -1; }
  const double base = finder ? incl[finder - 1] : 0.0;
  int lo = k0[finder], hi = k1[finder] - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (SER_ADD(base, cum[mid]) >= target) hi = mid; else lo = mid + 1;
  }
  const double rest = SER_SUB(target, lo > k0[finder] ? SER_ADD(base, cum[lo - 1]) : base);
  int q, n;
  const double le = SER_SUB(ser_item_eval(w, st, pos, lo, &q, &n), lm_all);
  return q - n + 1 + ser_run_pick(w, n, le, 0.0, rest);
}

int sample_ab(Chain &ch) {
  // item formulation: (1) postings of every column, (2) per step: per-taxon maximum, dense item
  // weights, per-taxon scan + pick -- the loops below are the kernel's phases run sequentially
  int changed = 0;
  const int M = ch.M, C = ch.C, W = ch.W, N = ch.N;
  std::vector<int> off(M + 1, 0);
  for (int m = 0; m < M; m++) off[m + 1] = off[m] + ch.ones[m] + 1;
  std::vector<uint16_t> pos(off[M] + 1);
  std::vector<double> val(off[M] + 1), lmax(M), ua(M), ub(M);
  std::vector<uint16_t> nrun(off[M] + 1);
  std::vector<SerStep> st(M);
  for (int m = 0; m < M; m++) { ua[m] = ch.next(); ub[m] = ch.next(); }
  for (int m = 0; m < M; m++) ser_expand_ones(ch.col(m), C, W, pos.data() + off[m]);
  for (int step = 0; step < 2; step++) {
    for (int m = 0; m < M; m++) {
      st[m] = step == 0 ? ser_step_a(ch.col(m), ch.pcol(m), C, W, N, ch.a[m], ch.b[m])
                        : ser_step_b(ch.col(m), ch.pcol(m), C, W, N, ch.a[m], ch.b[m]);
      lmax[m] = ser_step_lmax_cache(ch.WT(m), st[m], pos.data() + off[m], val.data() + off[m], nrun.data() + off[m]);
      if (lmax[m] != ser_step_lmax(ch.WT(m), st[m], pos.data() + off[m])) { fprintf(stderr, "emulator: cached maximum differs\n"); abort(); }
    }
    for (int m = 0; m < M; m++)          // dense over items in the kernel: from the cached log-weight and run length
      for (int kk = 0; kk <= st[m].kb; kk++) {
        const double w = ser_item_weight_cached(ch.WT(m), val[off[m] + kk], nrun[off[m] + kk], lmax[m]);
        if (w != ser_item_weight(ch.WT(m), st[m], pos.data() + off[m], kk, lmax[m])) { fprintf(stderr, "emulator: cached item weight differs\n"); abort(); }
        val[off[m] + kk] = w;
      }
    for (int m = 0; m < M; m++) {
      const int lsh = (m + step + (int)(ch.cur & 7)) % 6; // 1 .. 32 lanes per column, varied over columns, steps and sweeps
      const int cpick = chunked_pick(ch, ch.WT(m), st[m], pos.data() + off[m], val.data() + off[m], lmax[m], step == 0 ? ua[m] : ub[m], lsh);
      const int pick = ser_step_pick(ch.WT(m), st[m], pos.data() + off[m], val.data() + off[m], lmax[m], step == 0 ? ua[m] : ub[m]);
      ch.chunk_picks++; ch.chunk_mismatch += cpick != pick;
      if (step == 0) { changed += pick != ch.a[m]; ch.a[m] = pick; }
      else { changed += (N - pick) != ch.b[m]; ch.b[m] = N - pick; }
    }
  }
  ch.totals();
  return changed;
}

void sweep(Chain &ch) {
  sample_cd(ch);
  sample_ab(ch);
  pi2(ch, 1);
  for (int j = 0; j < 5; j++) { pi1(ch); pi2(ch, 0); pi3(ch); }
}

} // namespace

extern "C" {

void *emul_create_ex(int N, int M, const uint8_t *X, const uint8_t *hard, double c0, double cc0, double d0,
                     double dd0, double eps, int manycd);
void *emul_create(int N, int M, const uint8_t *X, const uint8_t *hard, double c0, double cc0, double d0,
                  double dd0, double eps) { return emul_create_ex(N, M, X, hard, c0, cc0, d0, dd0, eps, 0); }
void *emul_create_ex(int N, int M, const uint8_t *X, const uint8_t *hard, double c0, double cc0, double d0,
                     double dd0, double eps, int manycd) {
  Chain *ch = new Chain();
  ch->manycd = manycd != 0;
  ch->N = N; ch->M = M; ch->W = N / 32 + 1; ch->C = M + 1; ch->nh = 0;
  ch->X.assign(X, X + (size_t)N * M);
  ch->hard.assign(hard, hard + N);
  for (int n = 0; n < N; n++) ch->nh += hard[n] != 0;
  ch->V.assign((size_t)ch->W * ch->C, 0u);
  ch->a.assign(M, 0); ch->b.assign(M, 0); ch->ones.assign(M, 0);
  ch->pre.assign((size_t)(ch->W + 1) * ch->C, 0); ch->hp.assign(N + 1, 0);
  ch->rpi.resize(N);
  for (int n = 0; n < N; n++) ch->rpi[n] = (uint16_t)n;
  for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) ch->ones[m] += X[(size_t)n * M + m] != 0;
  ch->wt.eps = eps;
  ch->set_cd(c0, cc0, d0, dd0);
  if (ch->manycd) {
    ch->wts.resize(M);
    for (int m = 0; m < M; m++) { ser_set_weights_own(&ch->wts[m], c0, cc0, d0, dd0, N); ch->wts[m].eps = eps; }
  }
  ch->tape = nullptr; ch->tape_len = ch->cur = 0; ch->n_degenerate = 0;
  ch->build_columns();
  ch->initab();
  ch->totals();
  return ch;
}
void emul_free(void *p) { delete (Chain *)p; }
void emul_set_tape(void *p, const double *tape, size_t len) { Chain *ch = (Chain *)p; ch->tape = tape; ch->tape_len = len; ch->cur = 0; }

// mcmc_randomize (mcmc.c:477-578) in draws
void emul_randomize(void *p) {
  Chain &ch = *(Chain *)p;
  const int N = ch.N, nh = ch.nh;
  std::vector<int> pi(N);
  for (int n = 0; n < N; n++) pi[n] = n;
  if (nh == 0) {
    for (int i = N - 1; i > 0; i--) { int j = ser_draw_int(ch.next(), i + 1); std::swap(pi[i], pi[j]); }
    for (int n = 0; n < N; n++) ch.rpi[pi[n]] = (uint16_t)n;
    ch.build_columns(); // a/b deliberately kept (mcmc.c:486-494)
    ch.totals();
    return;
  }
  if (nh == N) return;
  std::vector<int> chosen, rest;
  for (int i = 0; i < N && (int)chosen.size() < nh; i++)
    if ((double)(N - i) * ch.next() < (double)(nh - (int)chosen.size())) chosen.push_back(i);
  { size_t j = 0; for (int i = 0; i < N; i++) { if (j < chosen.size() && chosen[j] == i) j++; else rest.push_back(i); } }
  for (int i = N - nh - 1; i > 0; i--) { int r = ser_draw_int(ch.next(), i + 1); std::swap(rest[i], rest[r]); }
  { size_t j = 0, k = 0; for (int i = 0; i < N; i++) pi[i] = ch.hard[i] ? chosen[j++] : rest[k++]; }
  for (int n = 0; n < N; n++) ch.rpi[pi[n]] = (uint16_t)n;
  ch.build_columns();
  ch.initab();
  ch.totals();
}

void emul_sweeps(void *p, int n) { for (int s = 0; s < n; s++) sweep(*(Chain *)p); }
int emul_step(void *p, int kind) {
  Chain &ch = *(Chain *)p;
  switch (kind) {
    case 10: sample_cd(ch); return 2; // c and d together (tape slots 0..5 of the sweep)
    case 12: return sample_ab(ch);
    case 13: return pi2(ch, 1);
    case 14: return pi1(ch);
    case 15: return pi2(ch, 0);
    case 16: return pi3(ch);
  }
  return -1;
}

void emul_get_state(void *p, int32_t *a, int32_t *b, int32_t *pi, int32_t *rpi, int32_t *t0, int32_t *f0,
                    int32_t *t1, int32_t *f1, int32_t tot[4], double cdl[3], long long *slots) {
  Chain &ch = *(Chain *)p;
  for (int m = 0; m < ch.M; m++) {
    a[m] = ch.a[m]; b[m] = ch.b[m];
    int x0, y0, x1, y1;
    ser_counts(ch.col(m), ch.pcol(m), ch.C, ch.N, ch.a[m], ch.b[m], ch.ones[m], &x0, &y0, &x1, &y1);
    t0[m] = x0; f0[m] = y0; t1[m] = x1; f1[m] = y1;
  }
  for (int n = 0; n < ch.N; n++) { rpi[n] = ch.rpi[n]; pi[ch.rpi[n]] = n; }
  tot[0] = ch.t0a; tot[1] = ch.f0a; tot[2] = ch.t1a; tot[3] = ch.f1a;
  cdl[0] = ch.WT(0).c; cdl[1] = ch.WT(0).d; cdl[2] = ch.loglik;
  *slots = (long long)ch.cur;
}
long long emul_degenerate(void *p) { return ((Chain *)p)->n_degenerate; }
void emul_chunk_stats(void *p, long long out[3]) { Chain *ch = (Chain *)p; out[0] = ch->chunk_picks; out[1] = ch->chunk_mismatch; out[2] = ch->chunk_finder_bad; }
void emul_get_cd(void *p, double *c, double *d) {
  Chain &ch = *(Chain *)p;
  for (int m = 0; m < ch.M; m++) { c[m] = ch.WT(m).c; d[m] = ch.WT(m).d; }
}


/* ---- the product's bit-reproducible math / samplers (csrc/ser_detmath.h, ser_chain_core.h), exported so that
 * tests can validate them INDEPENDENTLY of the oracle, which includes the same header (tests/test_detmath.py) */
void emul_detmath_log(const double *x, int n, double *out) { for (int i = 0; i < n; i++) out[i] = ser_log(x[i]); }
void emul_detmath_exp(const double *x, int n, double *out) { for (int i = 0; i < n; i++) out[i] = ser_exp(x[i]); }
void emul_exp_weight(const double *x, int n, double *out) { for (int i = 0; i < n; i++) out[i] = ser_exp_weight(x[i]); }
/* n Gamma(shape, 1) variates: the draw of chain `chain0 + i`, sweep 0, block 0 of the structured stream */
void emul_gamma(double shape, uint32_t seed, uint32_t chain0, int n, double *out)
{
  for (int i = 0; i < n; i++) out[i] = ser_gamma_ge1(shape, seed, chain0 + (uint32_t)i, 0u, 0u);
}
/* n Beta(a, b) variates exactly as the sweep forms them: two Gammas from blocks 0 and 1 */
void emul_beta(double a, double b, uint32_t seed, uint32_t chain0, int n, double *out)
{
  for (int i = 0; i < n; i++)
    out[i] = ser_beta_from_gammas(ser_gamma_ge1(a, seed, chain0 + (uint32_t)i, 0u, 0u), ser_gamma_ge1(b, seed, chain0 + (uint32_t)i, 0u, 1u));
}
void emul_uniform(uint32_t seed, uint32_t chain, uint32_t sweep, uint32_t block, int n, double *out)
{
  for (int i = 0; i < n; i++) out[i] = ser_stream_uniform(seed, chain, sweep, block, (uint32_t)i);
}

} // extern "C"

/* ---- coarse prefix tables (ser_pre_at<G>, the large-shape kernel's columns in global memory): a column with a per-word table
 * (G = 0) and one with a count per 2^G words go through the same random moves and range queries; every answer and every word must
 * agree, and the coarse table must equal a rebuild from the words.  Returns the number of disagreements. */
template <int G>
static int coarse_case(uint64_t seed, int N, int ops)
{
  const int W = (N + 31) / 32;
  std::vector<uint32_t> c0(W + 1, 0u), cg(W + 1, 0u);
  std::vector<uint16_t> p0(W + 2, 0), pg((W >> G) + 2, 0), chk((W >> G) + 2, 0), perm(N);
  uint64_t st = seed * 0x9E3779B97F4A7C15ull + 12345u;
  auto rnd = [&](int n) { st = st * 6364136223846793005ull + 1442695040888963407ull; return (int)((st >> 33) % (uint64_t)n); };
  const int dens = 1 + rnd(9);
  for (int q = 0; q < N; q++) if (rnd(10) < dens) { c0[q >> 5] |= 1u << (q & 31); cg[q >> 5] |= 1u << (q & 31); }
  ser_col_build_pre<0>(c0.data(), p0.data(), 1, W);
  ser_col_build_pre<G>(cg.data(), pg.data(), 1, W);
  int bad = 0;
  for (int op = 0; op < ops; op++) {
    int i = rnd(N), j = rnd(N);
    const int kind = rnd(4);
    if (kind == 0 && i != j) { ser_col_rotate<0>(c0.data(), 1, W, i, j, p0.data()); ser_col_rotate<G>(cg.data(), 1, W, i, j, pg.data()); }
    else {
      if (i > j) std::swap(i, j);
      if (kind == 1 || i == j) { ser_col_reverse<0>(c0.data(), 1, W, i, j, p0.data()); ser_col_reverse<G>(cg.data(), 1, W, i, j, pg.data()); }
      else if (kind == 2) { /* an involution on the window: reversal of the odd offsets, the even ones stay */
        for (int q = i; q <= j; q++) perm[q] = (uint16_t)q;
        for (int lo = i + 1, hi = j - ((j - i) % 2 == 0 ? 1 : 0); lo < hi; lo += 2, hi -= 2) { perm[lo] = (uint16_t)hi; perm[hi] = (uint16_t)lo; }
        ser_col_permute<0>(c0.data(), 1, W, i, j, perm.data(), p0.data()); ser_col_permute<G>(cg.data(), 1, W, i, j, perm.data(), pg.data());
      } else if (j == i + 1) { ser_col_reverse<0>(c0.data(), 1, W, i, j, p0.data()); ser_col_reverse<G>(cg.data(), 1, W, i, j, pg.data()); }
    }
    for (int w = 0; w < W; w++) bad += c0[w] != cg[w];
    ser_col_build_pre<G>(cg.data(), chk.data(), 1, W);
    for (int k = 0; k <= (W >> G); k++) bad += chk[k] != pg[k];
    for (int t = 0; t < 8; t++) {
      int lo = rnd(N + 1), hi = rnd(N + 1);
      bad += ser_rank1<0>(c0.data(), p0.data(), 1, lo) != ser_rank1<G>(cg.data(), pg.data(), 1, lo);
      bad += ser_col_popc<0>(c0.data(), p0.data(), 1, lo, hi) != ser_col_popc<G>(cg.data(), pg.data(), 1, lo, hi);
    }
    int a = rnd(N + 1), b = rnd(N + 1);
    if (a > b) std::swap(a, b);
    const SerStep s0 = ser_step_a<0>(c0.data(), p0.data(), 1, W, N, a, b), sg = ser_step_a<G>(cg.data(), pg.data(), 1, W, N, a, b);
    const SerStep t0 = ser_step_b<0>(c0.data(), p0.data(), 1, W, N, a, b), tg = ser_step_b<G>(cg.data(), pg.data(), 1, W, N, a, b);
    bad += s0.ocur != sg.ocur || s0.kb != sg.kb || s0.nones != sg.nones || t0.ocur != tg.ocur || t0.kb != tg.kb || t0.nones != tg.nones;
    if (i < j) {
      int x0, x1, y0, y1;
      ser_pi2_delta<0>(c0.data(), p0.data(), 1, a, b, i, j, rnd(2), rnd(2), &x0, &x1);
      st -= 0; /* same increments for both: re-draw deterministically below */
      const int i1 = rnd(2), i2 = rnd(2);
      ser_pi2_delta<0>(c0.data(), p0.data(), 1, a, b, i, j, i1, i2, &x0, &x1);
      ser_pi2_delta<G>(cg.data(), pg.data(), 1, a, b, i, j, i1, i2, &y0, &y1);
      bad += x0 != y0 || x1 != y1;
    }
  }
  return bad;
}

extern "C" int emul_coarse_pre_selftest(uint64_t seed, int N, int ops)
{
  return coarse_case<1>(seed, N, ops) + coarse_case<2>(seed + 1, N, ops) + coarse_case<3>(seed + 2, N, ops);
}
