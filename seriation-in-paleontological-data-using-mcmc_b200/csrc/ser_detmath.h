/*
 * ser_detmath.h -- bit-reproducible fp64 math + counter-based RNG shared by the
 * CUDA sweep kernel (device) and the host side of the free-running draw stream.
 *
 * Why this exists: the sampler's *decisions* must be reproducible between the
 * GPU and a CPU checker.  libm / libdevice transcendentals differ by an ulp now
 * and then, so every value that feeds a decision in free-running (Philox) mode
 * -- log of the Beta variate (c, d), log(1-exp(c)), the Gamma/Normal samplers --
 * is built here from IEEE-754 basic operations only (+ - * / sqrt, each
 * correctly rounded and never contracted into an FMA), which are bit-identical
 * on x86-64 and sm_100a.  In replay mode none of this is used for c/d: the tape
 * carries the reference's own libm values (see DESIGN.md "tape grammar").
 *
 * Everything is `static inline` and compiles as C99, C++ and CUDA.
 * The polynomial kernels follow the classic Sun fdlibm formulations
 * (e_log.c / e_exp.c, freely redistributable) restated with explicit
 * non-contracted operations.
 */
#ifndef SER_DETMATH_H
#define SER_DETMATH_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SER_HD __host__ __device__ __forceinline__
#else
#define SER_HD static inline
#include <math.h>
#endif

/* ---- non-contracted primitive operations -------------------------------- */
#if defined(__CUDA_ARCH__)
#define SER_ADD(a, b) __dadd_rn((a), (b))
#define SER_SUB(a, b) __dadd_rn((a), -(b))
#define SER_MUL(a, b) __dmul_rn((a), (b))
#define SER_DIV(a, b) __ddiv_rn((a), (b))
#define SER_SQRT(a) __dsqrt_rn((a))
#else
/* host: compile with -ffp-contract=off; volatile-free because x86-64 SSE2
 * double arithmetic is already strict IEEE with that flag. */
#define SER_ADD(a, b) ((a) + (b))
#define SER_SUB(a, b) ((a) - (b))
#define SER_MUL(a, b) ((a) * (b))
#define SER_DIV(a, b) ((a) / (b))
#define SER_SQRT(a) sqrt((a))
#endif

SER_HD uint64_t ser_d2u(double x)
{
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
SER_HD double ser_u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}

/* ---- log(x), x finite > 0 (normal or subnormal); <1 ulp ------------------ */
SER_HD double ser_log(double x)
{
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  int k = 0;
  uint64_t ux = ser_d2u(x);
  if ((ux >> 52) == 0) { /* subnormal: scale up by 2^54 */
    x = SER_MUL(x, 18014398509481984.0);
    ux = ser_d2u(x);
    k = -54;
  }
  int32_t hx = (int32_t)(ux >> 32);
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  int32_t i = (hx + 0x95f64) & 0x100000; /* normalise m into [sqrt(1/2), sqrt(2)) */
  ux = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (ux & 0xffffffffu);
  k += (i >> 20);
  double m = ser_u2d(ux);
  double f = SER_SUB(m, 1.0);
  double s = SER_DIV(f, SER_ADD(2.0, f));
  double dk = (double)k;
  double z = SER_MUL(s, s);
  double w = SER_MUL(z, z);
  double t1 = SER_MUL(w, SER_ADD(Lg2, SER_MUL(w, SER_ADD(Lg4, SER_MUL(w, Lg6)))));
  double t2 = SER_MUL(z, SER_ADD(Lg1, SER_MUL(w, SER_ADD(Lg3, SER_MUL(w, SER_ADD(Lg5, SER_MUL(w, Lg7)))))));
  double R = SER_ADD(t2, t1);
  double hfsq = SER_MUL(SER_MUL(0.5, f), f);
  /* dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f) */
  double inner = SER_ADD(SER_MUL(s, SER_ADD(hfsq, R)), SER_MUL(dk, ln2_lo));
  return SER_SUB(SER_MUL(dk, ln2_hi), SER_SUB(SER_SUB(hfsq, inner), f));
}

/* ---- exp(x) for -700 <= x <= 700 (no overflow/underflow handling); <1 ulp - */
SER_HD double ser_exp(double x)
{
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
               P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
               P5 = 4.13813679705723846039e-08;
  int k = (int)SER_ADD(SER_MUL(invln2, x), (x < 0.0 ? -0.5 : 0.5));
  double t = (double)k;
  double hi = SER_SUB(x, SER_MUL(t, ln2HI));
  double lo = SER_MUL(t, ln2LO);
  double r = SER_SUB(hi, lo);
  double tt = SER_MUL(r, r);
  double c = SER_SUB(r, SER_MUL(tt, SER_ADD(P1, SER_MUL(tt, SER_ADD(P2, SER_MUL(tt, SER_ADD(P3, SER_MUL(tt, SER_ADD(P4, SER_MUL(tt, P5))))))))));
  double y = SER_SUB(1.0, SER_SUB(SER_SUB(lo, SER_DIV(SER_MUL(r, c), SER_SUB(2.0, c))), hi));
  /* scale by 2^k through the exponent field (|k| < 1021 guaranteed by range) */
  uint64_t uy = ser_d2u(y);
  uy += ((uint64_t)(int64_t)k) << 52;
  return ser_u2d(uy);
}

/* ---- Philox4x32-10 (Salmon et al. 2011) ---------------------------------- */
SER_HD void ser_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* two 32-bit words -> uniform double in [0,1) with 53 random bits */
SER_HD double ser_u53(uint32_t hi, uint32_t lo)
{
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * 1.1102230246251565404e-16; /* exact: v < 2^53, times 2^-53 */
}

/*
 * Structured free-running stream.  A draw is addressed by
 *   key   = (seed, global chain id)
 *   ctr   = (index>>1, block, sweep, 0)  and the low bit of index picks the
 *           (w0,w1) or (w2,w3) half of the Philox output.
 * Blocks inside one sweep (DESIGN.md "free-running stream"):
 */
#define SER_BLK_C_GAMMA_A 0u /* Beta for c: Gamma(1+f1a) */
#define SER_BLK_C_GAMMA_B 1u /* Beta for c: Gamma(1+t0a) */
#define SER_BLK_D_GAMMA_A 2u /* Beta for d: Gamma(1+f0a) */
#define SER_BLK_D_GAMMA_B 3u /* Beta for d: Gamma(1+t1a) */
#define SER_BLK_AB 4u        /* index 2m -> U for a_m, 2m+1 -> U for b_m */
#define SER_BLK_PI 5u        /* sequential draws of the 16 pi proposals */
#define SER_BLK_INIT 6u      /* mcmc_randomize draws; sweep field = 0xFFFFFFFF */
#define SER_BLK_MANYCD 8u     /* manycd: taxon m uses blocks 8+4m .. 8+4m+3 (c: Gamma a, b; d: Gamma a, b) */
#define SER_SWEEP_INIT 0xFFFFFFFFu

SER_HD double ser_stream_uniform(uint32_t seed, uint32_t chain, uint32_t sweep, uint32_t block,
                                 uint32_t index)
{
  uint32_t o[4];
  ser_philox4x32_10(index >> 1, block, sweep, 0u, seed, chain, o);
  return (index & 1u) ? ser_u53(o[2], o[3]) : ser_u53(o[0], o[1]);
}

/* strictly positive variant (GSL's uniform_pos contract) */
SER_HD double ser_pos(double u) { return u == 0.0 ? 1.1102230246251565404e-16 : u; }

/*
 * Gamma(shape, 1), shape >= 1: Marsaglia & Tsang (2000) squeeze method with
 * Marsaglia polar normals.  Draws are taken sequentially from one block of the
 * structured stream; everything is built from SER_* primitives so host and
 * device return the same bits.
 */
SER_HD double ser_gamma_ge1(double shape, uint32_t seed, uint32_t chain, uint32_t sweep,
                            uint32_t block)
{
  uint32_t idx = 0;
  const double d = SER_SUB(shape, 1.0 / 3.0);
  const double c = SER_DIV(1.0 / 3.0, SER_SQRT(d));
  for (;;) {
    double x, v;
    do {
      double a, b, s;
      do { /* polar method: one normal per accepted pair */
        a = SER_SUB(SER_MUL(2.0, ser_stream_uniform(seed, chain, sweep, block, idx)), 1.0);
        b = SER_SUB(SER_MUL(2.0, ser_stream_uniform(seed, chain, sweep, block, idx + 1)), 1.0);
        idx += 2;
        s = SER_ADD(SER_MUL(a, a), SER_MUL(b, b));
      } while (s >= 1.0 || s == 0.0);
      x = SER_MUL(a, SER_SQRT(SER_DIV(SER_MUL(-2.0, ser_log(s)), s)));
      v = SER_ADD(1.0, SER_MUL(c, x));
    } while (v <= 0.0);
    v = SER_MUL(SER_MUL(v, v), v);
    double u = ser_pos(ser_stream_uniform(seed, chain, sweep, block, idx));
    idx += 1;
    double x2 = SER_MUL(x, x);
    if (u < SER_SUB(1.0, SER_MUL(0.0331, SER_MUL(x2, x2)))) return SER_MUL(d, v);
    /* log(u) < 0.5 x^2 + d (1 - v + log v) */
    double rhs = SER_ADD(SER_MUL(0.5, x2), SER_MUL(d, SER_ADD(SER_SUB(1.0, v), ser_log(v))));
    if (ser_log(u) < rhs) return SER_MUL(d, v);
  }
}

/* Beta(a, b) = Ga / (Ga + Gb); the two Gammas use blocks `blk` and `blk+1`. */
SER_HD double ser_beta_from_gammas(double ga, double gb) { return SER_DIV(ga, SER_ADD(ga, gb)); }

#endif /* SER_DETMATH_H */
