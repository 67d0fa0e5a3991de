/*
 * gs_numeric.h — the numeric contract shared by the CUDA path and the CPU oracle.
 *
 * Tile overlap lists and sort orderings must be bit-exact between the sm_100a
 * kernels and the oracle (BASELINE.json north_star).  libm's and CUDA's
 * logf/expf differ in the last ulp, so a borderline ellipse/tile test could
 * decide differently on the two sides.  This header therefore defines software
 * expf/logf built only from IEEE-754 binary32 add/mul/div and integer ops, which
 * are bit-reproducible on any conforming target PROVIDED the translation unit is
 * compiled without FMA contraction (nvcc: -fmad=false, g++: -ffp-contract=off).
 *
 * Algorithms: the classic fdlibm-style range reductions (exp: k*ln2 hi/lo split +
 * rational remainder; log: f/(2+f) series).  Accuracy < 1 ulp over the ranges the
 * projection / tile-mapper use (checked against libm in tests/test_numeric.py).
 *
 * Used by: csrc/geom_kernels.cu (projection, tile mapper) and oracle/oracle.cpp.
 * Reference math being restated: taichi_splatting/taichi_lib/generic.py:163-165
 * (sigmoid), perspective/projection.py:60 and taichi_lib/grid_query.py:76 (log).
 */
#ifndef GS_NUMERIC_H
#define GS_NUMERIC_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD static inline
#endif

GS_HD uint32_t gs_f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}

GS_HD float gs_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; memcpy(&x, &u, 4); return x;
#endif
}

/* exp(x) for finite or infinite x; NaN propagates. */
GS_HD float gs_expf(float x) {
  const float ln2_hi = 6.9314575195e-01f; /* 0x3f317200 */
  const float ln2_lo = 1.4286067653e-06f; /* 0x35bfbe8e */
  const float inv_ln2 = 1.4426950216e+00f;
  const float P1 = 1.6666625440e-1f;
  const float P2 = -2.7667332906e-3f;

  if (x != x) return x;
  if (x > 88.72168f) return gs_u2f(0x7f800000u);  /* +inf */
  if (x < -103.9720840f) return 0.0f;

  float hi = x, lo = 0.0f;
  int k = 0;
  uint32_t ax = gs_f2u(x) & 0x7fffffffu;
  if (ax > 0x3eb17218u) { /* |x| > 0.5 ln2 */
    float kf = inv_ln2 * x + (x < 0.0f ? -0.5f : 0.5f);
    k = (int)kf; /* truncation toward zero == round-to-nearest of x/ln2 */
    float t = (float)k;
    hi = x - t * ln2_hi; /* exact: ln2_hi has 8 trailing zero bits */
    lo = t * ln2_lo;
    x = hi - lo;
  } else if (ax < 0x39000000u) { /* |x| < 2^-14 */
    return 1.0f + x;
  }
  float t = x * x;
  float c = x - t * (P1 + t * P2);
  float y;
  if (k == 0) {
    y = 1.0f - ((x * c) / (c - 2.0f) - x);
    return y;
  }
  y = 1.0f - ((lo - (x * c) / (2.0f - c)) - hi);
  /* scale by 2^k in two safe steps (covers k in [-150, 128]) */
  if (k >= -125) {
    if (k == 128) return y * 2.0f * gs_u2f(0x7f000000u);
    return y * gs_u2f((uint32_t)(0x7f + k) << 23);
  }
  return y * gs_u2f((uint32_t)(0x7f + (k + 100)) << 23) * gs_u2f((uint32_t)(0x7f - 100) << 23);
}

/* log(x): x > 0 finite -> value; x == 0 -> -inf; x < 0 or NaN -> NaN; +inf -> +inf. */
GS_HD float gs_logf(float x) {
  const float ln2_hi = 6.9313812256e-01f; /* 0x3f317180 */
  const float ln2_lo = 9.0580006145e-06f; /* 0x3717f7d1 */
  const float Lg1 = 0.66666662693f;
  const float Lg2 = 0.40000972152f;
  const float Lg3 = 0.28498786688f;
  const float Lg4 = 0.24279078841f;

  uint32_t ix = gs_f2u(x);
  int k = 0;
  if (x != x) return x;
  if ((ix & 0x7fffffffu) == 0) return gs_u2f(0xff800000u); /* -inf */
  if (ix & 0x80000000u) return gs_u2f(0x7fc00000u);        /* NaN  */
  if (ix >= 0x7f800000u) return x;                          /* +inf */
  if (ix < 0x00800000u) { /* subnormal: scale up by 2^25 */
    x = x * 33554432.0f;
    ix = gs_f2u(x);
    k = -25;
  }
  k += (int)(ix >> 23) - 127;
  ix &= 0x007fffffu;
  uint32_t i = (ix + (0x95f64u << 3)) & 0x800000u;
  x = gs_u2f(ix | (i ^ 0x3f800000u)); /* normalise x to [sqrt(2)/2, sqrt(2)) */
  k += (int)(i >> 23);
  float f = x - 1.0f;
  float s = f / (2.0f + f);
  float dk = (float)k;
  float z = s * s;
  float w = z * z;
  float t1 = w * (Lg2 + w * Lg4);
  float t2 = z * (Lg1 + w * Lg3);
  float R = t2 + t1;
  float hfsq = 0.5f * f * f;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

/* sigmoid as the reference writes it: 1 / (1 + exp(-x))  (taichi_lib/generic.py:163-165) */
GS_HD float gs_sigmoidf(float x) { return 1.0f / (1.0f + gs_expf(-x)); }

#endif /* GS_NUMERIC_H */
