/* Minimal GSL-API shim (test infrastructure; see oracle/README.md).
 * Only what /root/reference/C_Implementation/mcmc.c uses. */
#ifndef SHIM_GSL_MATH_H
#define SHIM_GSL_MATH_H
#include <math.h>
#include <stddef.h>
#define GSL_MAX(a, b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a, b) ((a) < (b) ? (a) : (b))
#endif
