/* Minimal GSL-API shim: the three gsl_ran_* entry points mcmc.c calls. */
#ifndef SHIM_GSL_RANDIST_H
#define SHIM_GSL_RANDIST_H
#include <stddef.h>
#include "gsl_rng.h"

double gsl_ran_beta(const gsl_rng *r, const double a, const double b);
void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t nmembm, size_t size);
void *gsl_ran_choose(const gsl_rng *r, void *dest, size_t k, void *src, size_t n, size_t size);
#endif
