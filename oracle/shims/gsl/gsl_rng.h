/* oracle/shims/gsl/gsl_rng.h -- minimal stand-in for the part of GSL's RNG API the
 * reference uses (src/polycap-private.h:66-78): an MT19937 generator with GSL's
 * seeding rule and gsl_rng_uniform() = next 32-bit word / 2^32.
 * TEST INFRASTRUCTURE ONLY (used to build oracle/_ref). */
#ifndef ORACLE_SHIM_GSL_RNG_H
#define ORACLE_SHIM_GSL_RNG_H
#include <stdio.h>   /* the reference's private header relies on GSL pulling in FILE */
#include <stdlib.h>

typedef struct { const char *name; } gsl_rng_type;
typedef struct { unsigned long mt[624]; int mti; } gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(const gsl_rng *r, unsigned long int seed);
double gsl_rng_uniform(const gsl_rng *r);
#endif
