/* oracle/shims/gsl/gsl_multifit.h -- stand-in for the GSL calls behind the reference's
 * 4-point / 3-coefficient least-squares fit (src/polycap-profile.c:24-62).
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_GSL_MULTIFIT_H
#define ORACLE_SHIM_GSL_MULTIFIT_H
#include <stddef.h>

typedef struct { size_t size1, size2; double *data; } gsl_matrix;
typedef struct { size_t size; double *data; } gsl_vector;
typedef struct { size_t n, p; } gsl_multifit_linear_workspace;

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x);
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j);
gsl_vector *gsl_vector_alloc(size_t n);
void gsl_vector_free(gsl_vector *v);
void gsl_vector_set(gsl_vector *v, size_t i, double x);
double gsl_vector_get(const gsl_vector *v, size_t i);
gsl_multifit_linear_workspace *gsl_multifit_linear_alloc(size_t n, size_t p);
void gsl_multifit_linear_free(gsl_multifit_linear_workspace *w);
int gsl_multifit_linear(const gsl_matrix *X, const gsl_vector *y, gsl_vector *c, gsl_matrix *cov,
			double *chisq, gsl_multifit_linear_workspace *work);
#endif
