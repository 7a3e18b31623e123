/*
 * oracle/shims/shim_impl.c -- implementations behind the stand-in headers in this
 * directory (gsl/gsl_rng.h, gsl/gsl_multifit.h, xraylib.h).  Linked into oracle/_ref
 * together with the reference's own, unmodified hot-path sources.
 * TEST INFRASTRUCTURE ONLY: nothing in the product links or loads this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gsl/gsl_rng.h"
#include "gsl/gsl_multifit.h"
#include "xraylib.h"
#include "optconst.h" /* polycap_b200/csrc: built-in element model shared with the product */

/* ---------------- MT19937 with GSL's conventions ---------------- */

static const gsl_rng_type mt19937_type = {"mt19937"};
const gsl_rng_type *gsl_rng_mt19937 = &mt19937_type;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
	(void)T;
	gsl_rng *r = calloc(1, sizeof(gsl_rng));
	gsl_rng_set(r, 0);
	return r;
}

void gsl_rng_free(gsl_rng *r) { free(r); }

void gsl_rng_set(const gsl_rng *cr, unsigned long int s)
{
	gsl_rng *r = (gsl_rng *)cr;
	if (s == 0)
		s = 4357; /* GSL's default seed */
	r->mt[0] = s & 0xffffffffUL;
	for (int i = 1; i < 624; i++)
		r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
	r->mti = 624;
}

static unsigned long mt_next(gsl_rng *r)
{
	unsigned long *mt = r->mt;
	if (r->mti >= 624) {
		for (int k = 0; k < 624; k++) {
			unsigned long y = (mt[k] & 0x80000000UL) | (mt[(k + 1) % 624] & 0x7fffffffUL);
			mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
		}
		r->mti = 0;
	}
	unsigned long y = mt[r->mti++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680UL;
	y ^= (y << 15) & 0xefc60000UL;
	y ^= (y >> 18);
	return y & 0xffffffffUL;
}

double gsl_rng_uniform(const gsl_rng *r) { return mt_next((gsl_rng *)r) / 4294967296.0; }

/* ---------------- tiny dense least squares (Householder QR) ---------------- */

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2)
{
	gsl_matrix *m = malloc(sizeof *m);
	m->size1 = n1; m->size2 = n2; m->data = calloc(n1 * n2, sizeof(double));
	return m;
}
void gsl_matrix_free(gsl_matrix *m) { if (m) { free(m->data); free(m); } }
void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->size2 + j] = x; }
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->size2 + j]; }
gsl_vector *gsl_vector_alloc(size_t n)
{
	gsl_vector *v = malloc(sizeof *v);
	v->size = n; v->data = calloc(n, sizeof(double));
	return v;
}
void gsl_vector_free(gsl_vector *v) { if (v) { free(v->data); free(v); } }
void gsl_vector_set(gsl_vector *v, size_t i, double x) { v->data[i] = x; }
double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i]; }
gsl_multifit_linear_workspace *gsl_multifit_linear_alloc(size_t n, size_t p)
{
	gsl_multifit_linear_workspace *w = malloc(sizeof *w);
	w->n = n; w->p = p;
	return w;
}
void gsl_multifit_linear_free(gsl_multifit_linear_workspace *w) { free(w); }

int gsl_multifit_linear(const gsl_matrix *X, const gsl_vector *y, gsl_vector *c, gsl_matrix *cov,
			double *chisq, gsl_multifit_linear_workspace *work)
{
	(void)cov; (void)work;
	size_t n = X->size1, p = X->size2;
	double *A = malloc(n * p * sizeof(double)), *b = malloc(n * sizeof(double));
	memcpy(A, X->data, n * p * sizeof(double));
	memcpy(b, y->data, n * sizeof(double));
	for (size_t k = 0; k < p; k++) {
		double nrm = 0;
		for (size_t i = k; i < n; i++) nrm += A[i * p + k] * A[i * p + k];
		nrm = sqrt(nrm);
		if (nrm == 0) continue;
		double alpha = (A[k * p + k] > 0) ? -nrm : nrm;
		double *v = calloc(n, sizeof(double));
		for (size_t i = k; i < n; i++) v[i] = A[i * p + k];
		v[k] -= alpha;
		double vv = 0;
		for (size_t i = k; i < n; i++) vv += v[i] * v[i];
		if (vv > 0) {
			for (size_t j = k; j < p; j++) {
				double s = 0;
				for (size_t i = k; i < n; i++) s += v[i] * A[i * p + j];
				s = 2 * s / vv;
				for (size_t i = k; i < n; i++) A[i * p + j] -= s * v[i];
			}
			double s = 0;
			for (size_t i = k; i < n; i++) s += v[i] * b[i];
			s = 2 * s / vv;
			for (size_t i = k; i < n; i++) b[i] -= s * v[i];
		}
		free(v);
	}
	for (size_t kk = p; kk-- > 0;) {
		double s = b[kk];
		for (size_t j = kk + 1; j < p; j++) s -= A[kk * p + j] * c->data[j];
		c->data[kk] = s / A[kk * p + kk];
	}
	if (chisq) {
		double s = 0;
		for (size_t i = p; i < n; i++) s += b[i] * b[i];
		*chisq = s;
	}
	free(A); free(b);
	return 0;
}

/* ---------------- xraylib stand-in ---------------- */

double CS_Total(int Z, double E, xrl_error **error) { (void)error; return pcb_builtin_cs_total(Z, E); }
double Fi(int Z, double E, xrl_error **error) { (void)error; return pcb_builtin_fi(Z, E); }
double AtomicWeight(int Z, xrl_error **error) { (void)error; return pcb_builtin_atomic_weight(Z); }
