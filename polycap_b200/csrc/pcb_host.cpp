/*
 * pcb_host.cpp -- host-side objects of the polycap C API: errors, rng, profile, description,
 * source constructors and the .inp parser.  Pure setup code (runs once, O(1e3) flops); the
 * photon transport itself lives in pcb_trace.cuh and is driven from pcb_api.cpp.
 *
 * Behaviour (argument checks, error codes and messages, ownership) follows the reference
 * files cited at each function; the code is a restatement, not a copy.
 */
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cinttypes>
#include <ctime>

#include "pcb_internal.h"
#include "optconst.h"

#define FAIL(err, code, ...)                          \
	do {                                          \
		polycap_set_error(err, code, __VA_ARGS__); \
	} while (0)

/* ------------------------------------------------------------------ errors (src/polycap-error.c) */

extern "C" polycap_error *polycap_error_new_valist(enum polycap_error_code code, const char *format, va_list args)
{
	polycap_error *e = (polycap_error *)malloc(sizeof(polycap_error));
	if (!e)
		return NULL;
	e->code = code;
	e->message = NULL;
	if (vasprintf(&e->message, format, args) < 0)
		e->message = NULL;
	return e;
}

extern "C" polycap_error *polycap_error_new(enum polycap_error_code code, const char *format, ...)
{
	if (!format)
		return NULL;
	va_list ap;
	va_start(ap, format);
	polycap_error *e = polycap_error_new_valist(code, format, ap);
	va_end(ap);
	return e;
}

extern "C" polycap_error *polycap_error_new_literal(enum polycap_error_code code, const char *message)
{
	if (!message)
		return NULL;
	polycap_error *e = (polycap_error *)malloc(sizeof(polycap_error));
	if (!e)
		return NULL;
	e->code = code;
	e->message = strdup(message);
	return e;
}

extern "C" void polycap_error_free(polycap_error *error)
{
	if (!error)
		return;
	free(error->message);
	free(error);
}

extern "C" polycap_error *polycap_error_copy(const polycap_error *error)
{
	if (!error)
		return NULL;
	polycap_error *c = (polycap_error *)malloc(sizeof(polycap_error));
	*c = *error;
	c->message = error->message ? strdup(error->message) : NULL;
	return c;
}

extern "C" bool polycap_error_matches(const polycap_error *error, enum polycap_error_code code) { return error && error->code == code; }

static const char *OVERWRITE_WARNING =
	"polycap_error set over the top of a previous polycap_error or uninitialized memory.\n"
	"This indicates a bug in someone's code. You must ensure an error is NULL before it's set.\n"
	"The overwriting error message was: %s\n";

extern "C" void polycap_set_error(polycap_error **err, enum polycap_error_code code, const char *format, ...)
{
	if (err == NULL)
		return;
	va_list ap;
	va_start(ap, format);
	polycap_error *e = polycap_error_new_valist(code, format, ap);
	va_end(ap);
	if (*err == NULL) {
		*err = e;
	} else { /* error.c:125-153 */
		fprintf(stderr, OVERWRITE_WARNING, e && e->message ? e->message : "");
		polycap_error_free(e);
	}
}

extern "C" void polycap_set_error_literal(polycap_error **err, enum polycap_error_code code, const char *message)
{
	if (err == NULL)
		return;
	if (*err == NULL)
		*err = polycap_error_new_literal(code, message);
	else
		fprintf(stderr, OVERWRITE_WARNING, message);
}

extern "C" void polycap_propagate_error(polycap_error **dest, polycap_error *src)
{
	if (!src)
		return;
	if (dest == NULL) {
		polycap_error_free(src);
		return;
	}
	if (*dest != NULL) {
		fprintf(stderr, OVERWRITE_WARNING, src->message ? src->message : "");
		polycap_error_free(src);
	} else {
		*dest = src;
	}
}

extern "C" void polycap_clear_error(polycap_error **err)
{
	if (err && *err) {
		polycap_error_free(*err);
		*err = NULL;
	}
}

/* ------------------------------------------------------------------ aux (src/polycap-aux.c, transmission-efficiencies.c:1169) */

extern "C" char *polycap_strdup(const char *str) { return str ? strdup(str) : NULL; }
extern "C" char *polycap_strndup(const char *str, size_t len) { return str ? strndup(str, len) : NULL; }
extern "C" void polycap_free(void *data)
{
	if (data)
		free(data);
}

/* ------------------------------------------------------------------ rng (src/polycap-rng.c) */

extern "C" void pcb_mt_seed(struct pcb_mt19937 *r, unsigned long s)
{
	if (s == 0)
		s = 4357;
	r->mt[0] = (uint32_t)(s & 0xffffffffUL);
	for (int i = 1; i < 624; i++)
		r->mt[i] = 1812433253U * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
	r->mti = 624;
}

extern "C" double pcb_mt_uniform(struct pcb_mt19937 *r)
{
	if (r->mti >= 624) {
		for (int k = 0; k < 624; k++) {
			uint32_t y = (r->mt[k] & 0x80000000U) | (r->mt[(k + 1) % 624] & 0x7fffffffU);
			r->mt[k] = r->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
		}
		r->mti = 0;
	}
	uint32_t y = r->mt[r->mti++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680U;
	y ^= (y << 15) & 0xefc60000U;
	y ^= (y >> 18);
	return y / 4294967296.0;
}

extern "C" polycap_rng *polycap_rng_new_with_seed(unsigned long int seed)
{
	polycap_rng *rng = (polycap_rng *)calloc(1, sizeof(polycap_rng));
	rng->_rng = (struct pcb_mt19937 *)malloc(sizeof(struct pcb_mt19937));
	pcb_mt_seed(rng->_rng, seed);
	return rng;
}

extern "C" polycap_rng *polycap_rng_new(void) /* rng.c:50-71: seed from /dev/urandom, else time */
{
	unsigned long seed = 0;
	FILE *fp = fopen("/dev/urandom", "r");
	if (fp && fread(&seed, sizeof seed, 1, fp) == 1) {
		fclose(fp);
	} else {
		if (fp)
			fclose(fp);
		struct timespec ts;
		clock_gettime(CLOCK_REALTIME, &ts);
		seed = (unsigned long)ts.tv_sec ^ (unsigned long)ts.tv_nsec;
	}
	return polycap_rng_new_with_seed(seed);
}

extern "C" void polycap_rng_free(polycap_rng *rng)
{
	if (!rng)
		return;
	free(rng->_rng);
	free(rng);
}

extern "C" double polycap_rng_uniform(const polycap_rng *r) { return pcb_mt_uniform(r->_rng); }

/* ------------------------------------------------------------------ profile (src/polycap-profile.c) */

static polycap_profile *profile_alloc(int nmax, polycap_error **error, const char *who)
{
	polycap_profile *p = (polycap_profile *)malloc(sizeof(polycap_profile));
	if (!p) {
		FAIL(error, POLYCAP_ERROR_MEMORY, "%s: could not allocate memory for profile -> %s", who, strerror(errno));
		return NULL;
	}
	p->nmax = nmax;
	p->z = (double *)malloc(sizeof(double) * (size_t)(nmax + 1));
	p->cap = (double *)malloc(sizeof(double) * (size_t)(nmax + 1));
	p->ext = (double *)malloc(sizeof(double) * (size_t)(nmax + 1));
	if (!p->z || !p->cap || !p->ext) {
		FAIL(error, POLYCAP_ERROR_MEMORY, "%s: could not allocate memory for profile arrays -> %s", who, strerror(errno));
		polycap_profile_free(p);
		return NULL;
	}
	return p;
}

/* least-squares quadratic through 4 points (profile.c:24-62 uses gsl_multifit_linear);
 * Householder QR on the 4x3 Vandermonde system */
static void fit_quadratic4(const double x[4], const double y[4], double coeff[3])
{
	double A[4][3], b[4];
	for (int i = 0; i < 4; i++) {
		A[i][0] = 1.0;
		A[i][1] = x[i];
		A[i][2] = x[i] * x[i];
		b[i] = y[i];
	}
	for (int k = 0; k < 3; k++) {
		double nrm = 0;
		for (int i = k; i < 4; i++)
			nrm += A[i][k] * A[i][k];
		nrm = sqrt(nrm);
		if (nrm == 0)
			continue;
		double alpha = (A[k][k] > 0) ? -nrm : nrm;
		double v[4] = {0, 0, 0, 0};
		for (int i = k; i < 4; i++)
			v[i] = A[i][k];
		v[k] -= alpha;
		double vv = 0;
		for (int i = k; i < 4; i++)
			vv += v[i] * v[i];
		if (vv == 0)
			continue;
		for (int j = k; j < 3; j++) {
			double s = 0;
			for (int i = k; i < 4; i++)
				s += v[i] * A[i][j];
			s = 2 * s / vv;
			for (int i = k; i < 4; i++)
				A[i][j] -= s * v[i];
		}
		double s = 0;
		for (int i = k; i < 4; i++)
			s += v[i] * b[i];
		s = 2 * s / vv;
		for (int i = k; i < 4; i++)
			b[i] -= s * v[i];
	}
	for (int k = 2; k >= 0; k--) {
		double s = b[k];
		for (int j = k + 1; j < 3; j++)
			s -= A[k][j] * coeff[j];
		coeff[k] = s / A[k][k];
	}
}

extern "C" polycap_profile *polycap_profile_new(polycap_profile_type type, double length, double rad_ext_upstream, double rad_ext_downstream,
						double rad_int_upstream, double rad_int_downstream, double focal_dist_upstream,
						double focal_dist_downstream, polycap_error **error) /* profile.c:66-207 */
{
	const int nmax = 999;
	if (length <= 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: length must be greater than 0.0"); return NULL; }
	if (rad_ext_upstream <= 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: rad_ext_upstream must be greater than 0.0"); return NULL; }
	if (rad_ext_downstream <= 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: rad_ext_downstream must be greater than 0.0"); return NULL; }
	if (rad_int_upstream <= 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: rad_int_upstream must be greater than 0.0"); return NULL; }
	if (rad_int_downstream <= 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: rad_int_downstream must be greater than 0.0"); return NULL; }
	if (rad_int_upstream >= rad_ext_upstream) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: rad_ext_upstream must be greater than rad_int_upstream"); return NULL; }
	if (rad_int_downstream >= rad_ext_downstream) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: rad_ext_downstream must be greater than rad_int_downstream"); return NULL; }
	if (focal_dist_upstream <= 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: focal_dist_upstream must be greater than 0.0"); return NULL; }
	if (focal_dist_downstream <= 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: focal_dist_downstream must be greater than 0.0"); return NULL; }
	if (type != POLYCAP_PROFILE_CONICAL && type != POLYCAP_PROFILE_PARABOLOIDAL && type != POLYCAP_PROFILE_ELLIPSOIDAL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new: invalid profile type detected");
		return NULL;
	}
	polycap_profile *p = profile_alloc(nmax, error, "polycap_profile_new");
	if (!p)
		return NULL;
	/* the bore is conical for every type; z runs 0..length in nmax equal steps */
	for (int i = 0; i <= nmax; i++) {
		p->z[i] = length / nmax * i;
		p->cap[i] = (rad_int_downstream - rad_int_upstream) / length * p->z[i] + rad_int_upstream;
	}
	if (type == POLYCAP_PROFILE_CONICAL) {
		for (int i = 0; i <= nmax; i++)
			p->ext[i] = (rad_ext_downstream - rad_ext_upstream) / length * p->z[i] + rad_ext_upstream;
	} else if (type == POLYCAP_PROFILE_PARABOLOIDAL) { /* profile.c:149-169 */
		double px[4], py[4], coeff[3];
		px[0] = 0.;
		py[0] = rad_ext_upstream;
		px[3] = length;
		py[3] = rad_ext_downstream;
		px[1] = (focal_dist_upstream <= length) ? focal_dist_upstream / 10. : length / 10.;
		py[1] = (rad_ext_upstream - 0.) / (0. - (-1. * focal_dist_upstream)) * (px[1] - 0.) + rad_ext_upstream;
		px[2] = (focal_dist_downstream <= length) ? length - focal_dist_downstream / 10. : length - length / 10.;
		py[2] = (rad_ext_downstream - 0.) / (length - (length + focal_dist_downstream)) * (px[2] - length) + rad_ext_downstream;
		fit_quadratic4(px, py, coeff);
		for (int i = 0; i <= nmax; i++)
			p->ext[i] = coeff[0] + coeff[1] * p->z[i] + coeff[2] * p->z[i] * p->z[i];
	} else { /* ellipsoidal, profile.c:171-195 */
		if (rad_ext_downstream < rad_ext_upstream) {
			double slope = rad_ext_downstream / focal_dist_downstream;
			double dr = (rad_ext_downstream - rad_ext_upstream);
			double b = (-1. * dr * dr - slope * length * dr) / (slope * length + 2. * dr);
			double k = rad_ext_upstream - b;
			double a = sqrt((b * b * length) / (slope * (rad_ext_downstream - k)));
			for (int i = 0; i <= nmax; i++)
				p->ext[i] = sqrt(b * b - (b * b * p->z[i] * p->z[i]) / (a * a)) + k;
		} else {
			double slope = rad_ext_upstream / focal_dist_upstream;
			double dr = (rad_ext_upstream - rad_ext_downstream);
			double b = (-1. * dr * dr - slope * length * dr) / (slope * length + 2. * dr);
			double k = rad_ext_downstream - b;
			double a = sqrt(fabs((b * b * length) / (slope * (rad_ext_upstream - k))));
			for (int i = 0; i <= nmax; i++)
				p->ext[i] = sqrt(b * b - (b * b * p->z[nmax - i] * p->z[nmax - i]) / (a * a)) + k;
		}
	}
	return p;
}

static FILE *open_profile_file(const char *name, polycap_error **error)
{
	FILE *fp = fopen(name, "r");
	if (!fp)
		FAIL(error, POLYCAP_ERROR_IO, "polycap_profile_new_from_file: could not open %s -> %s", name, strerror(errno));
	return fp;
}

extern "C" polycap_profile *polycap_profile_new_from_file(const char *single_cap_profile_file, const char *central_axis_file,
							  const char *external_shape_file, polycap_error **error) /* profile.c:211-317 */
{
	if (!single_cap_profile_file) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new_from_file: single_cap_profile_file cannot be NULL"); return NULL; }
	if (!central_axis_file) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new_from_file: central_axis_file cannot be NULL"); return NULL; }
	if (!external_shape_file) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new_from_file: external_shape_file cannot be NULL"); return NULL; }

	FILE *fp = open_profile_file(single_cap_profile_file, error);
	if (!fp)
		return NULL;
	int n = 0;
	if (fscanf(fp, "%d", &n) != 1)
		n = 0;
	if (n <= 100) {
		fclose(fp);
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new_from_file: n_tmp must be greater than 100");
		return NULL;
	}
	polycap_profile *p = profile_alloc(n, error, "polycap_profile_new_from_file");
	if (!p) {
		fclose(fp);
		return NULL;
	}
	for (int i = 0; i <= n; i++)
		if (fscanf(fp, "%lf %lf", &p->z[i], &p->cap[i]) != 2)
			break;
	fclose(fp);

	/* the central-axis file only contributes its z column; x/y are read and dropped (profile.c:294) */
	fp = open_profile_file(central_axis_file, error);
	if (!fp) {
		polycap_profile_free(p);
		return NULL;
	}
	int n2 = 0;
	if (fscanf(fp, "%d", &n2) != 1 || n2 != n) {
		fclose(fp);
		FAIL(error, POLYCAP_ERROR_IO, "polycap_profile_new_from_file: Number of intervals inconsistent: %s", central_axis_file);
		polycap_profile_free(p);
		return NULL;
	}
	for (int i = 0; i <= n; i++) {
		double sx, sy;
		if (fscanf(fp, "%lf %lf %lf", &p->z[i], &sx, &sy) != 3)
			break;
	}
	fclose(fp);

	fp = open_profile_file(external_shape_file, error);
	if (!fp) {
		polycap_profile_free(p);
		return NULL;
	}
	if (fscanf(fp, "%d", &n2) != 1 || n2 != n) {
		fclose(fp);
		FAIL(error, POLYCAP_ERROR_IO, "polycap_profile_new_from_file: Number of intervals inconsistent: %s", external_shape_file);
		polycap_profile_free(p);
		return NULL;
	}
	for (int i = 0; i <= n; i++)
		if (fscanf(fp, "%lf %lf", &p->z[i], &p->ext[i]) != 2)
			break;
	fclose(fp);
	return p;
}

static int host_within_hexagon(double R, double x, double y) /* photon.c:139-169, setup-time use only */
{
	if (R <= 0.)
		return -1;
	double apothem = sqrt((R * R) - ((R / 2.) * (R / 2.)));
	double d1 = fabs(0 * x + 1 * y), d2 = fabs(COSPI_6 * x + 0.5 * y), d3 = fabs(COSPI_6 * x + -0.5 * y);
	return (d1 > apothem || d2 > apothem || d3 > apothem) ? 0 : 1;
}

extern "C" int polycap_profile_validate(polycap_profile *profile, int64_t n_cap, polycap_error **error) /* profile.c:321-423 */
{
	if (!profile) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_validate: profile cannot be NULL");
		return -1;
	}
	double n_shells = round(sqrt(12. * n_cap - 3.) / 6. - 0.5);
	if (n_shells == 0) {
		for (int i = 0; i <= profile->nmax; i++)
			if (profile->cap[i] >= profile->ext[i])
				return 0;
		return 1;
	}
	/* walk the outer shell: every capillary's outermost point must stay inside the hexagon at every z */
	static const int q_dir[6] = {1, 1, 0, -1, -1, 0};
	static const int r_dir[6] = {0, -1, -1, 0, 1, 1};
	double q = -1. * n_shells, r = n_shells;
	for (int side = 0; side < 6; side++) {
		for (int k = 0; k < n_shells; k++) {
			q += q_dir[side];
			r += r_dir[side];
			for (int i = 0; i <= profile->nmax; i++) {
				double hexz = profile->ext[i] / (2. * COSPI_6 * (n_shells + 1));
				double cy = r * (3. / 2) * hexz;
				double cx = (2 * q + r) * COSPI_6 * hexz;
				double angle = atan(cy / cx);
				cx += cos(angle) * profile->cap[i];
				cy += sin(angle) * profile->cap[i];
				int chk = host_within_hexagon(profile->ext[i], cx, cy);
				if (chk == 0)
					return 0;
				if (chk == -1) {
					polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_within_pc_boundary: polycap_radius must be greater than 0");
					return -1;
				}
			}
		}
	}
	return 1;
}

extern "C" polycap_profile *polycap_profile_new_from_arrays(int nid, double *ext, double *cap, double *z, polycap_error **error) /* profile.c:426-474 */
{
	if (!ext) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new_from_array: ext cannot be NULL"); return NULL; }
	if (!cap) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new_from_array: cap cannot be NULL"); return NULL; }
	if (!z) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new_from_array: z cannot be NULL"); return NULL; }
	if (nid <= 1) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_profile_new_from_array: nid must be greater than 1"); return NULL; }
	polycap_profile *p = profile_alloc(nid, error, "polycap_profile_new_from_array");
	if (!p)
		return NULL;
	memcpy(p->ext, ext, sizeof(double) * (size_t)(nid + 1));
	memcpy(p->cap, cap, sizeof(double) * (size_t)(nid + 1));
	memcpy(p->z, z, sizeof(double) * (size_t)(nid + 1));
	return p;
}

static bool profile_get(polycap_profile *profile, size_t *nid, double **out, const double *src)
{
	if (!profile || !nid)
		return false;
	*nid = (size_t)profile->nmax;
	*out = (double *)malloc(sizeof(double) * (size_t)(profile->nmax + 1));
	memcpy(*out, src, sizeof(double) * (size_t)(profile->nmax + 1));
	return true;
}
extern "C" bool polycap_profile_get_ext(polycap_profile *p, size_t *nid, double **ext, polycap_error **) { return profile_get(p, nid, ext, p ? p->ext : NULL); }
extern "C" bool polycap_profile_get_cap(polycap_profile *p, size_t *nid, double **cap, polycap_error **) { return profile_get(p, nid, cap, p ? p->cap : NULL); }
extern "C" bool polycap_profile_get_z(polycap_profile *p, size_t *nid, double **z, polycap_error **) { return profile_get(p, nid, z, p ? p->z : NULL); }

extern "C" void polycap_profile_free(polycap_profile *profile)
{
	if (!profile)
		return;
	free(profile->z);
	free(profile->cap);
	free(profile->ext);
	free(profile);
}

/* ------------------------------------------------------------------ description (src/polycap-description.c) */

extern "C" char *polycap_read_input_line(FILE *fptr, polycap_error **error) /* description.c:24-56 */
{
	if (!fptr) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_read_input_line: fptr cannot be NULL");
		return NULL;
	}
	size_t cap = 128, n = 0;
	char *s = (char *)malloc(cap);
	int ch;
	while ((ch = fgetc(fptr)) != '\n' && ch != EOF) {
		s[n++] = (char)ch;
		if (n == cap) {
			cap += 128;
			s = (char *)realloc(s, cap);
		}
	}
	s[n++] = '\0';
	return (char *)realloc(s, n);
}

extern "C" void polycap_description_check_weight(size_t nelem, double wi[], polycap_error **error) /* description.c:58-85 */
{
	double sum = 0;
	for (size_t i = 0; i < nelem; i++)
		sum += wi[i];
	if (sum > 1.) { /* given as percentages */
		sum = 0;
		for (size_t i = 0; i < nelem; i++) {
			wi[i] /= 100.0;
			if (wi[i] < 0.0) {
				polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_check_weight: Polycapillary element weights must be greater than 0.0");
				return;
			}
			sum += wi[i];
		}
	}
	if (sum != 1.)
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_check_weight: Polycapillary element weights do not sum to 1.");
}

static double open_area_of(const polycap_description *d) /* description.c:214-216 */
{
	double n = (round(sqrt(12. * d->n_cap - 3.) / 6. - 0.5) + 0.5) * 6.;
	n = (n * n + 3) / 12;
	return (d->profile->cap[0] * d->profile->cap[0] * M_PI) * n / (3. * sin(M_PI / 3) * d->profile->ext[0] * d->profile->ext[0]);
}

extern "C" polycap_description *polycap_description_new(polycap_profile *profile, double sig_rough, int64_t n_cap, unsigned int nelem, int iz[],
							double wi[], double density, polycap_error **error) /* description.c:89-231 */
{
	if (sig_rough < 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_new: sig_rough must be greater than or equal to zero"); return NULL; }
	if (n_cap <= 1) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_new: n_cap must be greater than 1"); return NULL; }
	if (nelem < 1) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_new: nelem must be 1 or greater"); return NULL; }
	if (!iz) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_new: iz cannot be NULL"); return NULL; }
	if (!wi) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_new: wi cannot be NULL"); return NULL; }
	if (density <= 0.0) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_new: density must be greater than 0.0"); return NULL; }
	for (unsigned int i = 0; i < nelem; i++)
		if (iz[i] < 1 || iz[i] > 111) {
			polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_new: iz[i] must be greater than 0 and smaller than 111");
			return NULL;
		}
	if (!profile) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_description_new: profile cannot be NULL");
		return NULL;
	}
	polycap_description *d = (polycap_description *)calloc(1, sizeof(polycap_description));
	d->iz = (int *)malloc(sizeof(int) * nelem);
	d->wi = (double *)malloc(sizeof(double) * nelem);
	d->sig_rough = sig_rough;
	d->n_cap = n_cap;
	d->nelem = nelem;
	d->density = density;
	for (unsigned int i = 0; i < nelem; i++) {
		d->iz[i] = iz[i];
		d->wi[i] = wi[i];
	}
	polycap_description_check_weight(d->nelem, d->wi, error); /* may set an error; construction continues (description.c:162) */

	d->profile = polycap_profile_new_from_arrays(profile->nmax, profile->ext, profile->cap, profile->z, NULL); /* deep copy */
	if (!d->profile) {
		polycap_set_error_literal(error, POLYCAP_ERROR_MEMORY, "polycap_description_new: could not allocate memory for description->profile");
		free(d->iz);
		free(d->wi);
		free(d);
		return NULL;
	}
	d->open_area = open_area_of(d);
	if (polycap_profile_validate(d->profile, d->n_cap, error) != 1) {
		polycap_clear_error(error);
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT,
					  "polycap_description_new: description->profile is faulty. Some capillary coordinates are outside of the external radius.");
		polycap_description_free(d);
		return NULL;
	}
	return d;
}

extern "C" const polycap_profile *polycap_description_get_profile(polycap_description *description) { return description->profile; }

extern "C" void polycap_description_free(polycap_description *description)
{
	if (!description)
		return;
	polycap_profile_free(description->profile);
	free(description->iz);
	free(description->wi);
	free(description);
}

/* ------------------------------------------------------------------ source constructors (src/polycap-source.c:147-445, 1090-1109) */

extern "C" polycap_source *polycap_source_new(polycap_description *description, double d_source, double src_x, double src_y, double src_sigx,
					      double src_sigy, double src_shiftx, double src_shifty, double hor_pol, size_t n_energies,
					      double *energies, polycap_error **error)
{
	if (!description) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new: description cannot be NULL"); return NULL; }
	if (d_source <= 0.) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new: d_source must be greater than 0"); return NULL; }
	if (src_x <= 0.) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new: src_x must be greater than 0"); return NULL; }
	if (src_y <= 0.) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new: src_y must be greater than 0"); return NULL; }
	if (fabs(hor_pol) > 1.) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new: hor_pol must be greater than or equal to -1 and smaller than or equal to 1"); return NULL; }
	if (n_energies <= 0.) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new: n_energies must be greater than 0"); return NULL; }
	if (!energies) { polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new: energies cannot be NULL"); return NULL; }
	for (size_t i = 0; i < n_energies; i++)
		if (energies[i] < 1. || energies[i] > 100.) {
			polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new: energies must be greater than 1 and smaller than 100");
			return NULL;
		}
	if (polycap_profile_validate(description->profile, description->n_cap, error) != 1) {
		polycap_clear_error(error);
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT,
					  "polycap_source_new: description->profile is faulty. Some capillary coordinates are outside of the external radius.");
		return NULL;
	}
	polycap_source *s = (polycap_source *)calloc(1, sizeof(polycap_source));
	s->energies = (double *)malloc(sizeof(double) * n_energies);
	s->d_source = d_source;
	s->src_x = src_x;
	s->src_y = src_y;
	s->src_sigx = src_sigx;
	s->src_sigy = src_sigy;
	s->src_shiftx = src_shiftx;
	s->src_shifty = src_shifty;
	s->hor_pol = hor_pol;
	s->n_energies = n_energies;
	memcpy(s->energies, energies, sizeof(double) * n_energies);
	s->rng = polycap_rng_new();
	s->description = polycap_description_new(description->profile, description->sig_rough, description->n_cap, description->nelem, description->iz,
						 description->wi, description->density, NULL);
	if (!s->description) {
		polycap_clear_error(error);
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT,
					  "polycap_source_new: description->profile is faulty. Some capillary coordinates are outside of the external radius.");
		polycap_source_free(s);
		return NULL;
	}
	return s;
}

/* positional, whitespace separated .inp (source.c:273-368); profile file names resolve against the CWD */
extern "C" polycap_source *polycap_source_new_from_file(const char *filename, polycap_error **error)
{
	if (!filename) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_new_from_file: filename cannot be NULL");
		return NULL;
	}
	FILE *fp = fopen(filename, "r");
	if (!fp) {
		FAIL(error, POLYCAP_ERROR_IO, "polycap_source_new_from_file: could not open %s -> %s", filename, strerror(errno));
		return NULL;
	}
	polycap_description *d = (polycap_description *)calloc(1, sizeof(polycap_description));
	polycap_source *s = (polycap_source *)calloc(1, sizeof(polycap_source));
	s->description = d;
	s->rng = polycap_rng_new();

	bool ok = true;
	ok = ok && fscanf(fp, "%lf", &d->sig_rough) == 1;
	ok = ok && fscanf(fp, "%lf", &s->d_source) == 1;
	ok = ok && fscanf(fp, "%lf %lf", &s->src_x, &s->src_y) == 2;
	ok = ok && fscanf(fp, "%lf %lf", &s->src_sigx, &s->src_sigy) == 2;
	ok = ok && fscanf(fp, "%lf %lf", &s->src_shiftx, &s->src_shifty) == 2;
	ok = ok && fscanf(fp, "%lf", &s->hor_pol) == 1;
	ok = ok && fscanf(fp, "%u", &d->nelem) == 1 && d->nelem >= 1 && d->nelem < 1000;
	if (!ok) {
		fclose(fp);
		FAIL(error, POLYCAP_ERROR_IO, "polycap_source_new_from_file: could not read the source/description header of %s", filename);
		polycap_source_free(s);
		return NULL;
	}
	d->iz = (int *)calloc(d->nelem, sizeof(int));
	d->wi = (double *)calloc(d->nelem, sizeof(double));
	for (unsigned int i = 0; i < d->nelem; i++) {
		if (fscanf(fp, "%d %lf", &d->iz[i], &d->wi[i]) != 2)
			ok = false;
		d->wi[i] /= 100.0;
	}
	double e_start = 0, e_final = 0, delta_e = 1;
	int nphotons = 0, type = 0;
	ok = ok && fscanf(fp, "%lf", &d->density) == 1;
	ok = ok && fscanf(fp, "%lf %lf %lf", &e_start, &e_final, &delta_e) == 3;
	if (ok) {
		s->n_energies = (size_t)((e_final - e_start) / delta_e + 1); /* truncation as in source.c:334 */
		s->energies = (double *)malloc(sizeof(double) * (s->n_energies ? s->n_energies : 1));
		for (size_t i = 0; i < s->n_energies; i++)
			s->energies[i] = e_start + i * delta_e;
	}
	ok = ok && fscanf(fp, "%d", &nphotons) == 1; /* read and ignored, like the reference (source.c:343) */
	ok = ok && fscanf(fp, "%d", &type) == 1;
	if (!ok) {
		fclose(fp);
		FAIL(error, POLYCAP_ERROR_IO, "polycap_source_new_from_file: could not parse %s", filename);
		polycap_source_free(s);
		return NULL;
	}
	if (type == 0 || type == 1 || type == 2) {
		double len, reu, red, riu, rid, fu, fd;
		if (fscanf(fp, "%lf %lf %lf %lf %lf %lf %lf", &len, &reu, &red, &riu, &rid, &fu, &fd) != 7) {
			fclose(fp);
			FAIL(error, POLYCAP_ERROR_IO, "polycap_source_new_from_file: could not parse the profile line of %s", filename);
			polycap_source_free(s);
			return NULL;
		}
		d->profile = polycap_profile_new((polycap_profile_type)type, len, reu, red, riu, rid, fu, fd, error);
	} else {
		fgetc(fp); /* the newline that ends the type line */
		char *f_cap = polycap_read_input_line(fp, error);
		char *f_axs = polycap_read_input_line(fp, error);
		char *f_ext = polycap_read_input_line(fp, error);
		d->profile = polycap_profile_new_from_file(f_cap, f_axs, f_ext, error);
		free(f_cap);
		free(f_axs);
		free(f_ext);
	}
	if (!d->profile) {
		fclose(fp);
		polycap_source_free(s);
		return NULL;
	}
	/* the capillary count is written as a real number ("200000.") */
	double ncap_real = 0;
	if (fscanf(fp, "%lf", &ncap_real) == 1)
		d->n_cap = (int64_t)ncap_real;
	fclose(fp);

	polycap_description_check_weight(d->nelem, d->wi, error);
	d->open_area = open_area_of(d);

	const char *bad = NULL;
	if (s->d_source < 0.0) bad = "polycap_source_new_from_file: source_temp->d_source must be greater than 0.0";
	else if (s->src_x < 0.0) bad = "polycap_source_new_from_file: source_temp->src_x must be greater than 0.0";
	else if (s->src_y < 0.0) bad = "polycap_source_new_from_file: source_temp->src_y must be greater than 0.0";
	else if (s->n_energies <= 0.) bad = "polycap_source_new_from_file: source->n_energies must be greater than 0";
	for (size_t i = 0; !bad && i < s->n_energies; i++)
		if (s->energies[i] < 1. || s->energies[i] > 100.)
			bad = "polycap_source_new_from_file: source->energies must be greater than 1 and smaller than 100";
	if (!bad && d->n_cap < 1) bad = "polycap_source_new_from_file: description->n_cap must be greater than 1";
	if (!bad && (d->open_area < 0 || d->open_area > 1)) bad = "polycap_source_new_from_file: description->open_area must be greater than 0 and less than 1";
	for (unsigned int i = 0; !bad && i < d->nelem; i++)
		if (d->iz[i] < 1 || d->iz[i] > 94)
			bad = "polycap_source_new_from_file: description->iz[i] must be greater than 0 and less than 94";
	if (!bad && d->density < 0.0) bad = "polycap_source_new_from_file: description->density must be greater than 0.0";
	if (bad) {
		polycap_clear_error(error);
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, bad);
		polycap_source_free(s);
		return NULL;
	}
	if (polycap_profile_validate(d->profile, d->n_cap, error) != 1) {
		polycap_clear_error(error);
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT,
					  "polycap_source_new_from_file: description->profile is faulty. Some capillary coordinates are outside of the external radius.");
		polycap_source_free(s);
		return NULL;
	}
	return s;
}

extern "C" void polycap_source_free(polycap_source *source)
{
	if (!source)
		return;
	polycap_description_free(source->description);
	polycap_rng_free(source->rng);
	free(source->energies);
	free(source);
}

extern "C" const polycap_description *polycap_source_get_description(polycap_source *source) { return source->description; }
