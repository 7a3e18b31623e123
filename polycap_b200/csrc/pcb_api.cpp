/*
 * pcb_api.cpp -- the photon-transport entry points of the polycap C API, served by the GPU.
 *
 *   polycap_photon_launch                          /root/reference/src/polycap-photon.c:390-955
 *   polycap_source_get_photon                      /root/reference/src/polycap-source.c:23-144
 *   polycap_source_get_transmission_efficiencies   /root/reference/src/polycap-source.c:448-1087
 *   result getters / frees                         /root/reference/src/polycap-transmission-efficiencies.c:782-1172
 *   TEST_BUILD-private single-item entry points    /root/reference/src/polycap-private.h:40-49,183-192
 *   additive polycap_b200_* entry points           include/polycap_b200.h
 *
 * There is no CPU implementation of the transport in this library: every function that traces,
 * samples or reflects a photon launches the CUDA kernels of pcb_trace.cuh / pcb_leak.cuh /
 * pcb_unit.cu and fails with POLYCAP_ERROR_RUNTIME when no CUDA device is usable.
 */
#include <cerrno>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <algorithm>
#include <mutex>
#include <vector>

#include "optconst.h"
#include "pcb_gpu.h"
#include "pcb_internal.h"
#include "pcb_unit.h"
#include "polycap_b200.h"

extern "C" void pcb_energy_constants(size_t n_energies, const double *energies, const double *amu, const double *scatf, double density, double *out);

static std::mutex g_gpu_mutex; /* the device objects below are not re-entrant; calls are serialised */

static bool quiet() { return getenv("POLYCAP_B200_QUIET") != NULL; }
/* no optical constants for this material: which provider was asked and what the caller can do (optconst.h) */
static void optconst_fail(polycap_error **error, const char *who)
{
	polycap_set_error(error, POLYCAP_ERROR_UNSUPPORTED,
			  "%s: optical constants unavailable for this composition with provider '%s' (install xraylib, supply "
			  "POLYCAP_B200_OPTCONST_TABLE for this material, or opt in to the built-in approximation with POLYCAP_B200_OPTCONST=builtin)",
			  who, pcb_optconst_provider_name());
}

/* POLYCAP_B200_PHASES=1: wall-clock phases of polycap_source_get_transmission_efficiencies on stderr (tuning aid) */
static bool phases_on() { return getenv("POLYCAP_B200_PHASES") != NULL; }
static double now_ms()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* ------------------------------------------------------------------ optic cache
 * A device optic = profile tables + per-energy constants.  polycap_photon_launch is called once
 * per photon by API users (tests/source.c:313-330, python Photon.launch), so the uploaded optic
 * is cached and keyed by a hash of everything it depends on. */
struct CachedOptic {
	uint64_t key = 0;
	pcb_gpu_optic *g = nullptr;
	uint64_t stamp = 0;
};
static CachedOptic g_cache[24];
static uint64_t g_stamp = 0;

static uint64_t fnv(uint64_t h, const void *p, size_t n)
{
	const unsigned char *b = (const unsigned char *)p;
	for (size_t i = 0; i < n; i++) {
		h ^= b[i];
		h *= 1099511628211ULL;
	}
	return h;
}

/* amu/scatf: NULL -> optical-constants provider */
static pcb_gpu_optic *optic_for(const polycap_description *d, size_t nE, const double *energies, const double *amu_in, const double *scatf_in,
				int device, polycap_error **error, const char *who, int replica = 0)
{
	const polycap_profile *p = d->profile;
	if (!p) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: description->profile cannot be NULL", who);
		return nullptr;
	}
	std::vector<double> amu(nE), scatf(nE);
	if (amu_in && scatf_in) {
		memcpy(amu.data(), amu_in, nE * 8);
		memcpy(scatf.data(), scatf_in, nE * 8);
	} else if (pcb_optconst_compound(d->nelem, d->iz, d->wi, d->density, nE, energies, amu.data(), scatf.data()) != 0) {
		optconst_fail(error, who);
		return nullptr;
	}
	uint64_t h = 1469598103934665603ULL;
	const size_t n = (size_t)p->nmax + 1;
	h = fnv(h, &p->nmax, sizeof p->nmax);
	h = fnv(h, p->z, n * 8);
	h = fnv(h, p->cap, n * 8);
	h = fnv(h, p->ext, n * 8);
	h = fnv(h, &d->sig_rough, 8);
	h = fnv(h, &d->n_cap, 8);
	h = fnv(h, &d->density, 8);
	h = fnv(h, energies, nE * 8);
	h = fnv(h, amu.data(), nE * 8);
	h = fnv(h, scatf.data(), nE * 8);
	const int resolved = pcb_gpu_resolve_device(device); /* -1 -> POLYCAP_B200_DEVICE / current device: an optic lives on ONE device */
	h = fnv(h, &resolved, sizeof resolved);
	h = fnv(h, &replica, sizeof replica); /* a device listed twice in POLYCAP_B200_DEVICES gets two independent optics (testing aid) */
	for (const char *k : {"POLYCAP_B200_EXACT_FRESNEL", "POLYCAP_B200_EXACT_CHECKS", "POLYCAP_B200_KERNEL", "POLYCAP_B200_DEFER_MIN",
			      "POLYCAP_B200_DEFER_HK"}) { /* modes read at optic creation */
		const char *v = getenv(k);
		h = fnv(h, v ? v : "", v ? strlen(v) + 1 : 1);
	}
	CachedOptic *slot = &g_cache[0];
	const bool no_cache = getenv("POLYCAP_B200_REUPLOAD") != NULL; /* re-send tables + constants on every call (end-to-end timing) */
	for (auto &c : g_cache) {
		if (c.g && c.key == h) {
			c.stamp = ++g_stamp;
			if (no_cache && pcb_gpu_optic_reupload(c.g, NULL)) { /* inputs travel host -> device on every call */
				polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: CUDA failure: %s", who, pcb_gpu_last_error());
				return nullptr;
			}
			return c.g;
		}
		if (c.stamp < slot->stamp)
			slot = &c;
	}
	pcb_optic_desc od;
	od.nmax = p->nmax;
	od.z = p->z;
	od.cap = p->cap;
	od.ext = p->ext;
	od.sig_rough = d->sig_rough;
	od.n_cap = d->n_cap;
	od.density = d->density;
	od.n_energies = nE;
	od.energies = energies;
	od.amu = amu.data();
	od.scatf = scatf.data();
	pcb_gpu_optic *g = pcb_gpu_optic_create(device, &od);
	if (!g) {
		polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: no usable CUDA device for the photon transport (%s); this library has no CPU path", who,
				  pcb_gpu_last_error());
		return nullptr;
	}
	if (slot->g)
		pcb_gpu_optic_destroy(slot->g);
	slot->g = g;
	slot->key = h;
	slot->stamp = ++g_stamp;
	return g;
}

static void gpu_fail(polycap_error **error, const char *who) { polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: CUDA failure: %s", who, pcb_gpu_last_error()); }

/* ------------------------------------------------------------------ photon object (photon.c:96-137, 956-1169) */

extern "C" polycap_photon *polycap_photon_new(polycap_description *description, polycap_vector3 start_coords, polycap_vector3 start_direction,
					      polycap_vector3 start_electric_vector, polycap_error **error)
{
	if (description == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_new: description cannot be NULL");
		return NULL;
	}
	if (start_coords.z < 0.) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_new: start_coords.z must be greater than 0");
		return NULL;
	}
	if (start_direction.z < 0.) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_new: start_direction.z must be greater than 0");
		return NULL;
	}
	polycap_photon *ph = (polycap_photon *)calloc(1, sizeof(polycap_photon));
	if (!ph) {
		polycap_set_error(error, POLYCAP_ERROR_MEMORY, "polycap_photon_new: could not allocate memory for photon -> %s", strerror(errno));
		return NULL;
	}
	ph->description = description; /* borrowed, photon.c:124 */
	ph->start_coords = ph->exit_coords = start_coords;
	ph->start_direction = ph->exit_direction = start_direction;
	ph->start_electric_vector = ph->exit_electric_vector = start_electric_vector;
	ph->d_travel = 0;
	return ph;
}

extern "C" double polycap_photon_get_dtravel(polycap_photon *photon) { return photon->d_travel; }
extern "C" int64_t polycap_photon_get_irefl(polycap_photon *photon) { return photon->i_refl; }
extern "C" polycap_vector3 polycap_photon_get_start_coords(polycap_photon *photon) { return photon->start_coords; }
extern "C" polycap_vector3 polycap_photon_get_start_direction(polycap_photon *photon) { return photon->start_direction; }
extern "C" polycap_vector3 polycap_photon_get_start_electric_vector(polycap_photon *photon) { return photon->start_electric_vector; }
extern "C" polycap_vector3 polycap_photon_get_exit_coords(polycap_photon *photon) { return photon->exit_coords; }
extern "C" polycap_vector3 polycap_photon_get_exit_direction(polycap_photon *photon) { return photon->exit_direction; }
extern "C" polycap_vector3 polycap_photon_get_exit_electric_vector(polycap_photon *photon) { return photon->exit_electric_vector; }

extern "C" polycap_leak *polycap_leak_new(polycap_vector3 leak_coords, polycap_vector3 leak_dir, polycap_vector3 leak_elecv, int64_t n_refl,
					  size_t n_energies, double *weights, polycap_error **error)
{
	polycap_leak *leak = (polycap_leak *)malloc(sizeof(polycap_leak));
	if (!leak) {
		polycap_set_error(error, POLYCAP_ERROR_MEMORY, "polycap_leak_new: could not allocate memory for leak -> %s", strerror(errno));
		return NULL;
	}
	leak->coords = leak_coords;
	leak->direction = leak_dir;
	leak->elecv = leak_elecv;
	leak->n_refl = n_refl;
	leak->n_energies = n_energies;
	leak->weight = (double *)malloc(sizeof(double) * n_energies);
	memcpy(leak->weight, weights, sizeof(double) * n_energies);
	return leak;
}

extern "C" void polycap_leak_free(polycap_leak *leak)
{
	if (!leak)
		return;
	free(leak->weight);
	free(leak);
}

static bool copy_leaks(polycap_leak **src, int64_t n, polycap_leak ***leaks, int64_t *n_leaks, polycap_error **error, const char *who, const char *what)
{
	*n_leaks = n;
	if (n == 0) {
		*leaks = NULL;
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: no %s events in photon", who, what);
		return false;
	}
	*leaks = (polycap_leak **)malloc(sizeof(polycap_leak *) * (size_t)n);
	for (int64_t i = 0; i < n; i++)
		(*leaks)[i] = polycap_leak_new(src[i]->coords, src[i]->direction, src[i]->elecv, src[i]->n_refl, src[i]->n_energies, src[i]->weight, NULL);
	return true;
}

extern "C" bool polycap_photon_get_extleak_data(polycap_photon *photon, polycap_leak ***leaks, int64_t *n_leaks, polycap_error **error)
{
	if (!photon) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_get_extleak_data: photon cannot be NULL");
		return false;
	}
	return copy_leaks(photon->extleak, photon->n_extleak, leaks, n_leaks, error, "polycap_photon_get_extleak_data", "extleak");
}

extern "C" bool polycap_photon_get_intleak_data(polycap_photon *photon, polycap_leak ***leaks, int64_t *n_leaks, polycap_error **error)
{
	if (!photon) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_get_intleak_data: photon cannot be NULL");
		return false;
	}
	return copy_leaks(photon->intleak, photon->n_intleak, leaks, n_leaks, error, "polycap_photon_get_intleak_data", "intleak");
}

static void photon_drop_leaks(polycap_photon *photon)
{
	if (photon->extleak) {
		for (int64_t i = 0; i < photon->n_extleak; i++)
			polycap_leak_free(photon->extleak[i]);
		free(photon->extleak);
		photon->extleak = NULL;
	}
	if (photon->intleak) {
		for (int64_t i = 0; i < photon->n_intleak; i++)
			polycap_leak_free(photon->intleak[i]);
		free(photon->intleak);
		photon->intleak = NULL;
	}
	photon->n_extleak = photon->n_intleak = 0;
}

extern "C" void polycap_photon_free(polycap_photon *photon)
{
	if (!photon)
		return;
	free(photon->energies);
	free(photon->weight);
	free(photon->amu);
	free(photon->scatf);
	photon_drop_leaks(photon);
	free(photon);
}

/* photon.c:22-94: fills photon->amu / photon->scatf for photon->energies (host table look-up, once) */
extern "C" void polycap_photon_scatf(polycap_photon *photon, polycap_error **error)
{
	if (photon == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_scatf: photon cannot be NULL");
		return;
	}
	polycap_description *d = photon->description;
	if (d == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_scatf: description cannot be NULL");
		return;
	}
	if (d->density <= 0) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_scatf: description->density must be greater than 0");
		return;
	}
	if (d->nelem <= 0) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_scatf: description->nelem must be greater than 0");
		return;
	}
	if (photon->n_energies <= 0) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_scatf: photon->n_energies must be greater than 0");
		return;
	}
	for (size_t i = 0; i < photon->n_energies; i++)
		if (photon->energies[i] < 1. || photon->energies[i] > 100.) {
			polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_scatf: photon->energies[i] must be greater than 1 and smaller than 100");
			return;
		}
	for (unsigned int i = 0; i < d->nelem; i++) {
		if (d->wi[i] < 0. || d->wi[i] > 1.) {
			polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_scatf: description->wi[i] must be greater than 0 and smaller than 1");
			return;
		}
		if (d->iz[i] < 1 || d->iz[i] > 111) {
			polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_scatf: description->iz[i] must be greater than 0 and smaller than 104");
			return;
		}
	}
	free(photon->amu);
	free(photon->scatf);
	photon->amu = (double *)malloc(sizeof(double) * photon->n_energies);
	photon->scatf = (double *)malloc(sizeof(double) * photon->n_energies);
	if (pcb_optconst_compound(d->nelem, d->iz, d->wi, d->density, photon->n_energies, photon->energies, photon->amu, photon->scatf) != 0)
		optconst_fail(error, "polycap_photon_scatf");
}

/* ------------------------------------------------------------------ small vector helpers
 * exported for the reference's tests; host arithmetic on a caller-owned struct, not transport */
extern "C" void polycap_norm(polycap_vector3 *v) /* photon.c:365-376 */
{
	double len = sqrt(v->x * v->x + v->y * v->y + v->z * v->z);
	v->x /= len;
	v->y /= len;
	v->z /= len;
}
extern "C" double polycap_scalar(polycap_vector3 a, polycap_vector3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* photon.c:379-386 */

/* ------------------------------------------------------------------ replay plumbing shared by launch / launch_batch */

/* photon.c:390-955 */
extern "C" int polycap_photon_launch(polycap_photon *photon, size_t n_energies, double *energies, double **weights, bool leak_calc,
				     polycap_error **error)
{
	if (photon == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_launch: photon cannot be NULL");
		return -1;
	}
	if (energies == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_launch: energies cannot be NULL");
		return -1;
	}
	if (n_energies < 1) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_launch: n_energies must be greater than 0");
		return -1;
	}
	for (size_t i = 0; i < n_energies; i++)
		if (energies[i] < 1. || energies[i] > 100.) {
			polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_launch: energies[i] must be greater than 1 and less than 100");
			return -1;
		}
	if (weights == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_launch: weights cannot be NULL");
		return -1;
	}
	photon_drop_leaks(photon);
	polycap_description *d = photon->description;
	if (d == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_launch: description cannot be NULL");
		return -1;
	}
	*weights = (double *)malloc(sizeof(double) * n_energies);
	if (!*weights) {
		polycap_set_error(error, POLYCAP_ERROR_MEMORY, "polycap_photon_launch: could not allocate memory for weights -> %s", strerror(errno));
		return -1;
	}
	photon->n_energies = n_energies;
	photon->i_refl = 0;

	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	pcb_gpu_optic *g = optic_for(d, n_energies, energies, NULL, NULL, -1, error, "polycap_photon_launch");
	if (!g) { /* no device / no optical constants: nothing was traced, the caller gets no weight vector to free */
		free(*weights);
		*weights = NULL;
		return -1;
	}
	int ret = -1;
	double st[11];
	int64_t irefl = 0;
	const double c[3] = {photon->start_coords.x, photon->start_coords.y, photon->start_coords.z};
	const double dir[3] = {photon->start_direction.x, photon->start_direction.y, photon->start_direction.z};
	const double ev[3] = {photon->start_electric_vector.x, photon->start_electric_vector.y, photon->start_electric_vector.z};
	int rc;
	int64_t lcount[2] = {0, 0}, n_rec = 0;
	double *lrec = NULL;
	if (leak_calc)
		rc = pcb_gpu_replay_leak(g, 1, c, dir, ev, &ret, st, &irefl, *weights, NULL, lcount, &lrec, &n_rec);
	else
		rc = pcb_gpu_replay(g, 1, c, dir, ev, &ret, st, &irefl, *weights, NULL);
	if (rc) {
		gpu_fail(error, "polycap_photon_launch");
		free(lrec);
		free(*weights);
		*weights = NULL;
		return -1;
	}
	/* the reference normalises start_direction in place: once in launch (photon.c:492) and once per capil_trace call (capil.c:1237) */
	int ntrace = (int)st[10];
	for (int k = 0; k < ntrace + 1; k++) {
		polycap_vector3 prev = photon->start_direction;
		polycap_norm(&photon->start_direction);
		if (prev.x == photon->start_direction.x && prev.y == photon->start_direction.y && prev.z == photon->start_direction.z)
			break;
	}
	photon->exit_coords.x = st[0]; photon->exit_coords.y = st[1]; photon->exit_coords.z = st[2];
	photon->exit_direction.x = st[3]; photon->exit_direction.y = st[4]; photon->exit_direction.z = st[5];
	photon->exit_electric_vector.x = st[6]; photon->exit_electric_vector.y = st[7]; photon->exit_electric_vector.z = st[8];
	photon->d_travel = st[9]; /* launch does not reset d_travel; photon_new / get_photon start it at 0 */
	photon->i_refl = irefl;
	if (leak_calc && n_rec > 0) {
		/* leak records: [kind(0 ext,1 int), coords(3), dir(3), elecv(3), n_refl, weights(nE)] in the reference's order */
		const size_t stride = 11 + n_energies;
		photon->extleak = (polycap_leak **)malloc(sizeof(polycap_leak *) * (size_t)(lcount[0] ? lcount[0] : 1));
		photon->intleak = (polycap_leak **)malloc(sizeof(polycap_leak *) * (size_t)(lcount[1] ? lcount[1] : 1));
		for (int64_t i = 0; i < n_rec; i++) {
			const double *r = lrec + (size_t)i * stride;
			polycap_vector3 lc = {r[1], r[2], r[3]}, ld = {r[4], r[5], r[6]}, le = {r[7], r[8], r[9]};
			polycap_leak *lk = polycap_leak_new(lc, ld, le, (int64_t)r[10], n_energies, (double *)(r + 11), NULL);
			if (r[0] == 0.)
				photon->extleak[photon->n_extleak++] = lk;
			else
				photon->intleak[photon->n_intleak++] = lk;
		}
		if (photon->n_extleak == 0) {
			free(photon->extleak);
			photon->extleak = NULL;
		}
		if (photon->n_intleak == 0) {
			free(photon->intleak);
			photon->intleak = NULL;
		}
	}
	free(lrec);
	if (ret == -2) { /* photon.c:519,555 */
		const double n_shells = round(sqrt(12. * d->n_cap - 3.) / 6. - 0.5);
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT,
					  n_shells == 0. ? "polycap_photon_launch: photon_pos_check: photon not within monocapillary boundaries"
							 : "polycap_photon_launch: photon_pos_check: photon not within optic boundaries");
	}
	/* photon.c:930-945: energies/weight/amu/scatf are NULL on return (this library never attaches them in launch) */
	free(photon->energies); photon->energies = NULL;
	free(photon->weight); photon->weight = NULL;
	free(photon->amu); photon->amu = NULL;
	free(photon->scatf); photon->scatf = NULL;
	return ret;
}

/* ------------------------------------------------------------------ source sampling (source.c:23-144) */

extern "C" polycap_photon *polycap_source_get_photon(polycap_source *source, polycap_rng *rng, polycap_error **error)
{
	if (source == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_get_photon: source cannot be NULL");
		return NULL;
	}
	polycap_description *d = source->description;
	if (d == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_get_photon: description cannot be NULL");
		return NULL;
	}
	if (rng == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_get_photon: rng cannot be NULL");
		return NULL;
	}
	/* The sampling arithmetic runs on the device (pcb_unit.cu) over uniforms drawn from the caller's
	 * MT19937 stream; the stream is advanced by exactly the number of draws the reference would make. */
	const double n_shells = round(sqrt(12. * d->n_cap - 3.) / 6. - 0.5);
	const double src[10] = {source->d_source, source->src_x, source->src_y, source->src_sigx, source->src_sigy,
				source->src_shiftx, source->src_shifty, source->hor_pol, d->profile->cap[0], d->profile->ext[0]};
	double out[11];
	int used = 0;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	for (int nu = 64;; nu *= 4) {
		struct pcb_mt19937 probe = *rng->_rng;
		std::vector<double> u((size_t)nu);
		for (int i = 0; i < nu; i++)
			u[i] = pcb_mt_uniform(&probe);
		int rc = pcb_unit_get_photon(src, n_shells == 0., u.data(), nu, out, &used);
		if (rc <= -100) {
			polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "polycap_source_get_photon: no usable CUDA device (%s); this library has no CPU path",
					  pcb_unit_last_error());
			return NULL;
		}
		if (rc == 0)
			break;
		if (nu > (1 << 20)) {
			polycap_set_error_literal(error, POLYCAP_ERROR_RUNTIME, "polycap_source_get_photon: entrance-window rejection loop did not terminate");
			return NULL;
		}
	}
	for (int i = 0; i < used; i++)
		pcb_mt_uniform(rng->_rng);
	polycap_vector3 c = {out[0], out[1], out[2]}, dir = {out[3], out[4], out[5]}, ev = {out[6], out[7], out[8]};
	polycap_photon *ph = polycap_photon_new(d, c, dir, ev, error);
	if (ph) {
		ph->src_start_coords.x = out[9];
		ph->src_start_coords.y = out[10];
		ph->src_start_coords.z = 0;
	}
	return ph;
}

/* ------------------------------------------------------------------ the run (source.c:448-1087) */

static uint64_t run_seed()
{
	const char *e = getenv("POLYCAP_B200_SEED");
	if (e && *e)
		return strtoull(e, NULL, 0);
	uint64_t s = 0;
	FILE *fp = fopen("/dev/urandom", "r"); /* the reference seeds every thread from /dev/urandom (source.c:740, rng.c:50-71) */
	if (!fp || fread(&s, sizeof s, 1, fp) != 1)
		s = (uint64_t)time(NULL);
	if (fp)
		fclose(fp);
	return s;
}

static bool check_source(polycap_source *source, polycap_error **error, const char *who)
{
	if (source == NULL) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: source cannot be NULL", who);
		return false;
	}
	if (source->description == NULL) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: description cannot be NULL", who);
		return false;
	}
	if (source->n_energies < 1) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: source->n_energies must be greater than or equal to 1", who);
		return false;
	}
	if (source->energies == NULL) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: source->energies cannot be NULL", who);
		return false;
	}
	for (size_t i = 0; i < source->n_energies; i++)
		if (source->energies[i] < 1. || source->energies[i] > 100.) {
			polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: source->energies[i] must be greater than 1 and less than 100", who);
			return false;
		}
	return true;
}

static pcb_source_desc source_desc(const polycap_source *s)
{
	pcb_source_desc sd = {s->d_source, s->src_x, s->src_y, s->src_sigx, s->src_sigy, s->src_shiftx, s->src_shifty, s->hor_pol};
	return sd;
}

extern "C" void polycap_transmission_efficiencies_free(polycap_transmission_efficiencies *efficiencies);

/* ---- result arenas.  The exit records of a call -- 16 double arrays, the reflection counts and the weight matrix of
 * struct _polycap_images -- are slices of ONE pinned host block laid out like the device's record buffer, so the device
 * copies straight into the arrays the caller reads: no staging copy, no first-touch page faults.  Pinning is slow
 * (~1 GB/s), so blocks released by polycap_transmission_efficiencies_free are kept for the next call. */
struct ResultArena {
	void *p;
	size_t bytes;
	bool pinned;
};
static std::mutex g_arena_mutex;
static std::vector<ResultArena> g_arena_pool;						/* released blocks, at most ARENA_KEEP */
static std::vector<std::pair<const void *, ResultArena>> g_arena_of;			/* live results -> their block */
enum { ARENA_KEEP = 3 };

static ResultArena arena_take(size_t bytes)
{
	{
		std::lock_guard<std::mutex> lk(g_arena_mutex);
		int best = -1;
		for (int i = 0; i < (int)g_arena_pool.size(); i++)
			if (g_arena_pool[i].bytes >= bytes && g_arena_pool[i].bytes <= 4 * bytes + (1 << 20) &&
			    (best < 0 || g_arena_pool[i].bytes < g_arena_pool[best].bytes))
				best = i;
		if (best >= 0) {
			ResultArena a = g_arena_pool[best];
			g_arena_pool.erase(g_arena_pool.begin() + best);
			return a;
		}
	}
	/* POLYCAP_B200_PAGEABLE_RESULTS=1: plain malloc (a one-shot caller with a multi-GB result does not pay for pinning) */
	const bool pageable = getenv("POLYCAP_B200_PAGEABLE_RESULTS") != NULL;
	ResultArena a = {pageable ? malloc(bytes) : pcb_gpu_host_alloc(bytes), bytes, !pageable};
	if (!a.p && !pageable) { /* pinning a multi-GB block can fail where plain memory is still there: slower downloads, same results */
		a.p = malloc(bytes);
		a.pinned = false;
	}
	return a;
}

static void arena_release(ResultArena a)
{
	ResultArena drop = {NULL, 0, false};
	{
		std::lock_guard<std::mutex> lk(g_arena_mutex);
		g_arena_pool.push_back(a);
		if (g_arena_pool.size() > ARENA_KEEP) { /* drop the smallest */
			size_t k = 0;
			for (size_t i = 1; i < g_arena_pool.size(); i++)
				if (g_arena_pool[i].bytes < g_arena_pool[k].bytes)
					k = i;
			drop = g_arena_pool[k];
			g_arena_pool.erase(g_arena_pool.begin() + k);
		}
	}
	if (drop.p) {
		if (drop.pinned)
			pcb_gpu_host_free(drop.p);
		else
			free(drop.p);
	}
}

/* Devices of a run.  POLYCAP_B200_DEVICES = "all" or a comma-separated list of CUDA ordinals shards the photon ids of
 * polycap_source_get_transmission_efficiencies over those devices from this one process (the C API has one caller
 * thread); unset: one device (POLYCAP_B200_DEVICE or the calling thread's current device). */
static std::vector<int> run_devices()
{
	std::vector<int> out;
	const char *e = getenv("POLYCAP_B200_DEVICES");
	if (!e || !*e) {
		out.push_back(-1);
		return out;
	}
	const int n = pcb_gpu_device_count();
	if (!strcmp(e, "all")) {
		for (int i = 0; i < n && i < 16; i++)
			out.push_back(i);
	} else {
		const char *p = e;
		while (*p) {
			char *end;
			long v = strtol(p, &end, 10);
			if (end == p)
				break;
			if (v >= 0 && v < n && out.size() < 16)
				out.push_back((int)v); /* a device may be listed more than once: each entry is an independent share (testing aid) */
			p = (*end == ',') ? end + 1 : end;
		}
	}
	if (out.empty())
		out.push_back(-1);
	return out;
}

static int64_t g_last_counters[PCB_C_COUNT];
static bool g_have_last_counters = false;

extern "C" bool polycap_b200_last_run_counters(int64_t *counters)
{
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	if (!counters || !g_have_last_counters)
		return false;
	memcpy(counters, g_last_counters, sizeof g_last_counters);
	return true;
}

struct DevRun { /* one device's share of a run */
	pcb_gpu_optic *g;
	int dev;
	double *d_sum;
	int64_t *d_cnt;
	double *d_rf, *d_rw;
	int64_t rec_cap, rec_used;
	int64_t n_ids, exits; /* current wave */
};

/* One id-ordered wave over the devices of a run, without leaks: ids [id0, id0 + *n_ids) are dealt to the devices as contiguous
 * sub-ranges in device order; phase 1 traces and ranks the exit photons on every device, then the "need-th exit in id order" is
 * found from the devices' exit counts: devices before it keep all their ids, the one that holds it cuts, those after it are
 * dropped (their ids count as not consumed).  Records of the kept ids go to host arrays laid out [17][host_cap] / [host_cap][nE] at
 * record host_base, downloaded on each device's copy stream while the caller launches the next wave.
 * Returns 0, 1 = CUDA failure (pcb_gpu_last_error), 2 = error already set. */
static int plain_wave(std::vector<DevRun> &R, const pcb_source_desc &sd, uint64_t seed, int64_t id0, int64_t *n_ids, int64_t need, double *h_fields,
		      double *h_weights, int64_t host_cap, int64_t host_base, int64_t *used_out, int64_t *got_out, polycap_error **error, const char *WHO)
{
	const int G = (int)R.size();
	int64_t used = 0, got = 0;
	/* phase 1 on every device: trace + rank the exit photons */
	const int64_t per = (*n_ids + G - 1) / G;
	int64_t off = 0;
	for (int k = 0; k < G; k++) {
		DevRun &r = R[k];
		r.n_ids = std::max<int64_t>(0, std::min<int64_t>(per, *n_ids - off));
		if (r.n_ids > 0 && pcb_gpu_wave_begin(r.g, &sd, seed, id0 + off, r.n_ids))
			return 1;
		off += r.n_ids;
	}
	*n_ids = off;
	for (int k = 0; k < G; k++) {
		R[k].exits = 0;
		if (R[k].n_ids > 0 && pcb_gpu_wave_count(R[k].g, &R[k].exits))
			return 1;
	}
	/* phase 2: place the cut, tally, compact, and send the records on their way to the host */
	for (int k = 0; k < G && got < need; k++) {
		DevRun &r = R[k];
		if (r.n_ids <= 0)
			continue;
		const bool cut_here = got + r.exits >= need;
		const int64_t n_wanted = cut_here ? need - got : 0;
		const int64_t n_rec = cut_here ? n_wanted : r.exits;
		const bool records = h_fields != NULL; /* tally-only callers keep no records: nothing is compacted or downloaded */
		if (records && n_rec > r.rec_cap) {
			polycap_set_error(error, POLYCAP_ERROR_MEMORY, "%s: %" PRId64 " exit records of one wave exceed the record buffer of device %d (%" PRId64
					  "); use more devices or fewer photons per call", WHO, n_rec, r.dev, r.rec_cap);
			return 2;
		}
		if (records && r.rec_used + n_rec > r.rec_cap) { /* reuse the record buffer from its start once its downloads are done */
			if (pcb_gpu_wave_reuse_records(r.g))
				return 1;
			r.rec_used = 0;
		}
		if (pcb_gpu_wave_finish(r.g, r.n_ids, n_wanted, r.d_sum, r.d_cnt, records ? r.d_rf : NULL, records ? r.d_rw : NULL, r.rec_cap, r.rec_used))
			return 1;
		int64_t ids_k = r.n_ids;
		if (cut_here) {
			int64_t ex_k = 0;
			if (pcb_gpu_wave_result(r.g, &ids_k, &ex_k))
				return 1;
			if (ex_k != n_rec) {
				polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: internal error: cut placed at %" PRId64 " exits, expected %" PRId64, WHO, ex_k, n_rec);
				return 2;
			}
		}
		if (records && pcb_gpu_wave_download(r.g, r.d_rf, r.d_rw, r.rec_cap, r.rec_used, n_rec, h_fields, h_weights, host_cap, host_base + got))
			return 1;
		if (records)
			r.rec_used += n_rec;
		got += n_rec;
		used += ids_k;
	}
	*used_out = used;
	*got_out = got;
	return 0;
}

extern "C" polycap_transmission_efficiencies *polycap_source_get_transmission_efficiencies(polycap_source *source, int max_threads, int n_photons,
											    bool leak_calc, polycap_progress_monitor *progress_monitor,
											    polycap_error **error)
{
	static const char *WHO = "polycap_source_get_transmission_efficiencies";
	(void)max_threads; /* host threads have no role: the trace runs on the GPU */
	if (source == NULL) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: source cannot be NULL", WHO);
		return NULL;
	}
	if (progress_monitor != NULL) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: progress_monitor must be NULL as polycap_progress_monitor currently has no implementation", WHO);
		return NULL;
	}
	if (!check_source(source, error, WHO))
		return NULL;
	if (n_photons < 1) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: n_photons must be greater than 1", WHO);
		return NULL;
	}
	polycap_description *d = source->description;
	const size_t nE = source->n_energies;
	const int64_t want = n_photons;

	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	const double t_begin = now_ms();
	/* one device optic per device of the run (leak runs: one device) */
	std::vector<int> devs = run_devices();
	if (leak_calc)
		devs.resize(1);
	const int G = (int)devs.size();
	std::vector<DevRun> R((size_t)G);
	for (int k = 0; k < G; k++) {
		const int replica = (int)std::count(devs.begin(), devs.begin() + k, devs[k]);
		R[k].g = optic_for(d, nE, source->energies, NULL, NULL, devs[k], error, WHO, replica);
		if (!R[k].g)
			return NULL;
		R[k].dev = pcb_gpu_optic_device(R[k].g);
	}
	pcb_gpu_optic *g = R[0].g;
	const int dev = R[0].dev;
	const uint64_t seed = run_seed();
	const pcb_source_desc sd = source_desc(source);
	const double t_optic = now_ms();

	/* Per device: tallies and a record buffer in one device block kept with the optic (no cudaMalloc / cudaFree per call).  With one
	 * device the buffer holds every record of the call; with several each holds a share and is reused from its start if it fills
	 * up (after its downloads have drained). */
	auto up256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
	const int64_t wave_bytes_per_id = 160 + 8 * (int64_t)nE;
	const int64_t wave_max = std::max<int64_t>(4096, std::min<int64_t>((int64_t)1 << 24, ((int64_t)6 << 30) / wave_bytes_per_id));
	for (int k = 0; k < G; k++) {
		DevRun &r = R[k];
		/* every device could hold all records; only calls beyond 32 GB of records per device split the capacity */
		r.rec_cap = (G == 1 || (double)want * (17 + (double)nE) * 8 <= 32e9) ? want : std::min<int64_t>(want, want / G + want / (2 * G) + 65536);
		r.rec_used = 0;
		const size_t o_cnt = up256(nE * 8), o_rf = o_cnt + up256(PCB_C_COUNT * 8), o_rw = o_rf + up256((size_t)r.rec_cap * 17 * 8),
			     res_bytes = o_rw + up256((size_t)r.rec_cap * nE * 8);
		char *d_res = (char *)pcb_gpu_optic_result(r.g, res_bytes);
		void *st = d_res ? pcb_gpu_optic_stream(r.g) : NULL;
		if (!d_res || !st) {
			polycap_set_error(error, POLYCAP_ERROR_MEMORY, "%s: could not allocate device memory for %d exit records -> %s", WHO, n_photons,
					  pcb_gpu_last_error());
			return NULL;
		}
		r.d_sum = (double *)d_res;
		r.d_cnt = (int64_t *)(d_res + o_cnt);
		r.d_rf = (double *)(d_res + o_rf);
		r.d_rw = (double *)(d_res + o_rw);
		pcb_gpu_memset(r.dev, r.d_sum, 0, nE * 8, leak_calc ? NULL : st);
		pcb_gpu_memset(r.dev, r.d_cnt, 0, PCB_C_COUNT * 8, leak_calc ? NULL : st);
	}
	/* leak records of counted photons, wave by wave (the malloc'd blocks the waves return, not copied) */
	struct LeakChunk {
		double *rec;
		int64_t n;
	};
	struct LeakChunks : std::vector<LeakChunk> {
		~LeakChunks()
		{
			for (LeakChunk &c : *this)
				free(c.rec);
		}
	} leak_chunks;
	int64_t n_leak_all = 0;

	/* the host block whose slices are the arrays of struct _polycap_images: the devices copy their records straight into it,
	 * wave by wave, while the next wave is traced */
	polycap_transmission_efficiencies *eff = (polycap_transmission_efficiencies *)calloc(1, sizeof(*eff));
	eff->images = (struct _polycap_images *)calloc(1, sizeof(struct _polycap_images));
	struct _polycap_images *im = eff->images;
	const size_t rf_bytes = (size_t)want * 17 * 8, rw_bytes = (size_t)want * nE * 8;
	const ResultArena arena = arena_take(rf_bytes + rw_bytes);
	double *stage = (double *)arena.p;
	if (!stage) {
		polycap_set_error(error, POLYCAP_ERROR_MEMORY, "%s: could not allocate host memory for %d exit records -> %s", WHO, n_photons, pcb_gpu_last_error());
		polycap_transmission_efficiencies_free(eff);
		return NULL;
	}
	auto fail = [&](bool cuda) {
		if (cuda)
			gpu_fail(error, WHO);
		for (int k = 0; k < G; k++)
			pcb_gpu_wave_download_sync(R[k].g);
		arena_release(arena);
		polycap_transmission_efficiencies_free(eff);
		return (polycap_transmission_efficiencies *)NULL;
	};
	double *h_fields = stage, *h_weights = stage + (size_t)want * 17;

	/* id-ordered waves until `want` photons reached the exit window; the wave that holds the want-th exit is cut right after it, so
	 * Sum(w)/started stays an unbiased estimator (source.c:745-884 redraws per slot).  A wave's ids are dealt to the devices as
	 * contiguous sub-ranges in device order, so "the want-th exit in id order" is found from the devices' exit counts: devices
	 * before it keep all their ids, the one that holds it cuts, those after it are dropped (their ids count as not consumed). */
	int64_t have = 0, next_id = 0, waves = 0;
	/* Wave size: ids needed at the exit fraction seen so far plus three standard deviations of that estimate and of the count
	 * itself (ids past the want-th exit are traced and discarded).  The first wave of a call uses the fraction the previous call
	 * with the same source geometry ended with; without one it is a probe of 2*need ids.  Calls that return many megabytes of
	 * records are split into up to 8 waves so that the download of one wave's records overlaps the trace of the next. */
	uint64_t src_key = 1469598103934665603ULL;
	for (size_t i = 0; i < sizeof sd; i++)
		src_key = (src_key ^ ((const unsigned char *)&sd)[i]) * 1099511628211ULL;
	int64_t est_exits = 0;
	double exit_frac = pcb_gpu_optic_hint(g, src_key, &est_exits);
	int64_t chunk_ids = 0, leak_shrink = 1;
	const double t_alloc = now_ms();
	while (have < want) {
		const double t_w0 = now_ms();
		int64_t need = want - have;
		int64_t n_ids;
		if (exit_frac > 0) {
			const double margin = std::min(0.25, 3.0 * sqrt(1.0 / (double)std::max<int64_t>(est_exits, 1) + 1.0 / (double)need) + 0.002);
			n_ids = (int64_t)((double)need / exit_frac * (1.0 + margin)) + 256;
		} else {
			n_ids = 2 * need + 1024;
		}
		if (chunk_ids == 0 && exit_frac > 0) {
			const char *ce = getenv("POLYCAP_B200_CHUNKS");
			int64_t chunks = ce ? atoll(ce) : (int64_t)((rf_bytes + rw_bytes) / ((size_t)48 << 20));
			/* splitting pays only where the records are a sizeable part of a wave's traffic: below ~256 record bytes per photon id
			 * (short grids, low exit fractions) the download hides behind nothing worth hiding and smaller waves only trace less well */
			if (!ce && exit_frac * (17.0 + (double)nE) * 8.0 < 256.0)
				chunks = 1;
			chunks = std::max<int64_t>(1, std::min<int64_t>(8, chunks));
			/* a wave should refill every CTA's photon pool (148 x 1152 slots) a few times: shorter launches are mostly ramp-up and tail */
			chunk_ids = std::max<int64_t>((n_ids + chunks - 1) / chunks, ce ? (int64_t)G * 65536 : (int64_t)G * 450000);
		}
		if (!leak_calc && chunk_ids > 0 && n_ids > chunk_ids + chunk_ids / 4)
			n_ids = chunk_ids;
		if (n_ids > wave_max * G)
			n_ids = wave_max * G;
		int64_t used = 0, got = 0;
		if (leak_calc) {
			double *lrec = NULL;
			int64_t n_lrec = 0;
			/* a leak wave also holds room for 96 leak records of 11+nE doubles per id (pcb_leak.cu): its own bytes-per-id budget, 24 GB per
			 * wave; a wave whose record pool still overflows (thin walls, hard photons) is retried with half the ids */
			const int64_t leak_bytes_per_id = wave_bytes_per_id + 96 * (11 + (int64_t)nE) * 8;
			const int64_t leak_wave_max = std::max<int64_t>(1024, std::min<int64_t>(wave_max, ((int64_t)24 << 30) / leak_bytes_per_id));
			if (n_ids > leak_wave_max)
				n_ids = leak_wave_max;
			if (leak_shrink > 1)
				n_ids = std::max<int64_t>(256, n_ids / leak_shrink);
			int rc = pcb_gpu_run_wave_leak(g, &sd, seed, next_id, n_ids, need, R[0].d_sum, R[0].d_cnt, R[0].d_rf, R[0].d_rw, want, have, &used, &got, &lrec, &n_lrec);
			if (rc == -2 && n_ids > 256) {
				free(lrec);
				leak_shrink *= 2;
				continue;
			}
			if (rc == 0 && n_lrec > 0) {
				leak_chunks.push_back(LeakChunk{lrec, n_lrec});
				n_leak_all += n_lrec;
				lrec = NULL;
			}
			free(lrec);
			if (rc)
				return fail(true);
		} else {
			const int prc = plain_wave(R, sd, seed, next_id, &n_ids, need, h_fields, h_weights, want, have, &used, &got, error, WHO);
			if (prc)
				return fail(prc == 1);
		}
		next_id += used;
		have += got;
		waves++;
		exit_frac = std::max(1e-4, (double)have / (double)next_id);
		est_exits = have;
		if (phases_on())
			fprintf(stderr, "polycap-b200 phases: wave %" PRId64 ": %" PRId64 " ids launched on %d device(s), %" PRId64 " used, %" PRId64 " exits, %.2f ms\n", waves,
				n_ids, G, used, got, now_ms() - t_w0);
		if (waves > 100000 || (next_id > ((int64_t)1 << 26) && have == 0)) {
			polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: no photon reaches the exit window (%" PRId64 " started)", WHO, next_id);
			return fail(false);
		}
	}

	pcb_gpu_optic_set_hint(g, src_key, (double)have / (double)next_id, have);
	const double t_waves = now_ms();
	std::vector<double> sum(nE);
	int64_t cnt[PCB_C_COUNT];
	int used_peer = 0;
	{
		/* tallies: the per-device sums are added on the first device over peer memory (NVLink) or, without peer access, on the host */
		std::vector<pcb_gpu_optic *> og((size_t)G);
		std::vector<double *> ds((size_t)G);
		std::vector<int64_t *> dc((size_t)G);
		for (int k = 0; k < G; k++) {
			og[k] = R[k].g;
			ds[k] = R[k].d_sum;
			dc[k] = R[k].d_cnt;
		}
		if (pcb_gpu_reduce_tallies(og.data(), ds.data(), dc.data(), G, sum.data(), cnt, &used_peer))
			return fail(true);
	}
	memcpy(g_last_counters, cnt, sizeof g_last_counters);
	g_have_last_counters = true;
	const double t_h2 = now_ms();
	if (leak_calc) { /* single wave buffer, default stream: one download at the end */
		if (pcb_gpu_copy_to_host(dev, stage, R[0].d_rf, rf_bytes, NULL) || pcb_gpu_copy_to_host(dev, stage + (size_t)want * 17, R[0].d_rw, rw_bytes, NULL))
			return fail(true);
	}
	for (int k = 0; k < G; k++)
		if (pcb_gpu_wave_download_sync(R[k].g))
			return fail(true);
	if (phases_on())
		fprintf(stderr, "polycap-b200 phases: tallies %.2f ms (%s), remaining record download %.2f ms\n", t_h2 - t_waves,
			G > 1 ? (used_peer ? "peer reduce" : "host reduce") : "one device", now_ms() - t_h2);
	{
		std::lock_guard<std::mutex> lk(g_arena_mutex);
		g_arena_of.emplace_back((const void *)eff, arena);
	}
	const double t_d2h = now_ms();
	double *rf = stage;
	/* the record fields of struct _polycap_images (private.h:156-181) in the device buffer's order; the reflection counts
	 * arrive as doubles and are converted in place */
	double **dst[16] = {&im->src_start_coords[0], &im->src_start_coords[1], &im->pc_start_coords[0], &im->pc_start_coords[1], &im->pc_start_dir[0],
			    &im->pc_start_dir[1], &im->pc_start_elecv[0], &im->pc_start_elecv[1], &im->pc_exit_coords[0], &im->pc_exit_coords[1],
			    &im->pc_exit_coords[2], &im->pc_exit_dir[0], &im->pc_exit_dir[1], &im->pc_exit_elecv[0], &im->pc_exit_elecv[1], &im->pc_exit_dtravel};
	const int src_field[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 16};
	for (int f = 0; f < 16; f++)
		*dst[f] = rf + (size_t)src_field[f] * want;
	im->exit_coord_weights = rf + (size_t)want * 17;
	im->pc_exit_nrefl = (int64_t *)(rf + (size_t)15 * want);
	for (int64_t i = 0; i < want; i++) {
		const double v = rf[(size_t)15 * want + i];
		im->pc_exit_nrefl[i] = (int64_t)v;
	}
	if (phases_on())
		fprintf(stderr, "polycap-b200 phases: optic %.2f ms, device buffers %.2f ms, waves %.2f ms, tallies + last download %.2f ms, record pointers %.2f ms\n",
			t_optic - t_begin, t_alloc - t_optic, t_waves - t_alloc, t_d2h - t_waves, now_ms() - t_d2h);
	/* leak records (source.c:799-879, 982-1031) */
	if (leak_calc) {
		const size_t stride = 11 + nE;
		int64_t n_ext = 0, n_int = 0;
		for (const LeakChunk &ch : leak_chunks)
			for (int64_t i = 0; i < ch.n; i++)
				(ch.rec[(size_t)i * stride] == 0. ? n_ext : n_int)++;
		im->i_extleak = n_ext;
		im->i_intleak = n_int;
		auto alloc_d = [](int64_t n) { return (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1)); };
		for (int k = 0; k < 3; k++) {
			im->extleak_coords[k] = alloc_d(n_ext);
			im->intleak_coords[k] = alloc_d(n_int);
		}
		for (int k = 0; k < 2; k++) {
			im->extleak_dir[k] = alloc_d(n_ext);
			im->intleak_dir[k] = alloc_d(n_int);
			im->intleak_elecv[k] = alloc_d(n_int);
		}
		im->extleak_coord_weights = alloc_d(n_ext * (int64_t)nE);
		im->intleak_coord_weights = alloc_d(n_int * (int64_t)nE);
		im->extleak_n_refl = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_ext > 0 ? n_ext : 1));
		im->intleak_n_refl = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_int > 0 ? n_int : 1));
		int64_t ie = 0, ii = 0;
		for (const LeakChunk &ch : leak_chunks)
		for (int64_t i = 0; i < ch.n; i++) {
			const double *r = ch.rec + (size_t)i * stride;
			if (r[0] == 0.) {
				for (int k = 0; k < 3; k++)
					im->extleak_coords[k][ie] = r[1 + k];
				im->extleak_dir[0][ie] = r[4];
				im->extleak_dir[1][ie] = r[5];
				im->extleak_n_refl[ie] = (int64_t)r[10];
				memcpy(im->extleak_coord_weights + (size_t)ie * nE, r + 11, nE * 8);
				ie++;
			} else {
				for (int k = 0; k < 3; k++)
					im->intleak_coords[k][ii] = r[1 + k];
				im->intleak_dir[0][ii] = r[4];
				im->intleak_dir[1][ii] = r[5];
				im->intleak_elecv[0][ii] = r[7];
				im->intleak_elecv[1][ii] = r[8];
				im->intleak_n_refl[ii] = (int64_t)r[10];
				memcpy(im->intleak_coord_weights + (size_t)ii * nE, r + 11, nE * 8);
				ii++;
			}
		}
	}

	const int64_t iexit = cnt[PCB_C_EXIT], ne = cnt[PCB_C_NOT_ENTERED], nt = cnt[PCB_C_NOT_TRANSMITTED];
	const double nsh = (round(sqrt(12. * d->n_cap - 3.) / 6. - 0.5) + 0.5) * 6.;
	printf("Average number of reflections: %lf, Simulated photons: %" PRId64 "\n", (double)cnt[PCB_C_REFL_EXIT] / n_photons, iexit + ne + nt);
	printf("Open area Calculated: %lf, Simulated: %lf\n",
	       nsh * nsh / 12. * (d->profile->cap[0] * d->profile->cap[0] * M_PI) / (3. * sin(M_PI / 3) * d->profile->ext[0] * d->profile->ext[0]),
	       (double)(iexit + nt) / (iexit + ne + nt));
	printf("iexit: %" PRId64 ", no enter: %" PRId64 ", no trans: %" PRId64 "\n", iexit, ne, nt);
	if (!quiet())
		fprintf(stderr, "polycap-b200: %" PRId64 " photon ids in %" PRId64 " wave(s) on CUDA device %d, seed %" PRIu64 ", optical constants: %s\n", next_id, waves,
			dev, seed, pcb_optconst_provider_name());

	d->open_area = (double)(iexit + nt) / (iexit + ne + nt); /* source.c:1066 */
	eff->n_energies = nE;
	eff->source = source; /* borrowed, source.c:682 */
	im->i_start = iexit + ne + nt;
	im->i_exit = iexit;
	eff->energies = (double *)malloc(sizeof(double) * nE);
	eff->efficiencies = (double *)malloc(sizeof(double) * nE);
	for (size_t i = 0; i < nE; i++) {
		eff->energies[i] = source->energies[i];
		eff->efficiencies[i] = (sum[i] / ((double)iexit + (double)nt)) * d->open_area; /* source.c:1075 */
	}
	return eff;
}

/* ------------------------------------------------------------------ results (transmission-efficiencies.c:782-1172) */

extern "C" void polycap_transmission_efficiencies_free(polycap_transmission_efficiencies *eff)
{
	if (!eff)
		return;
	ResultArena arena = {NULL, 0, false};
	{
		std::lock_guard<std::mutex> lk(g_arena_mutex);
		for (size_t i = 0; i < g_arena_of.size(); i++)
			if (g_arena_of[i].first == (const void *)eff) {
				arena = g_arena_of[i].second;
				g_arena_of.erase(g_arena_of.begin() + i);
				break;
			}
	}
	free(eff->energies);
	free(eff->efficiencies);
	if (eff->images) {
		struct _polycap_images *im = eff->images;
		if (!arena.p) { /* exit records are slices of the arena otherwise */
			for (int k = 0; k < 2; k++) {
				free(im->src_start_coords[k]); free(im->pc_start_coords[k]); free(im->pc_start_dir[k]); free(im->pc_start_elecv[k]);
				free(im->pc_exit_dir[k]); free(im->pc_exit_elecv[k]);
			}
			for (int k = 0; k < 3; k++)
				free(im->pc_exit_coords[k]);
			free(im->pc_exit_nrefl); free(im->pc_exit_dtravel); free(im->exit_coord_weights);
		}
		for (int k = 0; k < 2; k++) {
			free(im->extleak_dir[k]); free(im->intleak_dir[k]); free(im->intleak_elecv[k]);
		}
		for (int k = 0; k < 3; k++) {
			free(im->extleak_coords[k]); free(im->intleak_coords[k]);
		}
		free(im->extleak_coord_weights); free(im->extleak_n_refl); free(im->intleak_coord_weights); free(im->intleak_n_refl);
		free(im);
	}
	free(eff);
	if (arena.p)
		arena_release(arena);
}

extern "C" bool polycap_transmission_efficiencies_get_data(polycap_transmission_efficiencies *eff, size_t *n_energies, double **energies_arr,
							   double **efficiencies_arr, polycap_error **error)
{
	if (eff == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_transmission_efficiencies_get_data: efficiencies cannot be NULL");
		return false;
	}
	if (n_energies)
		*n_energies = eff->n_energies;
	if (energies_arr) {
		*energies_arr = (double *)malloc(sizeof(double) * eff->n_energies);
		memcpy(*energies_arr, eff->energies, sizeof(double) * eff->n_energies);
	}
	if (efficiencies_arr) {
		*efficiencies_arr = (double *)malloc(sizeof(double) * eff->n_energies);
		memcpy(*efficiencies_arr, eff->efficiencies, sizeof(double) * eff->n_energies);
	}
	return true;
}

static bool images_ok(polycap_transmission_efficiencies *eff, polycap_error **error, const char *who)
{
	if (eff == NULL) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: efficiencies cannot be NULL", who);
		return false;
	}
	if (eff->images == NULL) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: source->images cannot be NULL", who);
		return false;
	}
	return true;
}

static polycap_vector3 *vec_alloc(int64_t n) { return (polycap_vector3 *)malloc(sizeof(polycap_vector3) * (size_t)(n > 0 ? n : 1)); }

extern "C" bool polycap_transmission_efficiencies_get_start_data(polycap_transmission_efficiencies *eff, int64_t *n_start, int64_t *n_exit,
								 polycap_vector3 **start_coords, polycap_vector3 **start_direction,
								 polycap_vector3 **start_elecv, polycap_vector3 **src_start_coords, polycap_error **error)
{
	if (!images_ok(eff, error, "polycap_source_get_start_data"))
		return false;
	struct _polycap_images *im = eff->images;
	*n_start = im->i_start;
	if (im->i_start == 0) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_get_start_data: no photon start events in efficiencies");
		return false;
	}
	*n_exit = im->i_exit;
	if (im->i_exit == 0) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_get_start_data: no photon exit events in efficiencies");
		return false;
	}
	const int64_t n = im->i_exit;
	*start_coords = vec_alloc(n);
	*start_direction = vec_alloc(n);
	*start_elecv = vec_alloc(n);
	*src_start_coords = vec_alloc(n);
	for (int64_t i = 0; i < n; i++) {
		polycap_vector3 c = {im->pc_start_coords[0][i], im->pc_start_coords[1][i], 0.};
		polycap_vector3 dr = {im->pc_start_dir[0][i], im->pc_start_dir[1][i], 0.};
		dr.z = sqrt(1. - dr.x * dr.x - dr.y * dr.y); /* only two components are stored (transmission-efficiencies.c:832-838) */
		polycap_vector3 e = {im->pc_start_elecv[0][i], im->pc_start_elecv[1][i], 0.};
		e.z = sqrt(1. - e.x * e.x - e.y * e.y);
		polycap_vector3 s = {im->src_start_coords[0][i], im->src_start_coords[1][i], 0.};
		(*start_coords)[i] = c;
		(*start_direction)[i] = dr;
		(*start_elecv)[i] = e;
		(*src_start_coords)[i] = s;
	}
	return true;
}

extern "C" bool polycap_transmission_efficiencies_get_exit_data(polycap_transmission_efficiencies *eff, int64_t *n_exit, polycap_vector3 **exit_coords,
								polycap_vector3 **exit_direction, polycap_vector3 **exit_elecv, int64_t **n_refl,
								double **d_travel, size_t *n_energies, double ***exit_weights, polycap_error **error)
{
	if (!images_ok(eff, error, "polycap_source_get_exit_data"))
		return false;
	struct _polycap_images *im = eff->images;
	*n_exit = im->i_exit;
	*n_energies = eff->n_energies;
	if (im->i_start == 0) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_source_get_exit_data: no photon exit events in efficiencies");
		return false;
	}
	const int64_t n = im->i_exit;
	const size_t nE = eff->n_energies;
	*exit_coords = vec_alloc(n);
	*exit_direction = vec_alloc(n);
	*exit_elecv = vec_alloc(n);
	*n_refl = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
	*d_travel = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
	*exit_weights = (double **)malloc(sizeof(double *) * (size_t)(n > 0 ? n : 1));
	for (int64_t i = 0; i < n; i++) {
		(*n_refl)[i] = im->pc_exit_nrefl[i];
		(*d_travel)[i] = im->pc_exit_dtravel[i];
		polycap_vector3 c = {im->pc_exit_coords[0][i], im->pc_exit_coords[1][i], im->pc_exit_coords[2][i]};
		polycap_vector3 dr = {im->pc_exit_dir[0][i], im->pc_exit_dir[1][i], 0.};
		dr.z = sqrt(1. - dr.x * dr.x - dr.y * dr.y);
		polycap_vector3 e = {im->pc_exit_elecv[0][i], im->pc_exit_elecv[1][i], 0.};
		e.z = sqrt(1. - e.x * e.x - e.y * e.y);
		(*exit_coords)[i] = c;
		(*exit_direction)[i] = dr;
		(*exit_elecv)[i] = e;
		(*exit_weights)[i] = (double *)malloc(sizeof(double) * nE);
		memcpy((*exit_weights)[i], im->exit_coord_weights + (size_t)i * nE, sizeof(double) * nE);
	}
	return true;
}

static bool leaks_from_images(polycap_transmission_efficiencies *eff, bool ext, polycap_leak ***leaks, int64_t *n_leaks, polycap_error **error)
{
	const char *who = ext ? "polycap_source_get_extleak_data" : "polycap_source_get_intleak_data";
	if (!images_ok(eff, error, who))
		return false;
	struct _polycap_images *im = eff->images;
	const int64_t n = ext ? im->i_extleak : im->i_intleak;
	const size_t nE = eff->n_energies;
	*n_leaks = n;
	if (n == 0) {
		*leaks = NULL;
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: no %s events in efficiencies", who, ext ? "extleak" : "intleak");
		return false;
	}
	*leaks = (polycap_leak **)malloc(sizeof(polycap_leak *) * (size_t)n);
	for (int64_t i = 0; i < n; i++) {
		polycap_leak *lk = (polycap_leak *)calloc(1, sizeof(polycap_leak));
		double **co = ext ? im->extleak_coords : im->intleak_coords;
		double **di = ext ? im->extleak_dir : im->intleak_dir;
		lk->coords.x = co[0][i]; lk->coords.y = co[1][i]; lk->coords.z = co[2][i];
		lk->direction.x = di[0][i]; lk->direction.y = di[1][i];
		lk->direction.z = sqrt(1. - lk->direction.x * lk->direction.x - lk->direction.y * lk->direction.y);
		if (!ext) { /* the reference fills elecv for internal leaks only (transmission-efficiencies.c:960-963, 1031-1034) */
			lk->elecv.x = im->intleak_elecv[0][i]; lk->elecv.y = im->intleak_elecv[1][i];
			lk->elecv.z = sqrt(1. - lk->elecv.x * lk->elecv.x - lk->elecv.y * lk->elecv.y);
		}
		lk->n_energies = nE;
		lk->n_refl = ext ? im->extleak_n_refl[i] : im->intleak_n_refl[i];
		lk->weight = (double *)malloc(sizeof(double) * nE);
		memcpy(lk->weight, (ext ? im->extleak_coord_weights : im->intleak_coord_weights) + (size_t)i * nE, sizeof(double) * nE);
		(*leaks)[i] = lk;
	}
	return true;
}

extern "C" bool polycap_transmission_efficiencies_get_extleak_data(polycap_transmission_efficiencies *eff, polycap_leak ***leaks, int64_t *n_leaks,
								   polycap_error **error)
{
	return leaks_from_images(eff, true, leaks, n_leaks, error);
}
extern "C" bool polycap_transmission_efficiencies_get_intleak_data(polycap_transmission_efficiencies *eff, polycap_leak ***leaks, int64_t *n_leaks,
								   polycap_error **error)
{
	return leaks_from_images(eff, false, leaks, n_leaks, error);
}

/* ------------------------------------------------------------------ TEST_BUILD-private single-item entry points
 * Each runs the device function on one item (pcb_unit.cu); a CUDA failure is reported as POLYCAP_ERROR_RUNTIME. */

static bool unit_failed(int rc, polycap_error **error, const char *who)
{
	if (rc > -100)
		return false;
	polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: no usable CUDA device (%s); this library has no CPU path", who, pcb_unit_last_error());
	return true;
}

extern "C" int polycap_photon_within_pc_boundary(double polycap_radius, polycap_vector3 photon_coord, polycap_error **error) /* photon.c:139-169 */
{
	if (polycap_radius <= 0.) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_photon_within_pc_boundary: polycap_radius must be greater than 0");
		return -1;
	}
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	int rc = pcb_unit_within(polycap_radius, photon_coord.x, photon_coord.y);
	if (unit_failed(rc, error, "polycap_photon_within_pc_boundary"))
		return -1;
	return rc;
}

extern "C" int polycap_capil_segment(polycap_vector3 cap_coord0, polycap_vector3 cap_coord1, double cap_rad0, double cap_rad1, polycap_vector3 phot_coord0,
				     polycap_vector3 phot_coord1, polycap_vector3 photon_dir, polycap_vector3 *photon_coord, polycap_vector3 *surface_norm,
				     polycap_error **error) /* capil.c:52-255 */
{
	static const char *W = "polycap_capil_segment";
	if (cap_coord0.z < 0.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: cap_coord0.z must be greater than 0", W); return -1; }
	if (cap_coord1.z < 0.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: cap_coord1.z must be greater than 0", W); return -1; }
	if (cap_rad0 < 0.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: cap_rad0 must be greater than 0", W); return -1; }
	if (cap_rad1 < 0.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: cap_rad1 must be greater than 0", W); return -1; }
	if (photon_coord == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: photon_coord cannot be NULL", W); return -1; }
	if (photon_dir.z < 0.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: photon_dir.z must be greater than 0", W); return -1; }
	if (surface_norm == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: surface_norm cannot be NULL", W); return -1; }
	const double in[20] = {cap_coord0.x, cap_coord0.y, cap_coord0.z, cap_coord1.x, cap_coord1.y, cap_coord1.z, cap_rad0, cap_rad1,
			       phot_coord0.x, phot_coord0.y, phot_coord0.z, phot_coord1.x, phot_coord1.y, phot_coord1.z,
			       photon_dir.x, photon_dir.y, photon_dir.z, photon_coord->x, photon_coord->y, photon_coord->z};
	double out[6];
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	int rc = pcb_unit_segment(in, out);
	if (unit_failed(rc, error, W))
		return -1;
	surface_norm->x = out[3]; surface_norm->y = out[4]; surface_norm->z = out[5];
	if (rc == 1) {
		photon_coord->x = out[0]; photon_coord->y = out[1]; photon_coord->z = out[2];
	}
	return rc;
}

static void energy_constants_for(const polycap_photon *photon, const polycap_description *d, std::vector<double> &ec)
{
	ec.resize(6 * photon->n_energies);
	pcb_energy_constants(photon->n_energies, photon->energies, photon->amu, photon->scatf, d->density, ec.data());
}

extern "C" double polycap_refl_polar(double e, double density, double scatf, double lin_abs_coeff, polycap_vector3 surface_norm, polycap_photon *photon,
				     polycap_vector3 *electric_vector, polycap_error **error) /* capil.c:444-563 */
{
	static const char *W = "polycap_refl_polar";
	if (e < 1. || e > 100.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: e must be greater than 1 and smaller than 100.", W); return -1; }
	if (density <= 0.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: density must be greater than 0", W); return -1; }
	if (scatf < 0.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: scatf must be greater than or equal to 0", W); return -1; }
	if (lin_abs_coeff < 0.) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: lin_abs_coeff must be greater than or equal to 0", W); return -1; }
	if (photon == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: photon cannot be NULL", W); return -1; }
	double ec[6];
	pcb_energy_constants(1, &e, &lin_abs_coeff, &scatf, density, ec);
	const double in[13] = {surface_norm.x, surface_norm.y, surface_norm.z, photon->exit_direction.x, photon->exit_direction.y, photon->exit_direction.z,
			       photon->exit_electric_vector.x, photon->exit_electric_vector.y, photon->exit_electric_vector.z, ec[0], ec[1], ec[2], ec[3]};
	double out[7];
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	int rc = pcb_unit_refl_polar(in, out);
	if (unit_failed(rc, error, W))
		return -1;
	photon->exit_electric_vector.x = out[4]; photon->exit_electric_vector.y = out[5]; photon->exit_electric_vector.z = out[6];
	if (electric_vector) {
		electric_vector->x = out[1]; electric_vector->y = out[2]; electric_vector->z = out[3];
	}
	return out[0];
}

static void append_leaks(polycap_photon *photon, size_t nE, const double *lrec, int64_t n_rec)
{
	const size_t stride = 11 + nE;
	for (int64_t i = 0; i < n_rec; i++) {
		const double *r = lrec + (size_t)i * stride;
		polycap_vector3 lc = {r[1], r[2], r[3]}, ld = {r[4], r[5], r[6]}, le = {r[7], r[8], r[9]};
		polycap_leak *lk = polycap_leak_new(lc, ld, le, (int64_t)r[10], nE, (double *)(r + 11), NULL);
		if (r[0] == 0.) {
			photon->extleak = (polycap_leak **)realloc(photon->extleak, sizeof(polycap_leak *) * (size_t)(photon->n_extleak + 1));
			photon->extleak[photon->n_extleak++] = lk;
		} else {
			photon->intleak = (polycap_leak **)realloc(photon->intleak, sizeof(polycap_leak *) * (size_t)(photon->n_intleak + 1));
			photon->intleak[photon->n_intleak++] = lk;
		}
	}
}

/* polycap_capil_reflect / polycap_capil_trace with leak_calc = true: one step of the leak state machine of pcb_leak.cu on the
 * caller's photon (capil.c:565-891, 1197-1361).  mode 1: reflect on `normal`; mode 2: trace (cap_x / cap_y give the capillary axis). */
static int leak_unit_step(polycap_photon *photon, polycap_description *d, int mode, polycap_vector3 normal, int *ix, const double *cap_x, const double *cap_y,
			  polycap_error **error, const char *W)
{
	const polycap_profile *p = d->profile;
	const size_t nE = photon->n_energies;
	double A = 0, B = 0;
	if (mode == 2) { /* the axis arrays must describe a capillary of the hexagonal lattice: cap_x[i] = A hexz[i], cap_y[i] = B hexz[i] */
		const double n_shells = round(sqrt(12. * d->n_cap - 3.) / 6. - 0.5);
		const double den = 2. * 0.86602540378443864676 * (n_shells + 1);
		int k = -1;
		for (int i = 0; i <= p->nmax && k < 0; i++)
			if (p->ext[i] != 0.)
				k = i;
		if (k >= 0) {
			A = cap_x[k] / (p->ext[k] / den);
			B = cap_y[k] / (p->ext[k] / den);
		}
		for (int i = 0; i <= p->nmax; i++) {
			const double hz = p->ext[i] / den;
			if (fabs(A * hz - cap_x[i]) > 1e-12 * (fabs(cap_x[i]) + 1e-300) + 1e-300 || fabs(B * hz - cap_y[i]) > 1e-12 * (fabs(cap_y[i]) + 1e-300) + 1e-300) {
				polycap_set_error(error, POLYCAP_ERROR_UNSUPPORTED, "%s: leak_calc=true needs cap_x / cap_y that follow the external profile (a capillary of the lattice)", W);
				return -1;
			}
		}
	}
	pcb_gpu_optic *g = optic_for(d, nE, photon->energies, photon->amu, photon->scatf, -1, error, W);
	if (!g)
		return -1;
	const double in[16] = {photon->exit_coords.x, photon->exit_coords.y, photon->exit_coords.z, photon->exit_direction.x, photon->exit_direction.y,
			       photon->exit_direction.z, photon->exit_electric_vector.x, photon->exit_electric_vector.y, photon->exit_electric_vector.z,
			       normal.x, normal.y, normal.z, (double)photon->i_refl, photon->d_travel, A, B};
	double outs[13];
	double *lrec = NULL;
	int64_t n_lrec = 0;
	if (pcb_gpu_leak_unit_step(g, mode, in, ix ? *ix : 0, photon->weight, outs, &lrec, &n_lrec)) {
		gpu_fail(error, W);
		return -1;
	}
	const int rc = (int)outs[0];
	if (n_lrec > 0)
		append_leaks(photon, nE, lrec, n_lrec);
	free(lrec);
	photon->exit_direction.x = outs[4]; photon->exit_direction.y = outs[5]; photon->exit_direction.z = outs[6];
	if (mode == 2) {
		photon->exit_coords.x = outs[1]; photon->exit_coords.y = outs[2]; photon->exit_coords.z = outs[3];
		photon->d_travel = outs[11];
		photon->i_refl = (int64_t)outs[10];
		*ix = (int)outs[12];
	}
	if (mode == 2 || rc == 0 || rc == 1 || rc == -2) {
		photon->exit_electric_vector.x = outs[7]; photon->exit_electric_vector.y = outs[8]; photon->exit_electric_vector.z = outs[9];
	}
	return rc;
}

extern "C" int polycap_capil_reflect(polycap_photon *photon, polycap_vector3 surface_norm, bool leak_calc, polycap_error **error) /* capil.c:565-891 */
{
	static const char *W = "polycap_capil_reflect";
	if (photon == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: photon cannot be NULL", W); return -1; }
	polycap_description *d = photon->description;
	if (d == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: description cannot be NULL", W); return -1; }
	if (photon->n_energies < 1 || !photon->energies || !photon->weight) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: photon energies/weights are not set", W);
		return -1;
	}
	if (!photon->amu || !photon->scatf) {
		polycap_photon_scatf(photon, error);
		if (!photon->amu)
			return -1;
	}
	std::vector<double> ec;
	energy_constants_for(photon, d, ec);
	double out[6];
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	if (leak_calc) {
		int rc = leak_unit_step(photon, d, 1, surface_norm, NULL, NULL, NULL, error, W);
		if (rc == -1 && error && *error == NULL) /* capil.c:599-602: the one error a reflect reports itself */
			polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: alfa must be greater than 0", W);
		return rc;
	}
	const double in[9] = {photon->exit_direction.x, photon->exit_direction.y, photon->exit_direction.z, photon->exit_electric_vector.x,
			      photon->exit_electric_vector.y, photon->exit_electric_vector.z, surface_norm.x, surface_norm.y, surface_norm.z};
	int rc = pcb_unit_reflect(in, ec.data(), (int)photon->n_energies, d->sig_rough, photon->weight, out);
	if (unit_failed(rc, error, W))
		return -1;
	photon->exit_direction.x = out[0]; photon->exit_direction.y = out[1]; photon->exit_direction.z = out[2];
	if (rc == 0 || rc == 1 || rc == -2) {
		photon->exit_electric_vector.x = out[3]; photon->exit_electric_vector.y = out[4]; photon->exit_electric_vector.z = out[5];
	}
	return rc;
}

extern "C" int polycap_capil_trace(int *ix, polycap_photon *photon, polycap_description *description, double *cap_x, double *cap_y, bool leak_calc,
				   polycap_error **error) /* capil.c:1197-1361 */
{
	static const char *W = "polycap_capil_trace";
	if (ix == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: ix cannot be NULL", W); return -1; }
	if (photon == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: photon cannot be NULL", W); return -1; }
	if (description == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: description cannot be NULL", W); return -1; }
	if (cap_x == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: cap_x cannot be NULL", W); return -1; }
	if (cap_y == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: cap_y cannot be NULL", W); return -1; }
	if (photon->n_energies < 1 || !photon->energies || !photon->weight) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: photon energies/weights are not set", W);
		return -1;
	}
	if (!photon->amu || !photon->scatf) {
		polycap_photon_scatf(photon, error);
		if (!photon->amu)
			return -1;
	}
	std::vector<double> ec;
	energy_constants_for(photon, description, ec);
	const polycap_profile *p = description->profile;
	const double n_shells = round(sqrt(12. * description->n_cap - 3.) / 6. - 0.5);
	const double in[10] = {photon->exit_coords.x, photon->exit_coords.y, photon->exit_coords.z, photon->exit_direction.x, photon->exit_direction.y,
			       photon->exit_direction.z, photon->exit_electric_vector.x, photon->exit_electric_vector.y, photon->exit_electric_vector.z,
			       photon->d_travel};
	double out[10];
	long long irefl = photon->i_refl;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	if (leak_calc) {
		polycap_vector3 none = {0., 0., 0.};
		polycap_norm(&photon->start_direction); /* capil.c:1237 */
		return leak_unit_step(photon, description, 2, none, ix, cap_x, cap_y, error, W);
	}
	int rc = pcb_unit_trace(in, ix, &irefl, p->z, p->cap, p->ext, cap_x, cap_y, p->nmax, n_shells == 0., ec.data(), (int)photon->n_energies,
				description->sig_rough, photon->weight, out);
	if (unit_failed(rc, error, W))
		return -1;
	polycap_norm(&photon->start_direction); /* capil.c:1237 */
	photon->exit_coords.x = out[0]; photon->exit_coords.y = out[1]; photon->exit_coords.z = out[2];
	photon->exit_direction.x = out[3]; photon->exit_direction.y = out[4]; photon->exit_direction.z = out[5];
	photon->exit_electric_vector.x = out[6]; photon->exit_electric_vector.y = out[7]; photon->exit_electric_vector.z = out[8];
	photon->d_travel = out[9];
	photon->i_refl = irefl;
	return rc;
}

extern "C" int polycap_capil_trace_wall(polycap_photon *photon, double *d_travel, int *capx_id, int *capy_id, polycap_error **error) /* capil.c:896-1194 */
{
	static const char *W = "polycap_capil_trace_wall";
	if (photon == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: photon cannot be NULL", W); return -1; }
	if (photon->description == NULL) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: description cannot be NULL", W); return -1; }
	if (!d_travel || !capx_id || !capy_id) { polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: output pointers cannot be NULL", W); return -1; }
	polycap_description *d = photon->description;
	const double in[6] = {photon->exit_coords.x, photon->exit_coords.y, photon->exit_coords.z,
			      photon->exit_direction.x, photon->exit_direction.y, photon->exit_direction.z};
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	double e10 = 10.0, one = 1.0; /* the wall march is energy independent: any valid energy keys the optic */
	pcb_gpu_optic *g = optic_for(d, 1, &e10, &one, &one, -1, error, W);
	if (!g)
		return -1;
	int rc = pcb_gpu_leak_trace_wall(g, in, d_travel, capx_id, capy_id);
	if (rc == -1 && *pcb_gpu_last_error()) {
		gpu_fail(error, W);
		return -1;
	}
	polycap_norm(&photon->exit_direction); /* capil.c:919 */
	return rc;
}

/* ------------------------------------------------------------------ HDF5 (transmission-efficiencies.c:318-780) -> pcb_hdf5.cpp */
extern "C" int pcb_hdf5_write(polycap_transmission_efficiencies *eff, const char *filename, char *errbuf, size_t errlen);

extern "C" bool polycap_transmission_efficiencies_write_hdf5(polycap_transmission_efficiencies *efficiencies, const char *filename, polycap_error **error)
{
	if (efficiencies == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_transmission_efficiencies_write_hdf5: efficiencies cannot be NULL");
		return false;
	}
	if (filename == NULL) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_transmission_efficiencies_write_hdf5: filename cannot be NULL");
		return false;
	}
	char buf[512] = "";
	if (pcb_hdf5_write(efficiencies, filename, buf, sizeof buf) != 0) {
		polycap_set_error(error, POLYCAP_ERROR_IO, "polycap_transmission_efficiencies_write_hdf5: %s", buf);
		return false;
	}
	return true;
}

/* ------------------------------------------------------------------ additive entry points (include/polycap_b200.h) */

extern "C" int polycap_b200_device_count(void) { return pcb_gpu_device_count(); }
extern "C" const char *polycap_b200_last_error(void) { return pcb_gpu_last_error(); }
extern "C" const char *polycap_b200_optconst_provider(void) { return pcb_optconst_provider_name(); }

extern "C" bool polycap_b200_optical_constants(polycap_description *description, size_t n_energies, const double *energies, double *amu, double *scatf,
					       polycap_error **error)
{
	if (!description || !energies || !amu || !scatf || n_energies < 1) {
		polycap_set_error_literal(error, POLYCAP_ERROR_INVALID_ARGUMENT, "polycap_b200_optical_constants: NULL argument or n_energies < 1");
		return false;
	}
	if (pcb_optconst_compound(description->nelem, description->iz, description->wi, description->density, n_energies, energies, amu, scatf) != 0) {
		optconst_fail(error, "polycap_b200_optical_constants");
		return false;
	}
	return true;
}

static bool launch_batch_impl(const char *W, polycap_description *description, size_t n_energies, const double *energies, const double *amu,
			      const double *scatf, int64_t n, const double *start_coords, const double *start_direction, const double *start_electric_vector,
			      bool leak_calc, int *ret, double *exit_coords, double *exit_direction, double *exit_electric_vector, int64_t *i_refl,
			      double *d_travel, double *weights, int64_t *counters, int64_t *leak_count, double **leak_records, int64_t *n_leak_records,
			      polycap_error **error);

extern "C" bool polycap_b200_photon_launch_batch(polycap_description *description, size_t n_energies, const double *energies, const double *amu,
						 const double *scatf, int64_t n, const double *start_coords, const double *start_direction,
						 const double *start_electric_vector, bool leak_calc, int *ret, double *exit_coords, double *exit_direction,
						 double *exit_electric_vector, int64_t *i_refl, double *d_travel, double *weights, int64_t *counters,
						 polycap_error **error)
{
	static const char *W = "polycap_b200_photon_launch_batch";
	if (!description || !energies || n_energies < 1 || n < 0 || (n > 0 && (!start_coords || !start_direction || !start_electric_vector))) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: NULL argument, n_energies < 1 or n < 0", W);
		return false;
	}
	for (size_t i = 0; i < n_energies; i++)
		if (energies[i] < 1. || energies[i] > 100.) {
			polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: energies[i] must be greater than 1 and less than 100", W);
			return false;
		}
	if (leak_calc) {
		polycap_set_error(error, POLYCAP_ERROR_UNSUPPORTED, "%s: use polycap_b200_photon_launch_batch_leaks for leak_calc = true", W);
		return false;
	}
	return launch_batch_impl(W, description, n_energies, energies, amu, scatf, n, start_coords, start_direction, start_electric_vector, false, ret,
				 exit_coords, exit_direction, exit_electric_vector, i_refl, d_travel, weights, counters, NULL, NULL, NULL, error);
}

extern "C" bool polycap_b200_photon_launch_batch_leaks(polycap_description *description, size_t n_energies, const double *energies, const double *amu,
						       const double *scatf, int64_t n, const double *start_coords, const double *start_direction,
						       const double *start_electric_vector, int *ret, double *exit_coords, double *exit_direction,
						       double *exit_electric_vector, int64_t *i_refl, double *d_travel, double *weights,
						       int64_t *leak_count, double **leak_records, int64_t *n_leak_records, polycap_error **error)
{
	static const char *W = "polycap_b200_photon_launch_batch_leaks";
	if (!description || !energies || n_energies < 1 || n < 0 || !leak_count || !leak_records || !n_leak_records ||
	    (n > 0 && (!start_coords || !start_direction || !start_electric_vector))) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: NULL argument, n_energies < 1 or n < 0", W);
		return false;
	}
	for (size_t i = 0; i < n_energies; i++)
		if (energies[i] < 1. || energies[i] > 100.) {
			polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: energies[i] must be greater than 1 and less than 100", W);
			return false;
		}
	return launch_batch_impl(W, description, n_energies, energies, amu, scatf, n, start_coords, start_direction, start_electric_vector, true, ret,
				 exit_coords, exit_direction, exit_electric_vector, i_refl, d_travel, weights, NULL, leak_count, leak_records, n_leak_records,
				 error);
}

static bool launch_batch_impl(const char *W, polycap_description *description, size_t n_energies, const double *energies, const double *amu,
			      const double *scatf, int64_t n, const double *start_coords, const double *start_direction, const double *start_electric_vector,
			      bool leak_calc, int *ret, double *exit_coords, double *exit_direction, double *exit_electric_vector, int64_t *i_refl,
			      double *d_travel, double *weights, int64_t *counters, int64_t *leak_count, double **leak_records, int64_t *n_leak_records,
			      polycap_error **error)
{
	if (leak_records) {
		*leak_records = NULL;
		*n_leak_records = 0;
	}
	if (n == 0) {
		if (counters)
			memset(counters, 0, PCB_C_COUNT * 8);
		return true;
	}
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	pcb_gpu_optic *g = optic_for(description, n_energies, energies, amu, scatf, -1, error, W);
	if (!g)
		return false;
	std::vector<int> r((size_t)n);
	std::vector<double> st((size_t)n * 11);
	std::vector<int64_t> ir((size_t)n);
	std::vector<double> wtmp;
	double *w = weights;
	if (!w) {
		wtmp.resize((size_t)n * n_energies);
		w = wtmp.data();
	}
	int rc = leak_calc ? pcb_gpu_replay_leak(g, n, start_coords, start_direction, start_electric_vector, r.data(), st.data(), ir.data(), w, counters, leak_count,
						 leak_records, n_leak_records)
			   : pcb_gpu_replay(g, n, start_coords, start_direction, start_electric_vector, r.data(), st.data(), ir.data(), w, counters);
	if (rc) {
		gpu_fail(error, W);
		return false;
	}
	for (int64_t i = 0; i < n; i++) {
		const double *s = &st[(size_t)i * 11];
		if (ret) ret[i] = r[i];
		if (exit_coords) memcpy(exit_coords + 3 * i, s, 24);
		if (exit_direction) memcpy(exit_direction + 3 * i, s + 3, 24);
		if (exit_electric_vector) memcpy(exit_electric_vector + 3 * i, s + 6, 24);
		if (d_travel) d_travel[i] = s[9];
		if (i_refl) i_refl[i] = ir[i];
	}
	return true;
}

extern "C" bool polycap_b200_source_run_ids(polycap_source *source, uint64_t seed, int64_t id0, int64_t n_ids, bool leak_calc, double *sum_weights,
					    int64_t *counters, int8_t *status, double *rec, polycap_error **error)
{
	static const char *W = "polycap_b200_source_run_ids";
	if (!check_source(source, error, W))
		return false;
	if (!sum_weights || !counters || n_ids < 1) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: sum_weights/counters cannot be NULL and n_ids must be >= 1", W);
		return false;
	}
	if (leak_calc) {
		polycap_set_error(error, POLYCAP_ERROR_UNSUPPORTED, "%s: leak_calc runs go through polycap_source_get_transmission_efficiencies", W);
		return false;
	}
	const size_t nE = source->n_energies;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	pcb_gpu_optic *g = optic_for(source->description, nE, source->energies, NULL, NULL, -1, error, W);
	if (!g)
		return false;
	const int dev = pcb_gpu_optic_device(g);
	const pcb_source_desc sd = source_desc(source);
	double *d_sum = (double *)pcb_gpu_malloc(dev, nE * 8);
	int64_t *d_cnt = (int64_t *)pcb_gpu_malloc(dev, PCB_C_COUNT * 8);
	bool ok = d_sum && d_cnt;
	if (ok) {
		pcb_gpu_memset(dev, d_sum, 0, nE * 8, NULL);
		pcb_gpu_memset(dev, d_cnt, 0, PCB_C_COUNT * 8, NULL);
		int64_t used = 0, got = 0;
		ok = pcb_gpu_run_wave(g, &sd, seed, id0, n_ids, 0, d_sum, d_cnt, NULL, NULL, 0, 0, NULL, &used, &got) == 0 &&
		     pcb_gpu_copy_to_host(dev, sum_weights, d_sum, nE * 8, NULL) == 0 && pcb_gpu_copy_to_host(dev, counters, d_cnt, PCB_C_COUNT * 8, NULL) == 0;
		if (ok && (status || rec)) {
			std::vector<int8_t> st;
			int8_t *sp = status;
			if (!sp) {
				st.resize((size_t)n_ids);
				sp = st.data();
			}
			ok = pcb_gpu_last_wave_dense(g, n_ids, sp, rec) == 0;
		}
	}
	pcb_gpu_free(dev, d_sum);
	pcb_gpu_free(dev, d_cnt);
	if (!ok) {
		gpu_fail(error, W);
		return false;
	}
	return true;
}

extern "C" bool polycap_b200_source_sample_ids(polycap_source *source, uint64_t seed, int64_t id0, int64_t n_ids, double *out, polycap_error **error)
{
	static const char *W = "polycap_b200_source_sample_ids";
	if (!check_source(source, error, W))
		return false;
	if (!out || n_ids < 0) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: out cannot be NULL and n_ids must be >= 0", W);
		return false;
	}
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	pcb_gpu_optic *g = optic_for(source->description, source->n_energies, source->energies, NULL, NULL, -1, error, W);
	if (!g)
		return false;
	const pcb_source_desc sd = source_desc(source);
	if (pcb_gpu_sample_ids(g, &sd, seed, id0, n_ids, out)) {
		gpu_fail(error, W);
		return false;
	}
	return true;
}

/* ---- sessions ---- */
struct _polycap_b200_session {
	pcb_gpu_optic *g;
	pcb_source_desc sd;
	uint64_t seed;
	int device;
};

extern "C" polycap_b200_session *polycap_b200_session_new(polycap_source *source, int device, uint64_t seed, polycap_error **error)
{
	static const char *W = "polycap_b200_session_new";
	if (!check_source(source, error, W))
		return NULL;
	const polycap_description *d = source->description;
	const polycap_profile *p = d->profile;
	const size_t nE = source->n_energies;
	std::vector<double> amu(nE), scatf(nE);
	if (pcb_optconst_compound(d->nelem, d->iz, d->wi, d->density, nE, source->energies, amu.data(), scatf.data()) != 0) {
		optconst_fail(error, W);
		return NULL;
	}
	pcb_optic_desc od = {p->nmax, p->z, p->cap, p->ext, d->sig_rough, d->n_cap, d->density, nE, source->energies, amu.data(), scatf.data()};
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	pcb_gpu_optic *g = pcb_gpu_optic_create(device, &od);
	if (!g) {
		polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: no usable CUDA device (%s); this library has no CPU path", W, pcb_gpu_last_error());
		return NULL;
	}
	polycap_b200_session *s = new _polycap_b200_session();
	s->g = g;
	s->sd = source_desc(source);
	s->seed = seed;
	s->device = pcb_gpu_optic_device(g);
	return s;
}

/* ---- streams ---- */
struct _polycap_b200_stream {
	std::vector<DevRun> R;     /* own optics (not the cache's: they must outlive other calls) and own device buffers */
	std::vector<void *> d_mem; /* per device: tallies + record buffer */
	pcb_source_desc sd;
	uint64_t seed;
	size_t nE;
	int64_t next_id, exits, rec_cap;
};

static pcb_gpu_optic *own_optic(polycap_source *source, int device, polycap_error **error, const char *W)
{
	const polycap_description *d = source->description;
	const polycap_profile *p = d->profile;
	const size_t nE = source->n_energies;
	std::vector<double> amu(nE), scatf(nE);
	if (pcb_optconst_compound(d->nelem, d->iz, d->wi, d->density, nE, source->energies, amu.data(), scatf.data()) != 0) {
		optconst_fail(error, W);
		return NULL;
	}
	pcb_optic_desc od = {p->nmax, p->z, p->cap, p->ext, d->sig_rough, d->n_cap, d->density, nE, source->energies, amu.data(), scatf.data()};
	pcb_gpu_optic *g = pcb_gpu_optic_create(device, &od);
	if (!g)
		polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: no usable CUDA device (%s); this library has no CPU path", W, pcb_gpu_last_error());
	return g;
}

static void stream_release(polycap_b200_stream *st)
{
	for (size_t k = 0; k < st->R.size(); k++) {
		if (st->R[k].g) {
			pcb_gpu_wave_download_sync(st->R[k].g);
			if (k < st->d_mem.size() && st->d_mem[k])
				pcb_gpu_free(st->R[k].dev, st->d_mem[k]);
			pcb_gpu_optic_destroy(st->R[k].g);
		}
	}
	delete st;
}

extern "C" polycap_b200_stream *polycap_b200_stream_new(polycap_source *source, uint64_t seed, polycap_error **error)
{
	static const char *W = "polycap_b200_stream_new";
	if (!check_source(source, error, W))
		return NULL;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	polycap_b200_stream *st = new _polycap_b200_stream();
	st->sd = source_desc(source);
	st->seed = seed;
	st->nE = source->n_energies;
	st->next_id = st->exits = st->rec_cap = 0;
	const std::vector<int> devs = run_devices();
	for (int dv : devs) {
		DevRun r{};
		r.g = own_optic(source, dv, error, W);
		if (!r.g || !pcb_gpu_optic_stream(r.g)) {
			if (r.g) {
				gpu_fail(error, W);
				pcb_gpu_optic_destroy(r.g);
			}
			stream_release(st);
			return NULL;
		}
		r.dev = pcb_gpu_optic_device(r.g);
		st->R.push_back(r);
		st->d_mem.push_back(NULL);
	}
	return st;
}

extern "C" void polycap_b200_stream_free(polycap_b200_stream *stream)
{
	if (!stream)
		return;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	stream_release(stream);
}

/* (re)allocates the per-device block: tallies (kept) + a record buffer of rec_cap records */
static bool stream_buffers(polycap_b200_stream *st, int64_t rec_cap)
{
	auto up256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
	const size_t nE = st->nE;
	const size_t o_cnt = up256(nE * 8), o_rf = o_cnt + up256(PCB_C_COUNT * 8);
	if (rec_cap <= st->rec_cap && st->d_mem[0])
		return true;
	for (size_t k = 0; k < st->R.size(); k++) {
		DevRun &r = st->R[k];
		const size_t o_rw = o_rf + up256((size_t)rec_cap * 17 * 8), bytes = o_rw + up256((size_t)rec_cap * nE * 8);
		char *nw = (char *)pcb_gpu_malloc(r.dev, bytes);
		if (!nw)
			return false;
		void *strm = pcb_gpu_optic_stream(r.g);
		pcb_gpu_wave_download_sync(r.g);
		if (st->d_mem[k]) { /* carry the tallies over */
			if (pcb_gpu_sync(r.dev, strm) || pcb_gpu_copy_dev(r.dev, nw, st->d_mem[k], o_rf, strm))
				return false;
			pcb_gpu_sync(r.dev, strm);
			pcb_gpu_free(r.dev, st->d_mem[k]);
		} else {
			pcb_gpu_memset(r.dev, nw, 0, o_rf, strm);
		}
		st->d_mem[k] = nw;
		r.d_sum = (double *)nw;
		r.d_cnt = (int64_t *)(nw + o_cnt);
		r.d_rf = (double *)(nw + o_rf);
		r.d_rw = (double *)(nw + o_rw);
		r.rec_cap = rec_cap;
		r.rec_used = 0;
	}
	st->rec_cap = rec_cap;
	return true;
}

extern "C" int64_t polycap_b200_stream_next(polycap_b200_stream *stream, int64_t max_records, int64_t max_ids, double *fields, double *weights,
					    polycap_error **error)
{
	static const char *W = "polycap_b200_stream_next";
	if (!stream || max_records < 1 || max_ids < 0 || (fields == NULL) != (weights == NULL)) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: stream cannot be NULL, max_records must be >= 1, max_ids >= 0, fields and weights both or neither", W);
		return -1;
	}
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	polycap_b200_stream *st = stream;
	const int G = (int)st->R.size();
	const int64_t wave_bytes_per_id = 160 + 8 * (int64_t)st->nE;
	const int64_t wave_max = std::max<int64_t>(4096, std::min<int64_t>((int64_t)1 << 24, ((int64_t)6 << 30) / wave_bytes_per_id));
	/* tally-only chunks still rank and compact nothing: a one-record buffer is enough for them */
	if (!stream_buffers(st, fields ? max_records : std::max<int64_t>(st->rec_cap, 1))) {
		polycap_set_error(error, POLYCAP_ERROR_MEMORY, "%s: could not allocate device memory for %" PRId64 " records -> %s", W, max_records, pcb_gpu_last_error());
		return -1;
	}
	int64_t have = 0, ids_done = 0, waves = 0;
	while (have < max_records && (max_ids == 0 || ids_done < max_ids)) {
		const int64_t need = max_records - have;
		int64_t n_ids;
		if (st->next_id > 0 && st->exits > 0) {
			const double frac = (double)st->exits / (double)st->next_id;
			const double margin = std::min(0.25, 3.0 * sqrt(1.0 / (double)st->exits + 1.0 / (double)need) + 0.002);
			n_ids = (int64_t)((double)need / frac * (1.0 + margin)) + 256;
		} else {
			n_ids = 2 * need + 1024;
		}
		n_ids = std::min<int64_t>(n_ids, wave_max * G);
		if (max_ids > 0)
			n_ids = std::min<int64_t>(n_ids, max_ids - ids_done);
		int64_t used = 0, got = 0;
		/* tally-only: no cut is needed to bound a buffer, but the id-ordered cut keeps "n exits" semantics identical */
		const int prc = plain_wave(st->R, st->sd, st->seed, st->next_id, &n_ids, need, fields, weights, max_records, have, &used, &got, error, W);
		if (prc) {
			if (prc == 1)
				gpu_fail(error, W);
			for (int k = 0; k < G; k++)
				pcb_gpu_wave_download_sync(st->R[k].g);
			return -1;
		}
		st->next_id += used;
		st->exits += got;
		have += got;
		ids_done += used;
		if (++waves > 100000 || (used == 0 && got == 0)) {
			polycap_set_error(error, POLYCAP_ERROR_RUNTIME, "%s: the run does not advance (%" PRId64 " ids so far)", W, st->next_id);
			return -1;
		}
	}
	for (int k = 0; k < G; k++)
		if (pcb_gpu_wave_download_sync(st->R[k].g)) {
			gpu_fail(error, W);
			return -1;
		}
	return have;
}

extern "C" bool polycap_b200_stream_totals(polycap_b200_stream *stream, int64_t *ids_consumed, double *sum_weights, int64_t *counters, polycap_error **error)
{
	static const char *W = "polycap_b200_stream_totals";
	if (!stream || !sum_weights || !counters) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: NULL argument", W);
		return false;
	}
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	const int G = (int)stream->R.size();
	if (ids_consumed)
		*ids_consumed = stream->next_id;
	if (!stream->d_mem[0]) { /* nothing traced yet */
		memset(sum_weights, 0, stream->nE * 8);
		memset(counters, 0, PCB_C_COUNT * 8);
		return true;
	}
	std::vector<pcb_gpu_optic *> og((size_t)G);
	std::vector<double *> ds((size_t)G);
	std::vector<int64_t *> dc((size_t)G);
	for (int k = 0; k < G; k++) {
		og[k] = stream->R[k].g;
		ds[k] = stream->R[k].d_sum;
		dc[k] = stream->R[k].d_cnt;
	}
	if (pcb_gpu_reduce_tallies(og.data(), ds.data(), dc.data(), G, sum_weights, counters, NULL)) {
		gpu_fail(error, W);
		return false;
	}
	return true;
}

extern "C" void polycap_b200_session_free(polycap_b200_session *session)
{
	if (!session)
		return;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	pcb_gpu_optic_destroy(session->g);
	delete session;
}

extern "C" bool polycap_b200_session_run(polycap_b200_session *session, int64_t id0, int64_t n_ids, int64_t wave_size, void *cuda_stream, double *d_sum_weights,
					 int64_t *d_counters, polycap_error **error)
{
	static const char *W = "polycap_b200_session_run";
	if (!session || !d_sum_weights || !d_counters || n_ids < 0) {
		polycap_set_error(error, POLYCAP_ERROR_INVALID_ARGUMENT, "%s: NULL argument or n_ids < 0", W);
		return false;
	}
	if (wave_size < 1)
		wave_size = (int64_t)1 << 20;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	for (int64_t done = 0; done < n_ids; done += wave_size) {
		int64_t n = std::min(wave_size, n_ids - done);
		if (pcb_gpu_run_wave(session->g, &session->sd, session->seed, id0 + done, n, 0, d_sum_weights, d_counters, NULL, NULL, 0, 0, cuda_stream, NULL,
				     NULL)) {
			gpu_fail(error, W);
			return false;
		}
	}
	return true;
}

extern "C" void polycap_b200_session_tune(polycap_b200_session *session, int thr_cand, int thr_refl, int thr_new, int ctas_per_sm)
{
	if (!session)
		return;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	pcb_gpu_optic_tune(session->g, thr_cand, thr_refl, thr_new, ctas_per_sm, -1);
}

extern "C" void polycap_b200_session_timing(polycap_b200_session *session, int on)
{
	if (session)
		pcb_gpu_optic_timing(session->g, on);
}
extern "C" double polycap_b200_session_trace_ms(polycap_b200_session *session, int64_t *n_launches)
{
	if (!session)
		return 0;
	std::lock_guard<std::mutex> lock(g_gpu_mutex);
	return pcb_gpu_optic_trace_ms(session->g, n_launches);
}
extern "C" int64_t polycap_b200_session_kernels(polycap_b200_session *session, int64_t *n_trace_kernels)
{
	return session ? pcb_gpu_optic_kernels(session->g, n_trace_kernels) : 0;
}

extern "C" int64_t polycap_b200_selftest_division(int device, int64_t n, uint64_t seed, int span) { return pcb_gpu_div_selftest(device, n, seed, span); }
extern "C" double polycap_b200_dfma_peak_tflops(int device, int iters) { return pcb_gpu_dfma_peak(device, iters > 0 ? iters : 200000); }
