/*
 * include/polycap.h -- the drop-in C boundary of polycap-b200.
 *
 * One self-contained header declaring the complete public C API of polycap 1.2, so that the
 * reference's CLI (src/main.c), its Cython wrapper (python .pxd files) and user code compile and
 * link against libpolycap_b200.so unchanged.  Every prototype below cites the reference
 * declaration it replaces (paths relative to the reference tree).  Signatures, return codes,
 * ownership and error conventions are those of the reference; what differs is that the
 * photon-transport hot path behind polycap_photon_launch() and
 * polycap_source_get_transmission_efficiencies() runs as CUDA kernels on a B200 and fails
 * with POLYCAP_ERROR_RUNTIME when no CUDA device is usable (there is no CPU fallback).
 *
 * Additive, GPU-specific entry points (batched replay, tally-only runs, device-resident
 * tallies for multi-GPU reductions) live in include/polycap_b200.h.
 */
#ifndef POLYCAP_H
#define POLYCAP_H

#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifndef POLYCAP_EXTERN
#define POLYCAP_EXTERN __attribute__((visibility("default"))) extern
#endif

#define POLYCAP_VERSION_MAJOR 1 /* include/polycap.h:27 */
#define POLYCAP_VERSION_MINOR 2 /* include/polycap.h:31 */

/* physical constants, include/polycap.h:47-50 */
#define HC 1.23984193E-7      /* h*c [keV*cm] */
#define N_AVOG 6.022098e+23   /* Avogadro constant */
#define R0 2.8179403227e-13   /* classical electron radius [cm] */
#define EPSILON 1.0e-30

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ errors
 * include/polycap-error.h:66-171.  GError-style: pass NULL to ignore, or the address of a
 * NULL-initialised pointer; it is set only on failure and freed with polycap_error_free. */
enum polycap_error_code {
	POLYCAP_ERROR_MEMORY,
	POLYCAP_ERROR_INVALID_ARGUMENT,
	POLYCAP_ERROR_IO,
	POLYCAP_ERROR_OPENMP,
	POLYCAP_ERROR_TYPE,
	POLYCAP_ERROR_UNSUPPORTED,
	POLYCAP_ERROR_RUNTIME,
};

typedef struct _polycap_error {
	enum polycap_error_code code;
	char *message;
} polycap_error;

#if defined(__GNUC__)
#define POLYCAP_PRINTF(fmt, args) __attribute__((format(printf, fmt, args)))
#else
#define POLYCAP_PRINTF(fmt, args)
#endif

POLYCAP_EXTERN polycap_error *polycap_error_new(enum polycap_error_code code, const char *format, ...) POLYCAP_PRINTF(2, 3);
POLYCAP_EXTERN polycap_error *polycap_error_new_literal(enum polycap_error_code code, const char *message);
POLYCAP_EXTERN polycap_error *polycap_error_new_valist(enum polycap_error_code code, const char *format, va_list args) POLYCAP_PRINTF(2, 0);
POLYCAP_EXTERN void polycap_error_free(polycap_error *error);
POLYCAP_EXTERN polycap_error *polycap_error_copy(const polycap_error *error);
POLYCAP_EXTERN bool polycap_error_matches(const polycap_error *error, enum polycap_error_code code);
POLYCAP_EXTERN void polycap_set_error(polycap_error **err, enum polycap_error_code code, const char *format, ...) POLYCAP_PRINTF(3, 4);
POLYCAP_EXTERN void polycap_set_error_literal(polycap_error **err, enum polycap_error_code code, const char *message);
POLYCAP_EXTERN void polycap_propagate_error(polycap_error **dest, polycap_error *src);
POLYCAP_EXTERN void polycap_clear_error(polycap_error **err);

/* ------------------------------------------------------------------ profile
 * include/polycap-profile.h:43-160 */
typedef enum {
	POLYCAP_PROFILE_CONICAL,
	POLYCAP_PROFILE_PARABOLOIDAL,
	POLYCAP_PROFILE_ELLIPSOIDAL,
} polycap_profile_type;

typedef struct _polycap_profile polycap_profile;

POLYCAP_EXTERN polycap_profile *polycap_profile_new(polycap_profile_type type, double length, double rad_ext_upstream,
						    double rad_ext_downstream, double rad_int_upstream, double rad_int_downstream,
						    double focal_dist_upstream, double focal_dist_downstream, polycap_error **error); /* :71 */
POLYCAP_EXTERN polycap_profile *polycap_profile_new_from_file(const char *single_cap_profile_file, const char *central_axis_file,
							      const char *external_shape_file, polycap_error **error); /* :94 */
int polycap_profile_validate(polycap_profile *profile, int64_t n_cap, polycap_error **error); /* :107, not exported there either */
POLYCAP_EXTERN polycap_profile *polycap_profile_new_from_arrays(int nid, double *ext, double *cap, double *z, polycap_error **error); /* :120 */
POLYCAP_EXTERN bool polycap_profile_get_ext(polycap_profile *profile, size_t *nid, double **ext, polycap_error **error);
POLYCAP_EXTERN bool polycap_profile_get_cap(polycap_profile *profile, size_t *nid, double **cap, polycap_error **error);
POLYCAP_EXTERN bool polycap_profile_get_z(polycap_profile *profile, size_t *nid, double **z, polycap_error **error);
POLYCAP_EXTERN void polycap_profile_free(polycap_profile *profile);

/* ------------------------------------------------------------------ description
 * include/polycap-description.h:35-76 */
typedef struct _polycap_description polycap_description;

POLYCAP_EXTERN polycap_description *polycap_description_new(polycap_profile *profile, double sig_rough, int64_t n_cap, unsigned int nelem,
							    int iz[], double wi[], double density, polycap_error **error); /* :55 */
POLYCAP_EXTERN const polycap_profile *polycap_description_get_profile(polycap_description *description);
POLYCAP_EXTERN void polycap_description_free(polycap_description *description);

/* ------------------------------------------------------------------ rng
 * include/polycap-rng.h:33-59.  Host-side MT19937 stream used by polycap_source_get_photon;
 * whole runs draw from per-photon Philox4x32-10 streams on the device instead. */
typedef struct _polycap_rng polycap_rng;

POLYCAP_EXTERN polycap_rng *polycap_rng_new(void);
POLYCAP_EXTERN polycap_rng *polycap_rng_new_with_seed(unsigned long int seed);
POLYCAP_EXTERN void polycap_rng_free(polycap_rng *rng);

/* ------------------------------------------------------------------ photon
 * include/polycap-photon.h:40-203 */
typedef struct {
	double x, y, z;
} polycap_vector3;

typedef struct _polycap_photon polycap_photon;

typedef struct {
	polycap_vector3 coords;
	polycap_vector3 direction;
	polycap_vector3 elecv;
	size_t n_energies;
	double *weight;
	int64_t n_refl;
} polycap_leak;

POLYCAP_EXTERN polycap_photon *polycap_photon_new(polycap_description *description, polycap_vector3 start_coords,
						  polycap_vector3 start_direction, polycap_vector3 start_electric_vector,
						  polycap_error **error); /* :70 */
/* returns 1 (reached the end), 0 (absorbed), 2 (hit the glass at the entrance), -2 (started outside
 * the optic), -1 (error); *weights is malloc'd (free with polycap_free).  :89-98 */
POLYCAP_EXTERN int polycap_photon_launch(polycap_photon *photon, size_t n_energies, double *energies, double **weights,
					 bool leak_calc, polycap_error **error);
POLYCAP_EXTERN polycap_vector3 polycap_photon_get_start_coords(polycap_photon *photon);
POLYCAP_EXTERN polycap_vector3 polycap_photon_get_start_direction(polycap_photon *photon);
POLYCAP_EXTERN polycap_vector3 polycap_photon_get_start_electric_vector(polycap_photon *photon);
POLYCAP_EXTERN polycap_vector3 polycap_photon_get_exit_coords(polycap_photon *photon);
POLYCAP_EXTERN polycap_vector3 polycap_photon_get_exit_direction(polycap_photon *photon);
POLYCAP_EXTERN polycap_vector3 polycap_photon_get_exit_electric_vector(polycap_photon *photon);
POLYCAP_EXTERN double polycap_photon_get_dtravel(polycap_photon *photon);
POLYCAP_EXTERN int64_t polycap_photon_get_irefl(polycap_photon *photon);
POLYCAP_EXTERN bool polycap_photon_get_extleak_data(polycap_photon *photon, polycap_leak ***leaks, int64_t *n_leaks, polycap_error **error);
POLYCAP_EXTERN bool polycap_photon_get_intleak_data(polycap_photon *photon, polycap_leak ***leaks, int64_t *n_leaks, polycap_error **error);
POLYCAP_EXTERN void polycap_photon_free(polycap_photon *photon);
POLYCAP_EXTERN void polycap_leak_free(polycap_leak *leak);

/* ------------------------------------------------------------------ results
 * include/polycap-transmission-efficiencies.h:30-120 */
typedef struct _polycap_transmission_efficiencies polycap_transmission_efficiencies;
typedef struct _polycap_progress_monitor polycap_progress_monitor; /* include/polycap-progress-monitor.h:22 */

POLYCAP_EXTERN void polycap_transmission_efficiencies_free(polycap_transmission_efficiencies *efficiencies);
POLYCAP_EXTERN bool polycap_transmission_efficiencies_write_hdf5(polycap_transmission_efficiencies *efficiencies, const char *filename,
								 polycap_error **error);
POLYCAP_EXTERN bool polycap_transmission_efficiencies_get_data(polycap_transmission_efficiencies *efficiencies, size_t *n_energies,
							       double **energies_arr, double **efficiencies_arr, polycap_error **error);
POLYCAP_EXTERN bool polycap_transmission_efficiencies_get_extleak_data(polycap_transmission_efficiencies *efficiencies, polycap_leak ***leaks,
								       int64_t *n_leaks, polycap_error **error);
POLYCAP_EXTERN bool polycap_transmission_efficiencies_get_intleak_data(polycap_transmission_efficiencies *efficiencies, polycap_leak ***leaks,
								       int64_t *n_leaks, polycap_error **error);
POLYCAP_EXTERN bool polycap_transmission_efficiencies_get_start_data(polycap_transmission_efficiencies *efficiencies, int64_t *n_start,
								     int64_t *n_exit, polycap_vector3 **start_coords,
								     polycap_vector3 **start_direction, polycap_vector3 **start_elecv,
								     polycap_vector3 **src_start_coords, polycap_error **error);
POLYCAP_EXTERN bool polycap_transmission_efficiencies_get_exit_data(polycap_transmission_efficiencies *efficiencies, int64_t *n_exit,
								    polycap_vector3 **exit_coords, polycap_vector3 **exit_direction,
								    polycap_vector3 **exit_elecv, int64_t **n_refl, double **d_travel,
								    size_t *n_energies, double ***exit_weights, polycap_error **error);

/* ------------------------------------------------------------------ source
 * include/polycap-source.h:36-129 */
typedef struct _polycap_source polycap_source;

POLYCAP_EXTERN polycap_source *polycap_source_new(polycap_description *description, double d_source, double src_x, double src_y,
						  double src_sigx, double src_sigy, double src_shiftx, double src_shifty, double hor_pol,
						  size_t n_energies, double *energies, polycap_error **error); /* :60 */
POLYCAP_EXTERN void polycap_source_free(polycap_source *source);
POLYCAP_EXTERN polycap_photon *polycap_source_get_photon(polycap_source *source, polycap_rng *rng, polycap_error **error); /* :90 */
POLYCAP_EXTERN polycap_source *polycap_source_new_from_file(const char *filename, polycap_error **error); /* :99 */
/* n_photons = photons that must REACH THE EXIT WINDOW (:107); max_threads is accepted and ignored
 * (the trace runs on the GPU); progress_monitor must be NULL.  :114-121 */
POLYCAP_EXTERN polycap_transmission_efficiencies *polycap_source_get_transmission_efficiencies(polycap_source *source, int max_threads,
											       int n_photons, bool leak_calc,
											       polycap_progress_monitor *progress_monitor,
											       polycap_error **error);
POLYCAP_EXTERN const polycap_description *polycap_source_get_description(polycap_source *source);

/* include/polycap.h:60 -- frees anything this library malloc'd for the caller */
POLYCAP_EXTERN void polycap_free(void *data);

/* src/polycap-aux.h:22-26 */
POLYCAP_EXTERN char *polycap_strdup(const char *str);
POLYCAP_EXTERN char *polycap_strndup(const char *str, size_t len);

#ifdef __cplusplus
}
#endif
#endif /* POLYCAP_H */
