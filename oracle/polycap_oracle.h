/*
 * oracle/polycap_oracle.h -- CPU restatement of polycap's photon-transport hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product (polycap_b200/) never
 * does.  Parity status: PINNED -- every function below is checked (tests/test_oracle.py)
 * against the reference's own sources compiled as oracle/_ref and against the golden
 * vectors of the reference's tests (tests/capil.c, tests/photon.c, tests/source.c).
 *
 * Each function cites the reference file:line (relative to /root/reference) it follows.
 * Floating-point expressions keep the reference's association order on purpose: the
 * restatement is compared bit-for-bit with oracle/_ref.
 */
#ifndef POLYCAP_ORACLE_H
#define POLYCAP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double x, y, z; } orc_vec3;

/* read-only description of optic + material + energy grid (src/polycap-private.h:88-106) */
typedef struct {
	int nmax;              /* arrays hold nmax+1 entries */
	const double *z, *cap, *ext;
	double sig_rough;
	int64_t n_cap;
	double density;
	size_t n_energies;
	const double *energies, *amu, *scatf; /* amu/scatf: polycap_photon_scatf output */
} orc_optic;

/* leak record (include/polycap-photon.h:52-59) */
typedef struct {
	orc_vec3 coords, direction, elecv;
	int64_t n_refl;
	double *weight;
} orc_leak;

/* per-photon state (src/polycap-private.h:124-145) */
typedef struct {
	orc_vec3 start_coords, start_direction, start_electric_vector;
	orc_vec3 exit_coords, exit_direction, exit_electric_vector;
	orc_vec3 src_start_coords;
	int64_t i_refl;
	double d_travel;
	double *weight; /* [n_energies], owned by caller */
	orc_leak *extleak, *intleak;
	int64_t n_extleak, n_intleak;
} orc_photon;

/* source (src/polycap-private.h:108-122) */
typedef struct {
	double d_source, src_x, src_y, src_sigx, src_sigy, src_shiftx, src_shifty, hor_pol;
} orc_source;

/* uniform [0,1) generator handed to orc_source_get_photon */
typedef double (*orc_uniform_fn)(void *state);

/* --- helpers: src/polycap-photon.c:139-169, 365-386 --- */
int orc_within_pc_boundary(double polycap_radius, orc_vec3 p);
void orc_norm(orc_vec3 *v);
double orc_scalar(orc_vec3 a, orc_vec3 b);
double orc_n_shells(int64_t n_cap);

/* --- src/polycap-capil.c:52-255 --- */
int orc_capil_segment(orc_vec3 cap0, orc_vec3 cap1, double rad0, double rad1, orc_vec3 ph0, orc_vec3 ph1,
		      orc_vec3 dir, orc_vec3 *photon_coord, orc_vec3 *surface_norm);
/* --- src/polycap-capil.c:444-563 (photon->exit_electric_vector may be renormalised in place) --- */
double orc_refl_polar(double e, double density, double scatf, double amu, orc_vec3 surface_norm,
		      orc_photon *ph, orc_vec3 *electric_vector);
/* --- src/polycap-capil.c:565-891 --- */
int orc_capil_reflect(const orc_optic *o, orc_photon *ph, orc_vec3 surface_norm, int leak_calc);
/* --- src/polycap-capil.c:896-1194 --- */
int orc_capil_trace_wall(const orc_optic *o, orc_photon *ph, double *d_travel, int *r_cntr, int *q_cntr);
/* --- src/polycap-photon.c:174-362; returns 1 and fills *out when a point is produced --- */
int orc_pc_intersect(const orc_optic *o, orc_vec3 coord, orc_vec3 dir, orc_vec3 *out);
/* --- src/polycap-capil.c:1197-1361 --- */
int orc_capil_trace(const orc_optic *o, int *ix, orc_photon *ph, const double *cap_x, const double *cap_y, int leak_calc);
/* --- src/polycap-photon.c:390-955; ph->weight must hold n_energies doubles; returns 1/0/2/-2/-1 --- */
int orc_photon_launch(const orc_optic *o, orc_photon *ph, int leak_calc);
/* --- src/polycap-source.c:23-144 --- */
void orc_source_get_photon(const orc_source *s, const orc_optic *o, orc_uniform_fn uni, void *state, orc_photon *ph);

void orc_photon_clear_leaks(orc_photon *ph);

/* --- random streams --- */
typedef struct { uint32_t mt[624]; int mti; } orc_mt19937;
void orc_mt_seed(orc_mt19937 *r, unsigned long seed);      /* GSL seeding (seed 0 -> 4357) */
double orc_mt_uniform(void *r);                             /* next word / 2^32 */

/* Philox4x32-10, counter = (id_lo, id_hi, block, 0), key = (seed_lo, seed_hi);
 * draw d of photon `id` = 53 high bits of 64-bit half (d&1) of block d>>1. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
typedef struct { uint64_t seed, id; uint32_t draw; } orc_philox_stream;
double orc_philox_uniform(void *stream);

/* --- driver: src/polycap-source.c:745-1075 restated over photon ids, Philox stream per id --- */
typedef struct {
	int64_t n_ids;            /* photon ids [id0, id0+n_ids) were drawn */
	int64_t n_exit, n_not_entered, n_not_transmitted, n_uncounted;
	int64_t sum_irefl_exit;   /* reflections of exit photons (what the reference prints) */
	int64_t sum_irefl_all;    /* reflection events of every photon */
	int64_t n_segments;       /* capil_segment calls */
} orc_counters;

/* Runs ids [id0, id0+n_ids); sum_weights[nE] accumulates exit-photon weights; status[i] gets the
 * final classification per id (1 exit, 0 absorbed, 2 not entered, -2/-1 uncounted, 3 = returned 1 but
 * missed the exit window); if rec != NULL it must hold n_ids*(17+nE) doubles and receives, for exit
 * photons only and at the slot of their id, the 17 record values of struct _polycap_images followed
 * by the weights. Uses OpenMP over ids when threads > 1. */
void orc_run_ids(const orc_source *s, const orc_optic *o, uint64_t seed, int64_t id0, int64_t n_ids,
		 int leak_calc, int threads, double *sum_weights, orc_counters *cnt, int8_t *status, double *rec);

/* the same driver over caller-supplied start states starts[n][11] = coords(3) dir(3) elecv(3) src_x src_y */
void orc_run_starts(const orc_optic *o, int64_t n, const double *starts, int leak_calc, int threads, double *sum_weights, orc_counters *cnt,
		    int8_t *status, double *rec);

#ifdef __cplusplus
}
#endif
#endif
