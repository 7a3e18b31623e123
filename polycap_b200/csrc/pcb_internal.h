/*
 * pcb_internal.h -- host-side object layouts of libpolycap_b200.
 *
 * The field order of every struct follows /root/reference/src/polycap-private.h:88-181 so
 * that code which pokes at the reference's private structs (its white-box tests do) stays
 * source-compatible.  Nothing here is device code.
 */
#ifndef PCB_INTERNAL_H
#define PCB_INTERNAL_H

#include "polycap.h"
#include <stdio.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define COSPI_6 0.86602540378443864676 /* cos(pi/6), src/polycap-private.h:37 */

#ifdef __cplusplus
extern "C" {
#endif

struct _polycap_profile { /* private.h:88-94 */
	int nmax;
	double *z;
	double *cap;
	double *ext;
};

struct _polycap_description { /* private.h:96-106 */
	double sig_rough;
	int64_t n_cap;
	double open_area;
	unsigned int nelem;
	int *iz;
	double *wi;
	double density;
	polycap_profile *profile;
};

struct _polycap_rng { /* private.h:69-71: one pointer to generator state */
	struct pcb_mt19937 *_rng;
};

struct _polycap_source { /* private.h:108-122 */
	polycap_description *description;
	polycap_rng *rng;
	double d_source;
	double src_x;
	double src_y;
	double src_sigx;
	double src_sigy;
	double src_shiftx;
	double src_shifty;
	double hor_pol;
	size_t n_energies;
	double *energies;
};

struct _polycap_photon { /* private.h:124-145 */
	polycap_description *description;
	polycap_leak **extleak;
	polycap_leak **intleak;
	int64_t n_extleak;
	int64_t n_intleak;
	polycap_vector3 start_coords;
	polycap_vector3 start_direction;
	polycap_vector3 start_electric_vector;
	polycap_vector3 exit_coords;
	polycap_vector3 exit_direction;
	polycap_vector3 exit_electric_vector;
	polycap_vector3 src_start_coords;
	size_t n_energies;
	double *energies;
	double *weight;
	double *amu;
	double *scatf;
	int64_t i_refl;
	double d_travel;
};

struct _polycap_images { /* private.h:156-181 */
	int64_t i_start;
	int64_t i_exit;
	double *src_start_coords[2];
	double *pc_start_coords[2];
	double *pc_start_dir[2];
	double *pc_start_elecv[2];
	double *pc_exit_coords[3];
	double *pc_exit_dir[2];
	double *pc_exit_elecv[2];
	int64_t *pc_exit_nrefl;
	double *pc_exit_dtravel;
	double *exit_coord_weights;
	int64_t i_extleak;
	double *extleak_coords[3];
	double *extleak_dir[2];
	double *extleak_coord_weights;
	int64_t *extleak_n_refl;
	int64_t i_intleak;
	double *intleak_coords[3];
	double *intleak_dir[2];
	double *intleak_elecv[2];
	double *intleak_coord_weights;
	int64_t *intleak_n_refl;
};

struct _polycap_transmission_efficiencies { /* private.h:147-154 */
	size_t n_energies;
	double *energies;
	double *efficiencies;
	struct _polycap_images *images;
	polycap_source *source;
};

/* host MT19937 (GSL conventions: seed 0 -> 4357, uniform = word / 2^32) */
struct pcb_mt19937 {
	uint32_t mt[624];
	int mti;
};
void pcb_mt_seed(struct pcb_mt19937 *r, unsigned long seed);
double pcb_mt_uniform(struct pcb_mt19937 *r);

/* private entry points of the reference that its tests reach through -DTEST_BUILD
 * (private.h:40-49, 81-84, 183-192); exported here with the same names and signatures. */
POLYCAP_EXTERN int polycap_capil_segment(polycap_vector3 cap_coord0, polycap_vector3 cap_coord1, double cap_rad0, double cap_rad1,
					 polycap_vector3 phot_coord0, polycap_vector3 phot_coord1, polycap_vector3 photon_dir,
					 polycap_vector3 *photon_coord, polycap_vector3 *surface_norm, polycap_error **error);
POLYCAP_EXTERN double polycap_refl_polar(double e, double density, double scatf, double lin_abs_coeff, polycap_vector3 surface_norm,
					 polycap_photon *photon, polycap_vector3 *electric_vector, polycap_error **error);
POLYCAP_EXTERN int polycap_capil_reflect(polycap_photon *photon, polycap_vector3 surface_norm, bool leak_calc, polycap_error **error);
POLYCAP_EXTERN int polycap_capil_trace(int *ix, polycap_photon *photon, polycap_description *description, double *cap_x, double *cap_y,
				       bool leak_calc, polycap_error **error);
POLYCAP_EXTERN int polycap_capil_trace_wall(polycap_photon *photon, double *d_travel, int *capx_id, int *capy_id, polycap_error **error);
POLYCAP_EXTERN int polycap_photon_within_pc_boundary(double polycap_radius, polycap_vector3 photon_coord, polycap_error **error);
POLYCAP_EXTERN void polycap_photon_scatf(polycap_photon *photon, polycap_error **error);
POLYCAP_EXTERN void polycap_norm(polycap_vector3 *vect);
POLYCAP_EXTERN double polycap_scalar(polycap_vector3 vect1, polycap_vector3 vect2);
POLYCAP_EXTERN char *polycap_read_input_line(FILE *fptr, polycap_error **error);
POLYCAP_EXTERN void polycap_description_check_weight(size_t nelem, double wi[], polycap_error **error);
POLYCAP_EXTERN polycap_leak *polycap_leak_new(polycap_vector3 leak_coords, polycap_vector3 leak_dir, polycap_vector3 leak_elecv, int64_t n_refl,
					      size_t n_energies, double *weights, polycap_error **error);
POLYCAP_EXTERN double polycap_rng_uniform(const polycap_rng *r);

#ifdef __cplusplus
}
#endif
#endif
