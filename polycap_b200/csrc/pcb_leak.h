/* pcb_leak.h -- host interface of the leak / recap transport kernel (pcb_leak.cu). Internal. */
#ifndef PCB_LEAK_H
#define PCB_LEAK_H
#include <stddef.h>
#include <stdint.h>
#include "pcb_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

/* device-resident optic as the leak kernel sees it (filled by pcb_gpu_optic_leak_view) */
typedef struct {
	int nmax, nE, mono, monotonic, sm_count, device;
	double n_shells, sig_rough, cap0, ext0;
	const double *d_z, *d_cap, *d_ext, *d_hexz, *d_ec;
	const double *d_se, *d_sc, *d_sl, *d_isl; /* wall-march tables per z segment: d ext/dz, d cap/dz, cap/10, 10/cap */
} pcb_leak_optic;

const char *pcb_leak_last_error(void);

/* Runs n photons with leak_calc = true: ids [id0, id0+n) of `src` (src != NULL) or the replay states coords/dir/ev [n][3].
 * Host outputs: status[n], irefl[n], fields [17 or 11][n] (field-major, as the trace kernel writes them), weights [n][nE],
 * leak_n[n] = leak records of each photon, *leak_rec = malloc'd [*n_leak_rec][11+nE] records in photon order:
 * kind (0 ext, 1 int), coords(3), direction(3), elecv(3), n_refl, weights(nE). */
int pcb_leak_run(const pcb_leak_optic *o, const pcb_source_desc *src, uint64_t seed, int64_t id0, int64_t n, const double *coords, const double *dir,
		 const double *ev, int8_t *status, int *irefl, double *fields, double *weights, int *leak_n, double **leak_rec, int64_t *n_leak_rec,
		 int64_t *work /* [4] or NULL: capil_reflect events, wall-march steps (capil.c:1018-1068), steps actually evaluated, rounds of the segment scan */);

/* One polycap_capil_reflect (mode 1) or polycap_capil_trace (mode 2) with leak_calc = true on one photon state:
 * in[16] = coords(3) direction(3) elecv(3) wall normal(3) i_refl d_travel A B (capillary axis: cap_x[i] = A*hexz[i], cap_y[i] = B*hexz[i]),
 * ix = segment index; weights[nE] in/out; outs[13] = return code, coords(3), direction(3), elecv(3), i_refl, d_travel, ix;
 * *leak_rec = malloc'd leak records of the step (layout of pcb_leak_run).  Returns 0, or -1 with pcb_leak_last_error(). */
int pcb_leak_unit_step(const pcb_leak_optic *o, int mode, const double in[16], int ix, double *weights, double outs[13], double **leak_rec,
		       int64_t *n_leak_rec);

/* polycap_capil_trace_wall on one photon state in[6] = coords(3), direction(3) */
int pcb_leak_unit_trace_wall(const pcb_leak_optic *o, const double in[6], double *d_travel, int *r_cntr, int *q_cntr);

#ifdef __cplusplus
}
#endif
#endif
