/* pcb_unit.h -- single-item device entry points (see pcb_unit.cu). Internal. */
#ifndef PCB_UNIT_H
#define PCB_UNIT_H
#ifdef __cplusplus
extern "C" {
#endif
const char *pcb_unit_last_error(void);
/* all return <= -100 on a CUDA failure */
int pcb_unit_within(double R, double x, double y);
/* in: cap0(3) cap1(3) rad0 rad1 phot0(3) phot1(3) dir(3) photon_coord(3); out: hit(3) normal(3) */
int pcb_unit_segment(const double in[20], double out[6]);
/* in: normal(3) dir(3) elecv(3) n_re n_im i2_re i2_im; out: rtot new_ev(3) renormalised elecv(3) */
int pcb_unit_refl_polar(const double in[13], double out[7]);
/* in: dir(3) elecv(3) normal(3); ec: [6][nE]; w in/out; out: dir(3) elecv(3); returns iesc */
int pcb_unit_reflect(const double in[9], const double *ec, int nE, double sig_rough, double *w, double out[6]);
/* in: coords(3) dir(3) elecv(3) d_travel; out: coords(3) dir(3) elecv(3) d_travel; returns capil_trace code */
int pcb_unit_trace(const double in[10], int *ix, long long *irefl, const double *z, const double *cap, const double *ext, const double *cap_x,
		   const double *cap_y, int nmax, int mono, const double *ec, int nE, double sig_rough, double *w, double out[10]);
/* src: d_source src_x src_y sigx sigy shiftx shifty hor_pol cap0 ext0; out: coords(3) dir(3) ev(3) src_x src_y */
int pcb_unit_get_photon(const double src[10], int mono, const double *u, int nu, double out[11], int *used);
#ifdef __cplusplus
}
#endif
#endif
