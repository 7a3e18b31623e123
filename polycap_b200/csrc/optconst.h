/*
 * optconst.h -- optical-constants provider (plain C, host only).
 *
 * Replaces the three xraylib look-ups the reference performs per photon and per
 * energy (CS_Total, Fi, AtomicWeight; /root/reference/src/polycap-photon.c:83-92)
 * by a once-per-run table.  xraylib is a third-party dependency that is absent
 * from this image (reference pins libxrl >= 4.0.0, CI 4.1.3), so the provider is
 * chosen in this order and the choice is always reported (never silent):
 *
 *   1. dlopen("libxrl.so*") -- the real library, when it is installed;
 *   2. a user table  E[keV]  amu[1/cm]  scatf   (env POLYCAP_B200_OPTCONST_TABLE),
 *      linearly interpolated, applied at the compound level; a header line
 *      "# composition: 8:53.0 14:47.0 density: 2.23" (Z:weight-percent) keys it to one
 *      material, and descriptions of another material are then refused;
 *   3. ONLY ON REQUEST (env POLYCAP_B200_OPTCONST=builtin), a built-in approximate model:
 *      without xraylib, table or that request the transport entry points fail with
 *      POLYCAP_ERROR_UNSUPPORTED -- the reference cannot run without xraylib either.
 *      The model: NIST-XCOM-style mass attenuation nodes
 *      (log-log interpolation, K edges kept as discontinuities) for B, O, Na, Al,
 *      Si and a closed-form anomalous-scattering term, both anchored to the
 *      reference's own 10 keV golden (tests/photon.c:75-76: scatf 0.503696,
 *      amu 42.544635 for SiO2 53/47 wt%, rho 2.23).
 *
 * The hot path itself never calls this: it consumes amu[]/scatf[] arrays.
 */
#ifndef PCB_OPTCONST_H
#define PCB_OPTCONST_H

#ifdef __cplusplus
extern "C" {
#endif

enum pcb_optconst_provider {
	PCB_OPTCONST_NONE = -1, /* no xraylib, no table, built-in model not asked for: transport calls fail with UNSUPPORTED */
	PCB_OPTCONST_BUILTIN = 0,
	PCB_OPTCONST_XRAYLIB = 1,
	PCB_OPTCONST_USER_TABLE = 2
};

/* element-level (built-in model only; used by the oracle's xraylib stand-in) */
double pcb_builtin_cs_total(int Z, double e_kev);   /* cm^2/g, <0 on unsupported Z */
double pcb_builtin_fi(int Z, double e_kev);         /* f'(E) */
double pcb_builtin_atomic_weight(int Z);            /* g/mol, <0 on unsupported Z */

/* compound-level: fills amu[i] (1/cm) and scatf[i] for every energy.
 * Mirrors the accumulation order of polycap_photon_scatf (photon.c:83-92).
 * Returns 0 on success, -1 when an element is not covered by the active provider, -2 when there is no
 * provider at all, -3 when the user table is keyed to another material. */
int pcb_optconst_compound(unsigned int nelem, const int *iz, const double *wi, double density,
			  unsigned long n_energies, const double *energies,
			  double *amu, double *scatf);

/* which provider pcb_optconst_compound() uses in this process */
int pcb_optconst_active_provider(void);
const char *pcb_optconst_provider_name(void);

#ifdef __cplusplus
}
#endif
#endif
