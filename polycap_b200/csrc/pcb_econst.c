/*
 * pcb_econst.c -- per-energy optical constants of the Fresnel term, host side, plain C99.
 *
 * The reference recomputes, for every reflection and every energy, the complex refractive
 * index n = 1 - alfa + i*beta and its inverse (/root/reference/src/polycap-capil.c:493-495,499).
 * They depend on the energy only, so they are computed once per run here -- in C99 complex
 * arithmetic and in the reference's operation order, so the device consumes bit-identical
 * values -- and uploaded as a [6][nE] table.
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>

#define HC 1.23984193E-7
#define N_AVOG 6.022098e+23
#define R0 2.8179403227e-13

/* rows: Re n, Im n, Re (1/n)^2, Im (1/n)^2, 1.01358*E, amu */
void pcb_energy_constants(size_t n_energies, const double *energies, const double *amu, const double *scatf, double density, double *out)
{
	for (size_t i = 0; i < n_energies; i++) {
		double e = energies[i];
		double alfa = (HC / e) * (HC / e) * ((N_AVOG * R0 * density) / (2 * M_PI)) * scatf[i];
		double beta = (HC) / (4. * M_PI) * (amu[i] / e);
		double complex n = (1.0 - alfa + I * beta);
		double complex n_inv = 1.0 / (n);
		double complex i2 = (n_inv) * (n_inv);
		out[0 * n_energies + i] = creal(n);
		out[1 * n_energies + i] = cimag(n);
		out[2 * n_energies + i] = creal(i2);
		out[3 * n_energies + i] = cimag(i2);
		out[4 * n_energies + i] = (1.01358e0 * e);
		out[5 * n_energies + i] = amu[i];
	}
}
