/*
 * oracle/polycap_oracle.c -- CPU restatement of polycap's photon-transport hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see polycap_oracle.h).  The product never links this file.
 * Citations are relative to /root/reference.  Expressions keep the reference's
 * association order so results can be compared bit-for-bit with oracle/_ref.
 */
#include "polycap_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_HC 1.23984193E-7      /* include/polycap.h:47 */
#define ORC_N_AVOG 6.022098e+23   /* include/polycap.h:48 */
#define ORC_R0 2.8179403227e-13   /* include/polycap.h:49 */
#define ORC_COSPI_6 0.86602540378443864676 /* src/polycap-private.h:37 */

static _Thread_local int64_t t_segment_calls; /* work counter for the flop model */
static _Thread_local int64_t t_reflect_calls;

/* ------------------------------------------------------------------ helpers */

void orc_norm(orc_vec3 *v) /* photon.c:365-376 */
{
	double len = sqrt(v->x * v->x + v->y * v->y + v->z * v->z);
	v->x /= len;
	v->y /= len;
	v->z /= len;
}

double orc_scalar(orc_vec3 a, orc_vec3 b) /* photon.c:379-386 */
{
	return a.x * b.x + a.y * b.y + a.z * b.z;
}

double orc_n_shells(int64_t n_cap) /* photon.c:483 */
{
	return round(sqrt(12. * n_cap - 3.) / 6. - 0.5);
}

int orc_within_pc_boundary(double R, orc_vec3 p) /* photon.c:139-169 */
{
	if (R <= 0.)
		return -1;
	double apothem = sqrt((R * R) - ((R / 2.) * (R / 2.)));
	double d1 = fabs(0 * p.x + 1 * p.y);
	double d2 = fabs(ORC_COSPI_6 * p.x + 0.5 * p.y);
	double d3 = fabs(ORC_COSPI_6 * p.x + -0.5 * p.y);
	return (d1 > apothem || d2 > apothem || d3 > apothem) ? 0 : 1;
}

/* last index i < nmax with z[i] <= zq (0 when none); photon.c:508-509, capil.c:923-926 */
static int z_index(const orc_optic *o, double zq)
{
	int id = 0;
	for (int i = 0; i < o->nmax; i++)
		if (o->z[i] <= zq)
			id = i;
	return id;
}

/* hexagonal (q,r) cell of a point, cube rounding; photon.c:540-552, capil.c:961-973 */
static void hex_cell(double hexz, double x, double y, double *q_out, double *r_out)
{
	double r = y * (2. / 3) / hexz;
	double q = (x / (2. * ORC_COSPI_6) - y / 3) / hexz;
	if (fabs(q - round(q)) > fabs(r - round(r)) && fabs(q - round(q)) > fabs(-1. * q - r - round(-1. * q - r))) {
		q = -1. * round(r) - round(-1. * q - r);
		r = round(r);
	} else if (fabs(r - round(r)) > fabs(-1. * q - r - round(-1. * q - r))) {
		r = -1. * round(q) - round(-1. * q - r);
		q = round(q);
	} else {
		q = round(q);
		r = round(r);
	}
	*q_out = q;
	*r_out = r;
}

static orc_vec3 propagate_to_z(orc_vec3 p, orc_vec3 d, double zt)
{
	orc_vec3 r;
	r.x = p.x + d.x * (zt - p.z) / d.z;
	r.y = p.y + d.y * (zt - p.z) / d.z;
	r.z = zt;
	return r;
}

/* ------------------------------------------------------------------ segment */

int orc_capil_segment(orc_vec3 cap0, orc_vec3 cap1, double rad0, double rad1, orc_vec3 ph0, orc_vec3 ph1,
		      orc_vec3 dir, orc_vec3 *photon_coord, orc_vec3 *surface_norm) /* capil.c:52-255 */
{
	t_segment_calls++;
	if (cap0.z < 0. || cap1.z < 0 || rad0 < 0. || rad1 < 0. || photon_coord == NULL || dir.z < 0 || surface_norm == NULL)
		return -1;
	if (cap0.z != ph0.z || cap1.z != ph1.z)
		return -1;
	if (cap1.z <= cap0.z)
		return -1;

	surface_norm->x = surface_norm->y = surface_norm->z = 0.0;
	orc_norm(&dir);

	orc_vec3 axis = {cap1.x - cap0.x, cap1.y - cap0.y, cap1.z - cap0.z};
	double axis_len = sqrt(orc_scalar(axis, axis));

	/* quadratic in dz: |photon - axis|^2 = radius^2   (capil.c:119-124) */
	double ux = (dir.x / dir.z) - (axis.x / axis.z);
	double uy = (dir.y / dir.z) - (axis.y / axis.z);
	double sl = (rad1 - rad0) / (cap1.z - cap0.z);
	double a = (ux * ux + uy * uy - sl * sl);
	double b = (2. * (ph0.x - cap0.x) * ux + 2. * (ph0.y - cap0.y) * uy - 2. * rad0 * sl);
	double c = ((ph0.x - cap0.x) * (ph0.x - cap0.x) + (ph0.y - cap0.y) * (ph0.y - cap0.y) - rad0 * rad0);
	double discr = b * b - 4. * a * c;
	double zi;
	if (discr < 0)
		return -2;
	if (discr == 0) {
		double d1 = (-1. * b) / (2. * a);
		zi = ph0.z + d1;
	} else {
		double d1 = (-1. * b + sqrt(discr)) / (2. * a);
		double d2 = (-1. * b - sqrt(discr)) / (2. * a);
		/* a root is unusable when it lies outside the segment or is not >= 1e-5 beyond the last hit (capil.c:134-156) */
		int bad1 = (ph0.z + d1 < cap0.z || ph0.z + d1 - photon_coord->z < 1.e-5 || ph0.z + d1 > cap1.z);
		int bad2 = (ph0.z + d2 < cap0.z || ph0.z + d2 - photon_coord->z < 1.e-5 || ph0.z + d2 > cap1.z);
		if (bad1) {
			if (bad2)
				return -3;
			zi = ph0.z + d2;
		} else if (bad2) {
			zi = ph0.z + d1;
		} else if (ph0.z + d2 - photon_coord->z < ph0.z + d1 - photon_coord->z) {
			zi = ph0.z + d2;
		} else {
			zi = ph0.z + d1;
		}
	}
	if (zi > cap1.z)
		return -4;
	if (zi < cap0.z || zi - photon_coord->z < 1.e-5)
		return -5;

	double d_proj = (zi - ph0.z) / dir.z;
	if (d_proj < 1.e-10)
		return -6;
	orc_vec3 hit = {ph0.x + d_proj * dir.x, ph0.y + d_proj * dir.y, zi};

	/* foot point on the capillary axis, then the normal tilted by the wall taper (capil.c:225-246) */
	orc_vec3 rel = {ph0.x - cap0.x, ph0.y - cap0.y, ph0.z - cap0.z};
	double t = ((d_proj + (orc_scalar(rel, axis) / orc_scalar(dir, axis))) / (orc_scalar(axis, axis) / orc_scalar(dir, axis)));
	orc_vec3 foot = {cap0.x + t * axis.x, cap0.y + t * axis.y, cap0.z + t * axis.z};
	orc_vec3 radial = {hit.x - foot.x, hit.y - foot.y, hit.z - foot.z};
	double radial_len = sqrt(orc_scalar(radial, radial));
	double tga = (rad0 - rad1) / axis_len;
	double gam = atan(tga);
	double sga = sin(gam);
	double cga = cos(gam);
	surface_norm->x = cga * radial.x / radial_len + sga * axis.x / axis_len;
	surface_norm->y = cga * radial.y / radial_len + sga * axis.y / axis_len;
	surface_norm->z = cga * radial.z / radial_len + sga * axis.z / axis_len;
	orc_norm(surface_norm);

	*photon_coord = hit;
	return 1;
}

/* ------------------------------------------------------------------ Fresnel */

double orc_refl_polar(double e, double density, double scatf, double amu, orc_vec3 sn, orc_photon *ph,
		      orc_vec3 *electric_vector) /* capil.c:444-563 */
{
	if (e < 1. || e > 100. || density <= 0. || scatf < 0. || amu < 0. || ph == NULL)
		return -1;

	if (sqrt(sn.x * sn.x + sn.y * sn.y + sn.z * sn.z) != 1)
		orc_norm(&sn);
	double theta = acos(orc_scalar(sn, ph->exit_direction));
	if (theta < 0.)
		return -1;
	orc_vec3 *ev = &ph->exit_electric_vector;
	if (sqrt(ev->x * ev->x + ev->y * ev->y + ev->z * ev->z) != 1)
		orc_norm(ev);

	double alfa = (ORC_HC / e) * (ORC_HC / e) * ((ORC_N_AVOG * ORC_R0 * density) / (2 * M_PI)) * scatf;
	double beta = (ORC_HC) / (4. * M_PI) * (amu / e);
	double complex n = (1.0 - alfa + I * beta);

	double cos_theta = cos(theta);
	double sin_theta = sin(theta);
	double complex n_inv = 1.0 / (n);
	double complex tmp = ((n_inv) * (n_inv)) * (sin_theta * sin_theta);
	double complex root = csqrt((1.0 - creal(tmp) + I * -1.0 * cimag(tmp)));

	tmp = (n) * (root);
	double complex r_s = (cos_theta - creal(tmp) + I * -1.0 * cimag(tmp)) * (1.0 / (cos_theta + creal(tmp) + I * cimag(tmp)));
	double rs2 = cabs(r_s);
	rs2 *= rs2;

	tmp = (n) * (cos_theta);
	/* The reference builds these two complex numbers with an unparenthesised macro
	 * (capil.c:35 `real + I * imag`, used at capil.c:509), so the second imaginary term lands in the
	 * REAL part: num = (Re root - Re tmp - Im tmp) + i Im root, den = (Re root + Re tmp + Im tmp) + i Im root.
	 * Reproduced on purpose (bug-compatible). */
	double complex r_p = (creal(root) - creal(tmp) + I * cimag(root) - cimag(tmp)) *
			     (1.0 / (creal(root) + creal(tmp) + I * cimag(root) + cimag(tmp)));
	double rp2 = cabs(r_p);
	rp2 *= rp2;

	/* s = normal x direction, p = direction x s  (capil.c:516-524) */
	orc_vec3 d = ph->exit_direction, s, p;
	s.x = sn.y * d.z - d.y * sn.z;
	s.y = sn.z * d.x - d.z * sn.x;
	s.z = sn.x * d.y - d.x * sn.y;
	orc_norm(&s);
	p.x = d.y * s.z - s.y * d.z;
	p.y = d.z * s.x - s.z * d.x;
	p.z = d.x * s.y - s.x * d.y;
	orc_norm(&p);

	double ca = orc_scalar(*ev, s);
	double frac_s = ca * ca;
	double frac_p = 1. - frac_s;
	double rtot = rs2 * frac_s + rp2 * frac_p;

	double cb = orc_scalar(*ev, sn);
	double cc = orc_scalar(*ev, p);
	electric_vector->x = sqrt((ev->x * ca * frac_s) * (ev->x * ca * frac_s) + (ev->x * cb * frac_p) * (ev->x * cb * frac_p) +
				  (ev->x * cc * frac_p) * (ev->x * cc * frac_p));
	electric_vector->y = sqrt((ev->y * ca * frac_s) * (ev->y * ca * frac_s) + (ev->y * cb * frac_p) * (ev->y * cb * frac_p) +
				  (ev->y * cc * frac_p) * (ev->y * cc * frac_p));
	electric_vector->z = sqrt((ev->z * ca * frac_s) * (ev->z * ca * frac_s) + (ev->z * cb * frac_p) * (ev->z * cb * frac_p) +
				  (ev->z * cc * frac_p) * (ev->z * cc * frac_p));
	orc_norm(electric_vector);
	return rtot;
}

/* ------------------------------------------------------------------ leaks */

static void leak_append(orc_leak **list, int64_t *n, orc_vec3 coords, orc_vec3 dir, orc_vec3 elecv, int64_t n_refl,
			size_t nE, const double *w)
{
	*list = realloc(*list, sizeof(orc_leak) * (size_t)(*n + 1));
	orc_leak *lk = &(*list)[*n];
	lk->coords = coords;
	lk->direction = dir;
	lk->elecv = elecv;
	lk->n_refl = n_refl;
	lk->weight = malloc(sizeof(double) * nE);
	memcpy(lk->weight, w, sizeof(double) * nE);
	(*n)++;
}

void orc_photon_clear_leaks(orc_photon *ph)
{
	for (int64_t i = 0; i < ph->n_extleak; i++)
		free(ph->extleak[i].weight);
	for (int64_t i = 0; i < ph->n_intleak; i++)
		free(ph->intleak[i].weight);
	free(ph->extleak);
	free(ph->intleak);
	ph->extleak = ph->intleak = NULL;
	ph->n_extleak = ph->n_intleak = 0;
}

int orc_pc_intersect(const orc_optic *o, orc_vec3 coord, orc_vec3 direction, orc_vec3 *out) /* photon.c:174-362 */
{
	if (direction.z == 0.)
		return 0;
	orc_vec3 bd = {-1. * direction.x, -1. * direction.y, -1. * direction.z};
	orc_norm(&bd);

	int z_id = z_index(o, coord.z);
	double ext_here = (o->ext[z_id + 1] - o->ext[z_id]) / (o->z[z_id + 1] - o->z[z_id]) * (coord.z - o->z[z_id]) + o->ext[z_id];
	if (orc_within_pc_boundary(ext_here, coord) == 1)
		return 0;
	int step;
	if (bd.z < 0.) {
		z_id = z_id + 1;
		step = -1;
	} else {
		step = 1;
	}
	int found = 0;
	orc_vec3 probe;
	do {
		z_id += step;
		if (z_id < 0 || z_id > o->nmax) /* the reference reads out of bounds here; stop instead */
			break;
		probe.x = coord.x + bd.x * (o->z[z_id] - coord.z) / bd.z;
		probe.y = coord.y + bd.y * (o->z[z_id] - coord.z) / bd.z;
		probe.z = o->z[z_id];
		if (orc_within_pc_boundary(ext_here, coord) != orc_within_pc_boundary(o->ext[z_id], probe)) {
			found = 1;
			break;
		}
	} while (z_id >= 0 && z_id <= o->nmax);
	if (!found)
		return 0;
	int zb = z_id, ze = z_id - step;
	if (ze < 0 || ze > o->nmax)
		return 0;
	orc_vec3 beg, end;
	beg.x = coord.x + bd.x * (o->z[zb] - coord.z) / bd.z;
	beg.y = coord.y + bd.y * (o->z[zb] - coord.z) / bd.z;
	beg.z = coord.z + bd.z * (o->z[zb] - coord.z) / bd.z;
	end.x = coord.x + bd.x * (o->z[ze] - coord.z) / bd.z;
	end.y = coord.y + bd.y * (o->z[ze] - coord.z) / bd.z;
	end.z = coord.z + bd.z * (o->z[ze] - coord.z) / bd.z;

	double hb = sqrt((o->ext[zb] * o->ext[zb]) - ((o->ext[zb] / 2.) * (o->ext[zb] / 2.)));
	double he = sqrt((o->ext[ze] * o->ext[ze]) - ((o->ext[ze] / 2.) * (o->ext[ze] / 2.)));
	double pb[3], pe[3], zs[3];
	pb[0] = fabs(0 * beg.x + 1 * beg.y);
	pb[1] = fabs(ORC_COSPI_6 * beg.x + 0.5 * beg.y);
	pb[2] = fabs(ORC_COSPI_6 * beg.x + -0.5 * beg.y);
	pe[0] = fabs(0 * end.x + 1 * end.y);
	pe[1] = fabs(ORC_COSPI_6 * end.x + 0.5 * end.y);
	pe[2] = fabs(ORC_COSPI_6 * end.x + -0.5 * end.y);
	for (int k = 0; k < 3; k++) /* photon.c:261-263 (interpolates with ext, as the reference does) */
		zs[k] = (pb[k] - hb) / (hb - he - pb[k] + pe[k]) * (o->ext[zb] - o->ext[ze]) + o->ext[zb];

	/* viable roots lie between z[zb] and z[ze]; walking backwards keep the largest, forwards the smallest */
	double lo = (step < 0) ? o->z[zb] : o->z[ze];
	double hi = (step < 0) ? o->z[ze] : o->z[zb];
	int have = 0;
	double zf = 0;
	for (int k = 0; k < 3; k++) {
		if (!(zs[k] >= lo && zs[k] <= hi))
			continue;
		if (!have) {
			zf = zs[k];
			have = 1;
		} else if (step < 0 ? (zs[k] > zf) : (zs[k] < zf)) {
			zf = zs[k];
		}
	}
	if (!have) {
		*out = end;
		return 1;
	}
	out->x = coord.x + bd.x * (zf - coord.z) / bd.z;
	out->y = coord.y + bd.y * (zf - coord.z) / bd.z;
	out->z = coord.z + bd.z * (zf - coord.z) / bd.z;
	return 1;
}

/* classification once a wall march ends without a new bore: 3 = out the side, 2 = exit face in glass */
static int wall_exit_class(const orc_optic *o, orc_photon *ph, orc_vec3 fallback, double *d_travel)
{
	int nm = o->nmax;
	orc_vec3 at_exit = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[nm]);
	orc_vec3 rel;
	if (orc_within_pc_boundary(o->ext[nm], at_exit) == 0) {
		orc_vec3 side;
		if (orc_pc_intersect(o, at_exit, ph->exit_direction, &side)) {
			rel.x = side.x - ph->exit_coords.x; rel.y = side.y - ph->exit_coords.y; rel.z = side.z - ph->exit_coords.z;
		} else {
			rel.x = fallback.x - ph->exit_coords.x; rel.y = fallback.y - ph->exit_coords.y; rel.z = fallback.z - ph->exit_coords.z;
		}
		*d_travel = sqrt(orc_scalar(rel, rel));
		return 3;
	}
	rel.x = fallback.x - ph->exit_coords.x; rel.y = fallback.y - ph->exit_coords.y; rel.z = fallback.z - ph->exit_coords.z;
	*d_travel = sqrt(orc_scalar(rel, rel));
	return 2;
}

int orc_capil_trace_wall(const orc_optic *o, orc_photon *ph, double *d_travel, int *r_cntr, int *q_cntr) /* capil.c:896-1194 */
{
	int nm = o->nmax;
	*d_travel = 0.;
	*r_cntr = 0;
	*q_cntr = 0;

	orc_norm(&ph->exit_direction);
	if (ph->exit_coords.z >= o->z[nm])
		return -2;
	int z_id = z_index(o, ph->exit_coords.z);
	double ext_here;
	if (o->z[z_id] != ph->exit_coords.z)
		ext_here = ((o->ext[z_id + 1] - o->ext[z_id]) / (o->z[z_id + 1] - o->z[z_id])) * (ph->exit_coords.z - o->z[z_id]) + o->ext[z_id];
	else
		ext_here = o->ext[z_id];

	orc_vec3 hit = ph->exit_coords, sn;
	double n_shells = orc_n_shells(o->n_cap);
	if (n_shells == 0.) {
		if (sqrt((ph->exit_coords.x) * (ph->exit_coords.x) + (ph->exit_coords.y) * (ph->exit_coords.y)) > ext_here)
			return -2;
	} else if (orc_within_pc_boundary(ext_here, ph->exit_coords) == 0) {
		return -2;
	}

	double hexz = ext_here / (2. * ORC_COSPI_6 * (n_shells + 1));
	double q_i, r_i, q_new = 0, r_new = 0;
	hex_cell(hexz, ph->exit_coords.x, ph->exit_coords.y, &q_i, &r_i);

	int iesc = 0;
	orc_vec3 c0, c1, p0, p1;
	if (n_shells == 0.) { /* capil.c:992-1013 */
		do {
			p0 = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[z_id]);
			p1 = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[z_id + 1]);
			c0.x = 0.; c0.y = 0.; c0.z = o->z[z_id];
			c1.x = 0.; c1.y = 0.; c1.z = o->z[z_id + 1];
			iesc = orc_capil_segment(c0, c1, o->cap[z_id], o->cap[z_id + 1], p0, p1, ph->exit_direction, &hit, &sn);
			z_id++;
		} while (iesc != 1 && z_id < nm - 1);
	} else {
		double dist = 0;
		for (;;) {
			/* march in steps of a tenth of the bore radius until the hex cell changes (capil.c:1018-1068) */
			do {
				dist += o->cap[z_id] / 10.;
				p0.x = ph->exit_coords.x + dist * ph->exit_direction.x;
				p0.y = ph->exit_coords.y + dist * ph->exit_direction.y;
				p0.z = ph->exit_coords.z + dist * ph->exit_direction.z;
				z_id = z_index(o, p0.z);
				ext_here = ((o->ext[z_id + 1] - o->ext[z_id]) / (o->z[z_id + 1] - o->z[z_id])) * (p0.z - o->z[z_id]) + o->ext[z_id];
				double rad_here = ((o->cap[z_id + 1] - o->cap[z_id]) / (o->z[z_id + 1] - o->z[z_id])) * (p0.z - o->z[z_id]) + o->cap[z_id];
				hexz = ext_here / (2. * ORC_COSPI_6 * (n_shells + 1));
				hex_cell(hexz, p0.x, p0.y, &q_new, &r_new);
				c0.y = r_i * (3. / 2) * hexz;
				c0.x = (2. * q_i + r_i) * ORC_COSPI_6 * hexz;
				double d0 = sqrt((p0.x - c0.x) * (p0.x - c0.x) + (p0.y - c0.y) * (p0.y - c0.y));
				if (d0 < rad_here && fabs(q_i) <= n_shells && fabs(r_i) <= n_shells && fabs(-1. * q_i - r_i) <= n_shells) {
					orc_vec3 rel = {p0.x - ph->exit_coords.x, p0.y - ph->exit_coords.y, p0.z - ph->exit_coords.z};
					*d_travel = sqrt(orc_scalar(rel, rel));
					if (*d_travel > 1.e-5) {
						*r_cntr = r_i;
						*q_cntr = q_i;
						return 1;
					}
					*d_travel = 0.;
				}
			} while (q_new == q_i && r_new == r_i && p0.z <= o->z[nm]);

			if (fabs(q_new) > n_shells || fabs(r_new) > n_shells || fabs(-1. * q_new - r_new) > n_shells || p0.z > o->z[nm]) {
				*r_cntr = r_new;
				*q_cntr = q_new;
				return wall_exit_class(o, ph, p0, d_travel);
			}

			/* scan the segments of the new capillary from z_id on (capil.c:1110-1133) */
			iesc = 0;
			do {
				p0 = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[z_id]);
				p1 = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[z_id + 1]);
				hexz = o->ext[z_id] / (2. * ORC_COSPI_6 * (n_shells + 1));
				c0.y = r_new * (3. / 2) * hexz;
				c0.x = (2. * q_new + r_new) * ORC_COSPI_6 * hexz;
				c0.z = o->z[z_id];
				hexz = o->ext[z_id + 1] / (2. * ORC_COSPI_6 * (n_shells + 1));
				c1.y = r_new * (3. / 2) * hexz;
				c1.x = (2. * q_new + r_new) * ORC_COSPI_6 * hexz;
				c1.z = o->z[z_id + 1];
				iesc = orc_capil_segment(c0, c1, o->cap[z_id], o->cap[z_id + 1], p0, p1, ph->exit_direction, &hit, &sn);
				z_id++;
			} while (iesc != 1 && z_id < nm - 1);
			if (z_id >= nm && iesc != 0) { /* capil.c:1134-1140 */
				q_i = q_new;
				r_i = r_new;
				z_id = nm - 1;
				continue;
			}
			break;
		}
	}

	*r_cntr = r_new;
	*q_cntr = q_new;
	if (iesc != 1) {
		orc_vec3 at_exit = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[nm]);
		return wall_exit_class(o, ph, at_exit, d_travel);
	}
	orc_vec3 rel = {hit.x - ph->exit_coords.x, hit.y - ph->exit_coords.y, hit.z - ph->exit_coords.z};
	*d_travel = sqrt(orc_scalar(rel, rel));
	return (z_id >= nm) ? 2 : 1;
}

/* ------------------------------------------------------------------ reflect */

static void capillary_axis(const orc_optic *o, double n_shells, double q, double r, double *cap_x, double *cap_y)
{
	for (int i = 0; i <= o->nmax; i++) {
		double hexz = o->ext[i] / (2. * ORC_COSPI_6 * (n_shells + 1));
		cap_y[i] = r * (3. / 2) * hexz;
		cap_x[i] = (2. * q + r) * ORC_COSPI_6 * hexz;
	}
}

int orc_capil_reflect(const orc_optic *o, orc_photon *ph, orc_vec3 sn, int leak_calc) /* capil.c:565-891 */
{
	t_reflect_calls++;
	size_t nE = o->n_energies;
	orc_norm(&sn);
	orc_norm(&ph->exit_direction);
	double alfa = orc_scalar(ph->exit_direction, sn);
	if (alfa < 0.)
		return -1;

	int wall_trace = 0, r_cntr = 0, q_cntr = 0;
	double d_wall = 0;
	if (leak_calc) {
		wall_trace = orc_capil_trace_wall(o, ph, &d_wall, &r_cntr, &q_cntr);
		if (wall_trace <= 0)
			return -1;
	}

	double *w_leak = malloc(sizeof(double) * nE);
	int leak_flag = 0, weight_flag = 0;
	orc_vec3 new_ev = {0, 0, 0};
	for (size_t i = 0; i < nE; i++) {
		double cons1 = (1.01358e0 * o->energies[i]) * alfa * o->sig_rough;
		double r_rough = exp(-1. * cons1 * cons1);
		double rtot = orc_refl_polar(o->energies[i], o->density, o->scatf[i], o->amu[i], sn, ph, &new_ev);
		if (rtot < 0. || rtot > 1.) {
			free(w_leak);
			return -1;
		}
		if (wall_trace > 0) {
			w_leak[i] = (1. - rtot * r_rough) * ph->weight[i] * exp(-1. * d_wall * o->amu[i]);
			if (w_leak[i] >= 1.e-4)
				leak_flag = 1;
		}
		ph->weight[i] = ph->weight[i] * rtot * r_rough;
		if (ph->weight[i] >= 1.e-4)
			weight_flag = 1;
	}
	int iesc = weight_flag ? 1 : 0;
	ph->exit_electric_vector = new_ev;
	double n_shells = orc_n_shells(o->n_cap);

	if (leak_flag == 1) {
		int nm = o->nmax;
		double step = d_wall / sqrt(orc_scalar(ph->exit_direction, ph->exit_direction));
		orc_vec3 lc = {ph->exit_coords.x + step * ph->exit_direction.x, ph->exit_coords.y + step * ph->exit_direction.y,
			       ph->exit_coords.z + step * ph->exit_direction.z};
		if (wall_trace == 1) {
			int z_id = z_index(o, lc.z);
			double ext_here = ((o->ext[z_id + 1] - o->ext[z_id]) / (o->z[z_id + 1] - o->z[z_id])) * (lc.z - o->z[z_id]) + o->ext[z_id];
			if (n_shells == 0.) {
				if (sqrt(lc.x * lc.x + lc.y * lc.y) >= ext_here)
					wall_trace = 3;
			} else if (orc_within_pc_boundary(ext_here, lc) == 0) {
				wall_trace = 3;
			}
		}
		if (wall_trace == 3)
			leak_append(&ph->extleak, &ph->n_extleak, lc, ph->exit_direction, ph->exit_electric_vector, ph->i_refl, nE, w_leak);
		if (wall_trace == 2)
			leak_append(&ph->intleak, &ph->n_intleak, lc, ph->exit_direction, ph->exit_electric_vector, ph->i_refl, nE, w_leak);
		if (wall_trace == 1 && lc.z < o->z[nm]) {
			/* the transmitted part continues as a secondary photon in capillary (q_cntr, r_cntr) (capil.c:711-886) */
			orc_photon sec;
			memset(&sec, 0, sizeof sec);
			sec.start_coords = sec.exit_coords = lc;
			sec.start_direction = sec.exit_direction = ph->exit_direction;
			sec.start_electric_vector = sec.exit_electric_vector = ph->exit_electric_vector;
			sec.i_refl = ph->i_refl;
			sec.d_travel = ph->d_travel + d_wall;
			sec.weight = w_leak;
			double *cx = malloc(sizeof(double) * (size_t)(nm + 1)), *cy = malloc(sizeof(double) * (size_t)(nm + 1));
			int ix = 0;
			for (int i = 0; i <= nm; i++)
				if (o->z[i] <= sec.exit_coords.z)
					ix = i;
			if (n_shells == 0.) {
				for (int i = 0; i <= nm; i++)
					cx[i] = cy[i] = 0.;
			} else {
				capillary_axis(o, n_shells, q_cntr, r_cntr, cx, cy);
			}
			int esc = 0;
			for (int i = ix; i <= nm; i++) {
				esc = orc_capil_trace(o, &ix, &sec, cx, cy, leak_calc);
				if (esc != 1)
					break;
			}
			if (esc == -1 || esc == -3) {
				orc_photon_clear_leaks(&sec);
				free(cx); free(cy); free(w_leak);
				return -2;
			}
			for (int64_t i = 0; i < sec.n_extleak; i++)
				leak_append(&ph->extleak, &ph->n_extleak, sec.extleak[i].coords, sec.extleak[i].direction, sec.extleak[i].elecv,
					    sec.extleak[i].n_refl, nE, sec.extleak[i].weight);
			for (int64_t i = 0; i < sec.n_intleak; i++)
				leak_append(&ph->intleak, &ph->n_intleak, sec.intleak[i].coords, sec.intleak[i].direction, sec.intleak[i].elecv,
					    sec.intleak[i].n_refl, nE, sec.intleak[i].weight);
			if (esc == 1 || esc == -2) {
				double tt = ((o->z[nm] - sec.exit_coords.z) / sec.exit_direction.z);
				orc_vec3 e = {sec.exit_coords.x + sec.exit_direction.x * tt, sec.exit_coords.y + sec.exit_direction.y * tt,
					      sec.exit_coords.z + sec.exit_direction.z * tt};
				int inside = orc_within_pc_boundary(o->ext[nm], e);
				if (inside == 0)
					leak_append(&ph->extleak, &ph->n_extleak, e, ph->exit_direction, ph->exit_electric_vector, ph->i_refl, nE, sec.weight);
				else if (inside == 1)
					leak_append(&ph->intleak, &ph->n_intleak, e, ph->exit_direction, ph->exit_electric_vector, ph->i_refl, nE, sec.weight);
			}
			orc_photon_clear_leaks(&sec);
			free(cx);
			free(cy);
		}
	}
	free(w_leak);
	return iesc;
}

/* ------------------------------------------------------------------ trace */

int orc_capil_trace(const orc_optic *o, int *ix, orc_photon *ph, const double *cap_x, const double *cap_y, int leak_calc) /* capil.c:1197-1361 */
{
	int nm = o->nmax, iesc = 0;
	orc_norm(&ph->exit_direction);
	orc_norm(&ph->start_direction);
	orc_vec3 last_hit = ph->exit_coords; /* photon_coord: updated by every segment call that returns 1 */
	orc_vec3 dir = ph->exit_direction;
	orc_vec3 sn = {0, 0, 0};
	double cosalfa = 0;
	double n_shells = orc_n_shells(o->n_cap);

	for (int i = *ix; i < nm; i++) {
		orc_vec3 c0 = {cap_x[i], cap_y[i], o->z[i]}, c1 = {cap_x[i + 1], cap_y[i + 1], o->z[i + 1]};
		orc_vec3 p0 = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[i]);
		orc_vec3 p1 = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[i + 1]);
		if ((orc_within_pc_boundary(o->ext[i], c0) == 0) || (orc_within_pc_boundary(o->ext[i + 1], c1) == 0))
			return -1;
		iesc = orc_capil_segment(c0, c1, o->cap[i], o->cap[i + 1], p0, p1, dir, &last_hit, &sn);
		cosalfa = orc_scalar(sn, dir);
		if (acos(cosalfa) > M_PI / 2. || acos(cosalfa) < 0.)
			iesc = -5;
		if (iesc == 1) {
			double ext_here = ((o->ext[i] - o->ext[i + 1]) / (o->z[i] - o->z[i + 1])) * (last_hit.z - o->z[i + 1]) + o->ext[i + 1];
			if (n_shells == 0.) {
				if (sqrt(last_hit.x * last_hit.x + last_hit.y * last_hit.y) >= ext_here)
					return -3;
			} else if (orc_within_pc_boundary(ext_here, last_hit) == 0) {
				return -3;
			}
			*ix = i + 1;
			break;
		} else {
			orc_vec3 here = propagate_to_z(ph->exit_coords, ph->exit_direction, o->z[i]);
			if (orc_within_pc_boundary(o->ext[i], here) == 0)
				return -3;
		}
	}
	if (iesc != 1)
		return -2;

	orc_vec3 rel = {last_hit.x - ph->exit_coords.x, last_hit.y - ph->exit_coords.y, last_hit.z - ph->exit_coords.z};
	ph->d_travel += sqrt(orc_scalar(rel, rel));
	ph->exit_coords = last_hit;
	if (fabs(cosalfa) > 1.0)
		return -1;
	*ix = z_index(o, ph->exit_coords.z);
	double ext_here = ((o->ext[(*ix) + 1] - o->ext[(*ix)]) / (o->z[(*ix) + 1] - o->z[(*ix)])) * (last_hit.z - o->z[(*ix)]) + o->ext[(*ix)];
	if (n_shells == 0 && ph->exit_coords.x * ph->exit_coords.x + ph->exit_coords.y * ph->exit_coords.y >= ext_here)
		return -3;
	if (n_shells > 0 && orc_within_pc_boundary(ext_here, ph->exit_coords) == 0)
		return -3;
	iesc = orc_capil_reflect(o, ph, sn, leak_calc);
	if (iesc == 1) {
		ph->exit_direction.x = ph->exit_direction.x - 2.0 * cosalfa * sn.x;
		ph->exit_direction.y = ph->exit_direction.y - 2.0 * cosalfa * sn.y;
		ph->exit_direction.z = ph->exit_direction.z - 2.0 * cosalfa * sn.z;
		orc_norm(&ph->exit_direction);
		ph->i_refl++;
	} else if (iesc == -1 || iesc == -2) {
		iesc = -1;
	}
	return iesc;
}

/* ------------------------------------------------------------------ launch */

int orc_photon_launch(const orc_optic *o, orc_photon *ph, int leak_calc) /* photon.c:390-955 */
{
	int nm = o->nmax;
	size_t nE = o->n_energies;
	for (size_t i = 0; i < nE; i++)
		if (o->energies[i] < 1. || o->energies[i] > 100.)
			return -1;
	orc_photon_clear_leaks(ph);
	for (size_t i = 0; i < nE; i++)
		ph->weight[i] = 1.;
	ph->i_refl = 0;

	double n_shells = orc_n_shells(o->n_cap);
	orc_norm(&ph->start_direction);
	ph->exit_coords = ph->start_coords;
	ph->exit_direction = ph->start_direction;
	orc_norm(&ph->exit_direction);

	int z_id = 0;
	if (ph->start_coords.z > 0)
		z_id = z_index(o, ph->start_coords.z);
	double ext_here = ((o->ext[z_id] - o->ext[z_id + 1]) / (o->z[z_id] - o->z[z_id + 1])) * (ph->start_coords.z - o->z[z_id]) + o->ext[z_id];

	double q_i = 0, r_i = 0;
	if (n_shells == 0.) {
		if (sqrt((ph->start_coords.x) * (ph->start_coords.x) + (ph->start_coords.y) * (ph->start_coords.y)) > ext_here)
			return -2;
	} else {
		double hexz = ext_here / (2. * ORC_COSPI_6 * (n_shells + 1));
		hex_cell(hexz, ph->start_coords.x, ph->start_coords.y, &q_i, &r_i);
		if (orc_within_pc_boundary(ext_here, ph->start_coords) == 0)
			return -2;
	}

	double *cap_x = malloc(sizeof(double) * (size_t)(nm + 1)), *cap_y = malloc(sizeof(double) * (size_t)(nm + 1));
	int ix = 0;
	capillary_axis(o, n_shells, q_i, r_i, cap_x, cap_y);
	for (int i = 0; i <= nm; i++)
		if (o->z[i] <= ph->start_coords.z)
			ix = i;

	double rad_here, cx_here, cy_here;
	if (ph->start_coords.z > 0) {
		double dz = (o->z[z_id + 1] - o->z[z_id]), off = (ph->start_coords.z - o->z[z_id]);
		rad_here = ((o->cap[z_id + 1] - o->cap[z_id]) / dz) * off + o->cap[z_id];
		cx_here = ((cap_x[z_id + 1] - cap_x[z_id]) / dz) * off + cap_x[z_id];
		cy_here = ((cap_y[z_id + 1] - cap_y[z_id]) / dz) * off + cap_y[z_id];
	} else {
		rad_here = o->cap[0];
		cx_here = cap_x[0];
		cy_here = cap_y[0];
	}
	double d_axis = sqrt((ph->start_coords.x - cx_here) * (ph->start_coords.x - cx_here) + (ph->start_coords.y - cy_here) * (ph->start_coords.y - cy_here));
	int iesc = 0, ret;
	if (d_axis > rad_here) { /* photon starts in the glass (photon.c:645-907) */
		ret = 2;
		if (leak_calc && ph->start_coords.z == 0) {
			orc_vec3 face = {0, 0, 1};
			orc_capil_reflect(o, ph, face, leak_calc);
		} else if (leak_calc && ph->start_coords.z > 0) {
			double d_wall = 0;
			int r_cntr, q_cntr;
			int wall_trace = orc_capil_trace_wall(o, ph, &d_wall, &r_cntr, &q_cntr);
			if (wall_trace <= 0) {
				ret = -1;
			} else {
				for (size_t i = 0; i < nE; i++)
					ph->weight[i] = ph->weight[i] * exp(-1. * d_wall * o->amu[i]);
				double step = d_wall / sqrt(orc_scalar(ph->exit_direction, ph->exit_direction));
				ph->exit_coords.x = ph->exit_coords.x + step * ph->exit_direction.x;
				ph->exit_coords.y = ph->exit_coords.y + step * ph->exit_direction.y;
				ph->exit_coords.z = ph->exit_coords.z + step * ph->exit_direction.z;
				ret = 1;
				if (wall_trace == 3)
					leak_append(&ph->extleak, &ph->n_extleak, ph->exit_coords, ph->exit_direction, ph->exit_electric_vector, ph->i_refl, nE, ph->weight);
				if (wall_trace == 2)
					leak_append(&ph->intleak, &ph->n_intleak, ph->exit_coords, ph->exit_direction, ph->exit_electric_vector, ph->i_refl, nE, ph->weight);
				if (wall_trace == 1) {
					ph->d_travel = ph->d_travel + d_wall;
					capillary_axis(o, n_shells, q_cntr, r_cntr, cap_x, cap_y);
					for (int i = 0; i <= nm; i++)
						if (o->z[i] <= ph->exit_coords.z)
							ix = i;
					for (int i = 0; i <= nm; i++) {
						iesc = orc_capil_trace(o, &ix, ph, cap_x, cap_y, leak_calc);
						if (iesc != 1)
							break;
					}
					if (iesc == -1 || iesc == -3) {
						ret = -1;
					} else if (iesc == 1 || iesc == -2) {
						double tt = ((o->z[nm] - ph->exit_coords.z) / ph->exit_direction.z);
						ph->exit_coords.x = ph->exit_coords.x + ph->exit_direction.x * tt;
						ph->exit_coords.y = ph->exit_coords.y + ph->exit_direction.y * tt;
						ph->exit_coords.z = ph->exit_coords.z + ph->exit_direction.z * tt;
						int inside = orc_within_pc_boundary(o->ext[nm], ph->exit_coords);
						if (inside == 0)
							leak_append(&ph->extleak, &ph->n_extleak, ph->exit_coords, ph->exit_direction, ph->exit_electric_vector, ph->i_refl, nE, ph->weight);
						else if (inside == 1)
							leak_append(&ph->intleak, &ph->n_intleak, ph->exit_coords, ph->exit_direction, ph->exit_electric_vector, ph->i_refl, nE, ph->weight);
					}
				}
				if (ret == 1) { /* photon.c:856-885 */
					for (size_t i = 0; i < nE; i++)
						ph->weight[i] = 0.;
					ph->exit_coords.x = o->ext[nm] + 1.;
					ph->exit_coords.y = o->ext[nm] + 1.;
					ph->exit_coords.z = o->z[nm];
					ph->exit_direction = ph->start_direction;
					orc_norm(&ph->exit_direction);
				}
			}
		}
		free(cap_x);
		free(cap_y);
		return ret;
	}

	for (int i = 0; i <= nm; i++) {
		iesc = orc_capil_trace(o, &ix, ph, cap_x, cap_y, leak_calc);
		if (iesc != 1)
			break;
	}
	free(cap_x);
	free(cap_y);
	if ((iesc == -1) || (iesc == -3))
		return -1;
	return (iesc == 0) ? 0 : 1;
}

/* ------------------------------------------------------------------ source */

void orc_source_get_photon(const orc_source *s, const orc_optic *o, orc_uniform_fn uni, void *st, orc_photon *ph) /* source.c:23-144 */
{
	orc_vec3 start, dir, ev, src;
	double r = uni(st);
	double phi = atan(s->src_y / s->src_x * tan(2.0 * M_PI * r / 4.));
	r = uni(st);
	if ((r >= 0.25) && (r < 0.5))
		phi = M_PI - phi;
	if ((r >= 0.5) && (r < 0.75))
		phi = M_PI + phi;
	if (r >= 0.75)
		phi = -1.0 * phi;
	double max_rad = s->src_x * s->src_y / sqrt((s->src_y * cos(phi)) * (s->src_y * cos(phi)) + (s->src_x * sin(phi)) * (s->src_x * sin(phi)));
	r = uni(st);
	double sx = sqrt(r) * max_rad * cos(phi) + s->src_shiftx;
	double sy = sqrt(r) * max_rad * sin(phi) + s->src_shifty;
	src.x = sx;
	src.y = sy;
	src.z = 0;

	if (s->src_sigx < 0. || s->src_sigy < 0.) { /* uniform illumination of the entrance */
		double n_shells = orc_n_shells(o->n_cap);
		if (n_shells == 0.) {
			r = uni(st);
			start.x = (2. * r - 1.) * o->cap[0];
			r = uni(st);
			start.y = (2. * r - 1.) * o->cap[0];
		} else {
			int inside;
			do {
				r = uni(st);
				start.x = (2. * r - 1.) * o->ext[0];
				r = uni(st);
				start.y = (2. * r - 1.) * o->ext[0];
				start.z = 0.;
				inside = orc_within_pc_boundary(o->ext[0], start);
			} while (inside == 0);
		}
		dir.x = start.x - sx;
		dir.y = start.y - sy;
		dir.z = s->d_source;
	} else {
		r = uni(st);
		dir.x = s->src_sigx * (1. - 2. * fabs(r));
		r = uni(st);
		dir.y = s->src_sigy * (1. - 2. * fabs(r));
		dir.z = 1.;
		start.x = src.x + dir.x * s->d_source / dir.z;
		start.y = src.y + dir.y * s->d_source / dir.z;
	}
	start.z = 0.;
	orc_norm(&dir);

	double frac_hor = (1. + s->hor_pol) / 2.;
	r = uni(st);
	if (fabs(r) <= frac_hor) {
		ev.x = 1.; ev.y = 0.; ev.z = 0.;
	} else {
		ev.x = 0.; ev.y = 1.; ev.z = 0.;
	}
	double cosalpha = orc_scalar(ev, dir);
	double alpha = acos(cosalpha);
	double c_ae = 1. / sin(alpha);
	double c_be = -1. * c_ae * cosalpha;
	ev.x = ev.x * c_ae + dir.x * c_be;
	ev.y = ev.y * c_ae + dir.y * c_be;
	ev.z = ev.z * c_ae + dir.z * c_be;
	orc_norm(&ev);

	ph->start_coords = ph->exit_coords = start;
	ph->start_direction = ph->exit_direction = dir;
	ph->start_electric_vector = ph->exit_electric_vector = ev;
	ph->src_start_coords = src;
	ph->d_travel = 0;
	ph->i_refl = 0;
}

/* ------------------------------------------------------------------ random streams */

void orc_mt_seed(orc_mt19937 *r, unsigned long s)
{
	if (s == 0)
		s = 4357;
	r->mt[0] = (uint32_t)(s & 0xffffffffUL);
	for (int i = 1; i < 624; i++)
		r->mt[i] = 1812433253U * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
	r->mti = 624;
}

double orc_mt_uniform(void *state)
{
	orc_mt19937 *r = state;
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

void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
	uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
	uint32_t k0 = key_in[0], k1 = key_in[1];
	for (int round = 0; round < 10; round++) {
		uint64_t p0 = (uint64_t)0xD2511F53U * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57U * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9U;
		k1 += 0xBB67AE85U;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_philox_uniform(void *state)
{
	orc_philox_stream *s = state;
	uint32_t ctr[4] = {(uint32_t)s->id, (uint32_t)(s->id >> 32), s->draw >> 1, 0};
	uint32_t key[2] = {(uint32_t)s->seed, (uint32_t)(s->seed >> 32)};
	uint32_t o[4];
	orc_philox4x32_10(ctr, key, o);
	uint64_t bits = (s->draw & 1) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
	s->draw++;
	return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------ driver */

static void run_ids_impl(const orc_source *s, const orc_optic *o, uint64_t seed, int64_t id0, int64_t n_ids, int leak_calc, int threads,
			 double *sum_weights, orc_counters *cnt, int8_t *status, double *rec, const double *starts);

void orc_run_ids(const orc_source *s, const orc_optic *o, uint64_t seed, int64_t id0, int64_t n_ids, int leak_calc,
		 int threads, double *sum_weights, orc_counters *cnt, int8_t *status, double *rec)
{
	run_ids_impl(s, o, seed, id0, n_ids, leak_calc, threads, sum_weights, cnt, status, rec, NULL);
}

/* same driver over caller-supplied start states starts[n][11] = coords(3) dir(3) elecv(3) src_x src_y instead of sampling them */
void orc_run_starts(const orc_optic *o, int64_t n, const double *starts, int leak_calc, int threads, double *sum_weights, orc_counters *cnt,
		    int8_t *status, double *rec)
{
	run_ids_impl(NULL, o, 0, 0, n, leak_calc, threads, sum_weights, cnt, status, rec, starts);
}

static void run_ids_impl(const orc_source *s, const orc_optic *o, uint64_t seed, int64_t id0, int64_t n_ids, int leak_calc, int threads,
			 double *sum_weights, orc_counters *cnt, int8_t *status, double *rec, const double *starts) /* source.c:745-1075 */
{
	size_t nE = o->n_energies;
	int nm = o->nmax;
	int mono = (orc_n_shells(o->n_cap) == 0.);
	size_t stride = 17 + nE;
	memset(cnt, 0, sizeof *cnt);
	cnt->n_ids = n_ids;
	for (size_t k = 0; k < nE; k++)
		sum_weights[k] = 0.;
	if (threads < 1)
		threads = 1;

	int64_t n_exit = 0, n_ne = 0, n_nt = 0, n_unc = 0, refl_exit = 0, refl_all = 0, nseg = 0;
#pragma omp parallel num_threads(threads) reduction(+ : n_exit, n_ne, n_nt, n_unc, refl_exit, refl_all, nseg)
	{
		double *w = malloc(sizeof(double) * nE);
		double *acc = calloc(nE, sizeof(double));
		t_segment_calls = 0;
		t_reflect_calls = 0;
#pragma omp for schedule(dynamic, 64)
		for (int64_t j = 0; j < n_ids; j++) {
			orc_photon ph;
			memset(&ph, 0, sizeof ph);
			ph.weight = w;
			if (starts) {
				const double *q = starts + (size_t)j * 11;
				orc_vec3 c = {q[0], q[1], q[2]}, d = {q[3], q[4], q[5]}, e = {q[6], q[7], q[8]}, sp = {q[9], q[10], 0};
				ph.start_coords = ph.exit_coords = c;
				ph.start_direction = ph.exit_direction = d;
				ph.start_electric_vector = ph.exit_electric_vector = e;
				ph.src_start_coords = sp;
			} else {
				orc_philox_stream st = {seed, (uint64_t)(id0 + j), 0};
				orc_source_get_photon(s, o, orc_philox_uniform, &st, &ph);
			}
			int iesc = orc_photon_launch(o, &ph, leak_calc);
			int code = iesc;
			if (iesc == 1) { /* exit-window test, source.c:762-777 */
				orc_vec3 e = propagate_to_z(ph.exit_coords, ph.exit_direction, o->z[nm]);
				int inside;
				if (mono)
					inside = (sqrt((e.x) * (e.x) + (e.y) * (e.y)) > o->ext[nm]) ? 0 : 1;
				else
					inside = orc_within_pc_boundary(o->ext[nm], e);
				if (inside != 1)
					code = 3;
			}
			if (status)
				status[j] = (int8_t)code;
			if (code == 0) n_nt++;
			else if (code == 2) n_ne++;
			else if (code != 1) n_unc++;
			if (code == 1) {
				n_exit++;
				refl_exit += ph.i_refl;
				for (size_t k = 0; k < nE; k++)
					acc[k] += w[k];
				if (rec) {
					double *r = rec + (size_t)j * stride;
					double zt = o->z[nm];
					r[0] = ph.src_start_coords.x; r[1] = ph.src_start_coords.y;
					r[2] = ph.start_coords.x; r[3] = ph.start_coords.y;
					r[4] = ph.start_direction.x; r[5] = ph.start_direction.y;
					double ca = orc_scalar(ph.start_electric_vector, ph.start_direction);
					double al = acos(ca);
					double c_ae = 1. / sin(al), c_be = -1. * c_ae * ca;
					orc_vec3 t = {ph.start_electric_vector.x * c_ae + ph.start_direction.x * c_be,
						      ph.start_electric_vector.y * c_ae + ph.start_direction.y * c_be,
						      ph.start_electric_vector.z * c_ae + ph.start_direction.z * c_be};
					orc_norm(&t);
					r[6] = round(t.x); r[7] = round(t.y);
					r[8] = ph.exit_coords.x + ph.exit_direction.x * (zt - ph.exit_coords.z) / ph.exit_direction.z;
					r[9] = ph.exit_coords.y + ph.exit_direction.y * (zt - ph.exit_coords.z) / ph.exit_direction.z;
					r[10] = ph.exit_coords.z + ph.exit_direction.z * (zt - ph.exit_coords.z) / ph.exit_direction.z;
					r[11] = ph.exit_direction.x; r[12] = ph.exit_direction.y;
					orc_vec3 u = {ph.exit_electric_vector.x * c_ae + ph.exit_direction.x * c_be,
						      ph.exit_electric_vector.y * c_ae + ph.exit_direction.y * c_be,
						      ph.exit_electric_vector.z * c_ae + ph.exit_direction.z * c_be};
					orc_norm(&u);
					r[13] = round(u.x); r[14] = round(u.y);
					r[15] = (double)ph.i_refl;
					r[16] = ph.d_travel + sqrt((r[8] - ph.exit_coords.x) * (r[8] - ph.exit_coords.x) + (r[9] - ph.exit_coords.y) * (r[9] - ph.exit_coords.y) +
								   (zt - ph.exit_coords.z) * (zt - ph.exit_coords.z));
					memcpy(r + 17, w, sizeof(double) * nE);
				}
			}
			orc_photon_clear_leaks(&ph);
		}
		refl_all += t_reflect_calls;
		nseg += t_segment_calls;
#pragma omp critical
		for (size_t k = 0; k < nE; k++)
			sum_weights[k] += acc[k];
		free(w);
		free(acc);
	}
	cnt->n_exit = n_exit;
	cnt->n_not_entered = n_ne;
	cnt->n_not_transmitted = n_nt;
	cnt->n_uncounted = n_unc;
	cnt->sum_irefl_exit = refl_exit;
	cnt->sum_irefl_all = refl_all;
	cnt->n_segments = nseg;
}
