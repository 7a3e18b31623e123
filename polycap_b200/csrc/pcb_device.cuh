/*
 * pcb_device.cuh -- device functions of the photon-transport hot path (FP64, sm_100a).
 *
 * Each function states the reference lines (relative to /root/reference) whose result it
 * reproduces.  Arithmetic that decides a photon's fate (the quadratic ray/cone solve, the
 * surface normal, the Fresnel term, the specular direction) keeps the reference's operation
 * order and is compiled without FMA contraction (-fmad=false); only the conservative
 * "photon is clearly inside the bore" pre-test of the segment scan uses explicit FMAs, and
 * its verdict is never used to accept a hit -- only to skip segments that cannot contain one.
 */
#ifndef PCB_DEVICE_CUH
#define PCB_DEVICE_CUH

#include <cstdint>
#include <cuda_runtime.h>

#define PCB_COSPI_6 0.86602540378443864676
#define PCB_PI 3.14159265358979323846
#define PCB_PI_2 1.57079632679489661923

/* energy-constant rows (one double per energy each) */
enum { EC_NRE = 0, EC_NIM, EC_I2RE, EC_I2IM, EC_KR, EC_AMU, EC_ROWS };

/* status codes of a finished photon: the reference's launch return codes (photon.c:947-954)
 * plus 3 = "returned 1 but missed the exit window" (source.c:762-777). */
enum { ST_EXIT = 1, ST_ABSORBED = 0, ST_NOT_ENTERED = 2, ST_OUTSIDE = -2, ST_ERROR = -1, ST_MISSED_WINDOW = 3 };

struct PcbOptic {
	int nmax;          /* tables hold nmax+1 entries */
	int n_energies;
	int mono;          /* n_shells == 0 */
	int all_slow;      /* force the exact in-optic checks on every segment */
	int tstride;       /* doubles between consecutive tables (16-byte multiple, >= nmax+2) */
	double n_shells;
	double sig_rough;
	double hex_den;    /* 2*cos(pi/6)*(n_shells+1) */
	double m_safe;     /* capillaries with max|n_j.(A,B)| < m_safe never leave the hexagon */
	const double *z, *cap, *ext, *hexz, *rm; /* global copies of the tables */
	const double *ec;  /* [EC_ROWS][n_energies] */
};

struct PcbSource {
	double d_source, src_x, src_y, src_sigx, src_sigy, src_shiftx, src_shifty, hor_pol;
	double cap0, ext0;
	uint64_t seed;
};

struct V3 {
	double x, y, z;
};

__device__ __forceinline__ double dot3(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* photon.c:379-386 */

__device__ __forceinline__ void norm3(V3 &v) /* photon.c:365-376 */
{
	double len = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
	v.x /= len;
	v.y /= len;
	v.z /= len;
}

/* photon.c:139-169; returns 1 inside, 0 outside, -1 when R <= 0 */
__device__ __forceinline__ int within_pc(double R, double x, double y)
{
	if (R <= 0.)
		return -1;
	double apothem = sqrt((R * R) - ((R / 2.) * (R / 2.)));
	double d1 = fabs(0 * x + 1 * y);
	double d2 = fabs(PCB_COSPI_6 * x + 0.5 * y);
	double d3 = fabs(PCB_COSPI_6 * x + -0.5 * y);
	return (d1 > apothem || d2 > apothem || d3 > apothem) ? 0 : 1;
}

/* photon.c:540-552 (same code at capil.c:961-973, 1035-1047) */
__device__ __forceinline__ void hex_cell(double hexz, double x, double y, double &q, double &r)
{
	r = y * (2. / 3) / hexz;
	q = (x / (2. * PCB_COSPI_6) - y / 3) / hexz;
	double fq = fabs(q - round(q)), fr = fabs(r - round(r)), fs = fabs(-1. * q - r - round(-1. * q - r));
	if (fq > fr && fq > fs) {
		q = -1. * round(r) - round(-1. * q - r);
		r = round(r);
	} else if (fr > fs) {
		r = -1. * round(q) - round(-1. * q - r);
		q = round(q);
	} else {
		q = round(q);
		r = round(r);
	}
}

/* ------------------------------------------------------------------ ray / truncated-cone solve
 * capil.c:105-254 (argument checks of :65-100 are done by the callers that can violate them).
 * `last_z` is photon_coord->z on entry.  Returns the reference's code; on 1 fills hit and sn. */
static __device__ __noinline__ int segment_solve(const V3 &c0, const V3 &c1, double r0, double r1, const V3 &p0, V3 dir, double last_z, V3 &hit, V3 &sn)
{
	sn.x = sn.y = sn.z = 0.0;
	norm3(dir);
	V3 axis = {c1.x - c0.x, c1.y - c0.y, c1.z - c0.z};
	double axis_len = sqrt(dot3(axis, axis));

	double ux = (dir.x / dir.z) - (axis.x / axis.z);
	double uy = (dir.y / dir.z) - (axis.y / axis.z);
	double sl = (r1 - r0) / (c1.z - c0.z);
	double a = (ux * ux + uy * uy - sl * sl);
	double b = (2. * (p0.x - c0.x) * ux + 2. * (p0.y - c0.y) * uy - 2. * r0 * sl);
	double c = ((p0.x - c0.x) * (p0.x - c0.x) + (p0.y - c0.y) * (p0.y - c0.y) - r0 * r0);
	double discr = b * b - 4. * a * c;
	double zi;
	if (discr < 0)
		return -2;
	if (discr == 0) {
		zi = p0.z + (-1. * b) / (2. * a);
	} else {
		double sq = sqrt(discr);
		double d1 = (-1. * b + sq) / (2. * a);
		double d2 = (-1. * b - sq) / (2. * a);
		bool bad1 = (p0.z + d1 < c0.z || p0.z + d1 - last_z < 1.e-5 || p0.z + d1 > c1.z);
		bool bad2 = (p0.z + d2 < c0.z || p0.z + d2 - last_z < 1.e-5 || p0.z + d2 > c1.z);
		if (bad1) {
			if (bad2)
				return -3;
			zi = p0.z + d2;
		} else if (bad2) {
			zi = p0.z + d1;
		} else if (p0.z + d2 - last_z < p0.z + d1 - last_z) {
			zi = p0.z + d2;
		} else {
			zi = p0.z + d1;
		}
	}
	if (zi > c1.z)
		return -4;
	if (zi < c0.z || zi - last_z < 1.e-5)
		return -5;
	double d_proj = (zi - p0.z) / dir.z;
	if (d_proj < 1.e-10)
		return -6;
	hit.x = p0.x + d_proj * dir.x;
	hit.y = p0.y + d_proj * dir.y;
	hit.z = zi;

	V3 rel = {p0.x - c0.x, p0.y - c0.y, p0.z - c0.z};
	double da = dot3(dir, axis);
	double t = ((d_proj + (dot3(rel, axis) / da)) / (dot3(axis, axis) / da));
	V3 radial = {hit.x - (c0.x + t * axis.x), hit.y - (c0.y + t * axis.y), hit.z - (c0.z + t * axis.z)};
	double radial_len = sqrt(dot3(radial, radial));
	double tga = (r0 - r1) / axis_len;
	double gam = atan(tga);
	double sga = sin(gam);
	double cga = cos(gam);
	sn.x = cga * radial.x / radial_len + sga * axis.x / axis_len;
	sn.y = cga * radial.y / radial_len + sga * axis.y / axis_len;
	sn.z = cga * radial.z / radial_len + sga * axis.z / axis_len;
	norm3(sn);
	return 1;
}

/* ------------------------------------------------------------------ complex helpers
 * The reference evaluates the Fresnel term with C99 complex arithmetic (capil.c:493-511), i.e.
 * with libgcc's __divdc3 and glibc's csqrt/cabs.  These mirror those algorithms for finite,
 * normally scaled operands (third-party code, restated from the published algorithms:
 * libgcc "improved Smith" division; glibc s_csqrt_template.c; glibc e_hypot.c FMA kernel). */
__device__ __forceinline__ double hypot_g(double x, double y)
{
	double ax = fabs(x), ay = fabs(y);
	if (ax < ay) {
		double t = ax;
		ax = ay;
		ay = t;
	}
	if (!(ay > ax * 0x1p-54)) /* also catches ay == 0 and NaN ordering */
		return ax + ay;
	double t1 = ay + ay, t2 = ax - ay;
	if (t1 >= ax)
		return sqrt(__fma_rn(t1, ax, t2 * t2));
	return sqrt(__fma_rn(ax, ax, ay * ay));
}

/* 1.0 / (c + i d)  ==  __divdc3(1.0, 0.0, c, d) */
__device__ __forceinline__ void cinv_g(double c, double d, double &x, double &y)
{
	if (fabs(c) < fabs(d)) {
		double ratio = c / d;
		double denom = (c * ratio) + d;
		x = ((1.0 * ratio) + 0.0) / denom;
		y = ((0.0 * ratio) - 1.0) / denom;
	} else {
		double ratio = d / c;
		double denom = (d * ratio) + c;
		x = ((0.0 * ratio) + 1.0) / denom;
		y = (0.0 - (1.0 * ratio)) / denom;
	}
}

__device__ __forceinline__ void csqrt_g(double re, double im, double &r, double &s)
{
	if (im == 0.0) {
		if (re < 0) {
			r = 0.0;
			s = copysign(sqrt(-re), im);
		} else {
			r = fabs(sqrt(re));
			s = copysign(0.0, im);
		}
		return;
	}
	if (re == 0.0) {
		double h = sqrt(0.5 * fabs(im));
		r = h;
		s = copysign(h, im);
		return;
	}
	double d = hypot_g(re, im);
	if (re > 0) {
		r = sqrt(0.5 * (d + re));
		s = 0.5 * (im / r);
	} else {
		s = sqrt(0.5 * (d - re));
		r = fabs(0.5 * (im / s));
	}
	s = copysign(s, im);
}

/* Reflectivity for one energy, capil.c:497-511,538 with the per-energy constants
 * n = 1-alfa+i*beta and (1/n)^2 precomputed on the host in the reference's own order
 * (capil.c:493-495,499).  Keeps the reference's r_p, whose numerator/denominator real parts
 * absorb Im(n cos theta) because of the unparenthesised new_Dcomplex macro (capil.c:35,509). */
__device__ __forceinline__ double fresnel_rtot(double n_re, double n_im, double i2_re, double i2_im, double cos_t, double sin_t,
					       double frac_s, double frac_p)
{
	double s2 = sin_t * sin_t;
	double t_re = i2_re * s2, t_im = i2_im * s2;
	double q_re, q_im;
	csqrt_g(1.0 - t_re, -1.0 * t_im, q_re, q_im);

	double m_re = n_re * q_re - n_im * q_im; /* n * root */
	double m_im = n_re * q_im + n_im * q_re;
	double ix, iy;
	cinv_g(cos_t + m_re, m_im, ix, iy);
	double a_re = cos_t - m_re, a_im = -1.0 * m_im;
	double rs_re = a_re * ix - a_im * iy;
	double rs_im = a_re * iy + a_im * ix;
	double rs = hypot_g(rs_re, rs_im);
	rs *= rs;

	double c_re = n_re * cos_t, c_im = n_im * cos_t; /* n * cos(theta) */
	cinv_g(q_re + c_re + c_im, q_im, ix, iy);
	double b_re = q_re - c_re - c_im, b_im = q_im;
	double rp_re = b_re * ix - b_im * iy;
	double rp_im = b_re * iy + b_im * ix;
	double rp = hypot_g(rp_re, rp_im);
	rp *= rp;
	return rs * frac_s + rp * frac_p;
}

/* Energy-independent part of polycap_refl_polar (capil.c:482-490, 497-498, 516-553).
 * `sn` arrives normalised by capil_reflect (capil.c:596); ev is the photon's electric vector
 * (renormalised in place like the reference when its length is not exactly 1). */
struct ReflGeom {
	double cos_t, sin_t, frac_s, frac_p;
	V3 new_ev;
};

static __device__ __noinline__ void refl_geometry(V3 sn, const V3 &dir, V3 &ev, ReflGeom &g)
{
	if (sqrt(sn.x * sn.x + sn.y * sn.y + sn.z * sn.z) != 1)
		norm3(sn);
	double theta = acos(dot3(sn, dir));
	if (sqrt(ev.x * ev.x + ev.y * ev.y + ev.z * ev.z) != 1)
		norm3(ev);
	g.cos_t = cos(theta);
	g.sin_t = sin(theta);

	V3 s, p;
	s.x = sn.y * dir.z - dir.y * sn.z;
	s.y = sn.z * dir.x - dir.z * sn.x;
	s.z = sn.x * dir.y - dir.x * sn.y;
	norm3(s);
	p.x = dir.y * s.z - s.y * dir.z;
	p.y = dir.z * s.x - s.z * dir.x;
	p.z = dir.x * s.y - s.x * dir.y;
	norm3(p);
	double ca = dot3(ev, s);
	g.frac_s = ca * ca;
	g.frac_p = 1. - g.frac_s;
	double cb = dot3(ev, sn);
	double cc = dot3(ev, p);
	double fs = g.frac_s, fp = g.frac_p;
	g.new_ev.x = sqrt((ev.x * ca * fs) * (ev.x * ca * fs) + (ev.x * cb * fp) * (ev.x * cb * fp) + (ev.x * cc * fp) * (ev.x * cc * fp));
	g.new_ev.y = sqrt((ev.y * ca * fs) * (ev.y * ca * fs) + (ev.y * cb * fp) * (ev.y * cb * fp) + (ev.y * cc * fp) * (ev.y * cc * fp));
	g.new_ev.z = sqrt((ev.z * ca * fs) * (ev.z * ca * fs) + (ev.z * cb * fp) * (ev.z * cb * fp) + (ev.z * cc * fp) * (ev.z * cc * fp));
	norm3(g.new_ev);
}

/* ------------------------------------------------------------------ Philox4x32-10
 * counter = (id_lo, id_hi, block, 0), key = (seed_lo, seed_hi); draw d of photon `id` is the
 * 53 high bits of 64-bit half (d & 1) of block d >> 1.  Replaces polycap_rng (rng.c:93-95). */
struct Philox {
	uint32_t k0, k1, id_lo, id_hi;
	uint32_t draw;
	uint32_t cache[4];
	__device__ __forceinline__ void init(uint64_t seed, uint64_t id)
	{
		k0 = (uint32_t)seed;
		k1 = (uint32_t)(seed >> 32);
		id_lo = (uint32_t)id;
		id_hi = (uint32_t)(id >> 32);
		draw = 0;
	}
	__device__ __forceinline__ void block(uint32_t blk)
	{
		uint32_t c0 = id_lo, c1 = id_hi, c2 = blk, c3 = 0, a = k0, b = k1;
#pragma unroll
		for (int r = 0; r < 10; r++) {
			uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
			uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
			uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
			c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
			a += 0x9E3779B9u;
			b += 0xBB67AE85u;
		}
		cache[0] = c0; cache[1] = c1; cache[2] = c2; cache[3] = c3;
	}
	__device__ __forceinline__ double uniform()
	{
		if ((draw & 1u) == 0)
			block(draw >> 1);
		uint64_t bits = (draw & 1u) ? (((uint64_t)cache[3] << 32) | cache[2]) : (((uint64_t)cache[1] << 32) | cache[0]);
		draw++;
		return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
	}
};

/* ------------------------------------------------------------------ source sampling
 * source.c:52-137: point on the source ellipse, then either uniform illumination of the
 * entrance (sig < 0) or a direction inside +-sig; electric vector made perpendicular. */
struct StartState {
	V3 coords, dir, ev;
	double src_x, src_y;
};

static __device__ __noinline__ void source_sample(const PcbSource &s, int mono, uint64_t id, StartState &o)
{
	Philox rng;
	rng.init(s.seed, id);
	double r = rng.uniform();
	double phi = atan(s.src_y / s.src_x * tan(2.0 * PCB_PI * r / 4.));
	r = rng.uniform();
	if ((r >= 0.25) && (r < 0.5))
		phi = PCB_PI - phi;
	if ((r >= 0.5) && (r < 0.75))
		phi = PCB_PI + phi;
	if (r >= 0.75)
		phi = -1.0 * phi;
	double cphi = cos(phi), sphi = sin(phi);
	double max_rad = s.src_x * s.src_y / sqrt((s.src_y * cphi) * (s.src_y * cphi) + (s.src_x * sphi) * (s.src_x * sphi));
	r = rng.uniform();
	double sx = sqrt(r) * max_rad * cphi + s.src_shiftx;
	double sy = sqrt(r) * max_rad * sphi + s.src_shifty;
	o.src_x = sx;
	o.src_y = sy;
	if (s.src_sigx < 0. || s.src_sigy < 0.) {
		if (mono) {
			r = rng.uniform();
			o.coords.x = (2. * r - 1.) * s.cap0;
			r = rng.uniform();
			o.coords.y = (2. * r - 1.) * s.cap0;
		} else {
			int inside;
			do {
				r = rng.uniform();
				o.coords.x = (2. * r - 1.) * s.ext0;
				r = rng.uniform();
				o.coords.y = (2. * r - 1.) * s.ext0;
				inside = within_pc(s.ext0, o.coords.x, o.coords.y);
			} while (inside == 0);
		}
		o.dir.x = o.coords.x - sx;
		o.dir.y = o.coords.y - sy;
		o.dir.z = s.d_source;
	} else {
		r = rng.uniform();
		o.dir.x = s.src_sigx * (1. - 2. * fabs(r));
		r = rng.uniform();
		o.dir.y = s.src_sigy * (1. - 2. * fabs(r));
		o.dir.z = 1.;
		o.coords.x = sx + o.dir.x * s.d_source / o.dir.z;
		o.coords.y = sy + o.dir.y * s.d_source / o.dir.z;
	}
	o.coords.z = 0.;
	norm3(o.dir);
	double frac_hor = (1. + s.hor_pol) / 2.;
	r = rng.uniform();
	if (fabs(r) <= frac_hor) {
		o.ev.x = 1.; o.ev.y = 0.; o.ev.z = 0.;
	} else {
		o.ev.x = 0.; o.ev.y = 1.; o.ev.z = 0.;
	}
	double cosalpha = dot3(o.ev, o.dir);
	double alpha = acos(cosalpha);
	double c_ae = 1. / sin(alpha);
	double c_be = -1. * c_ae * cosalpha;
	o.ev.x = o.ev.x * c_ae + o.dir.x * c_be;
	o.ev.y = o.ev.y * c_ae + o.dir.y * c_be;
	o.ev.z = o.ev.z * c_ae + o.dir.z * c_be;
	norm3(o.ev);
}

#endif
