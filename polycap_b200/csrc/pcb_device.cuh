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


#include "pcb_codes.h"

struct PcbOptic {
	int nmax;          /* tables hold nmax+1 entries */
	int n_energies;
	int mono;          /* n_shells == 0 */
	int all_slow;      /* force the exact in-optic checks on every segment */
	int tstride;       /* doubles between consecutive tables (16-byte multiple, >= nmax+2) */
	double n_shells;
	double sig_rough;
	double hex_den;    /* 2*cos(pi/6)*(n_shells+1) */
	double m_safe;     /* capillaries with max|n_j.(A,B)| < m_safe never leave the hexagon (restart margin for tan(theta) <= 0.01) */
	double m_deep;     /* ... < m_deep: the same with the margin for tan(theta) <= 0.25 */
	const double *z, *cap, *ext, *hexz, *rm; /* global copies of the tables */
	const double *ec;  /* [EC_ROWS][n_energies] */
};

struct PcbSource {
	double d_source, src_x, src_y, src_sigx, src_sigy, src_shiftx, src_shifty, hor_pol;
	double cap0, ext0;
	uint64_t seed;
};

#include "pcb_math.h"

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
