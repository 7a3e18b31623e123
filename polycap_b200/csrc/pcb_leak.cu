/*
 * pcb_leak.cu -- photon transport WITH leak / recap wall transport (leak_calc = true), sm_100a, FP64.
 *
 * Reference behaviour reproduced (paths relative to /root/reference):
 *   polycap_capil_trace_wall      src/polycap-capil.c:896-1194   march through the glass to the next bore / out of the optic
 *   polycap_photon_pc_intersect   src/polycap-photon.c:174-362   where a photon left through the outer hexagonal surface
 *   polycap_capil_reflect (leak)  src/polycap-capil.c:565-891    leak weights, ext/int leak records, the secondary ("recap")
 *                                                                photon that continues in the neighbouring capillary
 *   polycap_capil_trace           src/polycap-capil.c:1197-1361  with the per-segment in-optic checks
 *   polycap_photon_launch (leak)  src/polycap-photon.c:645-907   photons that start in the glass
 *
 * One thread owns one primary photon (persistent threads pull ids from a global counter).  The reference recurses --
 * capil_reflect creates a secondary photon and runs capil_trace on it, whose capil_reflect may create another -- which
 * here is an explicit frame stack per thread: a frame is a photon with its weight vector, capillary axis, segment
 * index and trace-loop counter; a parent waits in its frame while its secondary runs.  Leak records are streamed in
 * depth-first order into the thread's record strip, which preserves the reference's order within the external and
 * the internal list (the two lists are independent), and are rolled back when the reference discards a secondary's
 * leaks (capil.c:799-805).  All trajectory arithmetic uses the bit-faithful primitives of pcb_math.h; weights use the
 * reference-order Fresnel term.  This path favours fidelity over speed (the reference's leak mode is ~1000x slower
 * than its plain mode); the voted, cheap-weights tracer of pcb_trace.cuh serves leak_calc = false.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pcb_device.cuh"
#include "pcb_gpu.h"
#include "pcb_leak.h"

/* Capacities per primary photon.  A hard photon at a steep angle crosses capillary after capillary, each crossing one
 * nested secondary (the reference recurses just as deep: its corpus shows 440 leak records and i_refl 568 on one photon). */
#define LK_DEPTH 1400    /* frames: primary + nested secondaries */
#define LK_REC_CAP 2048  /* leak records */
#define LK_THREADS 128

enum { LK_ERR_DEPTH = 1, LK_ERR_RECORDS = 2, LK_ERR_POOL = 4 }; /* RECORDS: one photon exceeds LK_REC_CAP leaks; POOL: the wave's record pool is full */

struct LkCtx {
	int nmax, nE, mono, monotonic;
	int no_fast_march; /* POLYCAP_B200_LEAK_FULL_MARCH=1: every step of the wall march takes the reference's full step (checking aid) */
	double n_shells, sig_rough;
	const double *z, *cap, *ext, *hexz; /* global */
	const double *ec;                   /* [EC_ROWS][nE] */
};

struct LkFrame {
	V3 ec, ed, ev;       /* exit_coords, exit_direction, exit_electric_vector */
	long long irefl;
	double dtrav;
	double A, B;         /* capillary axis: cap_x[i] = A*hexz[i], cap_y[i] = B*hexz[i] */
	int ix, it, it_end;  /* segment index, trace-loop counter (photon.c:912, capil.c:785) */
	int last_esc;
	/* saved while a secondary runs */
	V3 sn;
	double cosalfa;
	int iesc_after, rec_mark, cont;
	double *w;
};

enum { CONT_TRACE = 0, CONT_FACE = 1, CONT_UNIT = 2 };

struct LkOut {
	double *rec;    /* this thread's strip: cap x (11+nE) */
	int nrec;
	int err;
	int cap;        /* records the strip holds */
	int depth;      /* frames this thread's stack holds */
	/* pool of full-size workspaces (LK_DEPTH frames, LK_REC_CAP records) a photon moves to when it outgrows the thread's small one */
	LkFrame *big_frames;
	double *big_arena, *big_rec;
	unsigned long long *big_next;
	unsigned long long big_n;
};

/* The photon's nesting is about to exceed the thread's small workspace: move its frames [0, sp], their weight vectors and its
 * records to a full-size workspace from the pool and carry on there.  false: already full-size, or the pool is used up. */
__device__ bool lk_grow(const LkCtx &c, LkOut &o, LkFrame *&fr, double *&arena, int sp)
{
	if (o.depth >= LK_DEPTH || !o.big_next)
		return false;
	const unsigned long long idx = atomicAdd(o.big_next, 1ULL);
	if (idx >= o.big_n)
		return false;
	const size_t nE = (size_t)c.nE, stride = 11 + nE;
	LkFrame *nf = o.big_frames + idx * (unsigned long long)LK_DEPTH;
	double *na = o.big_arena + idx * (unsigned long long)(LK_DEPTH + 1) * nE;
	double *nr = o.big_rec + idx * (unsigned long long)LK_REC_CAP * stride;
	for (size_t k = 0; k < (size_t)(sp + 1) * nE; k++)
		na[k] = arena[k];
	for (int k = 0; k <= sp; k++) {
		nf[k] = fr[k];
		nf[k].w = na + (fr[k].w - arena);
	}
	for (size_t k = 0; k < (size_t)o.nrec * stride; k++)
		nr[k] = o.rec[k];
	fr = nf;
	arena = na;
	o.rec = nr;
	o.cap = LK_REC_CAP;
	o.depth = LK_DEPTH;
	return true;
}

__device__ __forceinline__ int lk_z_index(const LkCtx &c, double zq) /* last i < nmax with z[i] <= zq, 0 when none */
{
	if (c.monotonic) {
		if (!(c.z[0] <= zq))
			return 0;
		int lo = 0, hi = c.nmax - 1; /* invariant: z[lo] <= zq */
		while (lo < hi) {
			int mid = (lo + hi + 1) >> 1;
			if (c.z[mid] <= zq)
				lo = mid;
			else
				hi = mid - 1;
		}
		return lo;
	}
	int id = 0;
	for (int i = 0; i < c.nmax; i++)
		if (c.z[i] <= zq)
			id = i;
	return id;
}

__device__ __forceinline__ int lk_z_index_incl(const LkCtx &c, double zq) /* last i <= nmax with z[i] <= zq (photon.c:626, capil.c:747) */
{
	int ix = 0;
	if (c.monotonic) {
		ix = lk_z_index(c, zq);
		if (c.z[c.nmax] <= zq)
			ix = c.nmax;
		return ix;
	}
	for (int i = 0; i <= c.nmax; i++)
		if (c.z[i] <= zq)
			ix = i;
	return ix;
}

__device__ __forceinline__ V3 lk_propagate(const V3 &p, const V3 &d, double zt)
{
	V3 r;
	r.x = p.x + d.x * (zt - p.z) / d.z;
	r.y = p.y + d.y * (zt - p.z) / d.z;
	r.z = zt;
	return r;
}

__device__ void lk_append(const LkCtx &c, LkOut &o, int kind, const V3 &co, const V3 &di, const V3 &el, long long nrefl, const double *w)
{
	if (o.nrec >= o.cap) {
		o.err |= LK_ERR_RECORDS;
		return;
	}
	double *r = o.rec + (size_t)o.nrec * (11 + c.nE);
	r[0] = (double)kind;
	r[1] = co.x; r[2] = co.y; r[3] = co.z;
	r[4] = di.x; r[5] = di.y; r[6] = di.z;
	r[7] = el.x; r[8] = el.y; r[9] = el.z;
	r[10] = (double)nrefl;
	for (int e = 0; e < c.nE; e++)
		r[11 + e] = w[e];
	o.nrec++;
}

/* photon.c:174-362 */
__device__ int lk_pc_intersect(const LkCtx &c, V3 coord, V3 direction, V3 &out)
{
	if (direction.z == 0.)
		return 0;
	V3 bd = {-1. * direction.x, -1. * direction.y, -1. * direction.z};
	norm3(bd);
	int z_id = lk_z_index(c, coord.z);
	double ext_here = (c.ext[z_id + 1] - c.ext[z_id]) / (c.z[z_id + 1] - c.z[z_id]) * (coord.z - c.z[z_id]) + c.ext[z_id];
	if (within_pc(ext_here, coord.x, coord.y) == 1)
		return 0;
	int step;
	if (bd.z < 0.) {
		z_id = z_id + 1;
		step = -1;
	} else {
		step = 1;
	}
	bool found = false;
	const int in0 = within_pc(ext_here, coord.x, coord.y);
	do {
		z_id += step;
		if (z_id < 0 || z_id > c.nmax)
			break;
		double px = coord.x + bd.x * (c.z[z_id] - coord.z) / bd.z;
		double py = coord.y + bd.y * (c.z[z_id] - coord.z) / bd.z;
		if (in0 != within_pc(c.ext[z_id], px, py)) {
			found = true;
			break;
		}
	} while (z_id >= 0 && z_id <= c.nmax);
	if (!found)
		return 0;
	int zb = z_id, ze = z_id - step;
	if (ze < 0 || ze > c.nmax)
		return 0;
	V3 beg, end;
	beg.x = coord.x + bd.x * (c.z[zb] - coord.z) / bd.z;
	beg.y = coord.y + bd.y * (c.z[zb] - coord.z) / bd.z;
	beg.z = coord.z + bd.z * (c.z[zb] - coord.z) / bd.z;
	end.x = coord.x + bd.x * (c.z[ze] - coord.z) / bd.z;
	end.y = coord.y + bd.y * (c.z[ze] - coord.z) / bd.z;
	end.z = coord.z + bd.z * (c.z[ze] - coord.z) / bd.z;
	double hb = sqrt((c.ext[zb] * c.ext[zb]) - ((c.ext[zb] / 2.) * (c.ext[zb] / 2.)));
	double he = sqrt((c.ext[ze] * c.ext[ze]) - ((c.ext[ze] / 2.) * (c.ext[ze] / 2.)));
	double pb[3], pe[3], zs[3];
	pb[0] = fabs(0 * beg.x + 1 * beg.y);
	pb[1] = fabs(PCB_COSPI_6 * beg.x + 0.5 * beg.y);
	pb[2] = fabs(PCB_COSPI_6 * beg.x + -0.5 * beg.y);
	pe[0] = fabs(0 * end.x + 1 * end.y);
	pe[1] = fabs(PCB_COSPI_6 * end.x + 0.5 * end.y);
	pe[2] = fabs(PCB_COSPI_6 * end.x + -0.5 * end.y);
	for (int k = 0; k < 3; k++)
		zs[k] = (pb[k] - hb) / (hb - he - pb[k] + pe[k]) * (c.ext[zb] - c.ext[ze]) + c.ext[zb];
	double lo = (step < 0) ? c.z[zb] : c.z[ze];
	double hi = (step < 0) ? c.z[ze] : c.z[zb];
	bool have = false;
	double zf = 0;
	for (int k = 0; k < 3; k++) {
		if (!(zs[k] >= lo && zs[k] <= hi))
			continue;
		if (!have) {
			zf = zs[k];
			have = true;
		} else if (step < 0 ? (zs[k] > zf) : (zs[k] < zf)) {
			zf = zs[k];
		}
	}
	if (!have) {
		out = end;
		return 1;
	}
	out.x = coord.x + bd.x * (zf - coord.z) / bd.z;
	out.y = coord.y + bd.y * (zf - coord.z) / bd.z;
	out.z = coord.z + bd.z * (zf - coord.z) / bd.z;
	return 1;
}

/* classification once a wall march ends without a new bore: 3 = out the side, 2 = exit face in the glass (capil.c:1076-1107, 1150-1186) */
__device__ int lk_wall_exit_class(const LkCtx &c, const V3 &ec, const V3 &ed, const V3 &fallback, double &d_travel)
{
	const int nm = c.nmax;
	V3 at_exit = lk_propagate(ec, ed, c.z[nm]);
	V3 rel;
	if (within_pc(c.ext[nm], at_exit.x, at_exit.y) == 0) {
		V3 side;
		if (lk_pc_intersect(c, at_exit, ed, side)) {
			rel.x = side.x - ec.x; rel.y = side.y - ec.y; rel.z = side.z - ec.z;
		} else {
			rel.x = fallback.x - ec.x; rel.y = fallback.y - ec.y; rel.z = fallback.z - ec.z;
		}
		d_travel = sqrt(dot3(rel, rel));
		return 3;
	}
	rel.x = fallback.x - ec.x; rel.y = fallback.y - ec.y; rel.z = fallback.z - ec.z;
	d_travel = sqrt(dot3(rel, rel));
	return 2;
}

__device__ __forceinline__ int lk_segment_checked(const V3 &c0, const V3 &c1, double r0, double r1, const V3 &p0, const V3 &dir, V3 &hit, V3 &sn)
{
	if (c0.z < 0. || c1.z < 0 || r0 < 0. || r1 < 0. || dir.z < 0 || c1.z <= c0.z) { /* capil.c:65-100 */
		sn.x = sn.y = sn.z = 0.0;
		return -1;
	}
	V3 h = hit;
	int r = segment_solve(c0, c1, r0, r1, p0, dir, hit.z, h, sn);
	if (r == 1)
		hit = h;
	return r;
}

/* capil.c:896-1194; ed is normalised in place like the reference does with photon->exit_direction */
__device__ int lk_trace_wall(const LkCtx &c, const V3 &ec, V3 &ed, double &d_travel, int &r_cntr, int &q_cntr, unsigned long long &n_steps)
{
	const int nm = c.nmax;
	d_travel = 0.;
	r_cntr = 0;
	q_cntr = 0;
	norm3(ed);
	if (ec.z >= c.z[nm])
		return -2;
	int z_id = lk_z_index(c, ec.z);
	double ext_here;
	if (c.z[z_id] != ec.z)
		ext_here = ((c.ext[z_id + 1] - c.ext[z_id]) / (c.z[z_id + 1] - c.z[z_id])) * (ec.z - c.z[z_id]) + c.ext[z_id];
	else
		ext_here = c.ext[z_id];
	V3 hit = ec, sn;
	const double n_shells = c.n_shells;
	if (n_shells == 0.) {
		if (sqrt((ec.x) * (ec.x) + (ec.y) * (ec.y)) > ext_here)
			return -2;
	} else if (within_pc(ext_here, ec.x, ec.y) == 0) {
		return -2;
	}
	double hexz = ext_here / (2. * PCB_COSPI_6 * (n_shells + 1));
	double q_i, r_i, q_new = 0, r_new = 0;
	hex_cell(hexz, ec.x, ec.y, q_i, r_i);
	int iesc = 0;
	V3 c0, c1, p0;
	if (n_shells == 0.) { /* capil.c:992-1013 */
		do {
			p0 = lk_propagate(ec, ed, c.z[z_id]);
			c0.x = 0.; c0.y = 0.; c0.z = c.z[z_id];
			c1.x = 0.; c1.y = 0.; c1.z = c.z[z_id + 1];
			iesc = lk_segment_checked(c0, c1, c.cap[z_id], c.cap[z_id + 1], p0, ed, hit, sn);
			z_id++;
		} while (iesc != 1 && z_id < nm - 1);
	} else {
		double dist = 0;
		for (;;) {
			/* March in steps of a tenth of the bore radius until the hex cell changes (capil.c:1018-1068).  A step is eventful when the point
			 * falls back into the photon's own bore, leaves its hexagonal cell or passes the end of the optic; between the bore wall and
			 * the cell boundary thousands of steps are none of that.  Such a step is recognised from the squared distance to the cell's
			 * axis and the three edge normals of the hexagon (bounds taken 1e-9 inside the safe zone, rounding differences to the
			 * reference's own expressions are ~1e-16) and then only advances the path length and the z index, exactly as the reference
			 * does; everything else takes the reference's full step, so the outcome is the reference's step for step. */
			const bool fast_ok = c.monotonic && ed.z >= 0. && !c.no_fast_march;
			const double inv_den = 1.0 / (2. * PCB_COSPI_6 * (n_shells + 1));
			const double Ai = (2. * q_i + r_i) * PCB_COSPI_6, Bi = r_i * (3. / 2);
			int zc = -1;
			double seg_z0 = 0, seg_e0 = 0, seg_c0 = 0, seg_se = 0, seg_sc = 0;
			for (;;) {
				n_steps++;
				dist += c.cap[z_id] / 10.;
				if (fast_ok) {
					const double p0z = ec.z + dist * ed.z;
					int zi = z_id;
					if (!(c.z[zi] <= p0z))
						zi = lk_z_index(c, p0z);
					while (zi + 1 < nm && c.z[zi + 1] <= p0z)
						zi++;
					if (zi != zc) {
						zc = zi;
						seg_z0 = c.z[zi];
						const double dz = c.z[zi + 1] - seg_z0;
						seg_e0 = c.ext[zi];
						seg_c0 = c.cap[zi];
						seg_se = (c.ext[zi + 1] - seg_e0) / dz;
						seg_sc = (c.cap[zi + 1] - seg_c0) / dz;
					}
					const double t = p0z - seg_z0;
					const double hz = (seg_se * t + seg_e0) * inv_den;
					const double rr = (seg_sc * t + seg_c0) * (1. + 1.e-9);
					const double ux = (ec.x + dist * ed.x) - Ai * hz, uy = (ec.y + dist * ed.y) - Bi * hz;
					const double h = PCB_COSPI_6 * hz * (1. - 1.e-9);
					if (rr > 0. && hz > 0. && ux * ux + uy * uy > rr * rr && fabs(ux) < h && fabs(0.5 * ux + PCB_COSPI_6 * uy) < h &&
					    fabs(0.5 * ux - PCB_COSPI_6 * uy) < h && p0z <= c.z[nm]) {
						z_id = zi;
						continue;
					}
				}
				p0.x = ec.x + dist * ed.x;
				p0.y = ec.y + dist * ed.y;
				p0.z = ec.z + dist * ed.z;
				if (c.monotonic && ed.z >= 0.) { /* the march only moves forward: the index follows incrementally */
					if (!(c.z[z_id] <= p0.z))
						z_id = lk_z_index(c, p0.z);
					while (z_id + 1 < nm && c.z[z_id + 1] <= p0.z)
						z_id++;
				} else {
					z_id = lk_z_index(c, p0.z);
				}
				ext_here = ((c.ext[z_id + 1] - c.ext[z_id]) / (c.z[z_id + 1] - c.z[z_id])) * (p0.z - c.z[z_id]) + c.ext[z_id];
				double rad_here = ((c.cap[z_id + 1] - c.cap[z_id]) / (c.z[z_id + 1] - c.z[z_id])) * (p0.z - c.z[z_id]) + c.cap[z_id];
				hexz = ext_here / (2. * PCB_COSPI_6 * (n_shells + 1));
				hex_cell(hexz, p0.x, p0.y, q_new, r_new);
				c0.y = r_i * (3. / 2) * hexz;
				c0.x = (2. * q_i + r_i) * PCB_COSPI_6 * hexz;
				double d0 = sqrt((p0.x - c0.x) * (p0.x - c0.x) + (p0.y - c0.y) * (p0.y - c0.y));
				if (d0 < rad_here && fabs(q_i) <= n_shells && fabs(r_i) <= n_shells && fabs(-1. * q_i - r_i) <= n_shells) {
					V3 rel = {p0.x - ec.x, p0.y - ec.y, p0.z - ec.z};
					d_travel = sqrt(dot3(rel, rel));
					if (d_travel > 1.e-5) {
						r_cntr = (int)r_i;
						q_cntr = (int)q_i;
						return 1;
					}
					d_travel = 0.;
				}
				if (!(q_new == q_i && r_new == r_i && p0.z <= c.z[nm]))
					break;
			}

			if (fabs(q_new) > n_shells || fabs(r_new) > n_shells || fabs(-1. * q_new - r_new) > n_shells || p0.z > c.z[nm]) {
				r_cntr = (int)r_new;
				q_cntr = (int)q_new;
				return lk_wall_exit_class(c, ec, ed, p0, d_travel);
			}
			iesc = 0;
			do { /* scan the segments of the new capillary from z_id on (capil.c:1110-1133) */
				p0 = lk_propagate(ec, ed, c.z[z_id]);
				hexz = c.ext[z_id] / (2. * PCB_COSPI_6 * (n_shells + 1));
				c0.y = r_new * (3. / 2) * hexz;
				c0.x = (2. * q_new + r_new) * PCB_COSPI_6 * hexz;
				c0.z = c.z[z_id];
				hexz = c.ext[z_id + 1] / (2. * PCB_COSPI_6 * (n_shells + 1));
				c1.y = r_new * (3. / 2) * hexz;
				c1.x = (2. * q_new + r_new) * PCB_COSPI_6 * hexz;
				c1.z = c.z[z_id + 1];
				iesc = lk_segment_checked(c0, c1, c.cap[z_id], c.cap[z_id + 1], p0, ed, hit, sn);
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
	r_cntr = (int)r_new;
	q_cntr = (int)q_new;
	if (iesc != 1) {
		V3 at_exit = lk_propagate(ec, ed, c.z[nm]);
		return lk_wall_exit_class(c, ec, ed, at_exit, d_travel);
	}
	V3 rel = {hit.x - ec.x, hit.y - ec.y, hit.z - ec.z};
	d_travel = sqrt(dot3(rel, rel));
	return (z_id >= nm) ? 2 : 1;
}

/* one capil_trace call up to (not including) capil_reflect: capil.c:1236-1343.  Returns 1 when a wall hit passed the
 * checks (f.sn / f.cosalfa set, f.ec moved to the hit), otherwise the trace's return code (-1, -2, -3). */
__device__ int lk_trace_to_hit(const LkCtx &c, LkFrame &f)
{
	const int nm = c.nmax;
	int iesc = 0;
	norm3(f.ed);
	V3 last_hit = f.ec;
	const V3 dir = f.ed;
	V3 sn = {0, 0, 0};
	double cosalfa = 0;
	for (int i = f.ix; i < nm; i++) {
		V3 c0 = {f.A * c.hexz[i], f.B * c.hexz[i], c.z[i]}, c1 = {f.A * c.hexz[i + 1], f.B * c.hexz[i + 1], c.z[i + 1]};
		V3 p0 = lk_propagate(f.ec, f.ed, c.z[i]);
		if (within_pc(c.ext[i], c0.x, c0.y) == 0 || within_pc(c.ext[i + 1], c1.x, c1.y) == 0)
			return -1;
		iesc = lk_segment_checked(c0, c1, c.cap[i], c.cap[i + 1], p0, dir, last_hit, sn);
		cosalfa = dot3(sn, dir);
		{
			double ac = pcb_acos(cosalfa);
			if (ac > PCB_PI / 2. || ac < 0.)
				iesc = -5;
		}
		if (iesc == 1) {
			double ext_here = ((c.ext[i] - c.ext[i + 1]) / (c.z[i] - c.z[i + 1])) * (last_hit.z - c.z[i + 1]) + c.ext[i + 1];
			if (c.n_shells == 0.) {
				if (sqrt(last_hit.x * last_hit.x + last_hit.y * last_hit.y) >= ext_here)
					return -3;
			} else if (within_pc(ext_here, last_hit.x, last_hit.y) == 0) {
				return -3;
			}
			f.ix = i + 1;
			break;
		} else if (within_pc(c.ext[i], p0.x, p0.y) == 0) {
			return -3;
		}
	}
	if (iesc != 1)
		return -2;
	V3 rel = {last_hit.x - f.ec.x, last_hit.y - f.ec.y, last_hit.z - f.ec.z};
	f.dtrav += sqrt(dot3(rel, rel));
	f.ec = last_hit;
	if (fabs(cosalfa) > 1.0)
		return -1;
	{
		int found = lk_z_index(c, f.ec.z);
		if (c.z[found] <= f.ec.z) /* no z[i] <= z (NaN photon): the reference leaves *ix = i+1 */
			f.ix = found;
	}
	double ext_here = ((c.ext[f.ix + 1] - c.ext[f.ix]) / (c.z[f.ix + 1] - c.z[f.ix])) * (last_hit.z - c.z[f.ix]) + c.ext[f.ix];
	if (c.n_shells == 0 && f.ec.x * f.ec.x + f.ec.y * f.ec.y >= ext_here)
		return -3;
	if (c.n_shells > 0 && within_pc(ext_here, f.ec.x, f.ec.y) == 0)
		return -3;
	f.sn = sn;
	f.cosalfa = cosalfa;
	return 1;
}

struct LkResult {
	int ret;
	V3 ec, ed, ev;
	long long irefl;
	double dtrav;
};

/* polycap_photon_launch with leak_calc = true (photon.c:390-955), iterative over an explicit frame stack.
 * `arena` holds (out.depth+1) weight vectors of nE doubles; arena[0..nE) ends up with the primary's weights. */
/* Single-step use of the state machine (the reference's TEST_BUILD entry points with leak_calc = true): the photon state comes from
 * the caller instead of the launch preamble, and the machine stops when the primary frame has finished
 *   mode 1  one polycap_capil_reflect (capil.c:565-891) on wall normal `normal`: ret = 1 / 0 / -1 / -2
 *   mode 2  one polycap_capil_trace (capil.c:1197-1361):                        ret = 1 / 0 / -1 / -2 / -3 */
struct LkUnit {
	int mode;
	V3 normal;
	long long irefl;
	double A, B;
	int ix;
};

__device__ void lk_photon(const LkCtx &c, const V3 &sc, V3 sd, const V3 &sev, double d_travel0, LkFrame *&fr, double *&arena, LkOut &out, LkResult &res,
			  unsigned long long &n_refl_events, unsigned long long &n_wall_steps, const LkUnit *unit = nullptr)
{
	const int nm = c.nmax, nE = c.nE;
	int sp = 0;
	LkFrame *f = &fr[0];
	f->w = arena;
	if (!unit) /* single-step entries arrive with the photon's weights in arena[0, nE) */
		for (int e = 0; e < nE; e++)
			f->w[e] = 1.;
	f->irefl = unit ? unit->irefl : 0;
	f->dtrav = d_travel0;
	const V3 raw_dir = sd;
	norm3(sd);
	const V3 start_dir = sd;
	f->ec = sc;
	f->ed = sd;
	norm3(f->ed);
	f->ev = sev;
	f->last_esc = 0;
	f->cont = CONT_TRACE;
	bool glass_start = false; /* photon.c:673-887: started inside the glass at z > 0 */
	int ret = 0;

	/* ---- launch preamble, photon.c:480-645 ---- */
	int z_id = 0;
	if (sc.z > 0)
		z_id = lk_z_index(c, sc.z);
	double ext_here = ((c.ext[z_id] - c.ext[z_id + 1]) / (c.z[z_id] - c.z[z_id + 1])) * (sc.z - c.z[z_id]) + c.ext[z_id];
	double q_i = 0, r_i = 0;
	bool finished = false;
	if (c.n_shells == 0.) {
		if (sqrt((sc.x) * (sc.x) + (sc.y) * (sc.y)) > ext_here) {
			ret = -2;
			finished = true;
		}
	} else {
		double hz = ext_here / (2. * PCB_COSPI_6 * (c.n_shells + 1));
		hex_cell(hz, sc.x, sc.y, q_i, r_i);
		if (within_pc(ext_here, sc.x, sc.y) == 0) {
			ret = -2;
			finished = true;
		}
	}
	enum { S_TRACE_LOOP, S_REFLECT, S_AFTER_REFLECT, S_FRAME_DONE, S_FINISH };
	int state = S_FINISH, esc = 0, riesc = 0;
	V3 refl_normal = {0, 0, 1};
	int unit_ret = -1;
	if (unit) { /* the caller's photon as it is: no launch preamble */
		finished = true; /* keeps the launch epilogue out */
		f->ed = raw_dir; /* capil_reflect / capil_trace normalise it themselves, once */
		f->A = unit->A;
		f->B = unit->B;
		f->ix = unit->ix;
		f->it = 0;
		f->it_end = nm;
		f->sn = unit->normal;
		f->cosalfa = 0;
		if (unit->mode == 1) {
			f->cont = CONT_UNIT;
			refl_normal = unit->normal;
			state = S_REFLECT;
		} else {
			state = S_TRACE_LOOP;
		}
	} else if (!finished) {
		f->A = (2. * q_i + r_i) * PCB_COSPI_6;
		f->B = r_i * (3. / 2);
		f->ix = lk_z_index_incl(c, sc.z);
		double rad_here, cx_here, cy_here;
		if (sc.z > 0) {
			double dz = (c.z[z_id + 1] - c.z[z_id]), off = (sc.z - c.z[z_id]);
			double cx0 = f->A * c.hexz[z_id], cx1 = f->A * c.hexz[z_id + 1], cy0 = f->B * c.hexz[z_id], cy1 = f->B * c.hexz[z_id + 1];
			rad_here = ((c.cap[z_id + 1] - c.cap[z_id]) / dz) * off + c.cap[z_id];
			cx_here = ((cx1 - cx0) / dz) * off + cx0;
			cy_here = ((cy1 - cy0) / dz) * off + cy0;
		} else {
			rad_here = c.cap[0];
			cx_here = f->A * c.hexz[0];
			cy_here = f->B * c.hexz[0];
		}
		double d_axis = sqrt((sc.x - cx_here) * (sc.x - cx_here) + (sc.y - cy_here) * (sc.y - cy_here));
		if (d_axis > rad_here) { /* starts in the glass */
			ret = 2;
			if (sc.z == 0) { /* photon.c:647-672: reflect off the entrance face, normal (0,0,1) */
				f->cont = CONT_FACE;
				refl_normal.x = 0; refl_normal.y = 0; refl_normal.z = 1;
				state = S_REFLECT;
			} else { /* photon.c:673-887 */
				double d_wall = 0;
				int r_cntr, q_cntr;
				int wall_trace = lk_trace_wall(c, f->ec, f->ed, d_wall, r_cntr, q_cntr, n_wall_steps);
				if (wall_trace <= 0) {
					ret = -1;
					state = S_FINISH;
				} else {
					for (int e = 0; e < nE; e++)
						f->w[e] = f->w[e] * exp(-1. * d_wall * c.ec[EC_AMU * nE + e]);
					double step = d_wall / sqrt(dot3(f->ed, f->ed));
					f->ec.x = f->ec.x + step * f->ed.x;
					f->ec.y = f->ec.y + step * f->ed.y;
					f->ec.z = f->ec.z + step * f->ed.z;
					ret = 1;
					glass_start = true;
					if (wall_trace == 3)
						lk_append(c, out, 0, f->ec, f->ed, f->ev, f->irefl, f->w);
					if (wall_trace == 2)
						lk_append(c, out, 1, f->ec, f->ed, f->ev, f->irefl, f->w);
					if (wall_trace == 1) {
						f->dtrav = f->dtrav + d_wall;
						if (c.n_shells == 0.) {
							f->A = 0; f->B = 0;
						} else {
							f->A = (2. * (double)q_cntr + (double)r_cntr) * PCB_COSPI_6;
							f->B = (double)r_cntr * (3. / 2);
						}
						f->ix = lk_z_index_incl(c, f->ec.z);
						f->it = 0;
						f->it_end = nm;
						state = S_TRACE_LOOP;
					} else {
						state = S_FINISH;
					}
				}
			}
		} else {
			f->it = 0;
			f->it_end = nm;
			state = S_TRACE_LOOP;
		}
	}

	while (state != S_FINISH) {
		f = &fr[sp];
		if (state == S_TRACE_LOOP) {
			if (f->it > f->it_end) { /* loop exhausted with iesc == 1 */
				esc = f->last_esc;
				state = S_FRAME_DONE;
				continue;
			}
			f->it++;
			int r = lk_trace_to_hit(c, *f);
			if (r != 1) {
				f->last_esc = r;
				esc = r;
				state = S_FRAME_DONE;
				continue;
			}
			f->cont = CONT_TRACE;
			refl_normal = f->sn;
			state = S_REFLECT;
		} else if (state == S_REFLECT) {
			/* ---- polycap_capil_reflect with leak_calc, capil.c:565-891 ---- */
			n_refl_events++;
			V3 sn = refl_normal;
			norm3(sn);
			norm3(f->ed);
			double alfa = dot3(f->ed, sn);
			if (alfa < 0.) {
				riesc = -1;
				state = S_AFTER_REFLECT;
				continue;
			}
			double d_wall = 0;
			int r_cntr = 0, q_cntr = 0;
			int wall_trace = lk_trace_wall(c, f->ec, f->ed, d_wall, r_cntr, q_cntr, n_wall_steps);
			if (wall_trace <= 0) {
				riesc = -1;
				state = S_AFTER_REFLECT;
				continue;
			}
			if (sp + 1 >= out.depth) {
				if (lk_grow(c, out, fr, arena, sp)) {
					f = &fr[sp];
				} else {
					out.err |= LK_ERR_DEPTH;
					riesc = -1;
					state = S_AFTER_REFLECT;
					continue;
				}
			}
			double *w_leak = arena + (size_t)(sp + 1) * nE;
			ReflGeom g;
			refl_geometry(sn, f->ed, f->ev, g);
			bool leak_flag = false, weight_flag = false, bad = false;
			for (int e = 0; e < nE; e++) {
				double cons1 = c.ec[EC_KR * nE + e] * alfa * c.sig_rough;
				double r_rough = exp(-1. * cons1 * cons1);
				double rtot = fresnel_rtot(c.ec[EC_NRE * nE + e], c.ec[EC_NIM * nE + e], c.ec[EC_I2RE * nE + e], c.ec[EC_I2IM * nE + e], g.cos_t,
							   g.sin_t, g.frac_s, g.frac_p);
				if (rtot < 0. || rtot > 1.) {
					bad = true;
					break;
				}
				w_leak[e] = (1. - rtot * r_rough) * f->w[e] * exp(-1. * d_wall * c.ec[EC_AMU * nE + e]);
				if (w_leak[e] >= 1.e-4)
					leak_flag = true;
				f->w[e] = f->w[e] * rtot * r_rough;
				if (f->w[e] >= 1.e-4)
					weight_flag = true;
			}
			if (bad) {
				riesc = -1;
				state = S_AFTER_REFLECT;
				continue;
			}
			f->ev = g.new_ev;
			riesc = weight_flag ? 1 : 0;
			if (leak_flag) {
				double step = d_wall / sqrt(dot3(f->ed, f->ed));
				V3 lc = {f->ec.x + step * f->ed.x, f->ec.y + step * f->ed.y, f->ec.z + step * f->ed.z};
				if (wall_trace == 1) {
					int zi = lk_z_index(c, lc.z);
					double eh = ((c.ext[zi + 1] - c.ext[zi]) / (c.z[zi + 1] - c.z[zi])) * (lc.z - c.z[zi]) + c.ext[zi];
					if (c.n_shells == 0.) {
						if (sqrt(lc.x * lc.x + lc.y * lc.y) >= eh)
							wall_trace = 3;
					} else if (within_pc(eh, lc.x, lc.y) == 0) {
						wall_trace = 3;
					}
				}
				if (wall_trace == 3)
					lk_append(c, out, 0, lc, f->ed, f->ev, f->irefl, w_leak);
				if (wall_trace == 2)
					lk_append(c, out, 1, lc, f->ed, f->ev, f->irefl, w_leak);
				if (wall_trace == 1 && lc.z < c.z[nm]) {
					/* the transmitted part continues as a secondary photon in capillary (q_cntr, r_cntr): capil.c:711-886 */
					f->iesc_after = riesc;
					f->rec_mark = out.nrec;
					LkFrame *s = &fr[sp + 1];
					s->ec = lc;
					s->ed = f->ed;
					s->ev = f->ev;
					s->irefl = f->irefl;
					s->dtrav = f->dtrav + d_wall;
					s->w = w_leak;
					if (c.n_shells == 0.) {
						s->A = 0; s->B = 0;
					} else {
						s->A = (2. * (double)q_cntr + (double)r_cntr) * PCB_COSPI_6;
						s->B = (double)r_cntr * (3. / 2);
					}
					s->ix = lk_z_index_incl(c, lc.z);
					s->it = s->ix;
					s->it_end = nm;
					s->last_esc = 0;
					s->cont = CONT_TRACE;
					sp++;
					state = S_TRACE_LOOP;
					continue;
				}
			}
			state = S_AFTER_REFLECT;
		} else if (state == S_FRAME_DONE) {
			if (sp == 0) {
				unit_ret = esc; /* single polycap_capil_trace that found no reflection: -1 / -2 / -3 */
				state = S_FINISH;
				continue;
			}
			LkFrame *s = &fr[sp];
			sp--;
			f = &fr[sp];
			if (esc == -1 || esc == -3) { /* capil.c:799-805: the secondary's leaks are dropped, reflect returns -2 */
				out.nrec = f->rec_mark;
				riesc = -2;
			} else {
				if (esc == 1 || esc == -2) { /* the secondary reached the end: its weights leave as one more leak, capil.c:845-880 */
					double tt = ((c.z[nm] - s->ec.z) / s->ed.z);
					V3 e3 = {s->ec.x + s->ed.x * tt, s->ec.y + s->ed.y * tt, s->ec.z + s->ed.z * tt};
					int inside = within_pc(c.ext[nm], e3.x, e3.y);
					if (inside == 0)
						lk_append(c, out, 0, e3, f->ed, f->ev, f->irefl, s->w);
					else if (inside == 1)
						lk_append(c, out, 1, e3, f->ed, f->ev, f->irefl, s->w);
				}
				riesc = f->iesc_after;
			}
			state = S_AFTER_REFLECT;
		} else { /* S_AFTER_REFLECT */
			if (f->cont == CONT_FACE) { /* photon.c:652: return value of the entrance-face reflect is ignored */
				state = S_FINISH;
				continue;
			}
			if (f->cont == CONT_UNIT) { /* single polycap_capil_reflect: its return value is the result */
				unit_ret = riesc;
				state = S_FINISH;
				continue;
			}
			if (riesc == 1) { /* capil.c:1346-1352 */
				f->ed.x = f->ed.x - 2.0 * f->cosalfa * f->sn.x;
				f->ed.y = f->ed.y - 2.0 * f->cosalfa * f->sn.y;
				f->ed.z = f->ed.z - 2.0 * f->cosalfa * f->sn.z;
				norm3(f->ed);
				f->irefl++;
				esc = 1;
			} else if (riesc == -1 || riesc == -2) {
				esc = -1;
			} else {
				esc = riesc;
			}
			f->last_esc = esc;
			state = (esc == 1) ? S_TRACE_LOOP : S_FRAME_DONE;
			if (unit && sp == 0) { /* single polycap_capil_trace: one wall hit of the primary, reflected */
				unit_ret = esc;
				state = S_FINISH;
			}
		}
	}

	f = &fr[0];
	if (!finished && ret != 2 && ret != -1 && !glass_start) { /* photon.c:947-954 */
		esc = f->last_esc;
		if (esc == -1 || esc == -3)
			ret = -1;
		else
			ret = (esc == 0) ? 0 : 1;
	} else if (glass_start && ret == 1) {
		esc = f->last_esc;
		bool traced = (f->it > 0);
		if (traced && (esc == -1 || esc == -3)) {
			ret = -1;
		} else {
			if (traced && (esc == 1 || esc == -2)) { /* photon.c:829-854 */
				double tt = ((c.z[nm] - f->ec.z) / f->ed.z);
				f->ec.x = f->ec.x + f->ed.x * tt;
				f->ec.y = f->ec.y + f->ed.y * tt;
				f->ec.z = f->ec.z + f->ed.z * tt;
				int inside = within_pc(c.ext[nm], f->ec.x, f->ec.y);
				if (inside == 0)
					lk_append(c, out, 0, f->ec, f->ed, f->ev, f->irefl, f->w);
				else if (inside == 1)
					lk_append(c, out, 1, f->ec, f->ed, f->ev, f->irefl, f->w);
			}
			for (int e = 0; e < nE; e++) /* photon.c:856-885 */
				f->w[e] = 0.;
			f->ec.x = c.ext[nm] + 1.;
			f->ec.y = c.ext[nm] + 1.;
			f->ec.z = c.z[nm];
			f->ed = start_dir;
			norm3(f->ed);
		}
	}
	if (unit)
		ret = unit_ret;
	res.ret = ret;
	res.ec = f->ec;
	res.ed = f->ed;
	res.ev = f->ev;
	res.irefl = f->irefl;
	res.dtrav = f->dtrav;
}

/* ------------------------------------------------------------------ kernels */

struct LkParams {
	LkCtx c;
	PcbSource s;
	int mode;                 /* MODE_SOURCE / MODE_REPLAY (pcb_trace.cuh numbering: 0 / 1) */
	const double *rin;        /* replay: coords[n][3], dir[n][3], ev[n][3] */
	unsigned long long id0, n;
	unsigned long long *next; /* work counter */
	int depth, rec_cap;       /* per-thread capacities of this launch: frames, leak records */
	const unsigned long long *todo; /* second launch: the ids (local indices) the first one could not hold; NULL: all of [0, n) */
	unsigned long long n_todo;
	unsigned long long *retry;      /* first launch: ids whose nesting depth / record count exceeded the small capacities ... */
	unsigned long long *n_retry;    /* ... and how many */
	double *arena;            /* [threads][(depth+1)*nE] */
	LkFrame *frames;          /* [threads][depth] */
	double *rec;              /* [threads][rec_cap*(11+nE)] scratch strips */
	LkFrame *big_frames;      /* pool of full-size workspaces for the photons that outgrow the small ones (first launch) */
	double *big_arena, *big_rec;
	unsigned long long *big_next, big_n;
	/* dense per-id outputs */
	int8_t *status;
	int *irefl;
	double *fields;           /* SOURCE: [17][n]; REPLAY: [11][n] */
	double *weights;          /* [n][nE] */
	int *leak_n;              /* [n] records of this id */
	unsigned long long *leak_off; /* [n] offset of this id's records in leak_pool (records) */
	double *leak_pool;        /* append pool, leak_pool_cap records */
	unsigned long long leak_pool_cap;
	unsigned long long *leak_pool_used;
	unsigned long long *refl_events;
	int *err;
};

__global__ void __launch_bounds__(LK_THREADS) pcb_leak_kernel(const LkParams P)
{
	const LkCtx &c = P.c;
	const int nE = c.nE;
	const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	double *const arena0 = P.arena + tid * (unsigned long long)(P.depth + 1) * nE;
	LkFrame *const frames0 = P.frames + tid * (unsigned long long)P.depth;
	double *const rec0 = P.rec + tid * (unsigned long long)P.rec_cap * (11 + nE);
	LkOut out;
	out.big_frames = P.big_frames; out.big_arena = P.big_arena; out.big_rec = P.big_rec; out.big_next = P.big_next; out.big_n = P.big_n;
	unsigned long long refl_events = 0, wall_steps = 0;
	for (;;) {
		unsigned long long i = atomicAdd(P.next, 1ULL);
		if (P.todo) {
			if (i >= P.n_todo)
				break;
			i = P.todo[i];
		} else if (i >= P.n) {
			break;
		}
		StartState s0;
		if (P.mode == 0) {
			source_sample(P.s, c.mono, P.id0 + i, s0);
		} else {
			const double *q = P.rin;
			s0.coords.x = q[3 * i]; s0.coords.y = q[3 * i + 1]; s0.coords.z = q[3 * i + 2];
			s0.dir.x = q[3 * P.n + 3 * i]; s0.dir.y = q[3 * P.n + 3 * i + 1]; s0.dir.z = q[3 * P.n + 3 * i + 2];
			s0.ev.x = q[6 * P.n + 3 * i]; s0.ev.y = q[6 * P.n + 3 * i + 1]; s0.ev.z = q[6 * P.n + 3 * i + 2];
			s0.src_x = s0.src_y = 0;
		}
		out.nrec = 0;
		out.err = 0;
		out.rec = rec0; /* every photon starts in the thread's own (small) workspace */
		out.cap = P.rec_cap;
		out.depth = P.depth;
		double *arena = arena0;
		LkFrame *frames = frames0;
		LkResult r;
		unsigned long long ev_photon = 0, steps_photon = 0;
		lk_photon(c, s0.coords, s0.dir, s0.ev, 0.0, frames, arena, out, r, ev_photon, steps_photon);
		if (out.err) {
			if (P.retry) { /* did not fit the small per-thread capacities: the second launch redoes this id with the full ones */
				P.retry[atomicAdd(P.n_retry, 1ULL)] = i;
				P.leak_n[i] = 0;
				P.leak_off[i] = 0;
				continue;
			}
			atomicOr(P.err, out.err);
		}
		refl_events += ev_photon;
		wall_steps += steps_photon;
		int status = r.ret;
		const double zt = c.z[c.nmax];
		if (P.mode == 0 && r.ret == 1) { /* exit-window test, source.c:762-777 */
			double px = r.ec.x + r.ed.x * (zt - r.ec.z) / r.ed.z;
			double py = r.ec.y + r.ed.y * (zt - r.ec.z) / r.ed.z;
			int inside = c.mono ? ((sqrt(px * px + py * py) > c.ext[c.nmax]) ? 0 : 1) : within_pc(c.ext[c.nmax], px, py);
			if (inside != 1)
				status = ST_MISSED_WINDOW;
		}
		P.status[i] = (int8_t)status;
		P.irefl[i] = (int)r.irefl;
		double *F = P.fields;
		const unsigned long long n = P.n;
		if (P.mode != 0) {
			F[0 * n + i] = r.ec.x; F[1 * n + i] = r.ec.y; F[2 * n + i] = r.ec.z;
			F[3 * n + i] = r.ed.x; F[4 * n + i] = r.ed.y; F[5 * n + i] = r.ed.z;
			F[6 * n + i] = r.ev.x; F[7 * n + i] = r.ev.y; F[8 * n + i] = r.ev.z;
			F[9 * n + i] = r.dtrav;
			F[10 * n + i] = 0;
		} else if (status == ST_EXIT) { /* record of source.c:779-798, 893-923 */
			V3 sd = s0.dir;
			norm3(sd);
			norm3(sd);
			F[0 * n + i] = s0.src_x; F[1 * n + i] = s0.src_y;
			F[2 * n + i] = s0.coords.x; F[3 * n + i] = s0.coords.y;
			F[4 * n + i] = sd.x; F[5 * n + i] = sd.y;
			double ca = dot3(s0.ev, sd);
			double al = acos(ca);
			double c_ae = 1. / sin(al), c_be = -1. * c_ae * ca;
			V3 t = {s0.ev.x * c_ae + sd.x * c_be, s0.ev.y * c_ae + sd.y * c_be, s0.ev.z * c_ae + sd.z * c_be};
			norm3(t);
			F[6 * n + i] = round(t.x); F[7 * n + i] = round(t.y);
			double px = r.ec.x + r.ed.x * (zt - r.ec.z) / r.ed.z;
			double py = r.ec.y + r.ed.y * (zt - r.ec.z) / r.ed.z;
			double pz = r.ec.z + r.ed.z * (zt - r.ec.z) / r.ed.z;
			F[8 * n + i] = px; F[9 * n + i] = py; F[10 * n + i] = pz;
			F[11 * n + i] = r.ed.x; F[12 * n + i] = r.ed.y;
			V3 u = {r.ev.x * c_ae + r.ed.x * c_be, r.ev.y * c_ae + r.ed.y * c_be, r.ev.z * c_ae + r.ed.z * c_be};
			norm3(u);
			F[13 * n + i] = round(u.x); F[14 * n + i] = round(u.y);
			F[15 * n + i] = (double)r.irefl;
			F[16 * n + i] = r.dtrav + sqrt((px - r.ec.x) * (px - r.ec.x) + (py - r.ec.y) * (py - r.ec.y) + (zt - r.ec.z) * (zt - r.ec.z));
		}
		for (int e = 0; e < nE; e++)
			P.weights[i * nE + e] = arena[e];
		/* leak records of this id -> pool (the reference keeps the leaks of photons with codes 1, 0, 2; source.c:799-879) */
		int keep = out.nrec;
		if (P.mode == 0 && (status == ST_OUTSIDE || status == ST_ERROR))
			keep = 0;
		P.leak_n[i] = keep;
		if (keep > 0) {
			unsigned long long off = atomicAdd(P.leak_pool_used, (unsigned long long)keep);
			P.leak_off[i] = off;
			if (off + keep <= P.leak_pool_cap) {
				const size_t stride = 11 + nE;
				for (size_t k = 0; k < (size_t)keep * stride; k++)
					P.leak_pool[off * stride + k] = out.rec[k];
			} else {
				atomicOr(P.err, LK_ERR_POOL);
			}
		} else {
			P.leak_off[i] = 0;
		}
	}
	atomicAdd(P.refl_events, refl_events);
	atomicAdd(P.refl_events + 1, wall_steps);
}

/* single-item kernels for the reference's private test entry points */
__global__ void lk_unit_trace_wall_kernel(LkCtx c, const double *in, double *outd, int *outi)
{
	V3 ec = {in[0], in[1], in[2]}, ed = {in[3], in[4], in[5]};
	double d_travel;
	int r_cntr, q_cntr;
	unsigned long long steps = 0;
	outi[0] = lk_trace_wall(c, ec, ed, d_travel, r_cntr, q_cntr, steps);
	outi[1] = r_cntr;
	outi[2] = q_cntr;
	outd[0] = d_travel;
}

/* single polycap_capil_reflect / polycap_capil_trace with leak_calc = true on one photon state:
 * in[16] = coords(3) direction(3) elecv(3) normal(3) i_refl d_travel A B; ix; w[nE] in/out; outs[13] = ret coords(3) direction(3) elecv(3) i_refl d_travel ix */
__global__ void lk_unit_step_kernel(LkCtx c, int mode, const double *in, int ix, double *w, LkFrame *frames, double *arena, double *rec, int rec_cap, int depth,
				    double *outs, int *nrec_out, int *err_out)
{
	LkOut out;
	out.rec = rec; out.nrec = 0; out.err = 0; out.cap = rec_cap; out.depth = depth;
	out.big_frames = nullptr; out.big_arena = nullptr; out.big_rec = nullptr; out.big_next = nullptr; out.big_n = 0;
	for (int e = 0; e < c.nE; e++)
		arena[e] = w[e];
	LkUnit u;
	u.mode = mode;
	u.normal.x = in[9]; u.normal.y = in[10]; u.normal.z = in[11];
	u.irefl = (long long)in[12];
	u.A = in[14]; u.B = in[15];
	u.ix = ix;
	V3 sc = {in[0], in[1], in[2]}, sd = {in[3], in[4], in[5]}, sev = {in[6], in[7], in[8]};
	LkResult r;
	unsigned long long ev = 0, steps = 0;
	LkFrame *fr = frames;
	double *ar = arena;
	lk_photon(c, sc, sd, sev, in[13], fr, ar, out, r, ev, steps, &u);
	for (int e = 0; e < c.nE; e++)
		w[e] = ar[e];
	outs[0] = (double)r.ret;
	outs[1] = r.ec.x; outs[2] = r.ec.y; outs[3] = r.ec.z;
	outs[4] = r.ed.x; outs[5] = r.ed.y; outs[6] = r.ed.z;
	outs[7] = r.ev.x; outs[8] = r.ev.y; outs[9] = r.ev.z;
	outs[10] = (double)r.irefl;
	outs[11] = r.dtrav;
	outs[12] = (double)fr[0].ix;
	*nrec_out = out.nrec;
	*err_out = out.err;
}

/* ------------------------------------------------------------------ host side */

static thread_local std::string g_lk_err;
extern "C" const char *pcb_leak_last_error(void) { return g_lk_err.c_str(); }

#define LCK(call)                                                                                                   \
	do {                                                                                                        \
		cudaError_t e_ = (call);                                                                            \
		if (e_ != cudaSuccess) {                                                                            \
			char b_[512];                                                                               \
			snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
			g_lk_err = b_;                                                                              \
			return -1;                                                                                  \
		}                                                                                                   \
	} while (0)

struct DevMem {
	std::vector<void *> p;
	~DevMem()
	{
		for (void *q : p)
			cudaFree(q);
	}
	template <typename T> cudaError_t alloc(T **out, size_t bytes)
	{
		void *q = nullptr;
		cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
		if (e == cudaSuccess)
			p.push_back(q);
		*out = (T *)q;
		return e;
	}
};

/* shared driver: runs n photons (ids or replay states), fills host outputs */
extern "C" int pcb_leak_run(const pcb_leak_optic *o, const pcb_source_desc *src, uint64_t seed, int64_t id0, int64_t n, const double *coords,
			    const double *dir, const double *ev, int8_t *status, int *irefl, double *fields /* [F][n] field-major */, double *weights,
			    int *leak_n, double **leak_rec, int64_t *n_leak_rec, int64_t *work /* [2]: capil_reflect events, wall-march steps */)
{
	const int nE = o->nE;
	const int mode = src ? 0 : 1;
	const int nfields = mode == 0 ? 17 : 11;
	DevMem M;
	LkParams P{};
	P.c.nmax = o->nmax;
	P.c.nE = nE;
	P.c.mono = o->mono;
	P.c.monotonic = o->monotonic;
	P.c.no_fast_march = getenv("POLYCAP_B200_LEAK_FULL_MARCH") ? 1 : 0;
	P.c.n_shells = o->n_shells;
	P.c.sig_rough = o->sig_rough;
	P.c.z = o->d_z; P.c.cap = o->d_cap; P.c.ext = o->d_ext; P.c.hexz = o->d_hexz; P.c.ec = o->d_ec;
	if (src) {
		P.s.d_source = src->d_source; P.s.src_x = src->src_x; P.s.src_y = src->src_y; P.s.src_sigx = src->src_sigx; P.s.src_sigy = src->src_sigy;
		P.s.src_shiftx = src->src_shiftx; P.s.src_shifty = src->src_shifty; P.s.hor_pol = src->hor_pol;
		P.s.cap0 = o->cap0; P.s.ext0 = o->ext0; P.s.seed = seed;
	}
	P.mode = mode;
	P.id0 = (unsigned long long)id0;
	P.n = (unsigned long long)n;
	/* Per-thread workspace: frame stack, weight arena, record strip; the thread count follows from a memory budget.  Few photons
	 * nest deeper than a few dozen secondaries, but one that does needs LK_DEPTH frames and LK_REC_CAP records, i.e. megabytes per
	 * thread at long energy grids -- so the first launch gives every thread small capacities (many threads), collects the ids that
	 * did not fit, and a second launch redoes only those with the full capacities (few threads). */
	const size_t stride = 11 + (size_t)nE;
	auto per_thread_bytes = [&](int depth, int rcap) { return (size_t)depth * sizeof(LkFrame) + (size_t)(depth + 1) * nE * 8 + (size_t)rcap * stride * 8; };
	const int depth1 = getenv("POLYCAP_B200_LEAK_DEPTH1") ? atoi(getenv("POLYCAP_B200_LEAK_DEPTH1")) : 48;
	const int rcap1 = 4 * depth1;
	size_t free_b = 0, total_b = 0;
	LCK(cudaMemGetInfo(&free_b, &total_b));
	size_t budget = free_b / 3;
	if (budget > ((size_t)24 << 30))
		budget = (size_t)24 << 30;
	auto blocks_for = [&](int64_t items, int depth, int rcap) {
		int blocks = (int)((items + LK_THREADS - 1) / LK_THREADS);
		int max_blocks = (int)(budget / (per_thread_bytes(depth, rcap) * LK_THREADS));
		if (max_blocks > o->sm_count * 4)
			max_blocks = o->sm_count * 4;
		if (max_blocks < 1)
			max_blocks = 1;
		return blocks > max_blocks ? max_blocks : blocks;
	};
	int blocks = blocks_for(n, depth1, rcap1);
	size_t threads = (size_t)blocks * LK_THREADS;
	double *d_rin = nullptr;
	P.depth = depth1;
	P.rec_cap = rcap1;
	LCK(M.alloc(&P.next, 8));
	LCK(M.alloc(&P.retry, (size_t)n * 8));
	LCK(M.alloc(&P.n_retry, 8));
	LCK(cudaMemset(P.n_retry, 0, 8));
	LCK(M.alloc(&P.arena, threads * (size_t)(depth1 + 1) * nE * 8));
	LCK(M.alloc(&P.frames, threads * (size_t)depth1 * sizeof(LkFrame)));
	LCK(M.alloc(&P.rec, threads * (size_t)rcap1 * stride * 8));
	{ /* full-size workspaces for the deep photons of this launch: a quarter of what is left, at most 16 GB, at most one per id */
		size_t fb = 0, tb = 0;
		LCK(cudaMemGetInfo(&fb, &tb));
		size_t big_budget = fb / 4;
		if (big_budget > ((size_t)16 << 30))
			big_budget = (size_t)16 << 30;
		unsigned long long nb = big_budget / per_thread_bytes(LK_DEPTH, LK_REC_CAP);
		if (nb > (unsigned long long)n)
			nb = (unsigned long long)n;
		if (nb > 4096)
			nb = 4096;
		P.big_n = nb;
		LCK(M.alloc(&P.big_next, 8));
		LCK(cudaMemset(P.big_next, 0, 8));
		LCK(M.alloc(&P.big_frames, (size_t)nb * LK_DEPTH * sizeof(LkFrame)));
		LCK(M.alloc(&P.big_arena, (size_t)nb * (LK_DEPTH + 1) * nE * 8));
		LCK(M.alloc(&P.big_rec, (size_t)nb * LK_REC_CAP * stride * 8));
	}
	LCK(M.alloc(&P.status, (size_t)n));
	LCK(M.alloc(&P.irefl, (size_t)n * 4));
	LCK(M.alloc(&P.fields, (size_t)n * nfields * 8));
	LCK(M.alloc(&P.weights, (size_t)n * nE * 8));
	LCK(M.alloc(&P.leak_n, (size_t)n * 4));
	LCK(M.alloc(&P.leak_off, (size_t)n * 8));
	/* record pool of the wave: 24 leaks per id on average, within a third of the free memory; a wave that overflows it is
	 * reported with return code -2 and retried by the caller with fewer ids (the reference simply reallocs per photon) */
	P.leak_pool_cap = (unsigned long long)n * 24 + 65536;
	{
		const unsigned long long fit = (unsigned long long)(free_b / 3) / (stride * 8);
		if (P.leak_pool_cap > fit)
			P.leak_pool_cap = fit > 4096 ? fit : 4096;
	}
	LCK(M.alloc(&P.leak_pool, (size_t)P.leak_pool_cap * stride * 8));
	LCK(M.alloc(&P.leak_pool_used, 8));
	LCK(M.alloc(&P.refl_events, 16));
	LCK(M.alloc(&P.err, 4));
	LCK(cudaMemset(P.next, 0, 8));
	LCK(cudaMemset(P.leak_pool_used, 0, 8));
	LCK(cudaMemset(P.refl_events, 0, 16));
	LCK(cudaMemset(P.err, 0, 4));
	if (mode == 1) {
		LCK(M.alloc(&d_rin, (size_t)n * 9 * 8));
		LCK(cudaMemcpy(d_rin, coords, (size_t)n * 24, cudaMemcpyHostToDevice));
		LCK(cudaMemcpy(d_rin + 3 * n, dir, (size_t)n * 24, cudaMemcpyHostToDevice));
		LCK(cudaMemcpy(d_rin + 6 * n, ev, (size_t)n * 24, cudaMemcpyHostToDevice));
		P.rin = d_rin;
	}
	pcb_leak_kernel<<<blocks, LK_THREADS>>>(P);
	LCK(cudaGetLastError());
	LCK(cudaDeviceSynchronize());
	unsigned long long n_retry = 0;
	LCK(cudaMemcpy(&n_retry, P.n_retry, 8, cudaMemcpyDeviceToHost));
	if (n_retry > 0) { /* second launch: the deep photons with the full capacities */
		DevMem M2;
		LkParams Q = P;
		Q.depth = LK_DEPTH;
		Q.rec_cap = LK_REC_CAP;
		Q.todo = P.retry;
		Q.n_todo = n_retry;
		Q.retry = nullptr;
		Q.n_retry = nullptr;
		Q.big_next = nullptr;
		Q.big_n = 0;
		const int blocks2 = blocks_for((int64_t)n_retry, LK_DEPTH, LK_REC_CAP);
		const size_t threads2 = (size_t)blocks2 * LK_THREADS;
		LCK(M2.alloc(&Q.arena, threads2 * (size_t)(LK_DEPTH + 1) * nE * 8));
		LCK(M2.alloc(&Q.frames, threads2 * (size_t)LK_DEPTH * sizeof(LkFrame)));
		LCK(M2.alloc(&Q.rec, threads2 * (size_t)LK_REC_CAP * stride * 8));
		LCK(cudaMemset(Q.next, 0, 8));
		pcb_leak_kernel<<<blocks2, LK_THREADS>>>(Q);
		LCK(cudaGetLastError());
		LCK(cudaDeviceSynchronize());
	}
	if (getenv("POLYCAP_B200_PHASES"))
	{
		unsigned long long moved = 0;
		cudaMemcpy(&moved, P.big_next, 8, cudaMemcpyDeviceToHost);
		fprintf(stderr, "polycap-b200 phases: leak wave of %lld ids: %d x %d threads with %d frames each, %llu photons moved to a full-size workspace (%llu available), "
				"%llu ids redone in a second launch\n", (long long)n, blocks, LK_THREADS, depth1, moved, P.big_n, n_retry);
	}
	int err = 0;
	unsigned long long used = 0, refl[2] = {0, 0};
	LCK(cudaMemcpy(&err, P.err, 4, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(&used, P.leak_pool_used, 8, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(refl, P.refl_events, 16, cudaMemcpyDeviceToHost));
	if (err & (LK_ERR_DEPTH | LK_ERR_RECORDS)) {
		g_lk_err = (err & LK_ERR_DEPTH) ? "leak transport: more than 1400 nested secondary photons" : "leak transport: more than 2048 leak records from one photon";
		return -1;
	}
	if (err & LK_ERR_POOL) {
		g_lk_err = "leak transport: the wave's leak-record pool is full (retry with fewer photon ids per wave)";
		return -2;
	}
	if (work) {
		work[0] = (int64_t)refl[0];
		work[1] = (int64_t)refl[1];
	}
	LCK(cudaMemcpy(status, P.status, (size_t)n, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(irefl, P.irefl, (size_t)n * 4, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(fields, P.fields, (size_t)n * nfields * 8, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(weights, P.weights, (size_t)n * nE * 8, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(leak_n, P.leak_n, (size_t)n * 4, cudaMemcpyDeviceToHost));
	/* records: pool order is arrival order; return them in id order */
	std::vector<unsigned long long> off((size_t)n);
	LCK(cudaMemcpy(off.data(), P.leak_off, (size_t)n * 8, cudaMemcpyDeviceToHost));
	std::vector<double> pool((size_t)used * stride);
	if (used)
		LCK(cudaMemcpy(pool.data(), P.leak_pool, pool.size() * 8, cudaMemcpyDeviceToHost));
	double *recs = (double *)malloc((used ? (size_t)used : 1) * stride * 8);
	size_t w = 0;
	for (int64_t i = 0; i < n; i++)
		if (leak_n[i] > 0) {
			memcpy(recs + w * stride, pool.data() + (size_t)off[(size_t)i] * stride, (size_t)leak_n[i] * stride * 8);
			w += (size_t)leak_n[i];
		}
	*leak_rec = recs;
	*n_leak_rec = (int64_t)w;
	return 0;
}

extern "C" int pcb_leak_unit_step(const pcb_leak_optic *o, int mode, const double in[16], int ix, double *weights, double outs[13], double **leak_rec,
				  int64_t *n_leak_rec)
{
	LkCtx c{};
	c.nmax = o->nmax; c.nE = o->nE; c.mono = o->mono; c.monotonic = o->monotonic; c.n_shells = o->n_shells; c.sig_rough = o->sig_rough;
	c.no_fast_march = getenv("POLYCAP_B200_LEAK_FULL_MARCH") ? 1 : 0;
	c.z = o->d_z; c.cap = o->d_cap; c.ext = o->d_ext; c.hexz = o->d_hexz; c.ec = o->d_ec;
	const size_t nE = (size_t)o->nE, stride = 11 + nE;
	*leak_rec = nullptr;
	*n_leak_rec = 0;
	DevMem M;
	double *d_in, *d_w, *d_arena, *d_rec, *d_outs;
	LkFrame *d_frames;
	int *d_i;
	LCK(M.alloc(&d_in, 16 * 8));
	LCK(M.alloc(&d_w, nE * 8));
	LCK(M.alloc(&d_arena, (size_t)(LK_DEPTH + 1) * nE * 8));
	LCK(M.alloc(&d_frames, (size_t)LK_DEPTH * sizeof(LkFrame)));
	LCK(M.alloc(&d_rec, (size_t)LK_REC_CAP * stride * 8));
	LCK(M.alloc(&d_outs, 13 * 8));
	LCK(M.alloc(&d_i, 8));
	LCK(cudaMemcpy(d_in, in, 16 * 8, cudaMemcpyHostToDevice));
	LCK(cudaMemcpy(d_w, weights, nE * 8, cudaMemcpyHostToDevice));
	lk_unit_step_kernel<<<1, 1>>>(c, mode, d_in, ix, d_w, d_frames, d_arena, d_rec, LK_REC_CAP, LK_DEPTH, d_outs, d_i, d_i + 1);
	LCK(cudaGetLastError());
	int hi[2];
	LCK(cudaMemcpy(hi, d_i, 8, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(outs, d_outs, 13 * 8, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(weights, d_w, nE * 8, cudaMemcpyDeviceToHost));
	if (hi[1]) {
		g_lk_err = (hi[1] & LK_ERR_DEPTH) ? "leak transport: more than 1400 nested secondary photons" : "leak transport: more than 2048 leak records from one photon";
		return -1;
	}
	if (hi[0] > 0) {
		double *recs = (double *)malloc((size_t)hi[0] * stride * 8);
		LCK(cudaMemcpy(recs, d_rec, (size_t)hi[0] * stride * 8, cudaMemcpyDeviceToHost));
		*leak_rec = recs;
		*n_leak_rec = hi[0];
	}
	return 0;
}

extern "C" int pcb_leak_unit_trace_wall(const pcb_leak_optic *o, const double in[6], double *d_travel, int *r_cntr, int *q_cntr)
{
	LkCtx c{};
	c.nmax = o->nmax; c.nE = o->nE; c.mono = o->mono; c.monotonic = o->monotonic; c.n_shells = o->n_shells; c.sig_rough = o->sig_rough;
	c.no_fast_march = getenv("POLYCAP_B200_LEAK_FULL_MARCH") ? 1 : 0;
	c.z = o->d_z; c.cap = o->d_cap; c.ext = o->d_ext; c.hexz = o->d_hexz; c.ec = o->d_ec;
	DevMem M;
	double *d_in, *d_out;
	int *d_i;
	LCK(M.alloc(&d_in, 48));
	LCK(M.alloc(&d_out, 8));
	LCK(M.alloc(&d_i, 12));
	LCK(cudaMemcpy(d_in, in, 48, cudaMemcpyHostToDevice));
	lk_unit_trace_wall_kernel<<<1, 1>>>(c, d_in, d_out, d_i);
	LCK(cudaGetLastError());
	int hi[3];
	LCK(cudaMemcpy(hi, d_i, 12, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(d_travel, d_out, 8, cudaMemcpyDeviceToHost));
	*r_cntr = hi[1];
	*q_cntr = hi[2];
	return hi[0];
}
