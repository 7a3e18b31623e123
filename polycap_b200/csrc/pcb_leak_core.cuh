/*
 * pcb_leak_core.cuh -- photon transport WITH leak / recap wall transport (leak_calc = true): the per-photon state machine.
 *
 * Reference behaviour reproduced (paths relative to /root/reference):
 *   polycap_capil_trace_wall      src/polycap-capil.c:896-1194   march through the glass to the next bore / out of the optic
 *   polycap_photon_pc_intersect   src/polycap-photon.c:174-362   where a photon left through the outer hexagonal surface
 *   polycap_capil_reflect (leak)  src/polycap-capil.c:565-891    leak weights, ext/int leak records, the secondary ("recap")
 *                                                                photon that continues in the neighbouring capillary
 *   polycap_capil_trace           src/polycap-capil.c:1197-1361  with the per-segment in-optic checks
 *   polycap_photon_launch (leak)  src/polycap-photon.c:645-907   photons that start in the glass
 *
 * ONE WARP owns one primary photon.  Control flow and trajectory arithmetic are warp-uniform (every lane holds the same photon
 * state); what the lanes share out is the work that is wide:
 *   - the energy loop of a reflection (capil.c:625-650): lane l owns energies l, l+32, ... of the weight vectors, flags by vote;
 *   - the segment scan of a trace call (capil.c:1256-1343): 32 consecutive segments are solved at once, the first segment in
 *     table order with an outcome (hit, photon or axis outside the optic) decides, exactly as the sequential scan would;
 *   - copies of weight vectors and leak records.
 * The reference recurses -- capil_reflect creates a secondary photon and runs capil_trace on it, whose capil_reflect may create
 * another -- which here is an explicit frame stack per warp: a frame is a photon with its weight vector, capillary axis, segment
 * index and trace-loop counter; a parent waits in its frame while its secondary runs.  Leak records are streamed in depth-first
 * order into the warp's record strip, which preserves the reference's order within the external and the internal list, and are
 * rolled back when the reference discards a secondary's leaks (capil.c:799-805).
 *
 * The wall march (capil.c:1018-1068) advances the path length by a tenth of the bore radius per step, thousands of steps per
 * reflection, nearly all of them uneventful (the point is between the photon's own bore and the boundary of its hexagonal cell).
 * Uneventful steps are not walked: lk_march_jump() finds, from the ray's intersections with the cell's edge planes, the bore cone
 * and the end of the z segment, how many steps stay strictly inside that zone and advances the path length by exactly the value
 * the reference's repeated floating-point additions reach (see there); the steps around every event are taken one by one.
 *
 * All trajectory arithmetic uses the bit-faithful primitives of pcb_math.h; weights use the reference-order Fresnel term.
 *
 * This header also compiles for the HOST (g++, one lane per "warp") so that tests/ can run the same state machine on the CPU
 * against the reference's leak corpus (tests/host_leak.cpp, tests/test_host_leak.py).  That build is test infrastructure: the
 * library has no CPU path, its only user of this header is pcb_leak.cu.
 */
#ifndef PCB_LEAK_CORE_CUH
#define PCB_LEAK_CORE_CUH

#include "pcb_codes.h"
#include "pcb_math.h"

#ifdef __CUDACC__
/* LK_DEV functions are inlined by default.  -DLK_OUTLINE keeps one copy of each out of line (22 000 -> 13 000 instructions); measured
 * on the B200 that is 1.5x SLOWER per reflection at the same occupancy (calls, by-reference arguments through local memory) -- the
 * kernel is bound by the latency of its dependent FP64 chains, not by instruction supply. */
#ifdef LK_OUTLINE
#define LK_DEV __device__ __noinline__
#else
#define LK_DEV __device__
#endif
#define LK_INL __device__ __forceinline__
#define LK_LANES 32
#define LK_FULL 0xffffffffu
LK_INL int lk_lane() { return (int)(threadIdx.x & 31u); }
LK_INL bool lk_any(bool p) { return __any_sync(LK_FULL, p) != 0; }
LK_INL unsigned lk_ballot(bool p) { return __ballot_sync(LK_FULL, p); }
LK_INL int lk_first(unsigned m) { return __ffs((int)m) - 1; }
LK_INL int lk_bcast(int v, int src) { return __shfl_sync(LK_FULL, v, src); }
LK_INL double lk_bcast(double v, int src) { return __shfl_sync(LK_FULL, v, src); }
LK_INL unsigned long long lk_bcast(unsigned long long v, int src) { return __shfl_sync(LK_FULL, v, src); }
LK_INL int lk_min(int v)
{
	for (int o = 16; o > 0; o >>= 1) {
		const int t = __shfl_xor_sync(LK_FULL, v, o);
		v = t < v ? t : v;
	}
	return v;
}
LK_INL void lk_sync() { __syncwarp(); }
LK_INL unsigned long long lk_fetch_add(unsigned long long *p, unsigned long long v) /* one atomic per warp, result to every lane */
{
	unsigned long long r = 0;
	if (lk_lane() == 0)
		r = atomicAdd(p, v);
	return lk_bcast(r, 0);
}
LK_INL double lk_from_bits(unsigned long long u) { return __longlong_as_double((long long)u); }
#else
#define LK_DEV static
#define LK_INL static inline
#define LK_LANES 1
LK_INL int lk_lane() { return 0; }
LK_INL bool lk_any(bool p) { return p; }
LK_INL unsigned lk_ballot(bool p) { return p ? 1u : 0u; }
LK_INL int lk_first(unsigned m) { return m ? 0 : -1; }
LK_INL int lk_bcast(int v, int) { return v; }
LK_INL double lk_bcast(double v, int) { return v; }
LK_INL unsigned long long lk_bcast(unsigned long long v, int) { return v; }
LK_INL int lk_min(int v) { return v; }
LK_INL void lk_sync() {}
LK_INL unsigned long long lk_fetch_add(unsigned long long *p, unsigned long long v)
{
	unsigned long long r = *p;
	*p += v;
	return r;
}
LK_INL double lk_from_bits(unsigned long long u)
{
	double d;
	memcpy(&d, &u, 8);
	return d;
}
#endif

/* LK_PHASED (the kernel's default; -DLK_UNPHASED turns it off): the warps of a block pass through the three long phases of a reflection
 * event -- segment scan of the trace call (T), wall march (W), energy loop (E) -- in step: a block-wide clock ticks T, W, E, T, ... with a
 * barrier per tick, and a warp enters a phase only on that phase's tick, so the SM's instruction caches hold one phase's code at a time.
 * Measured on the B200: the phases themselves run 2-3x faster (wall march 177 000 -> 90 000 clocks per reflection event, scan 31 000 ->
 * 11 000, energy loop 15 000 -> 5 000 at 8 warps per SM) and a third of the warp time goes to waiting at the barriers: 1.3x net.
 * Every barrier of the kernel is the same __syncthreads_or(active): a warp that has run out of photons keeps ticking with active = 0
 * until no warp of the block is active. */
#if defined(LK_PHASED) && defined(__CUDACC__)
enum { LK_PH_T = 0, LK_PH_W = 1, LK_PH_E = 2 };
struct LkPhase {
	int clock; /* the phase of the current tick: identical in every warp of the block */
};
LK_INL void lk_phase_enter(LkPhase &ph, int want)
{
	for (;;) {
		__syncthreads_or(1);
		const int now = ph.clock;
		ph.clock = now == LK_PH_E ? LK_PH_T : now + 1;
		if (now == want)
			return;
	}
}
#define LK_PHASE(want) lk_phase_enter(wk.phase, want)
#else
#define LK_PHASE(want) ((void)0)
#endif

#ifndef LK_W_BUDGET
#define LK_W_BUDGET 0 /* tuning variant, not used: segment / binade crossings of a wall march per tick of the block's clock (0: the whole march in one tick).  Marches average 8 segments and 10 % exceed 16, so the block waits for its longest march; taking marches in stretches of 5 / 8 / 12 crossings was measured 41 % / 25 % / 24 % SLOWER on the B200 (a warp that comes back for the next stretch sits out a whole turn of the clock) */
#endif

/* Capacities per primary photon.  A hard photon at a steep angle crosses capillary after capillary, each crossing one
 * nested secondary (the reference recurses just as deep: its corpus shows 440 leak records and i_refl 568 on one photon). */
#define LK_DEPTH 1400    /* frames: primary + nested secondaries */
#define LK_REC_CAP 2048  /* leak records */

enum { LK_ERR_DEPTH = 1, LK_ERR_RECORDS = 2, LK_ERR_POOL = 4 }; /* RECORDS: one photon exceeds LK_REC_CAP leaks; POOL: the wave's record pool is full */

struct LkCtx {
	int nmax, nE, mono, monotonic;
	int no_fast_march; /* POLYCAP_B200_LEAK_FULL_MARCH=1: every step of the wall march takes the reference's full step (checking aid) */
	int no_jump;       /* POLYCAP_B200_LEAK_NO_JUMP=1: uneventful steps are recognised one by one instead of jumped over (checking aid) */
	double n_shells, sig_rough;
	const double *z, *cap, *ext, *hexz; /* global */
	/* per z segment i, computed once on the host with the expressions the march uses (IEEE division: the same bits):
	 * se = (ext[i+1]-ext[i])/(z[i+1]-z[i]), sc = (cap[i+1]-cap[i])/(z[i+1]-z[i]), sl = cap[i]/10 (the step length), isl = 1/sl */
	const double *se, *sc, *sl, *isl;
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
	double *rec;    /* this warp's strip: cap x (11+nE) */
	int nrec;
	int err;
	int cap;        /* records the strip holds */
	int depth;      /* frames this warp's stack holds */
	/* pool of full-size workspaces (LK_DEPTH frames, LK_REC_CAP records) a photon moves to when it outgrows the warp's small one */
	LkFrame *big_frames;
	double *big_arena, *big_rec;
	unsigned long long *big_next;
	unsigned long long big_n;
};

/* work counters of one photon */
struct LkWork {
	unsigned long long refl_events; /* capil_reflect calls incl. secondaries */
	unsigned long long wall_steps;  /* steps of the reference's wall march this photon stands for (capil.c:1018-1068) */
	unsigned long long wall_evals;  /* steps actually evaluated (the rest was jumped over) */
	unsigned long long scan_rounds; /* rounds of the lane-parallel segment scans */
	unsigned long long scan_solves; /* full solves of the scan for the next bore (per lane; the other segments were ruled out from the quadratic's coefficients) */
#if defined(LK_PHASED) && defined(__CUDACC__)
	LkPhase phase;
#endif
#ifdef LK_PROFILE
	unsigned long long t_total, t_wall, t_jump, t_scan, t_energy, n_jump_try; /* SM clocks per phase (tuning builds: -DLK_PROFILE) */
#endif
};

#if defined(LK_PROFILE) && defined(__CUDA_ARCH__)
#define LK_CLOCK() ((unsigned long long)clock64())
#define LK_TIME(field, t0) (wk.field += LK_CLOCK() - (t0))
#define LK_COUNT(field) (wk.field++)
#else
#define LK_CLOCK() 0ULL
#define LK_TIME(field, t0) ((void)(t0))
#define LK_COUNT(field) ((void)0)
#endif

/* The photon's nesting is about to exceed the warp's small workspace: move its frames [0, sp], their weight vectors and its
 * records to a full-size workspace from the pool and carry on there.  false: already full-size, or the pool is used up. */
LK_DEV bool lk_grow(const LkCtx &c, LkOut &o, LkFrame *&fr, double *&arena, int sp)
{
	if (o.depth >= LK_DEPTH || !o.big_next)
		return false;
	const unsigned long long idx = lk_fetch_add(o.big_next, 1ULL);
	if (idx >= o.big_n)
		return false;
	const size_t nE = (size_t)c.nE, stride = 11 + nE;
	LkFrame *nf = o.big_frames + idx * (unsigned long long)LK_DEPTH;
	double *na = o.big_arena + idx * (unsigned long long)(LK_DEPTH + 1) * nE;
	double *nr = o.big_rec + idx * (unsigned long long)LK_REC_CAP * stride;
	lk_sync(); /* the weight vectors and records were written lane-wise */
#pragma unroll 1
	for (size_t k = lk_lane(); k < (size_t)(sp + 1) * nE; k += LK_LANES)
		na[k] = arena[k];
#pragma unroll 1
	for (int k = lk_lane(); k <= sp; k += LK_LANES) {
		nf[k] = fr[k];
		nf[k].w = na + (fr[k].w - arena);
	}
#pragma unroll 1
	for (size_t k = lk_lane(); k < (size_t)o.nrec * stride; k += LK_LANES)
		nr[k] = o.rec[k];
	lk_sync();
	fr = nf;
	arena = na;
	o.rec = nr;
	o.cap = LK_REC_CAP;
	o.depth = LK_DEPTH;
	return true;
}

LK_DEV int lk_z_index(const LkCtx &c, double zq) /* last i < nmax with z[i] <= zq, 0 when none */
{
	if (c.monotonic) {
		if (!(c.z[0] <= zq))
			return 0;
		int lo = 0, hi = c.nmax - 1; /* invariant: z[lo] <= zq */
#pragma unroll 1
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
#pragma unroll 1
	for (int i = 0; i < c.nmax; i++)
		if (c.z[i] <= zq)
			id = i;
	return id;
}

LK_DEV int lk_z_index_incl(const LkCtx &c, double zq) /* last i <= nmax with z[i] <= zq (photon.c:626, capil.c:747) */
{
	int ix = 0;
	if (c.monotonic) {
		ix = lk_z_index(c, zq);
		if (c.z[c.nmax] <= zq)
			ix = c.nmax;
		return ix;
	}
#pragma unroll 1
	for (int i = 0; i <= c.nmax; i++)
		if (c.z[i] <= zq)
			ix = i;
	return ix;
}

LK_DEV int lk_within_pc(double R, double x, double y) { return within_pc(R, x, y); }

struct LkCell {
	double q, r;
};
LK_DEV LkCell lk_hex_cell(double hexz, double x, double y)
{
	LkCell h;
	hex_cell(hexz, x, y, h.q, h.r);
	return h;
}

LK_DEV V3 lk_propagate_xyz(double px, double py, double pz, double dx, double dy, double dz, double zt)
{
	V3 r;
	r.x = px + dx * (zt - pz) / dz;
	r.y = py + dy * (zt - pz) / dz;
	r.z = zt;
	return r;
}
LK_INL V3 lk_propagate(const V3 &p, const V3 &d, double zt) { return lk_propagate_xyz(p.x, p.y, p.z, d.x, d.y, d.z, zt); }

/* one leak record; the weight vector w is held lane-wise (lane l: entries l, l+32, ...) */
LK_DEV void lk_append(const LkCtx &c, LkOut &o, int kind, const V3 &co, const V3 &di, const V3 &el, long long nrefl, const double *w)
{
	if (o.nrec >= o.cap) {
		o.err |= LK_ERR_RECORDS;
		return;
	}
	double *r = o.rec + (size_t)o.nrec * (11 + c.nE);
	if (lk_lane() == 0) {
		r[0] = (double)kind;
		r[1] = co.x; r[2] = co.y; r[3] = co.z;
		r[4] = di.x; r[5] = di.y; r[6] = di.z;
		r[7] = el.x; r[8] = el.y; r[9] = el.z;
		r[10] = (double)nrefl;
	}
#pragma unroll 1
	for (int e = lk_lane(); e < c.nE; e += LK_LANES)
		r[11 + e] = w[e];
	o.nrec++;
}

/* photon.c:174-362 */
LK_DEV int lk_pc_intersect(const LkCtx &c, V3 coord, V3 direction, V3 &out)
{
	if (direction.z == 0.)
		return 0;
	V3 bd = {-1. * direction.x, -1. * direction.y, -1. * direction.z};
	norm3(bd);
	int z_id = lk_z_index(c, coord.z);
	double ext_here = (c.ext[z_id + 1] - c.ext[z_id]) / (c.z[z_id + 1] - c.z[z_id]) * (coord.z - c.z[z_id]) + c.ext[z_id];
	if (lk_within_pc(ext_here, coord.x, coord.y) == 1)
		return 0;
	int step;
	if (bd.z < 0.) {
		z_id = z_id + 1;
		step = -1;
	} else {
		step = 1;
	}
	bool found = false;
	const int in0 = lk_within_pc(ext_here, coord.x, coord.y);
	do {
		z_id += step;
		if (z_id < 0 || z_id > c.nmax)
			break;
		double px = coord.x + bd.x * (c.z[z_id] - coord.z) / bd.z;
		double py = coord.y + bd.y * (c.z[z_id] - coord.z) / bd.z;
		if (in0 != lk_within_pc(c.ext[z_id], px, py)) {
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
LK_DEV int lk_wall_exit_class(const LkCtx &c, const V3 &ec, const V3 &ed, const V3 &fallback, double &d_travel)
{
	const int nm = c.nmax;
	V3 at_exit = lk_propagate(ec, ed, c.z[nm]);
	V3 rel;
	if (lk_within_pc(c.ext[nm], at_exit.x, at_exit.y) == 0) {
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

LK_INL int lk_segment_checked(const V3 &c0, const V3 &c1, double r0, double r1, const V3 &p0, const V3 &dir, V3 &hit, V3 &sn)
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

/* ------------------------------------------------------------------ the wall march without its uneventful steps
 *
 * State of the march between two steps: the path length `dist` (a running floating-point sum, dist += cap[z_id]/10 per step) and
 * z_id, the z segment the last point fell into.  A step is UNEVENTFUL when its point lies in the same hexagonal cell, outside the
 * cell's bore and before the end of the optic; such a step changes nothing but dist and z_id.  lk_march_uneventful() recognises
 * one from the squared distance to the cell's axis and the three edge normals of the hexagon, with bounds taken 1e-9 inside the
 * zone (rounding differences to the reference's own expressions are ~1e-16).
 *
 * lk_march_jump(): with the current point inside that zone and inside z segment zs = z_id, all of the following are functions of
 * the path length s beyond the current point that are linear (cell edges, which scale with the optic's exterior; end of the
 * segment) or quadratic (distance to the axis against the bore radius, both linear in z within a segment), so the first s at which
 * the ray leaves the zone follows in closed form.  N = the number of steps that stay a whole step short of it.  Their path length is
 * reproduced exactly: while dist stays within one binade, every dist + c (c = cap[zs]/10, the same for all N steps because they
 * all land in segment zs) rounds to dist + inc with the SAME inc, a multiple of the binade's ulp -- unless c lies exactly half-way
 * between two multiples, which is excluded -- so after N steps the reference holds dist + N*inc, and that product and sum are
 * exact in double precision.  The landing point is verified with lk_march_uneventful() before the jump is accepted.  Points
 * between start and landing are inside the zone because the cell is convex and the bore test was solved for its first root.
 */
struct LkMarchSeg {
	int zi;                              /* segment these belong to, -1: none */
	double z0, e0, c0, se, sc;           /* z[zi], ext[zi], cap[zi], d ext/dz, d cap/dz */
};

LK_INL void lk_march_seg_load(const LkCtx &c, LkMarchSeg &g, int zi)
{
	if (zi == g.zi)
		return;
	g.zi = zi;
	g.z0 = c.z[zi];
	g.e0 = c.ext[zi];
	g.c0 = c.cap[zi];
	g.se = c.se[zi];
	g.sc = c.sc[zi];
}

/* the point at path length dist, z segment zi (seg loaded): strictly inside cell (Ai, Bi), outside its bore, before the exit face */
LK_INL bool lk_march_uneventful(const LkCtx &c, const LkMarchSeg &g, const V3 &ec, const V3 &ed, double Ai, double Bi, double inv_den, double dist, double p0z)
{
	const double t = p0z - g.z0;
	const double hz = (g.se * t + g.e0) * inv_den;
	const double rr = (g.sc * t + g.c0) * (1. + 1.e-9);
	const double ux = (ec.x + dist * ed.x) - Ai * hz, uy = (ec.y + dist * ed.y) - Bi * hz;
	const double h = PCB_COSPI_6 * hz * (1. - 1.e-9);
	return rr > 0. && hz > 0. && ux * ux + uy * uy > rr * rr && fabs(ux) < h && fabs(0.5 * ux + PCB_COSPI_6 * uy) < h &&
	       fabs(0.5 * ux - PCB_COSPI_6 * uy) < h && p0z <= c.z[c.nmax];
}

/* returns the number of steps jumped (0: none) and advances dist accordingly; z_id (= zs) stays what it is */
/* the jump within z segment zs and within the binade of dist; seg_end: the end of the segment or of the binade, not an event, is what
 * stopped it (the next step or two cross it) */
LK_INL unsigned long long lk_jump_in_segment(const LkCtx &c, const V3 &ec, const V3 &ed, double Ai, double Bi, double inv_den, int zs, double &dist, bool &seg_end)
{
	seg_end = false;
	/* Bounds are kept as fractions num/den (den > 0) and compared by cross-multiplication; the two quotients that are needed come from
	 * pcb_rcp_near() (a refined reciprocal, within an ulp): every bound is backed off by a whole step and the landing point is verified,
	 * so nothing below needs a correctly rounded quotient.  What has to be exact -- inc, rnd, nd -- is sums, differences and the product
	 * of an integer with a multiple of the binade's ulp. */
	if (!(ed.z > 0.) || !(dist > 0.))
		return 0;
	const double z0 = c.z[zs], z1 = c.z[zs + 1];
	const double pz = ec.z + dist * ed.z;
	if (!(z0 <= pz && pz < z1))
		return 0;
	const double cstep = c.sl[zs], inv_cstep = c.isl[zs];
	double num = z1 - pz, den = ed.z; /* smax = num / den: the end of the segment */
	bool at_seg_end = true;
	/* the increment the reference's additions realise in this binade */
	const unsigned long long db = pcb_bits(dist);
	const int ex = (int)((db >> 52) & 0x7ffULL);
	if (ex < 64 || ex > 2000)
		return 0;
	const double ulp = lk_from_bits((unsigned long long)(ex - 52) << 52), top = lk_from_bits((unsigned long long)(ex + 1) << 52);
	const double d1 = dist + cstep;
	if (!(d1 < top))
		return 0;
	const double inc = d1 - dist;            /* exact: both are multiples of ulp in one binade */
	if (!(inc >= 4. * ulp))
		return 0;
	const double rnd = cstep - inc;          /* exact (the two differ by at most ulp/2): what one addition rounds away */
	if (!(fabs(rnd) < 0.5 * ulp))
		return 0;                        /* half-way case: the additions alternate between two increments */
	/* the zone along the ray, s = path length beyond the current point */
	const double se = c.se[zs], sc = c.sc[zs];
	const double t = pz - z0;
	const double hz = (se * t + c.ext[zs]) * inv_den, hb = se * ed.z * inv_den;
	const double rr = (sc * t + c.cap[zs]) * (1. + 1.e-9), rb = sc * ed.z * (1. + 1.e-9);
	const double ux = (ec.x + dist * ed.x) - Ai * hz, uy = (ec.y + dist * ed.y) - Bi * hz;
	const double ux1 = ed.x - Ai * hb, uy1 = ed.y - Bi * hb;
	const double h = PCB_COSPI_6 * hz * (1. - 1.e-9), h1 = PCB_COSPI_6 * hb * (1. - 1.e-9);
	const double qc = ux * ux + uy * uy - rr * rr;
	if (!(rr > 0. && hz > 0. && qc > 0.))
		return 0;
	{ /* three pairs of edge planes |g + g1 s| < h + h1 s: g + g1 s reaches h + h1 s at (h - g)/(g1 - h1), -(g + g1 s) at (h + g)/(-g1 - h1) */
		const double g0 = ux, g0s = ux1;
		const double ga = 0.5 * ux + PCB_COSPI_6 * uy, gas = 0.5 * ux1 + PCB_COSPI_6 * uy1;
		const double gb = 0.5 * ux - PCB_COSPI_6 * uy, gbs = 0.5 * ux1 - PCB_COSPI_6 * uy1;
		if (!(fabs(g0) < h && fabs(ga) < h && fabs(gb) < h))
			return 0;
#define LK_EDGE(g, gs)                                              \
	{                                                           \
		const double dp = (gs)-h1, dm = -(gs)-h1;           \
		if (dp > 0. && (h - (g)) * den < num * dp) {        \
			num = h - (g);                              \
			den = dp;                                   \
			at_seg_end = false;                         \
		}                                                   \
		if (dm > 0. && (h + (g)) * den < num * dm) {        \
			num = h + (g);                              \
			den = dm;                                   \
			at_seg_end = false;                         \
		}                                                   \
	}
		LK_EDGE(g0, g0s)
		LK_EDGE(ga, gas)
		LK_EDGE(gb, gbs)
#undef LK_EDGE
	}
	{ /* bore: qa s^2 + qb s + qc > 0, qc > 0: the smallest positive root, if any */
		const double qa = ux1 * ux1 + uy1 * uy1 - rb * rb, qb = 2. * (ux * ux1 + uy * uy1 - rr * rb);
		double disc = qb * qb - 4. * qa * qc;
		if (disc < 0. && disc > -1.e-6 * (qb * qb))
			disc = 0.; /* a graze within rounding of a double root counts as one */
		if (disc >= 0. && (qb < 0. || qa < 0.)) {
			const double sq = sqrt(disc);
			double rn, rd; /* root = rn / rd, rd > 0 */
			if (qb < 0.) {
				rn = qc;
				rd = -0.5 * (qb - sq);
			} else {
				rn = 0.5 * (qb + sq);
				rd = -qa;
			}
			if (rd > 0. && rn >= 0. && rn * den < num * rd) {
				num = rn;
				den = rd;
				at_seg_end = false;
			}
		}
	}
	const double smax = num * pcb_rcp_near(den);
	if (!(smax > 0.) || !(smax < 1.e30))
		return 0;
	/* Steps that stay inside: one whole step short of an event (the rounding error of the bound is ~1e-12 of a step); when it is the
	 * end of the segment that bounds them, right up to it -- the landing point is checked against z1 with the march's own expression,
	 * and one step is given back if it fails. */
	double nf = floor(smax * inv_cstep * (1. - 1.e-7)) - (at_seg_end ? 0. : 1.);
	{ /* stay inside the binade */
		const double nb = floor((top - dist) * pcb_rcp_near(inc) * (1. - 1.e-7)) - 1.;
		if (nb < nf) { /* the end of the binade bounds the run: like the end of a segment, the next step or two cross it */
			nf = nb;
			at_seg_end = true;
		}
	}
	LkMarchSeg g;
	g.zi = -1;
	lk_march_seg_load(c, g, zs);
#pragma unroll 1
	for (int attempt = 0; attempt < 2; attempt++, nf -= 1.) {
		if (!(nf >= 1.) || nf > 4.e15)
			return 0;
		const double nd = dist + nf * inc;       /* exact */
		if (!(nd < top))
			continue;
		/* the landing point: same segment, inside the zone */
		const double lz = ec.z + nd * ed.z;
		if (!(z0 <= lz && lz < z1))
			continue;
		if (!lk_march_uneventful(c, g, ec, ed, Ai, Bi, inv_den, nd, lz))
			return 0;
		dist = nd;
		seg_end = at_seg_end;
		return (unsigned long long)nf;
	}
	return 0;
}

/* The march from (dist, z_id) on for as long as nothing happens, segment after segment: jump to the end of the segment, take the step
 * or two that cross into the next one -- each the reference's own addition dist + cap[z_id]/10, accepted only if its point is
 * uneventful -- and jump again; the end of the binade of dist (where the increment the additions realise changes) is crossed the
 * same way.  Returns the number of the reference's steps this stands for; on return (dist, z_id) is the state
 * after the last of them, and the caller's next ordinary step sees whatever stopped the run. */
LK_DEV unsigned long long lk_march_jump(const LkCtx &c, const V3 &ec, const V3 &ed, double Ai, double Bi, double inv_den, int &z_id, double &dist, unsigned long long &n_eval,
					 int max_cross, int &n_cross)
{
	n_cross = 0;
	const int nm = c.nmax;
	unsigned long long total = 0;
	int zs = z_id;
	LkMarchSeg g;
	g.zi = -1;
#pragma unroll 1
	for (;;) {
		bool seg_end = false;
		const unsigned long long nj = lk_jump_in_segment(c, ec, ed, Ai, Bi, inv_den, zs, dist, seg_end);
		total += nj;
		if (nj == 0 && total == 0)
			break; /* nothing to jump over where the march stands: the caller walks */
		if (nj != 0 && !seg_end)
			break; /* an event (or the end of the binade) is near: the caller's ordinary steps take over */
		/* cross into the next segment / the next binade: at most three single steps */
		bool crossed = false;
		const unsigned ex0 = pcb_hi_word(dist) >> 20;
#pragma unroll 1
		for (int k = 0; k < 3 && !crossed; k++) {
			const double d2 = dist + c.sl[zs];
			const double p2z = ec.z + d2 * ed.z;
			int zi = zs;
			while (zi + 1 < nm && c.z[zi + 1] <= p2z)
				zi++;
			lk_march_seg_load(c, g, zi);
			n_eval++;
			if (!lk_march_uneventful(c, g, ec, ed, Ai, Bi, inv_den, d2, p2z)) {
				n_eval--; /* not taken: the caller's ordinary step evaluates (and counts) it */
				z_id = zs;
				return total;
			}
			dist = d2;
			total++;
			crossed = zi != zs || (pcb_hi_word(d2) >> 20) != ex0;
			zs = zi;
		}
		if (!crossed)
			break;
		if (++n_cross >= max_cross)
			break; /* the caller's budget for one stretch (LK_W_BUDGET) */
	}
	z_id = zs;
	return total;
}

/* capil.c:896-1194; ed is normalised in place like the reference does with photon->exit_direction */
LK_DEV int lk_trace_wall(const LkCtx &c, const V3 &ec, V3 &ed, double &d_travel, int &r_cntr, int &q_cntr, LkWork &wk)
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
	} else if (lk_within_pc(ext_here, ec.x, ec.y) == 0) {
		return -2;
	}
	double hexz = ext_here / (2. * PCB_COSPI_6 * (n_shells + 1));
	double q_i, r_i, q_new = 0, r_new = 0;
	{ const LkCell h_ = lk_hex_cell(hexz, ec.x, ec.y); q_i = h_.q; r_i = h_.r; }
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
			/* March in steps of a tenth of the bore radius until the hex cell changes (capil.c:1018-1068).  A step is eventful when the
			 * point falls back into the photon's own bore, leaves its hexagonal cell or passes the end of the optic; everything eventful
			 * takes the reference's full step, so the outcome is the reference's step for step. */
			const bool fast_ok = c.monotonic && ed.z >= 0. && !c.no_fast_march;
			const bool jump_ok = fast_ok && !c.no_jump;
			const double inv_den = 1.0 / (2. * PCB_COSPI_6 * (n_shells + 1));
			const double Ai = (2. * q_i + r_i) * PCB_COSPI_6, Bi = r_i * (3. / 2);
			LkMarchSeg sg;
			sg.zi = -1;
			int budget = 0;
			int wait = 0; /* steps before the next attempt to jump (an attempt that found nothing to jump over is not repeated at once) */
			for (;;) {
				if (jump_ok) {
					if (wait == 0) {
						const unsigned long long tj = LK_CLOCK();
						int crossed = 0;
						const unsigned long long nj = lk_march_jump(c, ec, ed, Ai, Bi, inv_den, z_id, dist, wk.wall_evals, LK_W_BUDGET > 0 ? LK_W_BUDGET : 0x7fffffff, crossed);
						if (LK_W_BUDGET > 0 && (budget += crossed) >= LK_W_BUDGET) {
							/* a long march is taken in stretches of LK_W_BUDGET segment / binade crossings, one per tick of the block's clock, so that
							 * the other warps of the block do not wait at the barrier for the longest of their marches (LK_PHASED; a no-op otherwise) */
							budget = 0;
							LK_PHASE(LK_PH_W);
						}
						LK_TIME(t_jump, tj);
						LK_COUNT(n_jump_try);
						wk.wall_steps += nj;
						wait = (LK_W_BUDGET > 0 && crossed >= LK_W_BUDGET) ? 0 : (nj ? 2 : 4); /* after a jump that ran its course the next event or the end of the segment is one or two steps away */
					} else {
						wait--;
					}
				}
				wk.wall_steps++;
				wk.wall_evals++;
				dist += c.sl[z_id]; /* cap[z_id] / 10. */
				if (fast_ok) {
					const double p0z = ec.z + dist * ed.z;
					int zi = z_id;
					if (!(c.z[zi] <= p0z))
						zi = lk_z_index(c, p0z);
					while (zi + 1 < nm && c.z[zi + 1] <= p0z)
						zi++;
					lk_march_seg_load(c, sg, zi);
					if (lk_march_uneventful(c, sg, ec, ed, Ai, Bi, inv_den, dist, p0z)) {
						if (zi != z_id)
							wait = 0; /* a new segment: the next stretch can be jumped */
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
				{ const LkCell h_ = lk_hex_cell(hexz, p0.x, p0.y); q_new = h_.q; r_new = h_.r; }
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
				wait = 0; /* an eventful step that marches on (the first steps next to the bore wall) */
			}

			if (fabs(q_new) > n_shells || fabs(r_new) > n_shells || fabs(-1. * q_new - r_new) > n_shells || p0.z > c.z[nm]) {
				r_cntr = (int)r_new;
				q_cntr = (int)q_new;
				return lk_wall_exit_class(c, ec, ed, p0, d_travel);
			}
			/* scan the segments of the new capillary from z_id on for the point where the ray enters its bore (capil.c:1110-1133): the
			 * reference's do-while visits z_id, z_id+1, ... up to segment nmax-2 (at least one), on average more than a hundred segments
			 * per reflection; they are independent, so the lanes take LK_LANES of them at a time and the first hit in table order decides */
			iesc = 0;
			{
				const int z_first = z_id, z_last = (z_first < nm - 2) ? nm - 2 : z_first;
				const double Aq = (2. * q_new + r_new) * PCB_COSPI_6, Bq = r_new * (3. / 2);
				const double den = 2. * PCB_COSPI_6 * (n_shells + 1);
				bool found = false;
				int iesc_last = 0;
				/* Most of these segments are nowhere near the bore.  The solve's own quadratic a d^2 + b d + c in d = z - z[i] (same
				 * expressions, same bits: capil.c:105-124) tells without a root or a division more: no real root (the solve's own -2), or
				 * both ends of the segment outside the cone by a margin with the parabola's vertex outside the segment -- then no root
				 * lies within the segment and the solve cannot return 1 (capil.c:125-177); only the other segments are solved. */
				V3 dn = ed;
				norm3(dn);
				const PcbDivisor dn_z = pcb_divisor(dn.z);
				const double rx = pcb_div(dn.x, dn_z), ry = pcb_div(dn.y, dn_z);
				for (int i0 = z_first; i0 <= z_last && !found; i0 += LK_LANES) {
					wk.scan_rounds++;
					const int i = i0 + lk_lane();
					int mine = 0;
					V3 my_hit = hit, my_sn = {0, 0, 0};
					if (i <= z_last) {
						p0 = lk_propagate(ec, ed, c.z[i]);
						hexz = c.ext[i] / den;
						c0.y = Bq * hexz;
						c0.x = Aq * hexz;
						c0.z = c.z[i];
						hexz = c.ext[i + 1] / den;
						c1.y = Bq * hexz;
						c1.x = Aq * hexz;
						c1.z = c.z[i + 1];
						const double r0 = c.cap[i], r1 = c.cap[i + 1];
						bool solve = true;
						if (!c.no_fast_march && !(c0.z < 0. || c1.z < 0 || r0 < 0. || r1 < 0. || ed.z < 0 || c1.z <= c0.z)) {
							const double ax = c1.x - c0.x, ay = c1.y - c0.y, az = c1.z - c0.z;
							const PcbDivisor az_d = pcb_divisor(az);
							const double ux = rx - pcb_div(ax, az_d), uy = ry - pcb_div(ay, az_d);
							const double sl = (r1 - r0) / (c1.z - c0.z);
							const double qa = (ux * ux + uy * uy - sl * sl);
							const double qb = (2. * (p0.x - c0.x) * ux + 2. * (p0.y - c0.y) * uy - 2. * r0 * sl);
							const double qc = ((p0.x - c0.x) * (p0.x - c0.x) + (p0.y - c0.y) * (p0.y - c0.y) - r0 * r0);
							if (qb * qb - 4. * qa * qc < 0) {
								solve = false;
							} else {
								const double ql = (qa * az + qb) * az + qc, margin = 1.e-8 * r0 * r0;
								if (qc > margin && ql > margin && (qa <= 0. || qb >= 0. || 2. * qa * az + qb <= 0.))
									solve = false;
							}
						}
						mine = -3;
						if (solve) {
							wk.scan_solves++;
							mine = lk_segment_checked(c0, c1, r0, r1, p0, ed, my_hit, my_sn);
						}
					}
					const unsigned hits = lk_ballot(i <= z_last && mine == 1);
					if (hits) {
						const int src = lk_first(hits);
						found = true;
						iesc = 1;
						hit.x = lk_bcast(my_hit.x, src); hit.y = lk_bcast(my_hit.y, src); hit.z = lk_bcast(my_hit.z, src);
						z_id = i0 + src + 1;
					} else if (z_last - i0 < LK_LANES) {
						iesc_last = lk_bcast(mine, z_last - i0);
					}
				}
				if (!found) {
					iesc = iesc_last;
					z_id = z_last + 1;
				}
			}
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
 * checks (f.sn / f.cosalfa set, f.ec moved to the hit), otherwise the trace's return code (-1, -2, -3).
 * The segments are examined LK_LANES at a time, lane l taking segment i0 + l; per segment the reference's order of tests is
 *   axis outside the optic -> -1;  solve (+ obtuse-incidence veto);  hit -> in-optic test of the hit point (-3) or HIT;
 *   no hit -> photon outside the optic at the segment's start -> -3;  otherwise go on,
 * and the first segment in table order that does not "go on" decides. */
LK_DEV int lk_trace_to_hit(const LkCtx &c, LkFrame &f, LkWork &wk)
{
	const int nm = c.nmax;
	norm3(f.ed);
	const V3 from = f.ec;
	const V3 dir = f.ed;
	enum { GO_ON = 0, HIT = 1, AXIS_OUT = 2, PHOTON_OUT = 3 };
	V3 last_hit = from, sn = {0, 0, 0};
	double cosalfa = 0;
	int outcome = GO_ON;
	for (int i0 = f.ix; i0 < nm && outcome == GO_ON; i0 += LK_LANES) {
		wk.scan_rounds++;
		const int i = i0 + lk_lane();
		int mine = GO_ON;
		V3 my_hit = from, my_sn = {0, 0, 0};
		double my_cos = 0;
		if (i < nm) {
			V3 c0 = {f.A * c.hexz[i], f.B * c.hexz[i], c.z[i]}, c1 = {f.A * c.hexz[i + 1], f.B * c.hexz[i + 1], c.z[i + 1]};
			V3 p0 = lk_propagate(from, dir, c.z[i]);
			if (lk_within_pc(c.ext[i], c0.x, c0.y) == 0 || lk_within_pc(c.ext[i + 1], c1.x, c1.y) == 0) {
				mine = AXIS_OUT;
			} else {
				int iesc = lk_segment_checked(c0, c1, c.cap[i], c.cap[i + 1], p0, dir, my_hit, my_sn);
				my_cos = dot3(my_sn, dir);
				{
					double ac = pcb_acos(my_cos);
					if (ac > PCB_PI / 2. || ac < 0.)
						iesc = -5;
				}
				if (iesc == 1) {
					double ext_here = ((c.ext[i] - c.ext[i + 1]) / (c.z[i] - c.z[i + 1])) * (my_hit.z - c.z[i + 1]) + c.ext[i + 1];
					if (c.n_shells == 0.)
						mine = (sqrt(my_hit.x * my_hit.x + my_hit.y * my_hit.y) >= ext_here) ? PHOTON_OUT : HIT;
					else
						mine = (lk_within_pc(ext_here, my_hit.x, my_hit.y) == 0) ? PHOTON_OUT : HIT;
				} else if (lk_within_pc(c.ext[i], p0.x, p0.y) == 0) {
					mine = PHOTON_OUT;
				}
			}
		}
		const unsigned decided = lk_ballot(mine != GO_ON);
		if (decided) {
			const int src = lk_first(decided);
			outcome = lk_bcast(mine, src);
			if (outcome == HIT) {
				last_hit.x = lk_bcast(my_hit.x, src); last_hit.y = lk_bcast(my_hit.y, src); last_hit.z = lk_bcast(my_hit.z, src);
				sn.x = lk_bcast(my_sn.x, src); sn.y = lk_bcast(my_sn.y, src); sn.z = lk_bcast(my_sn.z, src);
				cosalfa = lk_bcast(my_cos, src);
				f.ix = i0 + src + 1;
			}
		}
	}
	if (outcome == AXIS_OUT)
		return -1;
	if (outcome == PHOTON_OUT)
		return -3;
	if (outcome != HIT)
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
	if (c.n_shells > 0 && lk_within_pc(ext_here, f.ec.x, f.ec.y) == 0)
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

/* polycap_photon_launch with leak_calc = true (photon.c:390-955), iterative over an explicit frame stack.
 * `arena` holds (out.depth+1) weight vectors of nE doubles, each held lane-wise; arena[0..nE) ends up with the primary's weights. */
LK_DEV void lk_photon(const LkCtx &c, const V3 &sc, V3 sd, const V3 &sev, double d_travel0, LkFrame *&fr, double *&arena, LkOut &out, LkResult &res,
		      LkWork &wk, const LkUnit *unit = nullptr)
{
	const int nm = c.nmax, nE = c.nE;
	const int lane = lk_lane();
	const unsigned long long t_begin = LK_CLOCK();
	int sp = 0;
	LkFrame *f = &fr[0];
	f->w = arena;
	if (!unit) /* single-step entries arrive with the photon's weights in arena[0, nE) */
#pragma unroll 1
		for (int e = lane; e < nE; e += LK_LANES)
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
		{ const LkCell h_ = lk_hex_cell(hz, sc.x, sc.y); q_i = h_.q; r_i = h_.r; }
		if (lk_within_pc(ext_here, sc.x, sc.y) == 0) {
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
				V3 ed = f->ed;
				LK_PHASE(LK_PH_W);
				const unsigned long long tw = LK_CLOCK();
				int wall_trace = lk_trace_wall(c, f->ec, ed, d_wall, r_cntr, q_cntr, wk);
				LK_TIME(t_wall, tw);
				f->ed = ed;
				if (wall_trace <= 0) {
					ret = -1;
					state = S_FINISH;
				} else {
#pragma unroll 1
					for (int e = lane; e < nE; e += LK_LANES)
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
			LK_PHASE(LK_PH_T);
			const unsigned long long ts = LK_CLOCK();
			int r = lk_trace_to_hit(c, *f, wk);
			LK_TIME(t_scan, ts);
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
			wk.refl_events++;
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
			V3 ed = f->ed;
			LK_PHASE(LK_PH_W);
			const unsigned long long tw = LK_CLOCK();
			int wall_trace = lk_trace_wall(c, f->ec, ed, d_wall, r_cntr, q_cntr, wk);
			LK_TIME(t_wall, tw);
			f->ed = ed;
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
			V3 ev = f->ev;
			refl_geometry(sn, f->ed, ev, g);
			f->ev = ev; /* renormalised in place like the reference's */
			/* the energy loop, capil.c:625-650, energies dealt over the lanes.  Pass 1 leaves the reflectivity in w_leak[e] (scratch) and finds
			 * the first energy at which it is out of range -- the reference stops there with the weights below it already updated --
			 * pass 2 applies the updates below that energy. */
			double *const fw = f->w;
			const double amu_d = -1. * d_wall;
			LK_PHASE(LK_PH_E);
			const unsigned long long te = LK_CLOCK();
			int my_bad = 0x7fffffff;
			for (int e = lane; e < nE; e += LK_LANES) {
				double rtot = fresnel_rtot(c.ec[EC_NRE * nE + e], c.ec[EC_NIM * nE + e], c.ec[EC_I2RE * nE + e], c.ec[EC_I2IM * nE + e], g.cos_t,
							   g.sin_t, g.frac_s, g.frac_p);
				if (rtot < 0. || rtot > 1.) {
					my_bad = e;
					break;
				}
				w_leak[e] = rtot;
			}
			const int first_bad = lk_min(my_bad);
			bool leak_flag = false, weight_flag = false;
			for (int e = lane; e < nE && e < first_bad; e += LK_LANES) {
				double cons1 = c.ec[EC_KR * nE + e] * alfa * c.sig_rough;
				double r_rough = exp(-1. * cons1 * cons1);
				const double rtot = w_leak[e], w = fw[e];
				const double wl = (1. - rtot * r_rough) * w * exp(amu_d * c.ec[EC_AMU * nE + e]);
				w_leak[e] = wl;
				if (wl >= 1.e-4)
					leak_flag = true;
				const double wn = w * rtot * r_rough;
				fw[e] = wn;
				if (wn >= 1.e-4)
					weight_flag = true;
			}
			LK_TIME(t_energy, te);
			if (first_bad != 0x7fffffff) {
				riesc = -1;
				state = S_AFTER_REFLECT;
				continue;
			}
			leak_flag = lk_any(leak_flag);
			weight_flag = lk_any(weight_flag);
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
					} else if (lk_within_pc(eh, lc.x, lc.y) == 0) {
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
					int inside = lk_within_pc(c.ext[nm], e3.x, e3.y);
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
				int inside = lk_within_pc(c.ext[nm], f->ec.x, f->ec.y);
				if (inside == 0)
					lk_append(c, out, 0, f->ec, f->ed, f->ev, f->irefl, f->w);
				else if (inside == 1)
					lk_append(c, out, 1, f->ec, f->ed, f->ev, f->irefl, f->w);
			}
#pragma unroll 1
			for (int e = lane; e < nE; e += LK_LANES) /* photon.c:856-885 */
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
	LK_TIME(t_total, t_begin);
	lk_sync(); /* weight vectors and records were written lane-wise; the caller reads them across lanes */
}

#endif
