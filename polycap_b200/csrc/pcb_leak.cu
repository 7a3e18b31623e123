/*
 * pcb_leak.cu -- kernels and host driver of the leak / recap transport (leak_calc = true), sm_100a, FP64.
 *
 * The per-photon state machine -- wall march, leak weights, secondary photons, the reference lines each part follows -- is
 * pcb_leak_core.cuh.  Here: pcb_leak_kernel (persistent warps, ONE WARP per primary photon pulling ids from a global counter;
 * a warp's frame stack, weight arena and record strip live in global memory), the single-item kernels behind the reference's
 * TEST_BUILD entry points, and the host side that sizes workspaces from a memory budget and returns the records in id order.
 * The voted, cheap-weights tracer of pcb_trace2.cuh serves leak_calc = false.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef LK_OUTLINE /* tuning builds: one copy of the solve, the reflection geometry and the normalisation (see LK_DEV in pcb_leak_core.cuh) */
#define PCB_OUTLINE_HOT 1
#define PCB_NORM3_CALL 1
#endif
#include "pcb_device.cuh"
#include "pcb_gpu.h"
#include "pcb_leak.h"
#if !defined(LK_UNPHASED) && !defined(LK_PHASED)
#define LK_PHASED 1 /* the warps of a block take the phases of a reflection event in step (pcb_leak_core.cuh): 1.3x on the B200 */
#endif
#include "pcb_leak_core.cuh"

#ifndef LK_THREADS
#define LK_THREADS 128
#endif
#define LK_WARPS (LK_THREADS / 32)
#ifndef LK_BLOCKS_PER_SM
#define LK_BLOCKS_PER_SM 1 /* registers as the code needs them: capping them at 128 (4 blocks, 16 warps per SM) measured 2.6x slower per warp */
#endif

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
	double *arena;            /* [warps][(depth+1)*nE] */
	LkFrame *frames;          /* [warps][depth] */
	double *rec;              /* [warps][rec_cap*(11+nE)] scratch strips */
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

__global__ void __launch_bounds__(LK_THREADS, LK_BLOCKS_PER_SM) pcb_leak_kernel(const LkParams P)
{
	const LkCtx &c = P.c;
	const int nE = c.nE;
	const int lane = lk_lane();
	const unsigned long long wid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	double *const arena0 = P.arena + wid * (unsigned long long)(P.depth + 1) * nE;
	LkFrame *const frames0 = P.frames + wid * (unsigned long long)P.depth;
	double *const rec0 = P.rec + wid * (unsigned long long)P.rec_cap * (11 + nE);
	LkOut out;
	out.big_frames = P.big_frames; out.big_arena = P.big_arena; out.big_rec = P.big_rec; out.big_next = P.big_next; out.big_n = P.big_n;
	LkWork total{};
#ifdef LK_PHASED
	int phase_clock = 0;
#endif
	for (;;) {
		unsigned long long i = lk_fetch_add(P.next, 1ULL);
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
		out.rec = rec0; /* every photon starts in the warp's own (small) workspace */
		out.cap = P.rec_cap;
		out.depth = P.depth;
		double *arena = arena0;
		LkFrame *frames = frames0;
		LkResult r;
		LkWork wk{};
#ifdef LK_PHASED
		wk.phase.clock = phase_clock;
#endif
		lk_sync(); /* the previous photon's copy-out has read this workspace across lanes */
		lk_photon(c, s0.coords, s0.dir, s0.ev, 0.0, frames, arena, out, r, wk);
#ifdef LK_PHASED
		phase_clock = wk.phase.clock;
#endif
		if (out.err) {
			if (P.retry) { /* did not fit the small per-warp capacities: the second launch redoes this id with the full ones */
				const unsigned long long slot = lk_fetch_add(P.n_retry, 1ULL);
				if (lane == 0) {
					P.retry[slot] = i;
					P.leak_n[i] = 0;
					P.leak_off[i] = 0;
				}
				continue;
			}
			if (lane == 0)
				atomicOr(P.err, out.err);
		}
		total.refl_events += wk.refl_events;
		total.wall_steps += wk.wall_steps;
		total.wall_evals += wk.wall_evals;
		total.scan_rounds += wk.scan_rounds;
		total.scan_solves += wk.scan_solves;
#ifdef LK_PROFILE
		total.t_total += wk.t_total; total.t_wall += wk.t_wall; total.t_jump += wk.t_jump; total.t_scan += wk.t_scan; total.t_energy += wk.t_energy;
		total.n_jump_try += wk.n_jump_try;
#endif
		int status = r.ret;
		const double zt = c.z[c.nmax];
		if (P.mode == 0 && r.ret == 1) { /* exit-window test, source.c:762-777 */
			double px = r.ec.x + r.ed.x * (zt - r.ec.z) / r.ed.z;
			double py = r.ec.y + r.ed.y * (zt - r.ec.z) / r.ed.z;
			int inside = c.mono ? ((sqrt(px * px + py * py) > c.ext[c.nmax]) ? 0 : 1) : lk_within_pc(c.ext[c.nmax], px, py);
			if (inside != 1)
				status = ST_MISSED_WINDOW;
		}
		double *F = P.fields;
		const unsigned long long n = P.n;
		if (lane == 0) {
			P.status[i] = (int8_t)status;
			P.irefl[i] = (int)r.irefl;
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
		}
		for (int e = lane; e < nE; e += 32)
			P.weights[i * nE + e] = arena[e];
		/* leak records of this id -> pool (the reference keeps the leaks of photons with codes 1, 0, 2; source.c:799-879) */
		int keep = out.nrec;
		if (P.mode == 0 && (status == ST_OUTSIDE || status == ST_ERROR))
			keep = 0;
		if (keep > 0) {
			const unsigned long long off = lk_fetch_add(P.leak_pool_used, (unsigned long long)keep);
			if (off + keep <= P.leak_pool_cap) {
				const size_t stride = 11 + nE;
				for (size_t k = lane; k < (size_t)keep * stride; k += 32)
					P.leak_pool[off * stride + k] = out.rec[k];
			} else if (lane == 0) {
				atomicOr(P.err, LK_ERR_POOL);
			}
			if (lane == 0)
				P.leak_off[i] = off;
		} else if (lane == 0) {
			P.leak_off[i] = 0;
		}
		if (lane == 0)
			P.leak_n[i] = keep;
	}
#ifdef LK_PHASED
	while (__syncthreads_or(0)) { /* out of photons: keep the block's clock ticking until its other warps are done too */
	}
#endif
	if (lane == 0) {
		atomicAdd(P.refl_events, total.refl_events);
		atomicAdd(P.refl_events + 1, total.wall_steps);
		atomicAdd(P.refl_events + 2, total.wall_evals);
		atomicAdd(P.refl_events + 3, total.scan_rounds);
		atomicAdd(P.refl_events + 10, total.scan_solves);
#ifdef LK_PROFILE
		atomicAdd(P.refl_events + 4, total.t_total);
		atomicAdd(P.refl_events + 5, total.t_wall);
		atomicAdd(P.refl_events + 6, total.t_jump);
		atomicAdd(P.refl_events + 7, total.t_scan);
		atomicAdd(P.refl_events + 8, total.t_energy);
		atomicAdd(P.refl_events + 9, total.n_jump_try);
#endif
	}
}

/* single-item kernels for the reference's private test entry points (one warp) */
__global__ void lk_unit_trace_wall_kernel(LkCtx c, const double *in, double *outd, int *outi)
{
	V3 ec = {in[0], in[1], in[2]}, ed = {in[3], in[4], in[5]};
	double d_travel;
	int r_cntr, q_cntr;
	LkWork wk{};
	const int rc = lk_trace_wall(c, ec, ed, d_travel, r_cntr, q_cntr, wk);
	if (lk_lane() == 0) {
		outi[0] = rc;
		outi[1] = r_cntr;
		outi[2] = q_cntr;
		outd[0] = d_travel;
	}
}

/* single polycap_capil_reflect / polycap_capil_trace with leak_calc = true on one photon state:
 * in[16] = coords(3) direction(3) elecv(3) normal(3) i_refl d_travel A B; ix; w[nE] in/out; outs[13] = ret coords(3) direction(3) elecv(3) i_refl d_travel ix */
__global__ void lk_unit_step_kernel(LkCtx c, int mode, const double *in, int ix, double *w, LkFrame *frames, double *arena, double *rec, int rec_cap, int depth,
				    double *outs, int *nrec_out, int *err_out)
{
	const int lane = lk_lane();
	LkOut out;
	out.rec = rec; out.nrec = 0; out.err = 0; out.cap = rec_cap; out.depth = depth;
	out.big_frames = nullptr; out.big_arena = nullptr; out.big_rec = nullptr; out.big_next = nullptr; out.big_n = 0;
	for (int e = lane; e < c.nE; e += 32) /* lane-wise, like every weight vector of the state machine */
		arena[e] = w[e];
	LkUnit u;
	u.mode = mode;
	u.normal.x = in[9]; u.normal.y = in[10]; u.normal.z = in[11];
	u.irefl = (long long)in[12];
	u.A = in[14]; u.B = in[15];
	u.ix = ix;
	V3 sc = {in[0], in[1], in[2]}, sd = {in[3], in[4], in[5]}, sev = {in[6], in[7], in[8]};
	LkResult r;
	LkWork wk{};
	LkFrame *fr = frames;
	double *ar = arena;
	lk_photon(c, sc, sd, sev, in[13], fr, ar, out, r, wk, &u);
	for (int e = lane; e < c.nE; e += 32)
		w[e] = ar[e];
	if (lane == 0) {
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
}

/* leak records of a wave from pool (arrival) order into id order: record k of id i goes to ordered[dst[i] + k] */
__global__ void lk_gather_kernel(const double *pool, const unsigned long long *src, const unsigned long long *dst, const int *leak_n, unsigned long long stride,
				 unsigned long long n, double *ordered)
{
	for (unsigned long long i = blockIdx.x; i < n; i += gridDim.x) {
		const int m = leak_n[i];
		if (m <= 0)
			continue;
		const double *from = pool + src[i] * stride;
		double *to = ordered + dst[i] * stride;
		const unsigned long long words = (unsigned long long)m * stride;
		for (unsigned long long k = threadIdx.x; k < words; k += blockDim.x)
			to[k] = from[k];
	}
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
			    int *leak_n, double **leak_rec, int64_t *n_leak_rec, int64_t *work /* [4]: capil_reflect events, wall-march steps, steps evaluated, scan rounds */)
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
	P.c.no_jump = getenv("POLYCAP_B200_LEAK_NO_JUMP") ? 1 : 0;
	P.c.n_shells = o->n_shells;
	P.c.sig_rough = o->sig_rough;
	P.c.z = o->d_z; P.c.cap = o->d_cap; P.c.ext = o->d_ext; P.c.hexz = o->d_hexz; P.c.ec = o->d_ec;
	P.c.se = o->d_se; P.c.sc = o->d_sc; P.c.sl = o->d_sl; P.c.isl = o->d_isl;
	if (src) {
		P.s.d_source = src->d_source; P.s.src_x = src->src_x; P.s.src_y = src->src_y; P.s.src_sigx = src->src_sigx; P.s.src_sigy = src->src_sigy;
		P.s.src_shiftx = src->src_shiftx; P.s.src_shifty = src->src_shifty; P.s.hor_pol = src->hor_pol;
		P.s.cap0 = o->cap0; P.s.ext0 = o->ext0; P.s.seed = seed;
	}
	P.mode = mode;
	P.id0 = (unsigned long long)id0;
	P.n = (unsigned long long)n;
	/* Per-warp workspace: frame stack, weight arena, record strip; the warp count follows from a memory budget.  Few photons
	 * nest deeper than a few dozen secondaries, but one that does needs LK_DEPTH frames and LK_REC_CAP records, i.e. megabytes per
	 * warp at long energy grids -- so every warp starts with small capacities, a photon that outgrows them moves to a full-size
	 * workspace from a pool, and the ids that found the pool empty are redone by a second launch with full capacities (few warps). */
	const size_t stride = 11 + (size_t)nE;
	auto per_warp_bytes = [&](int depth, int rcap) { return (size_t)depth * sizeof(LkFrame) + (size_t)(depth + 1) * nE * 8 + (size_t)rcap * stride * 8; };
	int depth1 = getenv("POLYCAP_B200_LEAK_DEPTH1") ? atoi(getenv("POLYCAP_B200_LEAK_DEPTH1")) : 96;
	int rcap1 = 4 * depth1;
	size_t free_b = 0, total_b = 0;
	LCK(cudaMemGetInfo(&free_b, &total_b));
	size_t budget = free_b / 3;
	if (budget > ((size_t)48 << 30))
		budget = (size_t)48 << 30;
	int per_sm = 0; /* resident blocks per SM: the persistent warps of a launch are exactly the ones the device holds at once */
	LCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcb_leak_kernel, LK_THREADS, 0));
	if (per_sm < 1)
		per_sm = 1;
	if (const char *ps = getenv("POLYCAP_B200_LEAK_BLOCKS_PER_SM")) /* tuning aid: fewer resident blocks than fit */
		if (atoi(ps) >= 1 && atoi(ps) < per_sm)
			per_sm = atoi(ps);
	auto blocks_for = [&](int64_t items, int depth, int rcap) {
		int blocks = (int)((items + LK_WARPS - 1) / LK_WARPS);
		int max_blocks = (int)(budget / (per_warp_bytes(depth, rcap) * LK_WARPS));
		if (max_blocks > o->sm_count * per_sm)
			max_blocks = o->sm_count * per_sm;
		if (max_blocks < 1)
			max_blocks = 1;
		return blocks > max_blocks ? max_blocks : blocks;
	};
	/* with one workspace per WARP (a few thousand per device) full capacities for everyone usually fit the budget -- 13.6 GB at 400
	 * energies -- and then no photon ever moves or is redone; the small-workspace scheme above remains for very long grids */
	bool all_full = false;
	if (!getenv("POLYCAP_B200_LEAK_DEPTH1")) {
		int want = (int)((n + LK_WARPS - 1) / LK_WARPS);
		if (want > o->sm_count * per_sm)
			want = o->sm_count * per_sm;
		if ((size_t)want * LK_WARPS * per_warp_bytes(LK_DEPTH, LK_REC_CAP) <= budget) {
			all_full = true;
			depth1 = LK_DEPTH;
			rcap1 = LK_REC_CAP;
		}
	}
	int blocks = blocks_for(n, depth1, rcap1);
	size_t warps = (size_t)blocks * LK_WARPS;
	double *d_rin = nullptr;
	P.depth = depth1;
	P.rec_cap = rcap1;
	LCK(M.alloc(&P.next, 8));
	LCK(M.alloc(&P.retry, (size_t)n * 8));
	LCK(M.alloc(&P.n_retry, 8));
	LCK(cudaMemset(P.n_retry, 0, 8));
	LCK(M.alloc(&P.arena, warps * (size_t)(depth1 + 1) * nE * 8));
	LCK(M.alloc(&P.frames, warps * (size_t)depth1 * sizeof(LkFrame)));
	LCK(M.alloc(&P.rec, warps * (size_t)rcap1 * stride * 8));
	{ /* full-size workspaces for the deep photons of this launch: a quarter of what is left, at most 16 GB, at most one per id */
		size_t fb = 0, tb = 0;
		LCK(cudaMemGetInfo(&fb, &tb));
		size_t big_budget = fb / 4;
		if (big_budget > ((size_t)16 << 30))
			big_budget = (size_t)16 << 30;
		unsigned long long nb = big_budget / per_warp_bytes(LK_DEPTH, LK_REC_CAP);
		if (nb > (unsigned long long)n)
			nb = (unsigned long long)n;
		const char *bm = getenv("POLYCAP_B200_LEAK_BIG_MAX"); /* profiling aid: fewer full-size workspaces = less memory for ncu to save and restore */
		const unsigned long long nb_max = bm ? (unsigned long long)atoll(bm) : 4096ULL;
		if (nb > nb_max)
			nb = nb_max;
		if (all_full)
			nb = 0;
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
	/* record pool of the wave: room for 96 leaks per id on average (the 1-40 keV ellipsoidal run leaves 39), within a third of the
	 * free memory; a wave that overflows it is reported with return code -2 and retried by the caller with fewer ids (the reference
	 * simply reallocs per photon) */
	P.leak_pool_cap = (unsigned long long)n * 96 + 65536;
	{
		const unsigned long long fit = (unsigned long long)(free_b / 3) / (stride * 8);
		if (P.leak_pool_cap > fit)
			P.leak_pool_cap = fit > 4096 ? fit : 4096;
	}
	LCK(M.alloc(&P.leak_pool, (size_t)P.leak_pool_cap * stride * 8));
	LCK(M.alloc(&P.leak_pool_used, 8));
	LCK(M.alloc(&P.refl_events, 128));
	LCK(M.alloc(&P.err, 4));
	LCK(cudaMemset(P.next, 0, 8));
	LCK(cudaMemset(P.leak_pool_used, 0, 8));
	LCK(cudaMemset(P.refl_events, 0, 128));
	LCK(cudaMemset(P.err, 0, 4));
	if (mode == 1) {
		LCK(M.alloc(&d_rin, (size_t)n * 9 * 8));
		LCK(cudaMemcpy(d_rin, coords, (size_t)n * 24, cudaMemcpyHostToDevice));
		LCK(cudaMemcpy(d_rin + 3 * n, dir, (size_t)n * 24, cudaMemcpyHostToDevice));
		LCK(cudaMemcpy(d_rin + 6 * n, ev, (size_t)n * 24, cudaMemcpyHostToDevice));
		P.rin = d_rin;
	}
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	float kernel_ms = 0, kernel2_ms = 0;
	LCK(cudaEventCreate(&ev0));
	LCK(cudaEventCreate(&ev1));
	LCK(cudaEventRecord(ev0));
	pcb_leak_kernel<<<blocks, LK_THREADS>>>(P);
	LCK(cudaGetLastError());
	LCK(cudaEventRecord(ev1));
	LCK(cudaDeviceSynchronize());
	cudaEventElapsedTime(&kernel_ms, ev0, ev1);
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
		const size_t warps2 = (size_t)blocks2 * LK_WARPS;
		LCK(M2.alloc(&Q.arena, warps2 * (size_t)(LK_DEPTH + 1) * nE * 8));
		LCK(M2.alloc(&Q.frames, warps2 * (size_t)LK_DEPTH * sizeof(LkFrame)));
		LCK(M2.alloc(&Q.rec, warps2 * (size_t)LK_REC_CAP * stride * 8));
		LCK(cudaMemset(Q.next, 0, 8));
		LCK(cudaEventRecord(ev0));
		pcb_leak_kernel<<<blocks2, LK_THREADS>>>(Q);
		LCK(cudaGetLastError());
		LCK(cudaEventRecord(ev1));
		LCK(cudaDeviceSynchronize());
		cudaEventElapsedTime(&kernel2_ms, ev0, ev1);
	}
	if (getenv("POLYCAP_B200_PHASES"))
	{
		unsigned long long moved = 0;
		cudaMemcpy(&moved, P.big_next, 8, cudaMemcpyDeviceToHost);
		fprintf(stderr, "polycap-b200 phases: leak wave of %lld ids: %d blocks x %d warps with %d frames each, %llu photons moved to a full-size workspace (%llu available), "
				"%llu ids redone in a second launch\n", (long long)n, blocks, LK_WARPS, depth1, moved, P.big_n, n_retry);
	}
	cudaEventDestroy(ev0);
	cudaEventDestroy(ev1);
	int err = 0;
	unsigned long long used = 0, refl[16] = {0};
	LCK(cudaMemcpy(&err, P.err, 4, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(&used, P.leak_pool_used, 8, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(refl, P.refl_events, 128, cudaMemcpyDeviceToHost));
	if (getenv("POLYCAP_B200_PHASES")) {
		fprintf(stderr, "polycap-b200 phases: leak kernel %.1f ms (+ %.1f ms second launch): %llu reflection events, %llu march steps (%llu evaluated), %llu scan rounds, %llu solves by lane 0 in the scans for the next bore\n",
			kernel_ms, kernel2_ms, refl[0], refl[1], refl[2], refl[3], refl[10]);
#ifdef LK_PROFILE
		const double tt = (double)refl[4] > 0 ? (double)refl[4] : 1.0;
		fprintf(stderr, "polycap-b200 phases: leak clocks per reflection event (warp-time): total %.0f = wall march %.0f (of which jumps %.0f, %.2f attempts) + segment scan %.0f + energy loop %.0f + rest %.0f\n",
			tt / (double)refl[0], (double)refl[5] / (double)refl[0], (double)refl[6] / (double)refl[0], (double)refl[9] / (double)refl[0], (double)refl[7] / (double)refl[0],
			(double)refl[8] / (double)refl[0], (tt - (double)refl[5] - (double)refl[7] - (double)refl[8]) / (double)refl[0]);
#endif
	}
	if (err & (LK_ERR_DEPTH | LK_ERR_RECORDS)) {
		g_lk_err = (err & LK_ERR_DEPTH) ? "leak transport: more than 1400 nested secondary photons" : "leak transport: more than 2048 leak records from one photon";
		return -1;
	}
	if (err & LK_ERR_POOL) {
		g_lk_err = "leak transport: the wave's leak-record pool is full (retry with fewer photon ids per wave)";
		return -2;
	}
	if (work) {
		for (int k = 0; k < 4; k++)
			work[k] = (int64_t)refl[k];
	}
	LCK(cudaMemcpy(status, P.status, (size_t)n, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(irefl, P.irefl, (size_t)n * 4, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(fields, P.fields, (size_t)n * nfields * 8, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(weights, P.weights, (size_t)n * nE * 8, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(leak_n, P.leak_n, (size_t)n * 4, cudaMemcpyDeviceToHost));
	/* records: pool order is arrival order; they are put in id order on the device (lk_gather_kernel) and come back in one copy */
	std::vector<unsigned long long> dst((size_t)n);
	size_t w = 0;
	for (int64_t i = 0; i < n; i++) {
		dst[(size_t)i] = (unsigned long long)w;
		w += (size_t)(leak_n[i] > 0 ? leak_n[i] : 0);
	}
	double *recs = (double *)malloc((w ? w : 1) * stride * 8);
	if (!recs) {
		g_lk_err = "leak transport: out of host memory for the leak records";
		return -1;
	}
	if (w) {
		unsigned long long *d_dst = nullptr;
		double *d_ordered = nullptr;
		LCK(M.alloc(&d_dst, (size_t)n * 8));
		LCK(M.alloc(&d_ordered, w * stride * 8));
		LCK(cudaMemcpy(d_dst, dst.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
		lk_gather_kernel<<<o->sm_count * 8, 256>>>(P.leak_pool, P.leak_off, d_dst, P.leak_n, (unsigned long long)stride, (unsigned long long)n, d_ordered);
		LCK(cudaGetLastError());
		LCK(cudaMemcpy(recs, d_ordered, w * stride * 8, cudaMemcpyDeviceToHost));
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
	c.no_jump = getenv("POLYCAP_B200_LEAK_NO_JUMP") ? 1 : 0;
	c.z = o->d_z; c.cap = o->d_cap; c.ext = o->d_ext; c.hexz = o->d_hexz; c.ec = o->d_ec;
	c.se = o->d_se; c.sc = o->d_sc; c.sl = o->d_sl; c.isl = o->d_isl;
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
	lk_unit_step_kernel<<<1, 32>>>(c, mode, d_in, ix, d_w, d_frames, d_arena, d_rec, LK_REC_CAP, LK_DEPTH, d_outs, d_i, d_i + 1);
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
	c.no_jump = getenv("POLYCAP_B200_LEAK_NO_JUMP") ? 1 : 0;
	c.z = o->d_z; c.cap = o->d_cap; c.ext = o->d_ext; c.hexz = o->d_hexz; c.ec = o->d_ec;
	c.se = o->d_se; c.sc = o->d_sc; c.sl = o->d_sl; c.isl = o->d_isl;
	DevMem M;
	double *d_in, *d_out;
	int *d_i;
	LCK(M.alloc(&d_in, 48));
	LCK(M.alloc(&d_out, 8));
	LCK(M.alloc(&d_i, 12));
	LCK(cudaMemcpy(d_in, in, 48, cudaMemcpyHostToDevice));
	lk_unit_trace_wall_kernel<<<1, 32>>>(c, d_in, d_out, d_i);
	LCK(cudaGetLastError());
	int hi[3];
	LCK(cudaMemcpy(hi, d_i, 12, cudaMemcpyDeviceToHost));
	LCK(cudaMemcpy(d_travel, d_out, 8, cudaMemcpyDeviceToHost));
	*r_cntr = hi[1];
	*q_cntr = hi[2];
	return hi[0];
}
