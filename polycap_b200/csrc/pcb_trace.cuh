/*
 * pcb_trace.cuh -- the persistent warp-synchronous photon tracer (sm_100a, FP64).
 *
 * One lane owns one photon (geometry state in registers); a warp advances its 32 photons
 * through a small state machine and VOTES on which phase to run next, so that the expensive
 * phases (full ray/cone solve, reflection geometry + per-energy Fresnel update) execute with
 * most lanes active even though photons reflect at different times:
 *
 *   NEW    fetch a photon id (warp-aggregated atomic), sample it from the source
 *          (Philox4x32-10 stream per id) or read its replay start state, run the
 *          launch logic of photon.c:480-645                                   -> SCAN / DONE
 *   SCAN   walk segments; a segment whose both ends are clearly inside the bore
 *          cannot contain a wall hit (convexity) and is skipped     -> CAND
 *   CAND   the reference's full solve for one segment (capil.c:52-255) plus the
 *          checks of capil.c:1270-1309                                         -> SCAN / REFL / DONE
 *   REFL   capil.c:1315-1352 + capil_reflect (capil.c:596-650): travel, in-optic
 *          check, Fresnel geometry, weight update for every energy, specular
 *          direction                                                          -> SCAN / DONE
 *   DONE   classify (source.c:758-777), write the dense per-id result          -> NEW
 *
 * Finished lanes refill immediately, which is the stream compaction of live photons.
 * Profile tables live in shared memory (bulk-copied with cp.async.bulk + mbarrier when they
 * fit); per-energy optical constants are read through L1/L2; the weight vector of a photon
 * sits in an energy-major scratch slab indexed by the persistent thread, so a warp's loads
 * and stores in the energy loop are fully coalesced and stay L2 resident.
 */
#ifndef PCB_TRACE_CUH
#define PCB_TRACE_CUH

#include "pcb_device.cuh"

enum { LS_NEW = 0, LS_SCAN = 1, LS_CAND = 2, LS_REFL = 3, LS_DONE = 4, LS_IDLE = 5 };

enum { MODE_SOURCE = 0, MODE_REPLAY = 1 };

/* dense per-id outputs of one wave */
struct PcbWaveOut {
	int8_t *status;      /* [n] */
	double *fields;      /* SOURCE: [17][n] record of struct _polycap_images; REPLAY: [11][n] exit state */
	int *irefl;          /* [n] */
	double *weights;     /* [n][nE] */
};

struct PcbReplayIn {
	const double *coords, *dir, *ev; /* [n][3] each */
};

struct PcbWorkCounters { /* device-side work counters for the flop model */
	unsigned long long next_id;
	unsigned long long seg_steps;   /* segments visited (skipped or solved) */
	unsigned long long solves;      /* full quadratic solves */
	unsigned long long reflections; /* capil_reflect events */
	unsigned long long launched;    /* photons sampled/launched */
	unsigned long long dbg[8];      /* scheduler statistics of the task-queue tracer: batches, lanes, empty polls, lost races, per type */
};

struct PcbTraceParams {
	PcbOptic o;
	PcbSource s;
	PcbReplayIn rin;
	PcbWaveOut out;
	PcbWorkCounters *wc;
	double *wscratch;          /* [total_threads][nE]: one weight vector per persistent lane */
	double *hist;              /* task-queue tracer, deferred weights: [CTA][slot][T2_HK][4] reflection history (NULL otherwise) */
	int hk;                    /* ... entries after which a slot's history is evaluated (<= T2_HK) */
	unsigned long long id0;    /* global id of local index 0 */
	unsigned long long n_ids;
	int mode;
	int tables_in_smem;
	int thr_cand, thr_refl, thr_new; /* phase-vote thresholds (lanes) */
	int record_all;            /* SOURCE mode: 0 = records only for exit photons */
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* Bulk global->shared copy of one table through the TMA engine (SASS: UBLKCP), completion
 * signalled on an mbarrier.  bytes must be a multiple of 16 and both addresses 16B aligned. */
__device__ __forceinline__ void tma_bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
		     "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
		     : "memory");
}

struct Tables {
	const double *z, *cap, *ext, *hexz, *rm;
};

template <int MODE, int EXACT_W>
#ifndef PCB_LB_THREADS
#define PCB_LB_THREADS 256
#define PCB_LB_CTAS 2
#endif
__global__ void __launch_bounds__(PCB_LB_THREADS, PCB_LB_CTAS) pcb_trace_kernel(const PcbTraceParams P)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ __align__(8) uint64_t tma_bar;
	enum { STG_N = 6 };
	__shared__ double stage[(PCB_LB_THREADS / 32) * 32 * STG_N]; /* per-warp strip: cos, sin, frac_s, frac_p, alfa, (w_init, lane) of reflecting photons */

	const PcbOptic &o = P.o;
	const int nmax = o.nmax;
	const int nE = o.n_energies;
	const int lane = threadIdx.x & 31;
	const unsigned long long slot = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;

	/* ---- stage the five profile tables in shared memory ------------------------------- */
	Tables T;
	if (P.tables_in_smem) {
		const uint32_t tb = (uint32_t)o.tstride * 8u;
		double *sz = reinterpret_cast<double *>(smem_raw);
		double *scap = reinterpret_cast<double *>(smem_raw + tb);
		double *sext = reinterpret_cast<double *>(smem_raw + 2 * tb);
		double *shex = reinterpret_cast<double *>(smem_raw + 3 * tb);
		double *srm = reinterpret_cast<double *>(smem_raw + 4 * tb);
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tma_bar)));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tma_bar)), "r"(5 * tb) : "memory");
			tma_bulk_load(sz, o.z, tb, &tma_bar);
			tma_bulk_load(scap, o.cap, tb, &tma_bar);
			tma_bulk_load(sext, o.ext, tb, &tma_bar);
			tma_bulk_load(shex, o.hexz, tb, &tma_bar);
			tma_bulk_load(srm, o.rm, tb, &tma_bar);
		}
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
				     : "=r"(done)
				     : "r"(smem_u32(&tma_bar))
				     : "memory");
		}
		T.z = sz; T.cap = scap; T.ext = sext; T.hexz = shex; T.rm = srm;
	} else {
		T.z = o.z; T.cap = o.cap; T.ext = o.ext; T.hexz = o.hexz; T.rm = o.rm;
	}

	const double *ec = o.ec;
	double *wsl = P.wscratch + slot * (unsigned long long)nE; /* this lane's weight vector: element e at wsl[e] */

	/* ---- per-lane photon state ---------------------------------------------------------- */
	int st = LS_NEW;
	unsigned long long lid = 0; /* local id within the wave */
	double ex = 0, ey = 0, ez = 0, dx = 0, dy = 0, dz = 1, vx = 0, vy = 0, vz = 0;
	double lhx = 0, lhy = 0, lhz = 0;   /* photon_coord of capil_trace */
	double snx = 0, sny = 0, snz = 0, cosalfa = 0;
	double dtrav = 0, A = 0, B = 0, inv_dz = 0;
	int irefl = 0, ix = 0, iseg = 0, ntrace = 0, code = 0;
	bool slow = false, deep = false, w_init = false, in_cur = false, in_nxt = false;
	unsigned int c_steps = 0, c_solves = 0, c_refl = 0, c_launch = 0;
	bool ids_left = true;

	/* begin one capil_trace call (capil.c:1236-1246): renormalise, reset photon_coord.
	 * on_wall: the call follows a reflection, i.e. the photon sits ON the bore wall inside segment ix and heads inwards. */
	auto trace_begin = [&](bool on_wall) {
		if (ntrace > nmax) { /* launch loop ran nmax+1 times with iesc==1 (photon.c:912-919) */
			code = 1;
			st = LS_DONE;
			return;
		}
		ntrace++;
		V3 d = {dx, dy, dz};
		norm3(d);
		dx = d.x; dy = d.y; dz = d.z;
		lhx = ex; lhy = ey; lhz = ez;
		iseg = ix;
		inv_dz = pcb_rcp_near(dz);
		if (iseg < nmax) {
			/* "No wall hit possible in segment iseg" needs the photon inside the bore at both ends (convexity).
			 * Restarting on the wall the same holds from the hit point on: distance-to-axis minus radius is convex in z,
			 * zero at the hit and negative at z[iseg+1], and the root at the hit itself is rejected by the reference's
			 * 1e-5 forward margin (capil.c:134-156).  The back-extrapolated point at z[iseg], which the reference tests
			 * against the outer hexagon (capil.c:1296-1308), lies at most dz*tan(theta) beyond the wall (<= 0.01 dz, or <= 0.25 dz for capillaries deep inside): inside the
			 * hexagon for interior capillaries (m_safe carries that margin) unless the photon is steep. */
			const bool steep = (dx * dx + dy * dy) > (deep ? 0.0625 : 1.e-4) * (dz * dz); /* tan(theta) > 0.25 (deep inside) or 0.01 */
			if (on_wall && !steep) {
				in_cur = true;
			} else {
				double t = (T.z[iseg] - ez) * inv_dz;
				double hz = T.hexz[iseg];
				double ux = pcb_fma(-A, hz, pcb_fma(dx, t, ex));
				double uy = pcb_fma(-B, hz, pcb_fma(dy, t, ey));
				in_cur = pcb_fma(ux, ux, uy * uy) < T.rm[iseg];
			}
		}
		st = LS_SCAN;
	};

	for (;;) {
		/* ---------------- vote ---------------- */
		unsigned m_new = __ballot_sync(0xffffffffu, st == LS_NEW);
		unsigned m_scan = __ballot_sync(0xffffffffu, st == LS_SCAN);
		unsigned m_cand = __ballot_sync(0xffffffffu, st == LS_CAND);
		unsigned m_refl = __ballot_sync(0xffffffffu, st == LS_REFL);
		unsigned m_done = __ballot_sync(0xffffffffu, st == LS_DONE);
		if (!(m_new | m_scan | m_cand | m_refl | m_done))
			break;
		int phase;
		if (m_done && (__popc(m_done) >= P.thr_new || !(m_scan | m_cand | m_refl)))
			phase = LS_DONE;
		else if (m_new && (__popc(m_new) >= P.thr_new || !(m_scan | m_cand | m_refl)))
			phase = LS_NEW;
		else if (m_refl && (__popc(m_refl) >= P.thr_refl || !(m_scan | m_cand)))
			phase = LS_REFL;
		else if (m_cand && (__popc(m_cand) >= P.thr_cand || !m_scan))
			phase = LS_CAND;
		else
			phase = LS_SCAN;

		if (phase == LS_SCAN) {
			/* ---------------- SCAN: skip segments that cannot hold a hit ---------------- */
#pragma unroll 1
			for (int rep = 0; rep < 8; rep++) {
				if (st == LS_SCAN) {
					if (iseg >= nmax) { /* capil.c:1312-1313 -> -2 -> launch returns 1 */
						code = 1;
						st = LS_DONE;
					} else {
						double t = (T.z[iseg + 1] - ez) * inv_dz;
						double hz = T.hexz[iseg + 1];
						double ux = pcb_fma(-A, hz, pcb_fma(dx, t, ex));
						double uy = pcb_fma(-B, hz, pcb_fma(dy, t, ey));
						bool in_next = pcb_fma(ux, ux, uy * uy) < T.rm[iseg + 1];
						if (in_cur && in_next && !slow) {
							iseg++;
							c_steps++;
							in_cur = in_next;
						} else {
							st = LS_CAND; /* in_cur keeps describing z[iseg]; in_nxt is adopted when the segment misses */
							in_nxt = in_next;
						}
					}
				}
			}
		} else if (phase == LS_CAND) {
			/* ---------------- CAND: the reference's processing of one segment ---------------- */
			if (st == LS_CAND) {
				const int i = iseg;
				c_steps++;
				c_solves++;
				double z0 = T.z[i], z1 = T.z[i + 1];
				V3 c0 = {A * T.hexz[i], B * T.hexz[i], z0};
				V3 c1 = {A * T.hexz[i + 1], B * T.hexz[i + 1], z1};
				const PcbDivisor dz_d = pcb_divisor(dz);
				V3 p0 = {ex + pcb_div(dx * (z0 - ez), dz_d), ey + pcb_div(dy * (z0 - ez), dz_d), z0};
				bool bail = false;
				if (slow) { /* capil.c:1263-1266 */
					if (within_pc(T.ext[i], c0.x, c0.y) == 0 || within_pc(T.ext[i + 1], c1.x, c1.y) == 0) {
						code = -1;
						st = LS_DONE;
						bail = true;
					}
				}
				if (!bail) {
					V3 hit = {lhx, lhy, lhz}, sn;
					V3 pd = {dx, dy, dz};
					int iesc;
					/* argument checks of capil.c:65-100 that data can violate */
					if (z0 < 0. || z1 < 0 || T.cap[i] < 0. || T.cap[i + 1] < 0. || dz < 0 || z1 <= z0) {
						iesc = -1;
						sn.x = sn.y = sn.z = 0.0;
					} else {
						iesc = segment_solve(c0, c1, T.cap[i], T.cap[i + 1], p0, pd, lhz, hit, sn);
					}
					double ca = dot3(sn, pd);
					if (iesc == 1) { /* capil.c:1270-1273; sn is zero otherwise and acos(0) passes */
						if (pcb_acos_is_obtuse(ca))
							iesc = -5;
						lhx = hit.x; lhy = hit.y; lhz = hit.z; /* photon_coord keeps the last accepted root */
					}
					if (iesc == 1) {
						bool outside = false;
						if (slow) { /* capil.c:1277-1292 */
							double ext_here = ((T.ext[i] - T.ext[i + 1]) / (z0 - z1)) * (hit.z - z1) + T.ext[i + 1];
							if (o.mono)
								outside = (sqrt(hit.x * hit.x + hit.y * hit.y) >= ext_here);
							else
								outside = (within_pc(ext_here, hit.x, hit.y) == 0);
						}
						if (outside) {
							code = -1; /* trace returns -3 -> launch returns -1 */
							st = LS_DONE;
						} else {
							ix = i + 1;
							snx = sn.x; sny = sn.y; snz = sn.z;
							cosalfa = ca;
							st = LS_REFL;
						}
					} else {
						/* capil.c:1296-1308: a photon inside the bore of an interior capillary is inside the optic; one that
						 * left the bore (rejected root, capil.c:134-177) must be tested against the outer hexagon */
						bool outside = false;
						if (slow || !in_cur)
							outside = (within_pc(T.ext[i], p0.x, p0.y) == 0);
						if (outside) {
							code = -1;
							st = LS_DONE;
						} else {
							iseg = i + 1;
							in_cur = in_nxt;
							st = LS_SCAN;
						}
					}
				}
			}
			/* a solve that hits is followed at once by the reflection: no second vote between the two */
			if (__any_sync(0xffffffffu, st == LS_REFL))
				phase = LS_REFL;
		}
		if (phase == LS_REFL) {
			/* ---------------- REFL: travel, checks, Fresnel update, specular direction ---------------- */
			bool mine = (st == LS_REFL);
			ReflGeom g;
			double alfa = 0;
			bool alive = false;
			if (mine) {
				c_refl++;
				double rx = lhx - ex, ry = lhy - ey, rz = lhz - ez;
				dtrav += sqrt(rx * rx + ry * ry + rz * rz); /* capil.c:1315-1319 */
				ex = lhx; ey = lhy; ez = lhz;
				if (fabs(cosalfa) > 1.0) {
					code = -1;
					st = LS_DONE;
					mine = false;
				}
			}
			if (mine) {
				/* *ix = last i < nmax with z[i] <= hit z (capil.c:1330-1333); the hit lies in segment ix-1 */
				int found = -1;
				if (ix < nmax && T.z[ix] <= ez) {
					found = ix;
				} else {
					for (int j = ix - 1; j >= 0; j--)
						if (T.z[j] <= ez) {
							found = j;
							break;
						}
				}
				if (found >= 0) /* no z[i] <= z (NaN photon): the reference leaves *ix = i+1 */
					ix = found;
				if (slow) { /* capil.c:1334-1343 */
					double ext_here = ((T.ext[ix + 1] - T.ext[ix]) / (T.z[ix + 1] - T.z[ix])) * (lhz - T.z[ix]) + T.ext[ix];
					bool outside;
					if (o.mono)
						outside = (ex * ex + ey * ey >= ext_here);
					else
						outside = (within_pc(ext_here, ex, ey) == 0);
					if (outside) {
						code = -1;
						st = LS_DONE;
						mine = false;
					}
				}
			}
			if (mine) {
				/* capil_reflect preamble, capil.c:596-602 */
				V3 sn = {snx, sny, snz};
				norm3(sn);
				V3 d = {dx, dy, dz};
				norm3(d);
				dx = d.x; dy = d.y; dz = d.z;
				alfa = dot3(d, sn);
				if (alfa < 0.) {
					code = -1;
					st = LS_DONE;
					mine = false;
				} else {
					V3 ev = {vx, vy, vz};
					if (EXACT_W)
						refl_geometry(sn, d, ev, g);
					else
						refl_geometry_fast(sn, d, ev, g);
					vx = g.new_ev.x; vy = g.new_ev.y; vz = g.new_ev.z;
					/* snx/sny/snz keep the segment's own normal: capil_reflect renormalises a COPY (capil.c:596), and the
					 * specular direction below uses capil_trace's un-renormalised one (capil.c:1347-1349) */
				}
			}
			/* energy loop (capil.c:625-645), spread over ALL lanes of the warp: the (photon, energy) pairs of the lanes that
			 * reflect in this round are dealt out 32 at a time, so the per-energy Fresnel term -- the bulk of the work at
			 * large nE -- runs fully converged however few photons reflect together.  Geometry travels by shuffle, the
			 * weight vectors sit photon-major in the warp's scratch slab (consecutive pairs -> consecutive words). */
			{
				const unsigned go = __ballot_sync(0xffffffffu, mine);
				const int n_go = __popc(go);
				const int total = n_go * nE;
				const bool rough = (o.sig_rough != 0.0);
				const double sig_rough = o.sig_rough;
				unsigned alive_bits = 0, bad_bits = 0;
				double *wwarp = P.wscratch + (slot - lane) * (unsigned long long)nE;
				/* stage the reflecting photons' geometry, rank-ordered, in this warp's shared-memory strip */
				double *stg = stage + (threadIdx.x >> 5) * (32 * STG_N);
				if (mine) {
					double *q = stg + __popc(go & ((1u << lane) - 1u)) * STG_N;
					q[0] = g.cos_t; q[1] = g.sin_t; q[2] = g.frac_s; q[3] = g.frac_p; q[4] = alfa;
					q[5] = __hiloint2double(w_init ? 1 : 0, lane);
				}
				__syncwarp();
				int pi = lane / nE, e = lane - pi * nE; /* pair index -> (rank of photon among `go`, energy) */
				const int dpi = 32 / nE, de = 32 - dpi * nE;
				int j0 = 0;
				/* rounds of 64 pairs while they last: two independent evaluations per lane keep the FP64 pipe busy across
				 * the dependency chains of one (4 warps per scheduler cannot hide them alone) */
				if (!EXACT_W) {
#ifndef PCB_ILP
#define PCB_ILP 2
#endif
					enum { U = PCB_ILP };
					int pu[U], eu[U];
					pu[0] = pi; eu[0] = e;
#pragma unroll
					for (int u = 1; u < U; u++) {
						pu[u] = pu[u - 1] + dpi; eu[u] = eu[u - 1] + de;
						if (eu[u] >= nE) { eu[u] -= nE; pu[u]++; }
					}
					const int dpiU = (32 * U) / nE, deU = 32 * U - dpiU * nE;
					for (; j0 + 32 * U <= total; j0 += 32 * U) {
						const double *q[U];
						double *wp[U];
						double w[U], r[U];
						int sl[U];
#pragma unroll
						for (int u = 0; u < U; u++) {
							q[u] = stg + pu[u] * STG_N;
							const double tag = q[u][5];
							sl[u] = __double2loint(tag);
							wp[u] = wwarp + (unsigned long long)sl[u] * nE + eu[u];
							w[u] = __double2hiint(tag) ? *wp[u] : 1.0;
						}
#pragma unroll
						for (int u = 0; u < U; u++)
							r[u] = fresnel_rtot_fast(__ldg(ec + EC_NRE * nE + eu[u]), __ldg(ec + EC_NIM * nE + eu[u]), __ldg(ec + EC_I2RE * nE + eu[u]),
										 __ldg(ec + EC_I2IM * nE + eu[u]), q[u][0], q[u][1], q[u][2], q[u][3]);
#pragma unroll
						for (int u = 0; u < U; u++) {
							if (rough) {
								const double cr = __ldg(ec + EC_KR * nE + eu[u]) * q[u][4] * sig_rough;
								w[u] = w[u] * r[u] * exp(-1. * cr * cr);
							} else {
								w[u] = w[u] * r[u] * 1.0;
							}
							*wp[u] = w[u];
							if (r[u] < 0. || r[u] > 1.)
								bad_bits |= 1u << sl[u];
							if (w[u] >= 1.e-4)
								alive_bits |= 1u << sl[u];
							pu[u] += dpiU; eu[u] += deU;
							if (eu[u] >= nE) { eu[u] -= nE; pu[u]++; }
						}
					}
					pi = pu[0];
					e = eu[0];
				}
				/* remainder: one pair per lane and round, the weight of the next pair loaded while the current one is evaluated */
				bool valid = (j0 + lane) < total;
				int src = 0, winit = 0;
				double w_next = 1.0;
				if (valid) {
					const double tag = stg[pi * STG_N + 5];
					src = __double2loint(tag);
					winit = __double2hiint(tag);
					if (winit)
						w_next = wwarp[(unsigned long long)src * nE + e];
				}
				for (; j0 < total; j0 += 32) {
					const bool v_cur = valid;
					const int src_cur = src, e_cur = e, pi_cur = pi;
					double w = w_next;
					/* advance to the next pair and issue its load */
					pi += dpi;
					e += de;
					if (e >= nE) {
						e -= nE;
						pi++;
					}
					valid = (j0 + 32 + lane) < total;
					w_next = 1.0;
					if (valid) {
						const double tag = stg[pi * STG_N + 5];
						src = __double2loint(tag);
						winit = __double2hiint(tag);
						if (winit)
							w_next = wwarp[(unsigned long long)src * nE + e];
					}
					if (v_cur) {
						const double *q = stg + pi_cur * STG_N;
						const double n_re = __ldg(ec + EC_NRE * nE + e_cur), n_im = __ldg(ec + EC_NIM * nE + e_cur);
						const double i2_re = __ldg(ec + EC_I2RE * nE + e_cur), i2_im = __ldg(ec + EC_I2IM * nE + e_cur);
						const double rtot = EXACT_W ? fresnel_rtot(n_re, n_im, i2_re, i2_im, q[0], q[1], q[2], q[3])
									    : fresnel_rtot_fast(n_re, n_im, i2_re, i2_im, q[0], q[1], q[2], q[3]);
						if (rough) {
							const double cons1 = __ldg(ec + EC_KR * nE + e_cur) * q[4] * sig_rough;
							const double r_rough = exp(-1. * cons1 * cons1);
							w = w * rtot * r_rough;
						} else {
							w = w * rtot * 1.0;
						}
						wwarp[(unsigned long long)src_cur * nE + e_cur] = w;
						if (rtot < 0. || rtot > 1.) /* capil.c:632-636 */
							bad_bits |= 1u << src_cur;
						if (w >= 1.e-4)
							alive_bits |= 1u << src_cur;
					}
				}
				alive_bits = __reduce_or_sync(0xffffffffu, alive_bits);
				bad_bits = __reduce_or_sync(0xffffffffu, bad_bits);
				if (mine) {
					w_init = true;
					alive = (alive_bits >> lane) & 1u;
					if ((bad_bits >> lane) & 1u) {
						code = -1;
						st = LS_DONE;
						mine = false;
					}
				}
			}
			if (mine) {
				if (alive) { /* capil.c:1346-1352 */
					dx = dx - 2.0 * cosalfa * snx;
					dy = dy - 2.0 * cosalfa * sny;
					dz = dz - 2.0 * cosalfa * snz;
					V3 d = {dx, dy, dz};
					norm3(d);
					dx = d.x; dy = d.y; dz = d.z;
					irefl++;
					trace_begin(true);
				} else {
					code = 0;
					st = LS_DONE;
				}
			}
		} else if (phase == LS_DONE) {
			/* ---------------- DONE: classify and write the dense result ---------------- */
			if (st == LS_DONE) {
				const unsigned long long n = P.n_ids;
				int status = code;
				double zt = T.z[nmax];
				if (MODE == MODE_SOURCE && code == 1) { /* exit-window test, source.c:762-777 */
					double px = ex + dx * (zt - ez) / dz;
					double py = ey + dy * (zt - ez) / dz;
					int inside;
					if (o.mono)
						inside = (sqrt((px) * (px) + (py) * (py)) > T.ext[nmax]) ? 0 : 1;
					else
						inside = within_pc(T.ext[nmax], px, py);
					if (inside != 1)
						status = ST_MISSED_WINDOW;
				}
				P.out.status[lid] = (int8_t)status;
				P.out.irefl[lid] = irefl;
				double *F = P.out.fields;
				bool wr_rec = (MODE == MODE_REPLAY) || status == ST_EXIT || P.record_all;
				if (MODE == MODE_REPLAY) {
					F[0 * n + lid] = ex; F[1 * n + lid] = ey; F[2 * n + lid] = ez;
					F[3 * n + lid] = dx; F[4 * n + lid] = dy; F[5 * n + lid] = dz;
					F[6 * n + lid] = vx; F[7 * n + lid] = vy; F[8 * n + lid] = vz;
					F[9 * n + lid] = dtrav;
					F[10 * n + lid] = (double)ntrace;
				} else if (wr_rec) {
					/* record of source.c:779-798, 893-923; start values are re-drawn from the id's stream */
					StartState s0;
					source_sample(P.s, o.mono, P.id0 + lid, s0);
					/* capil_trace renormalises start_direction once per call (capil.c:1237); launch once (photon.c:492) */
					V3 sd = s0.dir;
					for (int k = 0; k < ntrace + 1; k++) {
						V3 prev = sd;
						norm3(sd);
						if (sd.x == prev.x && sd.y == prev.y && sd.z == prev.z)
							break;
					}
					F[0 * n + lid] = s0.src_x; F[1 * n + lid] = s0.src_y;
					F[2 * n + lid] = s0.coords.x; F[3 * n + lid] = s0.coords.y;
					F[4 * n + lid] = sd.x; F[5 * n + lid] = sd.y;
					double ca = dot3(s0.ev, sd);
					double al = acos(ca);
					double c_ae = 1. / sin(al), c_be = -1. * c_ae * ca;
					V3 t = {s0.ev.x * c_ae + sd.x * c_be, s0.ev.y * c_ae + sd.y * c_be, s0.ev.z * c_ae + sd.z * c_be};
					norm3(t);
					F[6 * n + lid] = round(t.x); F[7 * n + lid] = round(t.y);
					double px = ex + dx * (zt - ez) / dz;
					double py = ey + dy * (zt - ez) / dz;
					double pz = ez + dz * (zt - ez) / dz;
					F[8 * n + lid] = px; F[9 * n + lid] = py; F[10 * n + lid] = pz;
					F[11 * n + lid] = dx; F[12 * n + lid] = dy;
					V3 u = {vx * c_ae + dx * c_be, vy * c_ae + dy * c_be, vz * c_ae + dz * c_be};
					norm3(u);
					F[13 * n + lid] = round(u.x); F[14 * n + lid] = round(u.y);
					F[15 * n + lid] = (double)irefl;
					F[16 * n + lid] = dtrav + sqrt((px - ex) * (px - ex) + (py - ey) * (py - ey) + (zt - ez) * (zt - ez));
				}
				if (wr_rec || status == ST_ABSORBED) {
					double *W = P.out.weights + lid * (unsigned long long)nE;
					for (int e = 0; e < nE; e++)
						W[e] = w_init ? wsl[e] : 1.0;
				}
				st = LS_NEW;
			}
		} else if (phase == LS_NEW) {
			/* ---------------- NEW: fetch an id, build the start state, launch ---------------- */
			bool want = (st == LS_NEW);
			unsigned m = __ballot_sync(0xffffffffu, want && ids_left);
			unsigned long long base = 0;
			if (m) {
				int leader = __ffs(m) - 1;
				if (lane == leader)
					base = atomicAdd(&P.wc->next_id, (unsigned long long)__popc(m));
				base = __shfl_sync(0xffffffffu, base, leader);
			}
			if (want) {
				if (!ids_left) {
					st = LS_IDLE;
				} else {
					lid = base + __popc(m & ((1u << lane) - 1u));
					if (lid >= P.n_ids) {
						ids_left = false;
						st = LS_IDLE;
					}
				}
			}
			if (want && st == LS_NEW) {
				c_launch++;
				StartState s0;
				if (MODE == MODE_SOURCE) {
					source_sample(P.s, o.mono, P.id0 + lid, s0);
				} else {
					s0.coords.x = P.rin.coords[3 * lid]; s0.coords.y = P.rin.coords[3 * lid + 1]; s0.coords.z = P.rin.coords[3 * lid + 2];
					s0.dir.x = P.rin.dir[3 * lid]; s0.dir.y = P.rin.dir[3 * lid + 1]; s0.dir.z = P.rin.dir[3 * lid + 2];
					s0.ev.x = P.rin.ev[3 * lid]; s0.ev.y = P.rin.ev[3 * lid + 1]; s0.ev.z = P.rin.ev[3 * lid + 2];
				}
				/* photon.c:476-504 */
				irefl = 0; ntrace = 0; dtrav = 0; w_init = false; code = 0;
				norm3(s0.dir);
				ex = s0.coords.x; ey = s0.coords.y; ez = s0.coords.z;
				V3 d = s0.dir;
				norm3(d);
				dx = d.x; dy = d.y; dz = d.z;
				vx = s0.ev.x; vy = s0.ev.y; vz = s0.ev.z;
				/* photon.c:507-512 */
				int z_id = 0;
				if (ez > 0) {
					for (int i = 0; i < nmax; i++)
						if (T.z[i] <= ez)
							z_id = i;
				}
				double ext_here = ((T.ext[z_id] - T.ext[z_id + 1]) / (T.z[z_id] - T.z[z_id + 1])) * (ez - T.z[z_id]) + T.ext[z_id];
				double q_i = 0, r_i = 0;
				bool outside;
				if (o.mono) {
					outside = (sqrt((ex) * (ex) + (ey) * (ey)) > ext_here);
				} else {
					double hz = ext_here / (2. * PCB_COSPI_6 * (o.n_shells + 1));
					hex_cell(hz, ex, ey, q_i, r_i);
					outside = (within_pc(ext_here, ex, ey) == 0);
				}
				if (outside) {
					code = -2;
					st = LS_DONE;
				} else {
					/* capillary axis: cap_x[i] = A*hexz[i], cap_y[i] = B*hexz[i] (photon.c:622-627) */
					A = (2. * q_i + r_i) * PCB_COSPI_6;
					B = r_i * (3. / 2);
					ix = 0;
					if (ez > 0) { /* last i <= nmax with z[i] <= start z */
						for (int i = 0; i <= nmax; i++)
							if (T.z[i] <= ez)
								ix = i;
					} else if (T.z[0] <= ez) {
						ix = 0;
						for (int i = 1; i <= nmax && T.z[i] <= ez; i++)
							ix = i;
					}
					double rad_here, cx_here, cy_here;
					if (ez > 0) { /* photon.c:629-638 */
						double dzs = (T.z[z_id + 1] - T.z[z_id]), off = (ez - T.z[z_id]);
						double cx0 = A * T.hexz[z_id], cx1 = A * T.hexz[z_id + 1];
						double cy0 = B * T.hexz[z_id], cy1 = B * T.hexz[z_id + 1];
						rad_here = ((T.cap[z_id + 1] - T.cap[z_id]) / dzs) * off + T.cap[z_id];
						cx_here = ((cx1 - cx0) / dzs) * off + cx0;
						cy_here = ((cy1 - cy0) / dzs) * off + cy0;
					} else {
						rad_here = T.cap[0];
						cx_here = A * T.hexz[0];
						cy_here = B * T.hexz[0];
					}
					double d_axis = sqrt((ex - cx_here) * (ex - cx_here) + (ey - cy_here) * (ey - cy_here));
					if (d_axis > rad_here) {
						code = 2; /* photon.c:906 (leak transport is handled by the leak kernel) */
						st = LS_DONE;
					} else {
						/* can this capillary ever touch the outer hexagon?  ("interior capillaries") */
						double m1 = fabs(B), m2 = fabs(PCB_COSPI_6 * A + 0.5 * B), m3 = fabs(PCB_COSPI_6 * A - 0.5 * B);
						double mm = fmax(m1, fmax(m2, m3));
						slow = o.all_slow || !(mm < o.m_safe);
						deep = mm < o.m_deep;
						trace_begin(false);
					}
				}
			}
		}
	}

	/* ---- flush work counters (warp-reduced) ---- */
	for (int off = 16; off > 0; off >>= 1) {
		c_steps += __shfl_down_sync(0xffffffffu, c_steps, off);
		c_solves += __shfl_down_sync(0xffffffffu, c_solves, off);
		c_refl += __shfl_down_sync(0xffffffffu, c_refl, off);
		c_launch += __shfl_down_sync(0xffffffffu, c_launch, off);
	}
	if (lane == 0) {
		atomicAdd(&P.wc->seg_steps, (unsigned long long)c_steps);
		atomicAdd(&P.wc->solves, (unsigned long long)c_solves);
		atomicAdd(&P.wc->reflections, (unsigned long long)c_refl);
		atomicAdd(&P.wc->launched, (unsigned long long)c_launch);
	}
}

#endif
