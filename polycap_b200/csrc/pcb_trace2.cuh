/*
 * pcb_trace2.cuh -- the CTA-level task-queue tracer (sm_100a, FP64): same arithmetic, different scheduling.
 *
 * pcb_trace.cuh keeps a photon in the registers of one lane and lets each warp vote on the phase to run, which
 * leaves ~40 % of the lanes idle in the expensive phases (a photon spends most of its life waiting for or doing the
 * wall-hit work, but the 32 photons of one warp are rarely there together).  Here a CTA owns a pool of T2_NP photons
 * whose state lives in SHARED MEMORY, and one ring buffer per task type:
 *
 *   NEW   slot is free: fetch an id, sample / read the start state, launch logic      -> SCAN / DONE / (idle)
 *   SCAN  walk segments with the convexity skip test until a candidate segment         -> HIT / DONE / SCAN
 *   HIT   the reference's solve for the candidate; on a hit the whole reflection: travel /
 *         in-optic checks, geometry, the per-energy Fresnel update with the batch's
 *         (photon, energy) pairs dealt over all lanes, specular direction, next trace
 *         call; then the scan to the next candidate while the photon is in registers   -> HIT / SCAN / DONE
 *   DONE  classify, write the dense per-id outputs                                     -> NEW
 *
 * Every warp loops: pick a queue (one with >= 32 entries if there is one), claim up to 32 slots, run that task with one
 * photon per lane, push each photon to its next queue.  Any 32 photons of the pool that wait for the same kind of work
 * form a full warp, so the lanes stay busy whatever the photons' histories; no CTA barrier is needed after start-up.
 * The device functions (segment_solve, refl_geometry*, fresnel_rtot*, source_sample) are the ones pcb_trace.cuh uses:
 * results are identical photon by photon (tests/test_gpu_parity.py runs both kernels).
 */
#ifndef PCB_TRACE2_CUH
#define PCB_TRACE2_CUH

#include "pcb_trace.cuh"

#ifndef T2_THREADS
#define T2_THREADS 640
#endif
#ifndef T2_NP
#define T2_NP 1152 /* photon slots per CTA */
#endif
#define T2_QCAP 2048 /* ring capacity, power of two > T2_NP */
#ifndef T2_LOCKSTEP
#define T2_LOCKSTEP 0 /* 1: barrier-separated rounds with a deterministic batch assignment; 0: every warp claims work on its own */
#endif
#ifndef T2_STICKY
#define T2_STICKY 0
#endif
#ifndef T2_CONSENSUS
#define T2_CONSENSUS 1 /* the warps of a CTA prefer the same long task (see the claim code); 0: fullest queue wins every time */
#endif
#ifndef T2_PREFETCH
#define T2_PREFETCH 64 /* energies ahead to prefetch in the dealt energy loop (0: off) */
#endif
#ifndef T2_SPREAD_ILP
#define T2_SPREAD_ILP 2 /* (photon, energy) pairs in flight per lane in the dealt energy loop */
#endif
#ifndef T2_HK
#define T2_HK 64 /* deferred weights: reflection-history entries per slot (the history is flushed at P.hk <= T2_HK entries) */
#endif
#ifndef T2_EU
#define T2_EU 2 /* deferred weights: energies per lane and pass in the energy task (independent FP64 chains per lane) */
#endif
#ifndef T2_EK
#define T2_EK 1 /* deferred weights: reflections per trip of the energy task's inner loop (independent Fresnel terms to overlap) */
#endif
#ifndef T2_SCAN_AFTER
#define T2_SCAN_AFTER 24 /* segment steps a photon takes inside the task that launched or reflected it before it goes to the scan queue */
#endif
#define T2_PRAGMA_(x) _Pragma(#x)
#define T2_PRAGMA(x) T2_PRAGMA_(x)
#ifndef T2_DOUBLE_MIN
#define T2_DOUBLE_MIN 0 /* tuning variant, not used: a warp claims two batches (64 slots) of a long task at once when the queue holds at least this many entries and runs them back to back -- half the claims, but the pool's photons wait longer in a warp's hands: measured -14 % (160) ... -24 % (64) on the conical workload */
#endif
#ifndef T2_CAS_TRIES
#define T2_CAS_TRIES 3 /* a claim that loses the race for the head of its queue re-reads the head and tries again before it starts over with the choice of a task: lost races per batch 0.87 -> 0.31, +2.6 % on the conical workload */
#endif
#ifndef T2_BACKOFF_MAX
#define T2_BACKOFF_MAX 4096u /* longest sleep of a warp that finds every queue empty [ns] */
#endif
#define T2_EMPTY 0xFFFFu

enum { TQ_NEW = 0, TQ_SCAN, TQ_HIT, TQ_REFL, TQ_DONE, TQ_ENERGY, TQ_N };

/* flag bits of a photon slot.  F_WINIT: the slot's weight vector in the scratch is valid.  Deferred weights only: F_FAILHIT = the
 * photon's last wall hit was rejected after it had been accepted as a reflection event (counted when the history is evaluated),
 * F_WOUT = the final weight vector has been written to the wave output */
enum { F_SLOW = 1, F_DEEP = 2, F_WINIT = 4, F_INCUR = 8, F_INNXT = 16, F_FAILHIT = 32, F_WOUT = 64 };

struct T2Dbl { /* the 16 double fields of the pool: ex ey ez dx dy dz vx vy vz g0 g1 lhz dtrav A B aux (g0, g1, lhz, aux double as the staging of the reflection geometry), each [T2_NP] */
	double *base;
	__device__ __forceinline__ double *operator[](int k) const { return base + k * T2_NP; }
};
struct T2Shared { /* pointers into dynamic shared memory */
	T2Dbl d;
	unsigned *lid;
	int *irefl, *ix, *iseg, *ntrace;
	unsigned char *flags;
	signed char *code;
	unsigned char *nh;   /* deferred weights: buffered reflections of the slot */
	signed char *resume; /* deferred weights: queue the photon continues in after its history has been evaluated */
	volatile unsigned short *qbase; /* TQ_N rings of T2_QCAP entries, back to back */
	__device__ __forceinline__ volatile unsigned short *q(int k) const { return qbase + k * T2_QCAP; }
};

enum { D_EX = 0, D_EY, D_EZ, D_DX, D_DY, D_DZ, D_VX, D_VY, D_VZ, D_G0, D_G1, D_LHZ, D_DTRAV, D_A, D_B, D_AUX, D_N };

__device__ __forceinline__ void t2_push(volatile unsigned short *buf, unsigned *tail, unsigned slot)
{
	__threadfence_block(); /* the slot's state is visible before its id is */
	const unsigned t = atomicAdd(tail, 1u) & (T2_QCAP - 1);
	while (buf[t] != T2_EMPTY) {
	}
	buf[t] = (unsigned short)slot;
}


/* Energy task of the deferred-weights tracer, one pass: the lane's U energies through all nh buffered reflections of one photon
 * (history H, 4 doubles per reflection, one entry per lane, broadcast by shuffle).  Updates w[], returns in `lead` the number of
 * leading reflections after which one of the lane's weights was still >= 1e-4 and in `first_bad` the first reflection with a
 * reflectivity outside [0, 1].  EXACT = false screens both tests on the high words of r and w and leaves u == 0 to chance: it
 * returns true (for the whole warp) when any lane met a value the screens cannot decide, and the caller repeats the pass with
 * EXACT = true.  Cost model of the inner loop (ncu, B200): an FP64 instruction holds the issue port for two cycles, any other
 * instruction for one, so the loop is written to minimise 2 F + O. */
template <int U, bool EXACT>
__device__ __forceinline__ bool t2_energy_pass(const double *H, int nh, int lane, bool rough, double sig_rough, const double (&c_nre)[U], const double (&c_nim)[U],
					       const double (&c_nsum)[U], const double (&c_i2re)[U], const double (&c_i2im)[U], const double (&c_kr)[U],
					       double (&w)[U], int &lead, int &first_bad)
{
	bool still = true, sus = false;
	_Pragma("unroll 1") for (int h0 = 0; h0 < nh; h0 += 32) {
		double h_ct = 0, h_s2 = 0, h_fs = 0, h_al = 0; /* entry h0 + lane */
		if (h0 + lane < nh) {
			const double2 *hp = reinterpret_cast<const double2 *>(H + (h0 + lane) * 4);
			const double2 a = hp[0], c = hp[1];
			h_ct = a.x; h_s2 = a.y; h_fs = c.x; h_al = c.y;
		}
		const int kn = (nh - h0) < 32 ? (nh - h0) : 32;
		T2_PRAGMA(unroll T2_EK) for (int k = 0; k < kn; k++) {
#ifdef T2_HIST_LDG /* tuning variant: every lane reads the entry itself (one broadcast transaction per 16 bytes) instead of 6-8 shuffles */
			const double2 *hq = reinterpret_cast<const double2 *>(H + (h0 + k) * 4);
			const double2 ha = hq[0], hb = hq[1];
			const double ct = ha.x, s2 = ha.y, fs = hb.x, al = hb.y;
			(void)h_ct; (void)h_s2; (void)h_fs; (void)h_al;
#else
			const double ct = __shfl_sync(0xffffffffu, h_ct, k), s2 = __shfl_sync(0xffffffffu, h_s2, k);
			const double fs = __shfl_sync(0xffffffffu, h_fs, k);
			double al = 0;
			if (rough)
				al = __shfl_sync(0xffffffffu, h_al, k);
#endif
			bool alive_k = false, bad_k = false;
			uint32_t hr = 0, hw = 0; /* largest high words of the lane's reflectivities / weights at this reflection */
#pragma unroll
			for (int u = 0; u < U; u++) {
				const double r = fresnel_rtot_fast_ns<EXACT>(c_nre[u], c_nim[u], c_nsum[u], c_i2re[u], c_i2im[u], ct, s2, fs);
				if (rough) {
					const double cr = c_kr[u] * al * sig_rough;
					w[u] = w[u] * r * exp(-1. * cr * cr);
				} else {
					w[u] = w[u] * r; /* "* 1.0" of the reference is exact */
				}
				if (EXACT) {
					bad_k |= pcb_refl_out_of_range(r);
					alive_k |= pcb_weight_alive(w[u]);
				} else {
					const uint32_t a = pcb_hi_word(r), b = pcb_hi_word(w[u]);
					hr = a > hr ? a : hr;
					hw = b > hw ? b : hw;
				}
			}
			if (!EXACT) {
				const int sc = pcb_weight_screen_hi(hw); /* the largest weight decides "any weight >= 1e-4" unless it sits on the edge */
				alive_k = sc > 0;
				sus |= pcb_refl_suspect_hi(hr) | (sc < 0);
			}
			if (bad_k && h0 + k < first_bad)
				first_bad = h0 + k;
			still = still && alive_k;
			lead += still ? 1 : 0;
		}
	}
	return EXACT ? false : __any_sync(0xffffffffu, sus);
}

template <int MODE, int EXACT_W, int DEFER>
__global__ void __launch_bounds__(T2_THREADS, 1) pcb_trace2_kernel(const PcbTraceParams P)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ __align__(8) uint64_t tma_bar;
	__shared__ unsigned q_head[TQ_N], q_tail[TQ_N];
	__shared__ int live, cur;
#if T2_LOCKSTEP
	__shared__ unsigned snap_h[TQ_N], snap_c[TQ_N];
#endif

	const PcbOptic &o = P.o;
	const int nmax = o.nmax;
	const int nE = o.n_energies;
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;

	/* ---- carve shared memory: tables | photon state | queues | Fresnel staging ---- */
	const uint32_t tb = (uint32_t)o.tstride * 8u;
	unsigned char *p = smem_raw;
	Tables T;
	T.z = reinterpret_cast<double *>(p); /* the three tables of the segment scan are staged; cap/ext (one read per wall hit) stay global */
	T.hexz = reinterpret_cast<double *>(p + tb);
	T.rm = reinterpret_cast<double *>(p + 2 * tb);
	T.cap = o.cap;
	T.ext = o.ext;
	p += 3 * tb;
	T2Shared S;
	S.d.base = reinterpret_cast<double *>(p);
	p += D_N * T2_NP * 8;
	S.lid = reinterpret_cast<unsigned *>(p); p += T2_NP * 4;
	S.irefl = reinterpret_cast<int *>(p); p += T2_NP * 4;
	S.ix = reinterpret_cast<int *>(p); p += T2_NP * 4;
	S.iseg = reinterpret_cast<int *>(p); p += T2_NP * 4;
	S.ntrace = reinterpret_cast<int *>(p); p += T2_NP * 4;
	S.qbase = reinterpret_cast<volatile unsigned short *>(p);
	p += TQ_N * T2_QCAP * 2;
	unsigned short *batch = reinterpret_cast<unsigned short *>(p) + warp * 32; /* slots of this warp's current batch */
	p += (T2_THREADS / 32) * 32 * 2;
	S.flags = p; p += T2_NP;
	S.code = reinterpret_cast<signed char *>(p); p += T2_NP;
	S.nh = p; p += T2_NP;
	S.resume = reinterpret_cast<signed char *>(p); p += T2_NP;
	p += (16 - (4 * T2_NP) % 16) % 16;
	unsigned *wstat = reinterpret_cast<unsigned *>(p);
	p += (T2_THREADS / 32) * 16 * 4;

	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tma_bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = threadIdx.x; i < TQ_N * T2_QCAP; i += blockDim.x)
		S.qbase[i] = T2_EMPTY;
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tma_bar)), "r"(3 * tb) : "memory");
		tma_bulk_load(const_cast<double *>(T.z), o.z, tb, &tma_bar);
		tma_bulk_load(const_cast<double *>(T.hexz), o.hexz, tb, &tma_bar);
		tma_bulk_load(const_cast<double *>(T.rm), o.rm, tb, &tma_bar);
		for (int k = 0; k < TQ_N; k++) {
			q_head[k] = 0;
			q_tail[k] = 0;
		}
		q_tail[TQ_NEW] = T2_NP; /* every slot starts free */
		live = T2_NP;
		cur = TQ_HIT;
	}
	for (int i = threadIdx.x; i < T2_NP; i += blockDim.x)
		S.q(TQ_NEW)[i] = (unsigned short)i;
	if (threadIdx.x < (T2_THREADS / 32) * 16)
		wstat[threadIdx.x] = 0;
	{
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
				     : "=r"(done)
				     : "r"(smem_u32(&tma_bar))
				     : "memory");
		}
	}
	__syncthreads();

	const double *ec = o.ec;
	double *wcta = P.wscratch + (unsigned long long)blockIdx.x * T2_NP * (unsigned long long)nE; /* weight vector of slot s at wcta + s*nE */
	/* deferred weights: reflection history of slot s, T2_HK entries of 4 doubles (cos theta, sin^2 theta, frac_s, cos alfa) */
	double *hcta = DEFER ? P.hist + (unsigned long long)blockIdx.x * T2_NP * (unsigned long long)(T2_HK * 4) : nullptr;
	/* work and scheduler counters live in shared memory, one row per warp, updated by lane 0 (they would cost 12 registers) */
	enum { W_STEPS = 0, W_SOLVES, W_REFL, W_LAUNCH, W_BATCHES, W_LANES, W_EMPTY, W_LOST, W_HIT_B, W_HIT_L, W_SCAN_B, W_SCAN_L, W_N = 16 };
	volatile unsigned *wst = wstat + warp * W_N;
#if !T2_LOCKSTEP
	volatile unsigned *vhead = q_head, *vtail = q_tail;
	volatile int *vlive = &live, *vcur = &cur;
#endif

	/* begin one capil_trace call for the photon in registers (capil.c:1236-1246); see pcb_trace.cuh::trace_begin */
#define T2_TRACE_BEGIN(on_wall)                                                                                       \
	do {                                                                                                          \
		if (ntrace > nmax) {                                                                                  \
			code = 1;                                                                                     \
			next = TQ_DONE;                                                                               \
		} else {                                                                                              \
			ntrace++;                                                                                     \
			V3 d_ = {dx, dy, dz};                                                                         \
			norm3(d_);                                                                                    \
			dx = d_.x; dy = d_.y; dz = d_.z;                                                              \
			lhz = ez;                                                                                     \
			iseg = ix;                                                                                    \
			flags &= ~F_INCUR;                                                                            \
			if (iseg < nmax) {                                                                            \
				const bool steep_ = (dx * dx + dy * dy) > ((flags & F_DEEP) ? 0.0625 : 1.e-4) * (dz * dz); \
				bool in_;                                                                             \
				if ((on_wall) && !steep_) {                                                           \
					in_ = true;                                                                   \
				} else {                                                                              \
					const double t_ = (T.z[iseg] - ez) * pcb_rcp_near(dz);                        \
					const double hz_ = T.hexz[iseg];                                              \
					const double ux_ = pcb_fma(-A, hz_, pcb_fma(dx, t_, ex));                     \
					const double uy_ = pcb_fma(-B, hz_, pcb_fma(dy, t_, ey));                     \
					in_ = pcb_fma(ux_, ux_, uy_ * uy_) < T.rm[iseg];                              \
				}                                                                                     \
				if (in_)                                                                              \
					flags |= F_INCUR;                                                             \
			}                                                                                             \
			next = TQ_SCAN;                                                                               \
		}                                                                                                     \
	} while (0)

	/* segment scan of the photon in registers: up to K steps with the convexity skip test (see pcb_trace.cuh, phase SCAN).
	 * Leaves next = TQ_HIT (candidate segment found), TQ_DONE (flew out of the end) or TQ_SCAN (budget used up). */
#define T2_SCAN_STEPS(K)                                                                                               \
	do {                                                                                                          \
		const double inv_dz_ = pcb_rcp_near(dz);                                                              \
		bool in_cur_ = flags & F_INCUR;                                                                       \
		const bool slow_ = flags & F_SLOW;                                                                    \
		next = TQ_SCAN;                                                                                       \
		_Pragma("unroll 1") for (int rep_ = 0; rep_ < (K); rep_++)                                            \
		{                                                                                                     \
			if (iseg >= nmax) { /* capil.c:1312-1313 -> -2 -> launch returns 1 */                         \
				code = 1;                                                                             \
				next = TQ_DONE;                                                                       \
				break;                                                                                \
			}                                                                                             \
			const double t_ = (T.z[iseg + 1] - ez) * inv_dz_;                                             \
			const double hz_ = T.hexz[iseg + 1];                                                          \
			const double ux_ = pcb_fma(-A, hz_, pcb_fma(dx, t_, ex));                                     \
			const double uy_ = pcb_fma(-B, hz_, pcb_fma(dy, t_, ey));                                     \
			const bool in_next_ = pcb_fma(ux_, ux_, uy_ * uy_) < T.rm[iseg + 1];                          \
			if (in_cur_ && in_next_ && !slow_) {                                                          \
				iseg++;                                                                               \
				t_steps++;                                                                            \
				in_cur_ = in_next_;                                                                   \
			} else {                                                                                      \
				flags = (flags & ~F_INNXT) | (in_next_ ? F_INNXT : 0);                                \
				next = TQ_HIT;                                                                        \
				break;                                                                                \
			}                                                                                             \
		}                                                                                                     \
		flags = (flags & ~F_INCUR) | (in_cur_ ? F_INCUR : 0);                                                 \
	} while (0)

#if T2_LOCKSTEP
	/* ---------------- rounds: between two CTA barriers no warp holds a photon, so every photon of the pool is queued.  All warps
	 * read the same queue counts and each takes the batch its warp index selects from one ordering (full batches of wall hits
	 * first), so nothing is raced for and the warps of the SM walk through the same task code at the same time: the instruction
	 * working set at any moment is a window of one task, not twenty unrelated places of a 150 KB kernel. ---------------- */
	for (;;) {
		int type = -1, n = 0;
		unsigned base = 0;
		__syncthreads(); /* the pushes of the previous round are complete and visible */
		if (threadIdx.x < TQ_N) {
			snap_h[threadIdx.x] = q_head[threadIdx.x];
			snap_c[threadIdx.x] = q_tail[threadIdx.x] - q_head[threadIdx.x];
		}
		__syncthreads();
		{
			const unsigned order_full = TQ_REFL | (TQ_HIT << 4) | (TQ_SCAN << 8) | (TQ_DONE << 12) | (TQ_NEW << 16); /* nibble k = k-th choice */
			const unsigned order_part = TQ_DONE | (TQ_NEW << 4) | (TQ_SCAN << 8) | (TQ_HIT << 12) | (TQ_REFL << 16);
			if (snap_c[0] + snap_c[1] + snap_c[2] + snap_c[3] + snap_c[4] == 0)
				break; /* nothing queued, so nothing alive */
			int idx = warp;
#pragma unroll
			for (int k = 0; k < TQ_N; k++) { /* full batches in priority order */
				const int qk = (order_full >> (4 * k)) & 15;
				const int f = (int)(snap_c[qk] >> 5);
				if (type < 0) {
					if (idx < f) {
						type = qk;
						base = snap_h[qk] + 32u * idx;
						n = 32;
					} else {
						idx -= f;
					}
				}
			}
#pragma unroll
			for (int k = 0; k < TQ_N; k++) { /* then the partial ones: recycle slots first */
				const int qk = (order_part >> (4 * k)) & 15;
				const unsigned c = snap_c[qk];
				if (type < 0 && (c & 31u)) {
					if (idx == 0) {
						type = qk;
						base = snap_h[qk] + (c & ~31u);
						n = (int)(c & 31u);
					}
					idx--;
				}
			}
			if (threadIdx.x == 0) { /* advance the heads by what the T2_THREADS/32 warps of this round take */
				int left = T2_THREADS / 32;
#pragma unroll
				for (int k = 0; k < TQ_N; k++) {
					const int qk = (order_full >> (4 * k)) & 15;
					const int f = (int)(snap_c[qk] >> 5);
					const int t = f < left ? f : left;
					left -= t;
					q_head[qk] += 32u * t;
				}
#pragma unroll
				for (int k = 0; k < TQ_N; k++) {
					const int qk = (order_part >> (4 * k)) & 15;
					const unsigned c = snap_c[qk];
					if (left > 0 && (c & 31u)) {
						q_head[qk] += c & 31u;
						left--;
					}
				}
			}
			if (type < 0) {
				if (lane == 0)
					wst[W_EMPTY]++;
				continue; /* fewer batches than warps in this round */
			}
		}
		if (lane == 0) {
			wst[W_BATCHES]++;
			wst[W_LANES] += n;
			if (type == TQ_HIT) { wst[W_HIT_B]++; wst[W_HIT_L] += n; wst[W_SOLVES] += n; }
			if (type == TQ_SCAN) { wst[W_SCAN_B]++; wst[W_SCAN_L] += n; }
		}
		unsigned t_steps = 0; /* this lane's segment steps, reflections (<< 16) and launches (<< 24) in this task */
		const bool act = lane < n;
		unsigned slot = 0;
		if (act)
			slot = S.q(type)[(base + lane) & (T2_QCAP - 1)];
#else
	unsigned backoff = 256u;
#if T2_STICKY
	int last = -1;
#endif
	unsigned pend = 0; /* second half of a double claim: count << 16 | task << 12 | ring position; 0: none */
	for (;;) {
		/* ---------------- claim a batch: lane k polls queue k, the fullest-first rule is evaluated by ballot ---------------- */
		int type = -1, n = 0;
		unsigned base = 0;
		if (T2_DOUBLE_MIN && pend) { /* the rest of what this warp claimed last time: same task, no new claim */
			type = (int)((pend >> 12) & 15u);
			base = pend & (T2_QCAP - 1);
			n = (int)(pend >> 16);
			pend = 0;
		} else {
			unsigned cnt = 0;
			if (lane < TQ_N)
				cnt = vtail[lane] - vhead[lane];
			/* Task choice.  The two long tasks (HIT: the solve, REFL: the reflection) are the kernel's hot code, ~20 KB of
			 * instructions each; with every warp free to run either, twenty warps roam over more code than the SM's
			 * instruction cache holds and the GPC-level cache that backs it saturates (ncu: 70 % hit rate, gcc instruction
			 * requests at 85 % of peak).  So the warps follow a common preference `cur`: whoever finds its queue short of a
			 * full batch moves the preference to the fullest queue.  The pool then flows through one task at a time, the
			 * SM's working set is one task's code, and no barrier is needed.  The short tasks (DONE, NEW, SCAN: slot
			 * recycling) run whenever they have a full batch. */
			const unsigned full = __ballot_sync(0xffffffffu, cnt >= 32u);
			const unsigned any = __ballot_sync(0xffffffffu, cnt > 0u);
			const unsigned pick = any;
			if (pick) {
				const unsigned shortq = full & ((1u << TQ_DONE) | (1u << TQ_NEW) | (1u << TQ_SCAN));
				const int c = *vcur;
				if (shortq & (1u << TQ_DONE))
					type = TQ_DONE;
				else if (shortq & (1u << TQ_NEW))
					type = TQ_NEW;
				else if (shortq & (1u << TQ_SCAN))
					type = TQ_SCAN;
#if T2_STICKY
				else if (last >= 0 && (full & (1u << last)))
					type = last; /* the code this warp ran last is the likeliest to be cached still */
#endif
				else if (T2_CONSENSUS && (full & (1u << c)))
					type = c;
				else {
					/* deferred weights: an entry of the energy queue is a whole photon's worth of full-lane work whatever the
					 * batch size, so it is taken before any PARTIAL batch of the per-photon tasks (and after their full ones) */
					const unsigned key = __reduce_max_sync(0xffffffffu, (lane < TQ_N && !(DEFER && lane == TQ_ENERGY)) ? ((cnt << 3) | (unsigned)lane) : 0u);
					if (DEFER && (any & (1u << TQ_ENERGY)) && (key >> 3) < 32u) {
						type = TQ_ENERGY;
					} else {
						type = (int)(key & 7u);
						if (lane == 0 && type != c && (type == TQ_HIT || type == TQ_REFL))
							*vcur = type;
					}
				}
				if (lane == 0) {
					const unsigned lim = (DEFER && type == TQ_ENERGY) ? 1u : 32u;
					_Pragma("unroll 1") for (int attempt = 0; attempt < T2_CAS_TRIES; attempt++) {
						const unsigned h = vhead[type];
						const unsigned avail = vtail[type] - h;
						unsigned want = avail < lim ? avail : lim;
						if (T2_DOUBLE_MIN && lim == 32u && avail >= (unsigned)T2_DOUBLE_MIN && (type == TQ_HIT || type == TQ_REFL))
							want = 64u;
						if (want == 0)
							break;
						if (atomicCAS(&q_head[type], h, h + want) == h) {
							n = (int)want;
							base = h;
							break;
						}
					}
				}
				n = __shfl_sync(0xffffffffu, n, 0);
				base = __shfl_sync(0xffffffffu, base, 0);
				if (n == 0) {
					if (lane == 0)
						wst[W_LOST]++;
					continue; /* lost the race */
				}
				if (T2_DOUBLE_MIN && n > 32) {
					pend = ((unsigned)(n - 32) << 16) | ((unsigned)type << 12) | ((base + 32u) & (T2_QCAP - 1));
					n = 32;
				}
			} else {
				if (*vlive == 0)
					break; /* nothing queued, nothing in flight */
				if (lane == 0)
					wst[W_EMPTY]++;
				__nanosleep(backoff); /* all photons are in other warps' hands: back off so the pollers leave the issue slots alone */
				if (backoff < T2_BACKOFF_MAX)
					backoff <<= 1;
				continue;
			}
		}
		backoff = 256u;
#if T2_STICKY
		last = (type == TQ_HIT || type == TQ_REFL) ? type : last;
#endif
		if (lane == 0) {
			wst[W_BATCHES]++;
			wst[W_LANES] += n;
			if (type == TQ_HIT) { wst[W_HIT_B]++; wst[W_HIT_L] += n; wst[W_SOLVES] += n; }
			if (type == TQ_SCAN) { wst[W_SCAN_B]++; wst[W_SCAN_L] += n; }
		}
		unsigned t_steps = 0; /* this lane's segment steps, reflections (<< 16) and launches (<< 24) in this task */
		const bool act = lane < n;
		unsigned slot = 0;
		if (act) {
			volatile unsigned short *e = &S.q(type)[(base + lane) & (T2_QCAP - 1)];
			unsigned short v;
			while ((v = *e) == T2_EMPTY) {
			}
			*e = T2_EMPTY;
			slot = v;
		}
		__threadfence_block();

#endif

		/* Photon fields are read from the slot when a phase needs them and written back as soon as they are final, so that
		 * few of them are live at any point (96 registers per thread, 640 threads). */
#define SD(k) S.d[k][slot]
		int next = -1; /* queue the photon goes to; -1: slot retired */

		if (type == TQ_SCAN) {
			/* ---------------- SCAN: continuation of a long free flight ---------------- */
			if (act) {
				const double ex = SD(D_EX), ey = SD(D_EY), ez = SD(D_EZ), dx = SD(D_DX), dy = SD(D_DY), dz = SD(D_DZ), A = SD(D_A), B = SD(D_B);
				int iseg = S.iseg[slot], code = 0;
				unsigned flags = S.flags[slot];
				T2_SCAN_STEPS(64);
				S.iseg[slot] = iseg;
				S.flags[slot] = (unsigned char)flags;
				if (next == TQ_DONE)
					S.code[slot] = (signed char)code;
			}
		} else if (type == TQ_HIT) {
			/* ---------------- HIT: solve the candidate segment (capil.c:1256-1309); on a hit the travel and in-optic checks
			 * (capil.c:1315-1343) and on to the reflection task; without one, scan on to the next candidate ---------------- */
			if (act) {
				bool hitting = false;
				double snx = 0, sny = 0, snz = 0, cosalfa = 0;
				unsigned flags = S.flags[slot];
				int iseg = S.iseg[slot], ix = 0, code = 0;
				double ex = SD(D_EX), ey = SD(D_EY), ez = SD(D_EZ);
				const double dx = SD(D_DX), dy = SD(D_DY), dz = SD(D_DZ);
				const double A = SD(D_A), B = SD(D_B);
				const bool slow = flags & F_SLOW;
				const int i = iseg;
				t_steps++;
				const double z0 = T.z[i], z1 = T.z[i + 1];
				V3 c0 = {A * T.hexz[i], B * T.hexz[i], z0};
				V3 c1 = {A * T.hexz[i + 1], B * T.hexz[i + 1], z1};
				const PcbDivisor dz_d = pcb_divisor(dz);
				V3 p0 = {ex + pcb_div(dx * (z0 - ez), dz_d), ey + pcb_div(dy * (z0 - ez), dz_d), z0};
				const double cap0 = __ldg(T.cap + i), cap1 = __ldg(T.cap + i + 1);
				if (slow && (within_pc(__ldg(T.ext + i), c0.x, c0.y) == 0 || within_pc(__ldg(T.ext + i + 1), c1.x, c1.y) == 0)) { /* capil.c:1263-1266 */
					code = -1;
					next = TQ_DONE;
				} else {
					V3 hit = {0., 0., 0.}, sn;
					V3 pd = {dx, dy, dz};
					int iesc;
					if (z0 < 0. || z1 < 0 || cap0 < 0. || cap1 < 0. || dz < 0 || z1 <= z0) {
						iesc = -1;
						sn.x = sn.y = sn.z = 0.0;
					} else {
						iesc = segment_solve(c0, c1, cap0, cap1, p0, pd, SD(D_LHZ), hit, sn);
					}
					const double ca = dot3(sn, pd);
					if (iesc == 1) {
						if (pcb_acos_is_obtuse(ca)) /* acos(cosalfa) > M_PI/2, capil.c:1270-1273 */
							iesc = -5;
						SD(D_LHZ) = hit.z; /* photon_coord keeps the last root, accepted or not; only its z is read again */
					}
					if (iesc == 1) {
						bool outside = false;
						if (slow) {
							const double e0 = __ldg(T.ext + i), e1 = __ldg(T.ext + i + 1);
							const double ext_here = ((e0 - e1) / (z0 - z1)) * (hit.z - z1) + e1;
							outside = o.mono ? (sqrt(hit.x * hit.x + hit.y * hit.y) >= ext_here) : (within_pc(ext_here, hit.x, hit.y) == 0);
						}
						if (outside) {
							code = -1;
							next = TQ_DONE;
						} else {
							ix = i + 1;
							snx = sn.x; sny = sn.y; snz = sn.z;
							cosalfa = ca;
							hitting = true;
							if (!DEFER)
								t_steps += 1u << 16; /* deferred weights: reflection events are counted when the history is evaluated */
							const double rx = hit.x - ex, ry = hit.y - ey, rz = hit.z - ez;
							SD(D_DTRAV) = SD(D_DTRAV) + sqrt(rx * rx + ry * ry + rz * rz); /* capil.c:1315-1319 */
							ex = hit.x; ey = hit.y; ez = hit.z;
							SD(D_EX) = ex; SD(D_EY) = ey; SD(D_EZ) = ez;
							if (fabs(cosalfa) > 1.0) {
								code = -1;
								next = TQ_DONE;
								hitting = false;
								flags |= F_FAILHIT;
							}
						}
					} else {
						bool outside = false;
						if (slow || !(flags & F_INCUR))
							outside = (within_pc(__ldg(T.ext + i), p0.x, p0.y) == 0);
						if (outside) {
							code = -1;
							next = TQ_DONE;
						} else { /* no wall hit in this segment (1 solve in 20): the scan task finds the next candidate */
							iseg = i + 1;
							flags = (flags & ~F_INCUR) | ((flags & F_INNXT) ? F_INCUR : 0);
							next = TQ_SCAN;
						}
					}
				}
				if (hitting) { /* *ix = last i < nmax with z[i] <= hit z (capil.c:1330-1333) */
					int found = -1;
					if (ix < nmax && T.z[ix] <= ez) {
						found = ix;
					} else {
						_Pragma("unroll 1") for (int j = ix - 1; j >= 0; j--)
							if (T.z[j] <= ez) {
								found = j;
								break;
							}
					}
					if (found >= 0)
						ix = found;
					if (slow) { /* capil.c:1334-1343 */
						const double e0 = __ldg(T.ext + ix), e1 = __ldg(T.ext + ix + 1);
						const double ext_here = ((e1 - e0) / (T.z[ix + 1] - T.z[ix])) * (ez - T.z[ix]) + e0;
						const bool outside = o.mono ? (ex * ex + ey * ey >= ext_here) : (within_pc(ext_here, ex, ey) == 0);
						if (outside) {
							code = -1;
							next = TQ_DONE;
							hitting = false;
							flags |= F_FAILHIT;
						}
					}
				}
				if (hitting) { /* the wall normal goes to the reflection task through the slot */
					SD(D_G0) = snx; SD(D_G1) = sny; SD(D_AUX) = snz;
					S.ix[slot] = ix;
					next = TQ_REFL;
				}
				S.iseg[slot] = iseg;
				S.flags[slot] = (unsigned char)flags;
				S.code[slot] = (signed char)code;
			}
		} else if (type == TQ_REFL) {
			/* ---------------- REFL: capil_reflect (capil.c:596-645) for a photon on the wall: preamble + geometry, the energy
			 * loop, the specular direction and the start of the next capil_trace call; then the scan to the next candidate ---------------- */
			bool hitting = act;
			double dx = 0, dy = 0, dz = 1, snx = 0, sny = 0, snz = 0, cosalfa = 0; /* live through the energy loop */
			double g_ct = 0, g_st = 0, g_fs = 0, alfa = 0;
			unsigned flags = 0;
			int iseg = 0, ix = 0, code = 0;
			if (act) {
				flags = S.flags[slot];
				iseg = S.iseg[slot];
				ix = S.ix[slot];
				dx = SD(D_DX); dy = SD(D_DY); dz = SD(D_DZ);
				snx = SD(D_G0); sny = SD(D_G1); snz = SD(D_AUX);
				{
					const V3 sn0 = {snx, sny, snz}, pd = {dx, dy, dz};
					cosalfa = dot3(sn0, pd); /* the same product the solve task tested */
				}
				{ /* capil_reflect preamble, capil.c:596-602, and the energy-independent geometry */
					V3 sn = {snx, sny, snz};
					norm3(sn);
					V3 d = {dx, dy, dz};
					norm3(d);
					dx = d.x; dy = d.y; dz = d.z;
					alfa = dot3(d, sn);
					if (alfa < 0.) {
						code = -1;
						next = TQ_DONE;
						hitting = false;
						flags |= F_FAILHIT;
					} else {
						V3 ev = {SD(D_VX), SD(D_VY), SD(D_VZ)};
						ReflGeom g;
						if (EXACT_W)
							refl_geometry(sn, d, ev, g);
						else
							refl_geometry_fast(sn, d, ev, g);
						SD(D_VX) = g.new_ev.x; SD(D_VY) = g.new_ev.y; SD(D_VZ) = g.new_ev.z;
						g_ct = g.cos_t; g_st = g.sin_t; g_fs = g.frac_s;
					}
				}
			}
			/* ---- energy loop (capil.c:625-645) ---- */
			const bool rough = (o.sig_rough != 0.0);
			const double sig_rough = o.sig_rough;
			bool w_alive = false, w_bad = false;
			bool flush = false;
			if (DEFER) {
				/* deferred weights: the reflection's energy-independent terms join the slot's history; the per-energy products are
				 * evaluated by the energy task for the whole history at once, weights in registers.  The photon is traced on as
				 * alive; a photon that the evaluation finds absorbed earlier is rolled back there (only its classification,
				 * reflection count and weights are ever read). */
				if (hitting) {
					int nh = S.nh[slot];
					double2 *hp = reinterpret_cast<double2 *>(hcta + ((unsigned long long)slot * T2_HK + nh) * 4);
					hp[0] = make_double2(g_ct, g_st * g_st);
					hp[1] = make_double2(g_fs, alfa);
					nh++;
					S.nh[slot] = (unsigned char)nh;
					w_alive = true;
					flush = nh >= P.hk;
				}
			} else {
				/* the batch's (photon, energy) pairs are dealt over all lanes, whatever the photons' energy count.  The geometry of each
				 * reflecting photon is staged in its slot (fields that are rewritten before they are read again), readable by every
				 * lane.  (A loop in which every lane walks its own photon's energies measured 5 % slower even for 10 energies.) */
				const unsigned go = __ballot_sync(0xffffffffu, hitting);
				const int n_go = __popc(go);
				if (hitting) {
					SD(D_G0) = g_ct; SD(D_G1) = EXACT_W ? g_st : g_st * g_st; SD(D_LHZ) = g_fs; SD(D_AUX) = alfa; /* cheap weights: sin^2 */
					batch[__popc(go & ((1u << lane) - 1u))] = (unsigned short)slot;
					if (!(flags & F_WINIT)) { /* first reflection: the weight vector starts at 1 */
						double *w0 = wcta + (unsigned long long)slot * nE;
						_Pragma("unroll 2") for (int e2 = 0; e2 < nE; e2++)
							w0[e2] = 1.0;
					}
				}
				__syncwarp();
				const int total = n_go * nE;
				unsigned alive_bits = 0, bad_bits = 0; /* bit k: k-th reflecting photon of the batch */
				int pi = lane / nE, e = lane - pi * nE;
				const int dpi = 32 / nE, de = 32 - dpi * nE;
				int j0 = 0;
				if (!EXACT_W) {
					enum { U = T2_SPREAD_ILP };
					int pu[U], eu[U];
					pu[0] = pi; eu[0] = e;
#pragma unroll
					for (int u = 1; u < U; u++) {
						pu[u] = pu[u - 1] + dpi; eu[u] = eu[u - 1] + de;
						if (eu[u] >= nE) { eu[u] -= nE; pu[u]++; }
					}
					const int dpiU = (32 * U) / nE, deU = 32 * U - dpiU * nE;
					for (; j0 + 32 * U <= total; j0 += 32 * U) {
						unsigned sl[U];
						double *wp[U];
						double w[U], r[U];
#pragma unroll
						for (int u = 0; u < U; u++) {
							sl[u] = batch[pu[u]];
							wp[u] = wcta + (unsigned long long)sl[u] * nE + eu[u];
							w[u] = *wp[u];
#if T2_PREFETCH
							/* the lane's pair of the round after next is 128 energies further in the same weight vector (when the
							 * vector is that long): pull it towards the L2 now -- at nE = 291 the pool's vectors (397 MB) live in DRAM */
							if (eu[u] + T2_PREFETCH < nE)
								asm volatile("prefetch.global.L2 [%0];" ::"l"(wp[u] + T2_PREFETCH));
#endif
						}
#pragma unroll
						for (int u = 0; u < U; u++)
							r[u] = fresnel_rtot_fast_s2(__ldg(ec + EC_NRE * nE + eu[u]), __ldg(ec + EC_NIM * nE + eu[u]), __ldg(ec + EC_I2RE * nE + eu[u]),
										    __ldg(ec + EC_I2IM * nE + eu[u]), S.d[D_G0][sl[u]], S.d[D_G1][sl[u]], S.d[D_LHZ][sl[u]]);
#pragma unroll
						for (int u = 0; u < U; u++) {
							if (rough) {
								const double cr = __ldg(ec + EC_KR * nE + eu[u]) * S.d[D_AUX][sl[u]] * sig_rough;
								w[u] = w[u] * r[u] * exp(-1. * cr * cr);
							} else {
								w[u] = w[u] * r[u] * 1.0;
							}
							*wp[u] = w[u];
							if (r[u] < 0. || r[u] > 1.)
								bad_bits |= 1u << pu[u];
							if (w[u] >= 1.e-4)
								alive_bits |= 1u << pu[u];
							pu[u] += dpiU; eu[u] += deU;
							if (eu[u] >= nE) { eu[u] -= nE; pu[u]++; }
						}
					}
					pi = pu[0];
					e = eu[0];
				}
				_Pragma("unroll 1") for (; j0 < total; j0 += 32) {
					if (j0 + lane < total) {
						const unsigned sl = batch[pi];
						double *wp = wcta + (unsigned long long)sl * nE + e;
						double w = *wp;
						const double n_re = __ldg(ec + EC_NRE * nE + e), n_im = __ldg(ec + EC_NIM * nE + e);
						const double i2_re = __ldg(ec + EC_I2RE * nE + e), i2_im = __ldg(ec + EC_I2IM * nE + e);
						const double c_t = S.d[D_G0][sl], s_t = S.d[D_G1][sl], f_s = S.d[D_LHZ][sl];
						const double rtot = EXACT_W ? fresnel_rtot(n_re, n_im, i2_re, i2_im, c_t, s_t, f_s, 1. - f_s)
									    : fresnel_rtot_fast_s2(n_re, n_im, i2_re, i2_im, c_t, s_t, f_s); /* s_t holds sin^2 here */
						if (rough) {
							const double cons1 = __ldg(ec + EC_KR * nE + e) * S.d[D_AUX][sl] * sig_rough;
							w = w * rtot * exp(-1. * cons1 * cons1);
						} else {
							w = w * rtot * 1.0;
						}
						*wp = w;
						if (rtot < 0. || rtot > 1.)
							bad_bits |= 1u << pi;
						if (w >= 1.e-4)
							alive_bits |= 1u << pi;
					}
					pi += dpi;
					e += de;
					if (e >= nE) {
						e -= nE;
						pi++;
					}
				}
				alive_bits = __reduce_or_sync(0xffffffffu, alive_bits);
				bad_bits = __reduce_or_sync(0xffffffffu, bad_bits);
				const int rank = __popc(go & ((1u << lane) - 1u));
				w_alive = (alive_bits >> rank) & 1u;
				w_bad = (bad_bits >> rank) & 1u;
			}
			if (hitting) {
				if (!DEFER)
					flags |= F_WINIT;
				if (w_bad) {
					code = -1;
					next = TQ_DONE;
				} else if (w_alive) { /* capil.c:1346-1352 */
					dx = dx - 2.0 * cosalfa * snx;
					dy = dy - 2.0 * cosalfa * sny;
					dz = dz - 2.0 * cosalfa * snz;
					V3 d = {dx, dy, dz};
					norm3(d);
					dx = d.x; dy = d.y; dz = d.z;
					S.irefl[slot] = S.irefl[slot] + 1;
					const double ex = SD(D_EX), ey = SD(D_EY), ez = SD(D_EZ), A = SD(D_A), B = SD(D_B);
					int ntrace = S.ntrace[slot];
					double lhz;
					T2_TRACE_BEGIN(true);
					if (next == TQ_SCAN) {
						S.ntrace[slot] = ntrace;
						SD(D_LHZ) = lhz;
						T2_SCAN_STEPS(T2_SCAN_AFTER);
					}
				} else {
					code = 0;
					next = TQ_DONE;
				}
				if (DEFER && flush && next != TQ_DONE) { /* history full: evaluate it, then carry on where the photon was headed */
					S.resume[slot] = (signed char)next;
					next = TQ_ENERGY;
				}
			}
			if (act) {
				SD(D_DX) = dx; SD(D_DY) = dy; SD(D_DZ) = dz;
				S.iseg[slot] = iseg; S.ix[slot] = ix;
				S.flags[slot] = (unsigned char)flags;
				S.code[slot] = (signed char)code;
			}
			__syncwarp();
		} else if (type == TQ_DONE) {
			/* ---------------- DONE: classify (source.c:758-777), dense per-id outputs ---------------- */
			if (act) {
				const double ex = SD(D_EX), ey = SD(D_EY), ez = SD(D_EZ), dx = SD(D_DX), dy = SD(D_DY), dz = SD(D_DZ);
				const double vx = SD(D_VX), vy = SD(D_VY), vz = SD(D_VZ), dtrav = SD(D_DTRAV);
				const unsigned lid = S.lid[slot], flags = S.flags[slot];
				const int irefl = S.irefl[slot], ntrace = S.ntrace[slot], code = S.code[slot];
				const unsigned long long nn = P.n_ids;
				const double *wsl = wcta + (unsigned long long)slot * nE;
				const bool w_init = flags & F_WINIT;
				int status = code;
				const double zt = T.z[nmax];
				const double ext_end = __ldg(T.ext + nmax);
				if (MODE == MODE_SOURCE && code == 1) {
					const double px = ex + dx * (zt - ez) / dz;
					const double py = ey + dy * (zt - ez) / dz;
					const int inside = o.mono ? ((sqrt((px) * (px) + (py) * (py)) > ext_end) ? 0 : 1) : within_pc(ext_end, px, py);
					if (inside != 1)
						status = ST_MISSED_WINDOW;
				}
				P.out.status[lid] = (int8_t)status;
				P.out.irefl[lid] = irefl;
				double *F = P.out.fields;
				const bool wr_rec = (MODE == MODE_REPLAY) || status == ST_EXIT || P.record_all;
				if (MODE == MODE_REPLAY) {
					F[0 * nn + lid] = ex; F[1 * nn + lid] = ey; F[2 * nn + lid] = ez;
					F[3 * nn + lid] = dx; F[4 * nn + lid] = dy; F[5 * nn + lid] = dz;
					F[6 * nn + lid] = vx; F[7 * nn + lid] = vy; F[8 * nn + lid] = vz;
					F[9 * nn + lid] = dtrav;
					F[10 * nn + lid] = (double)ntrace;
				} else if (wr_rec) {
					StartState s0;
					source_sample(P.s, o.mono, P.id0 + lid, s0);
					V3 sd = s0.dir;
					for (int k = 0; k < ntrace + 1; k++) {
						V3 prev = sd;
						norm3(sd);
						if (sd.x == prev.x && sd.y == prev.y && sd.z == prev.z)
							break;
					}
					F[0 * nn + lid] = s0.src_x; F[1 * nn + lid] = s0.src_y;
					F[2 * nn + lid] = s0.coords.x; F[3 * nn + lid] = s0.coords.y;
					F[4 * nn + lid] = sd.x; F[5 * nn + lid] = sd.y;
					const double ca = dot3(s0.ev, sd);
					const double al = acos(ca);
					const double c_ae = 1. / sin(al), c_be = -1. * c_ae * ca;
					V3 t = {s0.ev.x * c_ae + sd.x * c_be, s0.ev.y * c_ae + sd.y * c_be, s0.ev.z * c_ae + sd.z * c_be};
					norm3(t);
					F[6 * nn + lid] = round(t.x); F[7 * nn + lid] = round(t.y);
					const double px = ex + dx * (zt - ez) / dz;
					const double py = ey + dy * (zt - ez) / dz;
					const double pz = ez + dz * (zt - ez) / dz;
					F[8 * nn + lid] = px; F[9 * nn + lid] = py; F[10 * nn + lid] = pz;
					F[11 * nn + lid] = dx; F[12 * nn + lid] = dy;
					V3 u = {vx * c_ae + dx * c_be, vy * c_ae + dy * c_be, vz * c_ae + dz * c_be};
					norm3(u);
					F[13 * nn + lid] = round(u.x); F[14 * nn + lid] = round(u.y);
					F[15 * nn + lid] = (double)irefl;
					F[16 * nn + lid] = dtrav + sqrt((px - ex) * (px - ex) + (py - ey) * (py - ey) + (zt - ez) * (zt - ez));
				}
				if (DEFER) {
					if (flags & F_FAILHIT)
						t_steps += 1u << 16; /* a rejected last wall hit that the evaluation of the history did not roll back */
					if ((wr_rec || status == ST_ABSORBED) && !(flags & F_WOUT)) { /* no reflection since the last evaluation (if any): the scratch holds the weights */
						double *W = P.out.weights + (unsigned long long)lid * nE;
						_Pragma("unroll 2") for (int e2 = 0; e2 < nE; e2++)
							W[e2] = w_init ? wsl[e2] : 1.0;
					}
				} else if (wr_rec || status == ST_ABSORBED) {
					double *W = P.out.weights + (unsigned long long)lid * nE;
					_Pragma("unroll 2") for (int e2 = 0; e2 < nE; e2++)
						W[e2] = w_init ? wsl[e2] : 1.0;
				}
				next = TQ_NEW;
			}
		} else if (DEFER && type == TQ_ENERGY) {
			/* ---------------- ENERGY (deferred weights): capil_reflect's energy loop (capil.c:625-650) for every buffered reflection of a
			 * photon at once.  The warp works on ONE photon: lane l owns energies l, l+32, ... (T2_EU per pass) with their optical
			 * constants and weights in registers, the history entries sit one per lane and are broadcast by shuffle, so the inner loop
			 * is the Fresnel term and nothing else: no loads, no stores.  Weights are multiplied in the reference's order (reflection
			 * by reflection), so they equal the immediate update bit for bit.  The survival test (any weight >= 1e-4 after each
			 * reflection, capil.c:643-650) and the range test of the reflectivity are replayed per reflection: weights only shrink,
			 * so "alive" holds on a prefix of the history and the first reflection that absorbs the photon is the longest such
			 * prefix over all energies. ---------------- */
			const bool rough = (o.sig_rough != 0.0);
			const double sig_rough = o.sig_rough;
			enum { U = T2_EU };
			for (int b = 0; b < n; b++) {
				const unsigned sl = __shfl_sync(0xffffffffu, slot, b);
				const int nh = S.nh[sl];
				const unsigned fl = S.flags[sl];
				const int resume = S.resume[sl];
				const bool final = (resume == TQ_DONE);
				const bool winit = fl & F_WINIT;
				const double *H = hcta + (unsigned long long)sl * (T2_HK * 4);
				double *wscr = wcta + (unsigned long long)sl * nE;
				double *wdst = final ? P.out.weights + (unsigned long long)S.lid[sl] * nE : wscr;
				int k_dead = 0;		 /* reflections after which the photon is still alive (longest prefix over the passes) */
				int first_bad = 1 << 30; /* first reflection with a reflectivity outside [0, 1] (this lane's energies) */
				_Pragma("unroll 1") for (int e0 = 0; e0 < nE; e0 += 32 * U) {
					int eu[U];
					bool val[U];
					double w[U], c_nre[U], c_nim[U], c_nsum[U], c_i2re[U], c_i2im[U], c_kr[U];
#pragma unroll
					for (int u = 0; u < U; u++) {
						eu[u] = e0 + 32 * u + lane;
						val[u] = eu[u] < nE;
						const int ei = val[u] ? eu[u] : nE - 1;
						c_nre[u] = __ldg(ec + EC_NRE * nE + ei);
						c_nim[u] = __ldg(ec + EC_NIM * nE + ei);
						c_nsum[u] = c_nre[u] + c_nim[u];
						c_i2re[u] = __ldg(ec + EC_I2RE * nE + ei);
						c_i2im[u] = __ldg(ec + EC_I2IM * nE + ei);
						c_kr[u] = rough ? __ldg(ec + EC_KR * nE + ei) : 0.0;
						/* lanes past the end of the grid work on a copy of the last energy with weight 0: never alive, and out of range only
						 * where the last energy itself is -- no per-lane validity test inside the loop */
						w[u] = val[u] ? (winit ? wscr[eu[u]] : 1.0) : 0.0;
					}
					/* fast pass: the survival and range tests are screened on the high words (integer pipe), u == 0 is not looked for */
					int lead = 0, fb = 1 << 30;
					const bool redo = t2_energy_pass<U, false>(H, nh, lane, rough, sig_rough, c_nre, c_nim, c_nsum, c_i2re, c_i2im, c_kr, w, lead, fb);
					if (redo) { /* a value at the edge of a test, an infinity or a NaN: the same pass with the plain comparisons */
#pragma unroll
						for (int u = 0; u < U; u++)
							w[u] = val[u] ? (winit ? wscr[eu[u]] : 1.0) : 0.0;
						lead = 0;
						fb = 1 << 30;
						t2_energy_pass<U, true>(H, nh, lane, rough, sig_rough, c_nre, c_nim, c_nsum, c_i2re, c_i2im, c_kr, w, lead, fb);
					}
					first_bad = fb < first_bad ? fb : first_bad;
					const int lead_w = __reduce_max_sync(0xffffffffu, lead);
					k_dead = lead_w > k_dead ? lead_w : k_dead;
#pragma unroll
					for (int u = 0; u < U; u++)
						if (val[u])
							wdst[eu[u]] = w[u];
				}
				first_bad = __reduce_min_sync(0xffffffffu, first_bad);
				/* reflection k_end ended the photon: reflectivity out of range (-1, tested first: capil.c:636-641) or absorbed (0) */
				int k_end = nh, code_end = 1;
				if (first_bad < nh && first_bad <= k_dead) {
					k_end = first_bad;
					code_end = -1;
				} else if (k_dead < nh) {
					k_end = k_dead;
					code_end = 0;
				}
				int nq = resume;
				if (lane == 0) {
					unsigned f2 = fl | (final ? F_WOUT : F_WINIT);
					if (k_end < nh) { /* roll the optimistic bookkeeping back to reflection k_end (it counted nh - k_end reflections too many) */
						S.code[sl] = (signed char)code_end;
						S.irefl[sl] = S.irefl[sl] - (nh - k_end);
						f2 &= ~(unsigned)F_FAILHIT;
						wst[W_REFL] += (unsigned)(k_end + 1);
					} else {
						wst[W_REFL] += (unsigned)nh;
					}
					S.flags[sl] = (unsigned char)f2;
					S.nh[sl] = 0;
				}
				if (k_end < nh)
					nq = TQ_DONE;
				if (lane == b)
					next = nq;
			}
		} else {
			/* ---------------- NEW: fetch an id, build the start state, launch (photon.c:476-645) ---------------- */
			const unsigned m = __ballot_sync(0xffffffffu, act);
			unsigned long long idbase = 0;
			if (m) {
				const int leader = __ffs(m) - 1;
				if (lane == leader)
					idbase = atomicAdd(&P.wc->next_id, (unsigned long long)__popc(m));
				idbase = __shfl_sync(0xffffffffu, idbase, leader);
			}
			if (act) {
				const unsigned long long id = idbase + __popc(m & ((1u << lane) - 1u));
				if (id >= P.n_ids) {
					atomicSub(&live, 1); /* no ids left: the slot retires */
					next = -1;
				} else {
					double ex, ey, ez, dx, dy, dz, lhz = 0, A = 0, B = 0;
					int ix = 0, iseg = 0, ntrace = 0, code = 0;
					unsigned flags = 0;
					t_steps += 1u << 24;
					StartState s0;
					if (MODE == MODE_SOURCE) {
						source_sample(P.s, o.mono, P.id0 + id, s0);
					} else {
						s0.coords.x = P.rin.coords[3 * id]; s0.coords.y = P.rin.coords[3 * id + 1]; s0.coords.z = P.rin.coords[3 * id + 2];
						s0.dir.x = P.rin.dir[3 * id]; s0.dir.y = P.rin.dir[3 * id + 1]; s0.dir.z = P.rin.dir[3 * id + 2];
						s0.ev.x = P.rin.ev[3 * id]; s0.ev.y = P.rin.ev[3 * id + 1]; s0.ev.z = P.rin.ev[3 * id + 2];
					}
					norm3(s0.dir);
					ex = s0.coords.x; ey = s0.coords.y; ez = s0.coords.z;
					V3 d = s0.dir;
					norm3(d);
					dx = d.x; dy = d.y; dz = d.z;
					SD(D_VX) = s0.ev.x; SD(D_VY) = s0.ev.y; SD(D_VZ) = s0.ev.z;
					int z_id = 0;
					if (ez > 0) {
						_Pragma("unroll 1") for (int i = 0; i < nmax; i++)
							if (T.z[i] <= ez)
								z_id = i;
					}
					const double ext0 = __ldg(T.ext + z_id), ext1 = __ldg(T.ext + z_id + 1);
					const double ext_here = ((ext0 - ext1) / (T.z[z_id] - T.z[z_id + 1])) * (ez - T.z[z_id]) + ext0;
					double q_i = 0, r_i = 0;
					bool outside;
					if (o.mono) {
						outside = (sqrt((ex) * (ex) + (ey) * (ey)) > ext_here);
					} else {
						const double hz = ext_here / (2. * PCB_COSPI_6 * (o.n_shells + 1));
						hex_cell(hz, ex, ey, q_i, r_i);
						outside = (within_pc(ext_here, ex, ey) == 0);
					}
					if (outside) {
						code = -2;
						next = TQ_DONE;
					} else {
						A = (2. * q_i + r_i) * PCB_COSPI_6;
						B = r_i * (3. / 2);
						ix = 0;
						if (ez > 0) {
							_Pragma("unroll 1") for (int i = 0; i <= nmax; i++)
								if (T.z[i] <= ez)
									ix = i;
						} else if (T.z[0] <= ez) {
							ix = 0;
							_Pragma("unroll 1") for (int i = 1; i <= nmax && T.z[i] <= ez; i++)
								ix = i;
						}
						double rad_here, cx_here, cy_here;
						if (ez > 0) {
							const double dzs = (T.z[z_id + 1] - T.z[z_id]), off = (ez - T.z[z_id]);
							const double cx0 = A * T.hexz[z_id], cx1 = A * T.hexz[z_id + 1];
							const double cy0 = B * T.hexz[z_id], cy1 = B * T.hexz[z_id + 1];
							const double cp0 = __ldg(T.cap + z_id), cp1 = __ldg(T.cap + z_id + 1);
							rad_here = ((cp1 - cp0) / dzs) * off + cp0;
							cx_here = ((cx1 - cx0) / dzs) * off + cx0;
							cy_here = ((cy1 - cy0) / dzs) * off + cy0;
						} else {
							rad_here = __ldg(T.cap);
							cx_here = A * T.hexz[0];
							cy_here = B * T.hexz[0];
						}
						const double d_axis = sqrt((ex - cx_here) * (ex - cx_here) + (ey - cy_here) * (ey - cy_here));
						if (d_axis > rad_here) {
							code = 2;
							next = TQ_DONE;
						} else {
							const double m1 = fabs(B), m2 = fabs(PCB_COSPI_6 * A + 0.5 * B), m3 = fabs(PCB_COSPI_6 * A - 0.5 * B);
							const double mm = fmax(m1, fmax(m2, m3));
							if (o.all_slow || !(mm < o.m_safe))
								flags |= F_SLOW;
							if (mm < o.m_deep)
								flags |= F_DEEP;
							T2_TRACE_BEGIN(false);
							if (next == TQ_SCAN)
								T2_SCAN_STEPS(T2_SCAN_AFTER);
						}
					}
					SD(D_EX) = ex; SD(D_EY) = ey; SD(D_EZ) = ez;
					SD(D_DX) = dx; SD(D_DY) = dy; SD(D_DZ) = dz;
					SD(D_LHZ) = lhz;
					SD(D_DTRAV) = 0.0; SD(D_A) = A; SD(D_B) = B;
					S.lid[slot] = (unsigned)id; S.irefl[slot] = 0; S.ix[slot] = ix; S.iseg[slot] = iseg; S.ntrace[slot] = ntrace;
					S.flags[slot] = (unsigned char)flags;
					S.code[slot] = (signed char)code;
					if (DEFER)
						S.nh[slot] = 0;
				}
			}
		}
#undef SD
		if (DEFER && act && next == TQ_DONE && type != TQ_ENERGY && S.nh[slot] != 0) { /* buffered reflections are evaluated before the photon is classified */
			S.resume[slot] = (signed char)TQ_DONE;
			next = TQ_ENERGY;
		}

		{
			const unsigned tsum = __reduce_add_sync(0xffffffffu, t_steps);
			if (lane == 0 && tsum) {
				wst[W_STEPS] += tsum & 0xffffu;
				wst[W_REFL] += (tsum >> 16) & 0xffu;
				wst[W_LAUNCH] += tsum >> 24;
			}
		}

		/* ---------------- hand the photons on: one tail reservation per warp and target queue ---------------- */
#if T2_LOCKSTEP
#pragma unroll
		for (int qk = 0; qk < TQ_N; qk++) {
			const unsigned mq = __ballot_sync(0xffffffffu, act && next == qk);
			if (mq) {
				unsigned tb0 = 0;
				const int leader = __ffs(mq) - 1;
				if (lane == leader)
					tb0 = atomicAdd(&q_tail[qk], (unsigned)__popc(mq));
				tb0 = __shfl_sync(0xffffffffu, tb0, leader);
				if (act && next == qk)
					S.q(qk)[(tb0 + __popc(mq & ((1u << lane) - 1u))) & (T2_QCAP - 1)] = (unsigned short)slot;
			}
		}
	}
#else
		__threadfence_block(); /* the slots' state is visible before their ids are */
		{
			unsigned todo = __ballot_sync(0xffffffffu, act && next >= 0);
			while (todo) { /* one pass per target queue present in the batch */
				const int qk = __shfl_sync(0xffffffffu, next, __ffs(todo) - 1);
				const unsigned mq = __ballot_sync(0xffffffffu, act && next == qk);
				unsigned tb0 = 0;
				const int leader = __ffs(mq) - 1;
				if (lane == leader)
					tb0 = atomicAdd(&q_tail[qk], (unsigned)__popc(mq));
				tb0 = __shfl_sync(0xffffffffu, tb0, leader);
				if (act && next == qk) {
					volatile unsigned short *e = &S.q(qk)[(tb0 + __popc(mq & ((1u << lane) - 1u))) & (T2_QCAP - 1)];
					while (*e != T2_EMPTY) {
					}
					*e = (unsigned short)slot;
				}
				todo &= ~mq;
			}
		}
	}
#endif
#undef T2_SCAN_STEPS
#undef T2_TRACE_BEGIN

	if (lane == 0) {
		atomicAdd(&P.wc->seg_steps, (unsigned long long)wst[W_STEPS]);
		atomicAdd(&P.wc->solves, (unsigned long long)wst[W_SOLVES]);
		atomicAdd(&P.wc->reflections, (unsigned long long)wst[W_REFL]);
		atomicAdd(&P.wc->launched, (unsigned long long)wst[W_LAUNCH]);
		for (int k = 0; k < 5; k++) /* dbg[5] is the leak tracer's wall-step counter */
			atomicAdd(&P.wc->dbg[k], (unsigned long long)wst[W_BATCHES + k]);
	}
}

/* dynamic shared memory of pcb_trace2_kernel for a table stride of `tstride` doubles */
static inline size_t t2_smem_bytes(int tstride)
{
	return 3 * (size_t)tstride * 8 + 16 * (size_t)T2_NP * 8 + (size_t)(T2_THREADS / 32) * 32 * 2 + 5 * (size_t)T2_NP * 4 + (size_t)TQ_N * T2_QCAP * 2 +
	       4 * (size_t)T2_NP + 16 + (size_t)(T2_THREADS / 32) * 16 * 4 + 64;
}

#endif
